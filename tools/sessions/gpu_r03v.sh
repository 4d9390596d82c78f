#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r03v_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/r03v_pytest.log | tail -6
timeout 600 python bench.py --config 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/r03v_bench_c1.json 2> gpurun_out/r03v_bench_c1.err; python -c "
import json; d=json.load(open('gpurun_out/r03v_bench_c1.json')); print('c1', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['stages_ms_per_step'], d['gpu_launches'])"
