#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tcs_engine.py tests/test_gpu_tcgen05_engine.py -m gpu -q -x > gpurun_out/r03e_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r03e_pytest.log | tail -6
for c in 3 4; do timeout 600 python bench.py --config $c --steps 3 --warmup 2 --no-cpu > gpurun_out/r03e_bench_c$c.json 2> gpurun_out/r03e_bench_c$c.err; python -c "
import json; d=json.load(open('gpurun_out/r03e_bench_c$c.json')); print('c$c', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['stages_ms_per_step'], round(d['roofline']['avg_launch_ms'],2), round(d['roofline']['frac'],4))"; done
