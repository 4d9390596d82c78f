#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02i_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02i_pytest.log | tail -15
grep -E "^E  " gpurun_out/r02i_pytest.log | head -20
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02i_bench.json')); print('c2 ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['stages_ms_per_step'], d.get('speedup_vs_all_threads'))"
