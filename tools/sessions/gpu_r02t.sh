#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tcs_engine.py -m gpu -q -x > gpurun_out/r02t_pytest_tcs.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed|Error|rel" gpurun_out/r02t_pytest_tcs.log | tail -8
for k in 16 32 64; do timeout 600 python tools/tc_margin.py tcs $k 2>&1 | tail -2; done
for c in 3 4 5; do timeout 600 python bench.py --config $c --steps 3 --warmup 2 --no-cpu > gpurun_out/r02t_bench_c$c.json 2> gpurun_out/r02t_bench_c$c.err; python -c "
import json; d=json.load(open('gpurun_out/r02t_bench_c$c.json')); print('c$c', d['ms_per_step'], d['stages_ms_per_step'], d['roofline']['kernel'][:12], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"; done
