#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants
timeout 900 python -m pytest tests/test_gpu_tcs_engine.py -m gpu -q -x > gpurun_out/r02y_pytest.log 2>&1; tail -2 gpurun_out/r02y_pytest.log
FB200_TCS_TIMELINE=1 FB200_LIB=$V/tcs_tl.so timeout 600 python profiles/profile_cfg.py 4 148 4 > gpurun_out/r02y_tl4.log 2>&1; grep -A24 "^job kind" gpurun_out/r02y_tl4.log | tail -10
FB200_TCS_TIMELINE=1 FB200_LIB=$V/tcs_tl.so timeout 600 python profiles/profile_cfg.py 3 148 4 > gpurun_out/r02y_tl3.log 2>&1; grep -A12 "^job kind" gpurun_out/r02y_tl3.log | tail -10
for c in 3 4 5; do timeout 600 python bench.py --config $c --steps 3 --warmup 2 --no-cpu > gpurun_out/r02y_bench_c$c.json 2> gpurun_out/r02y_bench_c$c.err; python -c "
import json; d=json.load(open('gpurun_out/r02y_bench_c$c.json')); print('c$c', round(d['ms_per_step'],2), d['roofline']['kernel'][:12], round(d['roofline']['avg_launch_ms'],2), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"; done
