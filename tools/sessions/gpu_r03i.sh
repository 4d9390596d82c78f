#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tcs_engine.py tests/test_gpu_tcgen05_engine.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r03i_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r03i_pytest.log | tail -4
for k in tc; do timeout 600 python tools/tc_margin.py $k 2>&1 | tail -1; done
timeout 600 python tools/tc_margin.py tcs 32 2>&1 | tail -1
for c in 2 3 5; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu > gpurun_out/r03i_bench_c$c.json 2> gpurun_out/r03i_bench_c$c.err; python -c "
import json; d=json.load(open('gpurun_out/r03i_bench_c$c.json')); print('c$c', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), round(d['roofline']['avg_launch_ms']*d['roofline']['launches_per_step'],2), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"; done
