#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tcgen05_engine.py tests/test_gpu_parity.py tests/test_host_cpp.py tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r03d_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r03d_pytest.log | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r03d_bench_c2.json 2> gpurun_out/r03d_bench_c2.err; python -c "
import json; d=json.load(open('gpurun_out/r03d_bench_c2.json')); print('c2', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'e2e_cb', round(d['e2e_callback']['ms_per_step'],2), d['stages_ms_per_step'], d['roofline']['launches_per_step'], round(d['roofline']['avg_launch_ms'],2), round(d['roofline']['frac'],4), d['gpu_launches'])"
tail -3 gpurun_out/r03d_bench_c2.err
