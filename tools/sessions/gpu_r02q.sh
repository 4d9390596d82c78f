#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== 4wg + fence"; FB200_LIB=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants/tc4fence.so timeout 600 python tools/diag_tc_race2.py 0 1 1 150 2>&1 | tail -1
FB200_LIB=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants/tc4fence.so timeout 600 python -m pytest tests/test_gpu_tcgen05_engine.py -m gpu -q 2>&1 | tail -2
echo "== default (2wg + fence)"
timeout 900 python -m pytest tests/test_gpu_tcgen05_engine.py tests/test_gpu_tcs_engine.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02q_pytest.log 2>&1; tail -2 gpurun_out/r02q_pytest.log
V=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants
timeout 600 python tools/ab.py default $V/tc2wg.so $V/tc4fence.so > gpurun_out/r02q_ab.log 2>&1; cat gpurun_out/r02q_ab.log
for c in 3 5; do timeout 300 python bench.py --config $c --steps 3 --warmup 2 --no-cpu > gpurun_out/r02q_bench_c$c.json 2> gpurun_out/r02q_bench_c$c.err; python -c "
import json; d=json.load(open('gpurun_out/r02q_bench_c$c.json')); print('c$c', d['ms_per_step'], d['stages_ms_per_step'], d['roofline']['avg_launch_ms'])"; done
