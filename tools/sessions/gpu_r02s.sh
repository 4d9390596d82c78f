#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants
echo "== 6 terms + X_lo (round-1 arithmetic)"; FB200_LIB=$V/tc2wg.so timeout 600 python tools/tc_margin.py tc 2>&1 | tail -2
echo "== 3 terms only"; FB200_LIB=$V/tc_PTERMS3.so timeout 600 python tools/tc_margin.py tc 2>&1 | tail -2
echo "== default: 3 terms, no X_lo"; timeout 600 python tools/tc_margin.py tc 2>&1 | tail -2
timeout 600 python tools/tc_margin.py simt 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_tcgen05_engine.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/r02s_pytest.log 2>&1; tail -2 gpurun_out/r02s_pytest.log
timeout 900 python tools/ab.py default $V/tc2wg.so > gpurun_out/r02s_ab.log 2>&1; cat gpurun_out/r02s_ab.log
