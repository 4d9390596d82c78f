#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants
timeout 900 python -m pytest tests/test_gpu_tcs_engine.py -m gpu -q -x -k "64 or 40" > gpurun_out/r02x_pytest.log 2>&1; tail -2 gpurun_out/r02x_pytest.log
echo "== group 4 (default)"; timeout 600 python tools/tc_margin.py tcs 64 2>&1 | tail -2
for g in 1 2 8; do echo "== group $g"; FB200_LIB=$V/tcs_g$g.so timeout 600 python tools/tc_margin.py tcs 64 2>&1 | tail -2; done
timeout 1200 python tools/ab.py --config 4 default $V/tcs_g1.so $V/tcs_g2.so $V/tcs_g8.so 2>&1 | grep config
