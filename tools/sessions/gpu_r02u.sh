#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants
timeout 1200 python -m pytest tests/test_gpu_tcs_engine.py -m gpu -q -x > gpurun_out/r02u_pytest_tcs.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed|Error|rel" gpurun_out/r02u_pytest_tcs.log | tail -5
timeout 900 python tools/ab.py --config 3 default $V/tcs_old.so 2>&1 | grep config
timeout 900 python tools/ab.py --config 5 default $V/tcs_old.so 2>&1 | grep config
