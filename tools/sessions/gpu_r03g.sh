#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants
timeout 900 python -m pytest tests/test_gpu_tcs_engine.py -m gpu -q -x > gpurun_out/r03g_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r03g_pytest.log | tail -4
timeout 900 python tools/ab.py --config 3 default $V/tcs_pf0.so $V/tcs_pf6.so 2>&1 | grep config
timeout 1200 python tools/ab.py --config 4 default $V/tcs_pf0.so $V/tcs_pf6.so 2>&1 | grep config
timeout 900 python tools/ab.py --config 5 default $V/tcs_pf0.so $V/tcs_pf6.so 2>&1 | grep config
