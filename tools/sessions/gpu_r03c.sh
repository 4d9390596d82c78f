#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants
timeout 900 python -m pytest tests/test_gpu_tcs_engine.py -m gpu -q -x > gpurun_out/r03c_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r03c_pytest.log | tail -5
FB200_TCS_TIMELINE=1 FB200_LIB=$V/tcs_tl.so timeout 600 python profiles/profile_cfg.py 3 148 4 > gpurun_out/r03c_tl3.log 2>&1; grep -A20 "^job kind" gpurun_out/r03c_tl3.log | tail -12
timeout 900 python tools/ab.py --config 3 default $V/tcs_noupd.so 2>&1 | grep config
timeout 900 python tools/ab.py --config 5 default $V/tcs_noupd.so 2>&1 | grep config
