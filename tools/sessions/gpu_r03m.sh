#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants
FB200_LIB=$V/tc_pairs.so timeout 900 python -m pytest tests/test_gpu_tcgen05_engine.py -m gpu -q -x > gpurun_out/r03j_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r03j_pytest.log | tail -4
timeout 900 python -m pytest tests/test_gpu_tcgen05_engine.py -m gpu -q -x 2>&1 | tail -1
timeout 900 python tools/ab.py default $V/tc_pairs.so 2>&1 | grep config
