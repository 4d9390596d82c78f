#!/bin/bash
# soak: the engine test files five times in a row (rare races show up as a flaky run)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3 4 5; do timeout 900 python -m pytest tests/test_gpu_tcgen05_engine.py tests/test_gpu_tcs_engine.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -1; done
