#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02f_pytest.log
tail -25 gpurun_out/r02f_pytest.log
