#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02m_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02m_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02m_pytest.log | tail -12
grep -E "^E  " gpurun_out/r02m_pytest.log | head -24
