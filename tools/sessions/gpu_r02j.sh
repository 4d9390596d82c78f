#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02j_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02j_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02j_pytest.log | tail -8
grep -E "^E  " gpurun_out/r02j_pytest.log | head -10
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stft_fused -s 4 -c 1 -o gpurun_out/r02j_stft_fused python profiles/profile_cfg.py 2 1024 2 > gpurun_out/r02j_ncu_stft.log 2>&1
ls -la gpurun_out/r02j_stft_fused.ncu-rep
