#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02e_pytest.log
tail -4 gpurun_out/r02e_pytest.log
# ncu: streamed engine on config 3 (rank 32), 148 buffers x 20 iterations
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nmf_tcs -s 1 -c 1 -o gpurun_out/r02e_tcs_c3 python profiles/profile_cfg.py 3 148 20 > gpurun_out/r02e_ncu_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nmf_tcs -s 1 -c 1 -o gpurun_out/r02e_tcs_c5 python profiles/profile_cfg.py 5 1 10 > gpurun_out/r02e_ncu_c5.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
