#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nmf_tcs -c 1 -o gpurun_out/r02z_tcs32 python profiles/profile_cfg.py 3 148 12 > gpurun_out/r02z_ncu32.log 2>&1; tail -1 gpurun_out/r02z_ncu32.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nmf_tcs -c 1 -o gpurun_out/r02z_tcs16f python profiles/profile_cfg.py 5 1 10 > gpurun_out/r02z_ncu16.log 2>&1; tail -1 gpurun_out/r02z_ncu16.log
