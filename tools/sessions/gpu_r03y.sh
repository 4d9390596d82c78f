#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nmf_tcs -c 1 -o gpurun_out/r03y_tcs64 python profiles/profile_cfg.py 4 148 12 > gpurun_out/r03y_ncu.log 2>&1; tail -1 gpurun_out/r03y_ncu.log
