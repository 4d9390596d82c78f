#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants
FB200_TC_TIMELINE=1 FB200_LIB=$V/tc_tl.so timeout 600 python profiles/profile_cfg.py 2 148 6 > gpurun_out/r03h_tl2.log 2>&1
FB200_TC_TIMELINE=1 FB200_LIB=$V/tc4_tl.so timeout 600 python profiles/profile_cfg.py 2 148 6 > gpurun_out/r03h_tl4.log 2>&1
grep -A34 "^step |" gpurun_out/r03h_tl2.log | tail -36
grep -A34 "^step |" gpurun_out/r03h_tl4.log | tail -36
