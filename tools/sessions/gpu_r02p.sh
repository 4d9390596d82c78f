#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in LATE_VEMPTY LATE_PFREE NO_SETMAXNREG; do
echo "== $v"
FB200_LIB=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants/tc4_$v.so timeout 600 python tools/diag_tc_race2.py 0 1 1 100 2>&1 | tail -1
done
echo "== default"; timeout 600 python tools/diag_tc_race2.py 0 1 1 100 2>&1 | tail -1
