#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tool in memcheck synccheck; do
timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r03k_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|ok$|Error|error" gpurun_out/r03k_$tool.log | tail -12
done
