#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants
for n in PTERMS3 ROWPAD BOTH; do
echo "== $n"
FB200_LIB=$V/tc_$n.so timeout 900 python -m pytest tests/test_gpu_tcgen05_engine.py tests/test_gpu_parity.py -m gpu -q -k "tc_engine or config2_full_size or kl_divergence or config1" > gpurun_out/r02r_parity_$n.log 2>&1; tail -3 gpurun_out/r02r_parity_$n.log
done
timeout 900 python tools/ab.py default $V/tc_PTERMS3.so $V/tc_ROWPAD.so $V/tc_BOTH.so > gpurun_out/r02r_ab.log 2>&1; cat gpurun_out/r02r_ab.log
