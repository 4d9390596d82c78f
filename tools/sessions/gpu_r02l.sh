#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02l_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02l_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02l_pytest.log | tail -5
V=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants/noxlo.so
FB200_LIB=$V timeout 900 python -m pytest tests/test_gpu_tcgen05_engine.py tests/test_gpu_parity.py -m gpu -q -k "vs_oracle or config2_full_size or kl_divergence or matches_simt" > gpurun_out/r02l_noxlo_parity.log 2>&1
tail -4 gpurun_out/r02l_noxlo_parity.log
timeout 600 python tools/ab.py default $V > gpurun_out/r02l_ab.log 2>&1; cat gpurun_out/r02l_ab.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r02l_bench_c2.json 2> gpurun_out/r02l_bench_c2.err
python -c "
import json; d=json.load(open('gpurun_out/r02l_bench_c2.json')); print('c2', d['ms_per_step'], d['stages_ms_per_step'])"
