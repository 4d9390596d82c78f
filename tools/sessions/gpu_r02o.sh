#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tcgen05_engine.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02o_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02o_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r02o_pytest.log | tail -8
V=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants/tc2wg.so
timeout 600 python tools/ab.py default $V > gpurun_out/r02o_ab.log 2>&1; cat gpurun_out/r02o_ab.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r02o_bench_c2.json 2> gpurun_out/r02o_bench_c2.err
python -c "
import json; d=json.load(open('gpurun_out/r02o_bench_c2.json')); print('c2', d['ms_per_step'], d['stages_ms_per_step'])"
