#!/bin/bash
# round-2 GPU session B: first run of the streamed engine
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcs_engine.py -x -q -k "vs_oracle" > gpurun_out/r02b_tcs1.log 2>&1; rc1=$?; echo "rc=$rc1" >> gpurun_out/r02b_tcs1.log
tail -15 gpurun_out/r02b_tcs1.log
if [ $rc1 -eq 0 ]; then
  timeout 900 python -m pytest tests/test_gpu_tcs_engine.py -x -q -k "not vs_oracle" > gpurun_out/r02b_tcs2.log 2>&1; rc2=$?; echo "rc=$rc2" >> gpurun_out/r02b_tcs2.log
  tail -15 gpurun_out/r02b_tcs2.log
  if [ $rc2 -eq 0 ]; then
    timeout 600 python bench.py --config 3 --steps 3 --warmup 2 --no-cpu > gpurun_out/r02b_bench_c3.json 2> gpurun_out/r02b_bench_c3.err
    timeout 600 python bench.py --config 5 --steps 3 --warmup 2 --no-cpu > gpurun_out/r02b_bench_c5.json 2> gpurun_out/r02b_bench_c5.err
    head -c 1500 gpurun_out/r02b_bench_c3.json; head -c 1500 gpurun_out/r02b_bench_c5.json; tail -3 gpurun_out/r02b_bench_c5.err
  fi
fi
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02b_bench.json')); print('c2', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_callback'])"
