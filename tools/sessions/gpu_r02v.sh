#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02v_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02v_pytest.log | tail -5
for c in 2 3 4 5; do timeout 600 python bench.py --config $c --steps 3 --warmup 2 --no-cpu > gpurun_out/r02v_bench_c$c.json 2> gpurun_out/r02v_bench_c$c.err; python -c "
import json; d=json.load(open('gpurun_out/r02v_bench_c$c.json')); print('c$c', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['roofline']['kernel'][:12], round(d['roofline']['avg_launch_ms'],2), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"; done
