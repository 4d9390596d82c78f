#!/bin/bash
# HEAD verification: the three commands the driver runs at round end
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r03l_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r03l_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03l_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r03l_smoke.log | cut -c1-200
timeout 900 python bench.py --impl reference > gpurun_out/r03l_ref.json 2> gpurun_out/r03l_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r03l_bench.json 2> gpurun_out/r03l_bench.err; echo "bench rc=$?"
python -c "
import json
r=json.load(open('gpurun_out/r03l_ref.json')); d=json.load(open('gpurun_out/r03l_bench.json'))
print('reference', r['value'], 'ours', d['value'], 'e2e', d['e2e']['value'], 'ratio e2e', d['e2e']['value']/r['value'], 'frac', d['roofline']['frac'], d['clocks'])
print(sorted(d.keys()))"
