#!/bin/bash
# round-2 GPU session A: test suite, smoke, bench (config 2), A/B of the paired-reciprocal epilogue
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02a_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
timeout 600 python tools/ab.py default $GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants/rcppair.so > gpurun_out/r02a_ab.log 2>&1
tail -3 gpurun_out/r02a_pytest.log; tail -2 gpurun_out/r02a_smoke.log; cat gpurun_out/r02a_ab.log; head -c 3000 gpurun_out/r02a_bench.json
