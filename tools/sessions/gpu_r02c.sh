#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_pytest.log
tail -5 gpurun_out/r02c_pytest.log
FB200_BACKEND=3 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r02c_bench_c2_streamed.json 2> gpurun_out/r02c_bench_c2_streamed.err
timeout 600 python bench.py --config 3 --steps 3 --warmup 2 --no-cpu > gpurun_out/r02c_bench_c3.json 2> gpurun_out/r02c_bench_c3.err
timeout 600 python bench.py --config 5 --steps 3 --warmup 2 --no-cpu > gpurun_out/r02c_bench_c5.json 2> gpurun_out/r02c_bench_c5.err
python - <<'PY'
import json
for n in ("c2_streamed","c3","c5"):
    try:
        d=json.load(open(f"gpurun_out/r02c_bench_{n}.json")); r=d["roofline"]
        print(n, "ms/step %.2f"%d["ms_per_step"], "e2e %.2f"%d["e2e"]["ms_per_step"], r["kernel"][:10], "kernel ms %.2f"%r["avg_launch_ms"], "achieved %.1f %s frac %.4f"%(r["achieved"], r["unit"], r["frac"]))
    except Exception as e: print(n, "failed", e)
PY
