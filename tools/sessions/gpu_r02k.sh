#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02k_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02k_pytest.log | tail -8
grep -E "^E  " gpurun_out/r02k_pytest.log | head -10
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r02k_bench_c2.json 2> gpurun_out/r02k_bench_c2.err
FB200_BACKEND=3 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r02k_bench_c2_tcr.json 2> gpurun_out/r02k_bench_c2_tcr.err
FB200_BACKEND=3 FB200_TCS16=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r02k_bench_c2_tcs16.json 2> gpurun_out/r02k_bench_c2_tcs16.err
timeout 600 python bench.py --config 5 --steps 3 --warmup 2 --no-cpu > gpurun_out/r02k_bench_c5.json 2> gpurun_out/r02k_bench_c5.err
python - <<'PY'
import json
for n in ("c2","c2_tcr","c2_tcs16","c5"):
    try:
        d=json.load(open(f"gpurun_out/r02k_bench_{n}.json")); r=d["roofline"]
        print(n, "ms/step %.2f"%d["ms_per_step"], "e2e %.2f"%d["e2e"]["ms_per_step"], "stft %.2f"%d["stages_ms_per_step"]["ms_stft"], r["kernel"][:10], "kernel ms %.2f"%r["avg_launch_ms"], "achieved %.1f %s frac %.4f"%(r["achieved"], r["unit"], r["frac"]))
    except Exception as e: print(n, "failed", e)
PY
