#!/bin/bash
# HEAD verification after the late changes (fused inverse, SIMT finalisation in the tile kernel, sqrt magnitude)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r03w_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r03w_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03w_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/r03w_bench_c2_ref.json 2> gpurun_out/r03w_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r03w_bench_c2.json 2> gpurun_out/r03w_bench.err; echo "bench rc=$?"
for c in 1 3 4 5; do timeout 900 python bench.py --config $c --steps 5 --warmup 3 --no-cpu > gpurun_out/r03w_bench_c$c.json 2> gpurun_out/r03w_bench_c$c.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03w_bench_c*.json")):
    try:
        d=json.load(open(f))
        if d.get("impl")=="reference": print(f, "reference", d["value"]); continue
        print(f, "ms %.2f"%d["ms_per_step"], "value %.3g"%d["value"], "e2e ms %.2f"%d["e2e"]["ms_per_step"], "kernel ms %.2f"%(d["roofline"]["avg_launch_ms"]*d["roofline"]["launches_per_step"]), "frac %.4f"%d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["stages_ms_per_step"].get("ms_stft"))
    except Exception as e: print(f, "failed", e)
PY
