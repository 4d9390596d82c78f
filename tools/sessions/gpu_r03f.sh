#!/bin/bash
# final measurement session of round 2: tests, smoke, one bench line per config, the reference arm, ncu launch list and full captures
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r03f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r03f_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r03f_pytest.log | tail -5
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03f_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/r03f_bench_c2.json 2> gpurun_out/r03f_bench_c2.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r03f_bench_c2_ref.json 2> gpurun_out/r03f_bench_c2_ref.err
for c in 1 3 4 5; do timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r03f_bench_c$c.json 2> gpurun_out/r03f_bench_c$c.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03f_bench_c*.json")):
    try:
        d=json.load(open(f))
        if d.get("impl")=="reference": print(f, "reference", d["value"]); continue
        print(f, "ms %.2f"%d["ms_per_step"], "value %.3g"%d["value"], "e2e ms %.2f"%d["e2e"]["ms_per_step"], d["roofline"]["kernel"][:14], "kernel ms %.2f"%(d["roofline"]["avg_launch_ms"]*d["roofline"]["launches_per_step"]), "frac %.4f"%d["roofline"]["frac"], d.get("speedup_vs_all_threads"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e: print(f, "failed", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r03f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r03f_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nmf_tc\\b -c 1 -o gpurun_out/r03f_nmf_tc python profiles/profile_cfg.py 2 1024 200 > gpurun_out/r03f_ncu_tc.log 2>&1; tail -1 gpurun_out/r03f_ncu_tc.log
ls -la gpurun_out/r03f_*.ncu-rep
