#!/bin/bash
# round-2 final measurement session: tests, smoke, bench lines for all five BASELINE configs, ncu launch list + full captures
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02n_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02n_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02n_pytest.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02n_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02n_smoke.log; tail -1 gpurun_out/r02n_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02n_bench_c2.json 2> gpurun_out/r02n_bench_c2.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02n_bench_c2_ref.json 2> gpurun_out/r02n_bench_c2_ref.err
timeout 600 python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/r02n_bench_c1.json 2> gpurun_out/r02n_bench_c1.err
timeout 900 python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/r02n_bench_c3.json 2> gpurun_out/r02n_bench_c3.err
timeout 1200 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02n_bench_c4.json 2> gpurun_out/r02n_bench_c4.err
timeout 600 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r02n_bench_c5.json 2> gpurun_out/r02n_bench_c5.err
python - <<'PY'
import json
for n in ("c1","c2","c3","c4","c5"):
    try:
        d=json.load(open(f"gpurun_out/r02n_bench_{n}.json")); r=d["roofline"]
        print(n, "ms/step %.2f"%d["ms_per_step"], "value %.3g"%d["value"], "e2e %.3g"%d["e2e"]["value"], r["kernel"][:12], "kernel ms %.2f"%r["avg_launch_ms"], "achieved %.1f %s frac %.4f"%(r["achieved"], r["unit"], r["frac"]), d.get("speedup_vs_all_threads"), d.get("speedup_vs_1thread"))
    except Exception as e: print(n, "failed", e)
PY
# ncu: launch list of the bench command, then full captures of the two dominant kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02n_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02n_bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_nmf_tc -s 1 -c 1 -o gpurun_out/r02n_nmf_tc python profiles/profile_cfg.py 2 1024 200 > gpurun_out/r02n_ncu_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stft_fused -s 4 -c 1 -o gpurun_out/r02n_stft_fused python profiles/profile_cfg.py 2 1024 2 > gpurun_out/r02n_ncu_stft.log 2>&1
ls -la gpurun_out/r02n_*.ncu-rep
