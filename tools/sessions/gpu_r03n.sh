#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
N=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/r03n_bench_n${N}_weak.json 2> gpurun_out/r03n_bench_n${N}_weak.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 --no-cpu --config 3 --scaling strong > gpurun_out/r03n_bench_n${N}_c3strong.json 2> gpurun_out/r03n_bench_n${N}_c3strong.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03n_bench_*.json")):
    try:
        d=json.load(open(f)); print(f, d["n_gpus"], d["scaling"], "ms %.1f"%d["ms_per_step"], "value %.3g"%d["value"], "e2e %.3g"%d["e2e"]["value"], "e2e ms %.1f"%d["e2e"]["ms_per_step"])
    except Exception as e: print(f, "failed", e)
PY
