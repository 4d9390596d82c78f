#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02g_pytest.log
tail -12 gpurun_out/r02g_pytest.log
# weak and strong scaling lines at N = 2 (config 3 strong: 8192 buffers in total is the 8-GPU size; at 2 GPUs use config 2)
for N in 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 --no-cpu > gpurun_out/r02g_bench_n${N}_weak.json 2> gpurun_out/r02g_bench_n${N}_weak.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 --no-cpu --config 3 --scaling strong > gpurun_out/r02g_bench_n${N}_c3strong.json 2> gpurun_out/r02g_bench_n${N}_c3strong.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02g_bench_*.json")):
    try:
        d=json.load(open(f)); print(f, d["n_gpus"], d["scaling"], "ms %.1f"%d["ms_per_step"], "value %.3g"%d["value"], "e2e %.3g"%d["e2e"]["value"])
    except Exception as e: print(f, "failed", e)
PY
