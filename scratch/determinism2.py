"""Developer aid: which stage of bufnmf is not bitwise repeatable?
usage: python scratch/determinism2.py [reps] [batch] [iters] [backend]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200")]
import numpy as np, torch
import flucoma_b200 as fb
from bench import make_audio, WORKLOAD as w

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
backend = int(sys.argv[4]) if len(sys.argv) > 4 else 0
n, K = w["n"], w["rank"]
ad = torch.from_numpy(make_audio(batch, n)).cuda()
seeds = np.arange(batch, dtype=np.int64)
plan = fb.Plan(win=w["win"], hop=w["hop"], fft=w["fft"], max_rank=K, max_batch=batch, max_samples=n, backend=backend)
F, B = fb.num_frames(n, w["win"], w["hop"]), plan.bins
Vref = None
nbad = 0
for i in range(reps):
    _, V = plan.stft(ad, want_spectrum=False, want_magnitude=True)
    if Vref is None:
        Vref = V.clone()
    elif not torch.equal(V, Vref):
        d = (V - Vref).abs().amax(dim=(1, 2)); bad = torch.nonzero(d > 0).flatten().tolist()
        nbad += 1
        print("stft rep", i, "differs in buffers", len(bad), bad[:10], "max", float(d.max()), flush=True)
print("stft: %d of %d repeats differ" % (nbad, reps - 1), flush=True)
out = {"bases": torch.empty((batch, K, B), device="cuda"), "acts": torch.empty((batch, F, K), device="cuda")}
ref = None
for i in range(reps):
    plan.bufnmf(ad, K, iters, seeds=seeds, out=out)
    torch.cuda.synchronize()
    if ref is None:
        ref = {k: v.clone() for k, v in out.items()}
        continue
    msg = []
    for k in ("bases", "acts"):
        d = (out[k] - ref[k]).abs().amax(dim=(1, 2)); bad = torch.nonzero(d > 0).flatten().tolist()
        if bad:
            rel = float(((out[k] - ref[k]).norm() / ref[k].norm()))
            msg.append("%s: %d buffers %s max %.3g rel %.3g" % (k, len(bad), bad[:10], float(d.max()), rel))
    print("bufnmf rep", i, "; ".join(msg) if msg else "equal", flush=True)
