"""Developer aid: is the NMF engine bitwise repeatable?  Which buffers differ, by how much?
usage: python scratch/determinism.py [reps] [batch] [iters] [backend: 0 auto|1 simt|2 tc]"""
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
if len(sys.argv) > 5: fb.LIB_PATH = os.path.abspath(sys.argv[5])
n, K = w["n"], w["rank"]
ad = torch.from_numpy(make_audio(batch, n)).cuda()
seeds = np.arange(batch, dtype=np.int64)
plan = fb.Plan(win=w["win"], hop=w["hop"], fft=w["fft"], max_rank=K, max_batch=batch, max_samples=n, backend=backend)
_, V = plan.stft(ad, want_spectrum=False, want_magnitude=True)
_, V2 = plan.stft(ad, want_spectrum=False, want_magnitude=True)
print("stft repeatable:", bool(torch.equal(V, V2)), flush=True)
ref = None
for i in range(reps):
    t0 = time.perf_counter()
    try:
        W1, H1, _, st = plan.nmf_process(V, K, iters, seeds=seeds, want_v=False)
    except Exception as e:
        print("FAIL rep", i, e, flush=True)
        sys.exit(1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if ref is None:
        ref = (W1.clone(), H1.clone())
        print(i, "%.1f ms" % (1e3 * dt), "reference run; finite:", bool(torch.isfinite(W1).all() and torch.isfinite(H1).all()), flush=True)
        continue
    dW = (W1 - ref[0]).abs().amax(dim=(1, 2)); dH = (H1 - ref[1]).abs().amax(dim=(1, 2))
    bad = torch.nonzero((dW > 0) | (dH > 0)).flatten().tolist()
    nbad_runs = globals().get("nbad_runs", 0) + (1 if bad else 0)
    if bad or i == reps - 1:
        print(i, "%.1f ms" % (1e3 * dt), "differing buffers:", len(bad), bad[:12],
              "max dW %.3g dH %.3g" % (float(dW.max()), float(dH.max())), "| runs with a difference so far:", nbad_runs, flush=True)
print("done")
