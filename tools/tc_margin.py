"""Developer aid: actual errors of an engine build against the fp64 oracle (the tests only assert the 1e-4 bar).
usage: tc_margin.py [backend: tc|tcs|simt] [rank]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200"), os.path.join(ROOT, "oracle")]
import numpy as np
import flucoma_b200 as fb
import c_oracle as oracle
which = sys.argv[1] if len(sys.argv) > 1 else "tc"
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 16
be = {"tc": fb.BACKEND_TCGEN05, "tcs": fb.BACKEND_TCGEN05_STREAMED, "simt": fb.BACKEND_SIMT}[which]
def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
rng = np.random.default_rng(7)
nb = 4
F, B = 512, 513
# spectrogram-like magnitudes: low-rank structure + noise floor, and one buffer of real |STFT| of the synthetic bench audio
X = (rng.random((nb, F, 8)) ** 3) @ (rng.random((nb, 8, B)) ** 3) + 1e-3 * rng.random((nb, F, B))
for iters in (50, 200):
    with fb.Plan(win=64, backend=be) as plan:
        W, H, V, _ = plan.nmf_process(X, rank, iters, True, True, seeds=np.arange(nb))
        assert plan.stats()["backend_used"] == be
    ew = eh = ev = 0.0
    for b in range(nb):
        Wo, Ho, Vo, _ = oracle.nmf_process(X[b], rank, iters, True, True, b)
        ew = max(ew, rel(W[b], Wo)); eh = max(eh, rel(H[b], Ho)); ev = max(ev, rel(V[b], Vo))
    print(f"{which} rank {rank} iters {iters}: max rel err W {ew:.2e} H {eh:.2e} V {ev:.2e}", flush=True)
