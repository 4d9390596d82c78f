"""Developer aid: one iteration, many launches of identical buffers: which entries of W / H ever differ, and by how much?"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200")]
import numpy as np
import flucoma_b200 as fb
rng = np.random.default_rng(1)
F, B, NB = 512, 513, 148
base = ((rng.random((1, F, 6)) ** 3) @ (rng.random((1, 6, B)) ** 3) + 1e-3 * rng.random((1, F, B))).astype(np.float32)
X = np.concatenate([base] * NB)
uw, uh, iters = (sys.argv[1] == "1"), (sys.argv[2] == "1"), int(sys.argv[3])
events = 0
with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05) as plan:
    ref = None
    for rep in range(int(sys.argv[4])):
        W, H, _, _ = plan.nmf_process(X, 16, iters, uw, uh, seeds=np.zeros(NB, np.int64), want_v=False)
        if ref is None: ref = (W[0].copy(), H[0].copy())
        for b in range(NB):
            dw = np.argwhere(W[b] != ref[0]); dh = np.argwhere(H[b] != ref[1])
            if len(dw) or len(dh):
                events += 1
                if events <= 12:
                    print(f"rep {rep} buf {b}: W diffs {len(dw)}, H diffs {len(dh)}")
                    if len(dh):
                        rows = collections.Counter(int(f) for f, k in dh)
                        print("   H rows:", sorted(rows.items())[:16], " comps:", sorted(set(int(k) for f, k in dh)))
                        f, k = dh[0]; print("   e.g. H[%d][%d] = %.9g vs %.9g" % (f, k, H[b][f, k], ref[1][f, k]))
                    if len(dw):
                        cols = collections.Counter(int(bb) for k, bb in dw)
                        print("   W bins:", sorted(cols.items())[:16], " comps:", sorted(set(int(k) for k, bb in dw)))
print("events", events)
