"""Developer aid: where do identical buffers diverge on the tcgen05 engine?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200")]
import numpy as np
import flucoma_b200 as fb
rng = np.random.default_rng(1)
F, B = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 513
base = ((rng.random((1, F, 6)) ** 3) @ (rng.random((1, 6, B)) ** 3) + 1e-3 * rng.random((1, F, B))).astype(np.float32)

X = None
ITERS = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 2, 3, 6]
NB = int(sys.argv[3]) if len(sys.argv) > 3 else 16
for uw, uh in ((True, True), (False, True), (True, False)):
    for iters in (ITERS):
        X = np.concatenate([base] * NB)
        with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05) as plan:
            ref = None
            bad_w = bad_h = 0
            where_w, where_h = set(), set()
            for rep in range(6):
                W, H, _, _ = plan.nmf_process(X, 16, iters, uw, uh, seeds=np.zeros(NB, np.int64), want_v=False)
                if ref is None: ref = (W[0].copy(), H[0].copy())
                for b in range(NB):
                    dw = np.argwhere(W[b] != ref[0]); dh = np.argwhere(H[b] != ref[1])
                    bad_w += len(dw); bad_h += len(dh)
                    for k, bb in dw[:50]: where_w.add((int(k), int(bb)))
                    for f, k in dh[:50]: where_h.add((int(f), int(k)))
        print(f"uw={uw} uh={uh} iters={iters}: W diffs {bad_w} H diffs {bad_h}", flush=True)
        if where_w: print("   W (k,b):", sorted(where_w)[:24])
        if where_h: print("   H (f,k):", sorted(where_h)[:24])
