import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import flucoma_b200 as fb
from test_tcgen05_blocks import bf16
rng = np.random.default_rng(0)
stats = []
with fb.Plan(win=64) as plan:
    for trial in range(8):
        H1 = rng.random((128, 16)) * (10.0 ** rng.uniform(-3, 0, (128, 1))); W1 = rng.random((16, 64)) * (10.0 ** rng.uniform(-3, 0, (1, 64)))
        R1 = rng.random((128, 64)) * 4; W2 = rng.random((16, 128)); H2 = rng.random((64, 16)); R2 = rng.random((128, 64)) * 4
        V = rng.random((128, 68))
        inp = np.concatenate([x.ravel() for x in (H1, W1, R1, W2, H2, R2, V)]).astype(np.float32)
        out = plan.selftest_tcgen05(inp)
        o1, o2, o3, o4, o5 = np.split(out, np.cumsum([128 * 64, 128 * 16, 128 * 64, 128 * 16]))
        H1b, W1b, R1b, W2b, H2b, R2b = (bf16(x).astype(np.float64) for x in (H1, W1, R1, W2, H2, R2))
        for name, got, want in (("K16", o1.reshape(128, 64), H1b @ W1b), ("K64", o2.reshape(128, 16), R1b @ W1b.T)):
            want32 = want.astype(np.float32).astype(np.float64)   # correctly rounded fp32 of the exact value
            e = (got.astype(np.float64) - want) / want
            e32 = (want32 - want) / want
            stats.append((name, e.mean(), np.abs(e).mean(), np.abs(e).max(), np.abs(e32).mean()))
for name in ("K16", "K64"):
    s = np.array([x[1:] for x in stats if x[0] == name])
    print(name, "mean rel err %.3e  mean|err| %.3e  max|err| %.3e   (RN fp32 mean|err| %.3e)" % tuple(s.mean(0)))
