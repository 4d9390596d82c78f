"""Short driver for ncu: the SIMT update engine on BASELINE config 3 / 4 shapes (rank 32 / rank 64), few iterations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200")]
import numpy as np, torch
import flucoma_b200 as fb
from bench import make_audio

cfg = sys.argv[1] if len(sys.argv) > 1 else "3"
win, hop, K, batch = (1024, 256, 32, 296) if cfg == "3" else (4096, 1024, 64, 74)
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n = 511 * hop
a = torch.from_numpy(make_audio(batch, n, distinct=4)).cuda()
with fb.Plan(win=win, hop=hop, fft=win, max_rank=K) as plan:
    for _ in range(2):
        plan.bufnmf(a, K, iters, seeds=np.arange(batch))
    print(plan.stats())
