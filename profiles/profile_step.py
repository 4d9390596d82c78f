"""Short driver for ncu: one config-2 shaped BufNMF step with fewer iterations (same kernels, same grid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200")]
import numpy as np, torch
import flucoma_b200 as fb
from bench import make_audio, WORKLOAD as w

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
batch = int(sys.argv[2]) if len(sys.argv) > 2 else w["batch"]
a = torch.from_numpy(make_audio(batch, w["n"], distinct=4)).cuda()
with fb.Plan(win=w["win"], hop=w["hop"], fft=w["fft"]) as plan:
    for _ in range(2):
        plan.bufnmf(a, w["rank"], iters, seeds=np.arange(batch))
    print(plan.stats())
