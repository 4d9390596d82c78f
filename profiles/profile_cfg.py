"""Short driver for ncu: one call of a BASELINE config shape with a reduced batch / iteration count (same kernels).
usage: profile_cfg.py <config 2|3|4|5> <batch> <iters>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200")]
import numpy as np, torch
import flucoma_b200 as fb
from bench import make_audio, CONFIGS

cfg, batch, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
w = CONFIGS[cfg]
with fb.Plan(win=w["win"], hop=w["hop"], fft=w["fft"]) as plan:
    if w["kind"] == "bufnmf":
        a = torch.from_numpy(make_audio(batch, w["n"], distinct=min(batch, 8))).cuda()
        for _ in range(2):
            plan.bufnmf(a, w["rank"], iters, seeds=np.arange(batch))
    else:
        F = 200_000
        X = torch.rand((F, plan.bins), device="cuda") ** 2
        W = torch.rand((w["rank"], plan.bins), device="cuda")
        for _ in range(2):
            plan.nmf_process_frames(X, W, iters, seed=42)
    print(plan.stats())
