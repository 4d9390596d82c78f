"""A/B timing of libflucoma_b200 builds on a BASELINE config's update loop (device-resident).
usage: ab.py [--config 2|3|4|5] default|<path to .so> ..."""
import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
ROOT = sys.argv[1]; lib = sys.argv[2]; cfg = int(sys.argv[3])
sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200")]
import numpy as np, torch
import flucoma_b200 as fb
if lib != "default": fb.LIB_PATH = lib
from bench import make_audio, CONFIGS
w = CONFIGS[cfg]
with fb.Plan(win=w["win"], hop=w["hop"], fft=w["fft"]) as plan:
    ts = []
    if w["kind"] == "bufnmf":
        batch = w["batch"]
        a = torch.from_numpy(make_audio(batch, w["n"], distinct=8)).cuda()
        for _ in range(4):
            plan.bufnmf(a, w["rank"], w["iters"], seeds=np.arange(batch))
            ts.append(plan.stats()["ms_update_kernel"])
    else:
        X = torch.rand((1_000_000, plan.bins), device="cuda") ** 2
        W = torch.rand((w["rank"], plan.bins), device="cuda")
        for _ in range(4):
            plan.nmf_process_frames(X, W, w["iters"], seed=42)
            ts.append(plan.stats()["ms_update_kernel"])
    print("config", cfg, lib.split("/")[-1], "update kernel ms:", " ".join("%.2f" % t for t in ts), flush=True)
'''
args = sys.argv[1:]
cfg = 2
if args and args[0] == "--config": cfg = int(args[1]); args = args[2:]
for rep in range(2):
    for lib in args:
        subprocess.run([sys.executable, "-c", CHILD, ROOT, lib, str(cfg)])
