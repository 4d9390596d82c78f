"""A/B timing of libflucoma_b200 builds on the config-2 update loop (device-resident, 1024 buffers, 200 iterations)."""
import sys, os, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
ROOT = sys.argv[1]; lib = sys.argv[2]
sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200")]
import numpy as np, torch
import flucoma_b200 as fb
if lib != "default": fb.LIB_PATH = lib
from bench import make_audio, WORKLOAD as w
a = torch.from_numpy(make_audio(1024, w["n"], distinct=8)).cuda()
with fb.Plan(win=w["win"], hop=w["hop"], fft=w["fft"]) as plan:
    ts = []
    for _ in range(4):
        plan.bufnmf(a, w["rank"], w["iters"], seeds=np.arange(1024))
        ts.append(plan.stats()["ms_update_kernel"])
    print(lib, "update kernel ms:", " ".join("%.1f" % t for t in ts))
'''
for rep in range(2):
    for lib in sys.argv[1:]:
        subprocess.run([sys.executable, "-c", CHILD, ROOT, lib])
