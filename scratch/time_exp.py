import sys, os, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200")]
import numpy as np, torch
import flucoma_b200 as fb
lib = sys.argv[1] if len(sys.argv) > 1 else None
if lib: fb.LIB_PATH = lib
from bench import make_audio, WORKLOAD as w
a = torch.from_numpy(make_audio(148, w["n"], distinct=4)).cuda()
with fb.Plan(win=w["win"], hop=w["hop"], fft=w["fft"]) as plan:
    for _ in range(2):
        plan.bufnmf(a, w["rank"], 20, seeds=np.arange(148))
    st = plan.stats()
    print(lib, "ms_update_kernel %.3f -> per pass %.1f us = %.0f cycles @1.965GHz" % (st["ms_update_kernel"], 1e3*st["ms_update_kernel"]/21, 1.965e3*1e3*st["ms_update_kernel"]/21))
