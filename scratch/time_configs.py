"""Developer aid: device-resident timing of BASELINE configs 3, 4 (SIMT engine) and 5 (fixed-bases stream) on one GPU.
usage: python scratch/time_configs.py [batch3] [batch4]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200")]
import numpy as np, torch
import flucoma_b200 as fb
from bench import make_audio

def run(name, batch, win, hop, K, iters, F=512):
    n = (F - 1) * hop
    a = torch.from_numpy(make_audio(batch, n, distinct=4)).cuda()
    with fb.Plan(win=win, hop=hop, fft=win, max_rank=K) as plan:
        for _ in range(2):
            plan.bufnmf(a, K, iters, seeds=np.arange(batch))
            st = plan.stats()
    B = win // 2 + 1
    fl = 8.0 * B * K * iters * batch * F
    print("%s: batch %d fft %d K %d iters %d: total %.1f ms (stft %.1f, nmf %.1f) -> %.3e frames/s, update loop %.1f TFLOP/s algorithmic, backend %d" % (
        name, batch, win, K, iters, st["ms_total"], st["ms_stft"], st["ms_nmf"], batch * F / st["ms_total"] * 1e3, fl / st["ms_nmf"] * 1e-9, st["backend_used"]), flush=True)

b3 = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b4 = int(sys.argv[2]) if len(sys.argv) > 2 else 64
run("config2", 1024, 1024, 256, 16, 200)
run("config3 shard", b3, 1024, 256, 32, 200)
run("config4 (iters 100 of 500)", b4, 4096, 1024, 64, 100)
