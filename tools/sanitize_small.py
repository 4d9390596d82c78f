"""Small calls through every engine and entry point, meant to run under compute-sanitizer (memcheck / synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200")]
import numpy as np
import flucoma_b200 as fb
rng = np.random.default_rng(0)
def lowrank(b, F, B): return ((rng.random((b, F, 5)) ** 3) @ (rng.random((b, 5, B)) ** 3) + 1e-3 * rng.random((b, F, B))).astype(np.float32)
with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05) as p:
    p.nmf_process(lowrank(3, 256, 257), 16, 3, True, True, seeds=[1, 2, 3]); print("resident ok", flush=True)
for K in (16, 32, 64):
    with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05_STREAMED) as p:
        p.nmf_process(lowrank(2, 256, 257), K, 2, True, True, seeds=[1, 2])
        p.nmf_process(lowrank(2, 200, 129), K, 2, False, True, seeds=[1, 2])
        p.nmf_process_frames(np.abs(rng.standard_normal((300, 257))).astype(np.float32), rng.random((K, 257)).astype(np.float32), 3, seed=1)
    print("streamed", K, "ok", flush=True)
with fb.Plan(win=64, backend=fb.BACKEND_SIMT) as p:
    p.nmf_process(lowrank(2, 100, 33), 5, 3, True, True, seeds=[1, 2]); print("simt ok", flush=True)
a = (rng.standard_normal((2, 8192)) * 0.1).astype(np.float32)
with fb.Plan(win=256, hop=64, fft=256) as p:
    p.bufnmf(a, 4, 3, seeds=[1, 2], resynth=True); print("bufnmf ok", flush=True)
    p.stft(a); print("stft ok", flush=True)
