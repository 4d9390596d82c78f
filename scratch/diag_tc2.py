import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200")]
import numpy as np
import flucoma_b200 as fb
from tests.golden.make_golden import synth_audio
np.set_printoptions(linewidth=200, precision=2)
a = synth_audio(1000, 130816)[None]
with fb.Plan(win=1024, hop=256, fft=1024) as plan:
    _, mag = plan.stft(a, want_spectrum=False, want_magnitude=True)
X = mag[0].astype(np.float64)
res = {}
for name, be in (("simt", fb.BACKEND_SIMT), ("tc", fb.BACKEND_TCGEN05)):
    with fb.Plan(win=1024, hop=256, fft=1024, backend=be) as plan:
        res[name] = plan.nmf_process(X, 16, 100, True, True, seeds=0)
Ws, Hs, Vs, _ = res["simt"]; Wt, Ht, Vt, _ = res["tc"]
print("rel W %.2e H %.2e V %.2e" % (np.linalg.norm(Wt-Ws)/np.linalg.norm(Ws), np.linalg.norm(Ht-Hs)/np.linalg.norm(Hs), np.linalg.norm(Vt-Vs)/np.linalg.norm(Vs)))
dH = Ht - Hs
scale = (dH * Hs).sum(0) / (Hs * Hs).sum(0)            # per-component scale drift of H
resid = dH - Hs * scale
print("H per-component scale eps_k:", scale)
print("H comp energy share:", (Hs*Hs).sum(0)/ (Hs*Hs).sum())
print("H err: scale part %.2e residual %.2e" % (np.linalg.norm(Hs*scale)/np.linalg.norm(Hs), np.linalg.norm(resid)/np.linalg.norm(Hs)))
dW = Wt - Ws
print("W per-component rel err:", np.linalg.norm(dW,axis=1)/np.linalg.norm(Ws,axis=1))
# similarity between components (overlap)
G = Ws @ Ws.T
print("max offdiag W gram:", np.sort(np.abs(G - np.eye(16)).max(1))[::-1][:8])
# KL divergence of both
def kl(V, X):
    Vc = np.maximum(V, 1e-300); Xc = np.maximum(X, 1e-300)
    return np.sum(X*np.log(Xc/Vc) - X + V)
print("KL simt %.10e tc %.10e" % (kl(Vs, X), kl(Vt, X)))
