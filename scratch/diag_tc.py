import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200")]
import numpy as np
import flucoma_b200 as fb
from oracle import c_oracle as co
from tests.golden.make_golden import synth_audio
def rel(a,b):
    a=np.asarray(a,np.float64); b=np.asarray(b,np.float64); return np.linalg.norm(a-b)/np.linalg.norm(b)
a = np.stack([synth_audio(1000+b, 130816) for b in (0,3)])
for iters in (10, 50, 100, 200):
    res = {}
    for name, be in (("simt", fb.BACKEND_SIMT), ("tc", fb.BACKEND_TCGEN05)):
        with fb.Plan(win=1024, hop=256, fft=1024, backend=be) as plan:
            res[name] = plan.bufnmf(a, 16, iters, seeds=[0,3])
    for i,b in enumerate((0,3)):
        o = co.bufnmf_channel(a[i], 1024,1024,256,16,iters,b)
        print(iters, b, "simt W %.2e H %.2e | tc W %.2e H %.2e | tc-vs-simt W %.2e" % (
            rel(res["simt"]["bases"][i], o["bases"]), rel(res["simt"]["acts"][i], o["acts"]),
            rel(res["tc"]["bases"][i], o["bases"]), rel(res["tc"]["acts"][i], o["acts"]),
            rel(res["tc"]["bases"][i], res["simt"]["bases"][i])), flush=True)
