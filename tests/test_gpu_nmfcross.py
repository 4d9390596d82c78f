"""GPU parity of BufNMFCross (fb200_bufnmfcross) against the fp64 oracle: NMFCross activations (NMFCross.hpp:60-185),
Griffin-Lim resynthesis (GriffinLim.hpp:29-54) and the client glue (NMFCrossClient.hpp:85-185)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def fb():
    import flucoma_b200
    return flucoma_b200


def test_nmfcross_golden(fb, golden_dir):
    g = np.load(os.path.join(golden_dir, "nmfcross.npz"))
    with fb.Plan(win=256, hop=64, fft=256) as plan:
        out, H, st = plan.bufnmfcross(g["source"], g["target"], 7, 11, 7, 30, seed=5, griffinlim_iterations=50)
    assert st == 0 and H.shape == g["H"].shape
    assert rel(H, g["H"]) < 1e-4, rel(H, g["H"])
    # the zero pattern left by the last iteration's sparseness / polyphony steps (hard selections) must be the same
    assert np.array_equal(H > 0, g["H"] > 0)
    assert rel(out, g["out"]) < 1e-3, rel(out, g["out"])  # 50 Griffin-Lim iterations in fp32 vs fp64


@pytest.mark.parametrize("ns,nt,win,hop,r,p,c,iters", [(9000, 7000, 512, 128, 7, 11, 7, 50), (3000, 8000, 256, 64, 3, 5, 3, 20),
                                                        (20000, 6000, 1024, 256, 5, 3, 9, 10)])
def test_nmfcross_vs_oracle(fb, oracle, ns, nt, win, hop, r, p, c, iters):
    from tests.golden.make_golden import synth_audio
    src = synth_audio(ns, ns); tgt = synth_audio(nt + 1, nt)
    with fb.Plan(win=win, hop=hop, fft=win) as plan:
        out, H, _ = plan.bufnmfcross(src, tgt, r, p, c, iters, seed=11, griffinlim_iterations=10)
        out_d, H_d, _ = plan.bufnmfcross(__import__("torch").from_numpy(src).cuda(), __import__("torch").from_numpy(tgt).cuda(), r, p, c,
                                         iters, seed=11, griffinlim_iterations=10)
    out_o, H_o = oracle.bufnmfcross(src, tgt, win, win, hop, r, p, c, iters, 11, 10)
    assert rel(H, H_o) < 1e-4 and rel(out, out_o) < 1e-3, (rel(H, H_o), rel(out, out_o))
    assert np.array_equal(H_d.cpu().numpy(), H) and np.array_equal(out_d.cpu().numpy(), out)  # device arrays: same path


def test_nmfcross_progress_cancel_and_errors(fb):
    from tests.golden.make_golden import synth_audio
    src = synth_audio(1, 4000); tgt = synth_audio(2, 4000)
    seen = []
    with fb.Plan(win=256, hop=64, fft=256) as plan:
        _, _, st = plan.bufnmfcross(src, tgt, 7, 11, 7, 12, seed=1, griffinlim_iterations=2, progress=lambda it: seen.append(it) or True)
        assert st == 0 and seen == list(range(1, 16))  # iterations + 3 (NMFCrossClient.hpp:152)
        seen.clear()
        _, _, st = plan.bufnmfcross(src, tgt, 7, 11, 7, 12, seed=1, progress=lambda it: seen.append(it) or it < 4)
        assert st == fb.CANCELLED and seen == [1, 2, 3, 4]
        with pytest.raises(fb.FlucomaB200Error, match="Time Sparsity is larger than target frames"):
            plan.bufnmfcross(src, tgt[:100], 7, 11, 7, 5)
