"""GPU parity of the STFT epilogues against the fp64 oracle: fb200_melbands (MelBands.hpp:43-101) and
fb200_hpss (HPSS.hpp:47-162)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return np.linalg.norm(a.astype(np.complex128) - b.astype(np.complex128)) / max(np.linalg.norm(b.astype(np.complex128)), 1e-300)


@pytest.fixture(scope="module")
def fb():
    import flucoma_b200
    return flucoma_b200


def test_melbands_golden_and_from_audio(fb, oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "spectral.npz"))
    a = g["audio"]
    S = oracle.stft(a.astype(np.float64), 512, 512, 128)
    M = np.abs(S).astype(np.float32)
    with fb.Plan(win=512, hop=128, fft=512) as plan:
        b = plan.melbands(mags=M, n_bands=40)
        b_db = plan.melbands(mags=M, n_bands=40, mag_norm=False, use_power=True, log_output=True)
        b_audio = plan.melbands(audio=a, n_bands=40)
        batch = plan.melbands(mags=np.stack([M, 2 * M]), n_bands=13, lo=100.0, hi=8000.0)
    assert rel(b, g["mel_norm"]) < 1e-5 and rel(b_audio, g["mel_norm"]) < 1e-5
    assert np.abs(b_db - g["mel_pow_db"]).max() < 2e-3  # dB values: absolute tolerance (fp32 log10 of small band energies)
    ref13 = oracle.melbands(M.astype(np.float64), 100.0, 8000.0, 13, 44100.0, 512)
    assert rel(batch[0], ref13) < 1e-5 and rel(batch[1], 2 * ref13) < 1e-5  # magNorm output scales with the input


@pytest.mark.parametrize("mode,ht,pt", [(0, (0, 1, 1, 1), (0, 1, 1, 1)), (1, (0.005, 0.0, 0.1, -10.0), (0, 1, 1, 1)),
                                       (2, (0.005, 10.0, 0.1, 0.0), (0.005, 10.0, 0.1, 0.0))])
def test_hpss_golden(fb, golden_dir, oracle, mode, ht, pt):
    g = np.load(os.path.join(golden_dir, "spectral.npz"))
    S = oracle.stft(g["audio"].astype(np.float64), 512, 512, 128).astype(np.complex64)
    with fb.Plan(win=512, hop=128, fft=512) as plan:
        o = plan.hpss(S, 31, 17, mode, ht, pt)
    ref = g[f"hpss{mode}"]
    if mode == 0:
        assert rel(o, ref) < 1e-5
    else:  # binary masks: a bin whose ratio sits within fp32 rounding of the threshold may flip; all others are exact copies
        differs = np.abs(o - ref) > 1e-5 * np.abs(ref).max()
        assert differs.mean() < 1e-3, differs.mean()


def test_hpss_sizes_and_batch(fb, oracle):
    rng = np.random.default_rng(3)
    S = (rng.standard_normal((2, 40, 129)) + 1j * rng.standard_normal((2, 40, 129))).astype(np.complex64)
    with fb.Plan(win=256, hop=64, fft=256) as plan:
        o = plan.hpss(S, 5, 9, 0)
        with pytest.raises(fb.FlucomaB200Error):
            plan.hpss(S, 4, 9, 0)  # MedianFilter::init asserts an odd size >= 3
    for b in range(2):
        assert rel(o[b], oracle.hpss(S[b].astype(np.complex128), 5, 9, 0)) < 1e-5
