"""GPU parity of NMFSeed / NNDSVD (fb200_nmfseed) against the fp64 oracle (NNDSVD.hpp:30-131, NMFSeedClient.hpp:74-133)."""
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


@pytest.mark.parametrize("method", [0, 1, 2, 3])
@pytest.mark.parametrize("win,hop,n", [(256, 64, 9000), (1024, 256, 40000)])
def test_nmfseed_vs_oracle(fb, oracle, method, win, hop, n):
    from tests.golden.make_golden import synth_audio
    a = synth_audio(9 + win, n)
    M = np.abs(oracle.stft(a.astype(np.float64), win, win, hop))
    ko, Wo, Ho, so = oracle.nndsvd(M, 3, 12, 0.9, method, 7)
    with fb.Plan(win=win, hop=hop, fft=win) as plan:
        k, W, H, s = plan.nmfseed(audio=a, min_rank=3, max_rank=12, coverage=0.9, method=method, seed=7)
        k2, W2, H2, _ = plan.nmfseed(mags=M.astype(np.float32), min_rank=3, max_rank=12, coverage=0.9, method=method, seed=7)
    assert k == ko == k2
    assert np.abs(s - so).max() <= 1e-5 * so.max()          # fp32 magnitudes in, fp64 Jacobi
    assert rel(W, Wo) < 1e-4 and rel(H, Ho) < 1e-4, (rel(W, Wo), rel(H, Ho))
    assert rel(W2, Wo) < 1e-4 and rel(H2, Ho) < 1e-4


def test_nmfseed_feeds_bufnmf(fb, oracle):
    """The client pair NMFSeed -> BufNMF (seeded bases and activations): ranks and scaling as NMFSeedClient.hpp:104-128."""
    from tests.golden.make_golden import synth_audio
    a = synth_audio(77, 20000)
    with fb.Plan(win=512, hop=128, fft=512) as plan:
        k, W, H, _ = plan.nmfseed(audio=a, min_rank=2, max_rank=8, coverage=0.7, scale_acts=True)
        assert 2 <= k <= 8 and abs(H.max() - 1.0) < 1e-6
        r = plan.bufnmf(a, k, 20, seeds=1, bases_mode=1, bases_in=np.ascontiguousarray(W[None, :k]), acts_mode=1,
                        acts_in=np.ascontiguousarray(H[None, :, :k]))
    o = oracle.bufnmf_channel(a, 512, 512, 128, k, 20, 1, bases_mode=1, bases_in=W[:k].astype(np.float32), acts_mode=1,
                              acts_in=np.ascontiguousarray(H[:, :k]).astype(np.float32))
    assert rel(r["bases"][0], o["bases"]) < 1e-4 and rel(r["acts"][0], o["acts"]) < 1e-4
