"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libflucoma_b200.so), against the fp64 oracle.

Tolerance (BASELINE.json north_star): 1e-4 relative on W, H and the resynthesis; we measure it as the Frobenius-relative
error per buffer, ||gpu - oracle|| / ||oracle||, because individual near-zero entries of W/H carry no relative meaning.
The GPU computes in fp32 (the reference in fp64), so the STFT stage is held to 2e-6 and the long iterations to 1e-4.
Run on the B200 box:  python -m pytest tests -m gpu -x -q
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-4


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    ct = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a = a.astype(ct); b = b.astype(ct)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def fb():
    import flucoma_b200
    flucoma_b200.load()
    assert flucoma_b200.device_count() >= 1, "no CUDA device: the product path has no CPU fallback"
    return flucoma_b200


@pytest.fixture(scope="module")
def synth():
    from tests.golden.make_golden import synth_audio
    return synth_audio


# ------------------------------------------------------------------------------------------------ STFT / ISTFT
@pytest.mark.parametrize("n,win,fft,hop", [(3000, 256, 256, 64), (3000, 200, 256, 50), (1000, 128, 512, 128),
                                           (5, 64, 64, 16), (130816, 1024, 1024, 256)])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_stft_parity(fb, oracle, n, win, fft, hop, dt):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((3, n)).astype(dt)
    with fb.Plan(win=win, hop=hop, fft=fft) as plan:
        spec, mag = plan.stft(x, want_spectrum=True, want_magnitude=True)
    F = oracle.num_frames(n, win, hop)
    assert spec.shape == (3, F, fft // 2 + 1) and mag.shape == spec.shape
    for b in range(3):
        S = oracle.stft(x[b].astype(np.float64), win, fft, hop)
        assert rel(spec[b], S) < 2e-6
        assert rel(mag[b], np.abs(S)) < 2e-6
    assert np.all(spec[..., 0].imag == 0) and np.all(spec[..., -1].imag == 0)  # FFT.hpp:99-101


def test_stft_impulse_and_sine_kat(fb, oracle):
    n, win = 4096, 1024
    with fb.Plan(win=win, hop=256, fft=win) as plan:
        x = np.zeros(n, np.float32); x[0] = 1.0
        spec, _ = plan.stft(x)
        assert np.allclose(np.abs(spec[0, 0]), oracle.hann(win)[win // 2], atol=1e-6)  # first frame centred on sample 0
        k = 37
        s = np.sin(2 * np.pi * k * np.arange(n) / win).astype(np.float32)
        _, mag = plan.stft(s, want_spectrum=False, want_magnitude=True)
        f = 6  # a frame fully inside the signal
        assert abs(mag[0, f, k] - 0.25 * win) < 1e-2 and abs(mag[0, f, k + 1] - 0.125 * win) < 1e-2


@pytest.mark.parametrize("win,hop", [(256, 64), (1024, 256), (1024, 512), (64, 32)])
def test_stft_istft_identity(fb, win, hop):
    n = 20000
    x = np.random.default_rng(1).standard_normal((2, n)).astype(np.float32)
    with fb.Plan(win=win, hop=hop, fft=win) as plan:
        spec, _ = plan.stft(x)
        y = plan.istft(spec, n)
    assert np.abs(y - x).max() < 5e-6 * np.abs(x).max() * np.sqrt(win)


def test_istft_parity(fb, oracle):
    rng = np.random.default_rng(5)
    F, win, hop = 40, 256, 64
    S = (rng.standard_normal((2, F, 129)) + 1j * rng.standard_normal((2, F, 129)))
    n = (F - 1) * hop
    with fb.Plan(win=win, hop=hop, fft=win) as plan:
        y64 = plan.istft(S, n)
        y32 = plan.istft(S.astype(np.complex64), n)
    for b in range(2):
        ref = oracle.istft(S[b], win, win, hop, n)
        assert rel(y64[b], ref) < 2e-6 and rel(y32[b], ref) < 2e-6


# ------------------------------------------------------------------------------------------------ NMF::process
def test_nmf_reference_test_replica(fb, oracle):
    # tests/algorithms/public/TestNMF.cpp:11-46 through the C ABI
    X = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]], dtype=np.float64)
    with fb.Plan(win=64) as plan:
        r = [plan.nmf_process(X, 2, 1, True, True, seeds=s) for s in (42, 42, 5063, 5063)]
    for i, j in ((0, 1), (2, 3)):
        for a, b in zip(r[i][:3], r[j][:3]):
            assert np.array_equal(a, b)
    for a, b in zip(r[1][:3], r[2][:3]):
        assert not np.array_equal(a, b)
    for res, seed in ((r[0], 42), (r[2], 5063)):
        W, H, V, _ = oracle.nmf_process(X, 2, 1, True, True, seed)
        assert rel(res[0], W) < 1e-5 and rel(res[1], H) < 1e-5 and rel(res[2], V) < 1e-5
    assert r[0][0].shape == (2, 3) and r[0][1].shape == (3, 2) and r[0][2].shape == (3, 3)


@pytest.mark.parametrize("tag,uw,uh", [("wh", True, True), ("w", True, False), ("h", False, True)])
def test_nmf_golden_small(fb, golden_dir, tag, uw, uh):
    g = np.load(os.path.join(golden_dir, "nmf_small.npz"))
    with fb.Plan(win=256, hop=64) as plan:
        W, H, V, st = plan.nmf_process(g["mag"], 5, 40, uw, uh, seeds=3)
        W32, H32, V32, _ = plan.nmf_process(g["mag"].astype(np.float32), 5, 40, uw, uh, seeds=3)
    assert st == 0
    assert rel(W, g[f"W_{tag}"]) < TOL and rel(H, g[f"H_{tag}"]) < TOL and rel(V, g[f"V_{tag}"]) < TOL
    assert rel(W32, g[f"W_{tag}"]) < TOL and rel(H32, g[f"H_{tag}"]) < TOL


def test_nmf_seeded_w0_h0(fb, golden_dir):
    g = np.load(os.path.join(golden_dir, "nmf_small.npz"))
    with fb.Plan(win=256, hop=64) as plan:
        W, H, V, _ = plan.nmf_process(g["mag"], 5, 25, True, True, W0=g["W0"], H0=g["H0"])
    assert rel(W, g["W_seeded"]) < TOL and rel(H, g["H_seeded"]) < TOL and rel(V, g["V_seeded"]) < TOL


@pytest.mark.parametrize("F,B,K,iters", [(7, 5, 1, 3), (130, 33, 3, 20), (300, 257, 8, 30), (129, 513, 16, 30),
                                         (64, 2049, 20, 10), (256, 129, 64, 10)])
def test_nmf_shapes_vs_oracle(fb, oracle, F, B, K, iters):
    rng = np.random.default_rng(F * 1000 + B)
    kk = max(2, min(K, 6))
    X = (rng.random((2, F, kk)) ** 3) @ (rng.random((2, kk, B)) ** 3) + 1e-3 * rng.random((2, F, B))
    with fb.Plan(win=64) as plan:
        W, H, V, _ = plan.nmf_process(X, K, iters, True, True, seeds=[11, 12])
    for b in range(2):
        Wo, Ho, Vo, _ = oracle.nmf_process(X[b], K, iters, True, True, 11 + b)
        assert rel(W[b], Wo) < TOL and rel(H[b], Ho) < TOL and rel(V[b], Vo) < TOL


def test_nmf_zero_iterations_and_no_updates(fb, oracle):
    rng = np.random.default_rng(0)
    X = rng.random((6, 5)); W0 = rng.random((2, 5)); H0 = rng.random((6, 2)); H0[0, 0] = 0.0
    with fb.Plan(win=64) as plan:
        W, H, V, _ = plan.nmf_process(X, 2, 0, True, True, W0=W0, H0=H0)
        W2, H2, V2, _ = plan.nmf_process(X, 2, 5, False, False, W0=W0, H0=H0)
    Wo, Ho, Vo, _ = oracle.nmf_process(X, 2, 0, True, True, -1, W0=W0, H0=H0)
    for a, b in ((W, Wo), (H, Ho), (V, Vo), (W2, Wo), (H2, Ho), (V2, Vo)):
        assert rel(a, b) < 1e-6
    assert H[0, 0] > 0  # eps clamp (NMF.hpp:150)


def test_nmf_progress_and_cancel(fb, oracle):
    X = np.random.default_rng(0).random((40, 30))
    seen = []
    with fb.Plan(win=64) as plan:
        W, H, V, st = plan.nmf_process(X, 3, 10, True, True, seeds=1, progress=lambda it: seen.append(it) or it < 4)
        Wf, Hf, Vf, stf = plan.nmf_process(X, 3, 10, True, True, seeds=1, progress=lambda it: True)
        Wn, Hn, Vn, _ = plan.nmf_process(X, 3, 10, True, True, seeds=1)
    assert st == fb.CANCELLED and seen == [1, 2, 3, 4]
    assert np.array_equal(V, X)  # NMF.hpp:175-176 skips :182
    Wo, Ho, _, _ = oracle.nmf_process(X, 3, 4, True, True, 1)
    assert rel(W, Wo) < TOL and rel(H, Ho) < TOL
    assert stf == 0
    # the fused schedule (no callback) and the per-iteration schedule agree
    assert rel(Wf, Wn) < 1e-5 and rel(Hf, Hn) < 1e-5 and rel(Vf, Vn) < 1e-5


def test_nmf_device_arrays_and_repeatability(fb, oracle):
    import torch
    rng = np.random.default_rng(9)
    X = rng.random((4, 200, 65)).astype(np.float32)
    Xd = torch.from_numpy(X).cuda()
    with fb.Plan(win=128) as plan:
        Wd, Hd, Vd, _ = plan.nmf_process(Xd, 4, 50, True, True, seeds=[1, 2, 3, 4])
        Wh, Hh, Vh, _ = plan.nmf_process(X, 4, 50, True, True, seeds=[1, 2, 3, 4])
        Wd2, Hd2, _, _ = plan.nmf_process(Xd, 4, 50, True, True, seeds=[1, 2, 3, 4])
    assert Wd.is_cuda and np.array_equal(Wd.cpu().numpy(), Wh) and np.array_equal(Hd.cpu().numpy(), Hh)
    assert torch.equal(Wd, Wd2) and torch.equal(Hd, Hd2)  # bitwise repeatable (TestNMF.cpp:31-39)
    Wo, Ho, _, _ = oracle.nmf_process(X[2].astype(np.float64), 4, 50, True, True, 3)
    assert rel(Wh[2], Wo) < TOL and rel(Hh[2], Ho) < TOL


# ------------------------------------------------------------------------------------------------ NMF::processFrame
def test_process_frames_parity(fb, oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "nmf_small.npz"))
    mags, W = g["mag"], g["W_wh"]
    with fb.Plan(win=256, hop=64) as plan:
        H, V, Wn = plan.nmf_process_frames(mags, W, 10, 42, want_v=True, want_w=True)
        H0, _, _ = plan.nmf_process_frames(mags, W, 0, 42)
    acts = oracle.nmfmatch_frames(mags, W, 10, 42)
    assert rel(H, acts) < TOL
    h, v, Wm = oracle.nmf_process_frame(mags[20], W, 10, 42)
    assert rel(H[20], h) < TOL and rel(V[20], v) < TOL and rel(Wn, Wm) < 1e-6
    assert rel(H[20], g["pf_h"]) < TOL
    assert np.allclose(H0, np.maximum(oracle.random_uniform(42, 5), 2.2e-16)[None, :], rtol=1e-6)  # TestNMF.cpp:48-73


def test_process_frame_reference_test_replica(fb):
    x = np.array([[1, 0, 1, 0.0]]); W = np.array([[0, 0, 1, 0], [1, 0, 0, 0.0]])
    with fb.Plan(win=64) as plan:
        a, _, _ = plan.nmf_process_frames(x, W, 0, 42)
        b, _, _ = plan.nmf_process_frames(x, W, 0, 42)
        c, _, _ = plan.nmf_process_frames(x, W, 0, 7863)
    assert np.array_equal(a, b) and not np.array_equal(b, c)


# ------------------------------------------------------------------------------------------------ BufNMF
# ------------------------------------------------------------------------------------------------ streaming clients
def test_nmf_filter_stream_golden(fb, oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "nmffilter_stream.npz"))
    a, W = g["audio"], g["bases"]
    with fb.Plan(win=256, hop=64) as plan:
        out, acts = plan.nmf_filter(a, W, 10, 42)
        _, acts_only = plan.nmf_filter(a, W, 10, 42, want_out=False)          # NMFMatch on the same stream
        import torch
        out_d, acts_d = plan.nmf_filter(torch.from_numpy(a).cuda(), torch.from_numpy(W).cuda(), 10, 42)
    assert out.shape == g["out"].shape and acts.shape == g["acts"].shape
    assert rel(out, g["out"]) < TOL and rel(acts, g["acts"]) < TOL
    assert np.array_equal(acts, acts_only)
    assert np.array_equal(out_d.cpu().numpy(), out) and np.array_equal(acts_d.cpu().numpy(), acts)
    assert rel(out.sum(0)[256:], a[:-256]) < 1e-5 and np.abs(out[:, :64]).max() == 0.0


@pytest.mark.parametrize("n,win,fft,hop,K", [(5000, 200, 256, 50, 3), (3000, 128, 128, 128, 2), (777, 64, 64, 16, 4),
                                             (4000, 256, 512, 300, 2)])
def test_nmf_filter_stream_shapes(fb, oracle, synth, n, win, fft, hop, K):
    a = synth(77, n)
    rng = np.random.default_rng(3)
    W = (rng.random((K, fft // 2 + 1)) ** 2).astype(np.float32)
    ref_out, ref_acts = oracle.nmffilter_stream(a.astype(np.float64), win, fft, hop, W.astype(np.float64), 7, 5)
    with fb.Plan(win=win, hop=hop, fft=fft) as plan:
        out, acts = plan.nmf_filter(a, W, 7, 5)
    assert rel(out, ref_out) < TOL and rel(acts, ref_acts) < TOL


def test_nmf_filter_stream_multi_chunk(fb, oracle, synth):
    """A stream longer than one chunk of frames (8192 at rank 16, fft 1024): chunk seams must be invisible."""
    hop, win, K = 256, 1024, 16
    n = 256 * 20000 + 17
    a = np.tile(synth(11, 1 << 18), n // (1 << 18) + 1)[:n].copy()
    a *= np.linspace(0.2, 1.0, n, dtype=np.float32)                            # no two chunks see the same frames
    rng = np.random.default_rng(7)
    W = (rng.random((K, 513)) ** 3).astype(np.float32)
    with fb.Plan(win=win, hop=hop) as plan:
        out, acts = plan.nmf_filter(a, W, 10, 42)
    assert acts.shape == ((n + hop - 1) // hop, K)
    assert rel(out.sum(0)[win:], a[:-win]) < 1e-5
    for f0 in (0, 8192 - 3 - 40, 2 * (8192 - 3) - 40, 20000 - 60):             # start, both seams, tail
        lo = max(0, f0 * hop - win); hi = min(n, (f0 + 80) * hop)
        seg = a[lo:hi].astype(np.float64)
        r_out, r_acts = oracle.nmffilter_stream(seg, win, 1024, hop, W.astype(np.float64), 10, 42)
        skip = win // hop + 4                                                   # frames that saw zeros instead of history
        t0 = lo + skip * hop + win
        assert rel(out[:, t0:hi], r_out[:, t0 - lo:]) < TOL
        assert rel(acts[lo // hop + skip:(hi + hop - 1) // hop], r_acts[skip:]) < TOL


def test_config5_million_frames(fb, oracle, synth):
    """BASELINE config 5: 10^6 streaming frames against fixed rank-16 bases, 10 iterations, h0 from seed 42."""
    import torch
    win, hop, K, frames = 1024, 512, 16, 1_000_000
    n = frames * hop
    base = torch.from_numpy(synth(21, 1 << 20)).cuda()
    a = base.repeat(n // base.numel() + 1)[:n].contiguous()
    a *= torch.linspace(0.1, 1.0, n, device="cuda")
    rng = np.random.default_rng(7)
    W = (rng.random((K, 513)) ** 3).astype(np.float32)
    with fb.Plan(win=win, hop=hop) as plan:
        _, acts = plan.nmf_filter(a, torch.from_numpy(W).cuda(), 10, 42, want_out=False)
        ms = plan.stats()["ms_total"]
    acts = acts.cpu().numpy()
    assert acts.shape == (frames, K) and np.all(np.isfinite(acts)) and acts.min() >= 0
    print(f"config 5: {frames / ms * 1e3:.3e} frames/s")
    for f0 in (0, 8190, 131_070, 500_000, frames - 50):                         # includes a chunk seam (131072 frames per chunk)
        lo = max(0, f0 * hop - win); hi = min(n, (f0 + 50) * hop)
        seg = a[lo:hi].cpu().numpy().astype(np.float64)
        _, r = oracle.nmffilter_stream(seg, win, 1024, hop, W.astype(np.float64), 10, 42, want_out=False)
        skip = win // hop + 1
        assert rel(acts[lo // hop + skip:(hi + hop - 1) // hop], r[skip:]) < TOL


def test_bufnmf_golden_wav(fb, golden_dir):
    g = np.load(os.path.join(golden_dir, "bufnmf_wav.npz"))
    with fb.Plan(win=1024, hop=256, fft=1024) as plan:
        r = plan.bufnmf(g["audio"], 4, 100, seeds=42, resynth=True)
        st = plan.stats()
    assert r["status"] == 0
    assert rel(r["bases"][0], g["bases"]) < TOL and rel(r["acts"][0], g["acts"]) < TOL
    assert rel(r["resynth"][0], g["resynth"]) < TOL
    assert r["acts"].max() == pytest.approx(1.0, abs=1e-6)
    assert np.abs(r["resynth"][0].sum(0) - g["audio"]).max() < 1e-4  # masks sum to one
    assert st["launches_nmf"] > 0 and st["launches_total"] > st["launches_nmf"]


def test_bufnmf_batch_vs_oracle(fb, oracle, synth):
    a = np.stack([synth(1000 + b, 9000) for b in range(5)])
    with fb.Plan(win=512, hop=128, fft=512) as plan:
        r = plan.bufnmf(a, 6, 60, seeds=[0, 1, 2, 3, 4], resynth=True)
    for b in range(5):
        o = oracle.bufnmf_channel(a[b], 512, 512, 128, 6, 60, b, resynth=True)
        assert rel(r["bases"][b], o["bases"]) < TOL and rel(r["acts"][b], o["acts"]) < TOL
        assert rel(r["resynth"][b], o["resynth"]) < TOL


def test_bufnmf_seed_and_fixed_modes(fb, oracle, synth):
    a = synth(1003, 8000)[None, :]
    win, hop, K = 256, 64, 3
    base = oracle.bufnmf_channel(a[0], win, win, hop, K, 30, 7)
    with fb.Plan(win=win, hop=hop, fft=win) as plan:
        r_fix = plan.bufnmf(a, K, 30, seeds=7, bases_mode=2, bases_in=base["bases"][None])
        r_seed = plan.bufnmf(a, K, 30, seeds=7, bases_mode=1, bases_in=base["bases"][None], acts_mode=1,
                             acts_in=base["acts"][None], resynth=True)
        with pytest.raises(fb.FlucomaB200Error):
            plan.bufnmf(a, K, 30, seeds=7, bases_mode=1)
        r_none = plan.bufnmf(a, K, 30, seeds=7, bases_mode=2, bases_in=base["bases"][None], acts_mode=2,
                             acts_in=base["acts"][None])
    assert r_fix["bases"] is None
    o_fix = oracle.bufnmf_channel(a[0], win, win, hop, K, 30, 7, bases_mode=2, bases_in=base["bases"])
    assert rel(r_fix["acts"][0], o_fix["acts"]) < TOL
    o_seed = oracle.bufnmf_channel(a[0], win, win, hop, K, 30, 7, bases_mode=1, bases_in=base["bases"], acts_mode=1,
                                   acts_in=base["acts"], resynth=True)
    assert rel(r_seed["bases"][0], o_seed["bases"]) < TOL and rel(r_seed["acts"][0], o_seed["acts"]) < TOL
    assert rel(r_seed["resynth"][0], o_seed["resynth"]) < TOL
    assert r_none["status"] == fb.WARN_NO_WORK  # NMFClient.hpp:143-145


def test_bufnmf_device_memory_matches_host(fb, synth):
    import torch
    a = np.stack([synth(2000 + b, 6000) for b in range(3)])
    with fb.Plan(win=256, hop=64, fft=256) as plan:
        rh = plan.bufnmf(a, 4, 20, seeds=[5, 6, 7], resynth=True)
        rd = plan.bufnmf(torch.from_numpy(a).cuda(), 4, 20, seeds=[5, 6, 7], resynth=True)
    for k in ("bases", "acts", "resynth"):
        assert np.array_equal(rd[k].cpu().numpy(), rh[k])


# ------------------------------------------------------------------------------------------------ full size (config 2)
def test_config2_full_size_properties(fb, oracle, synth):
    """BASELINE config 2 shape (fft 1024, hop 256, F=512, K=16, 200 iters) at batch 128 on device memory:
    size-independent properties + two buffers spot-checked against the oracle."""
    import torch
    batch, n, K, iters = 128, 130816, 16, 200
    a = np.stack([synth(1000 + b, n) for b in range(8)])
    a = np.concatenate([a] * (batch // 8))  # 8 distinct buffers repeated; NMF seeds differ per buffer
    ad = torch.from_numpy(a).cuda()
    with fb.Plan(win=1024, hop=256, fft=1024) as plan:
        r = plan.bufnmf(ad, K, iters, seeds=np.arange(batch))
        r2 = plan.bufnmf(ad[:4], K, iters, seeds=np.arange(4), resynth=True)
    bases = r["bases"].cpu().numpy(); acts = r["acts"].cpu().numpy()
    assert bases.shape == (batch, K, 513) and acts.shape == (batch, 512, K)
    assert np.isfinite(bases).all() and np.isfinite(acts).all() and (bases >= 0).all() and (acts >= 0).all()
    assert np.allclose(np.linalg.norm(bases.astype(np.float64), axis=2), 1.0, atol=1e-4)   # NMF.hpp:162
    assert np.allclose(acts.reshape(batch, -1).max(1), 1.0, atol=1e-6)                    # NMFClient.hpp:289-298
    # same audio + same seed => same factorisation regardless of the position in the batch
    assert np.array_equal(r2["bases"].cpu().numpy(), bases[:4])
    rs = r2["resynth"].cpu().numpy()
    assert np.abs(rs.sum(1) - a[:4]).max() < 2e-4                                         # ratio masks sum to one
    for b in (0, 3):
        o = oracle.bufnmf_channel(a[b], 1024, 1024, 256, K, iters, b, resynth=True)
        assert rel(bases[b], o["bases"]) < TOL and rel(acts[b], o["acts"]) < TOL
        assert rel(rs[b], o["resynth"]) < TOL


# ------------------------------------------------------------------------- configs 3 and 4 (full shapes, reduced batch)
@pytest.mark.parametrize("name,win,hop,K,iters", [("config3", 1024, 256, 32, 40), ("config4", 4096, 1024, 64, 12)])
def test_config3_config4_shapes(fb, oracle, synth, name, win, hop, K, iters):
    """BASELINE configs 3 (rank 32) and 4 (fft 4096, rank 64) at their real frame/bin/rank shapes (F = 512), a few buffers
    and a reduced iteration count so that the fp64 CPU check stays in seconds; plus the size-independent properties."""
    import torch
    F = 512
    n = (F - 1) * hop
    batch = 3
    a = np.stack([synth(3000 + b, n) for b in range(batch)])
    with fb.Plan(win=win, hop=hop, fft=win) as plan:
        r = plan.bufnmf(torch.from_numpy(a).cuda(), K, iters, seeds=np.arange(batch))
        st = plan.stats()
    bases = r["bases"].cpu().numpy(); acts = r["acts"].cpu().numpy()
    assert bases.shape == (batch, K, win // 2 + 1) and acts.shape == (batch, F, K)
    assert np.isfinite(bases).all() and np.isfinite(acts).all() and (bases >= 0).all() and (acts >= 0).all()
    assert np.allclose(np.linalg.norm(bases.astype(np.float64), axis=2), 1.0, atol=1e-4)
    assert np.allclose(acts.reshape(batch, -1).max(1), 1.0, atol=1e-6)
    o = oracle.bufnmf_channel(a[1], win, win, hop, K, iters, 1)
    assert rel(bases[1], o["bases"]) < TOL and rel(acts[1], o["acts"]) < TOL
    print(f"{name}: {batch * F / st['ms_nmf'] * 1e3 * iters:.3e} frame-iterations/s on {batch} buffers")


# ------------------------------------------------------------------------------------------------ BufSTFT
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_bufstft_golden_and_parity(fb, oracle, golden_dir, mode):
    """BufSTFTClient processFwd / processInverse through fb200_bufstft against the committed oracle vectors; phases of
    noise-floor bins are ill-conditioned, so the forward result is compared as the complex spectrum mag*exp(i*phase)."""
    g = np.load(os.path.join(golden_dir, "bufstft.npz"))
    a = g["audio"]
    with fb.Plan(win=200, hop=50, fft=256) as plan:
        m, p = plan.bufstft(a, padding_mode=mode)
        r = plan.bufstft_inverse(g[f"mag{mode}"], g[f"phase{mode}"], padding_mode=mode)
        m_only, none = plan.bufstft(a, padding_mode=mode, want_phase=False)
    assert m.shape == g[f"mag{mode}"].shape and none is None and np.array_equal(m_only, m)
    z = m.astype(np.float64) * np.exp(1j * p.astype(np.float64))
    zg = g[f"mag{mode}"].astype(np.float64) * np.exp(1j * g[f"phase{mode}"].astype(np.float64))
    assert rel(m, g[f"mag{mode}"]) < 2e-6 and rel(z, zg) < 2e-6
    # The window^2 normaliser tends to zero at an unpadded edge (Hann), where acc/nrm amplifies the fp32 error of the
    # irFFT by 1/w: the first/last `win` samples are held to 1e-4, everything else to the STFT-stage bar of 2e-6.
    gi = g[f"inv{mode}"]
    assert rel(r[200:-200], gi[200:-200]) < 2e-6 and rel(r, gi) < 1e-4
    assert np.all(np.abs(p) <= np.float32(np.pi) + 1e-6)


@pytest.mark.parametrize("n,win,fft,hop,mode", [(3000, 256, 256, 64, 1), (5000, 1024, 1024, 256, 2), (777, 128, 512, 128, 0),
                                               (64, 64, 64, 16, 0)])
def test_bufstft_batch_device_roundtrip(fb, oracle, synth, n, win, fft, hop, mode):
    import torch
    a = np.stack([synth(500 + b, n) for b in range(3)])
    with fb.Plan(win=win, hop=hop, fft=fft) as plan:
        md, pd = plan.bufstft(torch.from_numpy(a).cuda(), padding_mode=mode)
        rd = plan.bufstft_inverse(md, pd, padding_mode=mode)
        mh, ph = plan.bufstft(a, padding_mode=mode)
    assert np.array_equal(md.cpu().numpy(), mh) and np.array_equal(pd.cpu().numpy(), ph)   # device and host paths agree
    rd = rd.cpu().numpy()
    for b in range(3):
        mo, po = oracle.bufstft_fwd(a[b], win, fft, hop, mode)
        assert rel(mh[b], mo) < 2e-6
        ro = oracle.bufstft_inv(mo, po, win, fft, hop, mode)
        assert rel(rd[b], ro) < 1e-4
        if ro.size > 3 * win and 2 * hop <= win:   # without overlap every frame edge divides by a vanishing window^2
            assert rel(rd[b][win:-win], ro[win:-win]) < 5e-6
    with fb.Plan(win=win, hop=hop, fft=fft) as plan:
        with pytest.raises(fb.FlucomaB200Error):
            plan.bufstft(a[:, :max(1, win // 4)], padding_mode=0)                            # shorter than one window


# ---------------------------------------------------------------------------------------------- round 2 additions
@pytest.mark.parametrize("name,frames", [("tremblay", 2013), ("nicol", 1774)])
def test_config1_full_length_wav(fb, oracle, golden_dir, name, frames):
    """BASELINE config 1 on the two reference WAVs at FULL length (F = 2013 / 1774), resynthesis included."""
    g = np.load(os.path.join(golden_dir, f"config1_{name}.npz"))
    a = (g["pcm"].astype(np.float64) / float(1 << (int(g["bits"]) - 1))).astype(np.float32)
    with fb.Plan(win=1024, hop=256, fft=1024) as plan:
        r = plan.bufnmf(a, 4, 100, seeds=42, resynth=True)
    assert r["acts"].shape == (1, frames, 4)
    assert rel(r["bases"][0], g["bases"]) < TOL and rel(r["acts"][0], g["acts"]) < TOL
    o = oracle.bufnmf_channel(a, 1024, 1024, 256, 4, 100, 42, resynth=True)
    for k in ("bases", "acts", "resynth"):
        assert rel(r[k][0], o[k]) < TOL, (k, rel(r[k][0], o[k]))
    assert np.abs(r["resynth"][0].sum(axis=0) - a).max() < 1e-4  # the masks sum to one: the components add up to the input


@pytest.mark.parametrize("backend,K", [("BACKEND_AUTO", 5), ("BACKEND_TCGEN05", 16), ("BACKEND_TCGEN05_STREAMED", 32),
                                       ("BACKEND_TCGEN05_STREAMED", 64)])
def test_kl_divergence_tracks_the_oracle_per_iteration(fb, oracle, backend, K):
    """A final-state Frobenius check can hide drift along flat directions; the objective cannot.  After j iterations the
    KL divergence of the GPU factors must equal the oracle's to 1e-5 relative (fp32 factors), and never increase."""
    from tests.test_oracle import kl_divergence
    rng = np.random.default_rng(12)
    V = (rng.random((256, 7)) ** 3) @ (rng.random((7, 257)) ** 3) + 1e-3 * rng.random((256, 257))
    prev = None
    with fb.Plan(win=64, backend=getattr(fb, backend)) as plan:
        for it in (1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144):
            W, H, _, _ = plan.nmf_process(V, K, it, True, True, seeds=4, want_v=False)
            Wo, Ho, _, _ = oracle.nmf_process(V, K, it, True, True, 4)
            d, do = kl_divergence(V, W.astype(np.float64), H.astype(np.float64)), kl_divergence(V, Wo, Ho)
            assert abs(d - do) <= 1e-5 * abs(do), (it, d, do)
            if prev is not None:
                assert d <= prev * (1 + 1e-6), (it, d, prev)
            prev = d


@pytest.mark.parametrize("backend", ["BACKEND_SIMT", "BACKEND_TCGEN05_STREAMED"])
def test_config4_rank64_500_iterations(fb, oracle, synth, backend):
    """BASELINE config 4 (fft 4096, hop 1024, rank 64) at its FULL iteration count, on 128 frames so that the fp64 oracle
    finishes in seconds: W, H within 1e-4 after 500 iterations, on the SIMT engine and on the streamed tcgen05 engine (the
    one a full-size config-4 batch runs on)."""
    a = np.stack([synth(3000 + b, 127 * 1024) for b in range(2)])
    with fb.Plan(win=4096, hop=1024, fft=4096, backend=getattr(fb, backend)) as plan:
        r = plan.bufnmf(a, 64, 500, seeds=[0, 1])
        assert plan.stats()["backend_used"] == getattr(fb, backend)
    bases, acts, _ = oracle.bufnmf_batch(a, 4096, 4096, 1024, 64, 500, np.arange(2), faithful=False)
    for b in range(2):
        assert rel(r["bases"][b], bases[b]) < TOL and rel(r["acts"][b], acts[b]) < TOL, (b, rel(r["bases"][b], bases[b]))


@pytest.mark.parametrize("n,win,fft,hop", [(8192, 256, 256, 64), (6000, 400, 512, 100), (130816, 1024, 1024, 256),
                                           (40000, 2048, 2048, 512), (70000, 4096, 4096, 1024), (12000, 1000, 4096, 250),
                                           (4096, 512, 512, 512)])
def test_fused_stft_kernel_vs_oracle_and_cufft_path(fb, oracle, n, win, fft, hop):
    """The fused TMA + shared-memory FFT kernel (kernels_stft_fused.cu) against the fp64 oracle, and against the cuFFT
    pipeline it replaces (FB200_STFT_CUFFT=1 forces the fallback)."""
    rng = np.random.default_rng(n + fft)
    x = rng.standard_normal((3, n)).astype(np.float32)
    with fb.Plan(win=win, hop=hop, fft=fft) as plan:
        spec, mag = plan.stft(x, want_spectrum=True, want_magnitude=True)
        launches_fused = plan.stats()["launches_total"]
        os.environ["FB200_STFT_CUFFT"] = "1"
        try:
            spec_c, mag_c = plan.stft(x, want_spectrum=True, want_magnitude=True)
            launches_cufft = plan.stats()["launches_total"]
        finally:
            del os.environ["FB200_STFT_CUFFT"]
    assert launches_fused < launches_cufft  # one kernel instead of three per wave
    for b in range(3):
        S = oracle.stft(x[b].astype(np.float64), win, fft, hop)
        assert rel(spec[b], S) < 2e-6 and rel(mag[b], np.abs(S)) < 2e-6, (rel(spec[b], S), rel(mag[b], np.abs(S)))
        assert rel(spec[b], spec_c[b]) < 2e-6
    assert np.all(spec[..., 0].imag == 0) and np.all(spec[..., -1].imag == 0)  # FFT.hpp:99-101


@pytest.mark.parametrize("n,win,fft,hop", [(8192, 256, 256, 64), (6001, 400, 512, 100), (130816, 1024, 1024, 256),
                                           (40000, 2048, 2048, 512), (70000, 4096, 4096, 1024), (12000, 1000, 4096, 250),
                                           (4096, 512, 512, 512), (300, 256, 256, 128)])
def test_fused_istft_kernel_vs_oracle_and_cufft_path(fb, oracle, n, win, fft, hop):
    """The fused inverse (spectrum rows -> shared-memory inverse FFT -> window -> overlap-add -> normalise, one kernel:
    kernels_stft_fused.cu) against the fp64 oracle's ISTFT (STFT.hpp:178-199) and against the cuFFT C2R + k_ola pipeline it
    replaces (FB200_ISTFT_CUFFT=1 forces the fallback); arbitrary (non-analysis) spectra, so the overlap-add really sums."""
    rng = np.random.default_rng(n + fft)
    F = n // hop + 1
    B = fft // 2 + 1
    S = (rng.standard_normal((3, F, B)) + 1j * rng.standard_normal((3, F, B))).astype(np.complex64)
    S[..., 0] = S[..., 0].real
    S[..., -1] = S[..., -1].real
    with fb.Plan(win=win, hop=hop, fft=fft) as plan:
        y = plan.istft(S, n)  # the first call also builds the plan's twiddle table
        y2 = plan.istft(S, n)
        launches_fused = plan.stats()["launches_total"]
        os.environ["FB200_ISTFT_CUFFT"] = "1"
        try:
            y_c = plan.istft(S, n)
            launches_cufft = plan.stats()["launches_total"]
        finally:
            del os.environ["FB200_ISTFT_CUFFT"]
    assert launches_fused < launches_cufft  # one kernel instead of cuFFT C2R + overlap-add
    assert np.array_equal(y, y2)  # fixed summation order
    for b in range(3):
        ref = oracle.istft(S[b].astype(np.complex128), win, fft, hop, n)
        assert rel(y[b], ref) < 3e-6, rel(y[b], ref)
        assert rel(y[b], y_c[b]) < 3e-6


def test_bufnmf_resynthesis_fused_mask_and_inverse_vs_cufft_pipeline(fb, synth):
    """BufNMF resynthesis applies the ratio masks while the fused inverse kernel loads its spectrum rows (the masked spectra
    are never written); FB200_ISTFT_CUFFT=1 runs mask kernel -> cuFFT C2R -> overlap-add instead.  Same masks bit for bit,
    inverse transforms within fp32 rounding; the components add up to the input either way (the masks sum to one)."""
    a = np.stack([synth(400 + b, 20000) for b in range(5)])
    with fb.Plan(win=512, hop=128, fft=512) as plan:
        r = plan.bufnmf(a, 6, 15, seeds=np.arange(5), resynth=True)
        os.environ["FB200_ISTFT_CUFFT"] = "1"
        try:
            rc = plan.bufnmf(a, 6, 15, seeds=np.arange(5), resynth=True)
        finally:
            del os.environ["FB200_ISTFT_CUFFT"]
    assert np.array_equal(r["bases"], rc["bases"]) and np.array_equal(r["acts"], rc["acts"])
    assert rel(r["resynth"], rc["resynth"]) < 3e-6
    assert np.abs(r["resynth"].sum(axis=1) - a).max() < 1e-4
