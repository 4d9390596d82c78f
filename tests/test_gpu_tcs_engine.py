"""GPU tests of the STREAMED tcgen05 NMF engine (kernels_nmf_tcs.cu), forced through FB200_BACKEND_TCGEN05_STREAMED:
rank 16, 32 and 64 (and ranks padded up to them), bins = 128 m + 1 up to fft 4096, frame counts beyond 512, fixed-dictionary
frame streams (NMF::processFrame over many frames).  Parity against the fp64 oracle at the north_star bar (1e-4
Frobenius-relative), agreement with the resident engine and the SIMT engine, repeatability."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-4


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def fb():
    import flucoma_b200
    return flucoma_b200


def lowrank(rng, batch, F, B, kk=6):
    return (rng.random((batch, F, kk)) ** 3) @ (rng.random((batch, kk, B)) ** 3) + 1e-3 * rng.random((batch, F, B))


@pytest.mark.parametrize("F,B,K,iters,uw,uh", [
    (128, 129, 16, 5, True, True), (200, 257, 16, 20, True, True), (512, 513, 16, 30, True, True),
    (384, 513, 16, 12, True, False), (256, 513, 16, 12, False, True), (130, 513, 16, 1, True, True),
    (128, 129, 32, 5, True, True), (300, 257, 32, 20, True, True), (512, 513, 32, 30, True, True),
    (640, 1025, 32, 10, True, True), (900, 513, 16, 10, True, True), (256, 2049, 32, 6, True, True),
    (384, 513, 32, 12, True, False), (256, 513, 32, 12, False, True), (200, 257, 20, 10, True, True), (200, 257, 9, 10, True, True),
    (128, 129, 64, 5, True, True), (300, 513, 64, 20, True, True), (256, 2049, 64, 6, True, True), (384, 513, 64, 12, True, False),
    (256, 513, 64, 12, False, True), (130, 257, 64, 1, True, True), (200, 257, 40, 10, True, True)])
def test_tcs_engine_vs_oracle(fb, oracle, F, B, K, iters, uw, uh):
    rng = np.random.default_rng(F + B + K)
    X = lowrank(rng, 3, F, B)
    with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05_STREAMED) as plan:
        W, H, V, _ = plan.nmf_process(X, K, iters, uw, uh, seeds=[3, 4, 5])
        assert plan.stats()["backend_used"] == fb.BACKEND_TCGEN05_STREAMED
    for b in range(3):
        Wo, Ho, Vo, _ = oracle.nmf_process(X[b], K, iters, uw, uh, 3 + b)
        assert rel(W[b], Wo) < TOL and rel(H[b], Ho) < TOL and rel(V[b], Vo) < TOL, (b, rel(W[b], Wo), rel(H[b], Ho))


def test_tcs_engine_matches_resident_and_simt_engines(fb):
    rng = np.random.default_rng(1)
    X = lowrank(rng, 5, 512, 513).astype(np.float32)
    seeds = np.arange(5)
    with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05_STREAMED) as pt, fb.Plan(win=64, backend=fb.BACKEND_TCGEN05) as pr, \
            fb.Plan(win=64, backend=fb.BACKEND_SIMT) as ps:
        Wt, Ht, Vt, _ = pt.nmf_process(X, 16, 50, True, True, seeds=seeds)
        Wt2, Ht2, _, _ = pt.nmf_process(X, 16, 50, True, True, seeds=seeds)
        Wr, Hr, Vr, _ = pr.nmf_process(X, 16, 50, True, True, seeds=seeds)
        assert pr.stats()["backend_used"] == fb.BACKEND_TCGEN05
        Ws, Hs, Vs, _ = ps.nmf_process(X, 16, 50, True, True, seeds=seeds)
    assert np.array_equal(Wt, Wt2) and np.array_equal(Ht, Ht2)
    for b in range(5):
        assert rel(Wt[b], Wr[b]) < TOL and rel(Ht[b], Hr[b]) < TOL and rel(Vt[b], Vr[b]) < TOL
        assert rel(Wt[b], Ws[b]) < TOL and rel(Ht[b], Hs[b]) < TOL and rel(Vt[b], Vs[b]) < TOL


@pytest.mark.parametrize("K", [16, 32, 64])
def test_tcs_engine_many_buffers_persistent_loop(fb, oracle, K):
    """more buffers than SMs: every CTA walks several buffers, barrier phases and job counters carry over"""
    rng = np.random.default_rng(2)
    base = lowrank(rng, 4, 256, 257)
    X = np.concatenate([base] * 80)[:310]
    seeds = np.arange(310) % 7
    with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05_STREAMED) as plan:
        W, H, _, _ = plan.nmf_process(X, K, 8, True, True, seeds=seeds, want_v=False)
    for b in (0, 5, 151, 309):
        Wo, Ho, _, _ = oracle.nmf_process(X[b], K, 8, True, True, int(seeds[b]))
        assert rel(W[b], Wo) < TOL and rel(H[b], Ho) < TOL
    assert np.array_equal(W[0], W[28]) and np.array_equal(H[3], H[31])


@pytest.mark.parametrize("K,iters,uw,uh,launches", [(16, 2, True, True, 30), (32, 2, True, True, 30), (16, 3, False, True, 30),
                                                    (32, 2, True, False, 20), (64, 2, True, True, 20), (64, 3, False, True, 20)])
def test_tcs_engine_identical_buffers_stay_identical_under_stress(fb, K, iters, uw, uh, launches):
    """Every CTA factorises copies of ONE spectrogram with ONE seed, several buffers per CTA, many launches: all buffers of
    all launches must be bit-identical (the reference's seed tests require repeatability, TestNMF.cpp:31-45) -- any race
    in the mbarrier / bulk-copy protocol shows up as a buffer that differs."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(11)
    X1 = lowrank(rng, 1, 512, 513).astype(np.float32)
    copies = 5 * 148
    X = torch.from_numpy(X1).cuda().expand(copies, -1, -1).contiguous()
    seeds = np.full(copies, 7, dtype=np.int64)
    ref = None
    with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05_STREAMED) as plan:
        for _ in range(launches):
            W, H, _, _ = plan.nmf_process(X, K, iters, uw, uh, seeds=seeds, want_v=False)
            if ref is None:
                ref = (W[0].clone(), H[0].clone())
                assert bool(torch.isfinite(ref[0]).all()) and bool(torch.isfinite(ref[1]).all())
            assert bool((W == ref[0]).all()) and bool((H == ref[1]).all())


@pytest.mark.parametrize("F,B,K", [(1000, 513, 16), (128, 257, 16), (3000, 513, 12), (700, 1025, 32), (600, 513, 64)])
def test_tcs_engine_process_frames(fb, oracle, F, B, K):
    """NMF::processFrame over many frames (NMF.hpp:45-89, NMFMatchClient.hpp:106-118): fixed dictionary, 10 iterations per
    frame, tile pairs per CTA (odd tile counts leave a single-tile unit)."""
    rng = np.random.default_rng(F + B)
    X = np.abs(rng.standard_normal((F, B))) * (rng.random((F, 1)) ** 2)
    X[5] = 0.0  # a silent frame: every magnitude is clamped to eps (NMF.hpp:60)
    W0 = rng.random((K, B)) ** 2
    with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05_STREAMED) as plan:
        H, V, Wn = plan.nmf_process_frames(X, W0, 10, seed=42, want_v=True, want_w=True)
        assert plan.stats()["backend_used"] == fb.BACKEND_TCGEN05_STREAMED
    Ho = oracle.nmfmatch_frames(X, W0, 10, 42)
    assert rel(H, Ho) < TOL
    for f in (0, 5, F // 2, F - 1):
        ho, vo, wo = oracle.nmf_process_frame(X[f], W0, 10, 42)
        assert rel(H[f], ho) < TOL and rel(V[f], vo) < TOL and rel(Wn, wo) < 1e-6


@pytest.mark.parametrize("K,iters", [(16, 200), (32, 200)])
def test_tcs_engine_full_iteration_counts(fb, oracle, K, iters):
    """BASELINE configs 2 and 3 at their full iteration counts on config-2 style audio: W, H within 1e-4 of the oracle."""
    from tests.golden.make_golden import synth_audio
    a = np.stack([synth_audio(1000 + b, 130816) for b in range(2)])
    with fb.Plan(win=1024, hop=256, fft=1024, backend=fb.BACKEND_TCGEN05_STREAMED) as plan:
        r = plan.bufnmf(a, K, iters, seeds=[0, 1])
        assert plan.stats()["backend_used"] == fb.BACKEND_TCGEN05_STREAMED
    for b in range(2):
        o = oracle.bufnmf_channel(a[b], 1024, 1024, 256, K, iters, b)
        assert rel(r["bases"][b], o["bases"]) < TOL and rel(r["acts"][b], o["acts"]) < TOL, (b, rel(r["bases"][b], o["bases"]))


def test_tcs_engine_exact_progress_and_cancel(fb, oracle):
    rng = np.random.default_rng(5)
    X = lowrank(rng, 2, 256, 257)
    seen = []
    with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05_STREAMED) as plan:
        W, H, V, st = plan.nmf_process(X, 32, 12, True, True, seeds=[1, 2], progress=lambda it: seen.append(it) or it < 5)
        assert plan.stats()["backend_used"] == fb.BACKEND_TCGEN05_STREAMED
    assert st == fb.CANCELLED and seen == [1, 2, 3, 4, 5] and np.array_equal(V, X)
    for b in range(2):
        Wo, Ho, _, _ = oracle.nmf_process(X[b], 32, 5, True, True, 1 + b)
        assert rel(W[b], Wo) < TOL and rel(H[b], Ho) < TOL


def test_bufnmf_host_call_split_on_the_streamed_engine(fb):
    """fb200_bufnmf with HOST audio at rank 32: the first sm_count buffers are started while the rest uploads (two launches,
    each with its own slice of the operand arrays); bit-identical to the unsplit device-memory call."""
    torch = pytest.importorskip("torch")
    from tests.golden.make_golden import synth_audio
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    batch, n = 3 * sm + 5, 8192
    a = np.stack([synth_audio(70 + (b % 7), n) * (1.0 + 0.01 * b) for b in range(batch)]).astype(np.float32)
    seeds = np.arange(batch)
    with fb.Plan(win=256, hop=64, fft=256) as plan:
        rd = plan.bufnmf(torch.from_numpy(a).cuda(), 32, 10, seeds=seeds)
        assert plan.stats()["backend_used"] == fb.BACKEND_TCGEN05_STREAMED and plan.stats()["update_kernel_launches"] == 1
        rh = plan.bufnmf(a, 32, 10, seeds=seeds)
        assert plan.stats()["backend_used"] == fb.BACKEND_TCGEN05_STREAMED and plan.stats()["update_kernel_launches"] == 2
    for k in ("bases", "acts"):
        assert np.array_equal(np.asarray(rd[k].cpu() if hasattr(rd[k], "cpu") else rd[k]), rh[k]), k
