"""GPU tests of the tcgen05 NMF engine (kernels_nmf_tc.cu) forced through FB200_BACKEND_TCGEN05: parity against the
fp64 oracle (1e-4 Frobenius-relative, the north_star bar), agreement with the SIMT engine, repeatability."""
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


@pytest.mark.parametrize("F,B,iters,uw,uh", [(128, 129, 5, True, True), (200, 257, 20, True, True), (512, 513, 30, True, True),
                                             (384, 513, 12, True, False), (256, 513, 12, False, True), (130, 513, 1, True, True)])
def test_tc_engine_vs_oracle(fb, oracle, F, B, iters, uw, uh):
    rng = np.random.default_rng(F + B)
    X = lowrank(rng, 3, F, B)
    with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05) as plan:
        W, H, V, _ = plan.nmf_process(X, 16, iters, uw, uh, seeds=[3, 4, 5])
        assert plan.stats()["backend_used"] == fb.BACKEND_TCGEN05
    for b in range(3):
        Wo, Ho, Vo, _ = oracle.nmf_process(X[b], 16, iters, uw, uh, 3 + b)
        assert rel(W[b], Wo) < TOL and rel(H[b], Ho) < TOL and rel(V[b], Vo) < TOL


def test_tc_engine_matches_simt_and_is_repeatable(fb):
    rng = np.random.default_rng(1)
    X = lowrank(rng, 5, 512, 513).astype(np.float32)
    with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05) as pt, fb.Plan(win=64, backend=fb.BACKEND_SIMT) as ps:
        Wt, Ht, Vt, _ = pt.nmf_process(X, 16, 50, True, True, seeds=np.arange(5))
        Wt2, Ht2, _, _ = pt.nmf_process(X, 16, 50, True, True, seeds=np.arange(5))
        Ws, Hs, Vs, _ = ps.nmf_process(X, 16, 50, True, True, seeds=np.arange(5))
        assert ps.stats()["backend_used"] == fb.BACKEND_SIMT
    assert np.array_equal(Wt, Wt2) and np.array_equal(Ht, Ht2)
    for b in range(5):
        assert rel(Wt[b], Ws[b]) < TOL and rel(Ht[b], Hs[b]) < TOL and rel(Vt[b], Vs[b]) < TOL


def test_tc_engine_many_buffers_persistent_loop(fb, oracle):
    """more buffers than SMs: every CTA walks several buffers, barrier phases carry over"""
    rng = np.random.default_rng(2)
    base = lowrank(rng, 4, 256, 257)
    X = np.concatenate([base] * 80)[:310]
    seeds = np.arange(310) % 7
    with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05) as plan:
        W, H, _, _ = plan.nmf_process(X, 16, 8, True, True, seeds=seeds, want_v=False)
    for b in (0, 5, 151, 309):
        Wo, Ho, _, _ = oracle.nmf_process(X[b], 16, 8, True, True, int(seeds[b]))
        assert rel(W[b], Wo) < TOL and rel(H[b], Ho) < TOL
    # identical (audio, seed) pairs give identical results wherever they sit in the batch
    assert np.array_equal(W[0], W[28]) and np.array_equal(H[3], H[31])


def test_tc_engine_rejects_unqualified_shapes(fb):
    X = np.random.default_rng(0).random((1, 64, 100))
    with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05) as plan:
        with pytest.raises(fb.FlucomaB200Error):
            plan.nmf_process(X, 16, 3, True, True, seeds=1)


@pytest.mark.parametrize("iters,uw,uh,launches", [(1, True, False, 150), (1, True, True, 100), (3, True, True, 60)])
def test_tc_engine_identical_buffers_stay_identical_under_stress(fb, iters, uw, uh, launches):
    """Regression for two synchronisation races found on the B200 (a parity wait on the V ring that could pass one phase
    early when TMA was slow; the two warps sharing a tile row in the H-update storing before the other had loaded).
    Every CTA factorises copies of ONE spectrogram with ONE seed, several buffers per CTA, many launches: all buffers of
    all launches must be bit-identical (the reference's seed tests require repeatability, TestNMF.cpp:31-45), and no
    launch may fault."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(11)
    X1 = lowrank(rng, 1, 512, 513).astype(np.float32)
    copies = 7 * 148
    X = torch.from_numpy(X1).cuda().expand(copies, -1, -1).contiguous()
    seeds = np.full(copies, 7, dtype=np.int64)
    ref = None
    with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05) as plan:
        for _ in range(launches):
            W, H, _, _ = plan.nmf_process(X, 16, iters, uw, uh, seeds=seeds, want_v=False)
            if ref is None:
                ref = (W[0].clone(), H[0].clone())
                assert bool(torch.isfinite(ref[0]).all()) and bool(torch.isfinite(ref[1]).all())
            assert bool((W == ref[0]).all()) and bool((H == ref[1]).all())


# ---------------------------------------------------------------------------------------------- progress / cancel
def test_tc_engine_exact_progress_and_cancel(fb, oracle):
    """A progress callback no longer pushes the call off the tensor-core engine (the reference ALWAYS installs one,
    NMFClient.hpp:261-267).  Exact mode at stride 1: a cancel at iteration 7 leaves W,H as after 7 reference iterations
    and V1 = X (NMF.hpp:175-176); stride 5: every iteration is reported, in order, and the result matches."""
    rng = np.random.default_rng(5)
    X = lowrank(rng, 2, 256, 257)
    seen = []
    with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05) as plan:
        W, H, V, st = plan.nmf_process(X, 16, 20, True, True, seeds=[1, 2], progress=lambda it: seen.append(it) or it < 7)
        assert plan.stats()["backend_used"] == fb.BACKEND_TCGEN05
        assert st == fb.CANCELLED and seen == list(range(1, 8)) and np.array_equal(V, X)
        seen5 = []
        W5, H5, V5, st5 = plan.nmf_process(X, 16, 20, True, True, seeds=[1, 2], progress=lambda it: seen5.append(it) or True,
                                           progress_stride=5)
        Wn, Hn, Vn, _ = plan.nmf_process(X, 16, 20, True, True, seeds=[1, 2])
    for b in range(2):
        Wo, Ho, _, _ = oracle.nmf_process(X[b], 16, 7, True, True, 1 + b)
        assert rel(W[b], Wo) < TOL and rel(H[b], Ho) < TOL
    assert st5 == 0 and seen5 == list(range(1, 21))
    assert rel(W5, Wn) < 1e-5 and rel(H5, Hn) < 1e-5 and rel(V5, Vn) < 1e-5


@pytest.mark.parametrize("uw,uh", [(True, True), (False, True), (True, False)])
def test_tc_engine_async_progress(fb, uw, uh):
    """FB200_PROGRESS_ASYNC: one uninterrupted persistent launch; iterations 1..n are reported exactly once, in order,
    and the outputs are bit-identical to the call without a callback."""
    rng = np.random.default_rng(6)
    X = lowrank(rng, 200, 256, 257).astype(np.float32)
    seeds = np.arange(200)
    seen = []
    with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05) as plan:
        Wa, Ha, _, st = plan.nmf_process(X, 16, 30, uw, uh, seeds=seeds, want_v=False,
                                         progress=lambda it: seen.append(it) or True, progress_stride=fb.PROGRESS_ASYNC)
        assert plan.stats()["backend_used"] == fb.BACKEND_TCGEN05 and plan.stats()["update_kernel_launches"] == 1
        Wn, Hn, _, _ = plan.nmf_process(X, 16, 30, uw, uh, seeds=seeds, want_v=False)
    assert st == 0 and seen == list(range(1, 31))
    assert np.array_equal(Wa, Wn) and np.array_equal(Ha, Hn)


def test_tc_engine_async_cancel_finishes_the_iteration_in_flight(fb):
    """Cancel in async mode: the call returns CANCELLED, every buffer holds the state after SOME completed iteration
    (W and H of the same iteration: the fused schedule runs one H-only pass after the cancel), buffers not yet started
    keep their initial state.  All buffers are copies of one problem, so a table of the uncancelled results after
    j = 0..n iterations must contain every buffer of the cancelled call, bit for bit."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(8)
    X1 = lowrank(rng, 1, 384, 257).astype(np.float32)
    copies, iters = 10 * 148, 24
    X = torch.from_numpy(X1).cuda().expand(copies, -1, -1).contiguous()
    seeds = np.full(copies, 3, dtype=np.int64)
    with fb.Plan(win=64, backend=fb.BACKEND_TCGEN05) as plan:
        table = []
        for j in range(iters + 1):
            Wj, Hj, _, _ = plan.nmf_process(X[:1], 16, j, True, True, seeds=seeds[:1], want_v=False)
            table.append((Wj[0].clone(), Hj[0].clone()))
        for attempt in range(3):  # the cancel races with a ~5 ms kernel: retry if the host thread was descheduled past its end
            calls = []
            W, H, V, st = plan.nmf_process(X, 16, iters, True, True, seeds=seeds, want_v=True,
                                           progress=lambda it: calls.append(it) or it < 3, progress_stride=fb.PROGRESS_ASYNC)
            assert st == fb.CANCELLED and calls == [1, 2, 3]
            assert bool((V == X).all())                      # NMF.hpp:175-176
            at = []
            for b in range(copies):
                hit = [j for j, (Wj, Hj) in enumerate(table) if bool((W[b] == Wj).all()) and bool((H[b] == Hj).all())]
                assert hit, f"buffer {b} is not at any completed iteration"
                at.append(hit[0])
            if min(at) < iters:
                break
    assert min(at) < iters, "the cancel arrived after everything had finished in three attempts"


def test_bufnmf_host_call_split_into_a_first_round_and_the_rest(fb):
    """fb200_bufnmf with HOST audio on the resident engine starts one buffer per SM while the rest of the batch is still
    being uploaded (two persistent launches).  Results must be bit-identical to the unsplit device-memory call, with and
    without an asynchronous progress callback, and the callback must still see iterations 1..n exactly once, in order."""
    torch = pytest.importorskip("torch")
    from tests.golden.make_golden import synth_audio
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    batch, n = 3 * sm + 7, 8192
    a = np.stack([synth_audio(50 + (b % 9), n) * (1.0 + 0.01 * b) for b in range(batch)]).astype(np.float32)
    seeds = np.arange(batch)
    with fb.Plan(win=256, hop=64, fft=256) as plan:
        rd = plan.bufnmf(torch.from_numpy(a).cuda(), 16, 12, seeds=seeds, resynth=True)      # device memory: one launch
        assert plan.stats()["backend_used"] == fb.BACKEND_TCGEN05 and plan.stats()["update_kernel_launches"] == 1
        rh = plan.bufnmf(a, 16, 12, seeds=seeds, resynth=True)                               # host memory: split
        assert plan.stats()["backend_used"] == fb.BACKEND_TCGEN05 and plan.stats()["update_kernel_launches"] == 2
        seen = []
        rp = plan.bufnmf(a, 16, 12, seeds=seeds, progress=lambda it: seen.append(it) or True, progress_stride=fb.PROGRESS_ASYNC)
        assert plan.stats()["update_kernel_launches"] == 2
    assert seen == list(range(1, 13))
    for k in ("bases", "acts", "resynth"):
        assert np.array_equal(np.asarray(rd[k].cpu() if hasattr(rd[k], "cpu") else rd[k]), rh[k]), k
    assert np.array_equal(rh["bases"], rp["bases"]) and np.array_equal(rh["acts"], rp["acts"])
