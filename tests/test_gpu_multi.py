"""Product-level multi-GPU entry fb200_bufnmf_sharded: one process, one plan (and host thread) per device, contiguous
shards, no data-path collective, one ncclAllGather of the final activations.  Runs on however many devices
the box has (1 on the default test box: the shard / gather plumbing still runs, with a world of one)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fb():
    import flucoma_b200
    return flucoma_b200


@pytest.mark.parametrize("batch", [7, 3])
def test_sharded_matches_single_plan_and_gathers(fb, batch):
    from tests.golden.make_golden import synth_audio
    torch = pytest.importorskip("torch")
    ndev = min(fb.device_count(), 4)
    a = np.stack([synth_audio(2000 + b, 8192) for b in range(batch)])
    seeds = np.arange(batch)
    plans = [fb.Plan(win=512, hop=128, fft=512, device=d) for d in range(ndev)]
    try:
        seen = []
        r = fb.bufnmf_sharded(plans, a, 4, 25, seeds=seeds, resynth=True, gather=True, progress=lambda it: seen.append(it) or True)
        one = plans[0].bufnmf(a, 4, 25, seeds=seeds, resynth=True)
    finally:
        for p in plans:
            p.close()
    assert r["status"] == 0 and seen == list(range(1, 26))
    # buffers are independent and seeded individually: sharding must not change a single bit
    for k in ("bases", "acts", "resynth"):
        assert np.array_equal(r[k], one[k]), k
    per = (batch + ndev - 1) // ndev
    F = r["acts"].shape[1]
    for d, g in enumerate(r["gathered"]):
        g = g.cpu().numpy()
        assert g.shape == (ndev * per, F, 4)
        for rank in range(ndev):
            b0, cnt = fb.shard_range(batch, ndev, rank)
            assert np.array_equal(g[rank * per:rank * per + cnt], r["acts"][b0:b0 + cnt]), (d, rank)


def test_sharded_cancel(fb):
    from tests.golden.make_golden import synth_audio
    ndev = min(fb.device_count(), 4)
    a = np.stack([synth_audio(2100 + b, 8192) for b in range(5)])
    plans = [fb.Plan(win=512, hop=128, fft=512, device=d) for d in range(ndev)]
    try:
        seen = []
        r = fb.bufnmf_sharded(plans, a, 4, 40, seeds=np.arange(5), progress=lambda it: seen.append(it) or it < 4,
                              progress_stride=1)
    finally:
        for p in plans:
            p.close()
    assert r["status"] == fb.CANCELLED and seen == [1, 2, 3, 4]
