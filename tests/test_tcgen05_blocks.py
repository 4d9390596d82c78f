"""GPU test of the tcgen05 / TMEM / TMA building blocks (fb200_selftest_tcgen05) in the four operand forms the
tensor-core NMF engine uses, against numpy on bf16-rounded inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bf16(x):
    """round-to-nearest-even to bfloat16, returned as float32"""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return u.astype(np.uint32).view(np.float32)


def test_tcgen05_blocks():
    import flucoma_b200 as fb
    rng = np.random.default_rng(0)
    H1 = rng.random((128, 16)); W1 = rng.random((16, 64)); R1 = rng.random((128, 64)) * 4
    W2 = rng.random((16, 128)); H2 = rng.random((64, 16)); R2 = rng.random((128, 64)) * 4
    V = rng.random((128, 68))
    inp = np.concatenate([x.ravel() for x in (H1, W1, R1, W2, H2, R2, V)]).astype(np.float32)
    with fb.Plan(win=64) as plan:
        out = plan.selftest_tcgen05(inp)
    o1, o2, o3, o4, o5 = np.split(out, np.cumsum([128 * 64, 128 * 16, 128 * 64, 128 * 16]))
    H1b, W1b, R1b, W2b, H2b, R2b = (bf16(x).astype(np.float64) for x in (H1, W1, R1, W2, H2, R2))

    def check(name, got, want):
        err = np.abs(got - want).max() / np.abs(want).max()
        assert err < 2e-6, (name, err, got[:2, :4], want[:2, :4])

    check("ss_Kmajor_MNmajor", o1.reshape(128, 64), H1b @ W1b)
    check("ts_Kmajor", o2.reshape(128, 16), R1b @ W1b.T)
    check("ss_MNmajor_Kmajor", o3.reshape(128, 64), W2b.T @ H2b.T)
    check("ts_MNmajor", o4.reshape(128, 16), R2b @ H2b)
    assert np.array_equal(o5.reshape(128, 32), V[:, 32:64].astype(np.float32)), "TMA swizzle read-back"
