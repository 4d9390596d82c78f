"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/flucoma_b200.h declares,
and its host-only helpers follow the reference's size rules.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fb():
    import flucoma_b200
    if not os.path.exists(flucoma_b200.LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("fb200_build", os.path.join(ROOT, "flucoma-core_b200", "build.py"))
        m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
        m.build()
    flucoma_b200.load()
    return flucoma_b200


def header_symbols():
    h = open(os.path.join(ROOT, "include", "flucoma_b200.h")).read()
    return sorted(set(re.findall(r"FB200_API\s+[\w\s\*]+?\b(fb200_\w+)\s*\(", h)))


def test_exports_every_declared_symbol(fb):
    syms = header_symbols()
    assert len(syms) >= 16
    L = C.CDLL(fb.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/flucoma_b200.h but not exported"
    assert sorted(fb.SYMBOLS) == syms  # the python binding covers the whole header


def test_abi_version_and_table(fb):
    L = fb.load()
    assert L.fb200_abi_version() == 2  # round 2: new entry points appended to the table, FB200_PROGRESS_ASYNC
    assert L.fb200_get_api(2) is not None and L.fb200_get_api(1) is None and L.fb200_get_api(999) is None


def test_struct_sizes_match_header_layout(fb):
    # natural alignment of the header's structs on LP64
    assert C.sizeof(fb.Config) == 48
    assert C.sizeof(fb.NmfArgs) == 136
    assert C.sizeof(fb.FramesArgs) == 88
    assert C.sizeof(fb.BufNmfArgs) == 120
    assert C.sizeof(fb.Stats) == 64
    assert C.sizeof(fb.FilterFramesArgs) == 64 and C.sizeof(fb.NmfCrossArgs) == 104 and C.sizeof(fb.ShardedArgs) == 32


def test_size_rules(fb, oracle):
    # ParameterTypes.hpp:295-313, NMFClient.hpp:111-113
    assert fb.resolve_fft(1024, -1, -1) == (512, 1024, 513)
    assert fb.resolve_fft(1000, -1, -1) == (500, 1024, 513)
    assert fb.resolve_fft(1024, 256, 4096) == (256, 4096, 2049)
    with pytest.raises(fb.FlucomaB200Error):
        fb.resolve_fft(1024, 256, 1000)  # not a power of two
    with pytest.raises(fb.FlucomaB200Error):
        fb.resolve_fft(1024, 256, 512)   # fft < win
    for n in (0, 1, 255, 256, 130816, 515088):
        assert fb.num_frames(n, 1024, 256) == oracle.num_frames(n, 1024, 256)
    assert fb.num_frames(130816, 1024, 256) == 512


def test_shard_range_partitions(fb):
    for total in (0, 1, 7, 1024, 8192, 8191):
        for world in (1, 2, 3, 8):
            got = [fb.shard_range(total, world, r) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == total
            for (b0, c0), (b1, _) in zip(got, got[1:]):
                assert b0 + c0 == b1
            counts = [c for _, c in got]
            assert max(counts) - min(counts) <= 1


def test_fails_loudly_without_gpu(fb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert fb.device_count() == 0
    with pytest.raises(fb.FlucomaB200Error) as e:
        fb.Plan(win=1024)
    assert e.value.code == -6  # FB200_ERR_NO_DEVICE: no CPU fallback


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "flucoma-core_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.lower() or f == "README.md", f"{f} mentions the oracle"


def test_bufstft_size_rules_match_the_oracle(fb):
    """fb200_bufstft_sizes is host-only: BufSTFTClient.hpp:121-131 / 241-242 against the CPU restatement, all padding modes."""
    from oracle import c_oracle
    c_oracle.build()
    for win, hop in ((1024, 256), (200, 50), (128, 128), (64, 16), (4096, 1024)):
        for mode in (0, 1, 2):
            for n in (win, win + 1, 3 * win - 7, 10 * win + hop // 2, 130816):
                try:
                    want = c_oracle.bufstft_sizes(win, hop, mode, False, n)
                except ValueError:
                    with pytest.raises(fb.FlucomaB200Error):
                        fb.bufstft_sizes(win, hop, mode, False, n)
                    continue
                assert fb.bufstft_sizes(win, hop, mode, False, n) == want
                frames = want[1]
                assert fb.bufstft_sizes(win, hop, mode, True, frames) == c_oracle.bufstft_sizes(win, hop, mode, True, frames)
    with pytest.raises(fb.FlucomaB200Error):
        fb.bufstft_sizes(1024, 256, 0, False, 100)        # shorter than one window
    with pytest.raises(fb.FlucomaB200Error):
        fb.bufstft_sizes(1024, 256, 3, False, 4096)       # no such padding mode
