"""C++17 host mirror of the reference interface (flucoma-core_b200/host/flucoma/...): containers on the CPU, the
algorithm/client shims on the GPU (TestNMF.cpp replica compiled against our headers)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "flucoma-core_b200", "host")
LIB = os.path.join(ROOT, "flucoma-core_b200", "lib", "libflucoma_b200.so")


def compile_cpp(src, out):
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", HOST, "-I", ROOT, os.path.join(ROOT, "tests", "cpp", src),
           "-ldl", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_host_containers_cpu():
    with tempfile.TemporaryDirectory() as t:
        exe = compile_cpp("test_host_containers.cpp", os.path.join(t, "a"))
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0 and "host containers ok" in r.stdout, r.stdout + r.stderr


def test_shims_compile_and_fail_loudly_without_gpu():
    import torch
    with tempfile.TemporaryDirectory() as t:
        exe = compile_cpp("test_shims_gpu.cpp", os.path.join(t, "a"))
        if torch.cuda.is_available():
            pytest.skip("GPU present: covered by the gpu test")
        r = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, FLUCOMA_B200_LIB=LIB))
        assert r.returncode != 0 and "no such CUDA device" in (r.stdout + r.stderr)  # no CPU fallback


def _read_dump(path):
    out = []
    with open(path, "rb") as f:
        for _ in range(4):
            r, c = np.frombuffer(f.read(16), np.int64)
            out.append(np.frombuffer(f.read(4 * r * c), np.float32).reshape(r, c))
    return out


@pytest.mark.gpu
def test_shims_gpu_vs_oracle(oracle):
    with tempfile.TemporaryDirectory() as t:
        exe = compile_cpp("test_shims_gpu.cpp", os.path.join(t, "a"))
        dump = os.path.join(t, "dump.bin")
        r = subprocess.run([exe, dump], capture_output=True, text=True, env=dict(os.environ, FLUCOMA_B200_LIB=LIB))
        assert r.returncode == 0 and "host shims ok" in r.stdout, r.stdout + r.stderr
        src, bases, acts, res = _read_dump(dump)
    n, chans = src.shape
    rank = bases.shape[1] // chans
    for c in range(chans):
        o = oracle.bufnmf_channel(np.ascontiguousarray(src[:, c]), 256, 256, 64, rank, 30, 7, resynth=True)
        for j in range(rank):
            ch = c * rank + j  # NMFClient.hpp:277-300: component j of channel c lives in buffer channel c*rank+j
            assert np.linalg.norm(bases[:, ch] - o["bases"][j]) / np.linalg.norm(o["bases"][j]) < 1e-4
            assert np.linalg.norm(acts[:, ch] - o["acts"][:, j]) / np.linalg.norm(o["acts"][:, j]) < 1e-4
            assert np.linalg.norm(res[:, ch] - o["resynth"][j]) / max(np.linalg.norm(o["resynth"][j]), 1e-12) < 1e-4


def test_bufstft_client_compiles_and_fails_loudly_without_gpu():
    import torch
    with tempfile.TemporaryDirectory() as t:
        exe = compile_cpp("test_bufstft_gpu.cpp", os.path.join(t, "a"))
        if torch.cuda.is_available():
            pytest.skip("GPU present: covered by the gpu test")
        r = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, FLUCOMA_B200_LIB=LIB))
        assert r.returncode != 0 and "no such CUDA device" in (r.stdout + r.stderr)  # no CPU fallback


@pytest.mark.gpu
def test_bufstft_client_gpu_vs_oracle(oracle):
    with tempfile.TemporaryDirectory() as t:
        exe = compile_cpp("test_bufstft_gpu.cpp", os.path.join(t, "a"))
        dump = os.path.join(t, "dump.bin")
        r = subprocess.run([exe, dump], capture_output=True, text=True, env=dict(os.environ, FLUCOMA_B200_LIB=LIB))
        assert r.returncode == 0 and "bufstft client ok" in r.stdout, r.stdout + r.stderr
        src, mag, phase, res = _read_dump(dump)
    mo, po = oracle.bufstft_fwd(np.ascontiguousarray(src[:, 1]), 256, 256, 64, 1)
    z = mag.astype(np.float64) * np.exp(1j * phase.astype(np.float64))
    zo = mo.astype(np.float64) * np.exp(1j * po.astype(np.float64))
    assert mag.shape == mo.shape and np.linalg.norm(z - zo) / np.linalg.norm(zo) < 2e-6
    ro = oracle.bufstft_inv(mo, po, 256, 256, 64, 1)
    assert res.shape == (ro.size, 1) and np.linalg.norm(res[:, 0] - ro) / np.linalg.norm(ro) < 1e-4


def test_rt_client_ring_buffers_cpu():
    """FluidSource / FluidSink / BufferedProcess mirrors: the reference's own ring-buffer tests restated
    (tests/clients/common/TestFluidSource.cpp:39-55, TestBufferedProcess.cpp:20-70) -- no device needed."""
    with tempfile.TemporaryDirectory() as t:
        exe = compile_cpp("test_rt_clients_gpu.cpp", os.path.join(t, "a"))
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0 and "rt clients cpu ok" in r.stdout, r.stdout + r.stderr


def _read_dumps(path, count):
    out = []
    with open(path, "rb") as f:
        for _ in range(count):
            r, c = np.frombuffer(f.read(16), np.int64)
            out.append(np.frombuffer(f.read(4 * r * c), np.float32).reshape(r, c))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("host", [64, 100, 256])
def test_rt_clients_gpu_vs_oracle_stream(oracle, host):
    """NMFFilterClient / NMFMatchClient mirrors driven block by block like an audio callback (host vectors of 64, 100 and
    256 samples) against the oracle's simulation of the reference's streaming clients."""
    win, hop, rank = 256, 64, 5
    with tempfile.TemporaryDirectory() as t:
        exe = compile_cpp("test_rt_clients_gpu.cpp", os.path.join(t, "a"))
        dump = os.path.join(t, "dump.bin")
        r = subprocess.run([exe, dump, str(host)], capture_output=True, text=True, env=dict(os.environ, FLUCOMA_B200_LIB=LIB))
        assert r.returncode == 0 and "rt clients ok" in r.stdout, r.stdout + r.stderr
        audio, W, fout, mout = _read_dumps(dump, 4)
    audio = audio[0]
    n_out = fout.shape[1]
    out_o, acts_o = oracle.nmffilter_stream(audio.astype(np.float64), win, win, hop, W.astype(np.float64), 10, 42)
    for k in range(rank):  # NMFFilterClient.hpp:98-117 through STFTBufferedProcess<true>
        assert np.linalg.norm(fout[k] - out_o[k, :n_out]) / np.linalg.norm(out_o[k, :n_out]) < 1e-4
    assert np.all(fout[rank:] == 0)  # channels beyond the rank of the bases buffer stay silent
    # NMFMatchClient.hpp:110-118: a block reports the activations of the last frame of the block before
    assert np.all(mout[0] == 0)
    for b in range(1, mout.shape[0]):
        f_last = -(-b * host // hop) - 1
        ref = acts_o[f_last]
        assert np.linalg.norm(mout[b, :rank] - ref) / max(np.linalg.norm(ref), 1e-30) < 1e-4, b
    assert np.all(mout[:, rank:] == 0)


def test_nrt_client_mirrors_compile_and_fail_loudly_without_gpu():
    import torch
    with tempfile.TemporaryDirectory() as t:
        exe = compile_cpp("test_nrt_clients_gpu.cpp", os.path.join(t, "a"))
        if torch.cuda.is_available():
            pytest.skip("GPU present: covered by the gpu test")
        r = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, FLUCOMA_B200_LIB=LIB))
        assert r.returncode != 0  # no CPU fallback: NMFSeed returns kError ("no such CUDA device") and the test's CHECK fails
        assert "no such CUDA device" in (r.stdout + r.stderr)


@pytest.mark.gpu
def test_nrt_client_mirrors_gpu_vs_oracle(oracle):
    """NMFSeedClient -> seeded NMFClient, and NMFCrossClient (with a FluidTask) through the C++ mirrors; the BufNMFCross
    output against the oracle (NMFCrossClient.hpp:85-185)."""
    with tempfile.TemporaryDirectory() as t:
        exe = compile_cpp("test_nrt_clients_gpu.cpp", os.path.join(t, "a"))
        dump = os.path.join(t, "dump.bin")
        r = subprocess.run([exe, dump], capture_output=True, text=True, env=dict(os.environ, FLUCOMA_B200_LIB=LIB))
        assert r.returncode == 0 and "nrt clients ok" in r.stdout, r.stdout + r.stderr
        src, tgt, out = _read_dumps(dump, 3)
    ref, _ = oracle.bufnmfcross(src[:, 0], tgt[:, 0], 256, 256, 64, 7, 11, 7, 20, 5, 50)
    assert np.linalg.norm(out[:, 0] - ref) / np.linalg.norm(ref) < 1e-3


def test_spectral_mirrors_compile():
    """MelBands / HPSS host mirrors (algorithms/public/MelBands.hpp:27-108, HPSS.hpp:30-186) compile on a CPU-only box."""
    with tempfile.TemporaryDirectory() as t:
        compile_cpp("test_spectral_mirrors_gpu.cpp", os.path.join(t, "a"))


@pytest.mark.gpu
def test_spectral_mirrors_gpu_vs_oracle(oracle):
    """MelBands::processFrame(s) and HPSS::processFrame(s) through the C++ mirrors: streaming calls equal the batched ones
    (checked inside the executable), the batched outputs equal the oracle."""
    with tempfile.TemporaryDirectory() as t:
        exe = compile_cpp("test_spectral_mirrors_gpu.cpp", os.path.join(t, "a"))
        dump = os.path.join(t, "dump.bin")
        r = subprocess.run([exe, dump], capture_output=True, text=True, env=dict(os.environ, FLUCOMA_B200_LIB=LIB))
        assert r.returncode == 0 and "spectral mirrors ok" in r.stdout, r.stdout + r.stderr
        M, bands, sp, hh = _read_dumps(dump, 4)
    ref = oracle.melbands(M.astype(np.float64), 20.0, 20000.0, 13, 44100.0, 256)
    assert np.linalg.norm(bands - ref) / np.linalg.norm(ref) < 1e-5
    S = sp[:, 0::2] + 1j * sp[:, 1::2]
    H = (hh[:, 0::2] + 1j * hh[:, 1::2]).reshape(3, S.shape[0], S.shape[1])
    href = oracle.hpss(S.astype(np.complex128), 7, 9, 0)
    assert np.linalg.norm(H - href) / np.linalg.norm(href) < 1e-5
