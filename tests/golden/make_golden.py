"""Generates the committed golden fixtures under tests/golden/.  Run HERE (container with /root/reference and g++):

    python tests/golden/make_golden.py

Sources of truth, in order of independence:
  * rng_kat.json        -- produced by a 6-line C++ program using libstdc++'s std::mt19937_64 +
                           std::uniform_real_distribution<double>, i.e. exactly what EigenRandom.hpp:73-110 runs.
  * fft_kat.npz         -- analytic DFT facts + numpy.fft (pocketfft) values.
  * nmf_small.npz       -- TestNMF.cpp's 3x3 matrix (tests/algorithms/public/TestNMF.cpp:18-27) and a small
                           synthetic spectrogram through the C oracle, cross-checked against the numpy oracle here.
  * bufnmf_wav.npz      -- BASELINE config 1 style: Resources/AudioFiles/Tremblay-AaS-SynthTwoVoices-M.wav
                           (first 32768 samples), fft 1024 hop 256 rank 4 iters 100 seed 42, incl. resynthesis.
  * nmffilter_stream.npz -- NMFFilter/NMFMatch over a 6000-sample stream (win 256, hop 64, the rank-5 bases of
                           nmf_small, 10 iterations, seed 42): C oracle output, cross-checked against a numpy
                           simulation of the client's ring buffers driven with host vectors of 64 and 100 samples.
  * bufstft.npz         -- BufSTFT forward (mag, phase) and inverse of a 4000-sample synthetic buffer for the three
                           padding modes (win 200, fft 256, hop 50): C oracle output, cross-checked against numpy.
The reference itself cannot be executed (Eigen/HISSTools absent), so these are oracle outputs: "parity unpinned".
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import c_oracle as co, np_oracle as no  # noqa: E402

REF_AUDIO = "/root/reference/Resources/AudioFiles"


def read_wav_mono(path):
    """Minimal PCM reader (16/24-bit), scaling as libsndfile/HISSTools do: int / 2^(bits-1)."""
    import struct
    with open(path, "rb") as f:
        d = f.read()
    assert d[:4] == b"RIFF" and d[8:12] == b"WAVE"
    pos = 12; fmt = None; data = None
    while pos + 8 <= len(d):
        cid = d[pos:pos + 4]; sz = struct.unpack("<I", d[pos + 4:pos + 8])[0]
        body = d[pos + 8:pos + 8 + sz]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
        elif cid == b"data":
            data = body
        pos += 8 + sz + (sz & 1)
    _, ch, sr, _, _, bits = fmt
    if bits == 16:
        x = np.frombuffer(data, "<i2").astype(np.float64) / 32768.0
    elif bits == 24:
        b = np.frombuffer(data, np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v - (1 << 24), v)
        x = v.astype(np.float64) / 8388608.0
    else:
        raise ValueError(bits)
    return x.reshape(-1, ch)[:, 0], sr


def rng_kat():
    src = r'''
#include <random>
#include <cstdio>
int main(){ unsigned long seeds[] = {42ul, 5063ul, 0ul, 4203ul, 7863ul};
  printf("{");
  for (int s = 0; s < 5; s++) { std::mt19937_64 g{seeds[s]}; std::uniform_real_distribution<double> d{0.0,1.0};
    printf("%s\"%lu\": [", s ? ", " : "", seeds[s]);
    for (int i = 0; i < 700; i++) printf("%s%.17g", i ? ", " : "", d(g));
    printf("]"); }
  printf("}\n"); }
'''
    with tempfile.TemporaryDirectory() as t:
        with open(os.path.join(t, "r.cpp"), "w") as f:
            f.write(src)
        subprocess.run(["/usr/bin/g++", "-O2", "-o", os.path.join(t, "r"), os.path.join(t, "r.cpp")], check=True)
        out = subprocess.run([os.path.join(t, "r")], check=True, capture_output=True, text=True).stdout
    kat = json.loads(out)
    for s, v in kat.items():
        assert np.array_equal(np.array(v), co.random_uniform(int(s), len(v))), s
        assert np.array_equal(np.array(v), no.random_uniform(int(s), len(v))), s
    with open(os.path.join(HERE, "rng_kat.json"), "w") as f:
        json.dump(kat, f)


def fft_kat():
    rng = np.random.default_rng(11)
    x = rng.standard_normal(1024)
    X = np.fft.rfft(x)
    X[0] = X[0].real; X[-1] = X[-1].real
    x200 = rng.standard_normal(200)
    X200 = np.fft.rfft(x200, n=256)
    np.savez_compressed(os.path.join(HERE, "fft_kat.npz"), x=x, X=X, x200=x200, X200=X200)


def synth_audio(seed, n, sr=44100.0):
    """BASELINE config-2 style synthetic buffer: 6 gated partials + noise at -40 dBFS, peak 0.9."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    y = np.zeros(n)
    for _ in range(6):
        f = np.exp(rng.uniform(np.log(80.0), np.log(8000.0)))
        a = rng.uniform(0.1, 1.0)
        env = np.zeros(n); pos = 0; on = bool(rng.integers(0, 2))
        while pos < n:
            seg = int(rng.uniform(0.05, 0.4) * sr)
            if on:
                env[pos:pos + seg] = 1.0
            on = not on; pos += seg
        y += a * env * np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi))
    y += 0.01 * rng.standard_normal(n)
    return (0.9 * y / np.abs(y).max()).astype(np.float32)


def nmf_small():
    X3 = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]], dtype=np.float64)
    out = {"X3": X3}
    for seed in (42, 5063):
        W, H, V, _ = co.nmf_process(X3, 2, 1, True, True, seed)
        W2, H2, V2, _ = no.nmf_process(X3, 2, 1, True, True, seed)
        assert np.allclose(W, W2, rtol=1e-12, atol=0) and np.allclose(H, H2, rtol=1e-12, atol=0)
        out[f"W3_{seed}"] = W; out[f"H3_{seed}"] = H; out[f"V3_{seed}"] = V
    a = synth_audio(1000, 6000)
    S = co.stft(a.astype(np.float64), 256, 256, 64)
    M = co.magnitude(S)
    out["audio"] = a; out["mag"] = M
    for tag, (uw, uh) in {"wh": (True, True), "w": (True, False), "h": (False, True)}.items():
        W, H, V, _ = co.nmf_process(M, 5, 40, uw, uh, 3)
        W2, H2, V2, _ = no.nmf_process(M, 5, 40, uw, uh, 3)
        assert np.allclose(W, W2, rtol=1e-9, atol=1e-14) and np.allclose(H, H2, rtol=1e-9, atol=1e-14)
        out[f"W_{tag}"] = W; out[f"H_{tag}"] = H; out[f"V_{tag}"] = V
    # seeded variant: W0/H0 supplied
    rng = np.random.default_rng(5)
    W0 = rng.random((5, M.shape[1])); H0 = rng.random((M.shape[0], 5))
    W, H, V, _ = co.nmf_process(M, 5, 25, True, True, -1, W0=W0, H0=H0)
    out["W0"] = W0; out["H0"] = H0; out["W_seeded"] = W; out["H_seeded"] = H; out["V_seeded"] = V
    # processFrame
    h, v, Wm = co.nmf_process_frame(M[20], out["W_wh"], 10, 42)
    out["pf_h"] = h; out["pf_v"] = v; out["pf_W"] = Wm
    np.savez_compressed(os.path.join(HERE, "nmf_small.npz"), **out)


def bufnmf_wav():
    x, sr = read_wav_mono(os.path.join(REF_AUDIO, "Tremblay-AaS-SynthTwoVoices-M.wav"))
    a = x[:32768].astype(np.float32)
    r = co.bufnmf_channel(a, 1024, 1024, 256, 4, 100, 42, resynth=True, debug=True)
    r2 = no.bufnmf_channel(a, 1024, 1024, 256, 4, 100, 42, resynth=True)
    for k in ("bases", "acts", "resynth"):
        err = np.abs(r[k].astype(np.float64) - r2[k]).max() / np.abs(r2[k]).max()
        assert err < 1e-6, (k, err)
    np.savez_compressed(os.path.join(HERE, "bufnmf_wav.npz"), audio=a, sr=sr, bases=r["bases"], acts=r["acts"],
                        resynth=r["resynth"].astype(np.float16 if False else np.float32), W=r["W"], H=r["H"])


def nmffilter_stream():
    g = np.load(os.path.join(HERE, "nmf_small.npz"))
    a = g["audio"]; W = g["W_wh"].astype(np.float32)
    out, acts = co.nmffilter_stream(a.astype(np.float64), 256, 256, 64, W.astype(np.float64), 10, 42)
    for hs in (64, 100):
        o2, h2 = no.nmffilter_stream(a.astype(np.float64), 256, 256, 64, W.astype(np.float64), 10, 42, host_size=hs)
        n2, f2 = o2.shape[1], h2.shape[0]
        assert np.abs(o2 - out[:, :n2]).max() < 1e-12 and np.abs(h2 - acts[:f2]).max() < 1e-12
    np.savez_compressed(os.path.join(HERE, "nmffilter_stream.npz"), audio=a, bases=W, out=out.astype(np.float32),
                        acts=acts)


def bufstft():
    a = synth_audio(77, 4000)
    out = {"audio": a}
    for mode in (0, 1, 2):
        m, p = co.bufstft_fwd(a, 200, 256, 50, mode)
        m2, p2 = no.bufstft_fwd(a, 200, 256, 50, mode)
        z, z2 = m.astype(np.float64) * np.exp(1j * p.astype(np.float64)), m2.astype(np.float64) * np.exp(1j * p2.astype(np.float64))
        assert np.abs(z - z2).max() <= 1e-6 * np.abs(z2).max()
        r = co.bufstft_inv(m, p, 200, 256, 50, mode)
        assert np.abs(r - no.bufstft_inv(m, p, 200, 256, 50, mode)).max() < 1e-6
        out[f"mag{mode}"] = m; out[f"phase{mode}"] = p; out[f"inv{mode}"] = r
    np.savez_compressed(os.path.join(HERE, "bufstft.npz"), **out)


def nmfcross():
    """BufNMFCross (NMFCrossClient.hpp:85-185) on two short synthetic buffers: win 256, hop 64, sparsity 7, polyphony 11,
    continuity 7, 30 iterations, seed 5, 50 Griffin-Lim iterations.  C oracle output, cross-checked against numpy."""
    src = synth_audio(31, 6000); tgt = synth_audio(32, 5000)
    out, H = co.bufnmfcross(src, tgt, 256, 256, 64, 7, 11, 7, 30, 5, 50)
    out2, H2 = no.bufnmfcross(src, tgt, 256, 256, 64, 7, 11, 7, 30, 5, 50)
    assert np.abs(H - H2).max() <= 1e-10 * np.abs(H2).max() and np.abs(out - out2).max() <= 1e-6 * np.abs(out2).max()
    np.savez_compressed(os.path.join(HERE, "nmfcross.npz"), source=src, target=tgt, out=out, H=H)


def spectral_epilogues():
    """MelBands (MelBands.hpp:43-101) and HPSS (HPSS.hpp:47-162) on the STFT of a short synthetic buffer: C oracle output,
    cross-checked against the numpy restatement (closed form of the delay lines instead of the streaming recursion)."""
    a = synth_audio(41, 12000)
    S = co.stft(a.astype(np.float64), 512, 512, 128)
    out = {"audio": a}
    for tag, args in {"norm": (True, False, False), "pow_db": (False, True, True)}.items():
        b1 = co.melbands(np.abs(S), 20.0, 20000.0, 40, 44100.0, 512, *args)
        b2 = no.melbands(np.abs(S), 20.0, 20000.0, 40, 44100.0, 512, *args)
        assert np.abs(b1 - b2).max() <= 1e-10 * np.abs(b2).max()
        out[f"mel_{tag}"] = b1
    for mode, ht, pt in [(0, (0, 1, 1, 1), (0, 1, 1, 1)), (1, (0.005, 0.0, 0.1, -10.0), (0, 1, 1, 1)),
                         (2, (0.005, 10.0, 0.1, 0.0), (0.005, 10.0, 0.1, 0.0))]:
        o1 = co.hpss(S, 31, 17, mode, ht, pt); o2 = no.hpss(S, 31, 17, mode, ht, pt)
        assert np.abs(o1 - o2).max() <= 1e-12 * np.abs(o2).max()
        out[f"hpss{mode}"] = o1.astype(np.complex64)
    np.savez_compressed(os.path.join(HERE, "spectral.npz"), **out)


def read_wav_mono_int(path):
    """Integer samples of a mono PCM file (16 / 24 bit) and the scale that turns them into the floats a host loads."""
    import struct
    with open(path, "rb") as f:
        d = f.read()
    pos = 12; fmt = None; data = None
    while pos + 8 <= len(d):
        cid = d[pos:pos + 4]; sz = struct.unpack("<I", d[pos + 4:pos + 8])[0]
        body = d[pos + 8:pos + 8 + sz]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
        elif cid == b"data":
            data = body
        pos += 8 + sz + (sz & 1)
    _, ch, sr, _, _, bits = fmt
    assert ch == 1
    if bits == 16:
        return np.frombuffer(data, "<i2").astype(np.int32), bits, sr
    b = np.frombuffer(data, np.uint8).reshape(-1, 3).astype(np.int32)
    v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
    return np.where(v & 0x800000, v - (1 << 24), v).astype(np.int32), bits, sr


def config1_full():
    """BASELINE config 1 at FULL length (the round-1 fixture held the first 32768 samples only): both reference
    WAVs, fft 1024 hop 256 rank 4, 100 iterations, seed 42.  The samples are stored losslessly as integers (the GPU box has
    no /root/reference); bases / activations and per-component resynthesis energies pin the oracle's full-length output."""
    for name in ("Tremblay-AaS-SynthTwoVoices-M", "Nicol-LoopE-M"):
        pcm, bits, sr = read_wav_mono_int(os.path.join(REF_AUDIO, name + ".wav"))
        a = (pcm.astype(np.float64) / float(1 << (bits - 1))).astype(np.float32)
        x, _ = read_wav_mono(os.path.join(REF_AUDIO, name + ".wav"))
        assert np.array_equal(a, x.astype(np.float32))
        r = co.bufnmf_channel(a, 1024, 1024, 256, 4, 100, 42, resynth=True)
        F = co.num_frames(a.size, 1024, 256)
        assert r["acts"].shape == (F, 4)
        np.savez_compressed(os.path.join(HERE, f"config1_{name.split('-')[0].lower()}.npz"),
                            pcm=pcm.astype(np.int32 if bits > 16 else np.int16), bits=bits, sr=sr, bases=r["bases"], acts=r["acts"],
                            resynth_energy=(r["resynth"].astype(np.float64) ** 2).sum(axis=1),
                            resynth_head=r["resynth"][:, :4096].astype(np.float32))
        print(name, a.size, "samples ->", F, "frames")


if __name__ == "__main__":
    co.build()
    if len(sys.argv) > 1 and sys.argv[1] == "config1":
        config1_full()
    elif len(sys.argv) > 1 and sys.argv[1] == "nmfcross":
        nmfcross()
    elif len(sys.argv) > 1 and sys.argv[1] == "spectral":
        spectral_epilogues()
    elif len(sys.argv) > 1 and sys.argv[1] == "stream":
        nmffilter_stream()
    elif len(sys.argv) > 1 and sys.argv[1] == "bufstft":
        bufstft()
    else:
        rng_kat(); fft_kat(); nmf_small(); bufnmf_wav(); nmffilter_stream(); bufstft(); config1_full(); nmfcross(); spectral_epilogues()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
