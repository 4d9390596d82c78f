"""ctypes wrapper over oracle/libflucoma_oracle.so (the CPU fp64 restatement of the reference path).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (flucoma-core_b200/) must never import this module.
Function-by-function reference citations live in flucoma_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i64 = C.c_int64
_pd = C.POINTER(C.c_double)
_pf = C.POINTER(C.c_float)
_pi64 = C.POINTER(C.c_int64)
PROGRESS_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64)


def build(force: bool = False, march: str = "x86-64-v3", out: str | None = None) -> str:
    """Compile the oracle with the system gcc (never $CC: the image exports a wrapper without libgomp)."""
    out = out or os.path.join(_HERE, "libflucoma_oracle.so")
    src = os.path.join(_HERE, "flucoma_oracle.c")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    base = ["/usr/bin/gcc", "-O3", f"-march={march}", "-fPIC", "-std=c11", "-fvisibility=hidden", "-shared", "-o", out, src, "-lm"]
    try:
        subprocess.run(base[:3] + ["-fopenmp"] + base[3:], check=True, capture_output=True)
    except (subprocess.CalledProcessError, FileNotFoundError):
        subprocess.run(base[:3] + ["-fopenmp-simd"] + base[3:], check=True, capture_output=True)
    return out


def lib(path: str | None = None):
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or os.path.join(_HERE, "libflucoma_oracle.so")
    if not os.path.exists(p):
        build(out=p)
    L = C.CDLL(p)
    L.fo_next_pow2.restype = _i64
    L.fo_next_pow2.argtypes = [_i64]
    L.fo_stft_num_frames.restype = _i64
    L.fo_stft_num_frames.argtypes = [_i64, _i64, _i64]
    L.fo_fft_params.argtypes = [_i64, _i64, _i64, _pi64, _pi64, _pi64, _pi64]
    L.fo_hann.argtypes = [_i64, _pd]
    L.fo_rfft.argtypes = [_pd, _i64, _i64, _pd]
    L.fo_irfft_unnorm.argtypes = [_pd, _i64, _pd]
    L.fo_stft.argtypes = [_pd, _i64, _i64, _i64, _i64, _pd]
    L.fo_stft_frame.argtypes = [_pd, _i64, _i64, _pd]
    L.fo_magnitude.argtypes = [_pd, _i64, _pd]
    L.fo_istft.argtypes = [_pd, _i64, _i64, _i64, _i64, _pd, _i64]
    L.fo_random_uniform.argtypes = [_i64, _i64, _pd]
    L.fo_nmf_random_init.argtypes = [_i64, _i64, _i64, _i64, _pd, _pd]
    L.fo_nmf_process.restype = C.c_int
    L.fo_nmf_process.argtypes = [_pd, _i64, _i64, _i64, _i64, C.c_int, C.c_int, _i64, _pd, _pd, _pd, _pd, _pd,
                                 C.c_int, C.c_void_p, C.c_void_p]
    L.fo_nmf_process_frame.argtypes = [_pd, _pd, _i64, _i64, _i64, _i64, _pd, _pd]
    L.fo_nmf_estimate.argtypes = [_pd, _pd, _i64, _i64, _i64, _i64, _pd]
    L.fo_ratio_mask.argtypes = [_pd, _pd, _pd, _i64, _pd]
    L.fo_bufnmf_channel.restype = C.c_int
    L.fo_bufnmf_channel.argtypes = [_pf, _i64, _i64, _i64, _i64, _i64, _i64, _i64, C.c_int, _pf, C.c_int, _pf, _pf,
                                    _pf, _pf, C.c_int, _pd, _pd, _pd]
    L.fo_bufnmf_batch.restype = C.c_int
    L.fo_bufnmf_batch.argtypes = [_pf, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _pi64, _pf, _pf, _pf, C.c_int,
                                  C.c_int]
    L.fo_nmfmatch_frames.argtypes = [_pd, _i64, _pd, _i64, _i64, _i64, _i64, _pd, C.c_int]
    L.fo_stream_frames_mag.argtypes = [_pd, _i64, _i64, _i64, _i64, _i64, _pd, _pd]
    L.fo_nmffilter_stream.argtypes = [_pd, _i64, _i64, _i64, _i64, _pd, _i64, _i64, _i64, _pd, _pd]
    L.fo_nmfcross_process.restype = C.c_int
    L.fo_nmfcross_process.argtypes = [_pd, _i64, _i64, _pd, _i64, _i64, _i64, _i64, _i64, _i64, _pd, C.c_void_p, C.c_void_p]
    L.fo_nmfcross_synthesize.argtypes = [_pd, _pd, _i64, _i64, _i64, _pd]
    L.fo_griffinlim.argtypes = [_pd, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64]
    L.fo_bufnmfcross.restype = C.c_int
    L.fo_bufnmfcross.argtypes = [_pf, _i64, _pf, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _pf, _pd]
    L.fo_melbands.argtypes = [_pd, _i64, _i64, C.c_double, C.c_double, _i64, C.c_double, _i64, C.c_int, C.c_int, C.c_int, _pd]
    L.fo_melbands_init.argtypes = [C.c_double, C.c_double, _i64, _i64, C.c_double, _i64, _pd, _pd, _pd]
    L.fo_hpss.argtypes = [_pd, _i64, _i64, _i64, _i64, _i64] + [C.c_double] * 8 + [_pd]
    L.fo_nndsvd.restype = _i64
    L.fo_nndsvd.argtypes = [_pd, _i64, _i64, _i64, _i64, C.c_double, _i64, _i64, _pd, _pd, _pd]
    L.fo_num_threads.restype = C.c_int
    L.fo_bufstft_sizes.restype = C.c_int
    L.fo_bufstft_sizes.argtypes = [_i64, _i64, _i64, C.c_int, _i64, _pi64, _pi64]
    L.fo_bufstft_fwd.restype = C.c_int
    L.fo_bufstft_fwd.argtypes = [_pf, _i64, _i64, _i64, _i64, _i64, _pf, _pf]
    L.fo_bufstft_inv.restype = C.c_int
    L.fo_bufstft_inv.argtypes = [_pf, _pf, _i64, _i64, _i64, _i64, _i64, _pf]
    if path is None:
        _LIB = L
    return L


def _d(a):
    return None if a is None else a.ctypes.data_as(_pd)


def _f(a):
    return None if a is None else a.ctypes.data_as(_pf)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def fft_params(win, hop=-1, fft=-1):
    o = [C.c_int64() for _ in range(4)]
    lib().fo_fft_params(win, hop, fft, *[C.byref(x) for x in o])
    return tuple(int(x.value) for x in o)  # win, hop, fft, bins


def num_frames(n, win, hop):
    return int(lib().fo_stft_num_frames(n, win, hop))


def hann(size):
    w = np.empty(size, np.float64)
    lib().fo_hann(size, _d(w))
    return w


def rfft(x, fft):
    x = _c64(x)
    out = np.empty(2 * (fft // 2 + 1), np.float64)
    lib().fo_rfft(_d(x), x.size, fft, _d(out))
    return out.view(np.complex128)


def irfft_unnorm(X, fft):
    X = np.ascontiguousarray(X, dtype=np.complex128)
    y = np.empty(fft, np.float64)
    lib().fo_irfft_unnorm(_d(X.view(np.float64)), fft, _d(y))
    return y


def stft(audio, win, fft, hop):
    a = _c64(audio)
    F = num_frames(a.size, win, hop)
    B = fft // 2 + 1
    out = np.empty((F, B), np.complex128)
    lib().fo_stft(_d(a), a.size, win, fft, hop, _d(out.view(np.float64)))
    return out


def stft_frame(frame, fft):
    a = _c64(frame)
    out = np.empty(fft // 2 + 1, np.complex128)
    lib().fo_stft_frame(_d(a), a.size, fft, _d(out.view(np.float64)))
    return out


def magnitude(spec):
    s = np.ascontiguousarray(spec, dtype=np.complex128)
    m = np.empty(s.shape, np.float64)
    lib().fo_magnitude(_d(s.view(np.float64)), s.size, _d(m))
    return m


def istft(spec, win, fft, hop, n_out):
    s = np.ascontiguousarray(spec, dtype=np.complex128)
    y = np.empty(n_out, np.float64)
    lib().fo_istft(_d(s.view(np.float64)), s.shape[0], win, fft, hop, _d(y), n_out)
    return y


def random_uniform(seed, count):
    o = np.empty(count, np.float64)
    lib().fo_random_uniform(seed, count, _d(o))
    return o


def nmf_random_init(seed, bins, rank, frames):
    W = np.empty((rank, bins), np.float64)
    H = np.empty((frames, rank), np.float64)
    lib().fo_nmf_random_init(seed, bins, rank, frames, _d(W), _d(H))
    return W, H


def nmf_process(X, rank, n_iter, update_w=True, update_h=True, seed=-1, W0=None, H0=None, faithful=False,
                progress=None):
    """NMF::process.  X[F][B] -> (W[K][B], H[F][K], V[F][B], cancelled)."""
    X = _c64(X)
    F, B = X.shape
    W0 = None if W0 is None else _c64(W0)
    H0 = None if H0 is None else _c64(H0)
    if W0 is not None:
        assert W0.shape == (rank, B)
    if H0 is not None:
        assert H0.shape == (F, rank)
    W = np.empty((rank, B)); H = np.empty((F, rank)); V = np.empty((F, B))
    cb = PROGRESS_CB(lambda user, it: int(bool(progress(it)))) if progress else None
    r = lib().fo_nmf_process(_d(X), F, B, rank, n_iter, int(update_w), int(update_h), seed, _d(W0), _d(H0), _d(W),
                             _d(H), _d(V), int(faithful), C.cast(cb, C.c_void_p) if cb else None, None)
    return W, H, V, bool(r)


def nmf_process_frame(x, W0, n_iter, seed, want_v=True):
    """NMF::processFrame.  Returns (h[K], v[B] or None, mutated W[K][B])."""
    x = _c64(x)
    W = _c64(W0).copy()
    K, B = W.shape
    h = np.empty(K); v = np.empty(B) if want_v else None
    lib().fo_nmf_process_frame(_d(x), _d(W), B, K, n_iter, seed, _d(h), _d(v))
    return h, v, W


def nmf_estimate(W, H, idx):
    W = _c64(W); H = _c64(H)
    K, B = W.shape; F = H.shape[0]
    E = np.empty((F, B))
    lib().fo_nmf_estimate(_d(W), _d(H), F, B, K, idx, _d(E))
    return E


def ratio_mask(mixture, target, denominator):
    m = np.ascontiguousarray(mixture, dtype=np.complex128)
    t = _c64(target); d = _c64(denominator)
    out = np.empty(m.shape, np.complex128)
    lib().fo_ratio_mask(_d(m.view(np.float64)), _d(t), _d(d), m.size, _d(out.view(np.float64)))
    return out


def bufnmf_channel(audio, win, fft, hop, rank, iters, seed, bases_mode=0, bases_in=None, acts_mode=0, acts_in=None,
                   resynth=False, faithful=False, debug=False):
    """One BufNMF channel.  Returns dict(bases[K][B] f32, acts[F][K] f32, resynth[K][n] f32|None, W,H,V f64 if debug)."""
    a = np.ascontiguousarray(audio, dtype=np.float32)
    n = a.size; B = fft // 2 + 1; F = num_frames(n, win, hop)
    bi = None if bases_in is None else np.ascontiguousarray(bases_in, dtype=np.float32)
    ai = None if acts_in is None else np.ascontiguousarray(acts_in, dtype=np.float32)
    bases = np.zeros((rank, B), np.float32); acts = np.zeros((F, rank), np.float32)
    rs = np.zeros((rank, n), np.float32) if resynth else None
    dW = np.empty((rank, B)) if debug else None
    dH = np.empty((F, rank)) if debug else None
    dV = np.empty((F, B)) if debug else None
    lib().fo_bufnmf_channel(_f(a), n, win, fft, hop, rank, iters, seed, bases_mode, _f(bi), acts_mode, _f(ai),
                            _f(bases), _f(acts), _f(rs), int(faithful), _d(dW), _d(dH), _d(dV))
    return dict(bases=bases, acts=acts, resynth=rs, W=dW, H=dH, V=dV)


def bufnmf_batch(audio, win, fft, hop, rank, iters, seeds, resynth=False, faithful=True, threads=0, lib_path=None):
    a = np.ascontiguousarray(audio, dtype=np.float32)
    batch, n = a.shape
    B = fft // 2 + 1; F = num_frames(n, win, hop)
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    bases = np.zeros((batch, rank, B), np.float32); acts = np.zeros((batch, F, rank), np.float32)
    rs = np.zeros((batch, rank, n), np.float32) if resynth else None
    lib(lib_path).fo_bufnmf_batch(_f(a), batch, n, win, fft, hop, rank, iters, seeds.ctypes.data_as(_pi64), _f(bases),
                                  _f(acts), _f(rs), int(faithful), threads)
    return bases, acts, rs


def nmfmatch_frames(mags, W, n_iter, seed, threads=0, lib_path=None):
    m = _c64(mags); W = _c64(W)
    K, B = W.shape
    acts = np.empty((m.shape[0], K))
    lib(lib_path).fo_nmfmatch_frames(_d(m), m.shape[0], _d(W), B, K, n_iter, seed, _d(acts), threads)
    return acts


def stream_frames_mag(audio, win, fft, hop, nframes, want_spec=False):
    a = _c64(audio)
    B = fft // 2 + 1
    mags = np.empty((nframes, B))
    spec = np.empty((nframes, B), np.complex128) if want_spec else None
    lib().fo_stream_frames_mag(_d(a), a.size, win, fft, hop, nframes, _d(mags),
                               _d(spec.view(np.float64)) if want_spec else None)
    return (mags, spec) if want_spec else mags


def nmffilter_stream(audio, win, fft, hop, W, n_iter, seed, want_out=True):
    """NMFFilter over a stream from reset state: (out[K][n] or None, acts[ceil(n/hop)][K])."""
    a = _c64(audio); W = _c64(W)
    K, B = W.shape
    assert B == fft // 2 + 1
    nframes = (a.size + hop - 1) // hop
    out = np.empty((K, a.size)) if want_out else None
    acts = np.empty((nframes, K))
    lib().fo_nmffilter_stream(_d(a), a.size, win, fft, hop, _d(W), K, n_iter, seed, _d(out) if want_out else None,
                              _d(acts))
    return out, acts


def bufstft_sizes(win, hop, mode, invert, count):
    """(padding, numHops) for the forward transform of `count` samples, (padding, samples) for the inverse of `count` frames."""
    pad, out = _i64(), _i64()
    if lib().fo_bufstft_sizes(win, hop, mode, int(invert), count, C.byref(pad), C.byref(out)) != 0:
        raise ValueError("bufstft_sizes: input shorter than one window or bad arguments")
    return pad.value, out.value


def bufstft_fwd(audio, win, fft, hop, mode=1):
    """BufSTFT processFwd of one mono float32 buffer -> (mag [hops][bins], phase [hops][bins]) float32."""
    a = np.ascontiguousarray(audio, dtype=np.float32)
    _, hops = bufstft_sizes(win, hop, mode, False, a.size)
    bins = fft // 2 + 1
    mag = np.empty((hops, bins), np.float32); ph = np.empty((hops, bins), np.float32)
    assert lib().fo_bufstft_fwd(_f(a), a.size, win, fft, hop, mode, _f(mag), _f(ph)) == 0
    return mag, ph


def bufstft_inv(mag, phase, win, fft, hop, mode=1):
    """BufSTFT processInverse: (mag, phase) [frames][bins] float32 -> float32 audio."""
    m = np.ascontiguousarray(mag, dtype=np.float32); p = np.ascontiguousarray(phase, dtype=np.float32)
    frames = m.shape[0]
    _, n_out = bufstft_sizes(win, hop, mode, True, frames)
    out = np.empty(n_out, np.float32)
    assert lib().fo_bufstft_inv(_f(m), _f(p), frames, win, fft, hop, mode, _f(out)) == 0
    return out


def nmfcross_process(X, W0, n_iter, r, p, c, seed):
    """NMFCross::process: X[F][B] target magnitudes, W0[R][B] source magnitudes -> H[F][R]."""
    X = _c64(X); W0 = _c64(W0)
    F, B = X.shape
    R = W0.shape[0]
    H = np.empty((F, R))
    lib().fo_nmfcross_process(_d(X), F, B, _d(W0), R, n_iter, r, p, c, seed, _d(H), None, None)
    return H


def nmfcross_synthesize(H, S):
    H = _c64(H); S = np.ascontiguousarray(S, dtype=np.complex128)
    F, R = H.shape
    B = S.shape[1]
    out = np.empty((F, B), np.complex128)
    lib().fo_nmfcross_synthesize(_d(H), _d(S.view(np.float64)), F, R, B, _d(out.view(np.float64)))
    return out


def griffinlim(spec, n_samples, n_iter, win, fft, hop, seed):
    S = np.ascontiguousarray(spec, dtype=np.complex128).copy()
    F, B = S.shape
    lib().fo_griffinlim(_d(S.view(np.float64)), F, B, n_samples, n_iter, win, fft, hop, seed)
    return S


def bufnmfcross(source, target, win, fft, hop, time_sparsity=7, polyphony=11, continuity=7, iters=50, seed=-1, gl_iters=50):
    """BufNMFCross client (NMFCrossClient.hpp:85-185) on mono float32 buffers -> (out float32 [n_target], H [Ft][Fs])."""
    s = np.ascontiguousarray(source, dtype=np.float32); t = np.ascontiguousarray(target, dtype=np.float32)
    Fs, Ft = num_frames(s.size, win, hop), num_frames(t.size, win, hop)
    out = np.empty(t.size, np.float32)
    H = np.empty((Ft, Fs))
    rc = lib().fo_bufnmfcross(_f(s), s.size, _f(t), t.size, win, fft, hop, time_sparsity, polyphony, continuity, iters, seed,
                              gl_iters, _f(out), _d(H))
    if rc != 0:
        raise ValueError("bufnmfcross: invalid arguments (empty buffer, or sparsity / continuity larger than the target)")
    return out, H


def melbands(mags, lo, hi, n_bands, sample_rate, win, mag_norm=True, use_power=False, log_output=False):
    """MelBands::init + processFrame per frame: mags[F][B] -> bands[F][nBands]."""
    m = _c64(mags)
    F, B = m.shape
    out = np.empty((F, n_bands))
    lib().fo_melbands(_d(m), F, B, lo, hi, n_bands, sample_rate, win, int(mag_norm), int(use_power), int(log_output), _d(out))
    return out


def melbands_filters(lo, hi, n_bands, n_bins, sample_rate, win):
    filt = np.empty((n_bands, n_bins)); s1 = np.empty(1); s2 = np.empty(1)
    lib().fo_melbands_init(lo, hi, n_bands, n_bins, sample_rate, win, _d(filt), _d(s1), _d(s2))
    return filt, float(s1[0]), float(s2[0])


def hpss(spec, v_size, h_size, mode=0, h_thresh=(0.0, 1.0, 1.0, 1.0), p_thresh=(0.0, 1.0, 1.0, 1.0)):
    """HPSS::processFrame over spec[F][B] from init state -> out[3][F][B] complex (harmonic, percussive, residual)."""
    S = np.ascontiguousarray(spec, dtype=np.complex128)
    F, B = S.shape
    out = np.empty((3, F, B), np.complex128)
    lib().fo_hpss(_d(S.view(np.float64)), F, B, v_size, h_size, mode, *[float(x) for x in h_thresh], *[float(x) for x in p_thresh],
                  _d(out.view(np.float64)))
    return out


def nndsvd(X, min_rank=1, max_rank=200, amount=0.5, method=0, seed=-1):
    """NNDSVD::process: X[F][B] -> (k, W[max_rank][B], H[F][max_rank], singular values)."""
    X = _c64(X)
    F, B = X.shape
    W = np.zeros((max_rank, B)); H = np.zeros((F, max_rank)); sv = np.zeros(min(F, B))
    k = lib().fo_nndsvd(_d(X), F, B, min_rank, max_rank, float(amount), method, seed, _d(W), _d(H), _d(sv))
    return int(k), W, H, sv


def num_threads():
    return int(lib().fo_num_threads())
