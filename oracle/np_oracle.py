"""Independent numpy (pocketfft, OpenBLAS) fp64 restatement of the reference path -- a second opinion on the C oracle.

TEST INFRASTRUCTURE ONLY (see flucoma_oracle.c header).  Written from the reference sources, not from the C oracle,
with different arithmetic building blocks (numpy.fft, BLAS matmul, numpy's own MT19937-64 is NOT used: the libstdc++
distribution is restated from its definition), so agreement between the two at ~1e-12 is meaningful.

Reference citations are relative to /root/reference/include/flucoma.
"""
from __future__ import annotations

import numpy as np

EPS = np.finfo(np.float64).eps  # algorithms/util/AlgorithmUtils.hpp:19


# ---- size rules: clients/common/ParameterTypes.hpp:295-313, clients/nrt/NMFClient.hpp:111-113 -----------------
def next_pow2(x: int) -> int:
    return 0 if x <= 0 else 1 << (int(x) - 1).bit_length()


def fft_params(win: int, hop: int = -1, fft: int = -1):
    f = next_pow2(win) if fft < 0 else fft
    h = hop if hop > 0 else win >> 1
    return win, h, f, (f >> 1) + 1


def num_frames(n: int, win: int, hop: int) -> int:
    return (n + win + hop - win) // hop


# ---- WindowFuncs.hpp:41-45 ---------------------------------------------------------------------------------------
def hann(size: int) -> np.ndarray:
    i = np.arange(size, dtype=np.float64)
    return 0.5 - 0.5 * np.cos((np.pi * 2 * i) / size)


# ---- STFT.hpp:90-108 + FFT.hpp:92-108 ----------------------------------------------------------------------------
def stft(audio: np.ndarray, win: int, fft: int, hop: int) -> np.ndarray:
    a = np.asarray(audio, dtype=np.float64)
    half = win // 2
    padded = np.zeros(a.size + win + hop)
    padded[half:half + a.size] = a
    F = (padded.size - win) // hop
    w = hann(win)
    idx = np.arange(F)[:, None] * hop + np.arange(win)[None, :]
    frames = padded[idx] * w
    S = np.fft.rfft(frames, n=fft, axis=1)  # zero-pads win -> fft, unnormalised forward DFT
    S[:, 0] = S[:, 0].real
    S[:, -1] = S[:, -1].real
    return S


def magnitude(S: np.ndarray) -> np.ndarray:  # STFT.hpp:61-66
    return np.abs(S)


# ---- STFT.hpp:178-199 + FFT.hpp:149-163 --------------------------------------------------------------------------
def istft(S: np.ndarray, win: int, fft: int, hop: int, n_out: int) -> np.ndarray:
    F = S.shape[0]
    half = win // 2
    osz = win + (F - 1) * hop + win + hop
    out = np.zeros(osz)
    nrm = np.zeros(osz)
    w = hann(win)
    S = np.array(S, dtype=np.complex128)
    S[:, 0] = S[:, 0].real   # htl::rifft ignores Im(DC)/Im(Nyquist): FFT.hpp:160
    S[:, -1] = S[:, -1].real
    y = np.fft.irfft(S, n=fft, axis=1) * fft  # = rifft result (fft * x)
    for i in range(F):
        out[i * hop:i * hop + win] += y[i, :win] * (1.0 / fft) * w
        nrm[i * hop:i * hop + win] += w * w
    out = out / np.maximum(nrm, EPS)
    return out[half:half + n_out]


# ---- EigenRandom.hpp:73-110 on libstdc++ -------------------------------------------------------------------------
def _mt19937_64(seed: int, count: int) -> np.ndarray:
    """Plain-Python MT19937-64 (Matsumoto & Nishimura 2004 reference constants)."""
    M = (1 << 64) - 1
    mt = [0] * 312
    mt[0] = seed & M
    for i in range(1, 312):
        mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & M
    out = np.empty(count, dtype=np.uint64)
    idx = 312
    for n in range(count):
        if idx >= 312:
            for i in range(312):
                x = (mt[i] & 0xFFFFFFFF80000000) | (mt[(i + 1) % 312] & 0x7FFFFFFF)
                mt[i] = mt[(i + 156) % 312] ^ (x >> 1) ^ (0xB5026F5AA96619E9 if x & 1 else 0)
            idx = 0
        x = mt[idx]
        idx += 1
        x ^= (x >> 29) & 0x5555555555555555
        x ^= (x << 17) & 0x71D67FFFEDA60000
        x ^= (x << 37) & 0xFFF7EEE000000000
        x ^= x >> 43
        out[n] = x & M
    return out


def random_uniform(seed: int, count: int) -> np.ndarray:
    raw = _mt19937_64(seed, count)
    r = raw.astype(np.float64) / 18446744073709551616.0  # generate_canonical<double,53> with a 64-bit engine
    r[r >= 1.0] = np.nextafter(1.0, 0.0)
    return r


def nmf_random_init(seed: int, bins: int, rank: int, frames: int):
    # column-major fill of (B x K) and (K x F); each restarts the stream: NMF.hpp:104-105, 116-117
    W = random_uniform(seed, bins * rank).reshape(rank, bins)      # W[k][b] = u[k*B + b]
    H = random_uniform(seed, rank * frames).reshape(frames, rank)  # H[f][k] = u[f*K + k]
    return W, H


# ---- NMF.hpp:91-134, 144-183 -------------------------------------------------------------------------------------
def nmf_process(X, rank, n_iter, update_w=True, update_h=True, seed=-1, W0=None, H0=None, progress=None):
    """X[F][B] -> W1[K][B], H1[F][K], V1[F][B], cancelled.  Written in the reference's own orientation (B x F)."""
    X = np.asarray(X, dtype=np.float64)
    F, B = X.shape
    V = X.T.copy()                                              # :125  (B x F)
    if W0 is None:
        W = nmf_random_init(seed, B, rank, F)[0].T.copy()       # B x K
    else:
        W = np.asarray(W0, dtype=np.float64).T.copy()           # :111
    if H0 is None:
        H = nmf_random_init(seed, B, rank, F)[1].T.copy()       # K x F
    else:
        H = np.asarray(H0, dtype=np.float64).T.copy()           # :123
    ones = np.ones_like(V)
    H = np.maximum(H, EPS)                                      # :150
    W = np.maximum(W, EPS)                                      # :151
    W = W / np.linalg.norm(W, axis=0, keepdims=True)            # :152
    H = H / np.linalg.norm(H, axis=1, keepdims=True)            # :153
    cancelled = False
    for it in range(n_iter):
        if update_w:
            V1 = np.maximum(W @ H, EPS)                         # :158
            wnum = (V / V1) @ H.T                               # :159
            wden = ones @ H.T                                   # :160
            W = W * wnum / np.maximum(wden, EPS)                # :161
            if W.max() > EPS:                                   # :162
                W = W / np.linalg.norm(W, axis=0, keepdims=True)
        V2 = np.maximum(W @ H, EPS)                             # :165
        if update_h:
            hnum = W.T @ (V / V2)                               # :168
            hden = W.T @ ones                                   # :169
            H = H * hnum / np.maximum(hden, EPS)                # :170
        if progress is not None and not progress(it + 1):       # :175-176
            cancelled = True
            break
    if not cancelled:
        V = W @ H                                               # :182
    return W.T.copy(), H.T.copy(), V.T.copy(), cancelled        # :127-133


def nmf_process_frame(x, W0, n_iter, seed):
    """NMF.hpp:45-89.  Returns (h, v, mutated W)."""
    x = np.asarray(x, dtype=np.float64)
    W = np.asarray(W0, dtype=np.float64).copy()                 # K x B
    K = W.shape[0]
    h = random_uniform(seed, K)                                 # :55
    v0 = np.maximum(x, EPS)                                     # :60
    W = np.maximum(W, EPS)                                      # :58
    h = np.maximum(h, EPS)                                      # :59
    W = W / np.linalg.norm(W, axis=1, keepdims=True)            # :63-64
    ones = np.ones_like(x)
    for _ in range(n_iter):
        v1 = np.maximum(W.T @ h, EPS)                           # :74-75
        r = v0 / v1                                             # :76
        h = h * (W @ r) / np.maximum(W @ ones, EPS)             # :77-79
    return h, W.T @ h, W


def nmf_estimate(W, H, idx):  # NMF.hpp:33-42
    return np.outer(np.asarray(H)[:, idx], np.asarray(W)[idx, :])


def ratio_mask(mixture, target, denominator):  # RatioMask.hpp:33-57, exponent 1
    mult = 1.0 / np.maximum(denominator, EPS)
    return mixture * np.minimum(1.0, target * mult)


# ---- NMFClient.hpp:233-335 ---------------------------------------------------------------------------------------
def bufnmf_channel(audio_f32, win, fft, hop, rank, iters, seed, bases_mode=0, bases_in=None, acts_mode=0, acts_in=None,
                   resynth=False):
    a = np.asarray(audio_f32, dtype=np.float32)
    n = a.size
    S = stft(a.astype(np.float64), win, fft, hop)
    M = magnitude(S)
    fix_w, fix_h = bases_mode == 2, acts_mode == 2
    needs = not (fix_w and fix_h)
    W0 = None if bases_mode == 0 else np.asarray(bases_in, dtype=np.float32).astype(np.float64)
    H0 = None if acts_mode == 0 else np.asarray(acts_in, dtype=np.float32).astype(np.float64)
    W, H, Vh, _ = nmf_process(M, rank, iters * int(needs), not fix_w, not fix_h, seed, W0, H0)
    bases = W.astype(np.float32)
    scale = np.float32(1.0 / H.max())
    acts = H.astype(np.float32) * scale
    rs = None
    if resynth:
        rs = np.zeros((rank, n), np.float32)
        for j in range(rank):
            est = nmf_estimate(W, H, j)
            rs[j] = istft(ratio_mask(S, est, Vh), win, fft, hop, n).astype(np.float32)
    return dict(bases=bases, acts=acts, resynth=rs, W=W, H=H, V=Vh)


# ---- streaming NMFFilter / NMFMatch: NMFFilterClient.hpp:98-117, NMFMatchClient.hpp:106-118 -----------------------
class _Source:
    """FluidSource.hpp:42-130: ring of size+host samples; pull(frame_time) reads `size` samples ending host-frame_time
    samples before the write head."""

    def __init__(self, size, host):
        self.size, self.host = size, host
        self.ring = np.zeros(size + host)
        self.counter = 0

    def push(self, block):
        n = self.ring.size
        idx = (self.counter + np.arange(block.size)) % n
        self.ring[idx] = block
        self.counter = (self.counter + block.size) % n

    def pull(self, blocksize, frame_time):
        n = self.ring.size
        offset = self.host - frame_time
        if offset > n:
            return np.zeros(blocksize)
        offset += blocksize
        start = self.counter - offset if offset <= self.counter else self.counter + n - offset
        return self.ring[(start + np.arange(blocksize)) % n].copy()


class _Sink:
    """FluidSink.hpp:40-130: accumulating ring; push adds at write head + frame_time, pull copies and clears."""

    def __init__(self, size, chans, host):
        self.ring = np.zeros((chans, size + host))
        self.counter = 0

    def push(self, x, frame_time):
        n = self.ring.shape[1]
        if frame_time + x.shape[1] > n:
            return
        idx = (self.counter + frame_time + np.arange(x.shape[1])) % n
        self.ring[:, idx] += x

    def pull(self, blocksize):
        n = self.ring.shape[1]
        idx = (self.counter + np.arange(blocksize)) % n
        out = self.ring[:, idx].copy()
        self.ring[:, idx] = 0
        self.counter = (self.counter + blocksize) % n
        return out


def nmffilter_stream(audio, win, fft, hop, W_in, n_iter, seed, host_size=64, want_out=True):
    """Drives the client exactly as a host would: `host_size` samples per call through BufferedProcess
    (BufferedProcess.hpp:49-73, 187-241).  Returns (out[K][n], acts[frames][K]); n is cut to whole host vectors."""
    a = np.asarray(audio, dtype=np.float64)
    W_in = np.asarray(W_in, dtype=np.float64)
    K = W_in.shape[0]
    nblocks = a.size // host_size
    n = nblocks * host_size
    src = _Source(win, host_size)
    snk = _Sink(win, K + 1, host_size)
    w = hann(win)
    frame_time = 0
    out = np.zeros((K, n))
    acts = []
    for blk in range(nblocks):
        src.push(a[blk * host_size:(blk + 1) * host_size])
        W = W_in.copy()                                                  # NMFFilterClient.hpp:94-95, per host vector
        while frame_time < host_size:                                    # BufferedProcess.hpp:57
            frame = src.pull(win, frame_time)
            S = np.fft.rfft(frame * w, n=fft)                            # STFT.hpp:110-118
            S[0] = S[0].real; S[-1] = S[-1].real
            h, est, W = nmf_process_frame(np.abs(S), W, n_iter, seed)    # mutates the filter copy (NMF.hpp:63-64)
            acts.append(h)
            frames_out = np.zeros((K + 1, win))
            if want_out:
                for k in range(K):
                    msp = ratio_mask(S, h[k] * W[k, :], est)             # :107-113
                    msp[0] = msp[0].real; msp[-1] = msp[-1].real
                    frames_out[k] = (np.fft.irfft(msp, n=fft) * fft)[:win] * (1.0 / fft) * w  # STFT.hpp:154-164
            frames_out[K] = w * w                                        # BufferedProcess.hpp:219-224
            snk.push(frames_out, frame_time)
            frame_time += hop
        frame_time -= host_size                                          # :72
        y = snk.pull(host_size)
        g = y[K]
        for k in range(K):                                               # :231-237
            x = y[k]
            nz = x != 0
            x[nz] = x[nz] / np.where(g[nz] > 0, g[nz], 1.0)
            out[k, blk * host_size:(blk + 1) * host_size] = x
    return out, np.array(acts)


# ---------------------------------------------------------------------------------------------------------------------
# BufSTFT (clients/nrt/BufSTFTClient.hpp:82-190, 192-279), independent numpy restatement
def bufstft_sizes(win, hop, mode, invert, count):
    hop = hop if hop > 0 else win >> 1
    pad = (0, win >> 1, win - hop)[mode]                                  # ParameterTypes.hpp:315-323
    if not invert:
        padded = count + 2 * pad                                          # :121-124
        if mode == 2:
            padded = -(-padded // hop) * hop                              # :126-128
        if padded < win:
            raise ValueError("input shorter than one window")
        return pad, 1 + (padded - win) // hop                             # :130-131
    return pad, (count - 1) * hop + win - pad                             # :241-242


def bufstft_fwd(audio_f32, win, fft, hop, mode=1):
    a = np.asarray(audio_f32, np.float32).astype(np.float64)
    pad, hops = bufstft_sizes(win, hop, mode, False, a.size)
    padded = np.zeros(max((hops - 1) * hop + win, pad + a.size))
    padded[pad:pad + a.size] = a
    w = hann(win)
    frames = np.stack([padded[i * hop:i * hop + win] * w for i in range(hops)])
    S = np.fft.rfft(frames, n=fft, axis=1)
    S[:, 0] = S[:, 0].real; S[:, -1] = S[:, -1].real                      # FFT.hpp:99-101
    return np.abs(S).astype(np.float32), np.angle(S).astype(np.float32)


def bufstft_inv(mag_f32, phase_f32, win, fft, hop, mode=1):
    m = np.asarray(mag_f32, np.float32).astype(np.float64); ph = np.asarray(phase_f32, np.float32).astype(np.float64)
    frames = m.shape[0]
    pad, n_out = bufstft_sizes(win, hop, mode, True, frames)
    S = m * np.exp(1j * ph)
    S[:, 0] = S[:, 0].real; S[:, -1] = S[:, -1].real                      # FFT.hpp:151-158 packs real DC / Nyquist
    y = np.fft.irfft(S, n=fft, axis=1)[:, :win]                            # = rifft * (1/fft)
    w = hann(win)
    plen = (frames - 1) * hop + win
    acc = np.zeros(plen); nrm = np.zeros(plen)
    for i in range(frames):
        acc[i * hop:i * hop + win] += y[i] * w
        nrm[i * hop:i * hop + win] += w * w
    out = acc / np.maximum(nrm, EPS)
    return out[pad:pad + n_out].astype(np.float32)


# ---- NMFCross.hpp:60-185, GriffinLim.hpp:29-54, NMFCrossClient.hpp:85-185 -----------------------------------------
def nmfcross_process(X, W0, n_iter, r, p, c, seed):
    """X[F][B] target magnitudes, W0[R][B] source magnitudes -> H[F][R]; written in the reference's orientation
    (W: B x R, H: R x F) with numpy building blocks (sliding windows, argsort, trace-like diagonal sums, BLAS)."""
    V = np.asarray(X, dtype=np.float64).T                      # B x F
    W = np.maximum(np.asarray(W0, dtype=np.float64).T, EPS)    # B x R   (:156)
    R, F = W.shape[1], V.shape[1]
    H = random_uniform(seed, R * F).reshape(F, R).T.copy()     # column-major fill of R x F (:72-73)
    energy = (W ** 2).sum(axis=0)                              # :159
    hden = np.maximum(W.sum(axis=0), EPS)[:, None]
    for i in range(n_iter):
        factor = 1.0 - float((i + 1) // n_iter)                # integer division as written (:119, :136)
        # temporal sparseness (:104-127)
        half = (r - 1) // 2
        pad = np.zeros((R, F + r)); pad[:, half:half + F] = H
        win = np.lib.stride_tricks.sliding_window_view(pad, r, axis=1)[:, :F, :]
        keep = win.argmax(axis=2) == half                      # first maximum, like Eigen's maxCoeff(&index)
        H = np.where(keep, H, H * factor)
        # polyphony (:130-143)
        score = H * energy[:, None]
        out = H * factor
        top = np.argsort(-score, axis=0, kind="stable")[:p, :]
        cols = np.arange(F)[None, :]
        out[top, cols] = H[top, cols]
        H = out
        # continuity (:86-102): sums along the diagonal
        half = (c - 1) // 2
        pad = np.zeros((R + c, F + c)); pad[half:half + R, half:half + F] = H
        H = sum(pad[d:d + R, d:d + F] for d in range(c))
        # KL update of H, W fixed (:167-170)
        V2 = np.maximum(W @ H, EPS)
        H = H * (W.T @ (V / V2)) / hden
    return H.T.copy()


def griffinlim(spec, n_samples, n_iter, win, fft, hop, seed):
    S = np.asarray(spec, dtype=np.complex128)
    F, B = S.shape
    mag = np.abs(S)
    theta = random_uniform(seed, F * B) * (2 * np.pi - 0.0) + 0.0     # uniform_real_distribution(0, 2 pi)
    phase = np.exp(1j * theta).reshape(B, F).T                        # column-major fill of an F x B array
    est = np.zeros((F, B), np.complex128)
    for _ in range(n_iter):
        prev = est
        est = stft(istft(mag * phase, win, fft, hop, n_samples), win, fft, hop)
        phase = est - (0.9 / 1.9) * prev
        phase = phase / (np.abs(phase) + EPS)
    return mag * phase


def bufnmfcross(source_f32, target_f32, win, fft, hop, time_sparsity=7, polyphony=11, continuity=7, iters=50, seed=-1,
                gl_iters=50):
    s = np.asarray(source_f32, dtype=np.float32).astype(np.float64)
    t = np.asarray(target_f32, dtype=np.float32).astype(np.float64)
    S = stft(s, win, fft, hop); T = stft(t, win, fft, hop)
    H = nmfcross_process(np.abs(T), np.abs(S), iters, time_sparsity, min(S.shape[0], polyphony), continuity, seed)
    res = griffinlim(H @ S, t.size, gl_iters, win, fft, hop, seed)
    return istft(res, win, fft, hop, t.size).astype(np.float32), H


# ---- MelBands.hpp:35-101 -------------------------------------------------------------------------------------------
def melbands(mags, lo, hi, n_bands, sample_rate, win, mag_norm=True, use_power=False, log_output=False):
    """mags[F][B] -> bands[F][nBands] (vectorised over frames; numpy.linspace == Eigen's LinSpaced for increasing ranges)."""
    M = np.array(mags, dtype=np.float64)
    B = M.shape[1]
    fft = 2 * (B - 1)
    s1 = 1.0 / (win / 4.0)
    s2 = 1.0 / (2.0 * fft / win)
    mel = np.linspace(1127.01048 * np.log(lo / 700.0 + 1.0), 1127.01048 * np.log(hi / 700.0 + 1.0), n_bands + 2)
    mel = 700.0 * (np.exp(mel / 1127.01048) - 1.0)
    freqs = np.linspace(0.0, sample_rate / 2.0, B)
    d = np.abs(mel[:-1] - mel[1:])
    ramps = mel[:, None] - freqs[None, :]
    filt = np.maximum(np.minimum(-ramps[:-2] / d[:-1, None], ramps[2:] / d[1:, None]), 0.0)
    if mag_norm:
        M = M * s1
    energy = M.sum(axis=1) * s2
    if use_power:
        M = M * M
    out = M @ filt.T
    if mag_norm:
        out = out * energy[:, None] / np.maximum(EPS, out.sum(axis=1))[:, None]
    if log_output:
        out = 20.0 * np.log10(np.maximum(out, EPS))
    return out


# ---- HPSS.hpp:47-162 (+ MedianFilter.hpp:36-57), closed form of the streaming recursion -------------------------------
def hpss(spec, v_size, h_size, mode=0, h_thresh=(0.0, 1.0, 1.0, 1.0), p_thresh=(0.0, 1.0, 1.0, 1.0)):
    """spec[F][B] -> out[3][F][B].  Instead of replaying the delay lines this writes down what they hold:
       output frame t carries input frame u = t - (hSize - 1);
       v0(t)[b] = median(|X[u]|[b .. b + vSize - 1]) (zeros past the last bin)        -- the causal filter over the padded frame
       h0(t)[b] = median(|X[w - hSize + 1 .. w]|[b]) with w = t - (h2 + 1) (zeros before the stream) -- column h2 + 1, shifted out later."""
    S = np.asarray(spec, dtype=np.complex128)
    F, B = S.shape
    A = np.abs(S)
    h2 = (h_size - 1) // 2
    vpad = np.concatenate([A, np.zeros((F, v_size - 1))], axis=1)
    vmed = np.median(np.lib.stride_tricks.sliding_window_view(vpad, v_size, axis=1), axis=2)          # [F][B]
    hpad = np.concatenate([np.zeros((h_size - 1, B)), A], axis=0)
    hmed = np.median(np.lib.stride_tricks.sliding_window_view(hpad, h_size, axis=0), axis=2)          # [F][B], causal in time

    def delayed(X, d):
        Y = np.zeros_like(X)
        if d < F:
            Y[d:] = X[:F - d]
        return Y
    X0 = delayed(S, h_size - 1)
    v0 = delayed(vmed, h_size - 1)
    h0 = delayed(hmed, h2 + 1)

    def threshold(x1, y1, x2, y2):
        th = np.ones(B)
        ks, ke = int(np.floor(x1 * B)), int(np.floor(x2 * B))
        th[:ks] = 10.0 ** (y1 / 20.0)
        if ke > ks:
            th[ks:ke] = 10.0 ** (np.linspace(y1, y2, ke - ks) / 20.0)
        th[ke:] = 10.0 ** (y2 / 20.0)
        return th
    with np.errstate(divide="ignore", invalid="ignore"):
        if mode == 0:
            mult = 1.0 / np.maximum(h0 + v0, EPS)
            hm, pm, rm = h0 * mult, v0 * mult, np.zeros_like(h0)
        elif mode == 1:
            hm = ((h0 / v0) > threshold(*h_thresh)[None, :]).astype(np.float64)
            pm, rm = 1.0 - hm, np.zeros_like(h0)
        else:
            hm = ((h0 / v0) > threshold(*h_thresh)[None, :]).astype(np.float64)
            pm = ((v0 / h0) > threshold(*p_thresh)[None, :]).astype(np.float64)
            rm = (1.0 - hm) * (1.0 - pm)
            nrm = np.maximum(1.0 / (hm + pm + rm), EPS)
            hm, pm, rm = hm * nrm, pm * nrm, rm * nrm
    return np.stack([X0 * np.minimum(hm, 1.0), X0 * np.minimum(pm, 1.0), X0 * np.minimum(rm, 1.0)])


# ---- NNDSVD.hpp:30-131 --------------------------------------------------------------------------------------------
def nndsvd(X, min_rank=1, max_rank=200, amount=0.5, method=0, seed=-1):
    """X[F][B] -> (k, W[max_rank][B], H[F][max_rank], s).  LAPACK SVD of X^T; the sign of each (u, v) pair is fixed by the
    convention of the C restatement (largest-magnitude entry of u positive) -- see the note there on :84."""
    X = np.asarray(X, dtype=np.float64)
    F, B = X.shape
    U, s, Vt = np.linalg.svd(X.T, full_matrices=False)      # U: B x r, Vt: r x F
    sign = np.sign(U[np.abs(U).argmax(axis=0), np.arange(U.shape[1])])
    sign[sign == 0] = 1.0
    U = U * sign[None, :]; Vt = Vt * sign[:, None]
    if amount == 0:
        k = min_rank
    else:
        k, cur, total = 0, 0.0, s.sum()
        while cur / total < amount and k < s.size:
            cur += s[k]; k += 1
    k = min(max(k, min_rank), max_rank, s.size)
    WT = np.zeros((B, max_rank)); HT = np.zeros((max_rank, F))
    if method == 0:
        WT[:, :k] = np.abs(U[:, :k]); HT[:k] = np.abs(s[:k, None] * Vt[:k])
    else:
        WT[:, 0] = np.abs(U[:, 0]); HT[0] = np.sqrt(s[0]) * np.abs(Vt[0])
        for j in range(1, k):
            x, y = U[:, j], Vt[j]
            xP, yP, xN, yN = np.maximum(x, 0), np.maximum(y, 0), np.abs(np.minimum(x, 0)), np.abs(np.minimum(y, 0))
            xPn, yPn, xNn = np.linalg.norm(xP), np.linalg.norm(yP), np.linalg.norm(xN)
            yNn = xNn                                        # :84 as written
            mP, mN = xPn * yPn, xNn * yNn
            if mP > mN:
                u, v, sigma = xP / xPn, yP / yPn, mP
            else:
                u, v, sigma = xN / xNn, yN / yNn, mN
            WT[:, j] = u; HT[j] = np.sqrt(s[j] * sigma) * v
        mean = X.mean()
        if method == 1:
            lo, hi = EPS, mean * 0.001
            Wr = (random_uniform(seed, B * max_rank) * (hi - lo) + lo).reshape(max_rank, B).T
            Hr = (random_uniform(seed, max_rank * F) * (hi - lo) + lo).reshape(F, max_rank).T
            WT = np.where(WT < EPS, Wr, WT); HT = np.where(HT < EPS, Hr, HT)
        elif method == 2:
            WT = np.where(WT < EPS, mean, WT); HT = np.where(HT < EPS, mean, HT)
    return k, WT.T.copy(), HT.T.copy(), s
