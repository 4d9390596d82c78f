"""ctypes binding of libflucoma_b200.so (include/flucoma_b200.h) for tests, smoke() and bench.py.

This is deliberately thin: every method forwards to one C-ABI entry point.  There is NO CPU fallback -- if the shared
library is missing or no CUDA device is present, calls raise.  Host arrays are numpy; device arrays are torch CUDA
tensors (torch is only used for device memory; the library never sees a torch type, only raw pointers).

Names mirror the reference interface the entry points replace:
  Plan.stft / Plan.magnitude      algorithm::STFT::process / STFT::magnitude   (algorithms/public/STFT.hpp:90-108,61-66)
  Plan.istft                      algorithm::ISTFT::process                    (STFT.hpp:178-199)
  Plan.nmf_process                algorithm::NMF::process                      (algorithms/public/NMF.hpp:91-134)
  Plan.nmf_process_frames         algorithm::NMF::processFrame over frames     (NMF.hpp:45-89, NMFMatchClient.hpp:106-118)
  Plan.bufnmf                     client::bufnmf::NMFClient::process           (clients/nrt/NMFClient.hpp:96-337)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("FB200_LIB") or os.path.join(_PKG, "lib", "libflucoma_b200.so")  # FB200_LIB: developer A/B builds

F32, F64 = 0, 1
HOST, DEVICE = 0, 1
OK, WARN_NO_WORK, CANCELLED = 0, 1, 2
BACKEND_AUTO, BACKEND_SIMT, BACKEND_TCGEN05, BACKEND_TCGEN05_STREAMED = 0, 1, 2, 3
PROGRESS_ASYNC = -1  # fb200_progress_fn: never interrupt the device loop, poll its pass counters

PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64)


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("win", C.c_int32), ("hop", C.c_int32),
                ("fft", C.c_int32), ("max_rank", C.c_int32), ("max_batch", C.c_int64), ("max_samples", C.c_int64),
                ("backend", C.c_int32), ("reserved", C.c_int32)]


class NmfArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("dtype", C.c_int32), ("mem", C.c_int32), ("batch", C.c_int64),
                ("frames", C.c_int64), ("bins", C.c_int64), ("rank", C.c_int32), ("iterations", C.c_int32),
                ("update_w", C.c_int32), ("update_h", C.c_int32), ("X", C.c_void_p), ("seeds", C.c_void_p),
                ("W0", C.c_void_p), ("H0", C.c_void_p), ("W1", C.c_void_p), ("H1", C.c_void_p), ("V1", C.c_void_p),
                ("progress", PROGRESS_FN), ("progress_user", C.c_void_p), ("progress_stride", C.c_int32),
                ("reserved", C.c_int32)]


class FramesArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("dtype", C.c_int32), ("mem", C.c_int32), ("frames", C.c_int64),
                ("bins", C.c_int64), ("rank", C.c_int32), ("iterations", C.c_int32), ("seed", C.c_int64),
                ("X", C.c_void_p), ("W0", C.c_void_p), ("W_norm", C.c_void_p), ("H", C.c_void_p), ("V", C.c_void_p)]


class BufNmfArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("mem", C.c_int32), ("batch", C.c_int64), ("n_samples", C.c_int64),
                ("rank", C.c_int32), ("iterations", C.c_int32), ("bases_mode", C.c_int32), ("acts_mode", C.c_int32),
                ("audio", C.c_void_p), ("seeds", C.c_void_p), ("bases_in", C.c_void_p), ("acts_in", C.c_void_p),
                ("bases_out", C.c_void_p), ("acts_out", C.c_void_p), ("resynth_out", C.c_void_p),
                ("progress", PROGRESS_FN), ("progress_user", C.c_void_p), ("progress_stride", C.c_int32),
                ("reserved", C.c_int32)]


class FilterArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("mem", C.c_int32), ("n_samples", C.c_int64), ("rank", C.c_int32),
                ("iterations", C.c_int32), ("seed", C.c_int64), ("audio", C.c_void_p), ("bases", C.c_void_p),
                ("out", C.c_void_p), ("acts_out", C.c_void_p)]


class FilterFramesArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("mem", C.c_int32), ("frames", C.c_int64), ("rank", C.c_int32),
                ("iterations", C.c_int32), ("seed", C.c_int64), ("in_", C.c_void_p), ("bases", C.c_void_p),
                ("out", C.c_void_p), ("acts_out", C.c_void_p)]


class NmfCrossArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("mem", C.c_int32), ("n_source", C.c_int64), ("n_target", C.c_int64),
                ("time_sparsity", C.c_int32), ("polyphony", C.c_int32), ("continuity", C.c_int32), ("iterations", C.c_int32),
                ("seed", C.c_int64), ("griffinlim_iterations", C.c_int32), ("reserved", C.c_int32), ("source", C.c_void_p),
                ("target", C.c_void_p), ("out", C.c_void_p), ("acts_out", C.c_void_p), ("progress", PROGRESS_FN),
                ("progress_user", C.c_void_p)]


class MelBandsArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("mem", C.c_int32), ("batch", C.c_int64), ("frames", C.c_int64),
                ("n_samples", C.c_int64), ("n_bands", C.c_int32), ("flags", C.c_int32), ("lo", C.c_double), ("hi", C.c_double),
                ("sample_rate", C.c_double), ("mags", C.c_void_p), ("audio", C.c_void_p), ("bands", C.c_void_p)]


class HpssArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("mem", C.c_int32), ("batch", C.c_int64), ("frames", C.c_int64),
                ("v_size", C.c_int32), ("h_size", C.c_int32), ("mode", C.c_int32), ("reserved", C.c_int32),
                ("thresholds", C.c_double * 8), ("spectrum", C.c_void_p), ("out", C.c_void_p)]


class NmfSeedArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("mem", C.c_int32), ("n_samples", C.c_int64), ("frames", C.c_int64),
                ("min_rank", C.c_int32), ("max_rank", C.c_int32), ("coverage", C.c_double), ("method", C.c_int32),
                ("scale_acts", C.c_int32), ("seed", C.c_int64), ("audio", C.c_void_p), ("mags", C.c_void_p),
                ("bases", C.c_void_p), ("acts", C.c_void_p), ("singular_values", C.c_void_p), ("rank_out", C.c_void_p)]


class ShardedArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_devices", C.c_int32), ("plans", C.POINTER(C.c_void_p)),
                ("job", C.POINTER(BufNmfArgs)), ("gathered_acts", C.POINTER(C.c_void_p))]


class BufStftArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("mem", C.c_int32), ("invert", C.c_int32), ("padding_mode", C.c_int32),
                ("batch", C.c_int64), ("n_samples", C.c_int64), ("frames", C.c_int64), ("audio", C.c_void_p),
                ("mag", C.c_void_p), ("phase", C.c_void_p), ("resynth", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("ms_h2d", "ms_stft", "ms_init", "ms_nmf", "ms_post", "ms_resynth", "ms_d2h",
                                         "ms_total")] + [("launches_total", C.c_int64), ("launches_nmf", C.c_int64),
                                                         ("backend_used", C.c_int32), ("update_kernel_launches", C.c_int32),
                                                         ("ms_update_kernel", C.c_float), ("reserved", C.c_float)]


# every symbol include/flucoma_b200.h declares (tests check the .so exports all of them)
SYMBOLS = ["fb200_abi_version", "fb200_device_count", "fb200_plan_create", "fb200_plan_destroy", "fb200_last_error",
           "fb200_num_frames", "fb200_resolve_fft", "fb200_shard_range", "fb200_stft", "fb200_istft",
           "fb200_nmf_process", "fb200_nmf_process_frames", "fb200_bufnmf", "fb200_nmf_filter", "fb200_get_stats",
           "fb200_get_api", "fb200_selftest_tcgen05", "fb200_bufstft_sizes", "fb200_bufstft", "fb200_nmf_filter_frames", "fb200_bufnmf_sharded", "fb200_bufnmfcross", "fb200_melbands", "fb200_hpss", "fb200_nmfseed"]

_lib = None


class FlucomaB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libflucoma_b200 error {code}: {msg}")
        self.code = code


def load(path: str | None = None):
    """dlopen the library.  Raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} not found: build it with `python flucoma-core_b200/build.py` "
                                "(there is no CPU fallback for the product path)")
    L = C.CDLL(p)
    L.fb200_abi_version.restype = C.c_uint32
    L.fb200_device_count.restype = C.c_int32
    L.fb200_plan_create.restype = C.c_int32
    L.fb200_plan_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.fb200_plan_destroy.argtypes = [C.c_void_p]
    L.fb200_plan_destroy.restype = None
    L.fb200_last_error.restype = C.c_char_p
    L.fb200_last_error.argtypes = [C.c_void_p]
    L.fb200_num_frames.restype = C.c_int64
    L.fb200_num_frames.argtypes = [C.c_int64, C.c_int32, C.c_int32]
    L.fb200_resolve_fft.restype = C.c_int32
    L.fb200_resolve_fft.argtypes = [C.c_int32, C.c_int32, C.c_int32] + [C.POINTER(C.c_int32)] * 3
    L.fb200_shard_range.restype = None
    L.fb200_shard_range.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.fb200_stft.restype = C.c_int32
    L.fb200_stft.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    L.fb200_istft.restype = C.c_int32
    L.fb200_istft.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32]
    for name, T in (("fb200_nmf_process", NmfArgs), ("fb200_nmf_process_frames", FramesArgs),
                    ("fb200_bufnmf", BufNmfArgs), ("fb200_nmf_filter", FilterArgs), ("fb200_bufstft", BufStftArgs),
                    ("fb200_nmf_filter_frames", FilterFramesArgs), ("fb200_bufnmfcross", NmfCrossArgs),
                    ("fb200_melbands", MelBandsArgs), ("fb200_hpss", HpssArgs), ("fb200_nmfseed", NmfSeedArgs)):
        fn = getattr(L, name)
        fn.restype = C.c_int32
        fn.argtypes = [C.c_void_p, C.POINTER(T)]
    L.fb200_bufnmf_sharded.restype = C.c_int32
    L.fb200_bufnmf_sharded.argtypes = [C.POINTER(ShardedArgs)]
    L.fb200_get_stats.restype = C.c_int32
    L.fb200_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.fb200_bufstft_sizes.restype = C.c_int32
    L.fb200_bufstft_sizes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_int64),
                                      C.POINTER(C.c_int64)]
    L.fb200_selftest_tcgen05.restype = C.c_int32
    L.fb200_selftest_tcgen05.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    L.fb200_get_api.restype = C.c_void_p
    L.fb200_get_api.argtypes = [C.c_uint32]
    if path is None:
        _lib = L
    return L


def num_frames(n_samples: int, win: int, hop: int) -> int:
    return int(load().fb200_num_frames(n_samples, win, hop))


def resolve_fft(win: int, hop: int = -1, fft: int = -1):
    o = [C.c_int32() for _ in range(3)]
    st = load().fb200_resolve_fft(win, hop, fft, *[C.byref(x) for x in o])
    if st != 0:
        raise FlucomaB200Error(st, "invalid FFT settings")
    return tuple(int(x.value) for x in o)  # hop, fft, bins


def shard_range(total: int, world: int, rank: int):
    b, c = C.c_int64(), C.c_int64()
    load().fb200_shard_range(total, world, rank, C.byref(b), C.byref(c))
    return int(b.value), int(c.value)


def bufstft_sizes(win: int, hop: int, padding_mode: int, invert: bool, count: int):
    """BufSTFT size rules (BufSTFTClient.hpp:121-131, 241-242): (padding, numHops) forward / (padding, samples) inverse."""
    pad, out = C.c_int64(), C.c_int64()
    st = load().fb200_bufstft_sizes(win, hop, padding_mode, int(bool(invert)), count, C.byref(pad), C.byref(out))
    if st != 0:
        raise FlucomaB200Error(st, "bufstft_sizes: input shorter than one window or bad arguments")
    return int(pad.value), int(out.value)


def device_count() -> int:
    return int(load().fb200_device_count())


def _is_torch(x):
    return x is not None and type(x).__module__.startswith("torch")


def _on_device(x):
    """True for a torch CUDA tensor (used in place, FB200_DEVICE); CPU torch tensors and numpy arrays are FB200_HOST."""
    return _is_torch(x) and bool(x.is_cuda)


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)


def _dtype_code(x):
    s = str(x.dtype)
    if s in ("float32", "torch.float32", "complex64", "torch.complex64"):
        return F32
    if s in ("float64", "torch.float64", "complex128", "torch.complex128"):
        return F64
    raise TypeError(f"unsupported dtype {s}")


class Plan:
    """Owns one fb200_plan (device, stream, cuFFT plans, workspaces)."""

    def __init__(self, win=1024, hop=-1, fft=-1, device=0, max_rank=64, max_batch=0, max_samples=0, backend=BACKEND_AUTO):
        self._L = load()
        backend = int(os.environ.get("FB200_BACKEND", backend))  # developer override (tools/, A/B runs)
        cfg = Config(C.sizeof(Config), device, win, hop, fft, max_rank, max_batch, max_samples, backend, 0)
        h = C.c_void_p()
        st = self._L.fb200_plan_create(C.byref(cfg), C.byref(h))
        if st != 0:
            raise FlucomaB200Error(st, self._L.fb200_last_error(None).decode())
        self._h = h
        self.device = device
        self.win = win
        self.hop, self.fft, self.bins = resolve_fft(win, hop, fft)

    def close(self):
        if getattr(self, "_h", None):
            self._L.fb200_plan_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _dev(self, x):
        """Memory space of an argument; a CUDA tensor must live on the plan's device (its pointer is used in place)."""
        if not _on_device(x):
            return False
        idx = x.device.index if x.device.index is not None else 0
        if idx != self.device:
            raise ValueError(f"tensor is on cuda:{idx} but the plan is bound to cuda:{self.device}")
        return True

    def _check(self, st):
        if st < 0:
            raise FlucomaB200Error(st, self._L.fb200_last_error(self._h).decode())
        return st

    def stats(self) -> dict:
        s = Stats()
        self._L.fb200_get_stats(self._h, C.byref(s))
        return {n: getattr(s, n) for n, _ in Stats._fields_ if n != "reserved"}

    # -- helpers ---------------------------------------------------------------------------------------------------
    def _empty_like_space(self, ref, shape, dtype_name):
        if _is_torch(ref):
            import torch
            return torch.empty(shape, dtype=getattr(torch, dtype_name), device=ref.device)
        return np.empty(shape, dtype=dtype_name)

    @staticmethod
    def _contig(x):
        if x is None:
            return None
        if _is_torch(x):
            return x.contiguous()
        return np.ascontiguousarray(x)

    def selftest_tcgen05(self, inputs: "np.ndarray") -> "np.ndarray":
        inp = np.ascontiguousarray(inputs, dtype=np.float32)
        out = np.zeros(128 * 64 + 128 * 16 + 128 * 64 + 128 * 16 + 128 * 32, np.float32)
        self._check(self._L.fb200_selftest_tcgen05(self._h, _ptr(inp), inp.size, _ptr(out), out.size))
        return out

    # -- STFT::process (+ magnitude) -------------------------------------------------------------------------------
    def stft(self, audio, want_spectrum=True, want_magnitude=False):
        a = self._contig(audio)
        if a.ndim == 1:
            a = a[None, :]
        batch, n = a.shape
        F = num_frames(n, self.win, self.hop)
        code = _dtype_code(a)
        real = "float32" if code == F32 else "float64"
        cplx = "complex64" if code == F32 else "complex128"
        spec = self._empty_like_space(a, (batch, F, self.bins), cplx) if want_spectrum else None
        mag = self._empty_like_space(a, (batch, F, self.bins), real) if want_magnitude else None
        self._check(self._L.fb200_stft(self._h, _ptr(a), batch, n, _ptr(spec), _ptr(mag), code,
                                       DEVICE if self._dev(a) else HOST))
        return spec, mag

    def istft(self, spectrum, n_samples):
        s = self._contig(spectrum)
        if s.ndim == 2:
            s = s[None]
        batch, F, B = s.shape
        assert B == self.bins
        code = _dtype_code(s)
        out = self._empty_like_space(s, (batch, n_samples), "float32" if code == F32 else "float64")
        self._check(self._L.fb200_istft(self._h, _ptr(s), batch, F, _ptr(out), n_samples, code,
                                        DEVICE if self._dev(s) else HOST))
        return out

    # -- BufSTFT (clients/nrt/BufSTFTClient.hpp:82-190, 192-279) -----------------------------------------------------
    def bufstft(self, audio, padding_mode=1, want_mag=True, want_phase=True):
        """audio [batch][n] (or [n]) float32 -> (mag, phase) [batch][numHops][bins] float32 (None where not wanted)."""
        a = self._contig(audio)
        squeeze = a.ndim == 1
        if squeeze:
            a = a[None, :]
        assert _dtype_code(a) == F32
        batch, n = a.shape
        _, hops = bufstft_sizes(self.win, self.hop, padding_mode, False, n)
        mag = self._empty_like_space(a, (batch, hops, self.bins), "float32") if want_mag else None
        ph = self._empty_like_space(a, (batch, hops, self.bins), "float32") if want_phase else None
        args = BufStftArgs(C.sizeof(BufStftArgs), DEVICE if self._dev(a) else HOST, 0, padding_mode, batch, n, hops, _ptr(a),
                           _ptr(mag), _ptr(ph), None)
        self._check(self._L.fb200_bufstft(self._h, C.byref(args)))
        if squeeze:
            mag = None if mag is None else mag[0]
            ph = None if ph is None else ph[0]
        return mag, ph

    def bufstft_inverse(self, mag, phase, padding_mode=1):
        """(mag, phase) [batch][frames][bins] (or [frames][bins]) float32 -> resynth [batch][(frames-1)*hop + win - padding]."""
        m = self._contig(mag); p = self._contig(phase)
        squeeze = m.ndim == 2
        if squeeze:
            m = m[None]; p = p[None]
        assert _dtype_code(m) == F32 and _dtype_code(p) == F32 and tuple(m.shape) == tuple(p.shape) and m.shape[2] == self.bins
        batch, frames, _ = m.shape
        _, n_out = bufstft_sizes(self.win, self.hop, padding_mode, True, frames)
        out = self._empty_like_space(m, (batch, n_out), "float32")
        args = BufStftArgs(C.sizeof(BufStftArgs), DEVICE if self._dev(m) else HOST, 1, padding_mode, batch, 0, frames, None,
                           _ptr(m), _ptr(p), _ptr(out))
        self._check(self._L.fb200_bufstft(self._h, C.byref(args)))
        return out[0] if squeeze else out

    # -- NMF::process ----------------------------------------------------------------------------------------------
    def nmf_process(self, X, rank, iterations, update_w=True, update_h=True, seeds=None, W0=None, H0=None,
                    want_v=True, progress=None, progress_stride=1):
        """X[batch][F][B] (or [F][B]) -> (W1[batch][K][B], H1[batch][F][K], V1[batch][F][B] | None, status)."""
        X = self._contig(X)
        squeeze = X.ndim == 2
        if squeeze:
            X = X[None]
            W0 = None if W0 is None else W0[None]
            H0 = None if H0 is None else H0[None]
        batch, F, B = X.shape
        code = _dtype_code(X)
        real = "float32" if code == F32 else "float64"
        W0 = self._contig(W0); H0 = self._contig(H0)
        for z in (W0, H0):
            if z is not None:
                assert _dtype_code(z) == code and _is_torch(z) == _is_torch(X)
        if W0 is not None:
            assert tuple(W0.shape) == (batch, rank, B)  # NMF.hpp:109-110
        if H0 is not None:
            assert tuple(H0.shape) == (batch, F, rank)  # NMF.hpp:121-122
        if seeds is None:
            seeds = [-1] * batch
        if np.isscalar(seeds):
            seeds = [int(seeds)] * batch
        seeds = np.ascontiguousarray(seeds, dtype=np.int64)
        assert seeds.shape == (batch,)
        W1 = self._empty_like_space(X, (batch, rank, B), real)
        H1 = self._empty_like_space(X, (batch, F, rank), real)
        V1 = self._empty_like_space(X, (batch, F, B), real) if want_v else None
        cb = PROGRESS_FN(lambda user, it: int(bool(progress(it)))) if progress else PROGRESS_FN()
        a = NmfArgs(C.sizeof(NmfArgs), code, DEVICE if self._dev(X) else HOST, batch, F, B, rank, iterations,
                    int(update_w), int(update_h), _ptr(X), _ptr(seeds), _ptr(W0), _ptr(H0), _ptr(W1), _ptr(H1),
                    _ptr(V1), cb, None, progress_stride, 0)
        st = self._check(self._L.fb200_nmf_process(self._h, C.byref(a)))
        if squeeze:
            W1, H1 = W1[0], H1[0]
            V1 = None if V1 is None else V1[0]
        return W1, H1, V1, st

    # -- NMF::processFrame over many frames --------------------------------------------------------------------------
    def nmf_process_frames(self, X, W0, iterations=10, seed=-1, want_v=False, want_w=False):
        X = self._contig(X); W0 = self._contig(W0)
        F, B = X.shape
        K = W0.shape[0]
        assert W0.shape[1] == B
        code = _dtype_code(X)
        assert _dtype_code(W0) == code
        real = "float32" if code == F32 else "float64"
        H = self._empty_like_space(X, (F, K), real)
        V = self._empty_like_space(X, (F, B), real) if want_v else None
        Wn = self._empty_like_space(X, (K, B), real) if want_w else None
        a = FramesArgs(C.sizeof(FramesArgs), code, DEVICE if self._dev(X) else HOST, F, B, K, iterations, seed,
                       _ptr(X), _ptr(W0), _ptr(Wn), _ptr(H), _ptr(V))
        self._check(self._L.fb200_nmf_process_frames(self._h, C.byref(a)))
        return H, V, Wn

    # -- BufNMF ------------------------------------------------------------------------------------------------------
    def bufnmf(self, audio, rank, iterations, seeds=None, bases_mode=0, bases_in=None, acts_mode=0, acts_in=None,
               resynth=False, out=None, progress=None, progress_stride=1):
        """audio float32 [batch][n] -> dict(bases[batch][K][B], acts[batch][F][K], resynth[batch][K][n] | None, status).

        `out` may hold preallocated 'bases'/'acts'/'resynth' arrays in the same memory space (bench reuses them)."""
        a_in = self._contig(audio)
        if a_in.ndim == 1:
            a_in = a_in[None, :]
        assert _dtype_code(a_in) == F32, "BufNMF host buffers are float32 (BufferAdaptor.hpp:49-66)"
        batch, n = a_in.shape
        F = num_frames(n, self.win, self.hop)
        B = self.bins
        out = out or {}
        fix_w, fix_h = bases_mode == 2, acts_mode == 2
        bases = out.get("bases") if out.get("bases") is not None else self._empty_like_space(a_in, (batch, rank, B), "float32")
        acts = out.get("acts") if out.get("acts") is not None else self._empty_like_space(a_in, (batch, F, rank), "float32")
        rs = None
        if resynth:
            rs = out.get("resynth") if out.get("resynth") is not None else self._empty_like_space(a_in, (batch, rank, n), "float32")
        bases_in = self._contig(bases_in); acts_in = self._contig(acts_in)
        if bases_in is not None:
            assert tuple(bases_in.shape) == (batch, rank, B) and _dtype_code(bases_in) == F32
        if acts_in is not None:
            assert tuple(acts_in.shape) == (batch, F, rank) and _dtype_code(acts_in) == F32
        if seeds is None:
            seeds = [-1] * batch
        if np.isscalar(seeds):
            seeds = [int(seeds)] * batch
        seeds = np.ascontiguousarray(seeds, dtype=np.int64)
        assert seeds.shape == (batch,)
        cb = PROGRESS_FN(lambda user, it: int(bool(progress(it)))) if progress else PROGRESS_FN()
        a = BufNmfArgs(C.sizeof(BufNmfArgs), DEVICE if self._dev(a_in) else HOST, batch, n, rank, iterations,
                       bases_mode, acts_mode, _ptr(a_in), _ptr(seeds), _ptr(bases_in), _ptr(acts_in),
                       None if fix_w else _ptr(bases), None if fix_h else _ptr(acts), _ptr(rs), cb, None,
                       progress_stride, 0)
        st = self._L.fb200_bufnmf(self._h, C.byref(a))
        self._check(st)
        return dict(bases=None if fix_w else bases, acts=None if fix_h else acts, resynth=rs, status=st)

    # -- streaming NMFFilter / NMFMatch ------------------------------------------------------------------------------
    def nmf_filter(self, audio, bases, iterations=10, seed=-1, want_out=True, want_acts=True):
        """Mono stream float32 [n] + bases float32 [K][bins] -> (out [K][n] | None, acts [ceil(n/hop)][K] | None).

        NMFFilterClient.hpp:98-117 (and NMFMatchClient.hpp:106-118 when want_out is False) from reset state; `out`
        carries the client's latency of `win` samples."""
        a_in = self._contig(audio); W = self._contig(bases)
        assert a_in.ndim == 1 and _dtype_code(a_in) == F32 and _dtype_code(W) == F32
        n = a_in.shape[0]
        K = W.shape[0]
        assert W.shape[1] == self.bins
        frames = (n + self.hop - 1) // self.hop
        out = self._empty_like_space(a_in, (K, n), "float32") if want_out else None
        acts = self._empty_like_space(a_in, (frames, K), "float32") if want_acts else None
        a = FilterArgs(C.sizeof(FilterArgs), DEVICE if self._dev(a_in) else HOST, n, K, iterations, seed, _ptr(a_in),
                       _ptr(W), _ptr(out), _ptr(acts))
        self._check(self._L.fb200_nmf_filter(self._h, C.byref(a)))
        return out, acts

    def bufnmfcross(self, source, target, time_sparsity=7, polyphony=11, continuity=7, iterations=50, seed=-1,
                    griffinlim_iterations=50, want_out=True, want_acts=True, progress=None):
        """BufNMFCross (NMFCrossClient.hpp:85-185): mono float32 source [ns], target [nt] -> (out [nt] | None,
        acts [target frames][source frames] | None, status)."""
        s_ = self._contig(source); t_ = self._contig(target)
        assert s_.ndim == 1 and t_.ndim == 1 and _dtype_code(s_) == F32 and _dtype_code(t_) == F32
        ns, nt = s_.shape[0], t_.shape[0]
        Fs, Ft = num_frames(ns, self.win, self.hop), num_frames(nt, self.win, self.hop)
        out = self._empty_like_space(t_, (nt,), "float32") if want_out else None
        acts = self._empty_like_space(t_, (Ft, Fs), "float32") if want_acts else None
        cb = PROGRESS_FN(lambda user, it: int(bool(progress(it)))) if progress else PROGRESS_FN()
        a = NmfCrossArgs(C.sizeof(NmfCrossArgs), DEVICE if self._dev(t_) else HOST, ns, nt, time_sparsity, polyphony, continuity,
                         iterations, seed, griffinlim_iterations, 0, _ptr(s_), _ptr(t_), _ptr(out), _ptr(acts), cb, None)
        st = self._check(self._L.fb200_bufnmfcross(self._h, C.byref(a)))
        return out, acts, st

    def nmfseed(self, mags=None, audio=None, min_rank=1, max_rank=200, coverage=0.5, method=0, seed=-1, scale_acts=False):
        """NMFSeed / NNDSVD (NNDSVD.hpp:30-131, NMFSeedClient.hpp:74-133): magnitudes float32 [F][bins] or mono audio [n] ->
        (rank, bases [max_rank][bins], acts [F][max_rank], singular values)."""
        x = self._contig(mags if mags is not None else audio)
        assert _dtype_code(x) == F32
        if mags is not None:
            F, n = x.shape[0], 0
            assert x.shape[1] == self.bins
        else:
            n = x.shape[0]
            F = num_frames(n, self.win, self.hop)
        bases = self._empty_like_space(x, (max_rank, self.bins), "float32")
        acts = self._empty_like_space(x, (F, max_rank), "float32")
        sv = np.zeros(min(F, self.bins))
        rank = C.c_int32(0)
        a = NmfSeedArgs(C.sizeof(NmfSeedArgs), DEVICE if self._dev(x) else HOST, n, F, min_rank, max_rank, float(coverage), method,
                        int(scale_acts), seed, _ptr(x) if mags is None else None, _ptr(x) if mags is not None else None, _ptr(bases),
                        _ptr(acts), _ptr(sv), C.cast(C.pointer(rank), C.c_void_p))
        self._check(self._L.fb200_nmfseed(self._h, C.byref(a)))
        return int(rank.value), bases, acts, sv

    def melbands(self, mags=None, audio=None, n_bands=40, lo=20.0, hi=20000.0, sample_rate=44100.0, mag_norm=True, use_power=False,
                 log_output=False):
        """MelBands (MelBands.hpp:43-101) of magnitudes float32 [batch][F][bins] (or [F][bins]) or of audio [batch][n]."""
        x = self._contig(mags if mags is not None else audio)
        assert _dtype_code(x) == F32
        squeeze = x.ndim == (2 if mags is not None else 1)
        if squeeze:
            x = x[None]
        batch = x.shape[0]
        if mags is not None:
            F, n = x.shape[1], 0
            assert x.shape[2] == self.bins
        else:
            n = x.shape[1]
            F = num_frames(n, self.win, self.hop)
        out = self._empty_like_space(x, (batch, F, n_bands), "float32")
        flags = (1 if mag_norm else 0) | (2 if use_power else 0) | (4 if log_output else 0)
        a = MelBandsArgs(C.sizeof(MelBandsArgs), DEVICE if self._dev(x) else HOST, batch, F, n, n_bands, flags, lo, hi, sample_rate,
                         _ptr(x) if mags is not None else None, _ptr(x) if mags is None else None, _ptr(out))
        self._check(self._L.fb200_melbands(self._h, C.byref(a)))
        return out[0] if squeeze else out

    def hpss(self, spectrum, v_size=31, h_size=17, mode=0, h_thresh=(0.0, 1.0, 1.0, 1.0), p_thresh=(0.0, 1.0, 1.0, 1.0)):
        """HPSS::processFrame over spectrum complex64 [batch][F][bins] (or [F][bins]) -> [batch][3][F][bins]."""
        s_ = self._contig(spectrum)
        assert _dtype_code(s_) == F32
        squeeze = s_.ndim == 2
        if squeeze:
            s_ = s_[None]
        batch, F, B = s_.shape
        assert B == self.bins
        out = self._empty_like_space(s_, (batch, 3, F, B), "complex64")
        a = HpssArgs(C.sizeof(HpssArgs), DEVICE if self._dev(s_) else HOST, batch, F, v_size, h_size, mode, 0,
                     (C.c_double * 8)(*[float(v) for v in (*h_thresh, *p_thresh)]), _ptr(s_), _ptr(out))
        self._check(self._L.fb200_hpss(self._h, C.byref(a)))
        return out[0] if squeeze else out

    def nmf_filter_frames(self, frames, bases, iterations=10, seed=-1, want_out=True, want_acts=True):
        """frames float32 [nf][win] (raw, as FluidSource::pull cuts them) + bases [K][bins] -> (out [nf][K][win] | None,
        acts [nf][K] | None): the per-frame body of NMFFilterClient / NMFMatchClient (fb200_nmf_filter_frames)."""
        x = self._contig(frames); W = self._contig(bases)
        assert x.ndim == 2 and x.shape[1] == self.win and _dtype_code(x) == F32 and _dtype_code(W) == F32
        nf, K = x.shape[0], W.shape[0]
        assert W.shape[1] == self.bins
        out = self._empty_like_space(x, (nf, K, self.win), "float32") if want_out else None
        acts = self._empty_like_space(x, (nf, K), "float32") if want_acts else None
        a = FilterFramesArgs(C.sizeof(FilterFramesArgs), DEVICE if self._dev(x) else HOST, nf, K, iterations, seed, _ptr(x),
                             _ptr(W), _ptr(out), _ptr(acts))
        self._check(self._L.fb200_nmf_filter_frames(self._h, C.byref(a)))
        return out, acts


def bufnmf_sharded(plans, audio, rank, iterations, seeds=None, resynth=False, gather=False, progress=None,
                   progress_stride=PROGRESS_ASYNC):
    """BufNMF of host audio float32 [batch][n] over several plans (one per device) in ONE call (fb200_bufnmf_sharded).
    Returns dict(bases, acts, resynth | None, status, gathered | None); `gathered` is a list of torch tensors, one per
    device, each [n_dev * ceil(batch / n_dev)][F][K]."""
    L = load()
    a_in = np.ascontiguousarray(audio, dtype=np.float32)
    batch, n = a_in.shape
    p0 = plans[0]
    F, B = num_frames(n, p0.win, p0.hop), p0.bins
    bases = np.empty((batch, rank, B), np.float32)
    acts = np.empty((batch, F, rank), np.float32)
    rs = np.empty((batch, rank, n), np.float32) if resynth else None
    if seeds is None:
        seeds = [-1] * batch
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    cb = PROGRESS_FN(lambda user, it: int(bool(progress(it)))) if progress else PROGRESS_FN()
    job = BufNmfArgs(C.sizeof(BufNmfArgs), HOST, batch, n, rank, iterations, 0, 0, _ptr(a_in), _ptr(seeds), None, None,
                     _ptr(bases), _ptr(acts), _ptr(rs), cb, None, progress_stride, 0)
    handles = (C.c_void_p * len(plans))(*[p._h for p in plans])
    gathered, gptr = None, None
    if gather:
        import torch
        per = (batch + len(plans) - 1) // len(plans)
        gathered = [torch.empty((len(plans) * per, F, rank), dtype=torch.float32, device=f"cuda:{p.device}") for p in plans]
        gptr = (C.c_void_p * len(plans))(*[g.data_ptr() for g in gathered])
    args = ShardedArgs(C.sizeof(ShardedArgs), len(plans), handles, C.pointer(job), gptr)
    st = L.fb200_bufnmf_sharded(C.byref(args))
    if st < 0:
        raise FlucomaB200Error(st, L.fb200_last_error(plans[0]._h).decode())
    return dict(bases=bases, acts=acts, resynth=rs, status=st, gathered=gathered)
