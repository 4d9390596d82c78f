#!/usr/bin/env python
"""bench.py -- benchmark of the STFT -> |X| -> NMF hot path (BASELINE.json metric: NMF frames/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5] [--scaling weak|strong]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload = BASELINE.json configs[1] ("config 2"): per GPU, batch=1024 distinct synthetic float32 buffers x 130816
samples (F=512 frames each), fft=win=1024, hop=256, rank K=16, 200 multiplicative-update iterations.  One "step" = one
BufNMF pass over that batch: STFT -> magnitude -> W/H init from per-buffer seeds -> 200 iterations -> bases + activations.
The other BASELINE configs are selectable with --config (builder-kept lines under profiles/):
  1  one 515088-sample buffer (the length of Tremblay-AaS-SynthTwoVoices-M.wav), fft 1024 hop 256 rank 4, 100 iterations
  3  batch 1024 per GPU (8192 over 8), rank 32, 200 iterations;  --scaling strong: 8192 buffers in total, split over N
  4  batch 256, fft 4096 hop 1024 (523264 samples), rank 64, 500 iterations
  5  NMFMatch: fixed W (16 x 513), 10^6 frames of |X| resident in HBM, 10 activation-only iterations

  value  : frames/s with the inputs already resident in HBM and outputs left in HBM (device pointers through the C ABI)
  e2e    : the same call with HOST (pinned) buffers: H2D of the inputs and D2H of the results inside the timed region
  e2e_callback : e2e again with a progress callback installed, as every real BufNMF job has (NMFClient.hpp:261-267)
  roofline / cpu_baseline : see DESIGN.md "Measurement"
  --impl reference : the CPU restatement of the reference (oracle/, fp64, faithful mode, all host threads) on a
                     bounded sample of the same workload (the reference itself needs Eigen/HISSTools: not buildable here)

Multi-GPU: buffers are independent, so each rank processes its own shard with no data-path collective; the final
activations are all-gathered with NCCL inside the timed region (north_star).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "flucoma-core_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

CONFIGS = {
    1: dict(name="config1: one buffer of 515088 samples (F=2013), fft=1024 hop=256 rank=4 iters=100",
            batch=1, n=515088, win=1024, fft=1024, hop=256, rank=4, iters=100, kind="bufnmf"),
    2: dict(name="config2: batch=1024 x 130816 samples (512 frames), fft=1024 hop=256 rank=16 iters=200",
            batch=1024, n=130816, win=1024, fft=1024, hop=256, rank=16, iters=200, kind="bufnmf"),
    3: dict(name="config3: batch=1024 per GPU (8192 over 8) x 130816 samples, fft=1024 hop=256 rank=32 iters=200",
            batch=1024, n=130816, win=1024, fft=1024, hop=256, rank=32, iters=200, kind="bufnmf", strong_total=8192),
    4: dict(name="config4: batch=256 x 523264 samples (512 frames), fft=4096 hop=1024 rank=64 iters=500",
            batch=256, n=523264, win=4096, fft=4096, hop=1024, rank=64, iters=500, kind="bufnmf"),
    5: dict(name="config5: NMFMatch, fixed W 16x513, 10^6 frames of |X| (fft=1024), 10 activation-only iterations",
            frames=1_000_000, win=1024, fft=1024, hop=512, rank=16, iters=10, kind="frames"),
}
WORKLOAD = CONFIGS[2]
METRIC = "NMF frames/sec (batch x frames) at rank=16, fft=1024"
UNIT = "frames/s"


def _synth_one(args):
    from tests.golden.make_golden import synth_audio
    seed, n = args
    return synth_audio(seed, n)


def make_audio(batch, n, base_seed=1000, distinct=None):
    """SURVEY 8d synthetic buffers: buffer b = synth_audio(base_seed + b) (6 gated partials + noise), all distinct.
    Generated once with a process pool and cached under $TMPDIR (1024 buffers take ~35 s of single-thread numpy)."""
    import tempfile
    distinct = batch if distinct is None else min(distinct, batch)
    path = os.path.join(tempfile.gettempdir(), f"fb200_audio_{base_seed}_{distinct}_{n}.npy")
    uniq = None
    if os.path.exists(path):
        try:
            uniq = np.load(path)
        except Exception:
            uniq = None
    if uniq is None or uniq.shape != (distinct, n):
        jobs = [(base_seed + i, n) for i in range(distinct)]
        if distinct >= 32:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(min(os.cpu_count() or 1, 32)) as pool:
                uniq = np.stack(pool.map(_synth_one, jobs, chunksize=8))
        else:
            uniq = np.stack([_synth_one(j) for j in jobs])
        try:
            tmp = path + f".{os.getpid()}.tmp.npy"
            np.save(tmp, uniq)
            os.replace(tmp, path)
        except Exception:
            pass
    if distinct == batch:
        return uniq
    reps = (batch + distinct - 1) // distinct
    return np.concatenate([uniq] * reps)[:batch].copy()


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop_evt.is_set():
            try:
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, util))
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        loaded = [m for m, u in self.samples if u > 0] or [m for m, _ in self.samples]
        return {"sm_mhz": float(np.median(loaded)) if loaded else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1344.4), d.get("hbm_gbs", 6539.9), "measured (MEASURED_PEAKS.json: sustained bf16, copy GB/s)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------- CPU baselines
def native_oracle():
    """CPU baseline library: the oracle rebuilt with -march=native for this box's cores (falls back to the shipped build)."""
    from oracle import c_oracle as co
    import tempfile
    try:
        out = os.path.join(tempfile.gettempdir(), "libflucoma_oracle_native.so")
        co.build(force=True, march="native", out=out)
        co.lib(out)
        return co, out
    except Exception:
        co.build()
        return co, None


def frames_of(w):
    from oracle import c_oracle as co
    return w["frames"] if w["kind"] == "frames" else co.num_frames(w["n"], w["win"], w["hop"])


def cpu_sample(co, lib_path, threads, w, nbuf, iters=None):
    """`nbuf` buffers of workload `w` (one per thread), fp64, reference-faithful (7 GEMMs per iteration); iterations may be
    cut for the large configs -- the per-iteration cost is constant, the caller scales.  Returns (frames/s at the full
    iteration count, wall seconds, iterations run)."""
    iters = iters or w["iters"]
    if w["kind"] == "frames":
        # NMFMatch: processFrame per frame (NMF.hpp:45-89); sample = nbuf * 4096 frames of |X|
        rng = np.random.default_rng(5)
        nfr = 4096 * nbuf
        X = np.abs(rng.standard_normal((nfr, w["fft"] // 2 + 1)))
        W = rng.random((w["rank"], w["fft"] // 2 + 1))
        t0 = time.perf_counter()
        co.nmfmatch_frames(X, W, w["iters"], 42, threads=threads, lib_path=lib_path)
        dt = time.perf_counter() - t0
        return nfr / dt, dt, w["iters"]
    audio = make_audio(nbuf, w["n"], distinct=min(nbuf, 64))
    seeds = np.arange(nbuf, dtype=np.int64)
    t0 = time.perf_counter()
    co.bufnmf_batch(audio, w["win"], w["fft"], w["hop"], w["rank"], iters, seeds, resynth=False, faithful=True,
                    threads=threads, lib_path=lib_path)
    dt = time.perf_counter() - t0
    F = co.num_frames(w["n"], w["win"], w["hop"])
    return nbuf * F / (dt * w["iters"] / iters), dt, iters


def blas_sample(w, threads, nbuf, iters):
    """B3, the second opinion of BASELINE.md 3: the same seven GEMMs per iteration through NumPy/OpenBLAS (fp64), `nbuf`
    buffers one after the other with `threads` BLAS threads.  Returns frames/s at the full iteration count."""
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        threadpool_limits = None
    B, K = w["fft"] // 2 + 1, w["rank"]
    F = frames_of(w)
    rng = np.random.default_rng(0)
    eps = np.finfo(np.float64).eps

    def run():
        for _ in range(nbuf):
            V = np.asfortranarray(rng.random((B, F))); W = np.asfortranarray(rng.random((B, K))); H = np.asfortranarray(rng.random((K, F)))
            ones = np.ones_like(V)
            for _ in range(iters):
                V1 = np.maximum(W @ H, eps)                      # NMF.hpp:158
                W = W * ((V / V1) @ H.T) / np.maximum(ones @ H.T, eps)   # :159-161
                W = W / np.linalg.norm(W, axis=0)                # :162
                V2 = np.maximum(W @ H, eps)                      # :165
                H = H * (W.T @ (V / V2)) / np.maximum(W.T @ ones, eps)   # :168-170
                np.maximum(W @ H, eps)                           # :173-174 (dead R)
    t0 = time.perf_counter()
    if threadpool_limits:
        with threadpool_limits(limits=threads):
            run()
    else:
        run()
    dt = time.perf_counter() - t0
    return nbuf * F / (dt * w["iters"] / iters), dt


def cpu_baselines(w, budget_iters=None):
    """B1 (1 thread, what one reference BufNMF job costs), B2 (all host threads, one buffer per thread), B3 (OpenBLAS)."""
    co, lib_path = native_oracle()
    threads = min(os.cpu_count() or 1, 64)
    it = budget_iters or w["iters"]
    b1, dt1, _ = cpu_sample(co, lib_path, 1, w, 1, it)
    b2, dt2, _ = cpu_sample(co, lib_path, threads, w, threads, it)
    out = {"value": b2, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{threads} buffers of the workload, one per thread, fp64 faithful mode (7 GEMMs/iter), {it} of {w['iters']} "
                     f"iterations timed and scaled, {dt2:.1f} s wall",
           "one_thread": {"value": b1, "cores": 1, "sample": f"1 buffer, {it} iterations, {dt1:.1f} s wall"},
           "all_threads": {"value": b2, "cores": threads}}
    if w["kind"] == "bufnmf":
        it3 = max(1, min(it, 50))
        v1, d1 = blas_sample(w, 1, 1, it3)
        vn, dn = blas_sample(w, threads, 2, it3)
        out["openblas"] = {"one_thread": v1, "all_threads": vn, "cores": threads,
                           "sample": f"NumPy/OpenBLAS fp64 restatement of NMF.hpp:144-183, update loop only, {it3} iterations scaled "
                                     f"({d1:.1f} s / {dn:.1f} s)"}
        out["port_vs_openblas_one_thread"] = b1 / v1 if v1 else None
    return out


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (restated in oracle/, fp64, 7 GEMMs/iter as NMF.hpp:144-183 runs
    them) with one buffer per host thread; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = CONFIGS[args.config]
    co, lib_path = native_oracle()
    threads = min(os.cpu_count() or 1, 64)
    nbuf = threads
    it = w["iters"] if args.config in (1, 2, 5) else max(1, w["iters"] // 10)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_sample(co, lib_path, threads, w, nbuf, it)
    vals, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_sample(co, lib_path, threads, w, nbuf, it)[0])
    dt = time.perf_counter() - t0
    value = float(np.mean(vals))
    sample = (f"{nbuf} buffers of the workload per step, one per thread ({threads} threads), fp64 faithful mode, "
              f"{it} of {w['iters']} iterations timed and scaled")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import flucoma_b200 as fb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner (and, with NCCL_DEBUG set, its log) on stdout when the communicator is created;
        # stdout must carry exactly one JSON line, so communicator creation runs with fd 1 pointed at stderr.
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    w = dict(CONFIGS[args.config])
    K, iters = w["rank"], w["iters"]
    strong = args.scaling == "strong"
    plan = fb.Plan(win=w["win"], hop=w["hop"], fft=w["fft"], device=local, max_rank=K)
    B = plan.bins

    if w["kind"] == "bufnmf":
        batch, n = w["batch"], w["n"]
        if strong:
            first, batch = fb.shard_range(w.get("strong_total", batch), world, rank)
        else:
            first = rank * batch
        F = fb.num_frames(n, w["win"], w["hop"])
        seeds = np.arange(first, first + batch, dtype=np.int64)
        # distinct buffers (default_rng(1000 + b)); beyond 1024 per rank the signals repeat, the NMF seeds do not
        distinct = min(batch, 1024)
        audio_h = torch.from_numpy(make_audio(batch, n, base_seed=1000 + (first % 8192), distinct=distinct)).pin_memory()
        audio_d = audio_h.cuda()
        out_d = {"bases": torch.empty((batch, K, B), dtype=torch.float32, device="cuda"),
                 "acts": torch.empty((batch, F, K), dtype=torch.float32, device="cuda")}
        out_h = {"bases": torch.empty((batch, K, B), dtype=torch.float32).pin_memory().numpy(),
                 "acts": torch.empty((batch, F, K), dtype=torch.float32).pin_memory().numpy()}
        frames_rank = batch * F
        if world > 1:
            counts = [fb.shard_range(w.get("strong_total", w["batch"]), world, r)[1] if strong else batch for r in range(world)]
            pad = max(counts)
            gathered = torch.empty((world * pad, F, K), dtype=torch.float32, device="cuda")
            send = out_d["acts"] if pad == batch else torch.zeros((pad, F, K), dtype=torch.float32, device="cuda")

        def step_device():
            plan.bufnmf(audio_d, K, iters, seeds=seeds, out=out_d)
            st = plan.stats()
            if world > 1:
                if send is not out_d["acts"]:
                    send[:batch].copy_(out_d["acts"])
                dist.all_gather_into_tensor(gathered, send)  # north_star: one allgather of the final activations
            return st

        def step_host():
            plan.bufnmf(audio_h.numpy(), K, iters, seeds=seeds, out=out_h)
            return plan.stats()

        ticks = [0]

        def step_host_cb():
            plan.bufnmf(audio_h.numpy(), K, iters, seeds=seeds, out=out_h,
                        progress=lambda it: ticks.__setitem__(0, ticks[0] + 1) or True, progress_stride=fb.PROGRESS_ASYNC)
            return plan.stats()

        h2d = int(batch * n * 4 + batch * 8)
        d2h = int(batch * K * B * 4 + batch * F * K * 4)
        flops_step = 8.0 * B * K * iters * batch * F
        l2_note = f"inputs ({batch * n * 4 / 1e6:.0f} MB audio, {batch * F * B * 4 / 1e6:.0f} MB |X|) vs the 126 MB L2"
    else:
        # config 5: |X| of a long stream already in HBM; shard the frames over the ranks
        total = w["frames"]
        first, F = fb.shard_range(total, world, rank) if (strong or world == 1) else (rank * total, total)
        batch = 1
        nb = 64
        a = torch.from_numpy(make_audio(nb, 130816, base_seed=5000 + rank, distinct=nb)).cuda()
        with fb.Plan(win=w["win"], hop=w["hop"], fft=w["fft"], device=local) as sp:
            _, m = sp.stft(a, want_spectrum=False, want_magnitude=True)
        m = m.reshape(-1, B)
        reps = (F + m.shape[0] - 1) // m.shape[0]
        X_d = m.repeat(reps, 1)[:F].contiguous()
        del m, a
        rng = np.random.default_rng(7)
        W_h = rng.random((K, B)).astype(np.float32)
        W_d = torch.from_numpy(W_h).cuda()
        X_h = torch.empty((F, B), dtype=torch.float32).pin_memory()
        X_h.copy_(X_d)
        frames_rank = F

        def step_device():
            plan.nmf_process_frames(X_d, W_d, iters, seed=42)
            return plan.stats()

        def step_host():
            plan.nmf_process_frames(X_h.numpy(), W_h, iters, seed=42)
            return plan.stats()

        step_host_cb = None
        h2d = int(F * B * 4 + K * B * 4)
        d2h = int(F * K * 4)
        flops_step = 4.0 * B * K * iters * F
        l2_note = f"|X| stream of {F * B * 4 / 1e9:.2f} GB vs the 126 MB L2"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, steps):
        agg = {"launches": 0, "ms_update_kernel": 0.0, "update_kernel_launches": 0, "ms_nmf": 0.0, "ms_stft": 0.0,
               "ms_total": 0.0, "backend": 0}
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            st = step()
            agg["launches"] += st["launches_total"]
            agg["backend"] = st["backend_used"]
            for k in ("ms_update_kernel", "update_kernel_launches", "ms_nmf", "ms_stft", "ms_total"):
                agg[k] += st[k]
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        # the C-ABI calls are synchronous, so both the torch events and the host clock bracket all device work;
        # take the larger (the events sit on torch's stream, the library launches on its own)
        ms = max(e0.elapsed_time(e1), 1e3 * wall)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, agg

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    ms, agg = timed(step_device, args.steps)
    clocks = sampler.stop()
    for _ in range(min(args.warmup, 2)):
        step_host()
    ms_e2e, _ = timed(step_host, args.steps)
    ms_cb = None
    if step_host_cb is not None:
        step_host_cb()
        ms_cb, _ = timed(step_host_cb, args.steps)

    frames_step = frames_rank
    if world > 1:
        t = torch.tensor([float(frames_rank)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        frames_step = float(t.item())
    value = frames_step * args.steps / (ms * 1e-3)
    e2e = frames_step * args.steps / (ms_e2e * 1e-3)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "per_gpu_batch": batch, "l2": l2_note + " (inputs larger than L2, no flush needed)",
                       "parallelism": (f"batch-sharded x{world}, allgather(H) at the end" if w["kind"] == "bufnmf" else f"frames sharded x{world}")
                       if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(agg["launches"]),
            "stages_ms_per_step": {k: agg[k] / args.steps for k in ("ms_stft", "ms_nmf", "ms_total")}}
    if ms_cb is not None:
        line["e2e_callback"] = {"value": frames_step * args.steps / (ms_cb * 1e-3), "unit": UNIT, "ms_per_step": ms_cb / args.steps,
                                "vs_e2e": ms_e2e / ms_cb, "callbacks_per_step": ticks[0] / (args.steps + 1),
                                "mode": "FB200_PROGRESS_ASYNC (what the host NMFClient mirror installs)"}

    if rank == 0:
        # roofline of the dominant kernel: the NMF update kernel.  Algorithmic flops per frame per iteration = 8*B*K
        # (4*B*K for the activation-only update), SURVEY 8d.
        peak_tf, peak_gbs, which = measured_peaks()
        ms_kernel = agg["ms_update_kernel"] / args.steps
        nl = max(1.0, agg["update_kernel_launches"] / args.steps)
        kname = {1: "k_nmf_tile (fp32 SIMT)", 2: "k_nmf_tc (tcgen05, split-bf16 operands, fp32 accumulate)",
                 3: "k_nmf_tcs (tcgen05, streamed split-bf16 operands, fp32 accumulate)"}.get(agg["backend"], "?")
        if w["kind"] == "frames":
            bytes_step = float(frames_rank) * (4 * B + 4 * K)
            achieved = bytes_step / (ms_kernel * 1e-3) / 1e9 if ms_kernel > 0 else None
            line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                                "frac": (achieved / peak_gbs) if achieved else None, "traffic": None, "kernel": kname,
                                "peak_source": which,
                                "note": "algorithmic bytes = 4*B (|X| in) + 4*K (H out) per frame (SURVEY 8d); also "
                                        f"{flops_step / (ms_kernel * 1e-3) / 1e12 if ms_kernel > 0 else 0:.1f} TFLOP/s of 4*B*K*iters flops per frame",
                                "launches_per_step": nl, "avg_launch_ms": ms_kernel / nl,
                                "kernel_share_of_step": ms_kernel / (agg["ms_total"] / args.steps)}
        else:
            achieved = flops_step / (ms_kernel * 1e-3) / 1e12 if ms_kernel > 0 else None
            traffic = None
            tp = os.path.join(ROOT, "profiles", "update_kernel_traffic.json")
            if os.path.exists(tp) and args.config == 2:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            line["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                                "frac": (achieved / peak_tf) if achieved else None, "traffic": traffic,
                                "kernel": kname, "peak_source": which,
                                "note": "achieved counts the algorithmic 8*B*K flops per frame per iteration; the tensor pipe "
                                        "executes 6x that (3 split terms in each of the two MMAs), see DESIGN.md 4.2",
                                "launches_per_step": nl, "avg_launch_ms": ms_kernel / nl,
                                "kernel_share_of_step": ms_kernel / (agg["ms_total"] / args.steps)}
            if agg["backend"] == 3 and ms_kernel > 0:
                # the streamed engine reads |X| twice per iteration (H and W half-iterations are not fused) and the batch
                # does not fit L2: for configs 3 / 4 HBM is the nearer roof
                xb = 2.0 * 4.0 * (B - 1) * frames_rank * w["iters"]
                line["roofline"]["hbm_view"] = {"algorithmic_x_bytes": xb, "achieved": xb / (ms_kernel * 1e-3) / 1e9, "peak": peak_gbs,
                                                "unit": "GB/s", "frac": xb / (ms_kernel * 1e-3) / 1e9 / peak_gbs}
        if world == 1 and not args.no_cpu:
            cb = cpu_baselines(w, None if args.config in (1, 2, 5) else max(1, w["iters"] // 10))
            line["cpu_baseline"] = cb
            line["speedup_vs_1thread"] = {"device_resident": value / cb["one_thread"]["value"], "e2e": e2e / cb["one_thread"]["value"]}
            line["speedup_vs_all_threads"] = {"device_resident": value / cb["all_threads"]["value"], "e2e": e2e / cb["all_threads"]["value"],
                                              "cores": cb["all_threads"]["cores"]}
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: a fixed total batch (config 3: 8192 buffers) is split over the ranks")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
