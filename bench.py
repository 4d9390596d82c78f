#!/usr/bin/env python
"""bench.py -- headline benchmark of the STFT -> |X| -> NMF hot path (BASELINE.json metric: NMF frames/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): per GPU, batch=1024 synthetic float32 buffers x 130816 samples (F=512 frames each),
fft=win=1024, hop=256, rank K=16, 200 multiplicative-update iterations.  One "step" = one BufNMF pass over that batch:
STFT -> magnitude -> W/H init from per-buffer seeds -> 200 iterations -> bases + activations out.

  value  : frames/s with the audio already resident in HBM and outputs left in HBM (device pointers through the C ABI)
  e2e    : the same call with HOST (pinned) buffers: H2D of the audio and D2H of bases/activations inside the timed region
  roofline / cpu_baseline : see DESIGN.md "Measurement"
  --impl reference : the CPU restatement of the reference (oracle/, fp64, faithful mode, all host threads) on a
                     bounded sample of the same workload (the reference itself needs Eigen/HISSTools: not buildable here)

Multi-GPU: buffers are independent, so each rank processes its own 1024-buffer shard (weak scaling) with no data-path
collective; the final activations are all-gathered with NCCL inside the timed region (north_star).
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

WORKLOAD = dict(name="config2: batch=1024 x 130816 samples (512 frames), fft=1024 hop=256 rank=16 iters=200",
                batch=1024, n=130816, win=1024, fft=1024, hop=256, rank=16, iters=200)
METRIC = "NMF frames/sec (batch x frames) at rank=16, fft=1024"
UNIT = "frames/s"


def make_audio(batch, n, base_seed=1000, distinct=16):
    """Config-2 style synthetic buffers.  `distinct` different signals tiled over the batch (generating
    1024 distinct ones on the host would take longer than the benchmark); every buffer still gets its own NMF seed."""
    from tests.golden.make_golden import synth_audio
    uniq = np.stack([synth_audio(base_seed + i, n) for i in range(min(distinct, batch))])
    reps = (batch + uniq.shape[0] - 1) // uniq.shape[0]
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
        return d.get("bf16_tflops_sustained", 1344.4), d.get("hbm_gbs", 6539.9), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


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


def cpu_sample(co, lib_path, threads, w, nbuf):
    audio = make_audio(nbuf, w["n"])
    seeds = np.arange(nbuf, dtype=np.int64)
    t0 = time.perf_counter()
    co.bufnmf_batch(audio, w["win"], w["fft"], w["hop"], w["rank"], w["iters"], seeds, resynth=False, faithful=True,
                    threads=threads, lib_path=lib_path)
    dt = time.perf_counter() - t0
    F = co.num_frames(w["n"], w["win"], w["hop"])
    return nbuf * F / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (restated in oracle/, fp64, 7 GEMMs/iter as NMF.hpp:144-183 runs
    them) with one buffer per host thread; each step is a bounded sample of the config-2 workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOAD
    co, lib_path = native_oracle()
    threads = min(os.cpu_count() or 1, 64)
    nbuf = threads
    F = co.num_frames(w["n"], w["win"], w["hop"])
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_sample(co, lib_path, threads, w, nbuf)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sample(co, lib_path, threads, w, nbuf)
    dt = time.perf_counter() - t0
    value = args.steps * nbuf * F / dt
    sample = f"{nbuf} buffers of the config-2 workload per step, one per thread ({threads} threads), fp64 faithful mode"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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

    w = WORKLOAD
    batch, n, K, iters = w["batch"], w["n"], w["rank"], w["iters"]
    plan = fb.Plan(win=w["win"], hop=w["hop"], fft=w["fft"], device=local, max_rank=K, max_batch=batch, max_samples=n)
    F, B = fb.num_frames(n, w["win"], w["hop"]), plan.bins
    # each rank owns the shard [rank*batch, (rank+1)*batch) of the global batch: distinct NMF seeds per buffer
    seeds = np.arange(rank * batch, (rank + 1) * batch, dtype=np.int64)
    audio_h = torch.from_numpy(make_audio(batch, n, base_seed=1000 + 16 * rank)).pin_memory()
    audio_d = audio_h.cuda()
    out_d = {"bases": torch.empty((batch, K, B), dtype=torch.float32, device="cuda"),
             "acts": torch.empty((batch, F, K), dtype=torch.float32, device="cuda")}
    out_h = {"bases": torch.empty((batch, K, B), dtype=torch.float32).pin_memory().numpy(),
             "acts": torch.empty((batch, F, K), dtype=torch.float32).pin_memory().numpy()}
    gathered = torch.empty((world * batch, F, K), dtype=torch.float32, device="cuda") if world > 1 else None

    def step_device():
        plan.bufnmf(audio_d, K, iters, seeds=seeds, out=out_d)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out_d["acts"])  # north_star: one allgather of the final activations
        return plan.stats()

    def step_host():
        plan.bufnmf(audio_h.numpy(), K, iters, seeds=seeds, out=out_h)
        return plan.stats()

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

    frames_per_step = world * batch * F
    value = frames_per_step * args.steps / (ms * 1e-3)
    e2e = frames_per_step * args.steps / (ms_e2e * 1e-3)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "per_gpu_batch": batch, "l2": "inputs (536 MB audio, 1.1 GB |X|) exceed the 126 MB L2",
                       "parallelism": f"batch-sharded x{world}, allgather(H) at the end" if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(batch * n * 4 + batch * 8),
                    "d2h_bytes_per_step": int(batch * K * B * 4 + batch * F * K * 4)},
            "gpu_launches": int(agg["launches"]),
            "stages_ms_per_step": {k: agg[k] / args.steps for k in ("ms_stft", "ms_nmf", "ms_total")}}

    if rank == 0:
        # roofline of the dominant kernel: the NMF update (tile) kernel.  Algorithmic flops per frame per iteration =
        # 8*B*K; one fused launch does one H-update + one W-numerator = one iteration's worth for batch*F
        # frames; the first/last launches of a step do half each, so a step's launches sum to exactly `iters` iterations.
        peak_tf, _, which = measured_peaks()
        flops_step = 8.0 * B * K * iters * batch * F
        ms_kernel = agg["ms_update_kernel"] / args.steps
        achieved = flops_step / (ms_kernel * 1e-3) / 1e12 if ms_kernel > 0 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "update_kernel_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        kname = {1: "k_nmf_tile (fp32 SIMT)", 2: "k_nmf_tc (tcgen05, split-bf16 operands, fp32 accumulate)"}.get(agg["backend"], "?")
        line["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                            "frac": (achieved / peak_tf) if achieved else None, "traffic": traffic,
                            "kernel": kname, "peak_source": which,
                            "note": "achieved counts the algorithmic 8*B*K flops per frame per iteration; the tensor pipe "
                                    "executes 7x that (6-term operand split + 4-term ratio split), see DESIGN.md",
                            "launches_per_step": agg["update_kernel_launches"] / args.steps,
                            "avg_launch_ms": ms_kernel / max(1, agg["update_kernel_launches"] / args.steps),
                            "kernel_share_of_step": ms_kernel / (agg["ms_total"] / args.steps)}
        if world == 1 and not args.no_cpu:
            co, lib_path = native_oracle()
            threads = min(os.cpu_count() or 1, 64)
            v, dt = cpu_sample(co, lib_path, threads, w, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{threads} config-2 buffers, one per thread, fp64 faithful mode (7 GEMMs/iter), {dt:.1f} s wall"}
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
