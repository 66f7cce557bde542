#!/usr/bin/env python
"""bench.py — Mpix/s corrected by ArrayPSFTransform.apply on 2048x2048 frames, 256-px patches.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames B]

One "step" = one apply() over a batch of B synthetic starfield frames (BASELINE.json config 2:
2048^2 PUNCH-WFI-like frame, 256-px patches, spatially varying coma source PSF -> Gaussian
target, alpha 1, epsilon 0.1).  For N > 1 (torchrun, one rank per GPU) every rank corrects its
own B frames with its own copy of the transform — frames are independent, so there is no
data-path collective (weak scaling); only the timing is reduced (max over ranks).

Printed JSON (rank 0, one line):
  value   device-resident throughput (frames already in HBM), CUDA events, max over ranks
  e2e     the same metric through the public API with HOST buffers (pinned numpy in, float64
          numpy out, H2D + D2H inside the timed region)
  roofline  dominant kernel (K2: column FFT x kernel x column IFFT) — algorithmic bytes per
          launch / its CUDA-event time inside the timed region, against the measured HBM peak;
          `kernels` carries the same figure for K1 and K3; `traffic` is the DRAM bytes per launch of
          the committed ncu capture (profiles/traffic.json) when it was taken at the same batch size
  cpu_baseline  the reference's own CPU apply (unmodified sources pip-installed into baseline/_ref
          by __graft_entry__.build(); kind "reference") or, if that copy is absent, the oracle port
          (bit-identical restatement; kind "port"), timed on this box's host cores with
          scipy.fft workers = all cores on a bounded sample of the same workload
`--impl reference` times that CPU path alone.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H = W = 2048
PATCH = 256
ALPHA, EPSILON = 1.0, 0.1
METRIC = "Mpix/s corrected (ArrayPSFTransform.apply, 2048^2, 256-px patches)"
WORKLOAD = "config2: 2048x2048 synthetic starfield frames, 256-px patches (289), coma source -> Gaussian target PSF"


def measured_peak_gbs() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel: str, frames: int):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the committed `ncu --set full`
    capture, if one exists for this kernel at this batch size; else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            table = json.load(f)
        entry = table[kernel]
        return int(entry["dram_bytes_per_launch"]) if int(entry["frames_per_launch"]) == frames else None
    except Exception:
        return None


def make_inputs(n_frames: int, seed0: int):
    from oracle import cpu_oracle as oracle   # input generators only; the oracle never computes on the product path
    import regularizepsf_b200 as rp
    coords = [tuple(int(v) for v in c) for c in rp.calculate_covering((H, W), PATCH)]
    src = oracle.coma_psf_cube(coords, PATCH, (H, W))
    tgt = oracle.gaussian_psf_cube(len(coords), PATCH, 3.0)
    base = oracle.starfield((H, W), seed=seed0)
    rng = np.random.default_rng(seed0 + 1)
    frames = np.stack([np.roll(base, (int(rng.integers(0, H)), int(rng.integers(0, W))), axis=(0, 1))
                       for _ in range(n_frames)])
    return coords, src, tgt, frames


class ClockSampler:
    """SM clock, power and throttle reasons of one GPU, polled through NVML (in-process, ~1 kHz) while
    the timed regions run; falls back to `nvidia-smi --query-gpu` (the recipe's clocks line) without pynvml."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int = 0):
        self.index, self.rows, self._stop, self._thread = index, [], threading.Event(), None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n, h = self._nvml, self._handle
        bits = int(n.nvmlDeviceGetCurrentClocksEventReasons(h))
        masks = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                 n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap]
        return [float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)), self._max,
                n.nvmlDeviceGetPowerUsage(h) / 1000.0] + ["active" if bits & m else "not active" for m in masks]

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [p.strip() for p in out.strip().split(",")]
        return parts if len(parts) >= 7 else None

    def _loop(self):
        while not self._stop.is_set():
            try:
                row = self._sample_nvml() if self._nvml else self._sample_smi()
                if row:
                    self.rows.append(row)
            except Exception:
                if self._nvml:
                    self._nvml = None          # e.g. an NVML build without the event-reason call
            self._stop.wait(0.001 if self._nvml else 0.1)

    def __enter__(self):
        self._stop.clear()
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self) -> dict:
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(str(r[3 + i]).lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows),
                "source": "nvml" if self._nvml else "nvidia-smi"}


def cpu_reference_run(frames: np.ndarray, coords, kernel, steps: int, warmup: int, workers: int):
    """Time the CPU path one frame per step; returns (times, kind).

    kind "reference": the UNMODIFIED reference's ArrayPSFTransform.apply (regularizepsf/transform.py:85-177),
    loaded from baseline/_ref (or /root/reference) by oracle/ref_loader.py; kind "port": the oracle
    restatement, bit-identical to it, when no copy of the reference is on the box.
    """
    from oracle import cpu_oracle as oracle
    from oracle import ref_loader
    if ref_loader.available():
        ref = ref_loader.load()
        transform = ref.transform.ArrayPSFTransform(ref.util.IndexedCube(coords, kernel))
        run, kind = (lambda frame: transform.apply(frame, workers=workers)), "reference"
    else:
        run, kind = (lambda frame: oracle.apply_transform(frame, coords, kernel, workers=workers)), "port"
    times = []
    for i in range(warmup + steps):
        frame = frames[i % len(frames)]
        t0 = time.perf_counter()
        run(frame)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, kind


def run_reference(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import cpu_oracle as oracle
    coords, src, tgt, frames = make_inputs(2, 1234)
    kernel = oracle.transfer_kernel(oracle.psf_fft(src.astype(np.float32)), oracle.psf_fft(tgt.astype(np.float32)),
                                    ALPHA, EPSILON)      # complex64 cube, as the reference's tests build it
    cores = os.cpu_count() or 1
    times, kind = cpu_reference_run(frames, coords, kernel, args.steps, args.warmup, cores)
    total = float(sum(times))
    value = args.steps * H * W / total / 1e6
    sample = f"{args.steps} single-frame apply() calls (1 frame of the batch per step), scipy.fft workers={cores}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": 1, "kernel_dtype": "complex64",
                   "note": ("unmodified reference ArrayPSFTransform.apply from baseline/_ref" if kind == "reference" else
                            "CPU oracle port of the pure-Python reference (bit-identical to it; "
                            "tests/test_oracle_vs_reference.py)")},
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args) -> int:
    import torch
    import torch.distributed as dist

    import regularizepsf_b200 as rp
    from regularizepsf_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    distributed = world > 1
    # several ranks share the host: keep each rank's pinned buffers on the socket its GPU hangs off
    # (N = 1 keeps every core for the CPU baseline)
    from regularizepsf_b200.distributed import bind_to_gpu_numa
    numa_cpus = bind_to_gpu_numa(local_rank) if distributed else None
    torch.cuda.set_device(local_rank)
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    B = args.frames

    coords, src, tgt, frames = make_inputs(B, 1234 + 17 * rank)
    source = rp.ArrayPSF(rp.IndexedCube(coords, src.astype(np.float32)))
    target = rp.ArrayPSF(rp.IndexedCube(coords, tgt.astype(np.float32)))
    transform = rp.ArrayPSFTransform.construct(source, target, ALPHA, EPSILON)
    lib = _native.load()

    # ---- parity gate before any timing: one frame against the oracle (rank 0)
    parity = None
    if rank == 0:
        from oracle import cpu_oracle as oracle
        kernel_host = transform._transfer_kernel.values
        want = oracle.apply_transform(frames[0], coords, kernel_host, workers=-1)
        got = transform.apply(frames[0])
        parity = float(np.max(np.abs(got - want)) / np.max(np.abs(frames[0])))
        if not parity <= 1e-5:
            raise SystemExit(f"parity gate failed: max|diff|/max|image| = {parity:.3e}")

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm
    dev_frames = torch.from_numpy(frames).cuda()
    dev_out = torch.empty_like(dev_frames)
    nt = transform._native_transform("float32")
    plan = nt.plan(H, W, 0, 0, H, B)
    for _ in range(args.warmup):
        transform._apply_device(dev_frames, "float32", 0, out=dev_out)
    barrier()
    lib.rpsf_plan_enable_timing(plan, 1)
    launches0 = _native.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local_rank)
    with clocks:
        barrier()
        start.record()
        for _ in range(args.steps):
            transform._apply_device(dev_frames, "float32", 0, out=dev_out)
        stop.record()
        barrier()
    launches = _native.launch_count() - launches0
    lib.rpsf_plan_enable_timing(plan, 0)
    ms_total = start.elapsed_time(stop)
    stage_ms = (ctypes.c_double * 3)()
    calls = ctypes.c_int()
    _native.check(lib.rpsf_plan_read_timing(plan, stage_ms, ctypes.byref(calls)))
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * H * W * args.steps / (ms_total * 1e-3) / 1e6

    # ---- end-to-end arm: public API, pinned host frames in, float64 host frames out
    from regularizepsf_b200.device import pinned_empty
    host_frames = pinned_empty(frames.shape, np.float32)
    host_frames[...] = frames
    # warm-up also fills torch's pinned-host cache: apply() returns a fresh pinned array per call,
    # and the first two calls pay cudaHostAlloc (~0.1 s for 268 MB) before the cache recycles blocks
    for _ in range(max(3, args.warmup)):
        out_host = transform.apply(host_frames)
    e2e_steps = args.steps
    with clocks:                                     # keep sampling clocks through this timed region too
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out_host = transform.apply(host_frames)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * B * H * W * e2e_steps / e2e_s / 1e6
    assert out_host.dtype == np.float64 and out_host.shape == frames.shape

    # secondary: the same call with the opt-in float32 result (half the device-to-host bytes)
    for _ in range(3):
        out32 = transform.apply(host_frames, out_dtype=np.float32)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out32 = transform.apply(host_frames, out_dtype=np.float32)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e32_value = world * B * H * W * e2e_steps / float(t.item()) / 1e6
    assert out32.dtype == np.float32 and np.array_equal(out32.astype(np.float64), out_host)

    if rank == 0:
        n_patches = len(coords)
        half = PATCH // 2
        # K2 algorithmic bytes per launch: read spectrum + write spectrum (in place) + read the
        # Hermitian-half kernel incl. its Nyquist column, complex64
        spec_bytes = B * n_patches * PATCH * half * 8
        kern_bytes = n_patches * PATCH * (half + 1) * 8
        k2_bytes = 2 * spec_bytes + kern_bytes
        k2_ms = stage_ms[1] / max(calls.value, 1)
        peak, peak_src = measured_peak_gbs()
        achieved = k2_bytes / (k2_ms * 1e-3) / 1e9
        # K1: unique frame bytes in + spectrum out; K3: spectrum in + frame out
        row_bytes = B * 4 * H * W + spec_bytes
        # whole apply, algorithmic: read frame + write frame + read kernel once per launch
        apply_bytes = B * 2 * 4 * H * W + kern_bytes
        per_stage = [stage_ms[i] / max(calls.value, 1) for i in range(3)]
        # CPU baseline on a bounded sample — at N = 1 only (at N > 1 the other ranks would idle behind it
        # and this rank's cores are pinned to one socket); `bench.py --impl reference` times it at every N
        cores = os.cpu_count() or 1
        if world == 1:
            kernel_host = transform._transfer_kernel.values
            n_cpu = args.cpu_frames
            cpu_times, cpu_kind = cpu_reference_run(frames, coords, kernel_host, n_cpu, 1, cores)
            cpu_baseline = {"value": n_cpu * H * W / float(sum(cpu_times)) / 1e6, "unit": "Mpix/s", "cores": cores,
                            "kind": cpu_kind, "sample": f"{n_cpu} single-frame apply() calls of the same frames, "
                                                        f"scipy.fft workers={cores}, after 1 warm-up"}
        else:
            cpu_baseline = {"value": None, "unit": "Mpix/s", "cores": cores, "kind": "reference",
                            "sample": "measured at N=1 only (see the N=1 line or --impl reference)"}
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": B, "alpha": ALPHA, "epsilon": EPSILON,
                       "l2": "no explicit flush: per-step working set (frames + spectrum workspace + kernel "
                             f"= {(2 * B * 4 * H * W + spec_bytes + kern_bytes) / 1e6:.0f} MB) exceeds the 126 MB L2",
                       "parallelism": f"frames sharded by rank (dp{world}), no data-path collective",
                       "parity_max_rel_err_vs_oracle": parity},
            "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": int(world * B * H * W * 4),
                    "d2h_bytes_per_step": int(world * B * H * W * 8), "steps": e2e_steps,
                    "bytes_note": "whole job (all ranks); each rank moves 1/n_gpus of it over its own PCIe link",
                    "host_numa_binding": (f"rank 0 pinned to {len(numa_cpus)} CPUs local to its GPU (NVML)"
                                          if numa_cpus else "none (single rank, or NVML reports no topology)"),
                    "api": "ArrayPSFTransform.apply(pinned float32 numpy (B,H,W)) -> float64 numpy",
                    "float32_out": {"value": e2e32_value, "unit": "Mpix/s",
                                    "d2h_bytes_per_step": int(world * B * H * W * 4),
                                    "api": "apply(..., out_dtype=np.float32): opt-in, same values, half the download"}},
            "gpu_launches": int(launches * world),
            "roofline": {"kernel": "k2_pipelined<256,float>", "bound": "hbm", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("k2_pipelined<256,float>", B),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": k2_bytes,
                         "ms_per_launch": k2_ms,
                         "stage_ms_per_step": {"k1": per_stage[0], "k2": per_stage[1], "k3": per_stage[2]},
                         "kernels": [
                             {"kernel": name, "algorithmic_bytes_per_launch": nbytes, "ms_per_launch": ms,
                              "achieved": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / peak,
                              "traffic": ncu_traffic(name, B)}
                             for name, nbytes, ms in (("k1_stream<256,float>", row_bytes, per_stage[0]),
                                                      ("k2_pipelined<256,float>", k2_bytes, per_stage[1]),
                                                      ("k3_stream<256,float>", row_bytes, per_stage[2]))],
                         "whole_apply": {"algorithmic_bytes_per_step": apply_bytes,
                                         "achieved_gbs": apply_bytes / (ms_total / args.steps * 1e-3) / 1e9,
                                         "frac": apply_bytes / (ms_total / args.steps * 1e-3) / 1e9 / peak}},
            "cpu_baseline": cpu_baseline,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=8, help="frames per step per GPU")
    ap.add_argument("--cpu-frames", type=int, default=10, help="frames in the bounded cpu_baseline sample")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
