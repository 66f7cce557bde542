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


def _host_threads() -> dict:
    """What decides the CPU arm's speed on this box: cores the process may use and thread-count overrides."""
    try:
        affinity = len(os.sched_getaffinity(0))
    except AttributeError:
        affinity = os.cpu_count() or 1
    return {"cpu_count": os.cpu_count() or 1, "affinity": affinity,
            "OMP_NUM_THREADS": os.environ.get("OMP_NUM_THREADS"), "MKL_NUM_THREADS": os.environ.get("MKL_NUM_THREADS")}


def run_reference(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if os.environ.get("RPSF_REFERENCE_CHILD") != "1":
        # torchrun exports OMP_NUM_THREADS=1 (and friends) to every rank; round 1's reference arm ran at half speed
        # under it.  Time the reference in a child with the launcher's thread and rendezvous variables removed and
        # the full CPU affinity, so the number does not depend on how bench.py was started.
        env = {k: v for k, v in os.environ.items()
               if not (k.endswith("_NUM_THREADS") or k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE",
                                                            "GROUP_RANK", "ROLE_RANK", "MASTER_ADDR", "MASTER_PORT",
                                                            "TORCHELASTIC_RUN_ID", "KMP_AFFINITY", "GOMP_CPU_AFFINITY"))}
        env["RPSF_REFERENCE_CHILD"] = "1"
        env["RPSF_LAUNCHER_ENV"] = json.dumps(_host_threads())
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))
        except Exception:
            pass
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--gpus", str(args.gpus), "--steps",
               str(args.steps), "--warmup", str(args.warmup)]
        return subprocess.run(cmd, env=env).returncode
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
                            "tests/test_oracle_vs_reference.py)"),
                   "host_threads": _host_threads(),
                   "launcher_host_threads": json.loads(os.environ.get("RPSF_LAUNCHER_ENV", "null"))},
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0



# ---------------------------------------------------------------------------- extra records (round 2)
def _random_transform(torch, rp, shape, patch, seed, dtype=None):
    """A covering with a random complex kernel built on the device (timing only: parity is the tests' business)."""
    from regularizepsf_b200.device import DeviceCube
    coords = [tuple(int(v) for v in c) for c in rp.calculate_covering(shape, patch)]
    g = torch.Generator(device="cuda").manual_seed(seed)
    kernel = torch.randn((len(coords), patch, patch), dtype=dtype or torch.complex64, device="cuda", generator=g)
    return coords, rp.ArrayPSFTransform(DeviceCube(coords, kernel))


def device_case(torch, rp, lib, shape, patch, frames_per_call, dtype_name, steps, warmup, peak):
    """Device-resident apply() of one configuration: time per frame, per-kernel CUDA-event times inside the timed
    region and their fraction of the measured HBM peak (algorithmic bytes: DESIGN.md section 4)."""
    from regularizepsf_b200 import _native
    h, w = shape
    cdt = torch.complex64 if dtype_name == "float32" else torch.complex128
    rdt = torch.float32 if dtype_name == "float32" else torch.float64
    coords, transform = _random_transform(torch, rp, shape, patch, seed=patch + h, dtype=cdt)
    g = torch.Generator(device="cuda").manual_seed(3)
    frames = (torch.rand((frames_per_call, h, w), device="cuda", generator=g) * 1000).to(rdt)
    out = torch.empty_like(frames)
    nt = transform._native_transform(dtype_name)
    plan = nt.plan(h, w, 0, 0, h, frames_per_call)
    for _ in range(warmup):
        transform._apply_device(frames, dtype_name, 0, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()                                      # the whole call, without the per-kernel event records in between
    for _ in range(steps):
        transform._apply_device(frames, dtype_name, 0, out=out)
    e1.record()
    torch.cuda.synchronize()
    lib.rpsf_plan_enable_timing(plan, 1)             # the same calls again with CUDA events around each kernel
    for _ in range(steps):
        transform._apply_device(frames, dtype_name, 0, out=out)
    torch.cuda.synchronize()
    lib.rpsf_plan_enable_timing(plan, 0)
    ms = (ctypes.c_double * 3)()
    calls = ctypes.c_int()
    _native.check(lib.rpsf_plan_read_timing(plan, ms, ctypes.byref(calls)))
    total_ms = e0.elapsed_time(e1) / steps
    s = 4 if dtype_name == "float32" else 8
    n, half = len(coords), patch // 2
    spec = frames_per_call * n * patch * half * 2 * s
    kern = n * patch * (half + 1) * 2 * s
    frame_bytes = frames_per_call * h * w * s
    alg = {"k1": frame_bytes + spec, "k2": 2 * spec + kern, "k3": spec + frame_bytes}
    col = (ctypes.c_int64 * 2)()
    _native.check(lib.rpsf_plan_column_info(plan, col))
    if col[0]:      # paired column pass: K2 writes, and K3 reads, the band sums (about half a spectrum)
        pair_bytes = frames_per_call * int(col[1])
        alg = {"k1": frame_bytes + spec, "k2": spec + pair_bytes + kern, "k3": pair_bytes + frame_bytes}
    rec = {"shape": [h, w], "patch": patch, "patches": n, "frames_per_call": frames_per_call, "dtype": dtype_name,
           "us_per_frame": 1e3 * total_ms / frames_per_call, "mpix_s": frames_per_call * h * w / total_ms / 1e3,
           "column_pass": "paired (k2_chain + k3_stream_paired)" if col[0] else "classic (k2_pipelined + k3_stream)",
           "kernels": {}}
    for i, name in enumerate(("k1", "k2", "k3")):
        k_ms = ms[i] / max(calls.value, 1)
        rec["kernels"][name] = {"us_per_launch": 1e3 * k_ms, "algorithmic_bytes": alg[name],
                                "frac_of_hbm_peak": alg[name] / (k_ms * 1e-3) / 1e9 / peak if k_ms > 0 else None}
    apply_bytes = 2 * frame_bytes + kern
    rec["whole_apply_frac_of_hbm_peak"] = apply_bytes / (total_ms * 1e-3) / 1e9 / peak
    # fp32 / fp64 pipe view: 5 n log2 n per complex FFT, r2c / c2r counted as half
    import math
    fft = 5 * patch * math.log2(patch)
    flops = frames_per_call * n * (patch / 2 * fft * 2 + (half + 1) * fft * 2 + 8 * patch * half)
    rec["algorithmic_tflops"] = flops / (total_ms * 1e-3) / 1e12
    del transform, frames, out
    torch.cuda.empty_cache()
    return rec


def e2e_variants(transform, frames: np.ndarray, steps: int) -> dict:
    """The calls a drop-in user makes: pageable arrays, one frame per call, integer pixels (all return float64)."""
    import torch
    out = {}

    def clock(fn, n):
        fn(); fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n

    b, h, w = frames.shape
    pageable = np.array(frames, dtype=np.float32, copy=True)
    single = pageable[0].copy()
    single_u16 = np.clip(single, 0, 65535).astype(np.uint16)
    cases = {
        "pageable_float32_batch": (lambda: transform.apply(pageable), b),
        "pageable_float32_single_frame": (lambda: transform.apply(single), 1),
        "pageable_uint16_single_frame": (lambda: transform.apply(single_u16), 1),
    }
    for name, (fn, n_frames) in cases.items():
        sec = clock(fn, max(3, steps // 2))
        out[name] = {"value": n_frames * h * w / sec / 1e6, "unit": "Mpix/s", "ms_per_call": 1e3 * sec,
                     "frames_per_call": n_frames}
    return out


def host_copy_ceiling(torch, dist, distributed: bool, frames_per_step: int, steps: int) -> dict:
    """Copy-only probe on every rank at once: the pinned H2D (float32 frames) and D2H (float64 frames) traffic of the
    e2e arm with no kernels in between, both directions concurrently.  What the host side of this box can move is the
    ceiling of the e2e number at N ranks."""
    n_in, n_out = frames_per_step * H * W * 4, frames_per_step * H * W * 8
    h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def step():
        with torch.cuda.stream(s_in):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_out.copy_(d_out, non_blocking=True)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    t = torch.tensor([sec], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    world = dist.get_world_size() if distributed else 1
    return {"value": world * frames_per_step * H * W * steps / sec / 1e6, "unit": "Mpix/s",
            "h2d_plus_d2h_gb_s_per_rank": (n_in + n_out) * steps / sec / 1e9,
            "what": "pinned H2D of the float32 frames + D2H of the float64 results on two streams, all ranks at once, no kernels"}


def multi_gpu_records(torch, dist, rp, rank: int, world: int, steps: int) -> dict:
    """The two splits BASELINE.json names, on this job's N ranks (N > 1): config 3 — 8 frames per rank corrected and
    GATHERED on rank 0 (NCCL gather, and peer stores fused into the overlap-add kernel); config 4 — one 8192^2 / 512-px
    frame in patch-row slabs (all-gather over NCCL; fused peer stores to every rank / to the root only / no exchange,
    the last two with the kernel cube sharded and only the needed frame rows resident).  Every result is compared with
    the single-GPU result bit for bit; times are CUDA events, max over ranks."""
    from regularizepsf_b200 import distributed as rdist

    def timed(fn):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(flag: bool) -> bool:
        t = torch.tensor([int(flag)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    rec = {}
    # ---- config 3
    per_rank = 8
    n_frames = per_rank * world
    _, transform = _random_transform(torch, rp, (H, W), PATCH, seed=1)       # same seed: same kernel on every rank
    g = torch.Generator(device="cuda").manual_seed(7)
    frames = torch.rand((n_frames, H, W), device="cuda", generator=g) * 1000
    want = None
    if rank == 0:
        want = torch.cat([transform.apply(frames[i:i + per_rank]) for i in range(0, n_frames, per_rank)])
    got = rdist.apply_frames_sharded(transform, frames, gather=True)
    ok_nccl = all_true(rank != 0 or bool(torch.equal(got, want)))
    got = rdist.apply_frames_fused(transform, frames)
    torch.cuda.synchronize()
    ok_fused = all_true(rank != 0 or bool(torch.equal(got, want)))
    del got, want
    ms_compute = timed(lambda: rdist.apply_frames_sharded(transform, frames, gather=False))
    ms_nccl = timed(lambda: rdist.apply_frames_sharded(transform, frames, gather=True))
    ms_fused = timed(lambda: rdist.apply_frames_fused(transform, frames))
    ingress_mb = (world - 1) * per_rank * H * W * 4 / 1e6
    mpix = n_frames * H * W / 1e6
    rec["config3_gather"] = {
        "frames": n_frames, "frames_per_rank": per_rank, "bit_identical": ok_nccl and ok_fused,
        "compute_only": {"ms": ms_compute, "mpix_s": mpix / ms_compute * 1e3},
        "nccl_gather": {"ms": ms_nccl, "mpix_s": mpix / ms_nccl * 1e3},
        "fused_peer_stores": {"ms": ms_fused, "mpix_s": mpix / ms_fused * 1e3},
        "root_ingress_mb": ingress_mb,
        "ms_link_bound": ingress_mb / 770.0,
        "link_bound_note": "the root receives (N-1) blocks of float32 frames; 770 GB/s = measured peer-copy rate per direction (B200_PROFILING.md)",
    }
    del frames, transform
    rdist._peer_frames.clear()
    torch.cuda.empty_cache()
    # ---- config 4
    hw, patch = 8192, 512
    _, transform = _random_transform(torch, rp, (hw, hw), patch, seed=2)
    g = torch.Generator(device="cuda").manual_seed(9)
    image = torch.rand((hw, hw), device="cuda", generator=g) * 1000
    single = transform.apply(image)
    lo, hi = rdist.slab_bounds(hw, patch, world)[rank]
    shard = rdist.shard_transform_rows(transform, hw, rank, world)
    first, last = rdist.rows_needed(transform.coordinates, patch, hw, (lo, hi))
    rows = image[first:last].clone()
    runs = {
        "nccl_all_gather": lambda: rdist.apply_slabs_sharded(transform, image),
        "fused_all": lambda: rdist.apply_slabs_fused(transform, image),
        "fused_root_sharded_cube": lambda: rdist.apply_slabs_fused(shard, rows, gather="root", frame_rows=(first, hw)),
        "no_exchange_sharded_cube": lambda: rdist.apply_slabs_fused(shard, rows, gather="none", frame_rows=(first, hw)),
    }
    ok = True
    for name, fn in runs.items():
        got = fn()
        torch.cuda.synchronize()
        whole = name in ("nccl_all_gather", "fused_all") or (name == "fused_root_sharded_cube" and rank == 0)
        ok = ok and bool(torch.equal(got, single if whole else single[lo:hi]))
        del got
    ok = all_true(ok)
    ms_single = timed(lambda: transform.apply(image))
    times = {name: timed(fn) for name, fn in runs.items()}
    rec["config4_slabs"] = {
        "frame": [hw, hw], "patch": patch, "bit_identical": ok, "single_gpu_ms": ms_single,
        "modes": {name: {"ms": ms, "speedup_vs_single_gpu": ms_single / ms, "mpix_s": hw * hw / ms / 1e3}
                  for name, ms in times.items()},
        "kernel_cube_mb": {"complete": len(transform) * patch * patch * 8 / 1e6,
                           "rank0_shard": len(shard) * patch * patch * 8 / 1e6},
        "frame_rows_resident_rank0": [first, last],
    }
    del image, rows, single, shard, transform
    rdist._peer_frames.clear()
    torch.cuda.empty_cache()
    return rec


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
    from regularizepsf_b200.distributed import bind_to_gpu_numa, numa_report
    numa_cpus = bind_to_gpu_numa(local_rank) if distributed else None
    numa_why = numa_report(local_rank)
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

    # ---- what the host side of this box can move at N ranks (copy-only), and the calls a drop-in user makes
    ceiling = host_copy_ceiling(torch, dist, distributed, B, e2e_steps)
    peak, peak_src = measured_peak_gbs()
    user_calls, other_configs, fp64 = None, None, None
    if world == 1:
        user_calls = e2e_variants(transform, frames, e2e_steps)
        del dev_frames, dev_out
        torch.cuda.empty_cache()
        ksteps = max(5, args.steps // 2)
        other_configs = {
            "config2_single_frame": device_case(torch, rp, lib, (H, W), PATCH, 1, "float32", ksteps, 3, peak),
            "config1_batch8": device_case(torch, rp, lib, (1024, 1024), 128, 8, "float32", ksteps, 3, peak),
            "config1_single_frame": device_case(torch, rp, lib, (1024, 1024), 128, 1, "float32", ksteps, 3, peak),
            "config4_single_frame": device_case(torch, rp, lib, (8192, 8192), 512, 1, "float32", ksteps, 3, peak),
        }
        fp64 = device_case(torch, rp, lib, (H, W), PATCH, B, "float64", max(3, ksteps // 2), 3, peak)
        fp64["note"] = ("float64 validation mode (1e-10 parity): the same kernels instantiated for double.  Every buffer is twice "
                        "as wide, and the kernels stay HBM-bound: `algorithmic_tflops` is far below the ~37 TFLOP/s of the "
                        "B200's non-tensor fp64 pipe, while the per-kernel fractions of the HBM peak match the float32 ones")
    multi = multi_gpu_records(torch, dist, rp, rank, world, max(5, args.steps // 2)) if distributed else None

    if rank == 0:
        n_patches = len(coords)
        half = PATCH // 2
        # K2 algorithmic bytes per launch: read spectrum + write spectrum (in place) + read the
        # Hermitian-half kernel incl. its Nyquist column, complex64
        spec_bytes = B * n_patches * PATCH * half * 8
        kern_bytes = n_patches * PATCH * (half + 1) * 8
        k2_bytes = 2 * spec_bytes + kern_bytes
        col = (ctypes.c_int64 * 2)()
        _native.check(lib.rpsf_plan_column_info(plan, col))
        paired = bool(col[0])
        pair_bytes = B * int(col[1])
        if paired:      # the column pass writes the band sums of overlapping patch rows instead of the spectrum
            k2_bytes = spec_bytes + pair_bytes + kern_bytes
        names = (("k1_stream<256,float>", "k2_chain<256,float>", "k3_stream_paired<256,float>") if paired else
                 ("k1_stream<256,float>", "k2_pipelined<256,float>", "k3_stream<256,float>"))
        k2_ms = stage_ms[1] / max(calls.value, 1)
        achieved = k2_bytes / (k2_ms * 1e-3) / 1e9
        # K1: unique frame bytes in + spectrum out; K3: spectrum (or band sums) in + frame out
        row_bytes = B * 4 * H * W + spec_bytes
        k3_bytes = B * 4 * H * W + (pair_bytes if paired else spec_bytes)
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
                                          if numa_cpus else f"none: {numa_why}"),
                    "host_ceiling": ceiling,
                    "frac_of_host_ceiling": e2e_value / ceiling["value"] if ceiling["value"] else None,
                    "user_calls": user_calls,
                    "api": "ArrayPSFTransform.apply(pinned float32 numpy (B,H,W)) -> float64 numpy",
                    "float32_out": {"value": e2e32_value, "unit": "Mpix/s",
                                    "d2h_bytes_per_step": int(world * B * H * W * 4),
                                    "api": "apply(..., out_dtype=np.float32): opt-in, same values, half the download"}},
            "gpu_launches": int(launches * world),
            "roofline": {"kernel": names[1], "bound": "hbm", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(names[1], B),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": k2_bytes,
                         "ms_per_launch": k2_ms,
                         "stage_ms_per_step": {"k1": per_stage[0], "k2": per_stage[1], "k3": per_stage[2]},
                         "kernels": [
                             {"kernel": name, "algorithmic_bytes_per_launch": nbytes, "ms_per_launch": ms,
                              "achieved": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / peak,
                              "traffic": ncu_traffic(name, B)}
                             for name, nbytes, ms in ((names[0], row_bytes, per_stage[0]),
                                                      (names[1], k2_bytes, per_stage[1]),
                                                      (names[2], k3_bytes, per_stage[2]))],
                         "column_pass": "paired" if paired else "classic",
                         "whole_apply": {"algorithmic_bytes_per_step": apply_bytes,
                                         "achieved_gbs": apply_bytes / (ms_total / args.steps * 1e-3) / 1e9,
                                         "frac": apply_bytes / (ms_total / args.steps * 1e-3) / 1e9 / peak}},
            "cpu_baseline": cpu_baseline,
            "clocks": clocks.summary(),
            "configs": other_configs,
            "fp64": fp64,
            "multi_gpu": multi,
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
