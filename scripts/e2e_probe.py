"""Time the host-buffer apply (public API) per call on config 2, to see pipeline efficiency."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import regularizepsf_b200 as rp
from regularizepsf_b200.device import DeviceCube, pinned_empty

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
P, HW = 256, 2048
coords = [tuple(int(v) for v in c) for c in rp.calculate_covering((HW, HW), P)]
g = torch.Generator(device="cuda").manual_seed(1)
kernel = torch.randn((len(coords), P, P), dtype=torch.complex64, device="cuda", generator=g)
t = rp.ArrayPSFTransform(DeviceCube(coords, kernel))
frames = pinned_empty((B, HW, HW), np.float32)
frames[...] = np.random.default_rng(0).random((B, HW, HW), dtype=np.float32)
for chunk_bytes in (8 << 20, 32 << 20, 1 << 30):
    t.HOST_CHUNK_BYTES = chunk_bytes
    times = []
    for i in range(8):
        t0 = time.perf_counter()
        out = t.apply(frames)
        times.append(time.perf_counter() - t0)
    print("chunk_bytes", chunk_bytes, "ms per call:", " ".join(f"{1e3 * x:.2f}" for x in times),
          "-> Mpix/s steady", B * HW * HW / min(times) / 1e6)
# float32 output through the internal entry point
for i in range(4):
    t0 = time.perf_counter()
    out = t._apply_host(frames, "float32", 0, out_dtype=np.float32)
    dt = time.perf_counter() - t0
print("f32 out: ms", 1e3 * dt, "Mpix/s", B * HW * HW / dt / 1e6)
pageable = np.array(frames)
for i in range(4):
    t0 = time.perf_counter()
    out = t.apply(pageable)
    dt = time.perf_counter() - t0
print("pageable in: ms", 1e3 * dt, "Mpix/s", B * HW * HW / dt / 1e6)
# direct C-ABI calls with preallocated pinned buffers
import ctypes
from regularizepsf_b200 import _native
nt = t._native_transform("float32")
lib = nt.lib
for chunk in (1, 2):
    plan = nt.plan(HW, HW, 0, 0, HW, chunk)
    for od, code in ((np.float32, _native.F32), (np.float64, _native.F64)):
        outbuf = pinned_empty((B, HW, HW), od)
        ts = []
        for i in range(6):
            t0 = time.perf_counter()
            _native.check(lib.rpsf_apply_host(plan, frames.ctypes.data, _native.F32, outbuf.ctypes.data, code, B))
            ts.append(time.perf_counter() - t0)
        print("direct chunk", chunk, od.__name__, " ".join(f"{1e3 * x:.2f}" for x in ts))
