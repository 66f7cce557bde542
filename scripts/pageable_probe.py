"""apply() on an ordinary (pageable) float32 numpy batch: time per call for a few RPSF_HOST_COPY_THREADS settings (set
in the environment before the first call; the library reads it per call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import regularizepsf_b200 as rp
from regularizepsf_b200.device import DeviceCube

HW, P, B = 2048, 256, 8
coords = [tuple(int(v) for v in c) for c in rp.calculate_covering((HW, HW), P)]
g = torch.Generator(device="cuda").manual_seed(1)
kernel = torch.randn((len(coords), P, P), dtype=torch.complex64, device="cuda", generator=g)
t = rp.ArrayPSFTransform(DeviceCube(coords, kernel))
frames = np.random.default_rng(0).random((B, HW, HW), dtype=np.float32)
for threads in (sys.argv[1:] or ["1", "4", "8", "12", "16"]):
    os.environ["RPSF_HOST_COPY_THREADS"] = threads
    for _ in range(3):
        out = t.apply(frames)
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        out = t.apply(frames)
    dt = (time.perf_counter() - t0) / n
    print(f"threads={threads}: {1e3 * dt:.2f} ms per call of {B} frames = {B * HW * HW / dt / 1e6:.0f} Mpix/s", flush=True)
pinned = torch.from_numpy(frames).pin_memory().numpy()
for _ in range(3):
    out = t.apply(pinned)
t0 = time.perf_counter()
for _ in range(10):
    out = t.apply(pinned)
dt = (time.perf_counter() - t0) / 10
print(f"pinned input: {1e3 * dt:.2f} ms per call = {B * HW * HW / dt / 1e6:.0f} Mpix/s")
one = frames[0]
for _ in range(3):
    out = t.apply(one)
t0 = time.perf_counter()
for _ in range(20):
    out = t.apply(one)
print(f"single pageable float32 frame: {1e3 * (time.perf_counter() - t0) / 20:.3f} ms per call")
u16 = (one * 60000).astype(np.uint16)
for _ in range(3):
    out = t.apply(u16)
t0 = time.perf_counter()
for _ in range(20):
    out = t.apply(u16)
print(f"single pageable uint16 frame: {1e3 * (time.perf_counter() - t0) / 20:.3f} ms per call")
