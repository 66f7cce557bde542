"""Probe host<->device copy bandwidth on the GPU box (pinned buffers), to size the e2e pipeline."""
import time
import torch

n = 256 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    print(name, 5 * n / (time.perf_counter() - t0) / 1e9, "GB/s")
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("duplex each way", 5 * n / dt / 1e9, "GB/s")
import os
print("cpus", os.cpu_count())
# host f32 -> f64 conversion speed
import numpy as np
a = np.ones((8, 2048, 2048), np.float32)
t0 = time.perf_counter(); b = a.astype(np.float64); print("astype f64 GB/s out", b.nbytes / (time.perf_counter() - t0) / 1e9)
t0 = time.perf_counter(); hp = torch.empty((8, 2048, 2048), dtype=torch.float64, pin_memory=True); print("pinned alloc 268MB s", time.perf_counter() - t0)
t0 = time.perf_counter(); hp2 = torch.empty((8, 2048, 2048), dtype=torch.float64, pin_memory=True); print("pinned alloc again s", time.perf_counter() - t0)
del hp
t0 = time.perf_counter(); hp3 = torch.empty((8, 2048, 2048), dtype=torch.float64, pin_memory=True); print("pinned alloc cached s", time.perf_counter() - t0)
