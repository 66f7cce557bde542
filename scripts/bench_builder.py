"""Device time of rpsf_average_patches (builder.py:45-125 on the GPU) next to numpy on the host cores.

    python scripts/bench_builder.py [n_cutouts] [patch] [image]

Synthetic stack: `n_cutouts` star cutouts of patch^2 float64 (2 % NaN) with uniformly random centres on an
image^2 frame, assigned to calculate_covering cells exactly as ArrayPSFBuilder does (each cutout joins the
~4 cells its centre falls in).  Algorithmic bytes = every cell reads its members once + writes one patch.
The numpy baseline (oracle.average_cutouts, the reference's arithmetic) is timed on a sample of cells.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import regularizepsf_b200 as rp
from oracle import cpu_oracle as oracle
from regularizepsf_b200 import _native
from regularizepsf_b200 import builder as b

M = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 32
HW = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
rng = np.random.default_rng(0)
corners = rp.calculate_covering((HW, HW), P)
keys = [(0, float(r), float(c)) for r, c in zip(rng.uniform(-P / 2, HW - P / 2, M), rng.uniform(-P / 2, HW - P / 2, M))]
offsets, items = b.assign_to_cells(keys, corners, P)
stack = rng.normal(1.0, 0.3, size=(M, P, P))
stack[rng.random(stack.shape) < 0.02] = np.nan
depth = np.diff(offsets)
lib = _native.load()
dev = torch.from_numpy(stack).cuda()
out = torch.empty((len(corners), P, P), dtype=torch.float64, device="cuda")
alg_bytes = (len(items) + len(corners)) * P * P * 8
sample = np.sort(rng.choice(len(corners), size=min(64, len(corners)), replace=False))
s_off = np.concatenate([[0], np.cumsum(depth[sample])]).astype(np.int64)
s_items = np.concatenate([items[offsets[c]:offsets[c + 1]] for c in sample]).astype(np.int32)
rows = []
for method, pct in (("mean", 50.0), ("median", 50.0), ("percentile", 30.0)):
    code = {"mean": 0, "median": 1, "percentile": 2}[method]
    def run():
        _native.check(lib.rpsf_average_patches(dev.data_ptr(), M, P, offsets.ctypes.data, items.ctypes.data, len(corners),
                                               code, pct, out.data_ptr(), 0, _native.current_stream_ptr(torch)))
    run(); run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    t0 = time.perf_counter()
    want = oracle.average_cutouts(stack, s_off, s_items, method, pct)
    cpu_s = (time.perf_counter() - t0) * len(corners) / len(sample)
    assert np.array_equal(out.cpu().numpy()[sample], want), method
    rows.append({"method": method, "percentile": pct, "gpu_ms": round(ms, 3), "numpy_s_extrapolated": round(cpu_s, 2),
                 "speedup": round(cpu_s * 1e3 / ms, 1), "algorithmic_GBps": round(alg_bytes / ms / 1e6, 1)})
print(json.dumps({"n_cutouts": M, "patch": P, "image": HW, "cells": len(corners), "stack_depth_min_med_max":
                  [int(depth.min()), int(np.median(depth)), int(depth.max())], "algorithmic_bytes": alg_bytes,
                  "note": "gpu_ms includes the call's own index upload, temporaries and stream sync; parity checked bit-exact "
                          "on the sampled cells", "rows": rows}))
