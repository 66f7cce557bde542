"""Round-2 builder stages on the device against the same stages through scipy on the host (the oracle's restatement of
image_processing.py:78-121 and builder.py:236-258).  Prints one JSON object."""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import cpu_oracle as oracle
from regularizepsf_b200 import builder as b

warnings.simplefilter("ignore", RuntimeWarning)
rng = np.random.default_rng(0)
frame = oracle.starfield((2048, 2048), seed=7, density=1 / 400).astype(np.float64)
out = {}
for width, n in ((32, 20000), (64, 5000)):
    keys = [(0, float(r), float(c)) for r, c in zip(rng.uniform(-width / 2, 2048 - width / 2, n), rng.uniform(-width / 2, 2048 - width / 2, n))]
    b.cutouts_at(frame, keys[:64], width, star_maximum=1e9)              # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter(); got = b.cutouts_at(frame, keys, width, star_maximum=1e9); t1 = time.perf_counter()
    sample = keys[:300]
    t2 = time.perf_counter(); oracle.cutouts_at(frame, sample, width, star_maximum=1e9); t3 = time.perf_counter()
    out[f"star_cutouts_{width}px"] = {"stars": n, "device_ms_incl_copies": 1e3 * (t1 - t0), "device_us_per_star": 1e6 * (t1 - t0) / n,
                                      "host_scipy_us_per_star": 1e6 * (t3 - t2) / len(sample), "accepted": len(got)}
for size, n in ((32, 1089), (64, 1089), (128, 289)):
    r, c = np.indices((size, size))
    stack = np.exp(-((r - size // 2) ** 2 + (c - size // 2) ** 2) / 18.0)[None] + rng.normal(0.1, 0.01, (n, size, size))
    b.isolate_cores(stack[:8]); torch.cuda.synchronize()
    t0 = time.perf_counter(); b.isolate_cores(stack); t1 = time.perf_counter()
    m = min(n, 100)
    t2 = time.perf_counter(); [oracle.isolate_core(p.copy()) for p in stack[:m]]; t3 = time.perf_counter()
    out[f"isolate_cores_{size}px"] = {"patches": n, "device_ms_incl_copies": 1e3 * (t1 - t0), "device_us_per_patch": 1e6 * (t1 - t0) / n,
                                      "host_scipy_us_per_patch": 1e6 * (t3 - t2) / m}
print(json.dumps(out, indent=1))
