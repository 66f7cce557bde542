"""Minimal driver for ncu: a few device-resident apply() calls on config 2 (batch of frames)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import regularizepsf_b200 as rp
from regularizepsf_b200.device import DeviceCube

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
P = int(sys.argv[3]) if len(sys.argv) > 3 else 256
HW = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
coords = [tuple(int(v) for v in c) for c in rp.calculate_covering((HW, HW), P)]
g = torch.Generator(device="cuda").manual_seed(1)
kernel = torch.randn((len(coords), P, P), dtype=torch.complex64, device="cuda", generator=g)
t = rp.ArrayPSFTransform(DeviceCube(coords, kernel))
frames = torch.rand((B, HW, HW), device="cuda", generator=g) * 1000
out = torch.empty_like(frames)
for _ in range(REPS):
    t._apply_device(frames, "float32", 0, out=out)
torch.cuda.synchronize()
print("done", float(out[0, 0, 0]))
