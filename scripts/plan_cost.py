"""Cost of the first call of a geometry (plan creation: work lists, tables, uploads) against a steady-state call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import regularizepsf_b200 as rp
from regularizepsf_b200.device import DeviceCube
for B, P, HW in ((1, 128, 1024), (8, 256, 2048), (1, 512, 8192), (1, 32, 2048)):
    coords = [tuple(int(v) for v in c) for c in rp.calculate_covering((HW, HW), P)]
    g = torch.Generator(device="cuda").manual_seed(1)
    kernel = torch.randn((len(coords), P, P), dtype=torch.complex64, device="cuda", generator=g)
    frames = torch.rand((B, HW, HW), device="cuda", generator=g)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t = rp.ArrayPSFTransform(DeviceCube(coords, kernel))
    nt = t._native_transform("float32")
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    plan = nt.plan(HW, HW, 0, 0, HW, B)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    t.apply(frames)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    t.apply(frames)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print(f"B={B} P={P} HW={HW} patches={len(coords)}: transform {1e3*(t1-t0):.2f} ms, plan {1e3*(t2-t1):.2f} ms, "
          f"first apply {1e3*(t3-t2):.2f} ms, second apply {1e3*(t4-t3):.2f} ms", flush=True)
