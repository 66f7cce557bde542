"""Single-CTA patch path (P <= 128) vs the three kernels: time per frame, device-resident."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import regularizepsf_b200 as rp
from regularizepsf_b200 import _native
from regularizepsf_b200.device import DeviceCube

def run(B, P, HW, steps=30):
    coords = [tuple(int(v) for v in c) for c in rp.calculate_covering((HW, HW), P)]
    g = torch.Generator(device="cuda").manual_seed(1)
    kernel = torch.randn((len(coords), P, P), dtype=torch.complex64, device="cuda", generator=g)
    t = rp.ArrayPSFTransform(DeviceCube(coords, kernel))
    frames = torch.rand((B, HW, HW), device="cuda", generator=g) * 1000
    out = torch.empty_like(frames)
    nt = t._native_transform("float32")
    plan = nt.plan(HW, HW, 0, 0, HW, B)
    lib = _native.load()
    res = {}
    for mode, name in ((1, "three kernels"), (2, "single CTA")):
        if lib.rpsf_plan_set_small_mode(plan, mode) != 0:
            print(name, "unavailable"); continue
        for _ in range(4):
            t._apply_device(frames, "float32", 0, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            t._apply_device(frames, "float32", 0, out=out)
        e1.record(); torch.cuda.synchronize()
        res[name] = 1e3 * e0.elapsed_time(e1) / steps / B
    print(f"B={B} P={P} HW={HW}: " + ", ".join(f"{k} {v:.1f} us/frame ({HW*HW/v:.0f} Mpix/s)" for k, v in res.items()), flush=True)

for a in (sys.argv[1:] or ["8,128,1024", "1,128,1024", "32,128,1024", "8,64,1024", "1,64,1024", "8,32,512", "8,128,2048", "8,64,2048"]):
    run(*[int(v) for v in a.split(",")])
