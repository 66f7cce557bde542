"""Whole-apply time per frame, device-resident, for a few (frames, patch, size) cases; run under different env
switches (RPSF_PDL=0/1 ...) to A/B a launch-level change."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import regularizepsf_b200 as rp
from regularizepsf_b200.device import DeviceCube


def run(B, P, HW, steps=40):
    coords = [tuple(int(v) for v in c) for c in rp.calculate_covering((HW, HW), P)]
    g = torch.Generator(device="cuda").manual_seed(1)
    kernel = torch.randn((len(coords), P, P), dtype=torch.complex64, device="cuda", generator=g)
    t = rp.ArrayPSFTransform(DeviceCube(coords, kernel))
    frames = torch.rand((B, HW, HW), device="cuda", generator=g) * 1000
    out = torch.empty_like(frames)
    for _ in range(5):
        t._apply_device(frames, "float32", 0, out=out)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            t._apply_device(frames, "float32", 0, out=out)
        e1.record(); torch.cuda.synchronize()
        best = min(best, 1e3 * e0.elapsed_time(e1) / steps / B)
    print(f"B={B} P={P} HW={HW}: {best:.2f} us/frame ({HW*HW/best:.0f} Mpix/s)  checksum {float(out.double().sum()):.6e}", flush=True)


for a in (sys.argv[1:] or ["1,256,2048", "8,256,2048", "1,128,1024", "8,128,1024", "1,512,8192", "1,64,1024"]):
    run(*[int(v) for v in a.split(",")])
