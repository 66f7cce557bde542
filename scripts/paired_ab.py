"""Paired vs classic column pass: stage times per frame (device-resident)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import regularizepsf_b200 as rp
from regularizepsf_b200 import _native
from regularizepsf_b200.device import DeviceCube

def run(B, P, HW, steps=20):
    coords = [tuple(int(v) for v in c) for c in rp.calculate_covering((HW, HW), P)]
    g = torch.Generator(device="cuda").manual_seed(1)
    kernel = torch.randn((len(coords), P, P), dtype=torch.complex64, device="cuda", generator=g)
    t = rp.ArrayPSFTransform(DeviceCube(coords, kernel))
    frames = torch.rand((B, HW, HW), device="cuda", generator=g) * 1000
    out = torch.empty_like(frames)
    nt = t._native_transform("float32")
    plan = nt.plan(HW, HW, 0, 0, HW, B)
    lib = _native.load()
    for mode, name in ((1, "classic"), (2, "paired")):
        if lib.rpsf_plan_set_column_mode(plan, mode) != 0:
            print(name, "unavailable"); continue
        for _ in range(4):
            t._apply_device(frames, "float32", 0, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            t._apply_device(frames, "float32", 0, out=out)
        e1.record(); torch.cuda.synchronize()
        tot = 1e3 * e0.elapsed_time(e1) / steps / B
        lib.rpsf_plan_enable_timing(plan, 1)
        for _ in range(steps):
            t._apply_device(frames, "float32", 0, out=out)
        torch.cuda.synchronize()
        ms = (ctypes.c_double * 3)(); calls = ctypes.c_int()
        lib.rpsf_plan_read_timing(plan, ms, ctypes.byref(calls))
        lib.rpsf_plan_enable_timing(plan, 0)
        per = [1e3 * ms[i] / calls.value / B for i in range(3)]
        print(f"B={B} P={P} HW={HW} {name:8s}: {tot:6.1f} us/frame  k1 {per[0]:.1f} k2 {per[1]:.1f} k3 {per[2]:.1f}  -> {HW*HW/tot:.0f} Mpix/s", flush=True)

for a in (sys.argv[1:] or ["8,256,2048", "1,256,2048", "2,256,2048", "32,256,2048", "8,128,1024", "1,128,1024", "8,64,1024"]):
    run(*[int(v) for v in a.split(",")])
