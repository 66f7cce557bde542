"""Host-side issue cost of a device-resident apply() call against its total time (is the GPU being starved?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import regularizepsf_b200 as rp
from regularizepsf_b200.device import DeviceCube
for B, P, HW in ((1,32,512),(1,128,1024),(1,256,2048)):
    coords = [tuple(int(v) for v in c) for c in rp.calculate_covering((HW, HW), P)]
    g = torch.Generator(device="cuda").manual_seed(1)
    kernel = torch.randn((len(coords), P, P), dtype=torch.complex64, device="cuda", generator=g)
    t = rp.ArrayPSFTransform(DeviceCube(coords, kernel))
    frames = torch.rand((B, HW, HW), device="cuda", generator=g)
    out = torch.empty_like(frames)
    for _ in range(5): t._apply_device(frames, "float32", 0, out=out)
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n): t._apply_device(frames, "float32", 0, out=out)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"P={P} HW={HW}: host issue {1e6*(t1-t0)/n:.1f} us/call, total {1e6*(t2-t0)/n:.1f} us/call")
    # public apply() path
    t0 = time.perf_counter()
    for _ in range(n): t.apply(frames[0])
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"   apply(tensor): host issue {1e6*(t1-t0)/n:.1f} us/call, total {1e6*(t2-t0)/n:.1f} us/call")
