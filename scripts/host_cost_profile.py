"""Where the host time of one device-resident apply() call goes (cProfile), and the native call alone."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import regularizepsf_b200 as rp
from regularizepsf_b200.device import DeviceCube
P, HW = 128, 1024
coords = [tuple(int(v) for v in c) for c in rp.calculate_covering((HW, HW), P)]
g = torch.Generator(device="cuda").manual_seed(1)
kernel = torch.randn((len(coords), P, P), dtype=torch.complex64, device="cuda", generator=g)
t = rp.ArrayPSFTransform(DeviceCube(coords, kernel))
frame = torch.rand((HW, HW), device="cuda", generator=g)
for _ in range(5): t.apply(frame)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(2000): t.apply(frame)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
