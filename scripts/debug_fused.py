"""Where does the fused slab output differ from the single-GPU output?  (torchrun, 2+ ranks)"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import regularizepsf_b200 as rp
from regularizepsf_b200 import distributed as rdist
from regularizepsf_b200.device import DeviceCube

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
for hw, patch in ((1024, 128), (4096, 256), (8192, 512)):
    coords = [tuple(int(v) for v in c) for c in rp.calculate_covering((hw, hw), patch)]
    g = torch.Generator(device="cuda").manual_seed(2)
    kernel = torch.randn((len(coords), patch, patch), dtype=torch.complex64, device="cuda", generator=g)
    t = rp.ArrayPSFTransform(DeviceCube(coords, kernel))
    image = torch.rand((hw, hw), device="cuda", generator=g) * 1000
    single = t.apply(image)
    fused = rdist.apply_slabs_fused(t, image).clone()
    torch.cuda.synchronize()
    bad = (fused != single)
    rows = torch.nonzero(bad.any(dim=1)).flatten()
    cols = torch.nonzero(bad.any(dim=0)).flatten()
    print(f"rank {rank} hw {hw} P {patch}: {int(bad.sum())} differing pixels; rows {rows[:6].tolist()}..{rows[-3:].tolist()} "
          f"({len(rows)}), cols {cols[:6].tolist()}..{cols[-3:].tolist()} ({len(cols)}); "
          f"max abs diff {float((fused - single).abs().max()):.3e}", flush=True)
    if int(bad.sum()):
        r = int(rows[0]); cs = torch.nonzero(bad[r]).flatten()
        print(f"   row {r}: cols {cs[:12].tolist()} fused {fused[r, cs[:4]].tolist()} single {single[r, cs[:4]].tolist()}", flush=True)
dist.barrier(); dist.destroy_process_group()
