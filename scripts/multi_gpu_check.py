"""Multi-GPU check on real hardware (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/multi_gpu_check.py [--big]

* config 3 (BASELINE.json): a batch of frames sharded by rank, one shared transform, outputs gathered
  on rank 0 over NCCL; rank 0 compares with its own single-GPU result, bit for bit.
* config 4: one large frame split into patch-row slabs with a one-patch halo, bands all-gathered;
  every rank compares the stitched frame with the single-GPU result, bit for bit, and the
  slab compute + all-gather is timed on the device (max over ranks).
Prints one JSON line per config on rank 0.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import regularizepsf_b200 as rp
from regularizepsf_b200 import distributed as rdist
from regularizepsf_b200.device import DeviceCube


def make_transform(shape, patch, seed):
    coords = [tuple(int(v) for v in c) for c in rp.calculate_covering(shape, patch)]
    g = torch.Generator(device="cuda").manual_seed(seed)
    kernel = torch.randn((len(coords), patch, patch), dtype=torch.complex64, device="cuda", generator=g)
    return rp.ArrayPSFTransform(DeviceCube(coords, kernel))


def timed(fn, steps, world):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    big = "--big" in sys.argv
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    # ---- config 3: frames sharded by rank, gather on rank 0
    hw, patch, per_rank = 2048, 256, 8
    n_frames = per_rank * world
    transform = make_transform((hw, hw), patch, seed=1)         # same seed -> same kernel on every rank
    g = torch.Generator(device="cuda").manual_seed(7)
    frames = torch.rand((n_frames, hw, hw), device="cuda", generator=g) * 1000
    full = rdist.apply_frames_sharded(transform, frames, gather=True)
    ok3 = True
    if rank == 0:
        want = torch.cat([transform.apply(frames[i:i + per_rank]) for i in range(0, n_frames, per_rank)])
        ok3 = bool(torch.equal(full, want))
    ms_compute = timed(lambda: rdist.apply_frames_sharded(transform, frames, gather=False), 10, world)
    ms_gather = timed(lambda: rdist.apply_frames_sharded(transform, frames, gather=True), 10, world)
    fused3 = rdist.apply_frames_fused(transform, frames)
    torch.cuda.synchronize()
    ok3f = bool(torch.equal(fused3, want)) if rank == 0 else True
    ms_fused3 = timed(lambda: rdist.apply_frames_fused(transform, frames), 10, world)
    by_chunk = {}
    for transport in ("stores", "copy"):
        for chunk in (0, 1, 2, 4):                               # 0 = the whole block in one launch (round 1)
            got = rdist.apply_frames_fused(transform, frames, chunk_frames=chunk, transport=transport)
            torch.cuda.synchronize()
            if rank == 0:
                ok3f = ok3f and bool(torch.equal(got, want))
            by_chunk[f"{transport}/{chunk}"] = timed(
                lambda: rdist.apply_frames_fused(transform, frames, chunk_frames=chunk, transport=transport), 10, world)
    failures = [] if (ok3 and ok3f) else ["config 3"]
    if rank == 0:
        root_ingress_mb = (world - 1) * per_rank * hw * hw * 4 / 1e6
        print(json.dumps({"config": 3, "world": world, "frames": n_frames, "bit_identical_to_single_gpu": ok3,
                          "ms_compute_only": ms_compute, "ms_with_nccl_gather": ms_gather,
                          "fused_bit_identical_to_single_gpu": ok3f, "ms_fused_peer_stores": ms_fused3,
                          "ms_fused_by_chunk_frames": by_chunk, "root_ingress_mb": root_ingress_mb,
                          "ms_root_ingress_at_770_gbs": root_ingress_mb / 770.0,
                          "mpix_s_fused": n_frames * hw * hw / ms_fused3 / 1e3,
                          "mpix_s_compute_only": n_frames * hw * hw / ms_compute / 1e3,
                          "mpix_s_with_gather": n_frames * hw * hw / ms_gather / 1e3}), flush=True)
    del frames, full, transform, fused3
    rdist._peer_frames.clear()
    torch.cuda.empty_cache()

    # ---- config 4: one mosaic, patch-row slabs + all-gather
    hw, patch = (8192, 512) if big else (4096, 256)
    transform = make_transform((hw, hw), patch, seed=2)
    g = torch.Generator(device="cuda").manual_seed(9)
    image = torch.rand((hw, hw), device="cuda", generator=g) * 1000
    stitched = rdist.apply_slabs_sharded(transform, image)
    single = transform.apply(image)
    ok4 = torch.tensor([int(torch.equal(stitched, single))], device="cuda")
    dist.all_reduce(ok4, op=dist.ReduceOp.MIN)
    ms_slab = timed(lambda: rdist.apply_slabs_sharded(transform, image), 10, world)
    ms_single = timed(lambda: transform.apply(image), 10, world)
    # the same slabs with the gather fused into the overlap-add kernel (peer stores over NVLink, no all-gather)
    fused = rdist.apply_slabs_fused(transform, image)
    torch.cuda.synchronize()
    ok4f = torch.tensor([int(torch.equal(fused, single))], device="cuda")
    dist.all_reduce(ok4f, op=dist.ReduceOp.MIN)
    ms_fused = timed(lambda: rdist.apply_slabs_fused(transform, image), 10, world)
    # the gather modes that do not send 7 copies of the frame around, with the transform sharded (each rank holds only
    # the kernels of its band + halo) and only the frame rows the rank reads resident
    lo, hi = rdist.slab_bounds(hw, patch, world)[rank]
    shard = rdist.shard_transform_rows(transform, hw, rank, world)
    first, last = rdist.rows_needed(transform.coordinates, patch, hw, (lo, hi))
    rows = image[first:last].clone()
    modes = {}
    ok_modes = True
    for gather in ("root", "none"):
        got = rdist.apply_slabs_fused(shard, rows, gather=gather, frame_rows=(first, hw))
        torch.cuda.synchronize()
        if gather == "root" and rank == 0:
            ok_modes = ok_modes and bool(torch.equal(got, single))
        else:
            ok_modes = ok_modes and bool(torch.equal(got, single[lo:hi]))
        modes["ms_fused_gather_" + gather] = timed(
            lambda: rdist.apply_slabs_fused(shard, rows, gather=gather, frame_rows=(first, hw)), 10, world)
        if gather == "root":
            for transport, subs in (("copy", 1), ("copy", 2), ("copy", 4), ("stores", 2)):
                got = rdist.apply_slabs_fused(shard, rows, gather="root", frame_rows=(first, hw), transport=transport, sub_bands=subs)
                torch.cuda.synchronize()
                ok_modes = ok_modes and bool(torch.equal(got, single if rank == 0 else single[lo:hi]))
                modes[f"ms_fused_gather_root_{transport}_{subs}"] = timed(
                    lambda: rdist.apply_slabs_fused(shard, rows, gather="root", frame_rows=(first, hw), transport=transport,
                                                    sub_bands=subs), 10, world)
        got = rdist.apply_slabs_sharded(transform, image, gather=gather)
        torch.cuda.synchronize()
        ok_modes = ok_modes and bool(torch.equal(got, single if (gather == "root" and rank == 0) else single[lo:hi]))
        modes["ms_nccl_gather_" + gather] = timed(lambda: rdist.apply_slabs_sharded(transform, image, gather=gather), 10, world)
    okm = torch.tensor([int(ok_modes)], device="cuda")
    dist.all_reduce(okm, op=dist.ReduceOp.MIN)
    kernel_mb = [len(transform) * patch * patch * 8 / 1e6, len(shard) * patch * patch * 8 / 1e6]
    if not (ok4.item() and ok4f.item() and okm.item()):
        failures.append("config 4")
    if rank == 0:
        print(json.dumps({"config": 4, "world": world, "frame": [hw, hw], "patch": patch,
                          "bit_identical_to_single_gpu": bool(ok4.item()),
                          "ms_slabs_plus_all_gather": ms_slab, "ms_single_gpu": ms_single,
                          "fused_bit_identical_to_single_gpu": bool(ok4f.item()), "ms_slabs_fused_peer_stores": ms_fused,
                          "sharded_modes_bit_identical": bool(okm.item()), **modes,
                          "kernel_cube_mb_full_vs_rank0_shard": kernel_mb, "frame_rows_held_by_rank0": [first, last],
                          "mpix_s_slabs_fused": hw * hw / ms_fused / 1e3,
                          "mpix_s_slabs": hw * hw / ms_slab / 1e3, "mpix_s_single_gpu": hw * hw / ms_single / 1e3}),
              flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if failures and "--assert" in sys.argv:
        raise SystemExit("not bit-identical to the single-GPU result: " + ", ".join(failures))


if __name__ == "__main__":
    main()
