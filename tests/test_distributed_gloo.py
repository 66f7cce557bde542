"""Host-side sharding logic over torch.distributed with the gloo backend (CPU, world_size > 1).

The data path has no collective; the only exchange is the gather of corrected frames (config 3)
or the all-gather of row slabs (config 4).  These tests run that plumbing with 2 and 3 CPU
ranks; the "corrected" data is a stand-in computed from the input so that ordering, ragged
shard sizes and empty shards are all visible in the result.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from regularizepsf_b200 import distributed as rdist


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, case: str, queue) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if case == "frames":
            n_frames = 5
            frames = np.arange(n_frames * 6 * 4, dtype=np.float64).reshape(n_frames, 6, 4)
            b, e = rdist.frame_shard(n_frames, rank, world)
            local = frames[b:e] * 2.0 + 1.0          # stand-in for the corrected block
            full = rdist.gather_frames(local, n_frames)
            if rank == 0:
                queue.put(("frames", np.array_equal(full, frames * 2.0 + 1.0)))
            else:
                queue.put(("frames", full is None))
        elif case == "slabs":
            height, width, patch = 40, 7, 16
            frame = np.arange(height * width, dtype=np.float32).reshape(height, width)
            bounds = rdist.slab_bounds(height, patch, world)
            b, e = bounds[rank]
            full = rdist.all_gather_slabs(frame[b:e] - 3.0, bounds)
            queue.put(("slabs", np.array_equal(full, frame - 3.0)))
        elif case == "empty":
            # fewer half-patch rows than ranks -> some rank owns an empty band
            height, width, patch = 8, 5, 16
            frame = np.arange(height * width, dtype=np.float32).reshape(height, width)
            bounds = rdist.slab_bounds(height, patch, world)
            b, e = bounds[rank]
            full = rdist.all_gather_slabs(frame[b:e], bounds)
            queue.put(("empty", np.array_equal(full, frame) and any(hi == lo for lo, hi in bounds)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["frames", "slabs", "empty"])
def test_gather_plumbing_over_gloo(world, case):
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in results), results


@pytest.mark.parametrize("n_frames,world", [(64, 8), (5, 2), (3, 4), (0, 2), (7, 7)])
def test_frame_shards_partition_the_batch(n_frames, world):
    spans = [rdist.frame_shard(n_frames, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n_frames
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    sizes = [e - b for b, e in spans]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("height,patch,world", [(8192, 512, 8), (8192, 512, 4), (2048, 256, 2), (100, 32, 3), (8, 16, 4)])
def test_slab_bounds_are_aligned_and_cover(height, patch, world):
    bounds = rdist.slab_bounds(height, patch, world)
    assert bounds[0][0] == 0 and bounds[-1][1] == height
    assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
    assert all(b % (patch // 2) == 0 for b, _ in bounds)


def test_numa_binding_is_a_harmless_no_op_without_nvml_topology():
    import os
    from regularizepsf_b200.distributed import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    cpus = bind_to_gpu_numa(0)                      # no GPU / no NVML here: must not raise, must not shrink to nothing
    after = os.sched_getaffinity(0)
    assert cpus is None or (len(cpus) > 0 and set(cpus) == after)
    os.sched_setaffinity(0, before)
