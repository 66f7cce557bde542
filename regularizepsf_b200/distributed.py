"""Multi-GPU sharding of the correction path: one process per GPU, torch.distributed for plumbing.

The path shards two ways and neither needs a reduction:

* by frame  — frames are independent; every rank holds the whole transform and corrects its
  own block of frames; the only collective is the optional gather of outputs;
* by patch-row slab — one huge frame; rank r owns a contiguous band of output rows aligned to
  P/2, computes every patch that intersects the band (patch rows on a boundary are computed by
  both neighbours: the one-patch halo) and reads only the frame rows those patches touch; the
  only collective is the all-gather of the bands.  Stitched output is bit-identical to the
  single-GPU result because the per-pixel summation order (colour order) does not change.
"""
from __future__ import annotations

import os

import numpy as np


def bind_to_gpu_numa(device_index: int) -> list[int] | None:
    """Pin this process to the CPUs NVML reports as local to GPU ``device_index``.

    The host path of ``apply`` moves 50 MB per frame through pinned buffers; on a multi-socket box a rank
    whose buffers were first touched on the other socket pays for every byte twice.  Call this once per
    rank BEFORE anything allocates pinned memory (Linux places pages on the node of the touching CPU).
    Returns the CPU list, or None when NVML has no topology to offer (a no-op then).
    """
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        allowed = sorted(os.sched_getaffinity(0))
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (max(allowed) // 64) + 1)
        local = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in local if c in set(allowed)]
        if not cpus or len(cpus) == len(allowed):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001 - NVML missing / VM without topology: leave the affinity alone
        return None


def numa_report(device_index: int) -> str:
    """Why ``bind_to_gpu_numa`` did or did not narrow the affinity: what NVML says about this GPU's local CPUs."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        allowed = sorted(os.sched_getaffinity(0))
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (max(allowed) // 64) + 1)
        local = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        nodes = "?"
        try:
            nodes = str(len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]))
        except OSError:
            pass
        return (f"NVML lists {len(local)} CPUs local to GPU {device_index}, the process may use {len(allowed)}; "
                f"{nodes} NUMA node(s) visible" + (": every CPU is local, nothing to bind" if len(local) >= len(allowed) else ""))
    except Exception as exc:  # noqa: BLE001
        return f"NVML unavailable ({type(exc).__name__})"


def frame_shard(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [begin, end) of frames for ``rank``; sizes differ by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_frames, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def slab_bounds(height: int, patch_size: int, world: int) -> list[tuple[int, int]]:
    """Output-row bands [begin, end) for ``world`` ranks, boundaries aligned to patch_size/2.

    Bands are as even as the alignment allows; trailing ranks may get an empty band when the
    frame has fewer half-patch rows than ranks.
    """
    if world < 1:
        raise ValueError("world must be >= 1")
    step = max(patch_size // 2, 1)
    units = -(-height // step)
    cuts = [min(height, ((units * r) // world) * step) for r in range(world)] + [height]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_frames(local: np.ndarray, n_frames: int, group=None) -> np.ndarray | None:
    """Gather per-rank frame blocks (frame_shard order) on rank 0; returns None elsewhere."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    tensor = local if hasattr(local, "is_cuda") else torch.from_numpy(np.ascontiguousarray(local))
    counts = [frame_shard(n_frames, r, world) for r in range(world)]
    most = max(e - b for b, e in counts)
    padded = tensor.new_zeros((most,) + tuple(tensor.shape[1:]))
    padded[: tensor.shape[0]] = tensor
    bucket = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, bucket, dst=0, group=group)
    if rank != 0:
        return None
    parts = [bucket[r][: counts[r][1] - counts[r][0]] for r in range(world)]
    out = torch.cat(parts, dim=0)
    return out if hasattr(local, "is_cuda") else out.numpy()


def all_gather_slabs(local_band, bounds: list[tuple[int, int]], group=None):
    """All-gather row bands (slab_bounds order) into the full frame on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    is_tensor = hasattr(local_band, "is_cuda")
    tensor = local_band if is_tensor else torch.from_numpy(np.ascontiguousarray(local_band))
    most = max(e - b for b, e in bounds)
    padded = tensor.new_zeros((most,) + tuple(tensor.shape[1:]))
    padded[: tensor.shape[0]] = tensor
    bucket = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bucket, padded, group=group)
    out = torch.cat([bucket[r][: bounds[r][1] - bounds[r][0]] for r in range(world)], dim=0)
    return out if is_tensor else out.numpy()


def gather_slabs(local_band, bounds: list[tuple[int, int]], root: int = 0, group=None):
    """Gather row bands (slab_bounds order) into the full frame on ``root``; returns None elsewhere."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    is_tensor = hasattr(local_band, "is_cuda")
    tensor = local_band if is_tensor else torch.from_numpy(np.ascontiguousarray(local_band))
    most = max(e - b for b, e in bounds)
    padded = tensor.new_zeros((most,) + tuple(tensor.shape[1:]))
    padded[: tensor.shape[0]] = tensor
    bucket = [torch.empty_like(padded) for _ in range(world)] if rank == root else None
    dist.gather(padded, bucket, dst=root, group=group)
    if rank != root:
        return None
    out = torch.cat([bucket[r][: bounds[r][1] - bounds[r][0]] for r in range(world)], dim=0)
    return out if is_tensor else out.numpy()


def apply_frames_sharded(transform, frames, *, gather: bool = True, group=None, **apply_kwargs):
    """Correct a batch of frames split by ``frame_shard`` over the ranks of ``group``.

    Every rank passes the same ``frames`` array (or at least its own block); each corrects only
    its block.  With ``gather`` rank 0 gets the full (B, H, W) result, other ranks ``None``;
    without it every rank gets its own block.
    """
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    begin, end = frame_shard(len(frames), rank, world)
    local = transform.apply(frames[begin:end], **apply_kwargs) if end > begin else frames[:0]
    return gather_frames(local, len(frames), group) if gather else local


class PeerFrames:
    """One full-frame output buffer per rank, each mapped into every other rank over NVLink (CUDA IPC).

    With these, the overlap-add kernel of a slab stores its band straight into every rank's frame while
    it runs (``rpsf_plan_set_output_mirrors``) and the all-gather disappears.  Ranks of one node only
    (at most 8).  ``tensor`` is this rank's buffer; it is overwritten by the next fused call.  Ranks may
    allocate different shapes (the fused frame gather gives only the root a full batch).
    """

    def __init__(self, shape, torch_dtype, group=None):
        import ctypes

        import torch
        import torch.distributed as dist

        from regularizepsf_b200 import _native

        self._lib, self._group = _native.load(), group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError("PeerFrames maps the frames of at most 8 ranks of one node")
        self.device = torch.cuda.current_device()
        self.shape, self.dtype = tuple(int(v) for v in shape), torch_dtype
        itemsize = torch.empty((), dtype=torch_dtype).element_size()
        nbytes = itemsize * int(np.prod(self.shape))
        own, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        _native.check(self._lib.rpsf_ipc_alloc(ctypes.byref(own), nbytes, self.device, handle))
        self.base = int(own.value)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self.ptrs = []                                    # device pointer of every rank's buffer, by rank
        for r, raw in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.base)
                continue
            ptr = ctypes.c_void_p()
            _native.check(self._lib.rpsf_ipc_open(ctypes.byref(ptr), raw, self.device))
            self.ptrs.append(int(ptr.value))
        self.peers = [p for r, p in enumerate(self.ptrs) if r != self.rank]   # the other ranks' buffers
        typestr = {torch.float32: "<f4", torch.float64: "<f8"}[torch_dtype]
        holder = type("_Frame", (), {})()
        holder.__cuda_array_interface__ = {"shape": self.shape, "typestr": typestr, "data": (self.base, False),
                                           "version": 2}
        self._holder = holder
        self.tensor = torch.as_tensor(holder, device=f"cuda:{self.device}")
        self._flag = torch.zeros(1, dtype=torch.int32, device=f"cuda:{self.device}")
        dist.barrier(group)

    def peer_tensor(self, rank: int, shape):
        """A tensor over rank ``rank``'s buffer as mapped into this process (for copy-engine transfers)."""
        import torch

        key = (rank, tuple(int(v) for v in shape))
        cache = self.__dict__.setdefault("_peer_tensors", {})
        if key not in cache:
            typestr = {torch.float32: "<f4", torch.float64: "<f8"}[self.dtype]
            holder = type("_PeerFrame", (), {})()
            holder.__cuda_array_interface__ = {"shape": key[1], "typestr": typestr, "data": (self.ptrs[rank], False), "version": 2}
            cache[key] = (holder, torch.as_tensor(holder, device=f"cuda:{self.device}"))
        return cache[key][1]

    def copy_stream(self):
        import torch

        if "_copy_stream" not in self.__dict__:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        return self._copy_stream

    def stream_barrier(self):
        """Order the ranks ON THE STREAM (a one-element all-reduce): no rank's later kernels start before every
        rank's earlier ones have finished, and the host is not blocked."""
        import torch.distributed as dist

        dist.all_reduce(self._flag, group=self._group)

    def close(self):
        import torch.distributed as dist

        if self.base is None:
            return
        for ptr in self.peers:
            _native_check(self._lib.rpsf_ipc_close(ptr, self.device))
        dist.barrier(self._group)                         # nobody maps this rank's frame any more
        _native_check(self._lib.rpsf_device_free(self.base, self.device))
        self.peers, self.ptrs, self.base, self.tensor = [], [], None, None


def _native_check(rc):
    from regularizepsf_b200 import _native
    _native.check(rc)


_peer_frames: dict = {}


def rows_needed(coordinates, patch_size: int, height: int, band: tuple[int, int], pad_mode: str = "symmetric") -> tuple[int, int]:
    """Frame rows [first, last) that the patches intersecting the output band ``band`` read (np.pad index map of
    ``pad_mode`` applied at the frame's top and bottom, transform.py:119-123): what a rank of a row-slab split must
    hold of the frame.  (The plan reports the same range: ``plan_info()["rows_read"]``.)"""
    from regularizepsf_b200 import _native

    if band[1] <= band[0]:
        return (0, 0)
    lib, code = _native.load(), _native.PAD_MODES[pad_mode]
    lo, hi = height, 0
    for corner in sorted({int(c[0]) for c in np.asarray(coordinates).reshape(-1, 2)}):
        if corner >= band[1] or corner + patch_size <= band[0]:
            continue
        for r in range(corner, corner + patch_size):
            y = r if 0 <= r < height else lib.rpsf_pad_index(r, height, code)
            if y >= 0:
                lo, hi = min(lo, y), max(hi, y + 1)
    return (lo, hi) if lo < hi else (0, 0)


def shard_transform_rows(transform, height: int, rank: int, world: int):
    """The part of ``transform`` that rank ``rank`` of a ``world``-way patch-row slab split needs: the kernels of the
    patches that intersect its band (``slab_bounds``), i.e. about 1 / world of the cube plus the one-patch halo.
    Returns an ``ArrayPSFTransform`` that can only be applied to that band (``row_range``); results are bit-identical
    to the complete transform's."""
    from regularizepsf_b200.device import DeviceCube
    from regularizepsf_b200.transform import ArrayPSFTransform
    from regularizepsf_b200.util import IndexedCube

    patch = transform.psf_shape[0]
    lo, hi = slab_bounds(height, patch, world)[rank]
    full = np.asarray(transform.coordinates).reshape(-1, 2)
    keep = (full[:, 0] < hi) & (full[:, 0] + patch > lo) if hi > lo else np.zeros(len(full), dtype=bool)
    coords = [tuple(int(v) for v in c) for c in full[keep]]
    cube = transform._transfer_kernel
    if isinstance(cube, DeviceCube):
        import torch
        sub = DeviceCube(coords, cube.tensor[torch.from_numpy(np.flatnonzero(keep)).to(cube.tensor.device)])
    else:
        sub = IndexedCube(coords, np.ascontiguousarray(cube.values[keep]))
    return ArrayPSFTransform.sharded(sub, full, keep)


def apply_slabs_fused(transform, image, *, group=None, pad_mode: str = "symmetric", dtype=None, gather: str = "all",
                      root: int = 0, frame_rows: tuple[int, int] | None = None, transport: str | None = None,
                      sub_bands: int | None = None):
    """Patch-row slabs with the output gather fused into the overlap-add kernel (no collective on the data path).

    Every rank computes its band (one-patch halo, as ``apply_slabs_sharded``).  ``gather``:

    * ``"all"``  — the kernel writes the band into the full frame of EVERY rank through peer-mapped memory while it
      runs; every rank returns the full frame (a buffer reused by the next call with the same shape);
    * ``"root"`` — only ``root``'s frame receives the bands (1 / world of the NVLink traffic per rank: a rank that
      goes on working on its own band, or a single consumer, does not need 7 copies of the frame); ``root``
      returns the full frame, the others their band;
    * ``"none"`` — nothing is exchanged: every rank returns its band.

    With ``gather="root"``, ``transport="copy"`` makes the kernel write locally and lets a copy engine carry the band to
    the root on a second stream, and ``sub_bands = k`` corrects the band in k pieces (each with its own halo, so some
    arithmetic is repeated) so that the first piece travels while the next is computed — the root's ingress (the
    other ranks' bands) is longer than a rank's arithmetic at 8 GPUs, so starting it early is what shortens the call.
    Defaults (None): peer stores and one piece up to 4 ranks; for ``gather="root"`` beyond 4 ranks the copy engine and
    two pieces (measured on 8 B200: 0.56 -> 0.48 ms for the 8192^2 frame).

    ``image`` is a 2-D CUDA tensor: the whole frame, or — with ``frame_rows = (first, height)`` — only the rows
    [first, first + image.shape[0]) of it (``rows_needed`` says which rows a rank reads).  ``transform`` may be a
    ``shard_transform_rows`` shard.  Two stream-ordered barriers (one-element all-reduces) order the ranks when bands
    travel.  Bit-identical to the single-GPU result.
    """
    import torch
    import torch.distributed as dist

    from regularizepsf_b200 import _native
    from regularizepsf_b200.transform import _normalize_dtype

    if not (hasattr(image, "is_cuda") and image.is_cuda and image.dim() == 2):
        raise ValueError("apply_slabs_fused needs one 2-D CUDA tensor (the frame, or the rows of it this rank holds)")
    if gather not in ("all", "root", "none"):
        raise ValueError(f"gather must be 'all', 'root' or 'none', got {gather!r}")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    name = _normalize_dtype(dtype)
    want = torch.float32 if name == "float32" else torch.float64
    height = frame_rows[1] if frame_rows is not None else int(image.shape[0])
    width = int(image.shape[1])
    lo, hi = slab_bounds(height, transform.psf_shape[0], world)[rank]
    code = _native.PAD_MODES[pad_mode]
    if gather == "none":
        return transform._apply_device(image, name, code, row_range=(lo, hi), frame_rows=frame_rows)
    if transport is None:
        transport = "copy" if (gather == "root" and world > 4) else "stores"
    if sub_bands is None:
        sub_bands = 2 if (gather == "root" and world > 4) else 1
    if transport not in ("stores", "copy") or (transport == "copy" and gather != "root"):
        raise ValueError("transport must be 'stores', or 'copy' together with gather='root'")
    key = ("slabs", gather, root, id(group), (height, width), want, torch.cuda.current_device())
    frames = _peer_frames.get(key)
    if frames is None:
        full_here = gather == "all" or rank == root
        frames = _peer_frames[key] = PeerFrames((height, width) if full_here else (max(hi - lo, 1), width), want, group)
    frames.stream_barrier()                               # everyone is done reading the previous result
    full_here = gather == "all" or rank == root
    if hi > lo:                                           # a rank past the last half-patch row owns nothing
        band = frames.tensor[lo:hi] if full_here else frames.tensor[: hi - lo]
        step = max(transform.psf_shape[0] // 2, 1)
        units = -(-(hi - lo) // step)
        k = max(1, min(int(sub_bands), units))
        cuts = [lo + min(hi - lo, ((units * i) // k) * step) for i in range(k)] + [hi]
        side = frames.copy_stream() if (transport == "copy" and rank != root) else None
        main = torch.cuda.current_stream()
        root_view = frames.peer_tensor(root, (height, width)) if side is not None else None
        if side is not None:
            side.wait_stream(main)                        # the barrier above orders the copies too
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b <= a:
                continue
            offset = a * width * frames.tensor.element_size()
            if gather == "all":
                mirrors = [p + offset for p in frames.peers]
            else:
                mirrors = [] if (rank == root or side is not None) else [frames.ptrs[root] + offset]
            piece = band[a - lo: b - lo]
            transform._apply_device(image, name, code, row_range=(a, b), out=piece.unsqueeze(0), mirrors=mirrors or None,
                                    frame_rows=frame_rows)
            if side is not None:
                done = torch.cuda.Event()
                done.record(main)
                side.wait_event(done)
                with torch.cuda.stream(side):
                    root_view[a:b].copy_(piece, non_blocking=True)
        if side is not None:
            main.wait_stream(side)
    frames.stream_barrier()                               # every band has landed
    return frames.tensor if full_here else frames.tensor[: hi - lo]


def apply_frames_fused(transform, frames, *, root: int = 0, group=None, pad_mode: str = "symmetric", dtype=None,
                       chunk_frames: int | None = None, transport: str | None = None):
    """Frames sharded by rank with the output gather fused into the overlap-add kernel (config 3, no NCCL gather).

    Every rank corrects its ``frame_shard`` block of the (B, H, W) CUDA tensor ``frames``; the kernel stores
    the block into this rank's buffer and, through the peer mapping, into its place in the root's (B, H, W)
    result while it runs.  With ``chunk_frames`` the block is corrected that many frames at a time: beyond ~4 ranks the
    root's NVLink ingress is the bound of this exchange (world - 1 blocks into one GPU), and with chunks it starts
    to fill after the first chunk's K1 / K2 instead of after the whole block's, so the transfer hides behind the
    remaining arithmetic.  Small chunks cost arithmetic efficiency (the transfer kernel is re-read per chunk), so
    the default (None) keeps the block whole up to 4 ranks and goes frame by frame beyond; 0 = always whole.

    ``transport``: "stores" — the overlap-add kernel itself stores every pixel to the root's buffer (peer stores from
    the SMs); "copy" — the kernel writes locally and a copy engine moves each finished chunk to the root over
    NVLink on a second stream, while the SMs go on with the next chunk (a DMA engine sustains the link better than
    16-byte stores do, and the arithmetic never waits for the link).  Default: "stores" up to 4 ranks, "copy" beyond.
    Returns the full result on ``root`` (a buffer reused by the next call with the same shape) and this rank's own
    block elsewhere.  Bit-identical to the single-GPU result.
    """
    import torch
    import torch.distributed as dist

    from regularizepsf_b200 import _native
    from regularizepsf_b200.transform import _normalize_dtype

    if not (hasattr(frames, "is_cuda") and frames.is_cuda and frames.dim() == 3):
        raise ValueError("apply_frames_fused needs a (B, H, W) CUDA tensor")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    name = _normalize_dtype(dtype)
    want = torch.float32 if name == "float32" else torch.float64
    n, h, w = (int(v) for v in frames.shape)
    begin, end = frame_shard(n, rank, world)
    key = ("frames", id(group), (n, h, w), want, root, torch.cuda.current_device())
    bufs = _peer_frames.get(key)
    if bufs is None:
        bufs = _peer_frames[key] = PeerFrames((n, h, w) if rank == root else (max(end - begin, 1), h, w), want, group)
    bufs.stream_barrier()                                 # the root is done reading the previous result
    itemsize = bufs.tensor.element_size()
    if transport is None:
        transport = "copy" if world > 4 else "stores"
    if transport not in ("stores", "copy"):
        raise ValueError(f"transport must be 'stores' or 'copy', got {transport!r}")
    if chunk_frames is None:
        chunk_frames = 1 if world > 4 else 0
    step = max(1, int(chunk_frames)) if chunk_frames else max(end - begin, 1)
    main = torch.cuda.current_stream()
    side = bufs.copy_stream() if (transport == "copy" and rank != root) else None
    root_view = bufs.peer_tensor(root, (n, h, w)) if side is not None else None
    if side is not None:
        side.wait_stream(main)                            # the barrier above orders the copies too
    for b0 in range(begin, end, step):
        b1 = min(end, b0 + step)
        if rank == root:
            out, mirrors = bufs.tensor[b0:b1], None
        else:
            out = bufs.tensor[b0 - begin: b1 - begin]
            mirrors = [bufs.ptrs[root] + b0 * h * w * itemsize] if side is None else None
        transform._apply_device(frames[b0:b1], name, _native.PAD_MODES[pad_mode], out=out, mirrors=mirrors)
        if side is not None:
            done = torch.cuda.Event()
            done.record(main)
            side.wait_event(done)
            with torch.cuda.stream(side):
                root_view[b0:b1].copy_(out, non_blocking=True)
    if side is not None:
        main.wait_stream(side)
    bufs.stream_barrier()                                 # every block has landed in the root's buffer
    return bufs.tensor if rank == root else bufs.tensor[: end - begin]


def apply_slabs_sharded(transform, image, *, group=None, pad_mode: str = "symmetric", dtype=None, gather: str = "all",
                        root: int = 0):
    """Correct one large frame split into patch-row slabs, bands exchanged over NCCL (the baseline of
    ``apply_slabs_fused``).  ``gather``: "all" — every rank returns the full frame; "root" — ``root`` returns the full
    frame, the others their band; "none" — every rank returns its band."""
    import torch.distributed as dist

    from regularizepsf_b200 import _native
    from regularizepsf_b200.transform import _is_torch_tensor, _normalize_dtype

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    bounds = slab_bounds(image.shape[-2], transform.psf_shape[0], world)
    name = _normalize_dtype(dtype)
    code = _native.PAD_MODES[pad_mode]
    if gather not in ("all", "root", "none"):
        raise ValueError(f"gather must be 'all', 'root' or 'none', got {gather!r}")
    if gather != "all":
        if _is_torch_tensor(image):
            band = transform._apply_device(image, name, code, row_range=bounds[rank])
        else:
            band = transform._apply_host(np.asarray(image), name, code, row_range=bounds[rank])
        if gather == "none":
            return band
        full = gather_slabs(band, bounds, root, group)
        return full if rank == root else band
    if _is_torch_tensor(image):
        sizes = {hi - lo for lo, hi in bounds}
        if image.dim() == 2 and len(sizes) == 1:
            # equal bands: every rank writes its band straight into its place in the full frame and the
            # all-gather runs in place (NCCL's in-place layout: input = output + rank * count) — no
            # staging copies either side of the collective
            import torch
            want = torch.float32 if name == "float32" else torch.float64
            full = torch.empty(tuple(image.shape), dtype=want, device=image.device)
            lo, hi = bounds[rank]
            transform._apply_device(image, name, code, row_range=(lo, hi), out=full[lo:hi].unsqueeze(0))
            dist.all_gather_into_tensor(full, full[lo:hi], group=group)
            return full
        band = transform._apply_device(image, name, code, row_range=bounds[rank])
    else:
        band = transform._apply_host(np.asarray(image), name, code, row_range=bounds[rank])
    return all_gather_slabs(band, bounds, group)
