"""``ArrayPSFTransform`` — the drop-in for the reference's correction path, running on a B200.

Mirrors regularizepsf/transform.py:25-177 (``__init__``, ``psf_shape``, ``coordinates``,
``__len__``, ``construct``, ``apply``, ``__eq__``) with the same names, argument order, defaults
and error types.  ``construct`` and ``apply`` are thin calls through the C ABI
(include/rpsf_b200.h) into hand-written sm_100a kernels; see DESIGN.md.  No CPU fallback.
"""
from __future__ import annotations

import ctypes
import math
import weakref

import numpy as np

from regularizepsf_b200 import _native
from regularizepsf_b200.device import DeviceCube, cube_tensor, pinned_empty
from regularizepsf_b200.exceptions import IncorrectShapeError, InvalidCoordinateError
from regularizepsf_b200.util import IndexedCube

_DTYPES = {"float32": _native.F32, "float64": _native.F64}
_default_dtype = "float32"


def set_default_dtype(name: str) -> None:
    """Select the arithmetic precision of ``apply``: "float32" (default) or "float64" (validation mode)."""
    global _default_dtype
    if name not in _DTYPES:
        raise ValueError(f"dtype must be one of {sorted(_DTYPES)}, got {name!r}")
    _default_dtype = name


def _normalize_dtype(dtype) -> str:
    if dtype is None:
        return _default_dtype
    name = np.dtype(dtype).name if not isinstance(dtype, str) else dtype
    if name not in _DTYPES:
        raise ValueError(f"dtype must be one of {sorted(_DTYPES)}, got {dtype!r}")
    return name


def _destroy(lib, kind: str, handle: int) -> None:
    try:
        getattr(lib, f"rpsf_{kind}_destroy")(handle)
    except Exception:  # pragma: no cover - interpreter shutdown
        pass


class _NativeTransform:
    """Owns one ``rpsf_transform`` (per compute dtype) and its plans."""

    def __init__(self, coords: np.ndarray, patch: int, dtype_name: str, device: int, kernel_tensor, stream: int,
                 keep: np.ndarray | None = None):
        self.lib = _native.load()
        self.dtype_name = dtype_name
        self.device = device
        handle = ctypes.c_void_p()
        if keep is None:
            _native.check(self.lib.rpsf_transform_create(
                ctypes.byref(handle), coords.ctypes.data, coords.shape[0], patch, _DTYPES[dtype_name], device))
        else:       # a row-slab shard: the whole coordinate list, kernels of the kept patches only
            keep = np.ascontiguousarray(keep, dtype=np.uint8)
            _native.check(self.lib.rpsf_transform_create_subset(
                ctypes.byref(handle), coords.ctypes.data, coords.shape[0], patch, _DTYPES[dtype_name], device,
                keep.ctypes.data))
        self.handle = handle.value
        # geometry key -> plan handle, most recently used last (bounded: see plan())
        self._plans: dict[tuple, int] = {}
        self._saturation: dict[int, tuple] = {}          # plan handle -> the saturation settings it was last given
        self._finalizer = weakref.finalize(self, _NativeTransform._cleanup, self.lib, self.handle, self._plans)
        kcode = _native.F32 if str(kernel_tensor.dtype) == "torch.complex64" else _native.F64
        _native.check(self.lib.rpsf_transform_set_kernel(self.handle, kernel_tensor.data_ptr(), kcode, stream))

    @staticmethod
    def _cleanup(lib, handle, plans):
        for plan in plans.values():
            _destroy(lib, "plan", plan)
        plans.clear()
        _destroy(lib, "transform", handle)

    #: at most this many plans stay alive per transform and compute dtype; each owns max_batch spectrum
    #: workspaces (75.8 MB per 2048^2 / 256-px frame), so an unbounded cache is an HBM leak
    MAX_PLANS = 6

    def plan(self, height: int, width: int, pad_mode: int, row_begin: int, row_end: int, max_batch: int) -> int:
        """Plan for this geometry able to take ``max_batch`` frames per call.

        A cached plan of the same geometry is reused when its capacity is in [max_batch, 2 * max_batch]
        (``rpsf_apply`` accepts any batch up to the plan's; the factor keeps a single-frame call off a plan
        whose work lists were cut for large batches).  Plans are evicted least-recently-used beyond
        ``MAX_PLANS``; eviction calls ``rpsf_plan_destroy`` (its ``cudaFree`` waits for work in flight).
        """
        geometry = (height, width, pad_mode, row_begin, row_end)
        exact = geometry + (max_batch,)
        if exact in self._plans:
            if next(reversed(self._plans)) != exact:
                self._plans[exact] = self._plans.pop(exact)
            return self._plans[exact]
        best = None
        for key in self._plans:
            if key[:5] == geometry and max_batch <= key[5] <= 2 * max_batch and (best is None or key[5] < best[5]):
                best = key
        if best is not None:
            plan = self._plans.pop(best)
            self._plans[best] = plan                      # most recently used last
            return plan
        out = ctypes.c_void_p()
        _native.check(self.lib.rpsf_plan_create(ctypes.byref(out), self.handle, height, width, pad_mode,
                                                row_begin, row_end, max_batch))
        self._plans[geometry + (max_batch,)] = out.value
        while len(self._plans) > self.MAX_PLANS:
            oldest = next(iter(self._plans))
            handle = self._plans.pop(oldest)
            self._saturation.pop(handle, None)
            _destroy(self.lib, "plan", handle)
        return out.value

    def set_saturation(self, plan: int, sat: tuple) -> None:
        """``rpsf_plan_set_saturation`` only when the settings of this plan change."""
        if self._saturation.get(plan) != sat:
            _native.check(self.lib.rpsf_plan_set_saturation(plan, *sat))
            self._saturation[plan] = sat

    def plan_info(self, plan: int) -> dict:
        info = (ctypes.c_int64 * 8)()
        _native.check(self.lib.rpsf_plan_info(plan, info))
        return {"active_patches": info[0], "colours": info[1], "workspace_bytes": info[2],
                "rows_read": (info[3], info[4]), "colour0_tiles_band": bool(info[5]),
                "overlap_add": {2: "streaming chains", 1: "row-pair gather", 0: "colour phases"}[int(info[6]) & 7],
                "fused_pipeline": bool(int(info[6]) & 8), "gather_teams": info[7]}


class ArrayPSFTransform:
    """A source→target PSF transform that can be applied to images (transform.py:25-51)."""

    #: host-buffer path: frames are streamed through the GPU in chunks of about this many input
    #: bytes (at least one frame), so upload, kernels and download of successive chunks overlap
    HOST_CHUNK_BYTES = 8 << 20

    def __init__(self, transfer_kernel: IndexedCube) -> None:
        self._transfer_kernel = transfer_kernel
        self._native: dict[tuple[str, int], _NativeTransform] = {}
        self._coords_i32: np.ndarray | None = None
        # row-slab shard (distributed.shard_transform_rows): coordinates of the COMPLETE transform and which of them
        # this object holds kernels for; None for an ordinary transform
        self._shard_coordinates = None
        self._shard_keep: np.ndarray | None = None

    @classmethod
    def sharded(cls, transfer_kernel: IndexedCube, all_coordinates, keep) -> "ArrayPSFTransform":
        """A transform that holds the kernels of only some patches of a larger one (one rank of a patch-row slab
        split).  ``transfer_kernel`` carries the kept patches in the order they have in
        ``all_coordinates``; ``keep`` is the boolean mask over ``all_coordinates``.  Colour classes, and with them
        the summation order, are those of the complete transform, so slabs stitch bit-identically."""
        keep = np.asarray(keep, dtype=bool)
        full = np.asarray(all_coordinates)
        if full.ndim != 2 or full.shape[1] != 2 or keep.shape != (full.shape[0],):
            raise InvalidCoordinateError("all_coordinates must be (N, 2) and keep a mask of length N")
        kept = [tuple(c) for c in np.asarray(transfer_kernel.coordinates).reshape(-1, 2).tolist()]
        if kept != [tuple(c) for c in full[keep].tolist()]:
            raise InvalidCoordinateError("the shard's coordinates are not all_coordinates[keep], in order")
        obj = cls(transfer_kernel)
        obj._shard_coordinates, obj._shard_keep = full, keep
        return obj

    # ------------------------------------------------------------------ container plumbing
    @property
    def psf_shape(self) -> tuple[int, int]:
        return self._transfer_kernel.sample_shape

    @property
    def coordinates(self):
        return self._transfer_kernel.coordinates

    def __len__(self) -> int:
        return len(self._transfer_kernel)

    def __eq__(self, other) -> bool:
        if not isinstance(other, ArrayPSFTransform):
            raise TypeError("Can only compare ArrayPSFTransform to another ArrayPSFTransform.")
        return self._transfer_kernel == other._transfer_kernel

    __hash__ = None

    # ------------------------------------------------------------------ persistence (transform.py:220-282)
    def save(self, path, overwrite: bool = False) -> None:
        """Save the transfer kernel: ``.h5`` (coordinates, transfer_kernel) or ``.fits`` (transfer_real/_imag)."""
        from regularizepsf_b200 import persistence
        persistence.write_cubes(path, self.coordinates, {"transfer_kernel": self._transfer_kernel.values},
                                exclusive=True, overwrite=overwrite)

    @classmethod
    def load(cls, path) -> "ArrayPSFTransform":
        """Load a transform written by this class or by the reference package; the kernel is uploaded to
        HBM (and re-laid out for the column kernel) at the first ``apply``."""
        from regularizepsf_b200 import persistence
        coordinates, cubes = persistence.read_cubes(path, {"transfer_kernel": True})
        return cls(IndexedCube(coordinates, cubes["transfer_kernel"]))

    # ------------------------------------------------------------------ construct
    @classmethod
    def construct(cls, source, target, alpha: float, epsilon: float) -> "ArrayPSFTransform":
        """Build the regularised transfer kernel on the device (transform.py:53-83).

        ``K = conj(S) |S|^(alpha-1) / (|S|^(alpha+1) + (epsilon |T|)^(alpha+1)) * T`` elementwise over
        the two FFT cubes, in the cubes' precision, with no zero guard (0/0 bins are NaN, as in
        the reference).
        """
        if np.any(np.array(source.coordinates) != np.array(target.coordinates)):
            raise InvalidCoordinateError("Source PSF coordinates do not match target PSF coordinates.")
        torch = _native.require_cuda()
        lib = _native.load()
        s_cube = source.fft_cube if hasattr(source, "fft_cube") else source._fft_cube
        t_cube = target.fft_cube if hasattr(target, "fft_cube") else target._fft_cube
        if s_cube.sample_shape != t_cube.sample_shape:
            raise IncorrectShapeError(f"source and target sample shapes differ: "
                                      f"{s_cube.sample_shape} != {t_cube.sample_shape}")
        s = cube_tensor(s_cube, torch)
        t = cube_tensor(t_cube, torch)
        if s.device != t.device:
            raise ValueError(f"source and target FFT cubes live on different devices: {s.device} and {t.device}")
        wide = torch.complex128 if torch.complex128 in (s.dtype, t.dtype) else torch.complex64
        with torch.cuda.device(s.device):                 # the launch stream must belong to the cubes' device
            s = s.to(wide) if s.dtype != wide else s
            t = t.to(wide) if t.dtype != wide else t
            kernel = torch.empty_like(s)
            code = _native.F32 if wide == torch.complex64 else _native.F64
            _native.check(lib.rpsf_construct_kernel(s.data_ptr(), t.data_ptr(), kernel.data_ptr(), s.numel(), code,
                                                    float(alpha), float(epsilon), s.device.index,
                                                    _native.current_stream_ptr(torch)))
        return cls(DeviceCube(source.coordinates, kernel))

    # ------------------------------------------------------------------ native state
    def _coords(self) -> np.ndarray:
        if self._coords_i32 is None:
            raw = np.asarray(self.coordinates if self._shard_coordinates is None else self._shard_coordinates)
            if raw.size == 0:
                raw = raw.reshape(0, 2)
            if raw.ndim != 2 or raw.shape[1] != 2:
                raise InvalidCoordinateError("coordinates must be (row, col) pairs")
            as_int = np.rint(raw).astype(np.int64)
            if not np.array_equal(as_int, raw):
                raise InvalidCoordinateError("patch corners must be integers to slice an image")
            if as_int.size and (as_int.min() < -2**30 or as_int.max() > 2**30):
                raise InvalidCoordinateError("patch corners out of range")
            self._coords_i32 = np.ascontiguousarray(as_int, dtype=np.int32)
        return self._coords_i32

    def _native_transform(self, dtype_name: str, device: int | None = None) -> _NativeTransform:
        """The native transform for (compute dtype, device); ``device`` defaults to the current CUDA device.
        Call with that device current (``_apply_device`` / ``_apply_host`` do)."""
        torch = _native.require_cuda()
        if device is None:
            device = torch.cuda.current_device()
        key = (dtype_name, device)
        nt = self._native.get(key)
        if nt is None:
            p0, p1 = self.psf_shape
            if p0 != p1:
                # the reference's apodization window only broadcasts for square patches (transform.py:151-155)
                raise IncorrectShapeError(f"patches must be square, got {(p0, p1)}")
            kernel = self._transfer_kernel
            if isinstance(kernel, DeviceCube):
                kt = kernel.tensor
                if kt.device.index != device:
                    kt = kt.to(f"cuda:{device}")
            else:
                values = np.ascontiguousarray(kernel.values)
                if values.dtype not in (np.complex64, np.complex128):
                    values = values.astype(np.complex128)
                kt = torch.from_numpy(values).to(f"cuda:{device}")
            kt = kt.contiguous()
            with torch.cuda.device(device):
                # set_kernel is stream-ordered on this device's current stream, the stream `kt` was made on, so
                # the caching allocator cannot hand kt's block to anyone before the layout kernel has read it
                nt = _NativeTransform(self._coords(), p0, dtype_name, device, kt, _native.current_stream_ptr(torch),
                                      keep=self._shard_keep)
            self._native[key] = nt
        return nt

    # ------------------------------------------------------------------ apply
    def apply(self, image, workers: int | None = None, pad_mode: str = "symmetric",
              saturation_threshold: float = math.inf, saturation_dilation: int = 1,
              neighborhood_width: int = 7, *, dtype=None, out_dtype=None):
        """Apply the transform to an image (transform.py:85-177).

        ``image`` may be a numpy array (any real dtype; returns a fresh float64 numpy array exactly
        like the reference) or a CUDA ``torch.Tensor`` (stays on the device; returns a tensor in
        the compute dtype).  A leading batch axis ``(B, H, W)`` is accepted.  ``workers`` is
        accepted for signature parity and ignored.  ``dtype`` picks the arithmetic precision:
        "float32" (default) or "float64" (validation mode).  ``out_dtype`` (numpy input only) is the
        dtype of the returned array: float64 like the reference by default; ``np.float32`` halves the
        device-to-host copy, which is what bounds a host-to-host call.
        """
        del workers
        if (pad_mode == "symmetric" and saturation_threshold == math.inf and dtype is None and out_dtype is None
                and _is_torch_tensor(image)):
            # the common device-resident call: nothing to normalise (the host side of such a call costs as much as
            # the kernels of a 512^2 frame, scripts/host_cost.py)
            return self._apply_device(image, "float32", 0)
        dtype_name = _normalize_dtype(dtype)
        # (threshold, dilation, neighbourhood width) of the saturation branch, transform.py:125-138;
        # +inf switches it off.  The mask, its dilation, the raster-ordered fill and the final
        # restore all run on the device, so no host pass over the frame decides anything.
        sat = (float(saturation_threshold), int(saturation_dilation), int(neighborhood_width))
        if math.isnan(sat[0]):
            sat = (math.inf,) + sat[1:]                    # `padded > nan` is all False in the reference
        out_np = np.dtype(np.float64 if out_dtype is None else out_dtype)
        if out_np not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise NotImplementedError(f"out_dtype must be float32 or float64, got {out_np}")
        if pad_mode not in _native.PAD_MODES:
            return self._apply_materialized_pad(image, dtype_name, pad_mode, out_np, sat)
        if _is_torch_tensor(image):
            return self._apply_device(image, dtype_name, _native.PAD_MODES[pad_mode], sat=sat)
        image = np.asarray(image)
        if image.ndim not in (2, 3):
            raise IncorrectShapeError(f"image must be (H, W) or (B, H, W), got shape {image.shape}")
        return self._apply_host(image, dtype_name, _native.PAD_MODES[pad_mode], out_dtype=out_np.type, sat=sat)

    _NO_SAT = (math.inf, 1, 7)
    _PAD_MATERIALIZED = 5                                  # RPSF_PAD_MATERIALIZED

    def _apply_materialized_pad(self, image, dtype_name: str, pad_mode, out_np, sat: tuple):
        """``pad_mode`` values without an on-device index map (np.pad's statistical modes, "linear_ramp", ...: the
        reference hands any mode string to np.pad, transform.py:119-123).  The margin is what np.pad makes of it, so it
        is computed by np.pad itself, on the host, in float64 like the reference (2P per side: "linear_ramp" depends on
        the width); the padded frames are uploaded and the kernels read the margin in place.  Slower than the five
        index-map modes (4.5 x the upload at 2048^2 / 256 px) — those stay the fast path."""
        torch = _native.require_cuda()
        is_tensor = _is_torch_tensor(image)
        host = image.detach().cpu().numpy() if is_tensor else np.asarray(image)
        if host.ndim not in (2, 3):
            raise IncorrectShapeError(f"image must be (H, W) or (B, H, W), got shape {host.shape}")
        squeeze = host.ndim == 2
        frames = host[np.newaxis] if squeeze else host
        pad = 2 * self.psf_shape[0]
        padded = np.stack([np.pad(f.astype(float), ((pad, pad), (pad, pad)), mode=pad_mode) for f in frames])
        want = torch.float32 if dtype_name == "float32" else torch.float64
        device = image.device if is_tensor else torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(device):
            dev = torch.from_numpy(padded).to(device).to(want)
            b, h, w = frames.shape
            inner = dev[:, pad:pad + h, pad:pad + w]            # a view: the kernels address the margin through its pitch
            out = self._apply_device(inner, dtype_name, self._PAD_MATERIALIZED, sat=sat)
        if is_tensor:
            return out[0] if squeeze else out
        result = out.cpu().numpy().astype(out_np.type, copy=False)
        return result[0] if squeeze else result

    def _apply_host(self, image: np.ndarray, dtype_name: str, pad_code: int, out_dtype=np.float64,
                    row_range: tuple[int, int] | None = None, sat: tuple = _NO_SAT) -> np.ndarray:
        nt = self._native_transform(dtype_name)
        squeeze = image.ndim == 2
        frames = image[np.newaxis] if squeeze else image
        code = _native.dtype_code(frames.dtype)
        if code is None:
            frames = frames.astype(np.float64)
            code = _native.F64
        frames = np.ascontiguousarray(frames)
        b, h, w = frames.shape
        r0, r1 = row_range if row_range is not None else (0, h)
        if r1 <= r0:                                       # an empty band (more ranks than half-patch rows)
            empty = np.empty((b, 0, w), out_dtype)
            return empty[0] if squeeze else empty
        chunk = max(1, min(b, self.HOST_CHUNK_BYTES // max(1, h * w * frames.dtype.itemsize)))
        plan = nt.plan(h, w, pad_code, r0, r1, chunk)
        nt.set_saturation(plan, sat)
        out = pinned_empty((b, r1 - r0, w), out_dtype)
        ocode = _native.dtype_code(out.dtype)
        _native.check(nt.lib.rpsf_apply_host(plan, frames.ctypes.data, code, out.ctypes.data, ocode, b))
        return out[0] if squeeze else out

    def _apply_device(self, image, dtype_name: str, pad_code: int, row_range: tuple[int, int] | None = None,
                      out=None, sat: tuple = _NO_SAT, mirrors: list[int] | None = None,
                      frame_rows: tuple[int, int] | None = None):
        """``frame_rows = (first, height)``: ``image`` holds only rows [first, first + image.shape[-2]) of frames that
        are ``height`` rows tall (a row slab keeps just the rows its patches read: ``plan_info()["rows_read"]``)."""
        torch = _native.require_cuda()
        if not image.is_cuda:
            raise ValueError("apply() takes a numpy array or a CUDA tensor; move the tensor to the GPU first")
        if image.device.index == torch._C._cuda_getDevice():
            return self._apply_device_on(torch, image, dtype_name, pad_code, row_range, out, sat, mirrors, frame_rows)
        with torch.cuda.device(image.device):             # plans, streams and launches follow the image's device
            return self._apply_device_on(torch, image, dtype_name, pad_code, row_range, out, sat, mirrors, frame_rows)

    def _apply_device_on(self, torch, image, dtype_name, pad_code, row_range, out, sat, mirrors, frame_rows=None):
        nt = self._native_transform(dtype_name, image.device.index)
        want = torch.float32 if dtype_name == "float32" else torch.float64
        squeeze = image.dim() == 2
        frames = image.unsqueeze(0) if squeeze else image
        if frames.dim() != 3:
            raise IncorrectShapeError(f"image must be (H, W) or (B, H, W), got shape {tuple(image.shape)}")
        if frames.dtype != want:
            frames = frames.to(want)
        if frames.stride(-1) != 1:
            frames = frames.contiguous()
        b, held, w = frames.shape
        first, h = frame_rows if frame_rows is not None else (0, held)
        r0, r1 = row_range if row_range is not None else (0, h)
        if out is None:
            out = torch.empty((b, r1 - r0, w), dtype=want, device=frames.device)
        else:
            if out.device != frames.device:
                raise ValueError(f"`out` lives on {out.device}, the frames on {frames.device}")
            if squeeze and out.dim() == 2:
                out = out.unsqueeze(0)
            if tuple(out.shape) != (b, r1 - r0, w) or out.dtype != want or out.stride(-1) != 1:
                raise IncorrectShapeError(f"`out` must be a {want} tensor of shape {(b, r1 - r0, w)} with unit column "
                                          f"stride, got {out.dtype} {tuple(out.shape)}")
        if r1 <= r0:                                       # an empty band: nothing to compute, no native call
            return out[0] if squeeze else out
        plan = nt.plan(h, w, pad_code, r0, r1, b)
        nt.set_saturation(plan, sat)
        if mirrors:        # device pointers of peer buffers laid out like `out` (distributed.PeerFrames)
            import ctypes
            arr = (ctypes.c_void_p * len(mirrors))(*mirrors)
            _native.check(nt.lib.rpsf_plan_set_output_mirrors(plan, len(mirrors), arr))
        try:
            rc = nt.lib.rpsf_apply(
                plan, frames.data_ptr(), frames.stride(1), frames.stride(0) if b > 1 else held * frames.stride(1), first, held,
                out.data_ptr(), out.stride(1), out.stride(0) if b > 1 else (r1 - r0) * out.stride(1), r0, b,
                _native.current_stream_ptr(torch, frames.device.index))
            if rc:
                _native.check(rc)
        finally:
            if mirrors:
                _native.check(nt.lib.rpsf_plan_set_output_mirrors(plan, 0, None))
        return out[0] if squeeze else out


def _is_torch_tensor(obj) -> bool:
    mod = type(obj).__module__
    return mod == "torch" or mod.startswith("torch.")
