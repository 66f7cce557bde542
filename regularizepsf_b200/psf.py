"""``ArrayPSF`` with its FFT cube resident in HBM.

Mirrors regularizepsf/psf.py:192-416 for the part of the class the correction path uses
(construction, accessors, equality).  ``fft_evaluations`` keeps the reference layout —
(N, P, P) complex, full unshifted spectrum, complex64 for float32 values and complex128 for
float64 — but is computed by the library's own batched 2-D FFT (``rpsf_psf_fft2``) instead of
``scipy.fft.fft2`` (psf.py:216-219) and stays on the device until a host accessor asks for it.
The cube is computed the first time something needs it (``fft_evaluations``, ``fft_at``,
``ArrayPSFTransform.construct``, ``save``, ``==``), so building models on a host without a GPU works
and only the spectrum itself needs the device.  ``save`` / ``load`` keep the reference's h5 and FITS
layouts (psf.py:262-330).
"""
from __future__ import annotations

import numpy as np

from regularizepsf_b200 import _native
from regularizepsf_b200.device import DeviceCube
from regularizepsf_b200.exceptions import IncorrectShapeError, InvalidCoordinateError
from regularizepsf_b200.util import IndexedCube


def _device_fft_cube(values_cube: IndexedCube) -> DeviceCube:
    torch = _native.require_cuda()
    lib = _native.load()
    values = values_cube.values
    if values.dtype not in (np.float32, np.float64):
        values = values.astype(np.float64)          # scipy.fft promotes the same way
    n, p0, p1 = values.shape
    if p0 != p1:
        raise IncorrectShapeError(f"PSF samples must be square for the device FFT, got {(p0, p1)}")
    if n and not 1 <= p0 <= 512:
        raise NotImplementedError(f"patch size {p0} has no device path (1..512: radix-2 kernels for powers of two "
                                  "16..512, a direct DFT otherwise)")
    dev_values = torch.from_numpy(np.ascontiguousarray(values)).cuda()
    cdtype = torch.complex64 if values.dtype == np.float32 else torch.complex128
    out = torch.empty((n, p0, p1), dtype=cdtype, device=dev_values.device)
    if n:
        code = _native.F32 if values.dtype == np.float32 else _native.F64
        _native.check(lib.rpsf_psf_fft2(dev_values.data_ptr(), out.data_ptr(), n, p0, code,
                                        dev_values.device.index, _native.current_stream_ptr(torch)))
    return DeviceCube(values_cube.coordinates, out)


class ArrayPSF:
    """A PSF model sampled on a grid of patches (regularizepsf/psf.py:192-237)."""

    def __init__(self, values_cube: IndexedCube, fft_cube: IndexedCube | None = None,
                 workers: int | None = None) -> None:
        self._values_cube = values_cube
        self._workers = workers                      # accepted for API parity; the GPU ignores it
        self._given_fft_cube = fft_cube              # None: computed on the device at first use
        if fft_cube is None:
            return                                   # a computed cube matches the values cube by construction
        if fft_cube.sample_shape != self._values_cube.sample_shape:
            raise IncorrectShapeError(
                f"Values cube and FFT cube have different sample shapes: "
                f"{self._values_cube.sample_shape} != {fft_cube.sample_shape}.")
        if len(fft_cube) != len(self._values_cube):
            raise IncorrectShapeError(
                f"Values cube and FFT cube have different sample counts: "
                f"{len(self._values_cube)} != {len(fft_cube)}.")
        if np.any(np.array(self._values_cube.coordinates) != np.array(fft_cube.coordinates)):
            raise InvalidCoordinateError("Values cube and FFT cube have different coordinates")

    @property
    def _fft_cube(self) -> IndexedCube:
        if self._given_fft_cube is None:
            self._given_fft_cube = _device_fft_cube(self._values_cube)
        return self._given_fft_cube

    @property
    def coordinates(self):
        return self._values_cube.coordinates

    @property
    def values(self) -> np.ndarray:
        return self._values_cube.values

    @property
    def fft_evaluations(self) -> np.ndarray:
        return self._fft_cube.values

    @property
    def fft_cube(self) -> IndexedCube:
        """The FFT cube object itself (a ``DeviceCube`` when it was computed on the GPU)."""
        return self._fft_cube

    def __getitem__(self, coord) -> np.ndarray:
        return self._values_cube[coord]

    def fft_at(self, coord) -> np.ndarray:
        return self._fft_cube[coord]

    @property
    def sample_shape(self) -> tuple[int, int]:
        return self._values_cube.sample_shape

    def __len__(self) -> int:
        return len(self._values_cube)

    def __eq__(self, other) -> bool:
        if not isinstance(other, ArrayPSF):
            raise TypeError("Can only compare ArrayPSF to other ArrayPSF.")
        return self._values_cube == other._values_cube and self._fft_cube == other._fft_cube

    __hash__ = None

    # ------------------------------------------------------------------ persistence (psf.py:262-330)
    def save(self, path) -> None:
        """Save the model: ``.h5`` (coordinates, values, fft_evaluations) or ``.fits``."""
        from regularizepsf_b200 import persistence
        persistence.write_cubes(path, self.coordinates, {"values": self.values, "fft_evaluations": self.fft_evaluations})

    @classmethod
    def load(cls, path) -> "ArrayPSF":
        """Load a model written by this class or by the reference package."""
        from regularizepsf_b200 import persistence
        coordinates, cubes = persistence.read_cubes(path, {"values": False, "fft_evaluations": True})
        return cls(IndexedCube(coordinates, cubes["values"]), IndexedCube(coordinates, cubes["fft_evaluations"]))
