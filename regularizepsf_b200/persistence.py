"""File formats of the reference package, so existing model files load straight into HBM.

h5 (psf.py:274-279,305-311; transform.py:237-241,265-269): datasets ``coordinates`` plus one dataset
per cube under its own name, complex stored natively.  FITS (psf.py:280-287,312-326;
transform.py:242-249,270-279): a ``coordinates`` CompImageHDU plus one CompImageHDU per real cube
and a ``<stem>_real`` / ``<stem>_imag`` pair (quantize_level 32) per complex cube.

h5py and astropy are optional: they are imported when a file of that type is actually read or
written.  Cubes are handed over as numpy arrays in the reference layout (N, P, P).
"""
from __future__ import annotations

import pathlib

import numpy as np

#: dataset name in h5 -> stem of the real/imag HDU pair in FITS, for the complex cubes
_FITS_STEM = {"fft_evaluations": "fft", "transfer_kernel": "transfer"}


def _unsupported(path: pathlib.Path) -> NotImplementedError:
    return NotImplementedError(f"Unsupported file type {path.suffix}. Change to .h5 or .fits.")


def _need(module: str):
    import importlib
    try:
        return importlib.import_module(module)
    except ImportError as exc:  # pragma: no cover - depends on the environment
        raise ImportError(f"{module} is required to read or write this file type") from exc


def write_cubes(path, coordinates, cubes: dict[str, np.ndarray], *, exclusive: bool = False, overwrite: bool = False) -> None:
    """Write ``coordinates`` and the named cubes.  ``exclusive`` fails if an h5 file exists (mode "w-")."""
    path = pathlib.Path(path)
    if path.suffix == ".h5":
        h5py = _need("h5py")
        with h5py.File(path, "w-" if exclusive and not overwrite else "w") as f:
            f.create_dataset("coordinates", data=coordinates)
            for name, cube in cubes.items():
                f.create_dataset(name, data=cube)
    elif path.suffix == ".fits":
        fits = _need("astropy.io.fits")
        hdus = [fits.PrimaryHDU(), fits.CompImageHDU(np.array(coordinates), name="coordinates")]
        for name, cube in cubes.items():
            if np.iscomplexobj(cube):
                stem = _FITS_STEM.get(name, name)
                hdus.append(fits.CompImageHDU(np.ascontiguousarray(cube.real), name=f"{stem}_real", quantize_level=32))
                hdus.append(fits.CompImageHDU(np.ascontiguousarray(cube.imag), name=f"{stem}_imag", quantize_level=32))
            else:
                hdus.append(fits.CompImageHDU(cube, name=name))
        fits.HDUList(hdus).writeto(path, overwrite=overwrite)
    else:
        raise _unsupported(path)


def read_cubes(path, names: dict[str, bool]) -> tuple[list[tuple], dict[str, np.ndarray]]:
    """Read ``coordinates`` and the cubes in ``names`` (name -> is complex)."""
    path = pathlib.Path(path)
    cubes: dict[str, np.ndarray] = {}
    if path.suffix == ".h5":
        h5py = _need("h5py")
        with h5py.File(path, "r") as f:
            coordinates = [tuple(c) for c in f["coordinates"][:]]
            for name in names:
                cubes[name] = f[name][:]
    elif path.suffix == ".fits":
        fits = _need("astropy.io.fits")
        with fits.open(path) as hdul:
            coordinates = [tuple(c) for c in hdul[hdul.index_of("coordinates")].data]
            for name, is_complex in names.items():
                if is_complex:
                    stem = _FITS_STEM.get(name, name)
                    real = hdul[hdul.index_of(f"{stem}_real")].data
                    imag = hdul[hdul.index_of(f"{stem}_imag")].data
                    cubes[name] = real + imag * 1j
                else:
                    cubes[name] = np.array(hdul[hdul.index_of(name)].data)
    else:
        raise _unsupported(path)
    return coordinates, cubes
