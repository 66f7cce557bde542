"""Host-side API around the hot path: functional PSF models and the h5 / FITS file layouts.

Behaviour follows the reference's tests/test_psf.py:17-31,93-237 and tests/test_transform.py:11-27,
85-97.  h5py and astropy are not installed in this image, so the file layouts are exercised through
minimal stand-in modules that record exactly which datasets / HDUs the product code writes and reads.
"""
import pickle
import sys
import types

import numpy as np
import pytest

import regularizepsf_b200 as rp
from regularizepsf_b200 import InvalidFunctionError, simple_functional_psf, varied_functional_psf
from regularizepsf_b200.functional import SimpleFunctionalPSF, VariedFunctionalPSF


# ------------------------------------------------------------------ functional PSFs
def test_simple_psf_parameters_and_call():
    plain = simple_functional_psf(lambda row, col: row + col)
    assert isinstance(plain, SimpleFunctionalPSF) and plain.parameters == set() and plain(1, 2) == 3
    two = simple_functional_psf(lambda row, col, sigma=3, mu=4: row + col + sigma + mu)
    assert two.parameters == {"sigma", "mu"} and two(1, 2) == 10 and two(1, 2, sigma=0) == 7


@pytest.mark.parametrize("bad", [lambda: 1, lambda y, x: x + y, lambda x, sigma: x + sigma, lambda row, sigma: row])
def test_simple_psf_signature_rules(bad):
    with pytest.raises(InvalidFunctionError):
        simple_functional_psf(bad)


def test_simple_psf_decorator_takes_no_arguments():
    with pytest.raises(TypeError):
        simple_functional_psf(3)


def test_varied_psf_valid_and_parameter_checks():
    base = simple_functional_psf(lambda row, col, sigma=5: row + col + sigma)
    varied = varied_functional_psf(base)(lambda row, col: {"sigma": 1})
    assert isinstance(varied, VariedFunctionalPSF) and varied.parameters == {"sigma"} and varied(0, 0) == 1
    with pytest.raises(InvalidFunctionError):       # too few arguments
        varied_functional_psf(base)(lambda: {"sigma": 0.1})
    with pytest.raises(InvalidFunctionError):       # too many
        varied_functional_psf(base)(lambda row, col, c: {"sigma": 0.1})
    with pytest.raises(InvalidFunctionError):       # wrong names
        varied_functional_psf(base)(lambda c, col: {"sigma": 0.1})
    with pytest.raises(InvalidFunctionError):
        varied_functional_psf(base)(lambda row, c: {"sigma": 0.1})
    with pytest.raises(InvalidFunctionError):       # parameters do not match the base model
        varied_functional_psf(base)(lambda row, col: {"n": 0, "sigma": 1})


def test_varied_psf_decorator_misuse():
    with pytest.raises(TypeError):
        varied_functional_psf()(lambda row, col: {"sigma": 0.2})
    with pytest.raises(TypeError):
        varied_functional_psf(None)
    with pytest.raises(TypeError):
        @varied_functional_psf
        def naked(row, col):
            return {"sigma": 0.1}


def test_varied_psf_validates_at_call_unless_switched_off():
    base = simple_functional_psf(lambda row, col, m: row + col + m)

    def wobbly(row, col):
        return {"m": 30} if (row == 0 and col == 0) else {"n": 100, "m": 30}

    with pytest.raises(InvalidFunctionError):
        varied_functional_psf(base)(wobbly)(10, 10)
    relaxed = varied_functional_psf(base)(check_at_call=False)(lambda row, col: {"m": row})
    assert relaxed(2, 3) == 2 + 3 + 2


def test_functional_psfs_sample_to_array_psfs_without_a_gpu():
    simple = simple_functional_psf(lambda row, col, a=10: row + col + a)
    psf = simple.as_array_psf([(0, 0), (1, 0)], 3)
    rr, cc = np.meshgrid(np.arange(3), np.arange(3))
    assert len(psf) == 2 and psf.sample_shape == (3, 3) and psf.coordinates == [(0, 0), (1, 0)]
    assert np.allclose(psf[(0, 0)], rr + cc + 10)
    base = simple_functional_psf(lambda row, col, sigma=5: row + col + sigma)
    varied = varied_functional_psf(base)(lambda row, col: {"sigma": row * col})
    psf = varied.as_array_psf([(0, 0), (3, 4)], 3)
    assert np.allclose(psf[(3, 4)], rr + cc + 12) and np.allclose(psf[(0, 0)], rr + cc)
    assert varied.simplify(3, 4)(1, 1) == 14


# ------------------------------------------------------------------ stand-ins for h5py / astropy.io.fits
class _FakeH5File:
    def __init__(self, path, mode):
        import os
        self.path, self.mode, self.data = str(path), mode, {}
        if mode == "w-" and os.path.exists(self.path):
            raise FileExistsError(self.path)
        if mode == "r":
            with open(self.path, "rb") as f:
                self.data = pickle.load(f)

    def create_dataset(self, name, data):
        self.data[name] = np.array(data)

    def __getitem__(self, name):
        return self.data[name]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.mode != "r":
            with open(self.path, "wb") as f:
                pickle.dump(self.data, f)


class _FakeHDU:
    def __init__(self, data=None, name="PRIMARY", quantize_level=None):
        self.data, self.name, self.quantize_level = (None if data is None else np.array(data)), name, quantize_level


class _FakeHDUList(list):
    def writeto(self, path, overwrite=False):
        import os
        if os.path.exists(path) and not overwrite:
            raise OSError(f"File {path} already exists.")
        with open(path, "wb") as f:
            pickle.dump([(h.name, h.data, h.quantize_level) for h in self], f)

    def index_of(self, name):
        return [h.name for h in self].index(name)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _fake_fits_open(path):
    with open(path, "rb") as f:
        return _FakeHDUList(_FakeHDU(d, n, q) for n, d, q in pickle.load(f))


@pytest.fixture()
def fake_io(monkeypatch):
    h5 = types.ModuleType("h5py")
    h5.File = _FakeH5File
    fits = types.ModuleType("astropy.io.fits")
    fits.PrimaryHDU, fits.CompImageHDU, fits.HDUList, fits.open = _FakeHDU, _FakeHDU, _FakeHDUList, _fake_fits_open
    astropy, aio = types.ModuleType("astropy"), types.ModuleType("astropy.io")
    astropy.io, aio.fits = aio, fits
    for name, mod in (("h5py", h5), ("astropy", astropy), ("astropy.io", aio), ("astropy.io.fits", fits)):
        monkeypatch.setitem(sys.modules, name, mod)


def _kernel_cube(coords, size=16):
    rng = np.random.default_rng(2)
    return rp.IndexedCube(coords, (rng.standard_normal((len(coords), size, size))
                                   + 1j * rng.standard_normal((len(coords), size, size))).astype(np.complex64))


@pytest.mark.parametrize("extension", ["h5", "fits"])
def test_transform_saves_and_loads_in_the_reference_layout(tmp_path, fake_io, extension):
    coords = [(0, 0), (8, 8), (0, 8)]
    transform = rp.ArrayPSFTransform(_kernel_cube(coords))
    path = tmp_path / f"transform.{extension}"
    transform.save(path)
    loaded = rp.ArrayPSFTransform.load(path)
    assert isinstance(loaded, rp.ArrayPSFTransform) and loaded == transform and loaded.coordinates == coords
    with pytest.raises((FileExistsError, OSError)):            # h5 mode "w-" / writeto(overwrite=False)
        transform.save(path)
    transform.save(path, overwrite=True)
    # exactly the datasets / HDUs the reference writes (transform.py:237-249)
    with open(path, "rb") as f:
        stored = pickle.load(f)
    if extension == "h5":
        assert sorted(stored) == ["coordinates", "transfer_kernel"] and np.iscomplexobj(stored["transfer_kernel"])
    else:
        assert [n for n, _, _ in stored] == ["PRIMARY", "coordinates", "transfer_real", "transfer_imag"]
        assert [q for _, _, q in stored][2:] == [32, 32]


@pytest.mark.parametrize("extension", ["h5", "fits"])
def test_arraypsf_saves_and_loads_in_the_reference_layout(tmp_path, fake_io, extension):
    coords = [(0, 0), (1, 1), (2, 2)]
    values = np.random.default_rng(1).random((3, 16, 16))
    psf = rp.ArrayPSF(rp.IndexedCube(coords, values), rp.IndexedCube(coords, np.fft.fft2(values)))
    path = tmp_path / f"psf.{extension}"
    psf.save(path)
    loaded = rp.ArrayPSF.load(path)
    assert isinstance(loaded, rp.ArrayPSF) and loaded == psf
    with open(path, "rb") as f:
        stored = pickle.load(f)
    if extension == "h5":
        assert sorted(stored) == ["coordinates", "fft_evaluations", "values"]
    else:
        assert [n for n, _, _ in stored] == ["PRIMARY", "coordinates", "values", "fft_real", "fft_imag"]


def test_unknown_suffix_is_not_implemented(tmp_path):
    transform = rp.ArrayPSFTransform(_kernel_cube([(0, 0)]))
    with pytest.raises(NotImplementedError):
        transform.save(tmp_path / "transform.npz")
    with pytest.raises(NotImplementedError):
        rp.ArrayPSFTransform.load(tmp_path / "transform.npz")
    with pytest.raises(NotImplementedError):
        rp.ArrayPSF.load(tmp_path / "psf.txt")
