"""Pin the oracle to the UNMODIFIED reference sources (/root/reference, or its pip-installed copy baseline/_ref)."""
import numpy as np
import pytest

from oracle import cpu_oracle as oracle
from oracle import ref_loader
from tests.helpers import make_gaussian

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present on this box")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


@pytest.mark.parametrize("shape,size", [((64, 64), 16), ((100, 90), 11), ((5, 5), 1), ((5, 5), 2), ((15, 15), 4),
                                        ((2048, 2048), 256), ((150, 150), np.ceil(150 * 0.3)), ((8192, 8192), 512)])
def test_covering_identical(ref, shape, size):
    a, b = oracle.covering(shape, size), ref.util.calculate_covering(shape, size)
    assert a.dtype == b.dtype and np.array_equal(a, b)


def test_gaussian_helper_identical():
    import importlib.util, os
    if not os.path.isfile(os.path.join(ref_loader.REF_ROOT, "tests", "helper.py")):
        pytest.skip("the installed copy of the reference (baseline/_ref) carries the package only, not its tests")
    spec = importlib.util.spec_from_file_location("ref_helper", os.path.join(ref_loader.REF_ROOT, "tests", "helper.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for size, fwhm in [(128, 3), (64, 4.5), (33, 2)]:
        assert np.array_equal(mod.make_gaussian(size, fwhm=fwhm), oracle.gaussian_psf(size, fwhm=fwhm))
        assert np.array_equal(mod.make_gaussian(size, fwhm=fwhm), make_gaussian(size, fwhm=fwhm))


@pytest.mark.parametrize("psf_dtype", [np.float32, np.float64])
@pytest.mark.parametrize("alpha,epsilon", [(1.0, 0.1), (3.0, 0.1), (0.5, 0.01), (2.0, 0.3)])
def test_construct_identical(ref, psf_dtype, alpha, epsilon):
    shape, size = (96, 80), 32
    coords = [tuple(int(v) for v in c) for c in oracle.covering(shape, size)]
    src = oracle.coma_psf_cube(coords, size, shape, dtype=psf_dtype)
    tgt = oracle.gaussian_psf_cube(len(coords), size, 3.0, dtype=psf_dtype)
    S = ref.psf.ArrayPSF(ref.util.IndexedCube(coords, src))
    T = ref.psf.ArrayPSF(ref.util.IndexedCube(coords, tgt))
    assert np.array_equal(S.fft_evaluations, oracle.psf_fft(src))
    t = ref.transform.ArrayPSFTransform.construct(S, T, alpha, epsilon)
    mine = oracle.transfer_kernel(oracle.psf_fft(src), oracle.psf_fft(tgt), alpha, epsilon)
    assert mine.dtype == t._transfer_kernel.values.dtype
    assert np.array_equal(mine, t._transfer_kernel.values, equal_nan=True)


def test_construct_nan_pattern_identical(ref):
    # a patch where source and target are both all-zero gives 0/0 bins (no guard in transform.py:78-82)
    coords = [(0, 0), (0, 16), (16, 0)]
    src = np.stack([np.zeros((32, 32)), oracle.gaussian_psf(32, 6.0), oracle.gaussian_psf(32, 3.0)]).astype(np.float32)
    tgt = np.stack([np.zeros((32, 32)), oracle.gaussian_psf(32, 3.0), oracle.gaussian_psf(32, 3.0)]).astype(np.float32)
    S = ref.psf.ArrayPSF(ref.util.IndexedCube(coords, src))
    T = ref.psf.ArrayPSF(ref.util.IndexedCube(coords, tgt))
    with np.errstate(all="ignore"):
        want = ref.transform.ArrayPSFTransform.construct(S, T, 1.0, 0.1)._transfer_kernel.values
        mine = oracle.transfer_kernel(oracle.psf_fft(src), oracle.psf_fft(tgt), 1.0, 0.1)
    assert np.isnan(want).any()
    assert np.array_equal(mine, want, equal_nan=True)


@pytest.mark.parametrize("kwargs", [{}, {"pad_mode": "reflect"}, {"pad_mode": "constant"}, {"pad_mode": "wrap"},
                                    {"pad_mode": "edge"}, {"saturation_threshold": 2000.0},
                                    {"saturation_threshold": 1500.0, "saturation_dilation": 0},
                                    {"saturation_threshold": 2500.0, "saturation_dilation": 3, "neighborhood_width": 5},
                                    {"workers": 2}])
def test_apply_identical(ref, kwargs):
    shape, size = (96, 80), 32
    coords = [tuple(int(v) for v in c) for c in oracle.covering(shape, size)]
    src = oracle.coma_psf_cube(coords, size, shape)
    tgt = oracle.gaussian_psf_cube(len(coords), size, 3.0)
    S = ref.psf.ArrayPSF(ref.util.IndexedCube(coords, src))
    T = ref.psf.ArrayPSF(ref.util.IndexedCube(coords, tgt))
    t = ref.transform.ArrayPSFTransform.construct(S, T, 1.0, 0.1)
    image = oracle.starfield(shape, seed=3)
    want = t.apply(image, **kwargs)
    got = oracle.apply_transform(image, coords, t._transfer_kernel.values, **kwargs)
    assert got.dtype == want.dtype == np.float64
    assert np.array_equal(got, want, equal_nan=True)


def test_apply_identical_config1(ref):
    """BASELINE config 1: 1024^2, 128-px patches, Gaussian fwhm 4 -> 3, alpha 3, eps 0.1."""
    shape, size = (1024, 1024), 128
    coords = [tuple(int(v) for v in c) for c in oracle.covering(shape, size)]
    src = oracle.gaussian_psf_cube(len(coords), size, 4.0)
    tgt = oracle.gaussian_psf_cube(len(coords), size, 3.0)
    S = ref.psf.ArrayPSF(ref.util.IndexedCube(coords, src))
    T = ref.psf.ArrayPSF(ref.util.IndexedCube(coords, tgt))
    t = ref.transform.ArrayPSFTransform.construct(S, T, 3.0, 0.1)
    image = oracle.starfield(shape, seed=1234)
    want = t.apply(image, workers=4)
    got = oracle.apply_transform(image, coords, t._transfer_kernel.values, workers=4)
    assert np.array_equal(got, want)


def test_reference_identity_test_holds_for_oracle(ref):
    """The reference's own test_transform_apply (tests/test_transform.py:29-49) on the oracle, at P=64."""
    size, shape = 64, (512, 512)
    gauss = make_gaussian(size, fwhm=3)
    coords = [tuple(t) for t in oracle.covering(shape, size)]
    values = np.stack([np.zeros((size, size), dtype=np.float32) for _ in coords])
    values[:] = gauss / np.sum(gauss)
    fft = oracle.psf_fft(values)
    kernel = oracle.transfer_kernel(fft, fft, 3.0, 0.1)
    image = np.zeros(shape, dtype=np.float32)
    image[100:300, 50:150] = 5
    out = oracle.apply_transform(image, coords, kernel)
    assert np.allclose(image, out, atol=1e-3)
