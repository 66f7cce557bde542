"""Host-side mirror of the reference API: geometry, containers, error behaviour (no GPU needed).

Modelled on the reference's tests/test_util.py and the error-path tests of tests/test_psf.py /
tests/test_transform.py.
"""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

import regularizepsf_b200 as rp
from regularizepsf_b200.exceptions import IncorrectShapeError, InvalidCoordinateError, NativeLibraryError
from tests.helpers import make_gaussian


def assert_four_covering(corners, img_shape, patch_size):
    counts = np.zeros(img_shape)
    for x, y in corners:
        counts[max(0, int(x)):int(min(img_shape[0], x + patch_size)),
               max(0, int(y)):int(min(img_shape[1], y + patch_size))] += 1
    assert np.all(counts == 4)


@pytest.mark.parametrize("img_shape, patch_size", [((5, 5), 1), ((5, 5), 2), ((15, 15), 3), ((15, 15), 4),
                                                   ((100, 100), 11), ((2048, 2048), 256), ((1024, 1000), 128)])
def test_covering_is_fourfold(img_shape, patch_size):
    assert_four_covering(rp.calculate_covering(img_shape, patch_size), img_shape, patch_size)


@given(img_dim=st.integers(min_value=100, max_value=200), patch_fraction=st.fractions(min_value=0.1, max_value=0.8))
@settings(max_examples=150, deadline=None)
def test_covering_random_square_images(img_dim, patch_fraction):
    patch_size = np.ceil(img_dim * patch_fraction)
    assert_four_covering(rp.calculate_covering((img_dim, img_dim), patch_size), (img_dim, img_dim), patch_size)


def test_covering_groups_are_disjoint_colour_classes():
    size = 16
    cov = rp.calculate_covering((64, 48), size)
    seen = 0
    for rows, cols in ((4, 3), (5, 4), (5, 3), (4, 4)):
        group = cov[seen:seen + rows * cols]
        seen += rows * cols
        for i in range(len(group)):
            for j in range(i + 1, len(group)):
                assert abs(group[i][0] - group[j][0]) >= size or abs(group[i][1] - group[j][1]) >= size
    assert seen == len(cov)


@pytest.mark.parametrize("num_layers, x_shape, y_shape", [(10, 10, 10), (15, 20, 25), (1, 15, 10), (1, 1, 1),
                                                          (0, 1, 1), (0, 0, 0)])
def test_indexed_cube_behaviour(num_layers, x_shape, y_shape):
    data = np.zeros((num_layers, x_shape, y_shape))
    for i in range(num_layers):
        data[i] = i
    coordinates = [(i, i + 1) for i in range(num_layers)]
    cube = rp.IndexedCube(coordinates, data)
    assert cube.sample_shape == (x_shape, y_shape)
    assert len(cube) == num_layers
    assert cube.coordinates == coordinates
    for i, c in enumerate(coordinates):
        assert np.all(cube[c] == i)
        cube[c] = np.full((x_shape, y_shape), -1.0)
        assert np.all(cube[c] == -1)
    assert np.all(cube.values == (-1 if num_layers else 0)) or num_layers == 0


def test_indexed_cube_errors():
    cube = rp.IndexedCube([(0, 0), (1, 1)], np.zeros((2, 4, 4)))
    with pytest.raises(InvalidCoordinateError):
        _ = cube[(5, 5)]
    with pytest.raises(InvalidCoordinateError):
        cube[(5, 5)] = np.zeros((4, 4))
    with pytest.raises(IncorrectShapeError):
        cube[(0, 0)] = np.zeros((3, 4))
    with pytest.raises(IncorrectShapeError):
        rp.IndexedCube([(0, 0)], np.zeros((4, 4)))
    with pytest.raises(IncorrectShapeError):
        rp.IndexedCube([(0, 0)], np.zeros((2, 4, 4)))
    with pytest.raises(TypeError):
        _ = cube == np.zeros((2, 4, 4))
    assert cube == rp.IndexedCube([(0, 0), (1, 1)], np.zeros((2, 4, 4)) + 5e-7)
    assert not (cube == rp.IndexedCube([(0, 0), (1, 1)], np.ones((2, 4, 4))))
    assert not (cube == rp.IndexedCube([(0, 0), (1, 2)], np.zeros((2, 4, 4))))


def _psf_with_host_fft(coords, size=32, dtype=np.float64):
    values = np.stack([make_gaussian(size) for _ in coords]).astype(dtype)
    fft = np.fft.fft2(values)
    return rp.ArrayPSF(rp.IndexedCube(coords, values), rp.IndexedCube(coords, fft))


def test_arraypsf_accessors_with_given_fft_cube():
    coords = [(0, 0), (1, 1), (2, 2)]
    psf = _psf_with_host_fft(coords)
    assert len(psf) == 3 and psf.sample_shape == (32, 32) and psf.coordinates == coords
    assert np.all(psf[(1, 1)] == make_gaussian(32))
    assert np.all(psf.fft_at((0, 0)) == np.fft.fft2(make_gaussian(32)))
    assert psf.fft_evaluations.shape == (3, 32, 32)
    assert psf == _psf_with_host_fft(coords)
    with pytest.raises(TypeError):
        _ = psf == 3


def test_arraypsf_consistency_errors():
    coords = [(0, 0), (1, 1), (2, 2)]
    values = np.stack([make_gaussian(32) for _ in coords])
    with pytest.raises(IncorrectShapeError):      # sample shape mismatch (psf.py:221-226)
        rp.ArrayPSF(rp.IndexedCube(coords, values), rp.IndexedCube(coords, np.zeros((3, 16, 16), complex)))
    with pytest.raises(IncorrectShapeError):      # sample count mismatch (psf.py:228-233)
        rp.ArrayPSF(rp.IndexedCube(coords, values), rp.IndexedCube(coords[:2], np.zeros((2, 32, 32), complex)))
    with pytest.raises(InvalidCoordinateError):   # coordinate mismatch (psf.py:235-237)
        rp.ArrayPSF(rp.IndexedCube(coords, values),
                    rp.IndexedCube([(0, 0), (1, 1), (2, 3)], np.zeros((3, 32, 32), complex)))


def test_construct_rejects_mismatched_coordinates_before_touching_the_gpu():
    # tests/test_transform.py:76-82 of the reference: float coordinate in the target
    source = _psf_with_host_fft([(0, 0), (1, 1), (2, 2)])
    target = _psf_with_host_fft([(0, 0), (1, 1), (0.5, 0.5)])
    with pytest.raises(InvalidCoordinateError):
        rp.ArrayPSFTransform.construct(source, target, 3.0, 0.1)


def test_transform_container_plumbing():
    coords = [(0, 0), (16, 16)]
    cube = rp.IndexedCube(coords, np.ones((2, 32, 32), dtype=np.complex64))
    t = rp.ArrayPSFTransform(cube)
    assert t.psf_shape == (32, 32) and len(t) == 2 and t.coordinates == coords
    assert t == rp.ArrayPSFTransform(rp.IndexedCube(coords, np.ones((2, 32, 32), dtype=np.complex64)))
    with pytest.raises(TypeError):
        _ = t == np.zeros((50, 50))


def test_unknown_dtypes_are_rejected_before_any_device_work():
    t = rp.ArrayPSFTransform(rp.IndexedCube([(0, 0)], np.ones((1, 32, 32), dtype=np.complex64)))
    with pytest.raises(ValueError):
        t.apply(np.zeros((32, 32)), dtype="float16")
    with pytest.raises(NotImplementedError):
        t.apply(np.zeros((32, 32)), out_dtype=np.int16)


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    t = rp.ArrayPSFTransform(rp.IndexedCube([(0, 0)], np.ones((1, 32, 32), dtype=np.complex64)))
    with pytest.raises(NativeLibraryError):
        t.apply(np.zeros((32, 32)))
    psf = rp.ArrayPSF(rp.IndexedCube([(0, 0)], np.ones((1, 32, 32))))     # building the model needs no GPU ...
    with pytest.raises(NativeLibraryError):
        _ = psf.fft_evaluations                                            # ... its spectrum does


def test_product_never_imports_the_oracle_or_a_cpu_fft():
    import os, re
    pkg = os.path.dirname(rp.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            text = open(os.path.join(root, f), errors="ignore").read() if f.endswith((".py", ".cu", ".cuh", ".h")) else ""
            if f.endswith(".py"):
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert not re.search(r"\b(scipy|numpy|np)\.fft\.\w+\(|^\s*(from|import)\s+(scipy|numpy)\.fft\b|"
                                     r"^\s*from\s+(scipy|numpy)\s+import\s+.*\bfft\b", text, re.M), f
                if f != "builder.py":       # the builder's host stages call scipy.ndimage / linalg / interpolate like the reference
                    assert not re.search(r"^\s*(from|import)\s+scipy\b", text, re.M), f
            else:
                assert "cufft" not in text.lower(), f


class _FakePlanLib:
    """Stands in for the C ABI so the plan-cache policy can be tested without a GPU."""

    def __init__(self):
        self.created, self.destroyed, self._next = [], [], 1000

    def rpsf_plan_create(self, out_ref, handle, h, w, pad, r0, r1, max_batch):
        self._next += 1
        out_ref._obj.value = self._next
        self.created.append((self._next, (h, w, pad, r0, r1, max_batch)))
        return 0

    def rpsf_plan_destroy(self, plan):
        self.destroyed.append(plan)
        return 0


def _cache_under_test():
    from regularizepsf_b200.transform import _NativeTransform
    nt = object.__new__(_NativeTransform)
    nt.lib, nt.handle, nt._plans, nt._saturation = _FakePlanLib(), 1, {}, {}
    return nt


def test_plan_cache_reuses_a_plan_with_enough_capacity():
    nt = _cache_under_test()
    p8 = nt.plan(64, 64, 0, 0, 64, 8)
    assert nt.plan(64, 64, 0, 0, 64, 8) == p8            # exact hit
    assert nt.plan(64, 64, 0, 0, 64, 5) == p8            # 5 <= 8 <= 10: the larger plan serves it
    assert nt.plan(64, 64, 0, 0, 64, 4) == p8
    p1 = nt.plan(64, 64, 0, 0, 64, 1)                    # 8 > 2 * 1: a single-frame call gets its own plan
    assert p1 != p8 and len(nt.lib.created) == 2
    assert nt.plan(64, 64, 0, 0, 32, 8) not in (p1, p8)  # another row band is another geometry
    assert len(nt.lib.created) == 3 and not nt.lib.destroyed


def test_plan_cache_is_bounded_and_evicts_least_recently_used():
    nt = _cache_under_test()
    first = nt.plan(32, 32, 0, 0, 32, 1)
    others = [nt.plan(32 + 16 * i, 32, 0, 0, 32, 1) for i in range(1, nt.MAX_PLANS)]
    assert nt.plan(32, 32, 0, 0, 32, 1) == first         # touch: `first` is now the most recently used
    nt.plan(512, 32, 0, 0, 32, 1)                        # one past the bound
    assert nt.lib.destroyed == [others[0]] and len(nt._plans) == nt.MAX_PLANS
    for b in range(1, 40):                               # a caller that varies its batch size does not leak plans
        nt.plan(32, 32, 0, 0, 32, b)
    assert len(nt._plans) == nt.MAX_PLANS
    assert len(nt.lib.created) - len(nt.lib.destroyed) == nt.MAX_PLANS


def test_row_slab_shards_of_a_transform_host_logic():
    """shard_transform_rows / rows_needed / ArrayPSFTransform.sharded: pure host bookkeeping."""
    from regularizepsf_b200 import distributed as rdist
    from regularizepsf_b200.transform import ArrayPSFTransform
    shape, size, world = (640, 384), 128, 4
    coords = [tuple(int(v) for v in c) for c in rp.calculate_covering(shape, size)]
    kernel = np.arange(len(coords), dtype=np.float64)[:, None, None] * np.ones((1, size, size), dtype=np.complex64)
    full = ArrayPSFTransform(rp.IndexedCube(coords, kernel))
    seen = np.zeros(len(coords), dtype=int)
    for rank in range(world):
        lo, hi = rdist.slab_bounds(shape[0], size, world)[rank]
        shard = rdist.shard_transform_rows(full, shape[0], rank, world)
        keep = shard._shard_keep
        assert shard._shard_coordinates.shape == (len(coords), 2) and keep.sum() == len(shard) < len(coords)
        for i, c in enumerate(coords):                                  # exactly the patches that touch the band
            assert keep[i] == (c[0] < hi and c[0] + size > lo)
        kept_ids = shard._transfer_kernel.values[:, 0, 0].real.astype(int)
        assert list(kept_ids) == list(np.flatnonzero(keep))             # kernels follow their coordinates, in order
        assert np.array_equal(shard._coords(), full._coords())          # the native layer still sees every coordinate
        seen += keep
        first, last = rdist.rows_needed(coords, size, shape[0], (lo, hi))
        rows = sorted({r for c in coords if c[0] < hi and c[0] + size > lo for r in range(c[0], c[0] + size)})
        mirrored = {r if 0 <= r < shape[0] else (-r - 1 if r < 0 else 2 * shape[0] - 1 - r) for r in rows}
        assert (first, last) == (min(mirrored), max(mirrored) + 1)
    assert seen.min() >= 1 and seen.max() <= 3                          # halo patch rows belong to two or three ranks
    with pytest.raises(InvalidCoordinateError):                         # kept coordinates must be all_coordinates[keep]
        ArrayPSFTransform.sharded(rp.IndexedCube(coords[:3], kernel[:3]), np.array(coords), np.arange(len(coords)) >= 3)
    assert rdist.rows_needed(coords, size, shape[0], (0, 0)) == (0, 0)
