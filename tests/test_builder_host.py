"""Host side of ArrayPSFBuilder: cutouts and cell assignment against the reference, and
the oracle's cutouts, averaging and core isolation against the reference's.  No GPU: the three device stages (star
cutouts, averaging, core isolation) are replaced by the oracle here and checked on their own in tests/test_gpu_builder.py."""
import os
import warnings

import numpy as np
import pytest

import regularizepsf_b200 as rp
from oracle import cpu_oracle as oracle
from oracle import fake_sep, ref_loader
from regularizepsf_b200 import builder as b
from regularizepsf_b200.exceptions import IncorrectShapeError, PSFBuilderError

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "builder", "c5_builder_p32.npz")
needs_reference = pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present on this box")


def _oracle_isolate(stack):
    return np.stack([oracle.isolate_core(np.array(p, dtype=np.float64)) for p in stack])


def _host_stages(monkeypatch):
    monkeypatch.setattr(b, "cutouts_at", oracle.cutouts_at)
    monkeypatch.setattr(b, "average_cutouts", oracle.average_cutouts)
    monkeypatch.setattr(b, "isolate_cores", _oracle_isolate)


@pytest.fixture(autouse=True)
def _sep_stand_in():
    fake_sep.install()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        yield


@pytest.fixture(scope="module")
def golden():
    with np.load(GOLDEN) as z:
        return {k: z[k] for k in z.files}


def test_fixture_inputs_are_reproducible(golden):
    frames, mask = oracle.builder_frames()
    assert frames.shape == (int(golden["n_frames"]), *golden["shape"])
    assert frames.sum() == float(golden["frames_checksum"]) and mask.sum() == int(golden["mask_checksum"])


@pytest.mark.parametrize("method,pct", [("median", 50), ("mean", 50), ("percentile", 30)])
def test_host_pipeline_reproduces_the_reference_builder(golden, monkeypatch, method, pct):
    """Our cutouts -> cells -> (oracle averaging) -> core isolation == the reference builder's model, bit for bit."""
    _host_stages(monkeypatch)
    frames, mask = oracle.builder_frames()
    model, counts, patches = rp.ArrayPSFBuilder(32).build(frames, num_workers=1, average_method=method, percentile=pct,
                                                          image_mask=mask, return_patches=True)
    assert isinstance(model, rp.ArrayPSF) and model.sample_shape == (32, 32)
    assert len(patches) == int(golden["n_cutouts"])
    assert np.array_equal(np.array(model.coordinates), golden["coords"])
    assert [counts[tuple(c)] for c in rp.calculate_covering(frames.shape[1:], 32)] == list(golden["counts"])
    assert np.array_equal(model.values, golden[f"values_{method}"], equal_nan=True)


def test_worker_pool_gives_the_same_model(golden, monkeypatch):
    _host_stages(monkeypatch)
    frames, mask = oracle.builder_frames()
    model, _ = rp.ArrayPSFBuilder(32).build(frames, num_workers=2, image_mask=mask)
    assert np.array_equal(model.values, golden["values_median"], equal_nan=True)


def test_assign_to_cells_against_brute_force():
    rng = np.random.default_rng(5)
    corners = rp.calculate_covering((100, 90), 16)
    keys = [(int(rng.integers(0, 3)), float(rng.uniform(-8, 100)), float(rng.uniform(-8, 90))) for _ in range(200)]
    keys += [(0, float(corners[5][0]) - 8.0, float(corners[5][1]) - 8.0)]          # centre exactly on a cell corner
    offsets, items = b.assign_to_cells(keys, corners, 16)
    for c, (r0, c0) in enumerate(corners):
        want = [i for i, k in enumerate(keys)
                if r0 <= k[1] + 8 < r0 + 16 and c0 <= k[2] + 8 < c0 + 16]
        assert list(items[offsets[c]:offsets[c + 1]]) == want
    assert offsets[-1] == len(items)
    off0, it0 = b.assign_to_cells([], corners, 16)
    assert off0.tolist() == [0] * (len(corners) + 1) and len(it0) == 0


def test_argument_errors(monkeypatch):
    with pytest.raises(PSFBuilderError):
        b.average_cutouts(np.zeros((1, 8, 8)), np.array([0, 1]), np.array([0]), method="mode")
    _host_stages(monkeypatch)
    frames, _ = oracle.builder_frames(n_frames=1, shape=(96, 96))
    with pytest.raises(PSFBuilderError):
        rp.ArrayPSFBuilder(32).build(frames, average_method="mode")
    with pytest.raises(TypeError):
        rp.ArrayPSFBuilder(32).build("frames.fits")
    with pytest.raises(IncorrectShapeError):
        rp.ArrayPSFBuilder(32).build(np.zeros((2, 2, 8, 8)))
    with pytest.raises(PSFBuilderError):                                          # frames of different shapes
        rp.ArrayPSFBuilder(32).build((f for f in (frames[0], frames[0][:64])), num_workers=1)
    assert rp.ArrayPSFBuilder(16).psf_size == 16


def test_single_frame_and_generator_inputs(monkeypatch):
    _host_stages(monkeypatch)
    frames, _ = oracle.builder_frames(n_frames=2, shape=(128, 128), density=1 / 150)
    one, _ = rp.ArrayPSFBuilder(32).build(frames[0])                                # a bare 2-D frame
    gen, _ = rp.ArrayPSFBuilder(32).build((f for f in frames[:1]))
    cube, _ = rp.ArrayPSFBuilder(32).build(frames[:1])
    assert np.array_equal(one.values, gen.values, equal_nan=True) and np.array_equal(gen.values, cube.values, equal_nan=True)
    with pytest.raises(PSFBuilderError):          # every detection masked (the reference dies on an empty dict here)
        rp.ArrayPSFBuilder(32).build(frames[0], sep_mask=np.ones((1, 128, 128), dtype=bool))


@needs_reference
def test_cutouts_background_and_matches_identical_to_reference(monkeypatch):
    """detect_stars + the oracle's per-star body == the reference's _find_patches, bit for bit (the device body is
    compared with the oracle's in tests/test_gpu_builder.py)."""
    _host_stages(monkeypatch)
    ref = ref_loader.load_builder()
    frames, mask = oracle.builder_frames(n_frames=2, shape=(160, 128))
    for i, frame in enumerate(frames):
        want = ref.image_processing._find_patches(frame, 3, None, 1, 32, i, image_mask=mask, star_maximum=1e9)
        got = b.star_cutouts(frame, i, 32, 3, None, image_mask=mask, star_maximum=1e9)
        assert list(got) == list(want) and len(got) > 20
        for key in want:
            assert np.array_equal(got[key], want[key], equal_nan=True)
        sat = float(np.median([np.nanmax(p) for p in want.values()]))           # about half the stars are "saturated"
        low = float(np.percentile([p[16, 16] for p in want.values()], 20))
        picky = ref.image_processing._find_patches(frame, 3, None, 1, 32, i, saturation_threshold=sat, star_minimum=low)
        assert list(b.star_cutouts(frame, i, 32, 3, None, saturation_threshold=sat, star_minimum=low)) == list(picky)
        assert 0 < len(picky) < len(want)
    star = next(iter(want.values())) + 50.0
    star[3, 4] = 0.0
    assert np.array_equal(oracle.planar_background(star), ref.image_processing.calculate_background(star), equal_nan=True)
    corners = ref.util.calculate_covering((160, 128), 32)
    offsets, items = b.assign_to_cells(list(want), corners, 32)
    bounds_r = np.stack([corners[:, 0], corners[:, 0] + 32], axis=-1)
    bounds_c = np.stack([corners[:, 1], corners[:, 1] + 32], axis=-1)
    members = [[] for _ in corners]
    for j, key in enumerate(want):
        for c in ref.builder._find_matches(key, bounds_r, bounds_c, 32):
            members[c].append(j)
    assert [list(items[offsets[c]:offsets[c + 1]]) for c in range(len(corners))] == members


@needs_reference
@pytest.mark.parametrize("method,pct", [("mean", None), ("median", None), ("percentile", 50), ("percentile", 30),
                                        ("percentile", 99.5), ("percentile", 0), ("percentile", 100)])
def test_averaging_oracle_identical_to_reference(method, pct):
    ref = ref_loader.load_builder()
    rng = np.random.default_rng(17)
    size, shape = 16, (64, 80)
    corners = ref.util.calculate_covering(shape, size)
    keys = [(0, float(rng.uniform(-6, shape[0] - 10)), float(rng.uniform(-6, shape[1] - 10))) for _ in range(300)]
    keys += [(1, 20.25, 30.5)] * 1                                                  # a cell with extra depth
    stack = rng.normal(1.0, 0.5, size=(len(keys), size, size))
    stack[rng.random(stack.shape) < 0.15] = np.nan
    stack[rng.random(stack.shape) < 0.01] = np.inf
    stack[7] = np.nan                                                               # a fully masked cutout
    stack[9, size // 2, size // 2] = np.nan                                         # NaN centre: whole cutout drops out
    patches = dict(zip(keys, stack))
    averages, counts = ref.builder._average_patches(patches, corners, method=method, percentile=pct)
    offsets, items = b.assign_to_cells(keys, corners, size)
    mine = oracle.average_cutouts(stack, offsets, items, method, pct)
    for c, corner in enumerate(corners):
        assert np.array_equal(mine[c], averages[(corner[0], corner[1])]), (c, method, pct)
        assert counts[tuple(corner)] == offsets[c + 1] - offsets[c]


@needs_reference
def test_oracle_core_isolation_identical_to_the_reference_tail(golden):
    """oracle.isolate_core restates builder.py:239-258; run the reference's own build() on the golden frames and
    compare the model it returns with the oracle's tail applied to the reference's averaged patches."""
    ref = ref_loader.load_builder()
    frames, mask = oracle.builder_frames()
    patches = {}
    for i, frame in enumerate(frames):
        patches.update(ref.image_processing._find_patches(frame, 3, None, 1, 32, i, image_mask=mask))
    corners = ref.util.calculate_covering(frames.shape[1:], 32)
    averages, _ = ref.builder._average_patches(patches, corners, method="median", percentile=50)
    got = np.stack([oracle.isolate_core(np.array(p, dtype=np.float64)) for p in averages.values()])
    assert np.array_equal(got, golden["values_median"], equal_nan=True)
