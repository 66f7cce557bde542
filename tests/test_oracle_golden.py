"""The CPU oracle must reproduce, bit for bit, outputs the reference itself generated
(tests/golden/*.npz, made by oracle/make_golden.py in the build container)."""
import numpy as np
import pytest

from oracle import cpu_oracle as oracle
from tests.helpers import golden_names, load_golden


def test_fixtures_exist():
    assert len(golden_names()) >= 15


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_generated_output(name):
    g = load_golden(name)
    s_fft = oracle.psf_fft(g["source"])
    t_fft = s_fft if np.array_equal(g["source"], g["target"]) else oracle.psf_fft(g["target"])
    kernel = oracle.transfer_kernel(s_fft, t_fft, g["alpha"], g["epsilon"])
    if "kernel" in g:
        assert np.array_equal(kernel, g["kernel"], equal_nan=True)
        assert np.array_equal(s_fft, g["source_fft"])
    out = oracle.apply_transform(g["image"], g["coords"], kernel, **g["apply_kwargs"])
    assert out.dtype == np.float64
    assert np.array_equal(out, g["out"], equal_nan=True)


def test_covering_known_answer():
    # util.py:27-53 worked by hand for a 4x4 frame and 2-px patches: h = 1
    got = oracle.covering((4, 4), 2)
    first_grid = [(0, 0), (2, 0), (0, 2), (2, 2)]
    assert [tuple(c) for c in got[:4]] == first_grid
    assert [tuple(c) for c in got[4:13]] == [(r, c) for c in (-1, 1, 3) for r in (-1, 1, 3)]
    assert len(got) == 4 + 9 + 6 + 6
