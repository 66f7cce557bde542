"""Shared helpers for the test-suite (not product code)."""
from __future__ import annotations

import ast
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name: str) -> dict:
    with np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")) as z:
        d = {k: z[k] for k in z.files}
    d["coords"] = [tuple(int(v) for v in c) for c in d["coords"]]
    d["size"] = int(d["size"])
    d["alpha"] = float(d["alpha"])
    d["epsilon"] = float(d["epsilon"])
    d["apply_kwargs"] = ast.literal_eval(str(d["apply_kwargs"]))
    return d


def make_gaussian(size, fwhm=3, center=None):
    """Unnormalised Gaussian, same recipe as the reference's tests/helper.py:4-21."""
    x = np.arange(0, size, 1, float)
    y = x[:, np.newaxis]
    x0, y0 = (size // 2, size // 2) if center is None else (center[0], center[1])
    return np.exp(-4 * np.log(2) * ((x - x0) ** 2 + (y - y0) ** 2) / fwhm ** 2)


def rel_err(got, want, scale) -> float:
    return float(np.nanmax(np.abs(np.asarray(got, dtype=np.float64) - want)) / scale)
