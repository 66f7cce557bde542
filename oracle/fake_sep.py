"""A deterministic stand-in for the ``sep`` source extractor (absent from this image).

TEST INFRASTRUCTURE ONLY.  The reference's star finder (image_processing.py:62-76) calls
``sep.Background(image)``, ``image - background``, ``background.globalrms`` and
``sep.extract(data, thresh, err=..., mask=...)`` and uses only the ``"x"`` / ``"y"`` fields of the
result.  This module provides exactly that surface with plain numpy/scipy so that the UNMODIFIED
reference builder and this repository's builder can be fed the *same* detections:

* ``Background``: a constant level (the median) and the robust rms (1.4826 x MAD);
* ``extract``: pixels above ``thresh * err`` that are the maximum of their 5x5 neighbourhood,
  refined to sub-pixel positions by a 3x3 intensity centroid (so the reference's spline shift is
  exercised with fractional offsets), returned in raster order.

``install()`` registers it as ``sys.modules["sep"]``.
"""
from __future__ import annotations

import sys
import types

import numpy as np
from scipy.ndimage import maximum_filter


class Background:
    __array_ufunc__ = None            # make ``ndarray - Background`` defer to __rsub__

    def __init__(self, image, **_ignored):
        data = np.asarray(image, dtype=np.float64)
        self.level = float(np.median(data))
        self.globalrms = float(1.4826 * np.median(np.abs(data - self.level)))
        self.globalback = self.level

    def __rsub__(self, image):
        return np.asarray(image, dtype=np.float64) - self.level

    def back(self):
        return self.level


def extract(data, thresh, err=None, mask=None, **_ignored):
    data = np.asarray(data, dtype=np.float64)
    level = thresh * (1.0 if err is None else float(err))
    peak = (data > level) & (data == maximum_filter(data, size=5, mode="constant", cval=-np.inf))
    if mask is not None:
        peak &= ~np.asarray(mask, dtype=bool)
    peak[0, :] = peak[-1, :] = False
    peak[:, 0] = peak[:, -1] = False
    rows, cols = np.nonzero(peak)
    out = np.zeros(len(rows), dtype=[("x", np.float64), ("y", np.float64), ("peak", np.float64)])
    for i, (r, c) in enumerate(zip(rows, cols)):
        win = np.clip(data[r - 1:r + 2, c - 1:c + 2], 0.0, None)
        total = win.sum()
        dy = float((win.sum(axis=1) * np.array([-1.0, 0.0, 1.0])).sum() / total)
        dx = float((win.sum(axis=0) * np.array([-1.0, 0.0, 1.0])).sum() / total)
        out[i] = (c + dx, r + dy, data[r, c])
    return out


def install() -> types.ModuleType:
    mod = types.ModuleType("sep")
    mod.Background = Background
    mod.extract = extract
    mod.__fake__ = True
    sys.modules["sep"] = mod
    return mod
