"""Host-side containers and patch geometry.

``IndexedCube`` is the layout contract shared with the reference (regularizepsf/util.py:56-172):
``values[i]`` is the (P, P) sample whose upper-left corner is ``coordinates[i]`` = (row, col) in
unpadded frame coordinates.  ``calculate_covering`` reproduces regularizepsf/util.py:10-53.
Neither touches the GPU.
"""
from __future__ import annotations

import numpy as np

from regularizepsf_b200.exceptions import IncorrectShapeError, InvalidCoordinateError


def calculate_covering(image_shape: tuple[int, int], size: int) -> np.ndarray:
    """Corners of the four half-offset patch grids that cover every pixel exactly four times.

    Grid order is (0,0), (-h,-h), (-h,0), (0,-h) with h = ceil(size/2) and stride ``size``;
    within a grid the row coordinate varies fastest (regularizepsf/util.py:27-53).  The four
    grids are the four colour classes the overlap-add kernel runs as phases.
    """
    half = int(np.ceil(size / 2))
    blocks = []
    for row_start, col_start in ((0, 0), (-half, -half), (-half, 0), (0, -half)):
        rows = np.arange(row_start, image_shape[0], size)
        cols = np.arange(col_start, image_shape[1], size)
        grid = np.empty((cols.size, rows.size, 2), dtype=np.result_type(rows, cols))
        grid[..., 0] = rows[np.newaxis, :]
        grid[..., 1] = cols[:, np.newaxis]
        blocks.append(grid.reshape(-1, 2))
    return np.concatenate(blocks)


class IndexedCube:
    """A stack of equally shaped samples keyed by patch corner (regularizepsf/util.py:56-172)."""

    def __init__(self, coordinates, values: np.ndarray) -> None:
        if len(values.shape) != 3:
            raise IncorrectShapeError("Values must be three dimensional")
        if len(coordinates) != values.shape[0]:
            raise IncorrectShapeError(f"{len(coordinates)} coordinates defined but {values.shape[0]} values found.")
        self._coordinates = coordinates
        self._values = values
        self._index = {tuple(c): i for i, c in enumerate(coordinates)}

    @property
    def sample_shape(self) -> tuple[int, int]:
        return self._values.shape[1], self._values.shape[2]

    @property
    def coordinates(self):
        return self._coordinates

    @property
    def values(self) -> np.ndarray:
        return self._values

    def _lookup(self, coordinate) -> int:
        if coordinate not in self._index:
            raise InvalidCoordinateError(f"Coordinate {coordinate} not in TransferKernel.")
        return self._index[coordinate]

    def __getitem__(self, coordinate) -> np.ndarray:
        return self.values[self._lookup(coordinate)]

    def __setitem__(self, coordinate, value: np.ndarray) -> None:
        i = self._lookup(coordinate)
        if value.shape != self.sample_shape:
            raise IncorrectShapeError(
                f"Cannot assign value of shape {value.shape} to transfer kernel of shape {self.sample_shape}.")
        self.values[i] = value

    def __len__(self) -> int:
        return len(self.coordinates)

    def __eq__(self, other) -> bool:
        if not isinstance(other, IndexedCube):
            raise TypeError("Can only compare IndexedCube instances.")
        same_coords = self.coordinates == other.coordinates
        if isinstance(same_coords, np.ndarray):        # ndarray coordinates compare elementwise
            same_coords = bool(np.all(same_coords))
        return (bool(same_coords) and self.sample_shape == other.sample_shape
                and bool(np.allclose(self.values, other.values, rtol=1e-04, atol=1e-06)))

    __hash__ = None
