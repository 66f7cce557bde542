"""Functional PSF models: user callables ``f(row, col, **params)`` turned into ``ArrayPSF`` cubes.

API of regularizepsf/psf.py:25-189 (``SimpleFunctionalPSF``, ``VariedFunctionalPSF`` and the two
decorators): same names, same signature rules, same exception types.  These are host-side model
*setup* helpers — arbitrary Python callables evaluated once per patch — so nothing here touches the
GPU; the ``ArrayPSF`` they return computes its FFT cube on the device the first time it is needed.
"""
from __future__ import annotations

import functools
import inspect
from typing import Any, Callable

import numpy as np

from regularizepsf_b200.exceptions import InvalidFunctionError
from regularizepsf_b200.psf import ArrayPSF
from regularizepsf_b200.util import IndexedCube


def _leading_names(function: Callable, what: str) -> list[str]:
    """Parameter names of ``function``; the first two must be ``row`` then ``col``."""
    names = list(inspect.signature(function).parameters)
    if len(names) < 2:
        raise InvalidFunctionError(f"row and col must be the first two arguments in your {what} equation.")
    for position, expected in enumerate(("row", "col")):
        if names[position] != expected:
            raise InvalidFunctionError(
                f"{expected} must be argument {position + 1} of your {what} equation, found {names[position]!r}.")
    return names


def _sample_grid(size: int) -> tuple[np.ndarray, np.ndarray]:
    # psf.py:67,162 evaluates on np.meshgrid(arange, arange): xy indexing, i.e. `row` varies along axis 1
    return tuple(np.meshgrid(np.arange(size), np.arange(size)))


class SimpleFunctionalPSF:
    """A PSF given by one function of (row, col) and optional keyword parameters (psf.py:25-75)."""

    def __init__(self, function: Callable) -> None:
        self._f = function
        self._signature = inspect.signature(function)
        self._parameters = set(_leading_names(function, "model")[2:])

    def __call__(self, row, col, **kwargs: Any):
        return self._f(row, col, **kwargs)

    @property
    def parameters(self) -> set[str]:
        return self._parameters

    @property
    def f(self) -> Callable:
        return self._f

    def as_array_psf(self, coordinates, size: int, **kwargs: Any) -> ArrayPSF:
        """The same sampled patch at every coordinate."""
        rr, cc = _sample_grid(size)
        patch = np.asarray(self(rr, cc, **kwargs))
        return ArrayPSF(IndexedCube(coordinates, np.stack([patch] * len(coordinates))))


def simple_functional_psf(arg: Any = None) -> SimpleFunctionalPSF:
    """Decorator: ``@simple_functional_psf`` (no arguments) over ``f(row, col, ...)``."""
    if not callable(arg):
        raise TypeError("psf decorator must have no arguments.")
    return SimpleFunctionalPSF(arg)


class VariedFunctionalPSF:
    """A base PSF whose parameters are a function of position in the image (psf.py:86-165)."""

    def __init__(self, vary_function: Callable, base_psf: SimpleFunctionalPSF, validate_at_call: bool = True) -> None:
        self._vary_function = vary_function
        self._base_psf = base_psf
        self.validate_at_call = validate_at_call
        self.parameterization_signature = inspect.signature(vary_function)
        names = _leading_names(vary_function, "parameterization")
        if len(names) > 2:
            raise InvalidFunctionError(
                f"Found function requiring {len(names)} arguments. Expected 2, only `row` and `col`.")
        supplied = set(vary_function(0, 0).keys())          # probe the parameter names at the origin
        if supplied != base_psf.parameters:
            raise InvalidFunctionError(
                f"The base PSF model has parameters {base_psf.parameters} while the varied psf supplies "
                f"{supplied} at the origin. These must match.")

    @property
    def parameters(self) -> set[str]:
        return self._base_psf.parameters

    def _parameters_at(self, row, col) -> dict[str, Any]:
        params = self._vary_function(row, col)
        if self.validate_at_call and set(params.keys()) != self.parameters:
            raise InvalidFunctionError(
                f"At (row, col) the varying parameters were {set(params.keys())} "
                f"when the parameters were expected as {self.parameters}.")
        return params

    def __call__(self, row, col):
        return self._base_psf(row, col, **self._parameters_at(row, col))

    def simplify(self, row: int, col: int) -> SimpleFunctionalPSF:
        """Freeze the parameters at (row, col)."""
        return SimpleFunctionalPSF(functools.partial(self._base_psf.f, **self._vary_function(row, col)))

    def as_array_psf(self, coordinates, size: int, **kwargs: Any) -> ArrayPSF:
        """One sampled patch per coordinate, with the parameters evaluated at that coordinate."""
        rr, cc = _sample_grid(size)
        patches = [np.asarray(self.simplify(row, col)(rr, cc, **kwargs)) for row, col in coordinates]
        return ArrayPSF(IndexedCube(coordinates, np.stack(patches)))


def varied_functional_psf(base_psf: SimpleFunctionalPSF = None):
    """Decorator factory: ``@varied_functional_psf(base)`` over ``f(row, col) -> dict of parameters``."""
    if not isinstance(base_psf, SimpleFunctionalPSF):
        if callable(base_psf):
            raise TypeError("varied_psf decorator must be called with an argument for the base_psf.")
        raise TypeError("varied_psf decorator expects exactly one argument of type PSF.")

    def decorate(function: Callable | None = None, *, check_at_call: bool = True):
        if function is None:
            return functools.partial(decorate, check_at_call=check_at_call)
        return VariedFunctionalPSF(function, base_psf, validate_at_call=check_at_call)

    return decorate
