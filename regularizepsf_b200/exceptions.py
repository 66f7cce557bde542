"""Error types at the API boundary.

Same names and hierarchy as the reference (regularizepsf/exceptions.py:4-23) so callers'
``except`` clauses keep working; ``NativeLibraryError`` is new and is raised when the CUDA
library is missing or no B200 is visible — this package has no CPU fallback.

``status`` is the C-ABI return code (include/rpsf_b200.h, ``RPSF_E_*``) that surfaces as this
exception; ``None`` marks the types only the Python layer raises.  ``for_status`` is the reverse
lookup used by ``_native.check``.
"""
from __future__ import annotations


class RegularizePSFError(Exception):
    """Root of every error this package raises on purpose."""

    status: int | None = None


class InvalidCoordinateError(RegularizePSFError):
    """A coordinate is not a key of the model, or source/target coordinates disagree."""

    status = -2                      # RPSF_E_INVALID_COORDINATE


class IncorrectShapeError(RegularizePSFError):
    """An array does not have the shape the model requires."""

    status = -3                      # RPSF_E_INCORRECT_SHAPE


class NativeLibraryError(RegularizePSFError):
    """librpsf_b200.so is missing/unloadable, no CUDA device is available, or a CUDA call failed."""

    status = -5                      # RPSF_E_CUDA (and RPSF_E_NO_KERNEL, -6)


class InvalidFunctionError(RegularizePSFError):
    """A functional PSF has an invalid signature."""


class FunctionParameterMismatchError(RegularizePSFError):
    """A functional PSF was evaluated with unknown keyword arguments."""


class PSFBuilderError(RegularizePSFError):
    """PSF model building failed."""


class InvalidDataError(RegularizePSFError):
    """Input data for PSF building is invalid."""


def for_status(code: int) -> type[Exception]:
    """Exception type for a negative C-ABI status: the reference's types where it has one, else built-ins."""
    table = {cls.status: cls for cls in (InvalidCoordinateError, IncorrectShapeError, NativeLibraryError)}
    table[-6] = NativeLibraryError          # RPSF_E_NO_KERNEL
    table[-4] = NotImplementedError         # RPSF_E_UNSUPPORTED
    return table.get(code, ValueError)      # RPSF_E_INVALID_ARGUMENT and anything unknown
