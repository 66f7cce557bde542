"""Error types at the API boundary.

Same names and hierarchy as the reference (regularizepsf/exceptions.py:4-23) so callers'
``except`` clauses keep working; ``NativeLibraryError`` is new and is raised when the CUDA
library is missing or no B200 is visible — this package has no CPU fallback.
"""


class RegularizePSFError(Exception):
    """Root of every error this package raises on purpose."""


class InvalidCoordinateError(RegularizePSFError):
    """A coordinate is not a key of the model, or source/target coordinates disagree."""


class IncorrectShapeError(RegularizePSFError):
    """An array does not have the shape the model requires."""


class InvalidFunctionError(RegularizePSFError):
    """A functional PSF has an invalid signature."""


class FunctionParameterMismatchError(RegularizePSFError):
    """A functional PSF was evaluated with unknown keyword arguments."""


class PSFBuilderError(RegularizePSFError):
    """PSF model building failed."""


class InvalidDataError(RegularizePSFError):
    """Input data for PSF building is invalid."""


class NativeLibraryError(RegularizePSFError):
    """librpsf_b200.so is missing/unloadable, or no CUDA device is available."""
