"""CUDA backend: same exported names as ``bionc/bionc_numpy/__init__.py:1-27`` for the
constrained forward-dynamics path (everything else of bioNC is out of scope).

``BiomechanicalModel`` (which needs torch and the CUDA library) is imported lazily so that the pure
numpy pieces -- the model table and the force set -- stay importable in light-weight processes.
"""

from .model_table import ModelTable, UnsupportedModelError  # noqa: F401
from .external_force import ExternalForceSet  # noqa: F401
from .natural_types import (  # noqa: F401
    NaturalAccelerations,
    NaturalCoordinates,
    NaturalVelocities,
    SegmentNaturalAccelerations,
    SegmentNaturalCoordinates,
    SegmentNaturalVelocities,
)

__all__ = [
    "BiomechanicalModel", "BioncCudaError", "NaturalSegment", "TimeSeriesUtils", "ExternalForceSet", "ModelTable", "UnsupportedModelError",
    "NaturalCoordinates", "NaturalVelocities", "NaturalAccelerations",
    "SegmentNaturalCoordinates", "SegmentNaturalVelocities", "SegmentNaturalAccelerations",
]


def __getattr__(name):
    if name == "BiomechanicalModel":
        from .biomechanical_model import BiomechanicalModel

        return BiomechanicalModel
    if name == "NaturalSegment":
        from .natural_segment import NaturalSegment

        return NaturalSegment
    if name == "TimeSeriesUtils":
        from .time_series_utils import TimeSeriesUtils

        return TimeSeriesUtils
    if name == "BioncCudaError":
        from ._lib import BioncCudaError

        return BioncCudaError
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
