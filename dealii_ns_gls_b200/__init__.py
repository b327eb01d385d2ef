"""B200-native matrix-free GLS Navier-Stokes operator (hot path of peterrum/dealii-ns-gls).

Importing the package does not import torch or load the CUDA library; constructing a
``NavierStokesOperator`` does, and raises if libglsb200.so or a CUDA device is missing."""

from .time_integration import (SolutionHistory, TimeIntegratorDataBDF, TimeIntegratorDataNone,  # noqa: F401
                               TimeIntegratorDataTheta)


def __getattr__(name):
    if name in ("NavierStokesOperator", "AffineConstraints", "build_desc"):
        from . import operator as _o
        return getattr(_o, name)
    raise AttributeError(name)
