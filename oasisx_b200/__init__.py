"""oasisx_b200 -- the oasisx IPCS fractional-step time loop on B200 (sm_100a).

Same public names as ``/root/reference/src/oasisx/__init__.py:12-18``.  Importing the package does
not need a GPU; constructing a solver does (there is no CPU fallback).
"""
import logging

from .bcs import DirichletBC, LocatorMethod, PressureBC
from .fracstep import FractionalStep_AB_CN
from .function import Projector

logger = logging.getLogger("oasisx")

__all__ = [
    "Projector",
    "FractionalStep_AB_CN",
    "DirichletBC",
    "LocatorMethod",
    "PressureBC",
]
