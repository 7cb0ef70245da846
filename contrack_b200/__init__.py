"""contrack_b200 -- B200-native implementation of ConTrack's run_contrack() tracking path.

``from contrack_b200 import contrack`` (or ``from contrack import contrack`` through the shim package at the repo root)
gives the reference's class interface; the work is done by ``lib/libcontrack_b200.so`` (include/contrack_b200.h).
"""
from .contrack import contrack, ContrackLibError          # noqa: F401
from .dataset import DataArray, Dataset, Variable          # noqa: F401
from .engine import Engine                                 # noqa: F401

__version__ = '0.1.0'
