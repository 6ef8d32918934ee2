"""prestige_b200 -- B200-native particle hot path (NNPS + SPH/DEM pair forces) behind prestige's
equation API.  The compute lives in libprestige_b200.so (hand-written sm_100a CUDA, C ABI in
include/prestige_b200.h); this package is the host-side mirror of the reference interface.
"""
from . import synth  # noqa: F401
from .equations import (EquationIR, FusedEquations, body_reduce, continuity, debug_equation, dem_contact, eq1, equation, fuse,  # noqa: F401
                        momentum, tait_eos, wall_pressure)
from . import codegen, decomp, io  # noqa: F401
from ._lib import PstError, LIB_PATH  # noqa: F401
from .context import Context, context_for_block  # noqa: F401

__version__ = "0.1.0"
