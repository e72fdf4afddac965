"""pyratbay_b200 -- B200-native line-by-line opacity engine behind Pyrat Bay's API surface
for the opacity path (TLI input, Voigt / Line_By_Line / Line_Sample objects, .npz tables).

Python host code calls hand-written sm_100a CUDA through the C ABI of
include/pb200_lbl.h (pyratbay_b200/libpb200_lbl.so).  There is no CPU fallback.
"""
from . import constants  # noqa: F401
from . import io  # noqa: F401
from . import tli  # noqa: F401
from .spectrum import Spectrum  # noqa: F401
from .atmosphere import Atmosphere  # noqa: F401
from .engine import Engine, voigt_grid, interp_ec, interp_ec_per_mol  # noqa: F401
from .voigt import Voigt  # noqa: F401
from .line_by_line import Line_By_Line  # noqa: F401
from .line_sampling import Line_Sample  # noqa: F401
from .pyrat import Pyrat, run  # noqa: F401
from . import extinction  # noqa: F401

__version__ = "0.1.0"
