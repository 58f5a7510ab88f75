"""score_b200 — B200-native implementation of SCORE's ``solve_score`` hot path."""
from . import pyfg_shim as _pyfg_shim

_pyfg_shim.install()

__version__ = "0.1.0"
