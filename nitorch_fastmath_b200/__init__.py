"""nitorch_fastmath_b200 -- the batched small-matrix hot path of
nitorch-fastmath (compact-symmetric ``sym`` routines and dense ``batched``
routines) rebuilt as hand-written CUDA for NVIDIA B200 (sm_100a) behind the
reference's Python signatures.  See DESIGN.md.
"""
from . import batched, multi_gpu, sugar, sym          # noqa: F401
from .batched import *                     # noqa: F401,F403
from .sugar import lmdiv, rmdiv, inv, matvec, solvevec   # noqa: F401
from .sym import *                         # noqa: F401,F403

__version__ = "0.1.0"
