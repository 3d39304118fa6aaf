"""bigseqkit_b200 -- B200-native engine for BigSeqKit's per-record hot path.

The package is a thin host layer over ``libbsk.so`` (CUDA, sm_100a; C ABI in
``include/bsk.h``).  Importing it does not need a GPU; running an operator does.
"""
from .api import (BskError, Library, Operator, Result, default_library, fq2fa, grep, locate, make_opts, rmDup, seq, stats,  # noqa: F401
                  subseq, translate, LIB_PATH, ABI_SYMBOLS)

from . import synth  # noqa: F401,E402

__all__ = ["BskError", "Library", "Operator", "Result", "default_library", "fq2fa", "grep", "locate", "make_opts", "rmDup", "seq",
           "stats", "subseq", "translate", "LIB_PATH", "ABI_SYMBOLS"]
