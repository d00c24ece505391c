"""bcalm_b200 -- B200-native compacted de Bruijn graph construction (the BCALM 2 reads -> unitigs path).

The product is the CUDA library bcalm_b200/libcdbg_b200.so behind the C ABI in include/cdbg.h; this package is the thin
Python host mirror (ctypes) plus the parity relation used to compare graphs.  See DESIGN.md.
"""
from .api import Cdbg, CdbgError, Graph, PinnedBuffer, build_unitigs, load_library  # noqa: F401

__version__ = "0.1"
