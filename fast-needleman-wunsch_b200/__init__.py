"""B200-native Needleman-Wunsch wavefront fill: Python glue over libnw_cuda.so (include/nw_cuda.h).

The product is the C-ABI shared library built from csrc/ (sm_100a CUDA); this package only loads it with ctypes for
the tests and bench.py.  Import it with importlib (the directory name has a hyphen):

    nw = importlib.import_module("fast-needleman-wunsch_b200")
"""
from .nwcuda import (  # noqa: F401
    NW_MODE_BOUNDARY, NW_MODE_FULL, NW_MODE_SCORE, NwCudaError, Plan, Batch, Scoring, best, align, plans_traceback, lib, lib_path,
    readSequence, printSequence, needlemanWunsch, score, boundaries, batch_scores, dpx_peak, device_count, device_info, init,
    strip_partition,
)
