"""gappadder_b200 -- B200-native overlap alignment for GAPPadder's ContigsMerger.

The product is the C-ABI shared library ``libgappadder_b200.so`` (CUDA kernels for sm_100a plus the
C++ host half; see include/gappadder_b200.h) and the ``ContigsMerger_b200`` drop-in binary built
from gappadder_b200/host.  This Python package is only a ctypes view of that ABI for tests and
bench.py.  There is no CPU fallback: loading fails loudly when the library has not been built, and
creating a context fails when no sm_100 device is present.
"""
from .capi import (  # noqa: F401
    GpError, Context, DpParams, Thresholds, Pair, Result, lib, lib_path,
    pack_sequences, candidate_pairs, revcomp, estimate_gap_cells, partition_gaps, is_score_significant, merged_concat,
    GAPPADDER_DP, gappadder_thresholds, dedup_unique_names, dedup_decide, dedup_records, DEDUP_RECORD_DTYPE,
    AffineParams, TEREFINER_AFFINE, LOCAL_DTYPE,
)
