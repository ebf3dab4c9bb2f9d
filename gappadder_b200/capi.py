"""ctypes binding of include/gappadder_b200.h (one function per exported symbol)."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GAPPADDER_B200_LIB: another build of the SAME library (the diagnostic `make trace` build, build/libgappadder_b200_trace.so);
# anything that is not a libgappadder_b200*.so is refused
lib_path = os.path.join(_HERE, "libgappadder_b200.so")
_override = os.environ.get("GAPPADDER_B200_LIB")
if _override:
    if not os.path.basename(_override).startswith("libgappadder_b200"):
        raise ImportError("GAPPADDER_B200_LIB must name a build of libgappadder_b200 (got %r)" % _override)
    lib_path = os.path.abspath(_override)

# every symbol include/gappadder_b200.h declares (tests check that each one is exported)
EXPORTS = [
    "gp_create", "gp_destroy", "gp_last_error", "gp_abi_version", "gp_stream",
    "gp_packed_size", "gp_pack_sequences", "gp_set_sequences", "gp_overlap_pairs", "gp_overlap_batch",
    "gp_upload_pairs", "gp_launch_resident", "gp_fetch_results", "gp_kernel_launches", "gp_pair_stats",
    "gp_is_score_significant", "gp_is_containment", "gp_merged_length", "gp_merged_concat",
    "gp_overlap_size", "gp_candidate_pairs", "gp_revcomp", "gp_int_peak",
    "gp_estimate_gap_cells", "gp_partition_gaps", "gp_pair_split", "gp_set_kernel_mask", "gp_last_timing",
    "gp_cert_stats", "gp_set_cert_system", "gp_kernel_times",
    "gp_closed_form_stats", "gp_set_team_mode", "gp_last_team", "gp_set_cert_layout", "gp_last_layout", "gp_quick_check_device", "gp_upload_sequences",
    "gp_quick_check_stats",
    "gp_reserve", "gp_relax_chains", "gp_relax_stats", "gp_set_orientation", "gp_transposed_pairs",
    "gp_semiglobal_batch", "gp_semiglobal_upload_pairs", "gp_semiglobal_launch", "gp_semiglobal_fetch", "gp_semiglobal_stats",
    "gp_dedup_unique_names", "gp_dedup_decide", "gp_dedup_records", "gp_quick_check_matrix", "gp_set_relax_launch_hook",
    "gp_affine_params_terefiner", "gp_local_affine_batch", "gp_local_affine_upload_pairs", "gp_local_affine_launch",
    "gp_local_affine_fetch", "gp_local_affine_stats",
]


class GpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("gappadder_b200 error %d: %s" % (code, msg))
        self.code = code


class Pair(C.Structure):
    _fields_ = [("row_seq", C.c_uint32), ("col_seq", C.c_uint32)]


class Result(C.Structure):
    _fields_ = [("score", C.c_int32), ("row_end", C.c_int32), ("col_end", C.c_int32),
                ("nclip", C.c_int32), ("flags", C.c_uint32)]


class DpParams(C.Structure):
    _fields_ = [("mismatch", C.c_int32), ("indel", C.c_int32), ("max_clip", C.c_int32)]


class AffineParams(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("n_score", C.c_int32),
                ("gap_open", C.c_int32), ("gap_ext", C.c_int32), ("band_width", C.c_int32)]


class Thresholds(C.Structure):
    _fields_ = [("fraction_loss_score", C.c_double), ("frac_min_overlap", C.c_double),
                ("min_overlap_len", C.c_double), ("min_overlap_len_with_scaffold", C.c_double)]


RESULT_DTYPE = np.dtype([("score", "<i4"), ("row_end", "<i4"), ("col_end", "<i4"), ("nclip", "<i4"), ("flags", "<u4")])
PAIR_DTYPE = np.dtype([("row_seq", "<u4"), ("col_seq", "<u4")])
DEDUP_RECORD_DTYPE = np.dtype([("q", "<u4"), ("r", "<u4"), ("single_m", "<u4"), ("m_len", "<u4"), ("other_len", "<u4")])
PLACE_DTYPE = np.dtype([("score", "<i4"), ("col_start", "<i4"), ("col_end", "<i4"), ("flags", "<u4")])
LOCAL_DTYPE = np.dtype([("score", "<i4"), ("start1", "<i4"), ("end1", "<i4"), ("start2", "<i4"), ("end2", "<i4"), ("flags", "<u4")])
LOCAL_NO_MATCH, LOCAL_UNDEFINED, LOCAL_POTENTIAL_BUG = 1, 2, 4
# aln_param_blast (TERefiner/algorithms/local_alignment.cpp:193-206), what LocalAlignment::optAlign uses
TEREFINER_AFFINE = AffineParams(1, -3, -2, 5, 2, 50)
FLAG_ROW0, FLAG_COL0, FLAG_CONTAINED, FLAG_KERNEL16, FLAG_CLOSED = 1, 2, 4, 8, 16
KERNEL_TABLE16, KERNEL_PRMT16, KERNEL_CERT16, KERNEL_CLOSED, KERNEL_ALL = 1, 2, 4, 8, 15
KERNEL_DP_ALL = KERNEL_TABLE16 | KERNEL_PRMT16 | KERNEL_CERT16      # every kernel, no closed form

# GAPPadder's command line (MergeContigs.py:85): -s 0.4 -i1 -2.0 -i2 -2.0 -x 12 -y 50 -k 10 -m 1
GAPPADDER_DP = DpParams(-2, -2, 50)


def gappadder_thresholds() -> Thresholds:
    # -s and -x pass through a float in CM/main.cpp:91-93,108-111; -c and -z keep their defaults
    return Thresholds(float(np.float32(0.4)), 0.005, float(np.float32(12)), 6.0)


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Loads libgappadder_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(lib_path):
            raise GpError(-100, "%s not found: run `make` (or __graft_entry__.build()) first; there is no CPU fallback" % lib_path)
        L = C.CDLL(lib_path)
        L.gp_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.gp_destroy.argtypes = [C.c_void_p]
        L.gp_destroy.restype = None
        L.gp_last_error.argtypes = [C.c_void_p]
        L.gp_last_error.restype = C.c_char_p
        L.gp_stream.argtypes = [C.c_void_p]
        L.gp_stream.restype = C.c_void_p
        L.gp_packed_size.argtypes = [C.c_void_p, C.c_uint32]
        L.gp_packed_size.restype = C.c_size_t
        L.gp_pack_sequences.argtypes = [C.POINTER(C.c_char_p), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
        L.gp_set_sequences.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.gp_overlap_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(DpParams), C.c_void_p]
        L.gp_overlap_batch.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(DpParams), C.c_void_p]
        L.gp_upload_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(DpParams)]
        L.gp_launch_resident.argtypes = [C.c_void_p]
        L.gp_fetch_results.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.gp_kernel_launches.argtypes = [C.c_void_p]
        L.gp_kernel_launches.restype = C.c_uint64
        L.gp_pair_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.gp_is_score_significant.argtypes = [C.POINTER(Thresholds)] + [C.c_int32] * 6
        L.gp_is_containment.argtypes = [C.c_int32, C.c_int32, C.POINTER(Result)]
        L.gp_merged_length.argtypes = [C.c_int32, C.c_int32, C.POINTER(Result)]
        L.gp_merged_length.restype = C.c_int32
        L.gp_merged_concat.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.POINTER(Result), C.c_char_p]
        L.gp_merged_concat.restype = C.c_int32
        L.gp_overlap_size.argtypes = [C.c_int32, C.c_int32, C.POINTER(Result)]
        L.gp_overlap_size.restype = C.c_int32
        L.gp_candidate_pairs.argtypes = [C.POINTER(C.c_char_p), C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p, C.c_uint64]
        L.gp_candidate_pairs.restype = C.c_int64
        L.gp_revcomp.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p]
        L.gp_revcomp.restype = None
        L.gp_int_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.gp_estimate_gap_cells.argtypes = [C.c_void_p, C.c_uint32]
        L.gp_estimate_gap_cells.restype = C.c_uint64
        L.gp_partition_gaps.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.c_void_p]
        L.gp_pair_split.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.gp_set_kernel_mask.argtypes = [C.c_void_p, C.c_uint32]
        L.gp_cert_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.gp_set_cert_system.argtypes = [C.c_void_p, C.c_uint32]
        L.gp_kernel_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.gp_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        L.gp_closed_form_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.gp_set_team_mode.argtypes = [C.c_void_p, C.c_uint32]
        L.gp_last_team.argtypes = [C.c_void_p]
        L.gp_set_cert_layout.argtypes = [C.c_void_p, C.c_uint32]
        L.gp_last_layout.argtypes = [C.c_void_p]
        L.gp_quick_check_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p, C.c_uint64]
        L.gp_upload_sequences.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_uint32]
        L.gp_quick_check_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.gp_set_orientation.argtypes = [C.c_void_p, C.c_uint32]
        L.gp_transposed_pairs.argtypes = [C.c_void_p]
        L.gp_transposed_pairs.restype = C.c_uint64
        L.gp_semiglobal_batch.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(DpParams), C.c_void_p]
        L.gp_semiglobal_upload_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(DpParams)]
        L.gp_semiglobal_launch.argtypes = [C.c_void_p]
        L.gp_affine_params_terefiner.argtypes = [C.POINTER(AffineParams)]
        L.gp_affine_params_terefiner.restype = None
        L.gp_local_affine_batch.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(AffineParams), C.c_void_p]
        L.gp_local_affine_upload_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(AffineParams)]
        L.gp_local_affine_launch.argtypes = [C.c_void_p]
        L.gp_local_affine_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.gp_local_affine_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.gp_semiglobal_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.gp_semiglobal_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _seq_arrays(seqs: Sequence[bytes]):
    n = len(seqs)
    arr = (C.c_char_p * n)(*seqs)
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    return arr, lens


def pack_sequences(seqs: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int]:
    """-> (packed uint32 words, word offsets, lengths, n_symbols)"""
    L = lib()
    arr, lens = _seq_arrays(seqs)
    nbytes = L.gp_packed_size(lens.ctypes.data, len(seqs))
    packed = np.zeros(nbytes // 4, dtype=np.uint32)
    off = np.zeros(len(seqs), dtype=np.uint32)
    nsym = C.c_uint32(0)
    rc = L.gp_pack_sequences(arr, lens.ctypes.data, len(seqs), packed.ctypes.data, off.ctypes.data, C.byref(nsym))
    if rc != 0:
        raise GpError(rc, "gp_pack_sequences")
    return packed, off, lens, int(nsym.value)


def candidate_pairs(nodes: Sequence[bytes], k: int = 10) -> np.ndarray:
    L = lib()
    arr, lens = _seq_arrays(nodes)
    n = len(nodes)
    cap = n * (n + 1) // 2
    out = np.zeros(cap, dtype=PAIR_DTYPE)
    cnt = L.gp_candidate_pairs(arr, lens.ctypes.data, n, k, out.ctypes.data, cap)
    if cnt < 0:
        raise GpError(int(cnt), "gp_candidate_pairs")
    return out[:cnt]


def _name_array(names: Sequence[bytes]):
    keep = [bytes(n) + b"\0" for n in names]
    arr = (C.c_char_p * max(1, len(names)))(*[C.c_char_p(k) for k in keep])
    return arr, keep


def dedup_unique_names(names: Sequence[bytes]) -> np.ndarray:
    """TERefiner_1 -U (refiner.cpp:1045-1140): keep[i] = 1 for the first record of every name."""
    L = lib()
    arr, _keep = _name_array(names)
    keep = np.zeros(max(1, len(names)), dtype=np.uint8)
    L.gp_dedup_unique_names.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    rc = L.gp_dedup_unique_names(arr, len(names), keep.ctypes.data)
    if rc != 0:
        raise GpError(rc, "gp_dedup_unique_names")
    return keep[:len(names)]


def dedup_decide(records: np.ndarray, names: Sequence[bytes], lens, cutoff: float, remove_contained: bool) -> np.ndarray:
    """TERefiner_1 -P [-g] -c cutoff (refiner.cpp:660-801) over alignment records (DEDUP_RECORD_DTYPE) -> removed[i]."""
    L = lib()
    arr, _keep = _name_array(names)
    recs = np.ascontiguousarray(records, dtype=DEDUP_RECORD_DTYPE)
    ln = np.ascontiguousarray(lens, dtype=np.uint32)
    removed = np.zeros(max(1, len(names)), dtype=np.uint8)
    L.gp_dedup_decide.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_int, C.c_void_p]
    rc = L.gp_dedup_decide(recs.ctypes.data, len(recs), arr, ln.ctypes.data, len(names), float(cutoff), int(bool(remove_contained)), removed.ctypes.data)
    if rc != 0:
        raise GpError(rc, "gp_dedup_decide")
    return removed[:len(names)]


def dedup_records(q: int, r: int, len_q: int, len_r: int, res, max_frac_score_loss: float = 0.4) -> np.ndarray:
    """The two alignment records one Evaluate result stands for (builder-defined; see the header)."""
    L = lib()
    out = np.zeros(2, dtype=DEDUP_RECORD_DTYPE)
    rr = Result(int(res["score"]), int(res["row_end"]), int(res["col_end"]), int(res["nclip"]), int(res["flags"]))
    L.gp_dedup_records.argtypes = [C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, C.c_void_p, C.c_double, C.c_void_p]
    n = L.gp_dedup_records(q, r, len_q, len_r, C.byref(rr), float(max_frac_score_loss), out.ctypes.data)
    return out[:n]


def estimate_gap_cells(contig_lens) -> int:
    a = np.ascontiguousarray(contig_lens, dtype=np.uint32)
    return int(lib().gp_estimate_gap_cells(a.ctypes.data, len(a)))


def partition_gaps(costs, n_parts: int) -> np.ndarray:
    """Longest-processing-time assignment of gaps to n_parts GPUs -> part index per gap."""
    c = np.ascontiguousarray(costs, dtype=np.uint64)
    part = np.zeros(len(c), dtype=np.int32)
    rc = lib().gp_partition_gaps(c.ctypes.data, len(c), n_parts, part.ctypes.data)
    if rc != 0:
        raise GpError(rc, "gp_partition_gaps")
    return part


def revcomp(s: bytes) -> bytes:
    out = C.create_string_buffer(len(s) + 1)
    lib().gp_revcomp(s, len(s), out)
    return out.raw[:len(s)]


def is_score_significant(t: Thresholds, score, len1, len2, row_end, col_end, nclip) -> int:
    return lib().gp_is_score_significant(C.byref(t), score, len1, len2, row_end, col_end, nclip)


def _as_result(r) -> Result:
    return Result(int(r["score"]), int(r["row_end"]), int(r["col_end"]), int(r["nclip"]), int(r["flags"]))


def merged_concat(s1: bytes, s2: bytes, r) -> bytes:
    res = _as_result(r)
    out = C.create_string_buffer(len(s1) + len(s2) + 1)
    n = lib().gp_merged_concat(s1, len(s1), s2, len(s2), C.byref(res), out)
    return out.raw[:n]


def is_containment(len1: int, len2: int, r) -> bool:
    res = _as_result(r)
    return bool(lib().gp_is_containment(len1, len2, C.byref(res)))


def overlap_size(len1: int, len2: int, r) -> int:
    res = _as_result(r)
    return int(lib().gp_overlap_size(len1, len2, C.byref(res)))


class HostBatch:
    """Host-side arguments of gp_overlap_batch in the C ABI's own shape: `const char* const*` sequence pointers
    (the bytes objects stay referenced here), lengths, gp_pair array, and the gp_result array the call fills.
    Building it is Python marshalling, not part of the library."""

    def __init__(self, seqs: Sequence[bytes], pairs):
        self.seqs = list(seqs)
        self.arr, self.lens = _seq_arrays(self.seqs)
        self.n_seq = len(self.seqs)
        p = np.zeros(len(pairs), dtype=PAIR_DTYPE)
        if len(pairs):
            pa = np.asarray(pairs)
            if pa.dtype == PAIR_DTYPE:
                p = np.ascontiguousarray(pa)
            else:
                p["row_seq"] = pa[:, 0]
                p["col_seq"] = pa[:, 1]
        self.pairs = p
        self.out = np.zeros(len(p), dtype=RESULT_DTYPE)


class Context:
    """One gp_ctx (one GPU, one stream)."""

    def __init__(self, device: int = 0):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.gp_create(device, C.byref(h))
        if rc != 0:
            raise GpError(rc, (self._L.gp_last_error(None) or b"").decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.gp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise GpError(rc, (self._L.gp_last_error(self._h) or b"").decode())

    @property
    def stream(self) -> int:
        return int(self._L.gp_stream(self._h) or 0)

    @property
    def kernel_launches(self) -> int:
        return int(self._L.gp_kernel_launches(self._h))

    def pair_stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._L.gp_pair_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(cells=a.value, pairs16=b.value, pairs32=c.value)

    def pair_split(self):
        """Pairs of the uploaded batch per kernel: table 16-bit, PRMT 16-bit, general 32-bit."""
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._L.gp_pair_split(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(table16=a.value, prmt16=b.value, wide32=c.value)

    def cert_stats(self):
        """Certificate kernel: pairs routed to it, and of the last fetched run the second passes and exact retries."""
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._L.gp_cert_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(cert16=a.value, second_passes=b.value, exact_retries=c.value)

    def kernel_times(self):
        """Device milliseconds of the last launch and host-routed DP cells per kernel: cert16, table16, prmt16, wide32."""
        ms, cells = (C.c_double * 4)(), (C.c_uint64 * 4)()
        self._check(self._L.gp_kernel_times(self._h, ms, cells))
        names = ("cert16", "table16", "prmt16", "wide32")
        return {k: dict(ms=ms[i], cells=int(cells[i])) for i, k in enumerate(names)}

    def set_cert_system(self, system: int):
        """Tests: 0 probe decides (default), 1 start with system U (column 0), 2 with system L (row 0)."""
        self._check(self._L.gp_set_cert_system(self._h, system))

    def closed_form_stats(self):
        """Pairs of the uploaded batch answered in closed form (a sequence against itself) and their m*n."""
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self._L.gp_closed_form_stats(self._h, C.byref(a), C.byref(b)))
        return dict(pairs=a.value, cells=b.value)

    def quick_check_matrix(self, gap_first, k: int = 10):
        """gp_quick_check_matrix with full_matrix = 1 -> list of (n_g, n_g) uint8 arrays: m[i, j] = 1 iff the ends of node j
        occur in node i (every ordered pair)."""
        gf = np.ascontiguousarray(np.asarray(gap_first, dtype=np.uint32))
        n = (gf[1:] - gf[:-1]).astype(np.int64)
        hit = np.zeros(max(1, int((n * n).sum())), dtype=np.uint8)
        self._L.gp_quick_check_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p, C.c_uint64, C.c_int]
        self._check(self._L.gp_quick_check_matrix(self._h, gf.ctypes.data, len(gf) - 1, k, hit.ctypes.data, hit.nbytes, 1))
        out, pos = [], 0
        for ng in n:
            out.append(hit[pos:pos + ng * ng].reshape(ng, ng).copy())
            pos += ng * ng
        return out

    def quick_check_device(self, gap_first, k: int = 10):
        """Candidate filter on the device for the gaps [gap_first[g], gap_first[g+1]) of the current sequence table.
        -> list of PAIR_DTYPE arrays (node indices local to the gap), in gp_candidate_pairs' order."""
        gf = np.ascontiguousarray(np.asarray(gap_first, dtype=np.uint32))
        n = (gf[1:] - gf[:-1]).astype(np.int64)
        hit = np.zeros(int((n * n).sum()), dtype=np.uint8)
        self._check(self._L.gp_quick_check_device(self._h, gf.ctypes.data, len(gf) - 1, k, hit.ctypes.data, hit.nbytes))
        out, pos = [], 0
        for ng in n:
            m = hit[pos:pos + ng * ng].reshape(ng, ng)
            pos += ng * ng
            i, j = np.nonzero(m)                       # row-major: i ascending, then j
            p = np.zeros(len(i), dtype=PAIR_DTYPE)
            p["row_seq"], p["col_seq"] = i, j
            out.append(p)
        return out

    # ---- flank placement (semi-global; parity unpinned: bit-exact against oracle gpo_semiglobal) ----
    def semiglobal_batch(self, seqs: Sequence[bytes], pairs, params: DpParams = GAPPADDER_DP) -> np.ndarray:
        """pairs: (flank index, contig index).  -> PLACE_DTYPE array (score, col_start, col_end, flags)."""
        hb = HostBatch(seqs, pairs)
        out = np.zeros(len(hb.pairs), dtype=PLACE_DTYPE)
        self._check(self._L.gp_semiglobal_batch(self._h, hb.arr, hb.lens.ctypes.data, hb.n_seq, hb.pairs.ctypes.data, len(hb.pairs),
                                                C.byref(params), out.ctypes.data))
        return out

    def semiglobal_host_batch(self, hb: "HostBatch", out: np.ndarray, params: DpParams = GAPPADDER_DP) -> np.ndarray:
        self._check(self._L.gp_semiglobal_batch(self._h, hb.arr, hb.lens.ctypes.data, hb.n_seq, hb.pairs.ctypes.data, len(hb.pairs),
                                                C.byref(params), out.ctypes.data))
        return out

    def semiglobal_upload_pairs(self, pairs: np.ndarray, params: DpParams = GAPPADDER_DP):
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        self._check(self._L.gp_semiglobal_upload_pairs(self._h, pairs.ctypes.data, len(pairs), C.byref(params)))
        self._n_place = len(pairs)

    def semiglobal_launch(self):
        self._check(self._L.gp_semiglobal_launch(self._h))

    def semiglobal_fetch(self) -> np.ndarray:
        out = np.zeros(self._n_place, dtype=PLACE_DTYPE)
        self._check(self._L.gp_semiglobal_fetch(self._h, out.ctypes.data, self._n_place))
        return out

    def semiglobal_stats(self) -> dict:
        cells, t, g, ms = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_double(0)
        self._check(self._L.gp_semiglobal_stats(self._h, C.byref(cells), C.byref(t), C.byref(g), C.byref(ms)))
        return dict(cells=cells.value, table_pairs=t.value, generic_pairs=g.value, kernel_ms=ms.value)

    # ---- TERefiner's affine local aligner (parity pinned: the reference's own local_alignment.cpp) ----
    def local_affine_batch(self, seqs: Sequence[bytes], pairs, params: AffineParams = TEREFINER_AFFINE) -> np.ndarray:
        """pairs: (sref index, ssgmt index).  -> LOCAL_DTYPE array (score, start1, end1, start2, end2, flags), 1-based."""
        hb = HostBatch(seqs, pairs)
        out = np.zeros(len(hb.pairs), dtype=LOCAL_DTYPE)
        self._check(self._L.gp_local_affine_batch(self._h, hb.arr, hb.lens.ctypes.data, hb.n_seq, hb.pairs.ctypes.data, len(hb.pairs),
                                                  C.byref(params), out.ctypes.data))
        return out

    def local_affine_host_batch(self, hb: "HostBatch", out: np.ndarray, params: AffineParams = TEREFINER_AFFINE) -> np.ndarray:
        self._check(self._L.gp_local_affine_batch(self._h, hb.arr, hb.lens.ctypes.data, hb.n_seq, hb.pairs.ctypes.data, len(hb.pairs),
                                                  C.byref(params), out.ctypes.data))
        return out

    def local_affine_upload_pairs(self, pairs: np.ndarray, params: AffineParams = TEREFINER_AFFINE):
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        self._check(self._L.gp_local_affine_upload_pairs(self._h, pairs.ctypes.data, len(pairs), C.byref(params)))
        self._n_local = len(pairs)

    def local_affine_launch(self):
        self._check(self._L.gp_local_affine_launch(self._h))

    def local_affine_fetch(self) -> np.ndarray:
        out = np.zeros(self._n_local, dtype=LOCAL_DTYPE)
        self._check(self._L.gp_local_affine_fetch(self._h, out.ctypes.data, self._n_local))
        return out

    def local_affine_stats(self) -> dict:
        cells, f, e = C.c_uint64(0), C.c_double(0), C.c_double(0)
        self._check(self._L.gp_local_affine_stats(self._h, C.byref(cells), C.byref(f), C.byref(e)))
        return dict(cells=cells.value, forward_ms=f.value, epilogue_ms=e.value)

    def quick_check_stats(self) -> dict:
        """Of the last quick_check_device: kernel ms (CUDA events), bases scanned, work items."""
        ms, bases, items = C.c_double(0), C.c_uint64(0), C.c_uint32(0)
        self._check(self._L.gp_quick_check_stats(self._h, C.byref(ms), C.byref(bases), C.byref(items)))
        return dict(kernel_ms=ms.value, bases=bases.value, items=items.value)

    def set_cert_layout(self, mode: int):
        """Certificate kernel: 0 free-moves layout whenever a launch allows it (default), 1 always the column potential."""
        self._check(self._L.gp_set_cert_layout(self._h, mode))

    @property
    def last_layout(self) -> int:
        return int(self._L.gp_last_layout(self._h))

    def set_orientation(self, mode: int):
        """Certificate kernel: 0 the longer sequence as rows (default), 1 never transpose, 2 always when allowed."""
        self._check(self._L.gp_set_orientation(self._h, mode))

    @property
    def transposed_pairs(self) -> int:
        return int(self._L.gp_transposed_pairs(self._h))

    def set_team_mode(self, mode: int):
        """Certificate kernel: 0 library chooses per launch (default), 1 one warp per pair, 2 one CTA of four warps per pair,
        3 one CTA of eight warps per pair."""
        self._check(self._L.gp_set_team_mode(self._h, mode))

    @property
    def last_team(self) -> bool:
        return bool(self._L.gp_last_team(self._h))

    def set_kernel_mask(self, mask: int):
        """Restricts the 16-bit kernels gp_upload_pairs may pick (tests, A/B timing); results never change."""
        self._check(self._L.gp_set_kernel_mask(self._h, mask))

    def last_timing(self):
        """Host wall-clock breakdown (ms) of the last overlap_batch: pack, prepare pairs, device, total."""
        a = (C.c_double * 4)()
        self._check(self._L.gp_last_timing(self._h, a, 4))
        return dict(pack_ms=a[0], prepare_ms=a[1], device_ms=a[2], total_ms=a[3])

    def int_peak(self):
        """-> (ALU-pipe, dual-pipe) thread-level packed-16x2 instructions per second, measured now."""
        a, d = C.c_double(), C.c_double()
        self._check(self._L.gp_int_peak(self._h, C.byref(a), C.byref(d)))
        return a.value, d.value

    def set_sequences(self, packed: np.ndarray, off: np.ndarray, lens: np.ndarray, n_symbols: int):
        self._check(self._L.gp_set_sequences(self._h, packed.ctypes.data, packed.nbytes, off.ctypes.data,
                                             lens.ctypes.data, len(lens), n_symbols))

    def upload_host_sequences(self, batch: "HostBatch"):
        """gp_upload_sequences: ASCII sequences -> 4-bit codes in the context's pinned buffer -> HBM.  Blocking."""
        self._check(self._L.gp_upload_sequences(self._h, batch.arr, batch.lens.ctypes.data, batch.n_seq))

    def upload_pairs(self, pairs: np.ndarray, params: DpParams = GAPPADDER_DP):
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        self._check(self._L.gp_upload_pairs(self._h, pairs.ctypes.data, len(pairs), C.byref(params)))
        self._n_pairs = len(pairs)

    def launch_resident(self):
        self._check(self._L.gp_launch_resident(self._h))

    def fetch_results(self) -> np.ndarray:
        out = np.zeros(self._n_pairs, dtype=RESULT_DTYPE)
        self._check(self._L.gp_fetch_results(self._h, out.ctypes.data, self._n_pairs))
        return out

    def overlap_pairs(self, pairs: np.ndarray, params: DpParams = GAPPADDER_DP) -> np.ndarray:
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        out = np.zeros(len(pairs), dtype=RESULT_DTYPE)
        self._check(self._L.gp_overlap_pairs(self._h, pairs.ctypes.data, len(pairs), C.byref(params), out.ctypes.data))
        return out

    def overlap_batch(self, seqs: Sequence[bytes], pairs, params: DpParams = GAPPADDER_DP) -> np.ndarray:
        """The call a user makes: host ASCII sequences + (row, col) index pairs -> results."""
        return self.overlap_host_batch(HostBatch(seqs, pairs), params)

    def overlap_host_batch(self, batch: "HostBatch", params: DpParams = GAPPADDER_DP) -> np.ndarray:
        """gp_overlap_batch on host buffers that already have the C ABI's shape (what a C++ caller holds: an
        array of sequence pointers, their lengths, the pair list).  Everything device-side -- packing into
        pinned memory, the copies in both directions, the kernels -- happens inside the call."""
        out = batch.out
        self._check(self._L.gp_overlap_batch(self._h, batch.arr, batch.lens.ctypes.data, batch.n_seq, batch.pairs.ctypes.data,
                                             len(batch.pairs), C.byref(params), out.ctypes.data))
        return out
