"""ctypes bindings to the checker libraries under oracle/ (tests and bench cpu_baseline only).

  oracle/_build/liboverlap_oracle.so  C restatement (oracle/overlap_oracle.c), built on demand with gcc
  oracle/_ref/libcm_ref.so            the reference's own Evaluate (oracle/build_ref.sh); optional
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

# GAPPadder's flags (MergeContigs.py:85) as main.cpp sees them: -s/-x/-y go through float.
import numpy as _np
GAPPADDER_PARAMS = dict(
    fractionLossScore=float(_np.float32(0.4)), fracMinOverlap=0.005, minOverlapLen=12.0,
    maxOverlapClipLen=50.0, minOverlapLenWithScaffold=6.0, scoreMismatch=-2.0, scoreIndel=-2.0)


class DPResult(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("score", "row_end", "col_end", "nclip", "tb_row", "tb_col", "bcontained")]

    def key(self):
        return (self.score, self.row_end, self.col_end, self.nclip, self.bcontained)


class Thresholds(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("fractionLossScore", "fracMinOverlap", "minOverlapLen", "minOverlapLenWithScaffold")]


def gappadder_thresholds() -> Thresholds:
    p = GAPPADDER_PARAMS
    return Thresholds(p["fractionLossScore"], p["fracMinOverlap"], p["minOverlapLen"], p["minOverlapLenWithScaffold"])


_oracle = None


class PlaceResult(C.Structure):
    _fields_ = [("score", C.c_int32), ("col_start", C.c_int32), ("col_end", C.c_int32)]


def oracle_lib() -> C.CDLL:
    global _oracle
    if _oracle is None:
        so = os.path.join(ORACLE_DIR, "_build", "liboverlap_oracle.so")
        src = os.path.join(ORACLE_DIR, "overlap_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "all"])
        lib = C.CDLL(so)
        for fn in ("gpo_evaluate", "gpo_evaluate_full"):
            getattr(lib, fn).argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(DPResult)]
            getattr(lib, fn).restype = C.c_int
        lib.gpo_is_score_significant.argtypes = [C.POINTER(Thresholds)] + [C.c_int] * 6
        lib.gpo_merged_concat.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]
        lib.gpo_is_containment.argtypes = [C.c_int] * 6
        lib.gpo_overlap_size.argtypes = [C.c_int] * 4
        lib.gpo_revcomp.argtypes = [C.c_char_p, C.c_int, C.c_char_p]
        lib.gpo_revcomp.restype = None
        lib.gpo_quickcheck.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int]
        lib.gpo_candidate_pairs.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int64]
        lib.gpo_candidate_pairs.restype = C.c_int64
        lib.gpo_semiglobal.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(PlaceResult)]
        _oracle = lib
    return _oracle


def oracle_evaluate(s1: bytes, s2: bytes, mismatch=-2, indel=-2, maxclip=50, full=False) -> DPResult:
    r = DPResult()
    fn = oracle_lib().gpo_evaluate_full if full else oracle_lib().gpo_evaluate
    rc = fn(s1, len(s1), s2, len(s2), mismatch, indel, maxclip, C.byref(r))
    assert rc == 0
    return r


def oracle_semiglobal(flank: bytes, contig: bytes, mismatch=-2, indel=-2) -> PlaceResult:
    """Flank end to end inside the contig (builder-written definition, BWA parity unpinned: see overlap_oracle.c)."""
    r = PlaceResult()
    rc = oracle_lib().gpo_semiglobal(flank, len(flank), contig, len(contig), mismatch, indel, C.byref(r))
    assert rc == 0
    return r


def oracle_revcomp(s: bytes) -> bytes:
    out = C.create_string_buffer(len(s) + 1)
    oracle_lib().gpo_revcomp(s, len(s), out)
    return out.value


def oracle_merged(s1: bytes, s2: bytes, r: DPResult) -> bytes:
    out = C.create_string_buffer(len(s1) + len(s2) + 1)
    n = oracle_lib().gpo_merged_concat(s1, len(s1), s2, len(s2), r.row_end, r.col_end, r.nclip, r.bcontained, out)
    return out.raw[:n]


def oracle_candidate_pairs(nodes, k=10):
    n = len(nodes)
    arr = (C.c_char_p * n)(*nodes)
    lens = (C.c_int32 * n)(*[len(x) for x in nodes])
    cap = n * (n + 1) // 2
    buf = (C.c_int32 * (2 * cap))()
    cnt = oracle_lib().gpo_candidate_pairs(arr, lens, n, k, buf, cap)
    return [(buf[2 * i], buf[2 * i + 1]) for i in range(cnt)]


_ref = None
_ref_tried = False


def ref_lib() -> Optional[C.CDLL]:
    """The reference's own code (oracle/_ref/libcm_ref.so) or None when it has not been built."""
    global _ref, _ref_tried
    if not _ref_tried:
        _ref_tried = True
        so = os.path.join(ORACLE_DIR, "_ref", "libcm_ref.so")
        if os.path.exists(so):
            lib = C.CDLL(so)
            lib.cmref_set_params.argtypes = [C.c_double] * 7
            lib.cmref_set_params.restype = None
            lib.cmref_evaluate.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int32)]
            lib.cmref_is_score_significant.argtypes = [C.c_int] * 6
            lib.cmref_last_merged.argtypes = [C.c_char_p, C.c_int64]
            lib.cmref_last_merged.restype = C.c_int64
            lib.cmref_quickcheck.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
            lib.cmref_revcomp.argtypes = [C.c_char_p, C.c_char_p]
            lib.cmref_revcomp.restype = None
            ref_set_params(lib)
            _ref = lib
    return _ref


def ref_set_params(lib=None, **kw):
    p = dict(GAPPADDER_PARAMS)
    p.update(kw)
    (lib or ref_lib()).cmref_set_params(p["fractionLossScore"], p["fracMinOverlap"], p["minOverlapLen"],
                                        p["maxOverlapClipLen"], p["minOverlapLenWithScaffold"],
                                        p["scoreMismatch"], p["scoreIndel"])


def ref_evaluate(s1: bytes, s2: bytes, relax: bool):
    """-> dict(res, score, row_end, col_end, nclip, bcontained, is_containment, overlap, merged)"""
    out = (C.c_int32 * 8)()
    ref_lib().cmref_evaluate(s1, s2, 1 if relax else 0, out)
    d = dict(zip(("res", "score", "row_end", "col_end", "nclip", "bcontained", "is_containment", "overlap"), list(out)))
    if d["bcontained"] >= 0:
        buf = C.create_string_buffer(len(s1) + len(s2) + 2)
        n = ref_lib().cmref_last_merged(buf, len(buf))
        d["merged"] = buf.raw[:n]
    else:
        d["merged"] = None
    return d


def ref_binary() -> Optional[str]:
    p = os.path.join(ORACLE_DIR, "_ref", "ContigsMerger")
    return p if os.path.exists(p) else None


# ---- TERefiner's affine local aligner: the reference's own code (oracle/_ref/libla_ref.so, oracle/la_harness.cpp) ----
_la_ref = None
_la_tried = False


def la_ref_lib() -> Optional[C.CDLL]:
    global _la_ref, _la_tried
    if not _la_tried:
        _la_tried = True
        so = os.path.join(ORACLE_DIR, "_ref", "libla_ref.so")
        if os.path.exists(so):
            lib = C.CDLL(so)
            lib.laref_opt_align.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int32)]
            lib.laref_opt_align.restype = None
            lib.laref_stdaln_local.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int32)]
            lib.laref_stdaln_local.restype = None
            lib.laref_forward_score.argtypes = [C.c_char_p, C.c_char_p]
            lib.laref_forward_score.restype = C.c_int32
            _la_ref = lib
    return _la_ref


def ref_local_affine(s1: bytes, s2: bytes):
    """(score, start1, end1, start2, end2) of the reference's aln_stdaln(s1, s2, &aln_param_blast, LOCAL, 1), or None when
    nothing aligns (the reference reads path[-1] there)."""
    lib = la_ref_lib()
    if not s1 or not s2 or lib.laref_forward_score(s1, s2) < 1:
        return None
    o = (C.c_int32 * 6)()
    lib.laref_stdaln_local(s1, s2, o)
    return (o[0], o[1], o[2], o[3], o[4])


_la_oracle = None


def la_oracle_lib() -> C.CDLL:
    """oracle/local_affine_oracle.c (the C restatement of the affine local aligner), built on demand with gcc."""
    global _la_oracle
    if _la_oracle is None:
        so = os.path.join(ORACLE_DIR, "_build", "liblocal_affine_oracle.so")
        src = os.path.join(ORACLE_DIR, "local_affine_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "all"])
        lib = C.CDLL(so)
        lib.lao_local_affine.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int32)]
        _la_oracle = lib
    return _la_oracle


def oracle_local_affine(s1: bytes, s2: bytes):
    """(score, start1, end1, start2, end2) by the oracle restatement, or None when nothing aligns."""
    o = (C.c_int32 * 5)()
    rc = la_oracle_lib().lao_local_affine(s1, len(s1), s2, len(s2), o)
    assert rc in (0, 1), "reverse band collapsed (the reference is undefined there)"
    return None if rc == 1 else tuple(o)
