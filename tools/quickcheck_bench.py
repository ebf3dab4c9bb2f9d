#!/usr/bin/env python
"""tools/quickcheck_bench.py -- the candidate filter (quick check) of N synthetic gaps: gp_candidate_pairs on the host
(one call per gap, as the drop-in does) next to gp_quick_check_device on the packed table already in HBM.  Prints one
JSON line; the two pair lists are compared."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gappadder_b200 as g  # noqa: E402
import synth_gaps  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaps", type=int, default=200)
    ap.add_argument("--config", default="cfg1")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    seqs, gap_first, per_gap = [], [0], []
    for gi in range(a.gaps):
        nodes = []
        for _, s in synth_gaps.make_gap(1 + gi, synth_gaps.CONFIGS[a.config]):
            nodes.append(s)
            nodes.append(g.revcomp(s))
        per_gap.append(nodes)
        seqs += nodes
        gap_first.append(len(seqs))
    bases = sum(len(s) for s in seqs)
    t0 = time.perf_counter()
    want = [g.candidate_pairs(nodes, 10) for nodes in per_gap]
    host_s = time.perf_counter() - t0
    packed, off, lens, nsym = g.pack_sequences(seqs)
    ctx = g.Context(0)
    ctx.set_sequences(packed, off, lens, nsym)
    got = ctx.quick_check_device(gap_first, 10)                  # warm-up + result
    same = all(np.array_equal(x["row_seq"], y["row_seq"]) and np.array_equal(x["col_seq"], y["col_seq"]) for x, y in zip(got, want))
    gf = np.asarray(gap_first, dtype=np.uint32)
    n = (gf[1:] - gf[:-1]).astype(np.int64)
    hit = np.zeros(int((n * n).sum()), dtype=np.uint8)
    ts = []
    for _ in range(a.reps):                                       # the blocking C call: meta H2D, kernel, hit matrices D2H
        t0 = time.perf_counter()
        ctx._check(ctx._L.gp_quick_check_device(ctx._h, gf.ctypes.data, len(gf) - 1, 10, hit.ctypes.data, hit.nbytes))
        ts.append(time.perf_counter() - t0)
    dev_s = float(np.median(ts))
    st = ctx.quick_check_stats()                                  # kernel alone (CUDA events), last call
    kernel_bytes = bases // 2 + bases // 8                        # packed codes once + 4 B carry-in per 32 bases
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6545.0
    algo_bytes = packed.nbytes + hit.nbytes                       # packed codes read once (0.5 B/base) + hit matrices written
    print(json.dumps({"config": a.config, "gaps": a.gaps, "nodes": len(seqs), "mbases": bases / 1e6, "pairs": int(sum(len(x) for x in want)),
                      "identical": bool(same), "host_ms": host_s * 1e3, "host_note": "gp_candidate_pairs, one call per gap, one thread",
                      "device_ms": dev_s * 1e3, "device_note": "blocking gp_quick_check_device call, sequences resident in HBM",
                      "device_gbases_per_s": bases / dev_s / 1e9, "algorithmic_bytes": int(algo_bytes),
                      "device_algorithmic_gb_per_s": algo_bytes / dev_s / 1e9,
                      "kernel_ms": st["kernel_ms"], "kernel_items": st["items"], "kernel_bytes": int(kernel_bytes),
                      "kernel_gb_per_s": kernel_bytes / (st["kernel_ms"] * 1e-3) / 1e9, "hbm_peak_gb_per_s": peak,
                      "kernel_frac_of_hbm_peak": kernel_bytes / (st["kernel_ms"] * 1e-3) / 1e9 / peak}))


if __name__ == "__main__":
    main()
