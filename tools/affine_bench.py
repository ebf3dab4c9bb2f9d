#!/usr/bin/env python3
"""Device-resident timing of TERefiner's affine local aligner (gp_local_affine_*) on cfg1-shaped contig pairs:
forward kernel and start-recovery kernel, CUDA events inside the library (gp_local_affine_stats).
    python tools/affine_bench.py [--gaps 20] [--reps 3] [--check 64]
--check N compares N sampled pairs with the reference's own aligner when oracle/_ref/libla_ref.so is present."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")]
import gappadder_b200 as g   # noqa: E402
import synth_gaps            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaps", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", type=int, default=64)
    a = ap.parse_args()
    seqs, pairs = [], []
    for seed in range(1, a.gaps + 1):
        nodes = []
        for _, s in synth_gaps.make_gap(seed, synth_gaps.CONFIGS["cfg1"]):
            nodes += [s, g.revcomp(s)]
        base = len(seqs)
        seqs += nodes
        cp = g.candidate_pairs(nodes, 10)
        pairs += [(base + int(x), base + int(y)) for x, y in cp if x != y]
    ctx = g.Context(0)
    hb = g.capi.HostBatch(seqs, pairs)
    t0 = time.perf_counter()
    ctx.upload_host_sequences(hb)
    ctx.local_affine_upload_pairs(hb.pairs)
    t1 = time.perf_counter()
    runs = []
    for _ in range(a.reps):
        ctx.local_affine_launch()
        res = ctx.local_affine_fetch()
        runs.append(ctx.local_affine_stats())
    best = min(runs, key=lambda r: r["forward_ms"] + r["epilogue_ms"])
    cells = best["cells"]
    out = dict(workload="cfg1 contig pairs (quick-check candidates), TERefiner affine local aligner", gaps=a.gaps, pairs=len(pairs), cells=cells,
               upload_ms=(t1 - t0) * 1e3, forward_ms=best["forward_ms"], epilogue_ms=best["epilogue_ms"],
               forward_gcups=cells / best["forward_ms"] / 1e6, total_gcups=cells / (best["forward_ms"] + best["epilogue_ms"]) / 1e6,
               mean_score=float(res["score"].mean()), flagged=int((res["flags"] != 0).sum()), runs=runs)
    if a.check:
        import _oracle
        if _oracle.la_ref_lib() is not None:
            idx = np.random.default_rng(1).choice(len(pairs), size=min(a.check, len(pairs)), replace=False)
            t2 = time.perf_counter()
            same = 0
            ref_cells = 0
            for k in idx:
                w = _oracle.ref_local_affine(seqs[pairs[k][0]], seqs[pairs[k][1]])
                r = res[k]
                ref_cells += len(seqs[pairs[k][0]]) * len(seqs[pairs[k][1]])
                same += (w is None and int(r["flags"]) & 1) or (w == (int(r["score"]), int(r["start1"]), int(r["end1"]), int(r["start2"]), int(r["end2"])))
            dt = time.perf_counter() - t2
            out["parity_sample"] = dict(pairs=int(len(idx)), identical=int(same), reference_gcups_1core=ref_cells / dt / 1e9)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
