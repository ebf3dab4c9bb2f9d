#!/usr/bin/env python
"""tools/lone_pair_bench.py -- how fast ONE long pair runs (the situation at the tail of a relax launch: a handful of long
pairs, the rest of the chip idle): rows x columns related sequences through gp_overlap_batch with one warp, a 4-warp team
and an 8-warp team; device time of the certificate kernel, clocks per step per warp (a step = one column of one 512-row
strip), cells per microsecond.  Also with 148 identical pairs (one per SM) to see what sharing the chip costs."""
import json
import random
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gappadder_b200 as g  # noqa: E402


def rand(rng, n):
    return bytes(rng.choice(b"ACGT") for _ in range(n))


def trace_one(m, n, mode):
    """Diagnostic build (make trace; GAPPADDER_B200_LIB=build/libgappadder_b200_trace.so): phase stamps of one pair."""
    import ctypes
    rng = random.Random(7)
    col = rand(rng, n)
    row = rand(rng, m - n // 2) + col[:n // 2]
    L = g.lib()
    with g.Context(0) as ctx:
        ctx.set_team_mode(mode)
        ctx.overlap_batch([row, col], [(0, 1)])
        L.gp_debug_trace_dump(b"/dev/null")
        ctx.overlap_batch([row, col], [(0, 1)])
        path = os.path.join(ROOT, "gpurun_out", "wf16c_trace_%d_%d_mode%d.txt" % (m, n, mode))
        k = L.gp_debug_trace_dump(path.encode())
        ctx.set_team_mode(0)
    rows = [tuple(int(x) for x in ln.split()) for ln in open(path)]
    rows.sort()
    t0 = rows[0][0]
    names = {1: "pair", 2: "probed", 3: "line", 4: "strips_done", 5: "pass_done", 10: "strip", 11: "table", 12: "waited", 13: "block", 14: "blocks_done", 15: "strip_end"}
    print("trace %d x %d mode %d: %d stamps" % (m, n, mode, k), file=sys.stderr)
    for t, blk, warp, tag, val in rows:
        print("  %9.2f us  warp %d  %-12s %d" % ((t - t0) / 1e3, warp, names.get(tag, str(tag)), val), file=sys.stderr)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--trace":
        trace_one(12000, 2000, 2)
        trace_one(12000, 2000, 1)
        return
    rng = random.Random(7)
    out = []
    with g.Context(0) as ctx:
        for m, n in ((12000, 2000), (12000, 500), (12000, 3700), (2400, 1700)):
            col = rand(rng, n)
            row = rand(rng, m - n // 2) + col[:n // 2]             # the column sequence's first half ends the row sequence
            for copies in (1,):
                seqs, pairs = [], []
                for c in range(copies):
                    seqs += [row, col]
                    pairs.append((2 * c, 2 * c + 1))
                for mode, name, warps in ((1, "warp", 1), (2, "team4", 4), (3, "team8", 8)):
                    ctx.set_team_mode(mode)
                    best = None
                    for _ in range(3):
                        ctx.overlap_batch(seqs, pairs)
                        ms = ctx.kernel_times()["cert16"]["ms"]
                        sp = ctx.cert_stats()
                        best = ms if best is None else min(best, ms)
                    strips = (m + 511) // 512
                    steps_per_warp = ((strips + warps - 1) // warps) * (n + 94)
                    out.append({"rows": m, "cols": n, "pairs": copies, "mode": name, "ms": best, "cert_stats": sp, "cells_per_us": m * n * 1e-3 / best,
                                "clk_per_step_per_warp_at_1965MHz": best * 1e-3 * 1.965e9 / steps_per_warp})
        ctx.set_team_mode(0)
    print(json.dumps(out))
    for r in out:
        print(r, file=sys.stderr)


if __name__ == "__main__":
    main()
