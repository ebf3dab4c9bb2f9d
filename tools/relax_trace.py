#!/usr/bin/env python
"""tools/relax_trace.py -- where a relax launch's time goes: runs the drop-in on synthetic gaps with GP_RELAX_TRACE set (every
item of gp_relax_chains records its begin / end time on the device) and prints the critical chain -- the item that ends last,
followed back through its parents -- step by step: rows, columns, duration, idle time between the parent's end and the
child's begin, and the cells per microsecond each step achieved."""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth_gaps  # noqa: E402

FLAGS = "-s 0.4 -i1 -2.0 -i2 -2.0 -x 12 -y 50 -k 10 -m 1 -t 5".split()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaps", type=int, default=200)
    ap.add_argument("--config", default="cfg1")
    args = ap.parse_args()
    with tempfile.TemporaryDirectory() as td:
        lst = os.path.join(td, "list.tsv")
        with open(lst, "w") as f:
            for g in range(args.gaps):
                fa = os.path.join(td, "g%d.fa" % g)
                synth_gaps.write_fasta(fa, synth_gaps.make_gap(1 + g, synth_gaps.CONFIGS[args.config]))
                f.write("%s\t%s\t%s\n" % (fa, os.path.join(td, "g%d.out" % g), os.path.join(td, "g%d.info" % g)))
        trace = os.path.join(td, "trace.txt")
        p = subprocess.run([os.path.join(ROOT, "build", "ContigsMerger_b200")] + FLAGS + ["--batch", lst, "--no-gml", "--stats", "--chunk-gaps", str(args.gaps)],
                           capture_output=True, text=True, env=dict(os.environ, GP_RELAX_TRACE=trace))
        if p.returncode != 0:
            print(json.dumps({"error": p.stderr[-400:]}))
            return 1
        items = {}
        for ln in open(trace):
            k, parent, t0, t1, m, n = (int(x) for x in ln.split())
            if t1:
                items[k] = (parent, t0, t1, m, n)
    start = min(v[1] for v in items.values())
    end_item = max(items, key=lambda k: items[k][2])
    chain = []
    k = end_item
    while k >= 0 and k in items:
        chain.append(k)
        k = items[k][0]
    chain.reverse()
    steps = []
    prev_end = None
    for k in chain:
        parent, t0, t1, m, n = items[k]
        steps.append({"item": k, "rows": m, "cols": n, "begin_us": (t0 - start) / 1e3, "dur_us": (t1 - t0) / 1e3,
                      "idle_before_us": None if prev_end is None else (t0 - prev_end) / 1e3, "mcells": m * n / 1e6,
                      "cells_per_us": m * n / max(1.0, (t1 - t0) / 1e3)})
        prev_end = t1
    total = (items[end_item][2] - start) / 1e3
    busy = sum(s["dur_us"] for s in steps)
    # how many items run at a time over the launch (20 slices)
    slices = 20
    conc = [0] * slices
    span = max(v[2] for v in items.values()) - start
    for parent, t0, t1, m, n in items.values():
        a, b = int((t0 - start) * slices / span), min(slices - 1, int((t1 - start) * slices / span))
        for q in range(a, b + 1):
            conc[q] += 1
    print(json.dumps({"items": len(items), "launch_us": total, "critical_chain_steps": len(steps), "critical_chain_busy_us": busy,
                      "critical_chain_first_begin_us": steps[0]["begin_us"], "items_active_per_slice": conc, "steps": steps}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
