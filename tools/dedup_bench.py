#!/usr/bin/env python
"""tools/dedup_bench.py -- the dedup stage timed whole: build/ContigsMerger_b200 --dedup-batch on the bench
workload's gaps as contig sets (cfg1: 200 sets x 40 contigs plus planted contained / duplicate contigs), both rules
(`g`: contained, `p`: duplicates).  What the reference spends here is six process spawns per set, a BWA index and a BWA
self-alignment (MergeContigs.py:15-70); none of them exists on the GPU box (BWA and samtools are not installed), so there is
no reference arm -- the line reports sets per second, pairs, DP cells and the kernel-phase share.  Correctness of the stage is
tests/test_gpu_dedup.py (oracle/dedup_oracle.py; rules pinned to TERefiner_1, alignment records builder-defined)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _dedupcases import FLAGS, make_set  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", type=int, default=200)
    ap.add_argument("--config", default="cfg1")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    binary = os.path.join(ROOT, "build", "ContigsMerger_b200")
    out = {"config": args.config, "sets": args.sets, "gpus": args.gpus,
           "parity": "rules pinned to TERefiner_1 (tests/golden/dedup); alignment records builder-defined, BWA parity unpinned (SURVEY.md 8c)"}
    with tempfile.TemporaryDirectory() as td:
        for k in range(args.sets):
            open(os.path.join(td, "s%d.fa" % k), "wb").write(make_set(1 + k, args.config))
        for mode in ("g", "p"):
            lst = os.path.join(td, "list_%s.tsv" % mode)
            with open(lst, "w") as f:
                for k in range(args.sets):
                    f.write("%s\t%s\t%s\t%s\n" % (os.path.join(td, "s%d.fa" % k), os.path.join(td, "s%d.%s.out" % (k, mode)), "0.99" if mode == "g" else "0.95", mode))
            best = None
            for _ in range(args.repeat):
                t0 = time.perf_counter()
                p = subprocess.run([binary] + FLAGS + ["--dedup-batch", lst, "--gpus", str(args.gpus), "--stats"], capture_output=True, text=True)
                wall = time.perf_counter() - t0
                if p.returncode != 0:
                    out[mode] = {"error": p.stderr[-300:]}
                    break
                st = json.loads(p.stderr.strip().splitlines()[-1])
                st["process_wall_s"] = wall
                if best is None or st["dedup_ms"] < best["dedup_ms"]:
                    best = st
            if best:
                best["sets_per_s"] = args.sets / (best["dedup_ms"] * 1e-3)
                best["gcups"] = best["dp_gcells"] / (best["dedup_ms"] * 1e-3)
                out[mode] = best
    print(json.dumps(out))


if __name__ == "__main__":
    main()
