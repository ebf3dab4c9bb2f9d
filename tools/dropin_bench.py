#!/usr/bin/env python
"""tools/dropin_bench.py -- whole-ContigsMerger timing: build/ContigsMerger_b200 --batch on N synthetic
cfg1 gaps (FASTA in, merged FASTA + info out, every phase: quick check, pairwise DP, graph, relax
chains, output) next to the reference binary (oracle/_ref/ContigsMerger, -t <cores>) on a few of the same
gaps, with byte comparison of the outputs of those gaps.  Prints one JSON line."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth_gaps  # noqa: E402

FLAGS = "-s 0.4 -i1 -2.0 -i2 -2.0 -x 12 -y 50 -k 10 -m 1".split()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaps", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--streams", type=int, default=1, help="workers (host thread + context + stream) per GPU")
    ap.add_argument("--ref-gaps", type=int, default=2, help="gaps also run through the reference binary (0: skip)")
    ap.add_argument("--config", default="cfg1")
    args = ap.parse_args()
    binary = os.path.join(ROOT, "build", "ContigsMerger_b200")
    ref = os.path.join(ROOT, "oracle", "_ref", "ContigsMerger")
    cores = len(os.sched_getaffinity(0))
    with tempfile.TemporaryDirectory() as td:
        lst = os.path.join(td, "list.tsv")
        with open(lst, "w") as f:
            for g in range(args.gaps):
                fa = os.path.join(td, "g%d.fa" % g)
                synth_gaps.write_fasta(fa, synth_gaps.make_gap(args.seed + g, synth_gaps.CONFIGS[args.config]))
                f.write("%s\t%s\t%s\n" % (fa, os.path.join(td, "g%d.out" % g), os.path.join(td, "g%d.info" % g)))
        t0 = time.perf_counter()
        p = subprocess.run([binary] + FLAGS + ["-t", "5", "--batch", lst, "--gpus", str(args.gpus), "--streams", str(args.streams), "--no-gml", "--stats"],
                           capture_output=True, text=True)
        wall = time.perf_counter() - t0
        if p.returncode != 0:
            print(json.dumps({"error": p.stderr[-500:]}))
            return 1
        stats = json.loads(p.stderr.strip().splitlines()[-1])
        line = {"impl": "b200", "config": args.config, "process_wall_s": wall, **stats,
                "gaps_per_s": args.gaps / (stats["merge_ms"] * 1e-3),
                # cells actually computed: not the closed-form pairs, not the relax steps shared between chains
                "gcups": (stats["dp_gcells"] - stats.get("closed_gcells", 0.0) - stats.get("relax_shared_gcells", 0.0)) / (stats["merge_ms"] * 1e-3)}
        if args.ref_gaps and os.path.exists(ref):
            t_ref, same = 0.0, True
            for g in range(min(args.ref_gaps, args.gaps)):
                fa = os.path.join(td, "g%d.fa" % g)
                info = os.path.join(td, "g%d.refinfo" % g)
                t0 = time.perf_counter()
                r = subprocess.run([ref] + FLAGS + ["-t", str(cores), "-o", info, fa], cwd=td, capture_output=True)
                t_ref += time.perf_counter() - t0
                same = same and r.stdout == open(os.path.join(td, "g%d.out" % g), "rb").read() \
                    and open(info, "rb").read() == open(os.path.join(td, "g%d.info" % g), "rb").read()
            n = min(args.ref_gaps, args.gaps)
            line["reference"] = {"gaps": n, "cores": cores, "seconds": t_ref, "gaps_per_s": n / t_ref, "outputs_identical": same}
        print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
