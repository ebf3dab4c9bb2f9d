#!/usr/bin/env python
"""tools/dropin_bench.py -- whole-ContigsMerger timing: build/ContigsMerger_b200 --batch [--gpus N] on synthetic gaps
(FASTA in, merged FASTA + info out, every phase: FASTA read, LPT partition over GPUs, quick check, pairwise DP, graph,
relax chains, output).  Checks, all by byte comparison of the output files:
  --ref-gaps K      K of the gaps also go through the reference binary (oracle/_ref/ContigsMerger, -t <cores>)
  --verify-gpus1 K  (with --gpus N > 1) the first K gaps are run again with --gpus 1
Prints one JSON line.  gaps_per_s is whole gaps merged per second over merge_ms = the whole batch inside the process: from the
parsed list to the last output file written (partition, FASTA reading, context creation, every phase, file writing);
process start-up and module load are reported separately as process_wall_s."""
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


def write_gaps(td, config, n_gaps, seed):
    lst = os.path.join(td, "list.tsv")
    with open(lst, "w") as f:
        for g in range(n_gaps):
            fa = os.path.join(td, "g%d.fa" % g)
            synth_gaps.write_fasta(fa, synth_gaps.make_gap(seed + g, synth_gaps.CONFIGS[config]))
            f.write("%s\t%s\t%s\n" % (fa, os.path.join(td, "g%d.out" % g), os.path.join(td, "g%d.info" % g)))
    return lst


def run_binary(binary, lst, gpus, streams, extra=(), chunk_gaps=256):
    t0 = time.perf_counter()
    p = subprocess.run([binary] + FLAGS + ["-t", "5", "--batch", lst, "--gpus", str(gpus), "--streams", str(streams), "--chunk-gaps", str(chunk_gaps),
                        "--no-gml", "--stats"] + list(extra),
                       capture_output=True, text=True)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(p.stderr[-500:])
    return wall, json.loads(p.stderr.strip().splitlines()[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaps", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--streams", type=int, default=2, help="mergers (host thread + context + stream) per GPU; they alternate on the device")
    ap.add_argument("--chunk-gaps", type=int, default=256, help="gaps per chunk of the batch pipeline")
    ap.add_argument("--ref-gaps", type=int, default=2, help="gaps also run through the reference binary (0: skip)")
    ap.add_argument("--verify-gpus1", type=int, default=0, help="with --gpus > 1: run the first K gaps again on one GPU and compare bytes")
    ap.add_argument("--ref-procs", type=int, default=0, help="also run this many gaps through the reference binary CONCURRENTLY at -t 1 (GAPPadder's process-pool shape, SURVEY 8d)")
    ap.add_argument("--config", default="cfg1")
    ap.add_argument("--repeat", type=int, default=1, help="timed runs of the batch; the fastest is reported (the first pays page-cache and module load)")
    args = ap.parse_args()
    binary = os.path.join(ROOT, "build", "ContigsMerger_b200")
    ref = os.path.join(ROOT, "oracle", "_ref", "ContigsMerger")
    cores = len(os.sched_getaffinity(0))
    with tempfile.TemporaryDirectory() as td:
        t0 = time.perf_counter()
        lst = write_gaps(td, args.config, args.gaps, args.seed)
        gen_s = time.perf_counter() - t0
        try:
            runs = [run_binary(binary, lst, args.gpus, args.streams, chunk_gaps=args.chunk_gaps) for _ in range(max(1, args.repeat))]
        except RuntimeError as e:
            print(json.dumps({"error": str(e)}))
            return 1
        wall, stats = min(runs, key=lambda r: r[1]["merge_ms"])
        walls = stats.get("worker_wall_ms", [stats["merge_ms"]])
        busy = [w for w in walls if w > 0]
        line = {"impl": "b200", "config": args.config, "streams": args.streams, "chunk_gaps": args.chunk_gaps, "process_wall_s": wall, "fasta_generation_s": gen_s, **stats,
                "gaps_per_s": args.gaps / (stats["merge_ms"] * 1e-3),
                "gaps_per_s_process": args.gaps / wall,
                "imbalance_max_over_mean": (max(busy) / (sum(busy) / len(busy))) if busy else None,
                # cells actually computed: not the closed-form pairs, not the relax steps shared between chains
                "gcups": (stats["dp_gcells"] - stats.get("closed_gcells", 0.0) - stats.get("relax_shared_gcells", 0.0)) / (stats["merge_ms"] * 1e-3)}
        outs = {g: (open(os.path.join(td, "g%d.out" % g), "rb").read(), open(os.path.join(td, "g%d.info" % g), "rb").read())
                for g in range(min(args.gaps, max(args.ref_gaps, args.verify_gpus1)))}
        if args.verify_gpus1 and args.gpus > 1:
            k = min(args.verify_gpus1, args.gaps)
            sub = os.path.join(td, "sub.tsv")
            with open(sub, "w") as f:
                for g in range(k):
                    f.write("%s\t%s\t%s\n" % (os.path.join(td, "g%d.fa" % g), os.path.join(td, "s%d.out" % g), os.path.join(td, "s%d.info" % g)))
            try:
                run_binary(binary, sub, 1, 1)
                same = all(outs[g] == (open(os.path.join(td, "s%d.out" % g), "rb").read(), open(os.path.join(td, "s%d.info" % g), "rb").read()) for g in range(k))
                line["vs_gpus1"] = {"gaps": k, "outputs_identical": same}
            except RuntimeError as e:
                line["vs_gpus1"] = {"error": str(e)}
        if args.ref_gaps and os.path.exists(ref):
            t_ref, same = 0.0, True
            n = min(args.ref_gaps, args.gaps)
            for g in range(n):
                fa = os.path.join(td, "g%d.fa" % g)
                info = os.path.join(td, "g%d.refinfo" % g)
                t0 = time.perf_counter()
                r = subprocess.run([ref] + FLAGS + ["-t", str(cores), "-o", info, fa], cwd=td, capture_output=True)
                t_ref += time.perf_counter() - t0
                same = same and (r.stdout, open(info, "rb").read()) == outs[g]
            line["reference"] = {"gaps": n, "cores": cores, "seconds": t_ref, "gaps_per_s": n / t_ref, "outputs_identical": same}
        if args.ref_procs and os.path.exists(ref):
            n = min(args.ref_procs, args.gaps)
            t0 = time.perf_counter()
            ps = []
            for g in range(n):
                wd = os.path.join(td, "rp%d" % g)                        # the reference drops ./tmp.gml: one directory per process
                os.makedirs(wd)
                ps.append(subprocess.Popen([ref] + FLAGS + ["-t", "1", "-o", os.path.join(wd, "x.info"), os.path.join(td, "g%d.fa" % g)],
                                           cwd=wd, stdout=open(os.path.join(wd, "x.out"), "wb"), stderr=subprocess.DEVNULL))
            for q in ps:
                q.wait()
            t_ref = time.perf_counter() - t0
            same = all((open(os.path.join(td, "rp%d" % g, "x.out"), "rb").read(), open(os.path.join(td, "rp%d" % g, "x.info"), "rb").read())
                       == (open(os.path.join(td, "g%d.out" % g), "rb").read(), open(os.path.join(td, "g%d.info" % g), "rb").read()) for g in range(n))
            line["reference_process_pool"] = {"gaps": n, "procs": n, "threads_each": 1, "cores": cores, "seconds": t_ref, "gaps_per_s": n / t_ref,
                                              "outputs_identical": same}
        print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
