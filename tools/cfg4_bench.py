#!/usr/bin/env python
"""tools/cfg4_bench.py -- BASELINE configs[3]: a mammalian-chromosome-scale draft, 50 000 gaps (10-80 contigs each, six k-mer
sets: the cfg3 distribution), sharded by gap across 1 / 2 / 4 / 8 B200s through the product's own path:
build/ContigsMerger_b200 --batch LIST --gpus N (chunk pipeline: stat-based LPT partition, reader pool, two mergers per GPU
behind a device lock, writers).  One fixed job for every N (strong scaling of whole gaps per second); the gap files are
generated once by a process pool; after every run a sample of gaps is byte-compared with the N = 1 outputs of the same
gaps.  One JSON line.  Run on an 8-GPU box: gpurun --gpus 8 -- python tools/cfg4_bench.py"""
import argparse
import hashlib
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth_gaps  # noqa: E402

FLAGS = "-s 0.4 -i1 -2.0 -i2 -2.0 -x 12 -y 50 -k 10 -m 1 -t 5".split()


def _write(job):
    td, g, seed = job
    synth_gaps.write_fasta(os.path.join(td, "in", "g%d.fa" % g), synth_gaps.make_gap(seed + g, synth_gaps.CONFIGS["cfg3"]))
    return g


def digest(td, sub, gaps):
    h = hashlib.sha256()
    for g in gaps:
        for ext in ("out", "info"):
            p = os.path.join(td, sub, "g%d.%s" % (g, ext))
            h.update(open(p, "rb").read() if os.path.exists(p) else b"<missing>")
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaps", type=int, default=50000)
    ap.add_argument("--seed", type=int, default=100000)
    ap.add_argument("--gpus", default="8,4,2,1", help="comma-separated GPU counts to run, in this order")
    ap.add_argument("--sample", type=int, default=500, help="gaps byte-compared between every N and the smallest N run")
    ap.add_argument("--tmp", default=None)
    args = ap.parse_args()
    binary = os.path.join(ROOT, "build", "ContigsMerger_b200")
    ns = [int(x) for x in args.gpus.split(",")]
    out = {"config": "cfg4 (BASELINE configs[3]): %d cfg3-shaped gaps, one fixed job for every N" % args.gaps, "gaps": args.gaps, "runs": []}
    with tempfile.TemporaryDirectory(dir=args.tmp) as td:
        os.makedirs(os.path.join(td, "in"))
        t0 = time.perf_counter()
        with mp.Pool(min(32, os.cpu_count() or 1)) as pool:
            for _ in pool.imap_unordered(_write, [(td, g, args.seed) for g in range(args.gaps)], chunksize=64):
                pass
        out["fasta_generation_s"] = time.perf_counter() - t0
        out["fasta_bytes"] = sum(os.path.getsize(os.path.join(td, "in", "g%d.fa" % g)) for g in range(args.gaps))
        step = max(1, args.gaps // max(1, args.sample))
        sample = list(range(0, args.gaps, step))
        ref_digest = None
        for n in ns:
            sub = "out%d" % n
            os.makedirs(os.path.join(td, sub))
            lst = os.path.join(td, "list%d.tsv" % n)
            with open(lst, "w") as f:
                for g in range(args.gaps):
                    f.write("%s\t%s\t%s\n" % (os.path.join(td, "in", "g%d.fa" % g), os.path.join(td, sub, "g%d.out" % g), os.path.join(td, sub, "g%d.info" % g)))
            t0 = time.perf_counter()
            p = subprocess.run([binary] + FLAGS + ["--batch", lst, "--gpus", str(n), "--no-gml", "--stats"], capture_output=True, text=True)
            wall = time.perf_counter() - t0
            if p.returncode != 0:
                out["runs"].append({"gpus": n, "error": p.stderr[-400:]})
                continue
            st = json.loads(p.stderr.strip().splitlines()[-1])
            d = digest(td, sub, sample)
            walls = [w for w in st["worker_wall_ms"] if w > 0]
            dt = st.get("detail_ms", {})
            run = {"gpus": n, "merge_ms": st["merge_ms"], "setup_ms": st.get("setup_ms"), "process_wall_s": wall,
                   "gaps_per_s": args.gaps / (st["merge_ms"] * 1e-3), "gaps_per_s_process": args.gaps / wall,
                   "dp_gcells": st["dp_gcells"], "gcups": (st["dp_gcells"] - st.get("closed_gcells", 0.0)) / (st["merge_ms"] * 1e-3),
                   "worker_wall_ms": st["worker_wall_ms"], "worker_gaps": st["worker_gaps"], "worker_chunks": st.get("worker_chunks"),
                   "imbalance_max_over_mean": max(walls) / (sum(walls) / len(walls)),
                   "slowest_gpu_device_lock_ms": sum(dt.get(k, 0) for k in ("read.upload", "read.quick_check", "pairwise.upload_pairs", "pairwise.kernels_fetch", "relax.device_call")),
                   "slowest_gpu_phase_sums_ms": {k: st[k] for k in ("read_ms", "pairwise_ms", "graph_ms", "relax_ms", "output_ms")},
                   "sample_digest": d}
            if ref_digest is None:
                ref_digest = d
            run["sample_identical_to_first_run"] = d == ref_digest
            out["runs"].append(run)
            # free the disk for the next run (keep nothing but the digest)
            subprocess.run(["rm", "-rf", os.path.join(td, sub)])
        ok = [r for r in out["runs"] if "error" not in r]
        base = next((r for r in ok if r["gpus"] == min(x["gpus"] for x in ok)), None)
        if base:
            for r in ok:
                r["speedup_vs_smallest_n"] = r["gaps_per_s"] / base["gaps_per_s"]
                r["efficiency_vs_smallest_n"] = r["speedup_vs_smallest_n"] / (r["gpus"] / base["gpus"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
