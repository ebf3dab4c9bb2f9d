#!/usr/bin/env python
"""tools/process_per_gap_bench.py -- GAPPadder's own call shape, unchanged Python side: one `ContigsMerger` process per gap,
`procs` of them at a time (Pool(nthreads).map(run_merge), /root/reference/assemble_gaps.py:296-318; MergeContigs.py:85).
  direct  every process is the drop-in binary in-process: CUDA context + module load per gap
  server  the same command line as a thin client of `ContigsMerger_b200 --serve` (GAPPADDER_B200_SOCKET): concurrent gaps share launches
Outputs of both modes are compared byte for byte with each other (and with the --batch form).  Prints one JSON line."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth_gaps  # noqa: E402

FLAGS = "-s 0.4 -i1 -2.0 -i2 -2.0 -x 12 -y 50 -k 10 -t 5 -m 1".split()
BIN = os.path.join(ROOT, "build", "ContigsMerger_b200")


def run_one(args):
    d, env = args
    p = subprocess.run([BIN] + FLAGS + ["-o", "x.merge.info", "contigs.fa"], cwd=d, capture_output=True, env=env)
    with open(os.path.join(d, "merged.fa"), "wb") as f:
        f.write(p.stdout)
    return p.returncode


def collect(dirs):
    return [(open(os.path.join(d, "merged.fa"), "rb").read(), open(os.path.join(d, "x.merge.info"), "rb").read()) for d in dirs]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaps", type=int, default=64)
    ap.add_argument("--procs", type=int, default=16, help="concurrent ContigsMerger processes (GAPPadder's nthreads)")
    ap.add_argument("--config", default="cfg1")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--direct-gaps", type=int, default=32, help="gaps of the direct (context per process) leg; 0 skips it")
    a = ap.parse_args()
    line = {"config": a.config, "gaps": a.gaps, "procs": a.procs}
    with tempfile.TemporaryDirectory() as td:
        dirs = []
        for g in range(a.gaps):
            d = os.path.join(td, "gap%d" % g)
            os.makedirs(d)
            synth_gaps.write_fasta(os.path.join(d, "contigs.fa"), synth_gaps.make_gap(a.seed + g, synth_gaps.CONFIGS[a.config]))
            dirs.append(d)
        env = dict(os.environ)
        env.pop("GAPPADDER_B200_SOCKET", None)
        nd = min(a.direct_gaps, a.gaps)
        direct = None
        if nd:
            t0 = time.perf_counter()
            with ThreadPoolExecutor(max_workers=a.procs) as ex:
                rcs = list(ex.map(run_one, [(d, env) for d in dirs[:nd]]))
            t = time.perf_counter() - t0
            direct = collect(dirs[:nd])
            line["direct"] = {"gaps": nd, "seconds": t, "gaps_per_s": nd / t, "exit_codes_ok": all(r == 0 for r in rcs),
                              "what": "one drop-in process per gap, CUDA context and module load in each"}
        sock = os.path.join(td, "gp.sock")
        srv = subprocess.Popen([BIN, "--serve", sock], stderr=subprocess.PIPE)
        try:
            t0 = time.perf_counter()
            while not os.path.exists(sock) and time.perf_counter() - t0 < 60:
                time.sleep(0.02)
            startup = time.perf_counter() - t0
            senv = dict(env, GAPPADDER_B200_SOCKET=sock)
            with ThreadPoolExecutor(max_workers=a.procs) as ex:           # warm-up wave (buffers, page cache)
                list(ex.map(run_one, [(d, senv) for d in dirs[:a.procs]]))
            t0 = time.perf_counter()
            with ThreadPoolExecutor(max_workers=a.procs) as ex:
                rcs = list(ex.map(run_one, [(d, senv) for d in dirs]))
            t = time.perf_counter() - t0
            served = collect(dirs)
            line["server"] = {"gaps": a.gaps, "seconds": t, "gaps_per_s": a.gaps / t, "server_startup_s": startup, "exit_codes_ok": all(r == 0 for r in rcs),
                              "what": "the same command line per gap as a thin client of ContigsMerger_b200 --serve; %d clients at a time" % a.procs}
            if direct is not None:
                line["direct_vs_server_identical"] = direct == served[:nd]
            subprocess.run([BIN, "--shutdown"], env=senv, timeout=30)
            srv.wait(timeout=30)
            line["server"]["log"] = srv.stderr.read().decode().strip().splitlines()[-1]
        finally:
            if srv.poll() is None:
                srv.kill()
        # the --batch form on the same gaps, for the bytes
        lst = os.path.join(td, "l.tsv")
        with open(lst, "w") as f:
            for g, d in enumerate(dirs):
                f.write("%s\t%s\t%s\n" % (os.path.join(d, "contigs.fa"), os.path.join(d, "b.out"), os.path.join(d, "b.info")))
        p = subprocess.run([BIN] + FLAGS + ["--batch", lst, "--no-gml"], capture_output=True)
        if p.returncode == 0:
            line["server_vs_batch_identical"] = served == [(open(os.path.join(d, "b.out"), "rb").read(), open(os.path.join(d, "b.info"), "rb").read()) for d in dirs]
    print(json.dumps(line))


if __name__ == "__main__":
    main()
