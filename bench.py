#!/usr/bin/env python
"""bench.py -- overlap-alignment throughput of the B200 path vs the reference's CPU path.

Metric (BASELINE.json): overlap-alignment GCUPS (1e9 DP cell updates / s; one Evaluate(s1,s2) call =
len(s1)*len(s2) cells) and gaps/s, at N GPUs, next to host-CPU ContigsMerger.

Workload (`config.workload`): BASELINE.json configs[0], the synthetic 200-gap ContigsMerger set
(40 Velvet-style contigs per gap, 300-3000 bp, 8 kb locus, GAPPadder's flags -i1 -2 -i2 -2 -y 50
-k 10): every candidate pair of the all-vs-all pairwise phase (contigs + reverse complements, k-mer
quick check) of every gap.  configs[1] ("TERefiner contig-to-flank") has no DP in the reference
(SURVEY.md section 0.3); its builder-defined semi-global form is `--config cfg2`.  One step = one pass of the hot
path over the whole 200-gap batch.  Each rank gets its own 200 gaps (seeds offset by rank): weak scaling, no collective
on the data path.

  value     GCUPS with sequences and pair lists already resident in HBM (CUDA events on the library's
            stream around gp_launch_resident, L2 flushed between steps, max over ranks)
  e2e       GCUPS through the public C call gp_overlap_batch on HOST buffers (ASCII sequences as char pointers,
            the pair list): packing into pinned memory, H2D, kernels, D2H of the results, every step (wall clock
            of the blocking call; the pointer array is built once, as a C++ caller holds it)
  roofline  integer issue-rate roofline of the dominant kernel (overlap_wf16c_kernel on cfg1): achieved =
            GCUPS * 6 integer ops per cell / peak = 2 lanes * measured dual-pipe packed
            16x2 instruction rate (gp_int_peak, measured live on this GPU)
            (traffic: DRAM bytes of one launch from the committed ncu capture; HBM is idle on this path)
  cpu_baseline  the reference's own Evaluate (oracle/_ref/libcm_ref.so, kind "reference") or the C
            restatement (kind "port") on the host cores, on a bounded sample of the same pair list
  parity_sample  (with cpu_baseline) 512 pairs (cfg5: 96) spread over the pair list: the timed end-to-end step's results
            against the oracle, field by field
  dropin    whole gaps per second through build/ContigsMerger_b200 --batch (the chunk pipeline: readers, two mergers per GPU
            behind a device gate, writers; FASTA reading and file writing inside the figure), outside every timed region
            above (tools/dropin_bench.py): `cfg1` (N = 1) the bench workload's own gaps as FASTA files, one of them also through
            the reference binary (bytes compared); `dedup` (N = 1) the dedup stage on the same gaps as contig sets
            (tools/dedup_bench.py; rules pinned to TERefiner_1, alignment records builder-defined); `affine` (N = 1) TERefiner's
            affine local aligner on 20 of the gaps (tools/affine_bench.py, 64 pairs against the reference's own code); `strong` (every N) one
            fixed 1 600-gap cfg3-shaped job through --gpus N: gaps/s, per-GPU wall, imbalance, device-phase time of the
            slowest GPU, first 32 gaps byte-compared with --gpus 1
  hbm       the HBM-bound phases beside the DP: pack + upload, the quick-check kernel (bytes, ms, GB/s against
            MEASURED_PEAKS.json), the result scatter

`--impl reference` times only that CPU path (rank 0 only under torchrun).  `--config cfg3|cfg5` runs the other
BASELINE shapes (cfg5: 10 kb contigs, 20 gaps); `--config cfg2` is BASELINE configs[1], flank placement: the semi-global kernel
(gp_semiglobal_batch) on 2 flanks x 80 nodes per gap, bit-exact against the oracle's builder-written definition -- the reference's
own arithmetic there is BWA's, absent from the reference tree (parity unpinned), so its CPU arm is the oracle port; `--cert-layout 1` forces the certificate kernel's
column-potential layout (A/B); `--config affine` is TERefiner's affine-gap local aligner (gp_local_affine_batch) on cfg1's pair list,
identical to the reference's own local_alignment.cpp, which is also its CPU arm (kind "reference"); the bench line the driver reads is
the default cfg1 run.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "overlap_alignment_gcups"
UNIT = "GCUPS"
OPS_PER_CELL = 6  # 1 compare/select + 3 adds + 2 max


def build_workload(n_gaps: int, first_seed: int, config: str = "cfg1"):
    """-> (seqs: list[bytes], pairs: structured array, cells, per-gap pair counts, per-gap node counts)"""
    import gappadder_b200 as g
    import synth_gaps
    from gappadder_b200.capi import PAIR_DTYPE
    seqs, chunks, per_gap, nodes_per_gap = [], [], [], []
    spec = synth_gaps.CONFIGS[config]
    for gi in range(n_gaps):
        recs = synth_gaps.make_gap(first_seed + gi, spec)
        base = len(seqs)
        nodes = []
        for _, s in recs:                       # graph nodes [c0, c0_R, c1, c1_R, ...] (ContigsCompactor.cpp:794-799)
            nodes.append(s)
            nodes.append(g.revcomp(s))
        cand = g.candidate_pairs(nodes, 10).copy()
        cand["row_seq"] += base
        cand["col_seq"] += base
        seqs.extend(nodes)
        chunks.append(cand)
        per_gap.append(len(cand))
        nodes_per_gap.append(len(nodes))
    pairs = np.concatenate(chunks) if chunks else np.zeros(0, dtype=PAIR_DTYPE)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    cells = int((lens[pairs["row_seq"]] * lens[pairs["col_seq"]]).sum()) if len(pairs) else 0
    return seqs, pairs, cells, per_gap, nodes_per_gap


def rank_first_seed(first_seed: int, rank: int, gaps_per_rank: int) -> int:
    """Weak scaling: rank r works on gaps seeded first_seed + r*gaps .. first_seed + (r+1)*gaps - 1."""
    return first_seed + rank * gaps_per_rank


def reduce_over_ranks(dist, device, cells: int, gaps: int, my_ms: float, my_e2e_ms: float):
    """Whole-job totals: units summed over ranks, time = max over ranks (no collective on the data path;
    this is the only communication in the benchmark).  dist=None for a single process."""
    if dist is None:
        return cells, gaps, my_ms, my_e2e_ms
    import torch
    t = torch.tensor([float(cells), float(gaps)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    mx = torch.tensor([my_ms, my_e2e_ms], dtype=torch.float64, device=device)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return int(t[0].item()), int(t[1].item()), float(mx[0].item()), float(mx[1].item())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU reference arm (oracle/_ref when built, else the oracle port)

def cpu_path():
    """-> (kind, callable(s1, s2))  each call runs one Evaluate on the CPU (GIL released)."""
    import ctypes as C
    import _oracle
    ref = _oracle.ref_lib()
    if ref is not None:
        def run(a, b, _buf=threading.local()):
            out = (C.c_int32 * 8)()
            ref.cmref_evaluate(a, b, 0, out)
        return "reference", run
    lib = _oracle.oracle_lib()

    def run(a, b):
        r = _oracle.DPResult()
        lib.gpo_evaluate(a, len(a), b, len(b), -2, -2, 50, C.byref(r))
    return "port", run


def cpu_run_pairs(run, seqs, pairs, cores: int):
    """Runs the given pairs on `cores` threads; returns elapsed seconds."""
    it = iter(range(len(pairs)))
    lock = threading.Lock()

    def worker():
        while True:
            with lock:
                k = next(it, None)
            if k is None:
                return
            run(seqs[int(pairs["row_seq"][k])], seqs[int(pairs["col_seq"][k])])
    ths = [threading.Thread(target=worker) for _ in range(cores)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return time.perf_counter() - t0


def cpu_sample(seqs, pairs, cores: int, budget_s: float, run):
    """Picks a prefix of the pair list worth about budget_s seconds on `cores` threads (calibrated on a
    small probe) and returns (sample pairs, cells)."""
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    pc = lens[pairs["row_seq"]] * lens[pairs["col_seq"]]
    csum = np.cumsum(pc)
    probe_n = int(np.searchsorted(csum, 3e7 * cores)) + 1          # ~0.03 Gcells per thread
    probe_n = max(cores, min(probe_n, len(pairs)))
    t = cpu_run_pairs(run, seqs, pairs[:probe_n], cores)
    rate = csum[probe_n - 1] / max(t, 1e-6)
    n = int(np.searchsorted(csum, rate * budget_s)) + 1
    n = max(probe_n, min(n, len(pairs)))
    return pairs[:n], int(csum[n - 1])


def parity_sample(seqs, pairs, res, cores: int, n: int = 512):
    """Checker (part of the cpu_baseline leg): n pairs spread over this run's pair list, the timed end-to-end
    step's results against the oracle's Evaluate, field by field."""
    import _oracle
    idx = np.unique(np.linspace(0, len(pairs) - 1, num=min(n, len(pairs)), dtype=np.int64)) if len(pairs) else []
    bad, lock, it = [], threading.Lock(), iter(list(idx))

    def worker():
        while True:
            with lock:
                k = next(it, None)
            if k is None:
                return
            a, b = seqs[int(pairs["row_seq"][k])], seqs[int(pairs["col_seq"][k])]
            o = _oracle.oracle_evaluate(a, b, -2, -2, 50)
            r = res[k]
            want = (o.score, o.row_end, o.col_end, o.nclip, int(o.tb_row == 0), int(o.tb_col == 0), int(o.bcontained))
            f = int(r["flags"])
            got = (int(r["score"]), int(r["row_end"]), int(r["col_end"]), int(r["nclip"]), f & 1, (f >> 1) & 1, (f >> 2) & 1)
            if got != want:
                with lock:
                    bad.append(int(k))
    ths = [threading.Thread(target=worker) for _ in range(max(1, min(cores, len(idx))))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return {"pairs": int(len(idx)), "mismatches": len(bad), "first_bad": bad[:4], "against": "oracle/overlap_oracle.c gpo_evaluate"}


def _dropin_tool(cli, timeout=600):
    tool = os.path.join(ROOT, "tools", "dropin_bench.py")
    if not os.path.exists(os.path.join(ROOT, "build", "ContigsMerger_b200")):
        return {"unavailable": "build/ContigsMerger_b200 not built"}
    try:
        p = subprocess.run([sys.executable, tool] + cli, capture_output=True, text=True, timeout=timeout)
        return json.loads(p.stdout.strip().splitlines()[-1])
    except Exception as e:                                   # the bench line must not depend on it
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


DROPIN_KEEP = ("gaps", "gpus", "workers", "dp_gcells", "pairwise_gcells", "merge_ms", "read_ms", "pairwise_ms", "graph_ms", "relax_ms",
               "relax_steps", "output_ms", "partition_ms", "qc_kernel_ms", "gaps_per_s", "gaps_per_s_process", "process_wall_s", "gcups",
               "worker_wall_ms", "worker_gaps", "worker_gcells", "worker_chunks", "mergers_per_gpu", "chunk_gaps", "detail_ms",
               "imbalance_max_over_mean", "vs_gpus1", "reference", "error", "unavailable")


def dropin_line(args, world):
    """Whole-gap throughput beside the bench line (not part of any timed region above): the drop-in binary
    build/ContigsMerger_b200 --batch on gaps as FASTA files -- read, LPT partition, quick check, pairwise phase, graph,
    relax chains, output -- timed by tools/dropin_bench.py.  gaps_per_s is whole gaps merged per second.
      cfg1    (N = 1) the bench workload's own 200 gaps on one GPU, one of them also through the reference binary (bytes compared)
      strong  the product's multi-GPU path: ONE fixed job (args.dropin_gaps cfg3-shaped gaps, 10-80 contigs each: BASELINE
              configs[2]/[3] shape) through `--batch --gpus N`, gaps assigned to GPUs by gp_partition_gaps (LPT on estimated
              cells), results gathered by the host; per-GPU wall, imbalance, and the first 32 gaps byte-compared with --gpus 1."""
    out = {}
    if world == 1 and args.config == "cfg1":
        d = _dropin_tool(["--gaps", str(args.gaps), "--seed", str(args.seed), "--ref-gaps", "1", "--repeat", "2"])
        out["cfg1"] = {k: d[k] for k in DROPIN_KEEP if k in d}
        # the dedup stage around every merge, same gaps as contig sets: sets per second for both rules
        try:
            p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dedup_bench.py"), "--sets", str(args.gaps), "--repeat", "2"],
                               capture_output=True, text=True, timeout=300)
            dd = json.loads(p.stdout.strip().splitlines()[-1])
            out["dedup"] = {"parity": dd.get("parity"), "sets": dd.get("sets"),
                            "contained_rule": {k: dd["g"][k] for k in ("contigs", "removed", "pairs", "dp_gcells", "dedup_ms", "device_ms", "sets_per_s", "gcups") if k in dd.get("g", {})},
                            "duplicate_rule": {k: dd["p"][k] for k in ("contigs", "removed", "pairs", "dp_gcells", "dedup_ms", "device_ms", "sets_per_s", "gcups") if k in dd.get("p", {})}}
        except Exception as e:                               # the bench line must not depend on it
            out["dedup"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
        # TERefiner's affine local aligner on 20 of these gaps (tools/affine_bench.py: the two kernels by CUDA events, 64 sampled
        # pairs against the reference's own aligner); `bench.py --config affine` is the full line for it
        try:
            p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "affine_bench.py"), "--gaps", "20", "--reps", "3", "--check", "64"],
                               capture_output=True, text=True, timeout=180)
            da = json.loads(p.stdout.strip().splitlines()[-1])
            out["affine"] = {k: da[k] for k in ("workload", "gaps", "pairs", "cells", "forward_ms", "epilogue_ms", "forward_gcups", "total_gcups",
                                                "mean_score", "flagged", "parity_sample") if k in da}
        except Exception as e:                               # the bench line must not depend on it
            out["affine"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    d = _dropin_tool(["--config", "cfg3", "--gaps", str(args.dropin_gaps), "--seed", "5000", "--gpus", str(world), "--ref-gaps", "0",
                      "--verify-gpus1", "32" if world > 1 else "0", "--repeat", "2"])
    out["strong"] = {k: d[k] for k in DROPIN_KEEP if k in d}
    out["strong"]["job"] = "%d cfg3-shaped gaps (seeds 5000..), fixed for every N: strong scaling of whole gaps/s" % args.dropin_gaps
    walls = out["strong"].get("worker_wall_ms")
    if walls:
        t = out["strong"]
        # phases of different chunks overlap (two mergers per GPU alternate on the device): what matters is how long the
        # slowest GPU's device lock was held
        dt = t.get("detail_ms", {})
        dev = sum(dt.get(k, 0) for k in ("pairwise.upload_pairs", "pairwise.kernels_fetch", "relax.device_call"))
        out["strong"]["device_phase_sum_ms"] = dev
        out["strong"]["limiter"] = ("device phases: pairwise + relax launches of the slowest GPU add up to %.0f ms for %.0f ms of wall (they overlap: a relax "
                                    "launch's tail runs beside the next chunk's pairwise kernel)" % (dev, t["merge_ms"])
                                    if dev > 0.6 * t["merge_ms"] else "host phases and pipeline fill (the slowest GPU's launches add up to only %.0f of %.0f ms)" % (dev, t["merge_ms"]))
    return out


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, run = cpu_path()
    cores = min(host_cores(), 64)
    n_gaps = max(1, min(args.gaps, 8))
    seqs, pairs, _, _, _ = build_workload(n_gaps, args.seed, args.config)
    sample, cells = cpu_sample(seqs, pairs, cores, args.cpu_budget, run)
    for _ in range(min(args.warmup, 1)):
        cpu_run_pairs(run, seqs, sample[:max(cores, len(sample) // 8)], cores)
    times = [cpu_run_pairs(run, seqs, sample, cores) for _ in range(args.steps)]
    t = float(np.mean(times))
    v = cells / t / 1e9
    sample_desc = "first %d candidate pairs (%.3f Gcells) of the %s pair list, seeds %d.., per step" % (len(sample), cells / 1e9, args.config, args.seed)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if kind == "reference" else "i32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample_desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


WORKLOADS = {
    "affine": "affine (TERefiner's LocalAlignment::optAlign on ContigsMerger's pairs): %d synthetic cfg1 gaps/GPU x 40 contigs and their reverse "
              "complements (300-3000 bp), affine-gap local alignment 1/-3/-2, open 5, ext 2 (aln_param_blast), score + start/end coordinates",
    "cfg2": "cfg2 (BASELINE configs[1], flank placement): %d synthetic gaps/GPU x 2 flanks (995 bp) x 40 contigs and their reverse "
            "complements (300-3000 bp), semi-global (flank end to end inside the contig); BWA parity unpinned",
    "cfg1": "cfg1: %d synthetic gaps/GPU x 40 contigs (300-3000 bp, 8 kb locus, 0.2%% subst, 50%% RC)",
    "cfg3": "cfg3 (BASELINE configs[2]/[3] shape): %d synthetic gaps/GPU x 10-80 contigs, six k-mer sets (300-3000 bp, 8 kb locus)",
    "cfg5": "cfg5 (BASELINE configs[4], long-contig stress): %d synthetic gaps/GPU x 200 contigs x 10 kb on a 40 kb repeat-rich locus",
}


def committed_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one overlap_wf16c_kernel launch on cfg1 (200 gaps), read from the
    newest committed `ncu --set full` summary under profiles/ (tools/ncu_summary.py output) -> (bytes, file name)."""
    import glob
    import re
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "wf16c_r*.ncu_summary.txt"))):
        txt = open(f).read()
        rd = re.search(r"dram__bytes_read\.sum\s+([0-9.]+) Mbyte", txt)
        wr = re.search(r"dram__bytes_write\.sum\s+([0-9.]+) Mbyte", txt)
        if rd and wr and "<1, 1, 1>" in txt:
            best = ((float(rd.group(1)) + float(wr.group(1))) * 1e6, os.path.relpath(f, ROOT))
    return best


def measured_hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6545.0, "B200_PROFILING.md fallback"


def workload_config(args):
    return {"workload": WORKLOADS[args.config] % args.gaps +
                        (", every flank against every node, scores +1 / -2 / -2" if args.config == "cfg2"
                         else ", all candidate pairs of the k = 10 quick check" if args.config == "affine"
                         else ", all candidate pairs of ContigsMerger's pairwise phase (-i1 -2 -i2 -2 -y 50 -k 10)"),
            "gaps_per_gpu": args.gaps, "first_seed": args.seed, "l2": "flushed between timed steps (256 MiB write)",
            "parallelism": "gaps sharded by rank, no collective"}


# ---------------------------------------------------------------------------------------------
# cfg2: flank placement (semi-global).  Same contract as the overlap arm; the CPU arm is the oracle's builder-written
# definition (kind "port"): GAPPadder's own arithmetic here is BWA's, which is not in /root/reference (parity unpinned).

def build_flank_workload(n_gaps: int, first_seed: int):
    import gappadder_b200 as g
    import synth_gaps
    from gappadder_b200.capi import PAIR_DTYPE
    spec = synth_gaps.CONFIGS["cfg1"]
    seqs, rows, cols = [], [], []
    for gi in range(n_gaps):
        base = len(seqs)
        seqs += [s for _, s in synth_gaps.make_flanks(first_seed + gi, spec)]
        for _, s in synth_gaps.make_gap(first_seed + gi, spec):
            seqs += [s, g.revcomp(s)]
            for f in (base, base + 1):
                rows += [f, f]
                cols += [len(seqs) - 2, len(seqs) - 1]
    pairs = np.zeros(len(rows), dtype=PAIR_DTYPE)
    pairs["row_seq"], pairs["col_seq"] = rows, cols
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    return seqs, pairs, int((lens[pairs["row_seq"]] * lens[pairs["col_seq"]]).sum())


def cpu_flank_path():
    import ctypes as C
    import _oracle
    lib = _oracle.oracle_lib()

    def run(a, b):
        r = _oracle.PlaceResult()
        lib.gpo_semiglobal(a, len(a), b, len(b), -2, -2, C.byref(r))
        return r
    return "port", run


def flank_parity_sample(seqs, pairs, res, cores, n=512):
    _, run = cpu_flank_path()
    idx = np.unique(np.linspace(0, len(pairs) - 1, num=min(n, len(pairs)), dtype=np.int64))
    from concurrent.futures import ThreadPoolExecutor

    def one(k):
        o = run(seqs[int(pairs["row_seq"][k])], seqs[int(pairs["col_seq"][k])])
        r = res[k]
        return (o.score, o.col_start, o.col_end) == (int(r["score"]), int(r["col_start"]), int(r["col_end"]))
    with ThreadPoolExecutor(max_workers=max(1, cores)) as ex:
        ok = list(ex.map(one, idx))
    bad = [int(k) for k, good in zip(idx, ok) if not good]
    return {"pairs": int(len(idx)), "mismatches": len(bad), "first_bad": bad[:4],
            "against": "oracle/overlap_oracle.c gpo_semiglobal (builder-written definition; BWA parity unpinned)"}


def run_reference_cfg2(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    kind, run = cpu_flank_path()
    cores = min(host_cores(), 64)
    seqs, pairs, _ = build_flank_workload(max(1, min(args.gaps, 4)), args.seed)
    sample, cells = cpu_sample(seqs, pairs, cores, args.cpu_budget, run)
    for _ in range(min(args.warmup, 1)):
        cpu_run_pairs(run, seqs, sample[:max(cores, len(sample) // 8)], cores)
    t = float(np.mean([cpu_run_pairs(run, seqs, sample, cores) for _ in range(args.steps)]))
    v = cells / t / 1e9
    desc = "first %d flank-node pairs (%.3f Gcells) of the cfg2 pair list, seeds %d.., per step" % (len(sample), cells / 1e9, args.seed)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc,
                         "note": "the reference's own arithmetic for this path is BWA's (pick_contigs.py:83-86), absent from /root/reference: "
                                 "this is the oracle's semi-global restatement, rolling rows, one pair per thread"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


def run_gpu_cfg2(args):
    import torch
    import gappadder_b200 as g
    from gappadder_b200.capi import HostBatch, PLACE_DTYPE

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = g.Context(local_rank)
    seqs, pairs, cells = build_flank_workload(args.gaps, rank_first_seed(args.seed, rank, args.gaps))
    hb = HostBatch(seqs, pairs)
    ctx.upload_host_sequences(hb)
    ctx.semiglobal_upload_pairs(pairs, g.GAPPADDER_DP)
    st = ctx.semiglobal_stats()
    assert st["cells"] == cells
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        ctx.semiglobal_launch()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    k_ms = 0.0
    barrier()
    for e0, e1 in ev:
        flush.zero_()
        torch.cuda.synchronize(dev)
        e0.record(stream)
        ctx.semiglobal_launch()
        e1.record(stream)
        stream.synchronize()
        k_ms += ctx.semiglobal_stats()["kernel_ms"]
    barrier()
    launches = ctx.kernel_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    my_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in ev]))
    k_ms /= args.steps
    # end to end: the blocking C call on host ASCII buffers (pack into pinned memory, H2D, kernels, D2H)
    out = np.zeros(len(pairs), dtype=PLACE_DTYPE)
    for _ in range(min(args.warmup, 2)):
        ctx.semiglobal_host_batch(hb, out, g.GAPPADDER_DP)
    barrier()
    e2e_t = []
    for _ in range(args.steps):
        out[:] = 0
        t0 = time.perf_counter()
        ctx.semiglobal_host_batch(hb, out, g.GAPPADDER_DP)
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    my_e2e_ms = float(np.mean(e2e_t)) * 1e3
    n_bases = int(sum(len(x) for x in seqs))
    h2d = int(n_bases // 2 + len(pairs) * (16 + 4))
    d2h = int(len(pairs) * 16)
    tot_cells, tot_gaps, max_ms, max_e2e_ms = reduce_over_ranks(dist, dev, cells, args.gaps, my_ms, my_e2e_ms)
    if rank == 0:
        alu, dual = ctx.int_peak()
        peak_lane_ops = 2.0 * dual
        achieved = cells / (k_ms * 1e-3) * OPS_PER_CELL
        line = {
            "metric": METRIC, "value": tot_cells / (max_ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s32", "data": "synthetic",
            "config": workload_config(args), "pairs_per_step": int(len(pairs)) * world, "gcells_per_step": tot_cells / 1e9,
            "placement_gaps_per_s": tot_gaps / (max_ms * 1e-3),
            "parity_status": "bit-exact against the builder-written semi-global definition (oracle gpo_semiglobal); BWA parity unpinned (SURVEY.md 8c)",
            "kernel_split": {"pairs_table": st["table_pairs"], "pairs_generic": st["generic_pairs"]},
            "e2e": {"value": tot_cells / (max_e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": max_e2e_ms, "timing": "wall clock of the blocking gp_semiglobal_batch call"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "int", "achieved": achieved / 1e12, "peak": peak_lane_ops / 1e12, "unit": "Tintop/s", "frac": achieved / peak_lane_ops,
                         "traffic": None, "kernel": "flank_place_kernel", "kernel_ms": k_ms, "kernel_gcells": cells / 1e9, "ops_per_cell": OPS_PER_CELL,
                         "peak_source": "measured live (gp_int_peak): 2 lanes x VIMNMX.S16x2+VIADD.16x2 dual-issue rate, the same denominator as the "
                                        "overlap kernels; this kernel computes one 32-bit cell per instruction pair (score and start column in one word)"},
            "result_checksum": int(out["score"].astype(np.int64).sum()),
        }
        if world == 1 and not args.no_cpu:
            kind, run = cpu_flank_path()
            cores = min(host_cores(), 64)
            sample, scells = cpu_sample(seqs, pairs, cores, args.cpu_budget, run)
            t = cpu_run_pairs(run, seqs, sample, cores)
            line["cpu_baseline"] = {"value": scells / t / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": "first %d flank-node pairs (%.3f Gcells) of this run's pair list, %.1f s" % (len(sample), scells / 1e9, t)}
            line["parity_sample"] = flank_parity_sample(seqs, pairs, out, cores)
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------
# affine: TERefiner's affine-gap local aligner (LocalAlignment::optAlign, TERefiner/algorithms/local_alignment.cpp) on the pair
# list of ContigsMerger's pairwise phase.  Parity is PINNED: the CPU arm is that file compiled as it lies
# (oracle/_ref/libla_ref.so, kind "reference").  GCUPS counts the forward table (len1 * len2 cells per pair); the start
# recovery (reverse pass + banded global fill) is inside every timed figure but adds no cells to the count.

def cpu_affine_path():
    import ctypes as C
    import _oracle
    lib = _oracle.la_ref_lib()
    if lib is None:
        return None, None

    def run(a, b):
        if lib.laref_forward_score(a, b) >= 1:          # the reference reads path[-1] when nothing aligns
            o = (C.c_int32 * 6)()
            lib.laref_stdaln_local(a, b, o)
            return (o[0], o[1], o[2], o[3], o[4])
        return None
    return "reference", run


def affine_parity_sample(seqs, pairs, res, cores, n=512):
    _, run = cpu_affine_path()
    idx = np.unique(np.linspace(0, len(pairs) - 1, num=min(n, len(pairs)), dtype=np.int64))
    from concurrent.futures import ThreadPoolExecutor

    def one(k):
        w = run(seqs[int(pairs["row_seq"][k])], seqs[int(pairs["col_seq"][k])])
        r = res[k]
        if w is None:
            return bool(int(r["flags"]) & 1)
        return w == (int(r["score"]), int(r["start1"]), int(r["end1"]), int(r["start2"]), int(r["end2"]))
    with ThreadPoolExecutor(max_workers=max(1, cores)) as ex:
        ok = list(ex.map(one, idx))
    bad = [int(k) for k, good in zip(idx, ok) if not good]
    return {"pairs": int(len(idx)), "mismatches": len(bad), "first_bad": bad[:4],
            "against": "the reference's own aln_stdaln(.., &aln_param_blast, LOCAL, 1) (oracle/_ref/libla_ref.so): score, start1, end1, start2, end2"}


def run_reference_affine(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    kind, run = cpu_affine_path()
    if run is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libla_ref.so not built (oracle/build_ref.sh needs /root/reference)"}))
        return 0
    cores = min(host_cores(), 64)
    seqs, pairs, _, _, _ = build_workload(max(1, min(args.gaps, 4)), args.seed, "cfg1")
    sample, cells = cpu_sample(seqs, pairs, cores, args.cpu_budget, run)
    for _ in range(min(args.warmup, 1)):
        cpu_run_pairs(run, seqs, sample[:max(cores, len(sample) // 8)], cores)
    t = float(np.mean([cpu_run_pairs(run, seqs, sample, cores) for _ in range(args.steps)]))
    v = cells / t / 1e9
    desc = "first %d pairs (%.3f Gcells) of the pair list, seeds %d.., per step" % (len(sample), cells / 1e9, args.seed)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


def run_gpu_affine(args):
    import torch
    import gappadder_b200 as g
    from gappadder_b200.capi import HostBatch, LOCAL_DTYPE

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = g.Context(local_rank)
    seqs, pairs, cells, _, _ = build_workload(args.gaps, rank_first_seed(args.seed, rank, args.gaps), "cfg1")
    hb = HostBatch(seqs, pairs)
    ctx.upload_host_sequences(hb)
    ctx.local_affine_upload_pairs(pairs)
    assert ctx.local_affine_stats()["cells"] == cells
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        ctx.local_affine_launch()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    f_ms = e_ms = 0.0
    barrier()
    for e0, e1 in ev:
        flush.zero_()
        torch.cuda.synchronize(dev)
        e0.record(stream)
        ctx.local_affine_launch()
        e1.record(stream)
        stream.synchronize()
        st = ctx.local_affine_stats()
        f_ms += st["forward_ms"]
        e_ms += st["epilogue_ms"]
    barrier()
    launches = ctx.kernel_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    my_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in ev]))
    f_ms /= args.steps
    e_ms /= args.steps
    # end to end: the blocking C call on host ASCII buffers (letter classes, pack into pinned memory, H2D, kernels, D2H)
    out = np.zeros(len(pairs), dtype=LOCAL_DTYPE)
    for _ in range(min(args.warmup, 2)):
        ctx.local_affine_host_batch(hb, out)
    barrier()
    e2e_t = []
    for _ in range(args.steps):
        out[:] = 0
        t0 = time.perf_counter()
        ctx.local_affine_host_batch(hb, out)
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    my_e2e_ms = float(np.mean(e2e_t)) * 1e3
    n_bases = int(sum(len(x) for x in seqs))
    h2d = int(n_bases // 2 + len(pairs) * (16 + 4))
    d2h = int(len(pairs) * 24)
    tot_cells, tot_gaps, max_ms, max_e2e_ms = reduce_over_ranks(dist, dev, cells, args.gaps, my_ms, my_e2e_ms)
    if rank == 0:
        alu, dual = ctx.int_peak()
        ops = 10                                    # per forward cell: 2 gap-state updates (2 add + max each), guard compare + select, add + max3/relu
        achieved = cells / (f_ms * 1e-3) * ops
        line = {
            "metric": METRIC, "value": tot_cells / (max_ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s32", "data": "synthetic",
            "config": workload_config(args), "pairs_per_step": int(len(pairs)) * world, "gcells_per_step": tot_cells / 1e9,
            "parity_status": "pinned: identical to the reference's TERefiner/algorithms/local_alignment.cpp (score and the four coordinates)",
            "kernel_ms": {"affine_forward_kernel": f_ms, "affine_epilogue_kernel": e_ms},
            "e2e": {"value": tot_cells / (max_e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": max_e2e_ms, "timing": "wall clock of the blocking gp_local_affine_batch call"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "int", "achieved": achieved / 1e12, "peak": dual / 1e12, "unit": "Tintop/s", "frac": achieved / dual,
                         "traffic": None, "kernel": "affine_forward_kernel", "kernel_ms": f_ms, "kernel_gcells": cells / 1e9, "ops_per_cell": ops,
                         "peak_source": "measured live (gp_int_peak): thread-level integer instruction rate with both pipes busy (%.2f Tinst/s; ALU pipe "
                                        "alone %.2f); this kernel computes ONE 32-bit cell per instruction (three states per cell), so no factor 2 "
                                        "for packed lanes as in the overlap kernels" % (dual / 1e12, alu / 1e12)},
            "mean_score": float(out["score"].mean()), "flagged_pairs": int((out["flags"] != 0).sum()),
            "result_checksum": int(out["score"].astype(np.int64).sum() + out["start1"].astype(np.int64).sum() + out["start2"].astype(np.int64).sum()),
        }
        if world == 1 and not args.no_cpu:
            kind, run = cpu_affine_path()
            if run is not None:
                cores = min(host_cores(), 64)
                sample, scells = cpu_sample(seqs, pairs, cores, args.cpu_budget, run)
                t = cpu_run_pairs(run, seqs, sample, cores)
                line["cpu_baseline"] = {"value": scells / t / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                                        "sample": "first %d pairs (%.3f Gcells) of this run's pair list, %.1f s" % (len(sample), scells / 1e9, t)}
                line["parity_sample"] = affine_parity_sample(seqs, pairs, out, cores)
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm

def run_gpu(args):
    import torch
    import gappadder_b200 as g

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    ctx = g.Context(local_rank)                 # fails loudly without the CUDA library / a B200
    ctx.set_cert_layout(args.cert_layout)
    ctx.set_orientation(args.orientation)
    seqs, pairs, cells, per_gap, nodes_per_gap = build_workload(args.gaps, rank_first_seed(args.seed, rank, args.gaps), args.config)
    packed, off, lens, nsym = g.pack_sequences(seqs)
    ctx.set_sequences(packed, off, lens, nsym)
    ctx.set_kernel_mask(args.kernel_mask)
    ctx.upload_pairs(pairs, g.GAPPADDER_DP)
    stats = ctx.pair_stats()
    split = ctx.pair_split()
    closed = ctx.closed_form_stats()
    ref_cells = cells                           # what the reference computes for this pair list
    cells = stats["cells"]                      # cells a kernel computes: closed-form pairs (s against s) are not counted
    assert cells + closed["cells"] == ref_cells
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # --- device-resident timing ---------------------------------------------------------------
    for _ in range(args.warmup):
        ctx.launch_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ktimes = {}                                 # per-kernel device ms, summed over the timed steps (library events)
    for e0, e1 in ev:
        flush.zero_()                           # L2 flush between timed steps
        torch.cuda.synchronize(dev)
        e0.record(stream)
        ctx.launch_resident()
        e1.record(stream)
        stream.synchronize()
        for k, v in ctx.kernel_times().items():
            ktimes.setdefault(k, dict(ms=0.0, cells=v["cells"]))["ms"] += v["ms"]
    barrier()
    launches = ctx.kernel_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    my_ms = float(np.mean(step_ms))

    # --- end to end through the public call on host buffers ------------------------------------
    # The host buffers are given the C ABI's shape once (pointer array, lengths, pair array: what a C++ caller
    # holds anyway); the timed region is the blocking C call: pack into pinned memory, H2D, kernels, D2H.
    from gappadder_b200.capi import HostBatch
    hb = HostBatch(seqs, pairs)
    for _ in range(min(args.warmup, 2)):
        ctx.overlap_host_batch(hb, g.GAPPADDER_DP)
    barrier()
    e2e_t = []
    res = None
    for _ in range(args.steps):
        hb.out[:] = 0                                        # every step must produce its results anew
        t0 = time.perf_counter()
        res = ctx.overlap_host_batch(hb, g.GAPPADDER_DP)
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    my_e2e_ms = float(np.mean(e2e_t)) * 1e3
    breakdown = ctx.last_timing()
    h2d = int(packed.nbytes + len(pairs) * (16 + 4))      # packed table + PairDesc + work order
    d2h = int(len(pairs) * 20)
    checksum = int(res["score"].astype(np.int64).sum()) if res is not None and len(res) else 0
    cert = ctx.cert_stats()

    # --- HBM / PCIe-bound byte phases (rank 0): packing + table upload, quick check on the device, result scatter ----
    hbm = None
    if rank == 0:
        n_bases = int(sum(len(x) for x in seqs))
        peak_gbs, peak_src = measured_hbm_peak()
        t_up = []
        for _ in range(3):
            t0 = time.perf_counter()
            ctx.upload_host_sequences(hb)                      # ASCII -> 4-bit codes in pinned memory (host threads) -> HBM
            t_up.append(time.perf_counter() - t0)
        up_ms = min(t_up) * 1e3
        gap_first = np.concatenate([[0], np.cumsum(nodes_per_gap)]).astype(np.uint32)
        qc_ms, qc_items, qc_equal = None, None, None
        if max(nodes_per_gap) <= 4096:
            flush.zero_()
            torch.cuda.synchronize(dev)
            qc = ctx.quick_check_device(gap_first, 10)
            st = ctx.quick_check_stats()
            qc_ms, qc_items = st["kernel_ms"], st["items"]
            qc_pairs = sum(len(x) for x in qc)
            qc_equal = bool(qc_pairs == len(pairs))               # same candidate list size as the host filter that built the workload
        qc_bytes = n_bases // 2 + n_bases // 8                     # packed codes once + 4 B of carry-in per 32 bases
        hbm = {
            "peak_GBps": peak_gbs, "peak_source": peak_src, "bases": n_bases,
            "pack_and_upload": {"what": "gp_upload_sequences: 1 B/base ASCII read + 0.5 B/base written to pinned memory (host threads), 0.5 B/base H2D",
                                "bytes": int(n_bases * 2), "ms": up_ms, "GBps": n_bases * 2 / (up_ms * 1e-3) / 1e9,
                                "bound": "host memory + PCIe, not HBM"},
            "quick_check": {"what": "quick_check_kernel on the packed table in HBM, L2 flushed before it", "bytes": int(qc_bytes), "kernel_ms": qc_ms,
                            "GBps": (qc_bytes / (qc_ms * 1e-3) / 1e9) if qc_ms else None,
                            "frac_of_hbm_peak": (qc_bytes / (qc_ms * 1e-3) / 1e9 / peak_gbs) if qc_ms else None,
                            "work_items": qc_items, "candidate_count_equals_host_filter": qc_equal},
            "result_scatter": {"what": "20 B per pair written by the DP kernels, one D2H copy", "bytes": d2h},
        }

    # --- reduce over ranks ---------------------------------------------------------------------
    tot_cells, tot_gaps, max_ms, max_e2e_ms = reduce_over_ranks(dist, dev, cells, args.gaps, my_ms, my_e2e_ms)

    if rank == 0:
        value = tot_cells / (max_ms * 1e-3) / 1e9
        e2e_v = tot_cells / (max_e2e_ms * 1e-3) / 1e9
        alu, dual = ctx.int_peak()
        traffic = committed_ncu_traffic()
        peak_lane_ops = 2.0 * dual                             # two 16-bit lanes per packed instruction
        # dominant kernel: the one with the most host-routed cells; its own launch time (library events around it)
        kname = max(ktimes, key=lambda k: ktimes[k]["cells"])
        k_ms = ktimes[kname]["ms"] / args.steps
        k_cells = ktimes[kname]["cells"]
        achieved = k_cells / (k_ms * 1e-3) * OPS_PER_CELL
        kernel_fn = {"cert16": "overlap_wf16c_kernel", "table16": "overlap_wf16t_kernel", "prmt16": "overlap_wf16_kernel",
                     "wide32": "overlap_wf32_kernel"}[kname]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s16x2", "data": "synthetic", "config": workload_config(args),
            "pairwise_gaps_per_s": tot_gaps / (max_ms * 1e-3),      # pairwise phase only; whole gaps/s is dropin.*.gaps_per_s
            "pairs_per_step": int(len(pairs)) * world, "gcells_per_step": tot_cells / 1e9,
            "kernel_split": {"pairs_cert16": cert["cert16"], "pairs_table16": split["table16"], "pairs_prmt16": split["prmt16"],
                             "pairs_wide32": split["wide32"], "cert_second_passes": cert["second_passes"],
                             "cert_exact_retries": cert["exact_retries"]},
            "closed_form": {"pairs_per_gpu": closed["pairs"], "gcells_per_gpu_not_counted": closed["cells"] / 1e9,
                            "note": "a node against itself (ContigsCompactor.cpp:1068-1100, j starts at i) has a proven closed form; "
                                    "its m*n cells are in neither value, e2e nor roofline"},
            "kernel_ms_per_step": {k: round(v["ms"] / args.steps, 4) for k, v in ktimes.items()},
            "kernel_gcells": {k: v["cells"] / 1e9 for k, v in ktimes.items()},
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": max_e2e_ms, "pairwise_gaps_per_s": tot_gaps / (max_e2e_ms * 1e-3), "timing": "wall clock of the blocking gp_overlap_batch call",
                    "last_call_breakdown_ms": {k: round(v, 3) for k, v in breakdown.items()}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "int", "achieved": achieved / 1e12, "peak": peak_lane_ops / 1e12, "unit": "Tintop/s",
                         "frac": achieved / peak_lane_ops, "traffic": traffic[0] if traffic and args.config == "cfg1" and args.gaps == 200 else None,
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch on this workload, read from the committed "
                                         "ncu --set full summary %s (ncu cannot run inside a timed bench); algorithmic bytes per launch: %d "
                                         "(packed table + pair descriptors + work order + results); HBM is idle, the bound is the integer issue rate"
                                         % (traffic[1] if traffic else "(none)", h2d + d2h),
                         "kernel": kernel_fn, "kernel_ms": k_ms, "kernel_gcells": k_cells / 1e9, "ops_per_cell": OPS_PER_CELL,
                         "peak_source": "measured live (gp_int_peak): 2 lanes x VIMNMX.S16x2+VIADD.16x2 dual-issue rate; "
                                        "ALU pipe alone %.2f Tinst/s, both pipes %.2f Tinst/s" % (alu / 1e12, dual / 1e12)},
            "result_checksum": checksum,
        }
        if world == 1 and not args.no_cpu:
            kind, run = cpu_path()
            cores = min(host_cores(), 64)
            sample, scells = cpu_sample(seqs, pairs, cores, args.cpu_budget, run)
            t = cpu_run_pairs(run, seqs, sample, cores)
            line["cpu_baseline"] = {"value": scells / t / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": "first %d candidate pairs (%.3f Gcells) of this run's pair list, %.1f s" % (len(sample), scells / 1e9, t)}
            line["parity_sample"] = parity_sample(seqs, pairs, res, cores, 96 if args.config == "cfg5" else 512)
        line["hbm"] = hbm
        print_line = line
    else:
        print_line = None
    ctx.close()
    # Whole-gap throughput through the drop-in binary (the product's own multi-GPU path at N > 1): rank 0 runs it as a
    # subprocess on all N GPUs while the other ranks wait on the CPU (a store key, not an NCCL barrier, whose kernel
    # would spin on their GPUs).
    if not args.no_dropin:
        if dist is not None:
            store = dist.distributed_c10d._get_default_store()
            torch.cuda.synchronize(dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
            if rank == 0:
                print_line["dropin"] = dropin_line(args, world)
                store.set("gp_dropin_done", "1")
            else:
                store.wait(["gp_dropin_done"])
        elif args.config == "cfg1":
            print_line["dropin"] = dropin_line(args, world)
    if rank == 0:
        print(json.dumps(print_line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gaps", type=int, default=None, help="gaps per GPU per step (default: cfg1/cfg3 200, cfg5 20 = BASELINE's counts)")
    ap.add_argument("--dropin-gaps", type=int, default=1600, help="size of the fixed cfg3-shaped job of the drop-in's strong-scaling run")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--config", default="cfg1", choices=sorted(WORKLOADS), help="workload shape (tools/synth_gaps.py); the bench line is cfg1")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work in the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-dropin", action="store_true", help="skip the whole-ContigsMerger (drop-in binary) timing beside the bench line")
    ap.add_argument("--cert-layout", type=int, default=0, help="A/B: certificate kernel value layout (0 free moves when a launch allows it, 1 column potential only)")
    ap.add_argument("--orientation", type=int, default=0, help="A/B: certificate kernel pair orientation (0 longer sequence as rows, 1 never transpose, 2 always)")
    ap.add_argument("--kernel-mask", type=int, default=15, help="A/B: what the library may use (1 table, 2 PRMT, 4 certificate kernel, 8 closed form for s-vs-s)")
    args = ap.parse_args()
    if args.gaps is None:
        args.gaps = 20 if args.config == "cfg5" else 50 if args.config == "affine" else 200
    if args.config == "cfg2":
        return run_reference_cfg2(args) if args.impl == "reference" else run_gpu_cfg2(args)
    if args.config == "affine":
        return run_reference_affine(args) if args.impl == "reference" else run_gpu_affine(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
