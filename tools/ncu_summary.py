#!/usr/bin/env python
"""tools/ncu_summary.py -- condenses an .ncu-rep (read here, no GPU needed) into the text summary kept
under profiles/: headline metrics, pipe utilisation, stall reasons, opcode mix and the hottest SASS
regions.  usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.summary.txt"""
import collections
import csv
import io
import subprocess
import sys


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    raw = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    col = {h: i for i, h in enumerate(hdr)}

    def get(name):
        return vals[col[name]] if name in col else "n/a"

    print("report:", rep)
    print("kernel:", get("Kernel Name"), " grid", get("Grid Size"), " block", get("Block Size"))
    keys = [
        "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_lsu.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    ]
    for k in keys:
        if k in col:
            print("  %-70s %s %s" % (k, vals[col[k]], units[col[k]]))
    print("stall reasons (warps per issue-active cycle):")
    st = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            st.append((float(vals[col[h]].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    for v, n in sorted(st, reverse=True)[:10]:
        print("  %-28s %.3f" % (n, v))

    src = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    h2 = src[1]
    c2 = {h: i for i, h in enumerate(h2)}
    rows = src[2:]
    ie = c2["Instructions Executed"]
    tot = sum(int(r[ie]) for r in rows)
    print("SASS instructions: %d static, %d executed (warp level)" % (len(rows), tot))
    byop = collections.Counter()
    for r in rows:
        t = r[c2["Source"]].split()
        op = t[1] if t[0].startswith("@") else t[0]
        byop[op] += int(r[ie])
    print("opcode mix (share of executed warp instructions):")
    for op, c in byop.most_common(24):
        print("  %-26s %6.2f%%" % (op, 100.0 * c / tot))
    # hot regions: maximal runs of instructions with (nearly) the same execution count
    regs, cur = [], None
    for n, r in enumerate(rows):
        c = int(r[ie])
        if cur and abs(c - cur["c"]) <= 0.12 * max(c, cur["c"], 1):
            cur["end"] = n; cur["tot"] += c; cur["n"] += 1
        else:
            if cur:
                regs.append(cur)
            cur = {"start": n, "end": n, "c": c, "tot": c, "n": 1}
    regs.append(cur)
    print("hot SASS regions (>= 1% of executed instructions):")
    for g in sorted(regs, key=lambda g: -g["tot"]):
        if g["tot"] < 0.01 * tot:
            break
        ops = collections.Counter()
        for r in rows[g["start"]:g["end"] + 1]:
            t = r[c2["Source"]].split()
            ops[t[1] if t[0].startswith("@") else t[0]] += 1
        print("  instr %5d-%5d  %3d instr/iter  %5.2f%% of all   %s" % (g["start"], g["end"], g["n"], 100.0 * g["tot"] / tot,
              ", ".join("%s x%d" % kv for kv in ops.most_common(8))))


if __name__ == "__main__":
    main()
