#!/usr/bin/env python
"""tools/ncu_summary.py <report.ncu-rep> [iterations] -- text summary of one ncu --set full
capture (raw + source pages) for profiles/.  `iterations` = solver iterations executed by
the profiled launch (to normalise instruction counts)."""
import collections
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    iters = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = page(rep, "raw")
    d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
    keys = [
        "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    ]
    print(f"# ncu summary of {rep}")
    for k in keys:
        if k in d:
            print(f"{k:85s} {d[k][1]} {d[k][0]}")
    print("\n# warp stall reasons (warps stalled per issue-active cycle)")
    for k in sorted(d):
        if "issue_stalled" in k and "per_issue_active" in k:
            print(f"  {k.split('stalled_')[1].split('_per')[0]:22s} {float(d[k][1]):.3f}")
    rows = page(rep, "source")
    hdr, body = rows[1], rows[2:]
    iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    byop, samp, tot = collections.Counter(), collections.Counter(), 0
    for r in body:
        try:
            n, s = int(r[iE]), int(r[iSamp])
        except ValueError:
            continue
        toks = r[iS].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        byop[op] += n
        samp[op] += s
        tot += n
    print(f"\n# executed warp instructions: {tot}  (SASS lines {len(body)})")
    if iters:
        print(f"# per solver iteration ({iters:.0f} iterations in this launch): {tot / iters:.1f}")
    for op, n in byop.most_common(16):
        per = f"{n / iters:8.1f}/iter" if iters else ""
        print(f"  {op:8s} {n:14d} {per}  samples {samp[op]}")


if __name__ == "__main__":
    main()
