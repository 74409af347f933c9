#!/usr/bin/env python3
"""Per-source-line view of an ncu report taken with --import-source on (runs without a GPU).

  tools/ncu_lines.py <report.ncu-rep> [--top N] [--metrics]

Joins the SASS page of the report (instructions executed, stall samples per instruction) with the
source lines recorded in it, and prints: the summary metrics that matter for this path, the share of
executed warp-instructions and of stall samples per source line, and the dominant stall reasons.
"""
import argparse, collections, csv, io, re, subprocess, sys


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--min", type=float, default=0.4, help="print lines with at least this percentage of instructions or samples")
    a = ap.parse_args()

    raw = ncu(["-i", a.report, "--page", "raw", "--csv"])
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) >= 3:
        hdr, units, vals = rows[0], rows[1], rows[2]
        want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size",
                "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
                "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum", "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_shared_atom_dot_alu.sum",
                "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
                "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
                "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active_mem_lg.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
                "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_inst0.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")
        for k in want:
            if k in hdr:
                i = hdr.index(k)
                print("%-72s %s %s" % (k, vals[i], units[i]))
        for i, k in enumerate(hdr):
            if k.startswith("smsp__average_warp") and "issue_stalled" in k and k.endswith("_per_warp_active.pct"):
                pass
        st = [(float(vals[i].replace(",", "")), k) for i, k in enumerate(hdr)
              if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and vals[i]]
        for v, k in sorted(st, reverse=True)[:8]:
            print("  stall %-60s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))

    cuda = ncu(["-i", a.report, "--page", "source", "--csv", "--print-source", "cuda,sass"])
    rows = list(csv.reader(io.StringIO(cuda)))
    hdr = None
    lines, sass = [], []
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr):
            if r[0]:
                lines.append(r)
            elif r[2].startswith("0x"):
                sass.append(r)
    if not hdr:
        print("no source page in the report (was it taken with --import-source on and built with -lineinfo?)")
        return 1
    iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    tot_e = sum(int(r[iE]) for r in sass) or 1
    tot_s = sum(int(r[iS]) for r in sass) or 1
    print("\n%d SASS instructions, %.4g warp-instructions executed, %d samples" % (len(sass), tot_e, tot_s))
    print("source lines with >= %.1f %% of the instructions or samples:" % a.min)
    for r in lines:
        e, sm = 100.0 * int(r[iE]) / tot_e, 100.0 * int(r[iS]) / tot_s
        if e >= a.min or sm >= a.min:
            top = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
            print("  L%-5s inst %5.2f%%  samp %5.2f%%  %-90s %s" % (r[0], e, sm, r[1].strip()[:90], ", ".join("%s %d" % (n, v) for v, n in top if v)))
    print("top instructions by samples:")
    for r in sorted(sass, key=lambda r: -int(r[iS]))[:a.top]:
        top = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print("  %s  inst %5.2f%%  samp %5.2f%%  %-58s %s" % (r[2][-5:], 100.0 * int(r[iE]) / tot_e, 100.0 * int(r[iS]) / tot_s, r[3].strip()[:58],
                                                           ", ".join("%s %d" % (n, v) for v, n in top if v)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
