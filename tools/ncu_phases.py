#!/usr/bin/env python3
"""Phase view of an ncu report of pileup_kernel (runs without a GPU): executed warp-instructions, stall samples and thread-level
lane efficiency of every SASS instruction, attributed to the phase of the kernel its address falls in. Phases are found from the
kernel's source-line markers (the "// ---- N." comments of the consumer loop); helper functions inlined into a phase count for it
because attribution goes by SASS order, not by source line.

  tools/ncu_phases.py <report.ncu-rep> <kernels.cuh>
"""
import csv, io, re, subprocess, sys


def main():
    rep, src = sys.argv[1], sys.argv[2]
    marks = []                      # (line, name)
    for n, l in enumerate(open(src), 1):
        m = re.search(r"// ---- (\d+)\. ([^:]+):", l)
        if m:
            marks.append((n, m.group(1) + " " + m.group(2).strip()))
        if "pileup_producer(smem, L, sh" in l and "__device__" not in l:
            marks.append((n, "producer call"))
        if "------ consumers" in l:
            marks.append((n, "consumer setup"))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    iE, iS, iT = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    out2 = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    # map SASS address -> source line through the correlated view
    rows2 = list(csv.reader(io.StringIO(out2)))
    hdr2 = next(r for r in rows2 if r and r[0] == "Line No")
    cur = 0
    addr_line = {}
    for r in rows2:
        if len(r) != len(hdr2) or r[0] == "Line No":
            continue
        if r[0]:
            cur = int(r[0])
        elif r[2].startswith("0x"):
            addr_line.setdefault(r[2], cur)
    phases, order = {}, []
    cur_phase = "prologue"
    for r in rows:
        if len(r) != len(hdr) or not r[0].startswith("0x"):
            continue
        ln = addr_line.get(r[0], 0)
        for n, name in marks:
            if ln == n or ln == n + 1 or ln == n + 2:
                cur_phase = name
        if cur_phase not in phases:
            phases[cur_phase] = [0, 0, 0, 0]
            order.append(cur_phase)
        p = phases[cur_phase]
        p[0] += int(r[iE]); p[1] += int(r[iS]); p[2] += int(r[iT]); p[3] += 1
    te = sum(p[0] for p in phases.values()) or 1
    ts = sum(p[1] for p in phases.values()) or 1
    print("%-40s %8s %8s %8s %10s" % ("phase (in SASS order)", "inst %", "samp %", "SASS", "lanes/inst"))
    for name in order:
        p = phases[name]
        print("%-40s %8.2f %8.2f %8d %10.1f" % (name, 100.0 * p[0] / te, 100.0 * p[1] / ts, p[3], p[2] / p[0] if p[0] else 0))


if __name__ == "__main__":
    main()
