#!/usr/bin/env python3
"""Dynamic instruction mix of one kernel from an ncu report (source page, read on the CPU) and the dispatch-port model of DESIGN.md
section 6 applied to it: a packed f32x2 instruction holds a scheduler's dispatch port for two cycles (tools/issue_bench.cu), so a
line costs 2 * packed + other warp instructions; compared with the cycles per line and scheduler the kernel actually takes.

    tools/ncu_dispatch_model.py report.ncu-rep lines warps_per_line [kernel_ms]

lines: A-scans the captured launch processed; kernel_ms: the kernel's duration in the un-profiled bench run (else the report's)."""
import csv
import io
import subprocess
import sys
from collections import Counter


def main():
    rep, lines, wpl = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    ms = float(sys.argv[4]) if len(sys.argv) > 4 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    col = {name: k for k, name in enumerate(rows[hdr])}
    ops, stalls = Counter(), Counter()
    for r in rows[hdr + 1:]:
        if len(r) <= col["Instructions Executed"] or not r[0].startswith("0x"):
            continue
        src = r[col["Source"]].strip()
        if src.startswith("@"):
            src = src.split(None, 1)[1]
        op = src.split()[0].split(".")[0].rstrip(";")
        ops[op] += int(r[col["Instructions Executed"]] or 0)
        stalls[op] += int(r[col["Warp Stall Sampling (All Samples)"]] or 0)
    total = sum(ops.values())
    packed = sum(v for k, v in ops.items() if k in ("FFMA2", "FADD2", "FMUL2"))
    per_line = total / lines
    p_line = packed / lines
    print(f"kernel: {rows[0][1]}")
    print(f"warp instructions executed: {total} = {per_line:.1f} per line ({per_line / wpl:.1f} per warp and line); packed f32x2: {p_line:.1f} per line")
    print("per line: " + " ".join(f"{k}:{v / lines:.1f}" for k, v in ops.most_common(24)))
    dispatch = 2 * p_line + (per_line - p_line)
    print(f"dispatch-port model: 2 x {p_line:.1f} + {per_line - p_line:.1f} = {dispatch:.0f} dispatch cycles per line")
    if ms is not None:
        # 148 SMs x 4 schedulers share the lines evenly (persistent grid)
        clock_ghz = float(sys.argv[5]) if len(sys.argv) > 5 else 1.965
        cyc = ms * 1e-3 * clock_ghz * 1e9 / (lines / (148 * 4))
        print(f"measured: {ms} ms at {clock_ghz} GHz = {cyc:.0f} cycles per line and scheduler -> the kernel runs at {dispatch / cyc:.3f} of its dispatch bound")
    s = sum(stalls.values())
    print("stall samples by opcode: " + " ".join(f"{k}:{100.0 * v / s:.1f}%" for k, v in stalls.most_common(10)))


if __name__ == "__main__":
    main()
