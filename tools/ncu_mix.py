#!/usr/bin/env python
"""Instruction mix / shared-memory wavefronts per opcode from `ncu --page source --csv`.

    ncu -i report.ncu-rep --page source --csv > src.csv ; python tools/ncu_mix.py src.csv [top]
"""
import csv
import sys
from collections import defaultdict


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    hdr = rows[1]
    ix = {k: i for i, k in enumerate(hdr)}
    cnt, wf, ideal, smp = defaultdict(int), defaultdict(int), defaultdict(int), defaultdict(int)
    total = 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        src = r[ix["Source"]].strip()
        parts = src.split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") else parts[0]
        op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "LD.", "ST.")) else op.split(".")[0]
        n = int(float(r[ix["Instructions Executed"]] or 0))
        cnt[op] += n
        total += n
        wf[op] += int(float(r[ix["L1 Wavefronts Shared"]] or 0))
        ideal[op] += int(float(r[ix["L1 Wavefronts Shared Ideal"]] or 0))
        smp[op] += int(float(r[ix["# Samples"]] or 0))
    print("total warp instr %d, static instructions %d" % (total, len(rows) - 2))
    for op, n in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]:
        print("%-12s %12d %5.1f%%  smem_wavefronts=%10d ideal=%10d samples=%d" % (op, n, 100.0 * n / total, wf[op], ideal[op], smp[op]))


if __name__ == "__main__":
    main()
