#!/usr/bin/env python
"""Aggregates an `ncu --set full --import-source on` report into basic blocks: share of executed warp instructions,
share of stall samples, average active lanes, memory/sync ops per block.
  ncu -i report.ncu-rep --page source --csv --print-source sass > sass.csv ; python scripts/ncu_sass_blocks.py sass.csv"""
import csv
import sys


def main(path, min_share=0.3):
    rows = list(csv.reader(open(path)))
    hdr, data = rows[1], rows[2:]
    i_src, i_exec = hdr.index("Source"), hdr.index("Instructions Executed")
    i_thr, i_smp = hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    tot = sum(int(r[i_exec]) for r in data)
    tots = sum(int(r[i_smp]) for r in data)
    blocks, cur = [], None
    for k, r in enumerate(data):
        e = int(r[i_exec])
        if cur is None or e != cur["e"]:
            cur = dict(e=e, start=k, n=0, thr=0, smp=0, ops=[])
            blocks.append(cur)
        cur["n"] += 1
        cur["thr"] += int(r[i_thr])
        cur["smp"] += int(r[i_smp])
        tok = r[i_src].split()
        op = tok[1] if tok[0].startswith("@") else tok[0]
        if any(x in op for x in ("LDG", "LDS", "STS", "ATOM", "BAR", "SHFL", "VOTE", "RED", "F2I", "FLO", "POPC")):
            cur["ops"].append(op.split(".")[0])
    print("# %s: %d SASS instructions, %d warp instructions executed, %d stall samples" % (rows[0][1][:60], len(data), tot, tots))
    print("# sass range    n  instr%%   cum%%  stall-samples%%  lanes  ops")
    cum = 0.0
    for b in blocks:
        share = 100.0 * b["e"] * b["n"] / tot
        cum += share
        if share < min_share:
            continue
        ops = {}
        for o in b["ops"]:
            ops[o] = ops.get(o, 0) + 1
        print("%4d-%-4d  %4d  %6.2f  %6.1f  %8.2f        %5.1f  %s" % (
            b["start"], b["start"] + b["n"] - 1, b["n"], share, cum, 100.0 * b["smp"] / tots,
            b["thr"] / float(b["e"] * b["n"]) if b["e"] else 0.0, " ".join("%s x%d" % kv for kv in sorted(ops.items()))))


if __name__ == "__main__":
    main(sys.argv[1])
