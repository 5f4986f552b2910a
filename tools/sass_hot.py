"""Summarise an `ncu --page source --csv --print-source sass` dump: opcode mix and hot instructions.

usage: python tools/sass_hot.py src.csv [threshold_pct] [kernel_index]
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = list(csv.reader(open(path)))
    # split per kernel ("Kernel Name" rows)
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    b = blocks[which]
    hdr, data = b["hdr"], b["data"]
    ia, ie, it, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed"), hdr.index("# Samples")
    tot = sum(int(r[ie]) for r in data)
    print(b["name"][:100])
    print("total warp instructions", tot, "sass rows", len(data))
    by, samp = collections.Counter(), collections.Counter()
    for r in data:
        toks = r[ia].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        by[op] += int(r[ie])
        samp[op] += int(r[isamp])
    ts = max(1, sum(samp.values()))
    for op, c in by.most_common(28):
        print("%-10s %6.2f%% instr  %6.2f%% samples" % (op, 100 * c / tot, 100 * samp[op] / ts))
    print("---- instructions above %.2f%% of executed" % thr)
    for i, r in enumerate(data):
        if int(r[ie]) > thr / 100 * tot:
            print("%5d %11s thr %5s smp %6s  %s" % (i, r[ie], r[it], r[isamp], r[ia][:100]))


if __name__ == "__main__":
    main()
