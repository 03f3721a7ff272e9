#!/usr/bin/env python
"""Basic-block execution breakdown from an .ncu-rep source page (SASS).
usage: python profiles/sass_blocks.py rep [min_share_pct]"""
import csv, subprocess, sys

def main():
    rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ie = hdr.index('Instructions Executed'); at = hdr.index('Avg. Threads Executed'); src = hdr.index('Source')
    data = [(float(r[ie]), float(r[at]), r[src].strip()) for r in rows[2:] if len(r) > at]
    tot = sum(d[0] for d in data)
    print("sass instrs", len(data), "total warp-inst %.4e" % tot, "avg lanes %.2f" % (sum(d[0] * d[1] for d in data) / tot))
    start = 0
    for i in range(1, len(data) + 1):
        if i == len(data) or data[i][0] != data[start][0] or data[i][1] != data[start][1]:
            n = i - start; c, a = data[start][0], data[start][1]
            share = 100 * n * c / tot
            if share > thr:
                ops = {}
                for d in data[start:i]:
                    t = d[2].split()
                    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
                    ops[op] = ops.get(op, 0) + 1
                top = sorted(ops.items(), key=lambda x: -x[1])[:7]
                print("blk @%4d n=%3d exec=%.3e lanes=%5.1f share=%5.1f%% %s" % (start, n, c, a, share, top))
            start = i

if __name__ == "__main__":
    main()
