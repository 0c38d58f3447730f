#!/usr/bin/env python
"""Per-opcode and per-instruction warp-stall samples of the first kernel in an .ncu-rep (SASS view).
Usage: python scripts/ncu_sass_hot.py gpurun_out/prof.ncu-rep [kernel index]"""
import collections
import csv
import io
import subprocess
import sys


def main(path, which=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    k, data, names, hdr = -1, [], [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            k += 1
            data.append([])
            names.append(r[1])
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if k >= 0 and len(r) > 5:
            data[k].append(r)
    d = data[which]
    si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    tot = sum(float(r[si]) for r in d) or 1.0
    byop, cnt = collections.Counter(), collections.Counter()
    for r in d:
        toks = r[src].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        byop[op] += float(r[si])
        cnt[op] += float(r[ie])
    ninst = sum(cnt.values()) or 1.0
    print("kernel:", names[which])
    print("total warp-stall samples %d, warp instructions executed %.3g" % (tot, ninst))
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not" not in h]
    agg = collections.Counter()
    for r in d:
        for i, h in stall_cols:
            agg[h] += float(r[i] or 0)
    print("stall reasons: " + ", ".join("%s %.1f%%" % (h, 100 * v / tot) for h, v in agg.most_common(8)))
    print("-- by opcode (samples%, instructions%)")
    for op, v in byop.most_common(14):
        print("   %-26s %5.1f%%  %5.1f%%" % (op, 100 * v / tot, 100 * cnt[op] / ninst))
    print("-- hottest instructions")
    for r in sorted(d, key=lambda r: -float(r[si]))[:14]:
        st = sorted(((h, float(r[i] or 0)) for i, h in stall_cols), key=lambda kv: -kv[1])[:2]
        print("   %5.1f%%  %-58s %s" % (100 * float(r[si]) / tot, r[src].strip()[:58], ["%s=%d" % x for x in st if x[1] > 0]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
