#!/usr/bin/env python
"""Turn an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` log of bench.py into
profiles/dram_traffic.json: mean DRAM bytes per launch per kernel (what bench.py reports as roofline.traffic).
Usage: python scripts/ncu_traffic.py gpurun_out/traffic.csv <qubits_per_gpu>"""
import collections
import csv
import json
import os
import sys


def short(name):
    if "fused_kernel" in name:
        return "fused_kernel"
    return name.split("(")[0].replace("void ", "").replace("qipb::", "")


def main(path, qubits):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for r in rows[1:]:
        if r[mi] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot[short(r[ki])] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
            if r[mi] == "dram__bytes_read.sum":
                cnt[short(r[ki])] += 1
    out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "dram_traffic.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for k in tot:
        out["%s@%d" % (k, qubits)] = {"bytes_per_launch": tot[k] / cnt[k], "launches": cnt[k],
                                       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on bench.py"}
    json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
    for k in sorted(tot):
        print("%-40s %3d launches  %.3f GB/launch" % (k, cnt[k], tot[k] / cnt[k] / 1e9))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
