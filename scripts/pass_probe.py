#!/usr/bin/env python
"""Per-launch anatomy of the fused passes of a workload: what every launch carries (QIPB_DEBUG line of the lowering) next
to its device time.  Usage: python scripts/pass_probe.py [--workload layered|qft] [--qubits 33] [--steps 4]"""
import argparse
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(args):
    import numpy as np
    import torch
    from qip_b200 import B200Backend
    from qip_b200.circuits import layered_stream, qfft_stream
    n = args.qubits
    b = B200Backend.make_state(n, [], [])
    steps = [list(qfft_stream(n)) if args.workload == "qft" else list(layered_stream(n, 1, 33 + s)) for s in range(args.steps + 1)]
    for m in steps[0]:
        b.kronselect_dot(m)
    b.flush()
    torch.cuda.synchronize()
    sys.stderr.write("[probe] begin\n")
    sys.stderr.flush()
    b.profile = []
    for s in range(1, args.steps + 1):
        for m in steps[s]:
            b.kronselect_dot(m)
        b.flush()
    torch.cuda.synchronize()
    for name, nbytes, e0, e1 in b.profile:
        print("T %s %.3f" % (name, e0.elapsed_time(e1)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="layered")
    ap.add_argument("--qubits", type=int, default=33)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return child(args)
    env = dict(os.environ, QIPB_DEBUG="1")
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--workload", args.workload, "--qubits", str(args.qubits),
                        "--steps", str(args.steps)], capture_output=True, text=True, env=env)
    err = r.stderr.split("[probe] begin\n")[-1]
    launches = [ln for ln in err.splitlines() if ln.startswith("[qipb] fused launch")]
    times = [ln.split() for ln in r.stdout.splitlines() if ln.startswith("T ")]
    ft = [float(t[2]) for t in times if t[1].startswith("fused_kernel")]
    floor = 2.0 * 16 * 2.0 ** args.qubits / 6556.8e9 * 1e3
    print("%s at %d qubits: %d fused launches, %.1f ms total, floor %.1f ms/launch (measured copy peak)" % (args.workload, args.qubits, len(ft), sum(ft), floor))
    for ln, t in zip(launches, ft):
        print("%7.2f ms (%.2f x floor)  %s" % (t, t / floor, ln.replace("[qipb] fused launch: ", "")))
    if len(launches) != len(ft):
        print("(launch lines %d != timed fused launches %d)" % (len(launches), len(ft)))
        print(r.stderr[-2000:])


if __name__ == "__main__":
    main()
