#!/usr/bin/env python
"""GPU-side probe of the fused tile pass: time qipb_apply_fused on hand-made gate lists so that the
cost of a pass decomposes into (tile load/store pipeline) + (per-gate sweeps).

    python scripts/fused_probe.py [--qubits 31] [--c64] [--reps 3]

Prints one line per case: ms per launch and algorithmic GB/s (2 * sizeof(amp) * 2^n / time)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=31)
    ap.add_argument("--c64", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cases", default="")
    ap.add_argument("--low", type=int, default=7, help="contiguous low tile bits of the scattered tile (7 -> 2^12 tile, 6 -> 2^11)")
    args = ap.parse_args()
    import torch
    from qip_b200 import B200Backend
    from qip_b200.circuits import H2, haar_unitary
    from qip_b200.ops import BitGate, Pass

    n = args.qubits
    dt = np.complex64 if args.c64 else np.complex128
    b = B200Backend.make_state(n, [], [], statetype=dt)
    rng = np.random.default_rng(5)
    amp = 8 if args.c64 else 16
    nbytes = 2.0 * amp * 2.0 ** n

    def u4():
        return haar_unitary(rng, 4)

    def u2():
        return haar_unitary(rng, 2)

    hi = [n - 10, n - 8, n - 5, n - 3, n - 1]
    tile_a = tuple(range(args.low)) + tuple(hi)          # the planner's usual shape: 7 low + 5 scattered high
    tile_c = tuple(range(args.low + 5))                  # contiguous tile
    ph = np.diag([np.exp(0.3j)])

    def dense2(bits):
        return BitGate("matrix", tuple(bits), 0, u4(), False)

    def dense1(bit):
        return BitGate("matrix", (bit,), 0, u2(), False)

    cases = []
    cases.append(("floor: k=0 phase under 3 outside controls", tile_a,
                  [BitGate("matrix", (), (1 << (n - 2)) | (1 << (n - 4)) | (1 << (n - 6)), ph, True)]))
    cases.append(("floor contiguous tile", tile_c,
                  [BitGate("matrix", (), (1 << (n - 2)) | (1 << (n - 4)) | (1 << (n - 6)), ph, True)]))
    for g in (1, 2, 4, 8):
        cases.append(("%d x dense2 on high tile bits" % g, tile_a,
                      [dense2((hi[(2 * i) % 5], hi[(2 * i + 1) % 5])) for i in range(g)]))
    for g in (1, 4):
        cases.append(("%d x dense2 on low bits (0,1)/(2,3)" % g, tile_a,
                      [dense2(((2 * i) % 4 + 1, (2 * i) % 4)) for i in range(g)]))
    for g in (1, 4):
        cases.append(("%d x dense2 mixed (low, high)" % g, tile_a,
                      [dense2((hi[i % 5], 3 + i % 4)) for i in range(g)]))
    for g in (1, 4, 8):
        cases.append(("%d x dense1 on high tile bits" % g, tile_a, [dense1(hi[i % 5]) for i in range(g)]))
    cases.append(("4 x dense1 on low bits", tile_a, [dense1(i) for i in range(4)]))
    cases.append(("layer-like: 3 dense2 + 2 phases", tile_a,
                  [dense2((hi[0], hi[1])), dense2((hi[2], 5)), dense2((hi[3], hi[4])),
                   BitGate("matrix", (), 1 << 9, ph, True), BitGate("matrix", (), 1 << (n - 7), ph, True)]))
    # a QFT step: H on a tile bit + controlled phases from every other bit
    def qft_step(t):
        out = [BitGate("matrix", (t,), 0, np.asarray(H2, dtype=np.complex128), False)]
        for c in range(n):
            if c != t and c < t:
                out.append(BitGate("matrix", (), (1 << c) | (1 << t), np.diag([np.exp(2j * np.pi / 2.0 ** (1 + t - c))]), True))
        return out
    cases.append(("QFT-like: 5 x (H + controlled phases)", tile_a, sum([qft_step(t) for t in hi[::-1]], [])))

    sel = [s for s in args.cases.split(",") if s]
    torch.cuda.synchronize()
    for name, tile, gates in cases:
        if sel and not any(s in name for s in sel):
            continue
        p = Pass(True, gates, tile)
        with torch.cuda.device(b.device):
            b._stream()
            b._launch_fused(p)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                b._launch_fused(p)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        print("%-52s %8.2f ms  %7.0f GB/s" % (name, ms, nbytes / ms / 1e6), flush=True)
    # reference point: the un-fused register kernel on one high bit
    b.fuse = False
    q = n - 1 - hi[0]
    b.kronselect_dot({q: u2()})
    b.flush()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        b.kronselect_dot({q: u2()})
        b.flush()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    print("%-52s %8.2f ms  %7.0f GB/s" % ("un-fused gate_kernel<K=1> (reference point)", ms, nbytes / ms / 1e6))
    return 0


if __name__ == "__main__":
    sys.exit(main())
