#!/usr/bin/env python
"""GPU-side probe (round 2): does a chunk of the state stay L2-resident between fused passes?

A layer of the layered benchmark needs ~4.5 fused passes because a pass can only make 12 index bits tile bits, and
every pass is one HBM round trip of the whole state (45 ms at 33 qubits).  B200 has 126 MB of L2: if a contiguous chunk
of 2^L amplitudes (L = 21: 32 MiB, L = 22: 64 MiB complex128) is processed by k passes back to back before the next
chunk is touched, passes 2..k should be served from L2 and the k passes together should cost about ONE HBM round trip.
That would turn "7 low bits + 5 free bits per HBM pass" into "7 + 5k bits per HBM pass" for gates on the low L bits.

No kernel change is needed to try it: qipb_apply_fused is called per chunk on `state + chunk offset` with nbits = L.
The probe times, for L in 20..24 and k in 1..4:   (all chunks) x (k passes per chunk)   against   k full-state passes.

    python scripts/l2_block_probe.py [--qubits 28] [--reps 3]

(28 qubits = 4 GiB keeps the python launch loop short; the HBM / L2 behaviour per chunk is the same as at 33.)"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=28)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    from qip_b200 import B200Backend
    from qip_b200.backend import pack_pass
    from qip_b200.circuits import haar_unitary
    from qip_b200.ops import BitGate, Pass

    n = args.qubits
    b = B200Backend.make_state(n, [], [])
    rng = np.random.default_rng(1)
    L_, ctx, code = b.L, b.ctx, b.code
    base_ptr = b.state.data_ptr()

    def make_pass(L, which):
        """two dense 2-qubit blocks on tile bits: 7 low bits + 5 high bits of the chunk (a different set per pass)"""
        hi = [7 + ((5 * which + j) % (L - 7)) for j in range(5)]
        tile = tuple(sorted(set(range(7)) | set(hi)))
        while len(tile) < 12:
            tile = tuple(sorted(set(tile) | {max(tile) + 1 if max(tile) + 1 < L else min(set(range(L)) - set(tile))}))
        gates = [BitGate("matrix", (tile[8], tile[3]), 0, haar_unitary(rng, 4), False),
                 BitGate("matrix", (tile[10], tile[9]), 0, haar_unitary(rng, 4), False)]
        return Pass(True, gates, tile)

    def run(L, k, chunked):
        passes = [make_pass(L, w) for w in range(k)]
        packed = [pack_pass(p) for p in passes]
        nchunks = 1 << (n - L) if chunked else 1
        nb = L if chunked else n
        def launch_all():
            if chunked:
                for s in range(nchunks):
                    ptr = ctypes.c_void_p(base_ptr + s * (16 << L))
                    for p, (arr, tb) in zip(passes, packed):
                        rc = L_.qipb_apply_fused(ctx, ptr, nb, code, len(p.tile_bits), tb, len(p.gates), arr)
                        assert rc == 0, L_.qipb_last_error()
            else:
                for p, (arr, tb) in zip(passes, packed):
                    rc = L_.qipb_apply_fused(ctx, ctypes.c_void_p(base_ptr), nb, code, len(p.tile_bits), tb, len(p.gates), arr)
                    assert rc == 0, L_.qipb_last_error()

        # the python launch loop (~30 us per call) would hide what is being measured: capture the launches into a CUDA
        # graph once (the library launches on the stream it is given) and time replays of the graph
        b._stream()
        launch_all()                                   # module load, attributes
        torch.cuda.synchronize()
        graph = None
        try:
            side = torch.cuda.Stream()
            with torch.cuda.stream(side):
                b._stream()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side, capture_error_mode="relaxed"):
                    b._stream()
                    launch_all()
                graph = g
        except Exception as exc:                       # noqa: BLE001 -- fall back to direct launches, say so
            print("graph capture failed (%s): timing direct launches, host overhead included" % exc)
        b._stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for rep in range(args.reps + 1):
            torch.cuda.synchronize()
            e0.record()
            if graph is not None:
                graph.replay()
            else:
                launch_all()
            e1.record()
            torch.cuda.synchronize()
            if rep:
                best = min(best, e0.elapsed_time(e1))
        return best

    full = {k: run(21, k, False) for k in (1, 2, 3, 4)}
    print("full-state passes at %d qubits: " % n + "  ".join("k=%d %.2f ms" % (k, v) for k, v in full.items()))
    for L in (20, 21, 22, 23, 24):
        if L >= n:
            continue
        row = []
        for k in (1, 2, 3, 4):
            t = run(L, k, True)
            row.append("k=%d %.2f ms (%.2fx of k full passes, %.2fx of ONE)" % (k, t, t / full[k], t / full[1]))
        print("chunks of 2^%d (%d MiB, %d launches per pass): " % (L, (16 << L) >> 20, 1 << (n - L)) + "  ".join(row))
    b.close()


if __name__ == "__main__":
    main()
