#!/usr/bin/env python
"""One GPU: what does a fused pass lose when the persistent chunk-remap kernel runs beside it?  The "peer" of the remap is a
second buffer on the SAME GPU, so the experiment separates SM / HBM contention from anything NVLink does.
   python scripts/corun_probe.py"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from qip_b200 import B200Backend, lib as qlib
    from qip_b200.circuits import layered_stream
    from qip_b200.ops import decode_mats, simplify, plan
    n = 31
    b = B200Backend.make_state(n, [], [])
    L = b.L
    gates = []
    for mats in layered_stream(n, 1, 33):
        for g in decode_mats(mats, n):
            s = simplify(g)
            if s is not None:
                gates.append(s)
    passes, _ = plan(gates, n, 16, strategy="tile")
    passes = [p for p in passes if p.fused]
    nb = 31
    B = torch.zeros(2 ** nb, dtype=torch.complex128, device="cuda")
    C = torch.zeros(2 ** nb, dtype=torch.complex128, device="cuda")
    cs = torch.cuda.Stream(priority=-1)
    xs = torch.cuda.Stream(priority=0)

    def remap(ctas):
        peers = (ctypes.c_void_p * 8)()
        peers[1] = ctypes.c_void_p(C.data_ptr())
        qlib.check(L.qipb_peer_remap_chunk(b.ctx, ctypes.c_void_p(B.data_ptr()), peers, nb, qlib.C128, 1, qlib.int_array([nb - 1]), 0,
                                           0, qlib.int_array([]), 0, ctas))

    def timed(fn_c, fn_x, reps=3):
        out = []
        for _ in range(reps):
            torch.cuda.synchronize()
            ec0, ec1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ex0, ex1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if fn_x:
                with torch.cuda.stream(xs):
                    b._stream()
                    ex0.record()
                    fn_x()
                    ex1.record()
            if fn_c:
                with torch.cuda.stream(cs):
                    b._stream()
                    ec0.record()
                    fn_c()
                    ec1.record()
            torch.cuda.synchronize()
            out.append((ec0.elapsed_time(ec1) if fn_c else 0.0, ex0.elapsed_time(ex1) if fn_x else 0.0))
        return min(o[0] for o in out), min(o[1] for o in out)

    def run_passes():
        for p in passes:
            b._launch_fused(p)

    print("%d fused passes at %d qubits; remap of 2^%d amplitudes against a local buffer" % (len(passes), n, nb))
    print("passes alone             : %.2f ms" % timed(run_passes, None)[0])
    for ctas in (148, 296, 74, 0):
        print("remap alone  (ctas %4d)  : %.2f ms" % (ctas, timed(None, lambda: remap(ctas))[1]))
        c, x = timed(run_passes, lambda: remap(ctas))
        print("together     (ctas %4d)  : passes %.2f ms, remap %.2f ms" % (ctas, c, x))
    b.close()


if __name__ == "__main__":
    main()
