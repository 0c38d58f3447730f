#!/usr/bin/env python
"""GPU-side: where does B200Backend.make_state spend its time at full size?  (bench.py's e2e leg)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from qip_b200 import B200Backend
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 33
    rng = np.random.default_rng(1)
    groups = [[q] for q in range(n)]
    feeds = [np.array([np.cos(t), np.sin(t)], dtype=np.complex128) for t in rng.uniform(0, np.pi, size=n)]
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b = B200Backend(n, np.complex128)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b._init_state(groups, feeds)
        e1.record()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        p = b.measure_probabilities(np.array([0, n // 2, n - 1], dtype=np.int32))
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        b.close()
        del b
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        print("rep %d: ctor %.1f ms, _init_state wall %.1f ms (device %.1f ms), measure %.1f ms, close %.1f ms, sum p = %.15f"
              % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), e0.elapsed_time(e1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), float(p.sum())))
    # the exact sequence of bench.py's e2e leg: make_state -> one layer -> measure_probabilities -> close
    from qip_b200.circuits import layered_stream
    for rep in range(3):
        ops = list(layered_stream(n, 1, 40 + rep))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b = B200Backend.make_state(n, groups, feeds, statetype=np.complex128)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        for mats in ops:
            b.kronselect_dot(mats)
        t3 = time.perf_counter()
        b.flush()
        t4 = time.perf_counter()
        torch.cuda.synchronize()
        t5 = time.perf_counter()
        b.measure_probabilities(np.array([0, n // 2, n - 1], dtype=np.int32))
        t6 = time.perf_counter()
        b.close()
        t7 = time.perf_counter()
        print("e2e rep %d: make_state host %.1f ms + wait %.1f ms | queue %.1f ms, plan+launch %.1f ms, device wait %.1f ms | measure %.1f ms | close %.1f ms | total %.1f ms"
              % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4), 1e3 * (t6 - t5), 1e3 * (t7 - t6), 1e3 * (t7 - t0)))
    # one 33-qubit group fed from the host is impossible (128 GiB); a few wide groups instead
    groups = [list(range(0, 11)), list(range(11, 22)), list(range(22, n))]
    feeds = []
    for g in groups:
        v = rng.normal(size=2 ** len(g)) + 1j * rng.normal(size=2 ** len(g))
        feeds.append(v / np.linalg.norm(v))
    b = B200Backend(n, np.complex128)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    b._init_state(groups, feeds)
    e1.record()
    torch.cuda.synchronize()
    print("three wide groups: _init_state device %.1f ms (%.0f GB/s written), total_prob %.15f"
          % (e0.elapsed_time(e1), 16 * 2.0 ** n / e0.elapsed_time(e1) / 1e6, b.total_prob()))
    return 0


if __name__ == "__main__":
    sys.exit(main())
