"""configs[2] (Grover, 27 + 1 qubits, K = 16 re-fed iterations) with per-phase wall-clock times: where an iteration's
milliseconds go (host graph replay, fused H layers, func_xor, the final soft_measure).  python scripts/grover_probe.py [repeats] [adopt 0|1]"""
import os
import sys
import time
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qip_b200 import B200Backend                              # noqa: E402
from qip_b200.circuits import H2                              # noqa: E402
from qip_b200.functions import equals, tabulated              # noqa: E402
from qip_b200.graph import CompiledCircuit                    # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    adopt = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
    ns, x0, K = 27, 42, 16
    n = ns + 1
    search, anc = list(range(ns)), [ns]
    h_all = {i: H2 for i in search}
    gops = [("f", search, anc, tabulated(equals(x0), ns)), ("k", h_all), ("f", search, anc, tabulated(equals(0), ns)), ("k", h_all)]
    plus = np.array([1.0, 1.0]) / np.sqrt(2.0)
    first = CompiledCircuit.from_ops(n, [[q] for q in search] + [anc], [plus] * ns + [[1 / np.sqrt(2), -1 / np.sqrt(2)]], gops)
    again = None
    sync = torch.cuda.synchronize
    for rep in range(reps):
        sync()
        t0 = time.perf_counter()
        state, _ = first.run(device_state=True)
        sync()
        t1 = time.perf_counter()
        if again is None:
            again = CompiledCircuit.from_ops(n, [search + anc], [state], gops)
        its = []
        for _ in range(K - 1):
            ta = time.perf_counter()
            state, _ = again.run(feed={(0,): state}, device_state=True, adopt_feed=adopt)
            sync()
            its.append(1e3 * (time.perf_counter() - ta))
        t2 = time.perf_counter()
        g = B200Backend.make_state(n, [search + anc], [state], adopt_feed=adopt)
        _, p = g.soft_measure(np.array(search, dtype=np.int32), measured=x0)
        g.close()
        sync()
        t3 = time.perf_counter()
        print("rep %d: first.run %.1f ms, 15 re-fed runs %.1f ms (min %.1f, median %.1f, max %.1f), soft_measure %.1f ms, total %.1f ms = %.1f ms/iteration, p=%.6e"
              % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), min(its), float(np.median(its)), max(its), 1e3 * (t3 - t2), 1e3 * (t3 - t0), 1e3 * (t3 - t0) / K, p), flush=True)
        if rep == 1:
            torch.cuda.empty_cache()                         # what bench.py's earlier blocks leave behind: a cold allocator
            print("(allocator cache emptied)")


if __name__ == "__main__":
    main()
