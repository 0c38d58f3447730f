#!/usr/bin/env python
"""Host-only statistics of the planners (no GPU): what the fused-pass planner and the shard scheduler produce for the
BASELINE workloads, with and without the host-side changes that were made after round 1's last GPU call.
Cost units: one fused pass over the state / shard = 1; a g-bit remap = 2.6 * (1 - 2^-g) (shardplan._move_cost, from the
measured 60-80 ms per pass and 100 / 155 ms per 1-bit / 3-bit exchange at 33 local qubits).

    python scripts/planner_stats.py > profiles/r01_planner_stats.txt"""
import collections
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qip_b200 import ops, shardplan as sp                      # noqa: E402
from qip_b200.circuits import layered_stream, qfft_stream       # noqa: E402


def logical(stream, n):
    out = []
    for mats in stream:
        for g in ops.decode_mats(mats, n):
            s = ops.simplify(g)
            if s is not None:
                out.append(s)
    return out


def single_gpu(n=33, layers=40):
    print("== one GPU, layered circuit at %d qubits, %d layers (seeds 33..): fused-pass planner ==" % (n, layers))
    for pack in ("0", "1"):
        os.environ["QIPB_PACK_1Q"] = pack
        pos = list(range(n - 1, -1, -1))
        passes = sweeps = lone1 = real2 = 0
        cost = 0.0
        for seed in range(33, 33 + layers):
            gates = []
            for g in logical(layered_stream(n, 1, seed), n):
                if g.kind == "swap" and not g.controls:            # relabel, like B200Backend._relabel
                    a, b = g.targets
                    pos[a], pos[b] = pos[b], pos[a]
                    continue
                gates.append(ops.Gate(g.kind, tuple(n - 1 - pos[q] for q in g.targets),
                                      tuple(n - 1 - pos[q] for q in g.controls), g.mat, g.diagonal))
            ps, _ = ops.plan(gates, n, 16)
            passes += len(ps)
            cost += sum(ops.pass_cost(p, n, 16) for p in ps) / (32 * 2.0 ** n)
            for p in ps:
                for g in p.gates:
                    dense = g.kind == "swap" or not (g.diagonal or g.k == 0)
                    sweeps += 1
                    lone1 += dense and g.k == 1 and g.ctrl_mask == 0
                    real2 += dense and g.kind == "matrix" and g.k == 2 and not g.mat.imag.any()
        print("QIPB_PACK_1Q=%s: %.2f passes/layer, %.1f gate sweeps/layer, %.2f lone dense 1-qubit sweeps/layer, "
              "%.2f real 2-qubit blocks/layer, model cost %.3f sweeps/layer"
              % (pack, passes / layers, sweeps / layers, lone1 / layers, real2 / layers, cost / layers))
    os.environ["QIPB_PACK_1Q"] = "1"


def sharded(layers=40):
    print("\n== sharded, 33 qubits per GPU: shard scheduler + rank-local planner (rank 0) ==")
    for hoist in ("0", "1"):
        os.environ["QIPB_SHARD_HOIST"] = hoist
        for G in (1, 2, 3):
            n, nl = 33 + G, 33

            def plan_local(batch):
                return ops.plan_passes(ops.merge_bitgates(batch, 2), nl, 16)
            lay = sp.Layout(n, G)
            tot = collections.Counter()
            for s in range(layers):
                gates = logical(layered_stream(n, 1, 33 + s), n)
                prog = sp.compile_program(sp.schedule(gates, lay, count_passes=lambda b: len(plan_local(b))), nl, 0, plan_local)
                for st in prog:
                    if isinstance(st, tuple):
                        tot["passes"] += len(st[1])
                    else:
                        g = len(st.pairs) if isinstance(st, sp.MultiExchange) else 1
                        tot["moves"] += 1
                        tot["move_cost"] += sp._move_cost(g)
            print("layered %d qubits on %d GPUs, QIPB_SHARD_HOIST=%s: %.2f passes + %.2f moves per layer = %.2f cost units"
                  % (n, 1 << G, hoist, tot["passes"] / layers, tot["moves"] / layers, (tot["passes"] + tot["move_cost"]) / layers))
        for G in (1, 2, 3):
            n, nl = 33 + G, 33

            def plan_local(batch):
                return ops.plan_passes(ops.merge_bitgates(batch, 2), nl, 16)
            gates = logical(qfft_stream(n), n)
            lay = sp.Layout(n, G)
            sp.choose_initial_layout(gates, lay)
            prog = sp.compile_program(sp.schedule(gates, lay, count_passes=lambda b: len(plan_local(b))), nl, 0, plan_local)
            shape = []
            for st in prog:
                if isinstance(st, tuple):
                    shape.append("%d passes" % len(st[1]))
                else:
                    shape.append("%d-bit exchange" % (len(st.pairs) if isinstance(st, sp.MultiExchange) else 1))
            print("QFFT %d qubits on %d GPUs, QIPB_SHARD_HOIST=%s: %s" % (n, 1 << G, hoist, " -> ".join(shape)))
    os.environ["QIPB_SHARD_HOIST"] = "1"


if __name__ == "__main__":
    single_gpu()
    sharded()
