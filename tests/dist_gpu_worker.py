"""Multi-GPU parity worker (launched by tests/test_gpu_sharded.py with torch.distributed.run, one rank
per GPU, NCCL): ShardedB200Backend vs the CPU oracle on the same seeded circuits."""
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as orc                      # noqa: E402
from qip_b200.circuits import H2, X2, haar_unitary, layered_stream, qfft_stream, rm_mat   # noqa: E402
from qip_b200.mats import CMat, SwapMat                # noqa: E402
from qip_b200.sharded import ShardedB200Backend        # noqa: E402


REPORT = {"cases": 0, "max_err": 0.0}


def check(name, got, want, tol=1e-12):
    """max |got - want| / max |want| <= tol, and -- for amplitudes above 1e-3 of the largest -- the elementwise
    RELATIVE error (the tolerance north_star states) <= 10 * tol."""
    got, want = np.asarray(got), np.asarray(want)
    big = max(1e-300, float(np.max(np.abs(want))))
    err = float(np.max(np.abs(got - want))) / big
    assert err <= tol, (name, err)
    sel = np.abs(want) > 1e-3 * big
    if np.any(sel):
        rel = float(np.max(np.abs(got[sel] - want[sel]) / np.abs(want[sel])))
        assert rel <= 10 * tol, (name, "elementwise relative", rel)
        err = max(err, rel / 10)
    REPORT["cases"] += 1
    REPORT["max_err"] = max(REPORT["max_err"], err)
    return err


def run_cases(quick=False, log=print):
    """All cases of the multi-GPU parity tier; torch.distributed (NCCL, one rank per GPU) must be initialised.
    quick: the subset bench.py runs before its timed region at N > 1 (every action kind of the shard scheduler,
    measurement across shards, func_apply, range access, the production-size QFFT closed form); the reduce_measure
    / compiled-circuit / lazy-init cases stay with the pytest tier.  Returns {"cases": checks passed, "max_err": ...}."""
    REPORT["cases"], REPORT["max_err"] = 0, 0.0
    rank, world = dist.get_rank(), dist.get_world_size()
    # the exchange / compute pipeline is on for large shards only; the parity tier runs it at every size
    saved = os.environ.get("QIPB_OVERLAP_MIN_BYTES")
    os.environ["QIPB_OVERLAP_MIN_BYTES"] = "0"
    try:
        return _run_cases(quick, log, rank, world)
    finally:
        if saved is None:
            os.environ.pop("QIPB_OVERLAP_MIN_BYTES", None)
        else:
            os.environ["QIPB_OVERLAP_MIN_BYTES"] = saved


def _run_cases(quick, log, rank, world):
    n = 12
    rng = np.random.default_rng(5)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    groups, feeds = [list(range(n))], [psi]
    rngu = np.random.default_rng(6)
    extra = [{(0, 5): CMat(X2)}, {(1, 0, 6): CMat(CMat(haar_unitary(rngu, 2)))}, {0: rm_mat(3)},
             {(1, 0): CMat(rm_mat(2))}, {(0, 7): SwapMat(1)}, {(0, 1): haar_unitary(rngu, 4)},
             {(11, 0, 3): CMat(SwapMat(1))}, {0: H2}, {(2, 9, 0): haar_unitary(rngu, 8)},
             {(2, 9, 0, 4, 7, 1): haar_unitary(rngu, 64)}]   # six qubits, three of them global at P = 8: the batched dense kernel on a shard
    cases = {"layered": list(layered_stream(n, 3, 2)), "qfft": list(qfft_stream(n)), "mixed": extra,
             "layered+qfft": list(layered_stream(n, 1, 9)) + list(qfft_stream(n))}
    if quick:
        cases = {"layered": cases["layered"], "mixed": extra, "layered+qfft": cases["layered+qfft"]}
    for name, ops_ in cases.items():
        for fuse, peer in (((True, False), (True, True)) if quick else ((True, False), (False, False), (True, True))):
            g = ShardedB200Backend.make_state(n, groups, feeds, statetype=np.complex128, fuse=fuse, peer_gates=peer)
            c = orc.OracleBackend.make_state(n, groups, feeds)
            for mats in ops_:
                g.kronselect_dot(mats)
                c.kronselect_dot(mats)
            for idx in ([0], [n - 1, 0], [3, 1, 7], list(range(n))):
                check(name + " probs", g.measure_probabilities(np.array(idx, dtype=np.int32)), c.measure_probabilities(idx), 1e-13)
            ia, pa = g.measure_probabilities(np.array([0, 4, 9], dtype=np.int32), top_k=3)
            ib, pb = c.measure_probabilities([0, 4, 9], top_k=3)
            assert ia == ib, (ia, ib)
            assert abs(g.total_prob() - 1.0) < 1e-12
            err = check(name + " state", g.get_state(), c.get_state())
            random.seed(3)
            mg, pg = g.measure(np.array([0, 6], dtype=np.int32))
            random.seed(3)
            mc, pc = c.measure([0, 6])
            assert mg == mc and abs(pg - pc) < 1e-13, (mg, mc, pg, pc)
            check(name + " collapsed", g.get_state(), c.get_state())
            f = lambda x: (3 * x + 1) % 4
            g.func_apply([0, 5, 2], [1, 8], f)
            c.func_apply([0, 5, 2], [1, 8], f)
            check(name + " func", g.get_state(), c.get_state())
            check(name + " range", g.get_relative_range(100, 2300), c.get_relative_range(100, 2300))
            g.addto_relative_range(2040, 2056, np.arange(16) * (0.5 + 0.25j))
            c.addto_relative_range(2040, 2056, np.arange(16) * (0.5 + 0.25j))
            g.overwrite_relative_range(5, 9, np.array([1, 2, 3, 4], dtype=np.complex128))
            c.overwrite_relative_range(5, 9, np.array([1, 2, 3, 4], dtype=np.complex128))
            check(name + " range writes", g.get_state(), c.get_state())
            if quick:
                if rank == 0:
                    log("OK %-14s fuse=%-5s peer=%-5s err=%.1e exchanges=%d peer_gates=%d" % (name, fuse, peer, err, g.stats["exchanges"], g.stats["peer_gates"]))
                g.close()
                continue
            random.seed(4)
            mg, pg = g.reduce_measure(np.array([0, 7, 3], dtype=np.int32))
            random.seed(4)
            mc, pc = c.reduce_measure([0, 7, 3])
            assert mg == mc and abs(pg - pc) < 1e-12 * max(1.0, pc) and g.n == c.n == n - 3, (mg, mc, pg, pc)
            check(name + " reduced", g.get_state(), c.get_state())
            g.kronselect_dot({0: H2, (1, 8): CMat(X2)})
            c.kronselect_dot({0: H2, (1, 8): CMat(X2)})
            check(name + " after reduce", g.get_state(), c.get_state())
            if rank == 0:
                log("OK %-14s fuse=%-5s peer=%-5s err=%.1e exchanges=%d peer_gates=%d" % (name, fuse, peer, err, g.stats["exchanges"], g.stats["peer_gates"]))
            g.close()
    # kron-product init per shard (rank bits select the top sub-indices) and empty feed
    groups2 = [[7, 0, 5], [1], [11, 10]]
    feeds2 = [rng.normal(size=8) + 1j * rng.normal(size=8), [0.6, 0.8j], rng.normal(size=4)]
    g = ShardedB200Backend.make_state(n, groups2, feeds2)
    c = orc.OracleBackend.make_state(n, groups2, feeds2)
    check("kron init", g.get_state(), c.get_state(), 1e-15)
    g.close()
    g = ShardedB200Backend.make_state(n, [], [])
    want = np.zeros(2 ** n)
    want[0] = 1
    assert np.array_equal(g.get_state(), want)
    g.close()
    # one-hot int feeds (qip/distributed/backend.py:42-45) mixed with a vector group: only index bits are fixed
    groups3, hot = [[3, 0, 9], [1, 2], [11, 10, 4, 5]], np.zeros(8)
    hot[6] = 1.0
    v3 = rng.normal(size=4) + 1j * rng.normal(size=4)
    hot16 = np.zeros(16)
    hot16[9] = 1.0
    g = ShardedB200Backend.make_state(n, groups3, [6, v3, 9])
    c = orc.OracleBackend.make_state(n, groups3, [hot, v3, hot16])
    check("one-hot feeds", g.get_state(), c.get_state(), 1e-15)
    g.close()
    g = ShardedB200Backend.make_state(n, [list(range(n))], [2741])            # a basis state, no vector at all
    assert int(np.argmax(np.abs(g.get_state()))) == 2741 and abs(g.total_prob() - 1.0) < 1e-15
    g.close()
    if not quick:
        _compiled_and_lazy_cases(n, groups, feeds, rng, rank, world, log)
    # production tile counts against the ORACLE (n = 20): 2^(20-G) amplitudes per shard, i.e. many
    # tiles per CTA on the specialised fused kernels (TMA staging, the grid-stride loop, the mbarrier parity flip),
    # plus at least one multi-bit remap per layer
    n2 = 20
    g = ShardedB200Backend.make_state(n2, [], [])
    c = orc.RefBackend.make_state(n2, [], []) if _have_ref() else orc.OracleBackend.make_state(n2, [], [])
    for mats in layered_stream(n2, 3, 20):
        g.kronselect_dot(mats)
        c.kronselect_dot(mats)
    err = check("layered n=20 vs oracle", g.get_state(), c.get_state())
    if rank == 0:
        log("OK layered n=%d vs %s err=%.1e exchanges=%d" % (n2, type(c).__name__, err, g.stats["exchanges"]))
    g.close()
    # production-size shards (2^24 amplitudes each: the specialised fused kernels, the multi-bit remap): QFFT of
    # a basis state |j> against its closed form e^{+2 pi i j k / N} / sqrt(N)
    G = int(np.log2(world))
    nb = 24 + G
    j = 0x5A5A5A5 & ((1 << nb) - 1)
    g = ShardedB200Backend.make_state(nb, [list(range(nb))], [j])
    for mats in qfft_stream(nb):
        g.kronselect_dot(mats)
    assert abs(g.total_prob() - 1.0) < 1e-12
    err = qfft_closed_form_error(g, nb, j)
    assert err <= 1e-12, ("big qfft", err)
    REPORT["cases"] += 1
    REPORT["max_err"] = max(REPORT["max_err"], err)
    if rank == 0:
        log("OK big qfft n=%d err=%.1e exchanges=%d" % (nb, err, g.stats["exchanges"]))
    g.close()
    # the sharded engine against the SINGLE-GPU engine at a size no CPU oracle reaches ("1-GPU vs
    # sharded agree"): rank 0 also runs the circuit on its own GPU alone; sampled windows and a histogram must agree
    err = sharded_vs_single_gpu(26 if quick else 28, log)
    dist.barrier()
    return dict(REPORT)


def run_single_gpu_cases(log=print):
    """The single-GPU subset bench.py runs before its timed region at N = 1: the production fused kernels at many
    tiles per CTA against the oracle (layered, n = 20), dense 5-, 6- and 8-qubit gates (n = 16), the QFFT closed form at 26 qubits."""
    from qip_b200 import B200Backend
    REPORT["cases"], REPORT["max_err"] = 0, 0.0
    n2 = 20
    g = B200Backend.make_state(n2, [], [])
    c = orc.RefBackend.make_state(n2, [], []) if _have_ref() else orc.OracleBackend.make_state(n2, [], [])
    for mats in layered_stream(n2, 3, 20):
        g.kronselect_dot(mats)
        c.kronselect_dot(mats)
    idx = [0, 7, n2 - 1]
    check("layered n=20 probs", g.measure_probabilities(np.array(idx, dtype=np.int32)), c.measure_probabilities(idx), 1e-13)
    err = check("layered n=20 vs oracle", g.get_state(), c.get_state())
    random.seed(5)
    mg, pg = g.measure(np.array([3, 11], dtype=np.int32))
    random.seed(5)
    mc, pc = c.measure([3, 11])
    assert mg == mc and abs(pg - pc) < 1e-13, (mg, mc, pg, pc)
    check("layered n=20 collapsed", g.get_state(), c.get_state())
    log("OK layered n=%d vs %s err=%.1e" % (n2, type(c).__name__, err))
    g.close()
    # dense gates on 6 and 8 qubits (complex128: the FP64-tensor kernel) against the oracle, scattered targets, un-fused
    n3 = 16
    rng = np.random.default_rng(11)
    from qip_b200.circuits import haar_unitary
    g = B200Backend.make_state(n3, [], [])
    c = orc.RefBackend.make_state(n3, [], []) if _have_ref() else orc.OracleBackend.make_state(n3, [], [])
    for k in (1, 6, 8, 5):
        qs = tuple(int(q) for q in rng.permutation(n3)[:k])
        u = haar_unitary(rng, 2 ** k)
        g.kronselect_dot({qs if k > 1 else qs[0]: u})
        c.kronselect_dot({qs if k > 1 else qs[0]: u})
    err = check("dense 5/6/8-qubit gates n=16 vs oracle", g.get_state(), c.get_state())
    log("OK dense K=5,6,8 gates n=%d vs %s err=%.1e" % (n3, type(c).__name__, err))
    g.close()
    nb = 26
    j = 0x5A5A5A5 & ((1 << nb) - 1)
    g = B200Backend.make_state(nb, [list(range(nb))], [j])
    for mats in qfft_stream(nb):
        g.kronselect_dot(mats)
    err = qfft_closed_form_error(g, nb, j)
    assert err <= 1e-12 and abs(g.total_prob() - 1.0) < 1e-12, ("qfft closed form", err)
    REPORT["cases"] += 1
    REPORT["max_err"] = max(REPORT["max_err"], err)
    log("OK qfft n=%d closed form err=%.1e" % (nb, err))
    g.close()
    return dict(REPORT)


def _have_ref():
    from oracle.ref_loader import have_ref_ext
    return have_ref_ext()


def qfft_closed_form_error(g, n, j, windows=None, width=2048):
    """max over sampled windows of |amp - e^{+2 pi i j k / N} / sqrt(N)| * sqrt(N) for a state that should be QFFT|j>
    (qip/qfft.py:8-43; + sign: SURVEY 8g-1).  j * k is reduced mod N with python integers (it overflows int64 at n > 31)."""
    N = 1 << n
    if windows is None:
        windows = (0, 12345, N - 2 * width, (N >> 1) - 7, (N >> 2) + 3 * width + 1)
    worst = 0.0
    for start in windows:
        frac = np.array([((j * k) % N) / float(N) for k in range(start, start + width)], dtype=np.float64)
        want = np.exp(2j * np.pi * frac)
        got = np.asarray(g.get_relative_range(start, start + width)) * np.sqrt(float(N))
        worst = max(worst, float(np.max(np.abs(got - want))))
    return worst


def sharded_vs_single_gpu(n, log=print):
    from qip_b200 import B200Backend
    rank = dist.get_rank()
    ops_ = list(layered_stream(n, 2, 77)) + list(qfft_stream(n))
    idx = np.array([0, n // 3, n - 1, 5, n // 2], dtype=np.int32)
    g = ShardedB200Backend.make_state(n, [], [])
    for mats in ops_:
        g.kronselect_dot(mats)
    pg = g.measure_probabilities(idx)
    N = 1 << n
    starts = [0, 4097, N // 2 - 1000, N - 4096, (N // 8) * 5 + 17]
    wins = [g.get_relative_range(s, s + 4096) for s in starts]
    exchanges = g.stats["exchanges"]
    g.close()
    err = 0.0
    if rank == 0:
        b = B200Backend.make_state(n, [], [])
        for mats in ops_:
            b.kronselect_dot(mats)
        pb = b.measure_probabilities(idx)
        err = float(np.max(np.abs(pg - pb)))
        scale = float(np.sqrt(N))
        for s, w in zip(starts, wins):
            err = max(err, float(np.max(np.abs(w - b.get_relative_range(s, s + 4096)))) * scale)
        b.close()
    t = torch.tensor([err], dtype=torch.float64, device="cuda")
    dist.broadcast(t, 0)
    err = float(t.item())
    assert err <= 1e-12, ("sharded vs single GPU", n, err)
    REPORT["cases"] += 1
    REPORT["max_err"] = max(REPORT["max_err"], err)
    if rank == 0:
        log("OK sharded == single GPU at n=%d (err %.1e x sqrt(N), %d exchanges)" % (n, err, exchanges))
    return err


def _compiled_and_lazy_cases(n, groups, feeds, rng, rank, world, log):
    # compiled circuit replayed on the sharded engine: the flush of every gate segment caches its rank-local
    # program (schedule + planned passes) and the replays must reproduce the first run
    from qip_b200.graph import CompiledCircuit
    seg = list(layered_stream(n, 2, 4)) + list(qfft_stream(n))
    ops_c = [("k", m) for m in seg] + [("p", [0, n - 1, 5])] + [("k", m) for m in layered_stream(n, 1, 8)]
    circ = CompiledCircuit.from_ops(n, groups, feeds, ops_c)
    c = orc.OracleBackend.make_state(n, groups, feeds)
    for m in seg:
        c.kronselect_dot(m)
    want_p = c.measure_probabilities([0, n - 1, 5])
    for m in layered_stream(n, 1, 8):
        c.kronselect_dot(m)
    for replay in range(3):
        state, classic = circ.run(backend_constructor=ShardedB200Backend.make_state)
        check("compiled state %d" % replay, state, c.get_state())
        check("compiled probs %d" % replay, classic[len(seg)], want_p, 1e-13)
        assert circ.last_stats.get("cached_flushes", 0) == (0 if replay == 0 else 2), circ.last_stats
    if rank == 0:
        log("OK compiled circuit replays (cached sharded programs)")
    # lazy product-state init on every rank (QIPB_LAZY_INIT): the first rank-local fused pass writes its tiles
    nb = 16 + int(np.log2(world))
    pgroups = [[q] for q in range(nb)]
    pfeeds = []
    for _ in range(nb):
        v = rng.normal(size=2) + 1j * rng.normal(size=2)
        pfeeds.append(v / np.linalg.norm(v))
    g = ShardedB200Backend.make_state(nb, pgroups, pfeeds, lazy_init=True)
    c = orc.OracleBackend.make_state(nb, pgroups, pfeeds)
    for mats in layered_stream(nb, 2, 6):
        g.kronselect_dot(mats)
        c.kronselect_dot(mats)
    check("lazy init", g.get_state(), c.get_state())
    assert g.stats.get("fill_passes") == 1, g.stats
    g.close()
    if rank == 0:
        log("OK lazy product-state init (fill pass)")


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rep = run_cases(quick=False)
    if dist.get_rank() == 0:
        print("SHARDED PARITY OK world=%d cases=%d max_err=%.2e" % (dist.get_world_size(), rep["cases"], rep["max_err"]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
