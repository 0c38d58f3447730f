"""CPU tier: the real qip_b200.sharded.ShardedB200Backend on P = 2, 4, 8 VIRTUAL ranks (threads of one process).

tests/hostlib.py supplies the doubles: host buffers as shards, the other thread's buffer as peer memory (the
qipb_peer_* kernels restated in numpy), thread rendezvous for torch.distributed, the fused-pass emulator for the
rank-local passes.  What runs is the product's own multi-GPU host code -- lazy layout choice, kron init per shard,
the scheduler (defer_global, exchanges, multi-bit remaps, peer gates), rank-local merging / planning / launching,
measurement across shards, func_apply, range access, canonicalisation, shard pooling, the compiled-circuit program
cache -- against the CPU oracle.  The multi-GPU box runs the same scenario over NVLink (tests/dist_gpu_worker.py)."""
import random

import numpy as np
import pytest

import hostlib
from oracle import oracle as orc
from qip_b200.circuits import H2, X2, haar_unitary, layered_stream, qfft_stream, rm_mat
from qip_b200.mats import CMat, SwapMat


def _check(name, got, want, tol=1e-12):
    err = float(np.max(np.abs(np.asarray(got) - np.asarray(want)))) / max(1e-300, float(np.max(np.abs(want))))
    assert err <= tol, (name, err)


@pytest.mark.parametrize("P", [2, 4, 8])
def test_sharded_backend_matches_oracle_on_virtual_ranks(monkeypatch, P):
    from qip_b200.sharded import ShardedB200Backend
    n = 11
    rng = np.random.default_rng(5)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    groups, feeds = [list(range(n))], [psi]
    rngu = np.random.default_rng(6)
    extra = [{(0, 5): CMat(X2)}, {(1, 0, 6): CMat(CMat(haar_unitary(rngu, 2)))}, {0: rm_mat(3)},
             {(1, 0): CMat(rm_mat(2))}, {(0, 7): SwapMat(1)}, {(0, 1): haar_unitary(rngu, 4)},
             {(10, 0, 3): CMat(SwapMat(1))}, {0: H2}, {(2, 9, 0): haar_unitary(rngu, 8)},
             {(2, 9, 0, 4, 7, 1): haar_unitary(rngu, 64)}]   # six qubits, three of them global at P = 8: the batched dense kernel on a shard
    cases = {"layered": list(layered_stream(n, 3, 2)), "qfft": list(qfft_stream(n)), "mixed": extra}

    def body(rank):
        for name, ops_ in cases.items():
            for fuse, peer in ((True, False), (False, False), (True, True)):
                g = ShardedB200Backend.make_state(n, groups, feeds, statetype=np.complex128, fuse=fuse, peer_gates=peer,
                                                  tile_bits=5, min_low_bits=2)
                c = orc.OracleBackend.make_state(n, groups, feeds)
                for mats in ops_:
                    g.kronselect_dot(mats)
                    c.kronselect_dot(mats)
                for idx in ([0], [n - 1, 0], [3, 1, 7], list(range(n))):
                    _check(name + " probs", g.measure_probabilities(np.array(idx, dtype=np.int32)), c.measure_probabilities(idx), 1e-13)
                ia, pa = g.measure_probabilities(np.array([0, 4, 9], dtype=np.int32), top_k=3)
                ib, pb = c.measure_probabilities([0, 4, 9], top_k=3)
                assert ia == ib
                assert abs(g.total_prob() - 1.0) < 1e-12
                _check(name + " state", g.get_state(), c.get_state())
                random.seed(3)
                mg, pg = g.measure(np.array([0, 6], dtype=np.int32))
                random.seed(3)
                mc, pc = c.measure([0, 6])
                assert mg == mc and abs(pg - pc) < 1e-13
                _check(name + " collapsed", g.get_state(), c.get_state())
                f = lambda x: (3 * x + 1) % 4
                g.func_apply([0, 5, 2], [1, 8], f)
                c.func_apply([0, 5, 2], [1, 8], f)
                _check(name + " func", g.get_state(), c.get_state())
                _check(name + " range", g.get_relative_range(100, 1300), c.get_relative_range(100, 1300))
                g.addto_relative_range(1016, 1032, np.arange(16) * (0.5 + 0.25j))
                c.addto_relative_range(1016, 1032, np.arange(16) * (0.5 + 0.25j))
                g.overwrite_relative_range(5, 9, np.array([1, 2, 3, 4], dtype=np.complex128))
                c.overwrite_relative_range(5, 9, np.array([1, 2, 3, 4], dtype=np.complex128))
                _check(name + " range writes", g.get_state(), c.get_state())
                random.seed(4)
                mg, pg = g.reduce_measure(np.array([0, 7, 3], dtype=np.int32))
                random.seed(4)
                mc, pc = c.reduce_measure([0, 7, 3])
                assert mg == mc and abs(pg - pc) < 1e-12 * max(1.0, pc) and g.n == c.n == n - 3
                _check(name + " reduced", g.get_state(), c.get_state())
                g.close()
        # kron-product init per shard, one-hot feeds, empty feed
        g = ShardedB200Backend.make_state(n, [[3, 0, 9], [1, 2], [10, 8, 4, 5]], [6, np.array([0.5, 0.5j, -0.5, 0.5]), 9])
        c = orc.OracleBackend.make_state(n, [[3, 0, 9], [1, 2], [10, 8, 4, 5]], [np.eye(8)[6], np.array([0.5, 0.5j, -0.5, 0.5]), np.eye(16)[9]])
        _check("one-hot feeds", g.get_state(), c.get_state(), 1e-15)
        g.close()
        g = ShardedB200Backend.make_state(n, [], [])
        want = np.zeros(2 ** n)
        want[0] = 1
        assert np.array_equal(g.get_state(), want)
        g.close()

    hostlib.run_virtual_ranks(monkeypatch, P, body)


@pytest.mark.parametrize("P", [2, 8])
def test_compiled_circuit_replays_on_virtual_ranks_with_cached_programs(monkeypatch, P):
    from qip_b200.graph import CompiledCircuit
    from qip_b200.sharded import ShardedB200Backend
    n = 11
    rng = np.random.default_rng(1)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    groups, feeds = [list(range(n))], [psi]
    seg = list(layered_stream(n, 2, 4)) + list(qfft_stream(n))
    tail = list(layered_stream(n, 1, 8))
    ops_c = [("k", m) for m in seg] + [("p", [0, n - 1, 5])] + [("k", m) for m in tail]
    c = orc.OracleBackend.make_state(n, groups, feeds)
    for m in seg:
        c.kronselect_dot(m)
    want_p = c.measure_probabilities([0, n - 1, 5])
    for m in tail:
        c.kronselect_dot(m)
    want_state = c.get_state()

    def body(rank):
        circ = CompiledCircuit.from_ops(n, groups, feeds, ops_c)          # one compiled circuit (and cache) per rank/process
        for replay_no in range(3):
            state, classic = circ.run(backend_constructor=ShardedB200Backend.make_state, tile_bits=5, min_low_bits=2)
            _check("compiled state", state, want_state)
            _check("compiled probs", classic[len(seg)], want_p, 1e-13)
            assert circ.last_stats.get("cached_flushes", 0) == (0 if replay_no == 0 else 2), circ.last_stats

    hostlib.run_virtual_ranks(monkeypatch, P, body)


@pytest.mark.parametrize("P", [2, 4])
def test_sharded_backend_at_production_tile_size_on_virtual_ranks(monkeypatch, P):
    # shards of 2^13 amplitudes, default tile shape (2^12-amplitude tiles, 2 KiB runs): the rank-local passes take the
    # specialised sweeps of the fused kernel (through the emulator); QFFT against its closed form sqrt(N) * ifft
    from qip_b200.sharded import ShardedB200Backend
    n = 13 + int(np.log2(P))
    rng = np.random.default_rng(n)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    want_qft = np.fft.ifft(psi) * np.sqrt(2 ** n)
    c = orc.OracleBackend.make_state(n, [list(range(n))], [psi])
    layer = list(layered_stream(n, 2, 3))
    for m in layer:
        c.kronselect_dot(m)
    want_layered = c.get_state()

    def body(rank):
        g = ShardedB200Backend.make_state(n, [list(range(n))], [psi])
        for m in qfft_stream(n):
            g.kronselect_dot(m)
        _check("qfft", g.get_state(), want_qft)
        assert g.stats["exchanges"] >= 1
        g.close()
        g = ShardedB200Backend.make_state(n, [list(range(n))], [psi])
        for m in layer:
            g.kronselect_dot(m)
        g.flush()
        _check("layered", g.get_state(), want_layered)
        g.close()

    L = hostlib.run_virtual_ranks(monkeypatch, P, body)
    assert ("apply_fused" in L.log or "apply_fused_chunk" in L.log) and \
        ("peer_remap" in L.log or "peer_swap_bit" in L.log or "peer_remap_chunk" in L.log)


@pytest.mark.parametrize("P", [2, 4])
@pytest.mark.parametrize("workload", ["layered", "qft"])
def test_sharded_lazy_product_state_init_on_virtual_ranks(monkeypatch, P, workload):
    # QIPB_LAZY_INIT: every rank keeps its slice of a product state virtual (the rank bits' factors become a scalar of
    # the shard) and the first rank-local fused pass writes its tiles instead of loading them
    from qip_b200.sharded import ShardedB200Backend
    n = 13 + int(np.log2(P))
    rng = np.random.default_rng(3)
    groups = [[q] for q in range(n - 3)] + [[n - 1, n - 3]]              # qubit n-2 stays un-fed
    feeds = []
    for _ in range(n - 3):
        v = rng.normal(size=2) + 1j * rng.normal(size=2)
        feeds.append(v / np.linalg.norm(v))
    stream = list(layered_stream(n, 2, 1)) if workload == "layered" else list(qfft_stream(n))
    c = orc.OracleBackend.make_state(n, groups, feeds + [np.eye(4)[1]])
    want0 = c.get_state().copy()
    for m in stream:
        c.kronselect_dot(m)
    want = c.get_state()

    def body(rank):
        g = ShardedB200Backend.make_state(n, groups, feeds + [1], lazy_init=True)
        for m in stream:
            g.kronselect_dot(m)
        _check("lazy " + workload, g.get_state(), want)
        assert g.stats.get("fill_passes") == 1, g.stats
        g.close()
        g = ShardedB200Backend.make_state(n, groups, feeds + [1], lazy_init=True)     # observed before any gate
        _check("lazy, no gates", g.get_state(), want0, 1e-15)
        assert "fill_passes" not in g.stats
        g.close()

    L = hostlib.run_virtual_ranks(monkeypatch, P, body)
    assert L.log.count("apply_fused_fill") == P and L.log.count("init_kron") == P


@pytest.mark.parametrize("seed", range(9))
def test_random_sessions_through_the_real_sharded_backend_on_virtual_ranks(monkeypatch, seed):
    import fuzzlib
    from qip_b200.sharded import ShardedB200Backend
    P = [2, 4, 8][seed % 3]
    n = int(np.log2(P)) + 5 + seed % 4

    def body(rank):
        fuzzlib.session(ShardedB200Backend.make_state, 100 + seed, n, lazy_init=bool(seed % 2), fuse=bool(seed % 5),
                        peer_gates=(seed % 4 == 0), tile_bits=5, min_low_bits=2)

    hostlib.run_virtual_ranks(monkeypatch, P, body)


@pytest.mark.parametrize("P", [2, 4])
def test_unclosed_sharded_states_do_not_leak_their_shards(monkeypatch, P):
    # the reference's run() returns get_state() and forgets the backend (qip/pipeline.py:248): a dropped state hands its
    # shard back (pool, or the orphan list that the next construction frees collectively) -- ADVICE r01
    import gc
    from qip_b200.sharded import ShardedB200Backend
    n = 9

    def body(rank):
        for rep in range(6):
            g = ShardedB200Backend.make_state(n, [], [], tile_bits=5, min_low_bits=2)
            h = ShardedB200Backend.make_state(n, [], [], tile_bits=5, min_low_bits=2)      # two live states of one size
            g.kronselect_dot({0: H2})
            h.kronselect_dot({1: H2})
            assert abs(g.total_prob() - 1.0) < 1e-12 and abs(h.total_prob() - 1.0) < 1e-12
            del g, h                                                                       # never closed
            gc.collect()
        g = ShardedB200Backend.make_state(n, [], [], tile_bits=5, min_low_bits=2)
        g.close()

    before = set(hostlib._PeerMixin._bufs)
    hostlib.run_virtual_ranks(monkeypatch, P, body)
    # P ranks x (one pooled shard each) stay allocated; everything else was returned
    left = set(hostlib._PeerMixin._bufs) - before
    assert len(left) <= P, len(left)


def test_get_state_of_a_large_sharded_state_is_a_lazy_handle(monkeypatch):
    # run() calls get_state() unconditionally; beyond the host limit it must return a handle, not raise
    from qip_b200 import sharded as sh
    n, P = 10, 4
    rng = np.random.default_rng(2)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    c = orc.OracleBackend.make_state(n, [list(range(n))], [psi])
    ops_ = list(layered_stream(n, 2, 5))
    for m in ops_:
        c.kronselect_dot(m)
    want = c.get_state()
    monkeypatch.setattr(sh, "_HOST_STATE_MAX_QUBITS", 8)

    def body(rank):
        g = sh.ShardedB200Backend.make_state(n, [list(range(n))], [psi], tile_bits=5, min_low_bits=2)
        for m in ops_:
            g.kronselect_dot(m)
        st = g.get_state()
        assert isinstance(st, sh.ShardedState) and len(st) == 2 ** n and st.shape == (2 ** n,)
        _check("slice", st[100:900], want[100:900])
        _check("strided", st[3:700:7], want[3:700:7])
        assert abs(st[517] - want[517]) < 1e-12 and abs(st[-1] - want[-1]) < 1e-12
        _check("array", np.asarray(st), want)
        lo, hi = st.local_range
        _check("local shard", st.local_shard.numpy(), want[lo:hi])
        g.close()

    hostlib.run_virtual_ranks(monkeypatch, P, body)


@pytest.mark.parametrize("P,nl", [(4, 2), (4, 3), (8, 2), (2, 1)])
def test_tiny_shards_do_not_coalesce_more_exchanges_than_local_bits(monkeypatch, P, nl):
    # qipb_peer_remap needs nbits > g; with nl <= 3 a 3-pair MultiExchange was built and failed after the
    # layout had been mutated.  One local qubit cannot host a 2-qubit gate: a clear error, raised by the scheduler.
    from qip_b200.sharded import ShardedB200Backend
    G = int(np.log2(P))
    n = G + nl
    ops_ = [{q: H2} for q in range(n)] + [{(q, (q + 1) % n): CMat(X2)} for q in range(n)] + [{q: H2} for q in range(n)]
    c = orc.OracleBackend.make_state(n, [], [])
    for m in ops_:
        c.kronselect_dot(m)
    want = c.get_state()

    def body(rank):
        g = ShardedB200Backend.make_state(n, [], [], tile_bits=min(5, nl), min_low_bits=min(2, nl))
        for m in ops_:
            g.kronselect_dot(m)
        _check("tiny shards", g.get_state(), want)
        g.close()

    hostlib.run_virtual_ranks(monkeypatch, P, body)


@pytest.mark.parametrize("P", [2, 4, 8])
@pytest.mark.parametrize("chunk_bits", [1, 2, 3])
def test_exchange_compute_overlap_pipeline_on_virtual_ranks(monkeypatch, P, chunk_bits):
    # the chunk pipeline (passes A -> exchange -> passes B per chunk: qipb_apply_fused_chunk / qipb_peer_remap_chunk)
    # must give what the un-overlapped program gives; the doubles also assert that a chunked pass leaves every other
    # chunk untouched
    from qip_b200.sharded import ShardedB200Backend
    monkeypatch.setenv("QIPB_OVERLAP_CHUNK_BITS", str(chunk_bits))
    monkeypatch.setenv("QIPB_OVERLAP_MIN_BYTES", "0")
    n = 9 + int(np.log2(P))
    rng = np.random.default_rng(P + chunk_bits)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    streams = {"layered": list(layered_stream(n, 4, 11)), "qfft": list(qfft_stream(n)),
               "controls": [{(0, n - 1): CMat(X2)}, {(n - 1, 0, 4): CMat(CMat(H2))}, {0: H2}, {(1, 0): CMat(rm_mat(2))}, {1: H2},
                            {(2, n - 2): haar_unitary(rng, 4)}, {(n - 1, 1): CMat(rm_mat(3))}, {n - 1: H2}, {(3, 0): haar_unitary(rng, 4)}],
               # dense gates on 6 and 5 qubits (never fused: the batched kernel on a shard) between fusable layers, global targets
               "dense6": list(layered_stream(n, 1, 5)) + [{(2, n - 1, 0, 4, n - 2, 1): haar_unitary(rng, 64)}] + list(layered_stream(n, 1, 6)) +
                         [{(0, 3, 5, 6, n - 3): haar_unitary(rng, 32)}, {(1, 0): CMat(rm_mat(2))}] + list(layered_stream(n, 1, 7))}
    wants = {}
    for name, ops_ in streams.items():
        c = orc.OracleBackend.make_state(n, [list(range(n))], [psi])
        for m in ops_:
            c.kronselect_dot(m)
        wants[name] = c.get_state().copy()
    seen = []

    def body(rank):
        for name, ops_ in streams.items():
            for overlap in (True, False):
                g = ShardedB200Backend.make_state(n, [list(range(n))], [psi], tile_bits=5, min_low_bits=2, overlap=overlap)
                for m in ops_:
                    g.kronselect_dot(m)
                    if name == "layered" and len(g.queue) > 3 * n:      # several flushes: layouts evolve between them
                        g.flush()
                _check(name, g.get_state(), wants[name])
                if rank == 0:
                    seen.append((name, overlap, g.stats.get("overlapped_exchanges", 0), g.stats["exchanges"]))
                g.close()

    L = hostlib.run_virtual_ranks(monkeypatch, P, body)
    assert any(ov and k > 0 for _, ov, k, _ in seen), seen
    assert all(k == 0 for _, ov, k, _ in seen if not ov), seen
    assert "apply_fused_chunk" in L.log and "peer_remap_chunk" in L.log
