"""CPU tier: host-side logic of the product (op decoding, exact simplification, lowering, fusion
planner, sampling scan, func tabulation, circuit generators) -- no device code involved.
The planner is executed by a tests-only numpy simulator (tests/bitsim.py) and compared with the
golden streams recorded from the reference."""
import numpy as np
import pytest

import bitsim
import replay
from oracle import oracle as orc
from qip_b200 import ops
from qip_b200.backend import scan_outcome, tabulate, top_probabilities
from qip_b200.circuits import H2, X2, layered_stream, qfft_stream, rm_mat
from qip_b200.mats import CMat, SwapMat

META, STREAMS, ARRAYS = replay.load_streams()


class PlannedHostBackend(orc.OracleBackend):
    """Oracle state container whose kronselect_dot goes through the PRODUCT's host pipeline
    (decode -> simplify -> lower -> plan_passes) and a numpy executor of the resulting plan."""
    fuse = True
    strategy = "auto"

    @classmethod
    def make_state(cls, n, index_groups, feed_list, statetype=np.complex128, **_):
        ob = orc.OracleBackend.make_state(n, index_groups, feed_list)
        b = cls(n, ob.state)
        b.queue = []
        return b


    def kronselect_dot(self, mats, input_offset=0, output_offset=0):
        for g in ops.decode_mats(mats, self.n):
            s = ops.simplify(g)
            if s is not None:
                self.queue.append(s)

    def _flush(self):
        if self.queue:
            passes, _ = ops.plan(self.queue, self.n, 16, fuse=self.fuse, tile_bits=min(12, max(2, self.n - 1)),
                                 min_low_bits=min(6, max(0, self.n - 3)), strategy=self.strategy)
            self.state = bitsim.run_passes(self.state, passes, self.n)
            self.arena = np.empty_like(self.state)
            self.queue = []

    def get_state(self):
        self._flush()
        return self.state

    def func_apply(self, *a, **k):
        self._flush()
        return super().func_apply(*a, **k)

    def measure(self, *a, **k):
        self._flush()
        return super().measure(*a, **k)

    def soft_measure(self, *a, **k):
        self._flush()
        return super().soft_measure(*a, **k)

    def reduce_measure(self, *a, **k):
        self._flush()
        return super().reduce_measure(*a, **k)

    def measure_probabilities(self, *a, **k):
        self._flush()
        return super().measure_probabilities(*a, **k)

    def total_prob(self):
        self._flush()
        return super().total_prob()

    def apply_gates(self, gates, cache=None, key=None):       # compiled circuits (qip_b200.graph)
        self.queue.extend(gates)

    def close(self):
        pass


class TileHostBackend(PlannedHostBackend):
    strategy = "tile"


class Dense4HostBackend(PlannedHostBackend):
    strategy = "dense4"


SMALL = [i for i, s in enumerate(STREAMS) if s["n"] <= 12]


@pytest.mark.parametrize("i", SMALL, ids=[STREAMS[i]["label"] for i in SMALL])
def test_host_pipeline_replays_golden_stream(i):
    assert replay.replay(STREAMS[i], ARRAYS, TileHostBackend.make_state, tol=1e-12)
    assert replay.replay(STREAMS[i], ARRAYS, Dense4HostBackend.make_state, tol=1e-12)


def test_decode_validation_matches_reference_errors():       # qip/util.py:35-58, kronprod.pyx:114-116
    with pytest.raises(Exception, match="Type of indices must be tuple"):
        ops.decode_mats({"a": np.eye(2)}, 3)
    with pytest.raises(Exception, match="Shape of square submatrix"):
        ops.decode_mats({(0, 1): np.eye(2)}, 3)
    with pytest.raises(ValueError, match="not numpy, SwapMat, or CMat"):
        class Odd:
            shape = (2, 2)
        ops.decode_mats({0: Odd()}, 3)
    with pytest.raises(ValueError):
        ops.decode_mats({5: np.eye(2)}, 3)
    with pytest.raises(ValueError, match="more than one entry"):
        ops.decode_mats({0: H2, (0, 1): np.eye(4)}, 3)
    # entries may share CONTROL qubits: C(C(Not)) onto a 2-qubit register
    g = ops.decode_mats({(0, 1, 2): CMat(CMat(X2)), (0, 1, 3): CMat(CMat(X2))}, 4)
    assert [x.controls for x in g] == [(0, 1), (0, 1)] and [x.targets for x in g] == [(2,), (3,)]
    # list values and int keys are normalised
    g = ops.decode_mats({1: [[0, 1], [1, 0]]}, 2)
    assert g[0].targets == (1,) and g[0].mat.dtype == np.complex128


def test_swapmat_decodes_to_bit_swaps():
    g = ops.decode_mats({(0, 1, 2, 3, 4): CMat(SwapMat(2))}, 5)
    assert [(x.kind, x.targets, x.controls) for x in g] == [("swap", (1, 3), (0,)), ("swap", (2, 4), (0,))]


def test_simplify_promotes_controls_and_detects_diagonals():
    rm = ops.simplify(ops.decode_mats({2: rm_mat(3)}, 4)[0])
    assert rm.targets == () and rm.controls == (2,) and rm.mat.shape == (1, 1) and rm.diagonal
    cp = ops.simplify(ops.decode_mats({(3, 1): CMat(rm_mat(2))}, 4)[0])
    assert cp.targets == () and set(cp.controls) == {3, 1}
    cx = ops.simplify(ops.decode_mats({(0, 1): np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])}, 2)[0])
    assert cx.targets == (1,) and cx.controls == (0,) and np.array_equal(cx.mat, X2)
    assert ops.simplify(ops.decode_mats({0: np.eye(2)}, 1)[0]) is None
    z = ops.simplify(ops.decode_mats({0: np.array([[1, 0], [0j, -1]])}, 1)[0])
    assert z.targets == () and z.mat[0, 0] == -1
    d = ops.simplify(ops.decode_mats({0: np.diag([1j, -1j])}, 1)[0])
    assert d.targets == (0,) and d.diagonal
    h = ops.simplify(ops.decode_mats({0: H2}, 1)[0])
    assert h.targets == (0,) and not h.diagonal


def test_planner_fuses_qft_into_few_passes():
    n = 24
    gates = []
    for mats in qfft_stream(n):
        for g in ops.decode_mats(mats, n):
            gates.append(ops.simplify(g))
    assert len(gates) == n + n * (n - 1) // 2 + n // 2
    passes, name = ops.plan(gates, n, 16, strategy="tile")
    assert len(passes) <= 12, len(passes)
    for p in passes:
        if p.fused:
            assert len(p.tile_bits) == 12 and p.tile_bits[:6] == (0, 1, 2, 3, 4, 5)
    unfused, name = ops.plan(gates, n, 16, fuse=False)
    assert name == "unfused" and len(unfused) == len(gates) and not any(p.fused for p in unfused)
    _, auto = ops.plan(gates, n, 16)
    assert auto == "tile"                                   # diagonal-heavy: tile passes beat dense blocks


def test_merge_blocks_is_exact_and_shrinks_layered_circuits():
    n = 10
    gates = []
    for mats in layered_stream(n, 2, 5):
        for g in ops.decode_mats(mats, n):
            s = ops.simplify(g)
            if s is not None:
                gates.append(s)
    rng = np.random.default_rng(0)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    ref = psi.copy()
    for g in gates:
        ref = bitsim.apply_bitgate(ref, ops.lower(g, n), n)
    for mk in (2, 3, 4):
        merged = ops.merge_blocks(gates, mk)
        assert len(merged) < len(gates) and all(len(g.qubits()) <= mk for g in merged)
        out = psi.copy()
        for g in merged:
            out = bitsim.apply_bitgate(out, ops.lower(g, n), n)
        assert float(np.max(np.abs(out - ref))) <= 1e-13
    # a lone swap stays a (half-traffic) swap, CX stays controlled after merging
    keep = ops.merge_blocks([ops.Gate("swap", (0, 1)), ops.simplify(ops.decode_mats({(2, 3): CMat(X2)}, 4)[0])], 2)
    assert keep[0].kind == "swap" and keep[1].controls == (2,)


def test_planner_keeps_cheap_controlled_gates_unfused():
    n = 20
    g = ops.lower(ops.simplify(ops.decode_mats({(3, 1): CMat(rm_mat(2))}, n)[0]), n)
    g2 = ops.lower(ops.simplify(ops.decode_mats({(5, 7): CMat(rm_mat(2))}, n)[0]), n)
    passes = ops.plan_passes([g, g2], n, 16)
    assert [p.fused for p in passes] == [False, False]      # 2 x 1/4 of the state < one full sweep


@pytest.mark.parametrize("seed", range(8))
def test_planned_execution_equals_oracle_on_random_circuits(seed):
    rng = np.random.default_rng(seed)
    n = 9
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    a = orc.OracleBackend.make_state(n, [list(range(n))], [psi])
    b = PlannedHostBackend.make_state(n, [list(range(n))], [psi])
    for mats in layered_stream(n, 3, seed):
        a.kronselect_dot(mats)
        b.kronselect_dot(mats)
    assert replay.close(b.get_state(), a.get_state(), 1e-12)


def test_scan_outcome_matches_reference_semantics():         # qip/ext/kronprod.pyx:364-381
    probs = np.array([0.0, 0.64, 0.0, 0.36])
    assert scan_outcome(probs, 0.0) == (0, 0.0)              # r == 0 stops at the first outcome
    assert scan_outcome(probs, 0.5) == (1, 0.64)
    assert scan_outcome(probs, 0.64 + 1e-9)[0] == 3
    assert scan_outcome(probs, 1.0)[0] == 3
    assert scan_outcome(np.array([0.5, 0.5]), 0.999999)[0] == 1
    # a draw that rounding leaves positive after the last outcome returns the last outcome
    assert scan_outcome(np.array([0.25, 0.25]), 1.0 + 1e-12)[0] == 1


def test_top_probabilities():
    assert top_probabilities([0.0, 0.0, 0.64, 0.36], 4) == ([2, 3, 0, 1], [0.64, 0.36, 0.0, 0.0])
    assert top_probabilities([0.1, 0.2, 0.3, 0.4], 3)[0] == [3, 2, 1]
    assert top_probabilities([0.5, 0.5], 9)[0] == [0, 1]


def test_tabulate_vectorised_and_fallback():
    assert np.array_equal(tabulate(lambda x: (x + 1) % 4, 2), [1, 2, 3, 0])
    assert np.array_equal(tabulate(lambda x: 1, 3), np.ones(8))
    assert np.array_equal(tabulate(lambda x: int(x == 5), 3), [0, 0, 0, 0, 0, 1, 0, 0])      # int() of an array raises
    assert np.array_equal(tabulate(lambda x: pow(3, int(x), 8), 3), [pow(3, x, 8) for x in range(8)])
    assert np.array_equal(tabulate(lambda x: (x == 5) * 1, 12)[:8], [0, 0, 0, 0, 0, 1, 0, 0])


def _stream_of(label):
    return [s for s in STREAMS if s["label"] == label][0]


def _same_mats(a, b):
    if list(a.keys()) != list(b.keys()):
        return False
    for k in a:
        x, y = a[k], b[k]
        while getattr(x, "_kron_struct", None) == 2:
            if getattr(y, "_kron_struct", None) != 2:
                return False
            x, y = x.m, y.m
        if getattr(x, "_kron_struct", None) == 3:
            if getattr(y, "_kron_struct", None) != 3 or x.n != y.n:
                return False
        elif not np.array_equal(np.asarray(x, dtype=np.complex128), np.asarray(y, dtype=np.complex128)):
            return False
    return True


def test_qfft_generator_equals_reference_front_end_stream():
    s = _stream_of("config/qfft8")
    ref_ops = [replay.dec_mats(ARRAYS, op["mats"]) for op in s["ops"]]
    mine = list(qfft_stream(8))
    assert len(ref_ops) == len(mine) == 8 + 28 + 4
    assert all(_same_mats(a, b) for a, b in zip(mine, ref_ops))


def test_layered_generator_equals_reference_front_end_stream():
    s = _stream_of("config/layered_n8_d4_s33")
    ref_ops = [replay.dec_mats(ARRAYS, op["mats"]) for op in s["ops"]]
    mine = list(layered_stream(8, 4, 33))
    assert len(ref_ops) == len(mine)
    assert all(_same_mats(a, b) for a, b in zip(mine, ref_ops))


def test_sink_lone_diagonals_only_reorders_commuting_gates():
    # isolated diagonal gates move behind the gates they commute with and meet in one run (one table sweep in
    # the fused kernel); runs that already fold (QFT) stay put; the circuit is unchanged
    n = 7
    rng = np.random.default_rng(12)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    from qip_b200.circuits import haar_unitary
    mats_list = [{0: rm_mat(3)}, {(1, 2): haar_unitary(rng, 4)}, {3: rm_mat(2)}, {(0, 4): CMat(X2)}, {5: rm_mat(4)},
                 {(3, 6): haar_unitary(rng, 4)}, {(2, 5): CMat(rm_mat(3))}, {5: H2}, {1: rm_mat(6)}, {(6, 1): CMat(X2)}]
    gates = [s for mats in mats_list for s in (ops.simplify(g) for g in ops.decode_mats(mats, n)) if s is not None]
    moved = ops.sink_lone_diagonals(gates)
    assert sorted(map(id, moved)) == sorted(map(id, gates)) and [id(g) for g in moved] != [id(g) for g in gates]
    a, b = psi.copy(), psi.copy()
    for g in gates:
        a = bitsim.apply_bitgate(a, ops.lower(g, n), n)
    for g in moved:
        b = bitsim.apply_bitgate(b, ops.lower(g, n), n)
    assert float(np.max(np.abs(a - b))) <= 1e-13
    # the phases on qubits 0 and 3 cannot pass the gates that target those qubits; the other lone phases (qubits 5: no,
    # H(5) targets it; qubit 1: CX targets it) are emitted right in front of their blockers -- at least two end up adjacent
    diag = [g.kind == "matrix" and (g.diagonal or g.k == 0) for g in moved]
    assert any(diag[i] and diag[i + 1] for i in range(len(diag) - 1))
    # a QFT keeps its structure: every H whose controlled phases fold into a stage (runs of >= 3) is still
    # directly followed by them (H0 + 5, H1 + 4, H2 + 3 phases = the first 15 gates of a 6-qubit QFFT)
    q = [s for mats in qfft_stream(6, rev=False) for s in (ops.simplify(g) for g in ops.decode_mats(mats, 6)) if s is not None]
    assert [id(g) for g in ops.sink_lone_diagonals(q)[:15]] == [id(g) for g in q[:15]]


def test_pack_lone_1q_tensors_commuting_gates_only():
    from qip_b200.circuits import haar_unitary
    # inside one pass: pairs of un-controlled dense 1-qubit gates become one 2-qubit block; the state is unchanged
    n = 8
    rng = np.random.default_rng(11)
    for trial in range(30):
        gates = []
        for _ in range(14):
            kind = int(rng.integers(0, 5))
            a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
            if kind == 0:
                gates.append(ops.BitGate("matrix", (a,), 0, haar_unitary(rng, 2), False))
            elif kind == 1:
                gates.append(ops.BitGate("matrix", (a,), 0, H2.astype(np.complex128), False))
            elif kind == 2:
                gates.append(ops.BitGate("matrix", (a, b), 0, haar_unitary(rng, 4), False))
            elif kind == 3:
                gates.append(ops.BitGate("matrix", (a,), 1 << b, haar_unitary(rng, 2), False))            # controlled: never packed
            else:
                gates.append(ops.BitGate("matrix", (), (1 << a) | (1 << b), np.array([[np.exp(0.3j)]]), True))
        packed = ops.pack_lone_1q(gates)
        assert len(packed) <= len(gates)
        psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
        a_state, b_state = psi.copy(), psi.copy()
        for g in gates:
            a_state = bitsim.apply_bitgate(a_state, g, n)
        for g in packed:
            b_state = bitsim.apply_bitgate(b_state, g, n)
        assert float(np.max(np.abs(a_state - b_state))) <= 1e-13
    # two Hadamards with nothing in between on the second one's bit -> one REAL block
    two = ops.pack_lone_1q([ops.BitGate("matrix", (3,), 0, H2.astype(np.complex128), False),
                            ops.BitGate("matrix", (), 1 << 3, np.array([[1j]]), True),
                            ops.BitGate("matrix", (5,), 0, H2.astype(np.complex128), False)])
    assert len(two) == 2 and two[0].bits == (3, 5) and not two[0].mat.imag.any()
    # a QFT step (H followed by its run of controlled phases) keeps its Hadamard: the kernel rides the run on its sweep
    q = [s for mats in qfft_stream(10, rev=False) for s in (ops.simplify(g) for g in ops.decode_mats(mats, 10)) if s is not None]
    q = [ops.lower(g, 10) for g in ops.merge_blocks(q, 2, cost_aware=True)]
    kept = ops.pack_lone_1q(q)
    assert [id(g) for g in kept[:35]] == [id(g) for g in q[:35]] and sum(g.k == 2 and not g.diagonal for g in kept) <= 2


class _RecordingLib(object):
    """Stands in for libqipb200 in host-logic tests: every entry point records its call and returns a status."""

    def __init__(self, fill_status=0):
        self.calls = []
        self.fill_status = fill_status

    def __getattr__(self, name):
        if not name.startswith("qipb_"):
            raise AttributeError(name)

        def call(*args):
            self.calls.append((name, args))
            if name == "qipb_apply_fused_fill":
                return self.fill_status
            if name.endswith("_count"):
                return 0
            return 0
        return call

    def names(self):
        return [c[0] for c in self.calls if c[0] not in ("qipb_set_stream",)]


def _host_only_b200(n, fill_status=0, monkeypatch=None):
    """A B200Backend whose device side is cut off (recording library, CPU tensor as the buffer): the real
    _init_state / kronselect_dot / flush / _fill_first_pass control flow runs."""
    import torch
    from qip_b200 import backend as be
    b = object.__new__(be.B200Backend)
    b.L = _RecordingLib(fill_status)
    b.n, b.code, b.tdtype, b.amp_bytes, b.np_dtype = n, 0, torch.complex128, 16, np.dtype(np.complex128)
    b.device = -1                                  # torch.cuda.device(-1) is a no-op context
    b.fuse, b.strategy, b.tile_bits, b.min_low_bits = True, "tile", 12, 7
    b.relabel_swaps, b.pos = True, [n - 1 - q for q in range(n)]
    b.ctx, b.state, b.queue, b.plan_cache, b.profile = None, None, [], None, None
    b.stats = {"gates": 0, "passes": 0, "fused_passes": 0, "flushes": 0}
    b._pending_init, b.lazy_init = None, True
    b.host_state_max_qubits = 30
    b._stream = lambda: None
    monkeypatch.setattr(be, "feeds_to_device", lambda vfeeds, device: torch.zeros(4, dtype=torch.complex128))
    monkeypatch.setattr(torch, "empty", lambda *a, **k: torch.zeros(8, dtype=torch.complex128))
    return b


def test_lazy_product_state_init_rides_on_the_first_fused_pass(monkeypatch):
    n = 14
    groups = [[q] for q in range(n - 2)] + [[n - 2, n - 1]]
    feeds = [np.array([0.6, 0.8j]) for _ in range(n - 2)] + [2]            # one-qubit vectors + a one-hot pair
    stream = list(layered_stream(n, 2, 3))
    # 1. supported: no stand-alone init, the first pass goes through qipb_apply_fused_fill with n extra gates
    b = _host_only_b200(n, 0, monkeypatch)
    b._init_state(groups, feeds)
    assert b._pending_init is not None and b.L.names() == []
    for mats in stream:
        b.kronselect_dot(mats)
    b.flush()
    names = b.L.names()
    assert names[0] == "qipb_apply_fused_fill" and "qipb_init_kron" not in names and b._pending_init is None
    fill_args = b.L.calls[[c[0] for c in b.L.calls].index("qipb_apply_fused_fill")][1]
    assert fill_args[2] == n and fill_args[6] > n and b.stats["fill_passes"] == 1
    assert b.stats["passes"] == len([x for x in names if x.startswith("qipb_apply")])
    # 2. the library cannot serve it (status 3): stand-alone init first, then every pass the ordinary way
    b = _host_only_b200(n, 3, monkeypatch)
    b._init_state(groups, feeds)
    for mats in stream:
        b.kronselect_dot(mats)
    b.flush()
    names = b.L.names()
    assert names[:3] == ["qipb_apply_fused_fill", "qipb_init_kron", "qipb_apply_fused"] and b._pending_init is None
    assert "fill_passes" not in b.stats
    # 3. nothing queued when the state is observed: stand-alone init
    b = _host_only_b200(n, 0, monkeypatch)
    b._init_state(groups, feeds)
    b.flush()
    assert b.L.names() == ["qipb_init_kron"] and b._pending_init is None
    # 4. an entangled (multi-qubit vector) feed is initialised at once
    b = _host_only_b200(n, 0, monkeypatch)
    b._init_state([[0, 1]] + [[q] for q in range(2, n)], [np.ones(4) / 2] + [np.array([1.0, 0.0])] * (n - 2))
    assert b._pending_init is None and b.L.names() == ["qipb_init_kron"]
    # 5. any other library error in fill mode is raised, not swallowed
    from qip_b200 import lib as qlib
    b = _host_only_b200(n, 1, monkeypatch)
    monkeypatch.setattr(qlib, "check", lambda rc: (_ for _ in ()).throw(qlib.QipbError("boom")) if rc else None)
    b._init_state(groups, feeds)
    for mats in stream:
        b.kronselect_dot(mats)
    with pytest.raises(qlib.QipbError):
        b.flush()
