"""CPU tier: the graph-level op-stream compiler (qip_b200/graph.py), the native
one-hot / device feed path (backend.split_feeds), vectorised oracle functions (8f row 3) and the
DistributedBackend front door (8f row 4).

The device kernels cannot run here; compiled circuits are replayed on the tests-only numpy executor behind
the PRODUCT's planner (test_host_logic.TileHostBackend) and compared with the UNMODIFIED reference
(front-end + Cython kernels, oracle/_ref) on the same graphs -- the GPU tier replays compiled circuits on
the CUDA path (tests/test_gpu_parity.py)."""
import random

import numpy as np
import pytest

from oracle.ref_loader import have_ref_ext, have_reference_tree
from qip_b200.backend import split_feeds
from qip_b200.functions import controlled, equals, modexp, tabulated
from qip_b200.graph import CompiledCircuit, compile_circuit
from test_host_logic import TileHostBackend

needs_ref = pytest.mark.skipif(not (have_reference_tree() and have_ref_ext()),
                               reason="needs /root/reference and oracle/_ref (build container only)")


def _ref():
    from oracle.ref_loader import import_reference_qip
    import_reference_qip()
    import qip.operators as O
    import qip.pipeline as P
    import qip.qip as Q
    from qip.qfft import QFFT
    return O, P, Q, QFFT


def _classic_equal(a, b, tol=1e-12):
    if isinstance(a, tuple) and isinstance(b, tuple):
        return len(a) == len(b) and all(_classic_equal(x, y, tol) for x, y in zip(a, b))
    if isinstance(a, (list, np.ndarray)) or isinstance(b, (list, np.ndarray)):
        return np.allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), atol=tol, rtol=0)
    if isinstance(a, float) or isinstance(b, float):
        return abs(a - b) <= tol
    return a == b


@needs_ref
def test_compiled_cswap_matches_reference_run_for_every_seed():
    O, P, Q, _ = _ref()
    q1, q2, q3 = Q.Qubit(n=1), Q.Qubit(n=5), Q.Qubit(n=5)      # README.md:8-39 / tests/qiptest.py:194-228
    c1, c2, c3 = O.C(O.Swap)(O.H(q1), q2, q3)
    m = Q.Measure(O.H(c1))
    s2, s3 = np.zeros(32), np.zeros(32)
    s2[0] = s3[1] = 1.0
    feed = {q1: [1.0, 0.0], q2: s2, q3: s3}
    circ = compile_circuit(m, c2, c3, feed=feed)
    assert circ.n == 11 and circ.ngates >= 3
    for seed in range(8):
        random.seed(seed)
        want_state, want_c = P.run(m, c2, c3, feed=feed)
        random.seed(seed)
        got_state, got_c = circ.run(feed=feed, backend_constructor=TileHostBackend.make_state)
        assert got_c[m][0] == want_c[m][0] and abs(got_c[m][1] - want_c[m][1]) <= 1e-12
        assert np.allclose(got_state, want_state, atol=1e-12, rtol=0)
    # new feed values through the same compiled circuit
    s3b = np.zeros(32)
    s3b[0] = 1.0
    random.seed(1)
    want_state, want_c = P.run(m, c2, c3, feed={q1: [1.0, 0.0], q2: s2, q3: s3b})
    random.seed(1)
    got_state, got_c = circ.run(feed={q3: s3b}, backend_constructor=TileHostBackend.make_state)
    assert got_c[m][0] == want_c[m][0] == 0 and abs(got_c[m][1] - 1.0) <= 1e-12
    assert np.allclose(got_state, want_state, atol=1e-12, rtol=0)
    with pytest.raises(ValueError, match="not part of the compiled circuit"):
        circ.run(feed={Q.Qubit(n=1): [1.0, 0.0]}, backend_constructor=TileHostBackend.make_state)


@needs_ref
def test_compiled_grover_iteration_with_function_nodes_and_defaults():
    O, P, Q, _ = _ref()
    n, x0 = 6, 42
    q = Q.Qubit(n=n, default=np.ones(2 ** n) / np.sqrt(2 ** n))       # examples/grovers_iterative.py:33-39
    anc = Q.Qubit(n=1, default=[1 / np.sqrt(2), -1 / np.sqrt(2)])
    os_, oa = O.F(lambda x: int(x == x0), q, anc)
    fs, fa = O.F(lambda x: int(x == 0), O.H(os_), oa)
    ds, da = O.H(fs), fa
    sm, sa = Q.StochasticMeasure(ds), Q.StochasticMeasure(da)
    want_state, want_c = P.run(sm, sa)
    circ = compile_circuit(sm, sa)
    got_state, got_c = circ.run(backend_constructor=TileHostBackend.make_state)
    assert np.allclose(got_state, want_state, atol=1e-12, rtol=0)
    assert _classic_equal(got_c[sm], want_c[sm]) and _classic_equal(got_c[sa], want_c[sa])
    # iterate by re-feeding the whole register (tuple key), the working form of the example's loop
    circ2 = compile_circuit(sm, sa, feed={(q, anc): want_state})
    state_ref, state_got = want_state, got_state
    for _ in range(3):
        state_ref, c_ref = P.run(sm, sa, feed={(q, anc): state_ref})
        state_got, c_got = circ2.run(feed={(q, anc): state_got}, backend_constructor=TileHostBackend.make_state)
        assert np.allclose(state_got, state_ref, atol=1e-12, rtol=0)
        assert _classic_equal(c_got[sm], c_ref[sm])
    # the function tables were built once per F node, the gate segments planned once
    assert len(circ2._tables) == 2


@needs_ref
def test_compiled_shor_circuit_end_to_end():
    O, P, Q, QFFT = _ref()
    m_bits, n_bits, x, N = 8, 4, 7, 15                               # examples/shors.py:102-124
    reg1 = Q.Qubit(n=m_bits, default=np.ones(2 ** m_bits) * pow(2 ** m_bits, -0.5))
    reg2 = Q.Qubit(n=n_bits)
    u1, u2 = O.F(lambda i: pow(x, i, N), reg1, reg2)
    qft = QFFT(u1)
    out1, out2 = Q.Qubit(qft, u2).split()
    mq, m = Q.StochasticMeasure(out1), Q.Measure(qft)
    random.seed(5)
    want_state, want_c = P.run(mq, m)
    circ = compile_circuit(mq, m)
    random.seed(5)
    got_state, got_c = circ.run(backend_constructor=TileHostBackend.make_state)
    assert got_c[m][0] == want_c[m][0] and abs(got_c[m][1] - want_c[m][1]) <= 1e-12
    assert _classic_equal(got_c[mq], want_c[mq])
    assert np.allclose(got_state, want_state, atol=1e-12, rtol=0)
    # period 4 of 7^i mod 15: the distribution peaks at multiples of 2^8 / 4 (StochasticMeasure's index is
    # little-endian over the measured qubits, SURVEY 8g-2: reverse the 8 bits to read the register value)
    peaks = np.argsort(-np.asarray(got_c[mq]))[:4]
    assert sorted(int("{:08b}".format(int(p))[::-1], 2) % 64 for p in peaks) == [0, 0, 0, 0]
    # int default -> native one-hot feed in the compiled circuit (never a 2^k vector on replay)
    r2 = Q.Qubit(n=n_bits, default=3)
    v1, v2 = O.F(lambda i: pow(x, i, N), reg1, r2)
    circ3 = compile_circuit(Q.StochasticMeasure(v2))
    assert any(isinstance(f, int) and f == 3 for f in circ3.default_feeds)


def test_compiled_circuit_from_plain_op_list():
    from qip_b200.circuits import H2, qfft_stream
    from oracle import oracle as orc
    n = 6
    rng = np.random.default_rng(0)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    ops_ = [("k", mats) for mats in qfft_stream(n)] + [("p", [0, 1, 2], 0), ("f", [0, 1, 2], [3, 4, 5], lambda x: (x * 3) % 8),
                                                     ("k", {5: H2}), ("m", [0, 5]), ("p", [4, 2], 3)]
    circ = CompiledCircuit.from_ops(n, [list(range(n))], [psi], ops_)
    c = orc.OracleBackend.make_state(n, [list(range(n))], [psi])
    want = {}
    random.seed(3)
    for i, op in enumerate(ops_):
        if op[0] == "k":
            c.kronselect_dot(op[1])
        elif op[0] == "f":
            c.func_apply(np.array(op[1], dtype=np.int32), np.array(op[2], dtype=np.int32), op[3])
        elif op[0] == "m":
            want[i] = c.measure(np.array(op[1], dtype=np.int32))
        else:
            want[i] = c.measure_probabilities(np.array(op[1], dtype=np.int32), top_k=op[2])
    for _ in range(2):                                           # second replay uses the cached tables
        random.seed(3)
        state, classic = circ.run(backend_constructor=TileHostBackend.make_state)
        assert np.allclose(state, c.get_state(), atol=1e-12, rtol=0)
        assert set(classic) == set(want)
        for i in want:
            assert _classic_equal(classic[i], want[i]), i


def test_split_feeds_fixes_bits_for_one_hot_and_unfed_qubits():
    n = 6
    bit = lambda q: n - 1 - q
    vg, vf, mask, value = split_feeds([[0, 3], [4], [2, 1]], [2, [0.6, 0.8], 1], n, bit)
    assert vg == [[4]] and np.allclose(vf[0], [0.6, 0.8])
    # qubits 0,3 fixed to (1,0); qubits 2,1 fixed to (0,1); qubit 5 un-fed -> 0
    assert mask == sum(1 << bit(q) for q in (0, 3, 2, 1, 5))
    assert value == (1 << bit(0)) | (1 << bit(1))
    with pytest.raises(ValueError):
        split_feeds([[0, 1]], [4], n, bit)
    with pytest.raises(ValueError):
        split_feeds([[0, 1]], [[1, 0, 0]], n, bit)
    vg, vf, mask, value = split_feeds([], [], 3, lambda q: 2 - q)
    assert vg == [] and mask == 7 and value == 0


def test_vectorised_oracle_functions():
    from qip_b200.backend import tabulate
    t = tabulate(modexp(11, 21), 9)
    assert all(int(t[i]) == pow(11, i, 21) for i in range(512))
    assert modexp(11, 21)(5) == pow(11, 5, 21)
    t = tabulate(equals(42), 10)
    assert t[42] == 1 and int(t.sum()) == 1
    calls = []

    def slow(x):
        calls.append(x)
        return int(x) % 3
    tb = tabulated(slow, 5)
    ncalls = len(calls)
    assert (tabulate(tb, 5) == np.arange(32) % 3).all() and len(calls) == ncalls      # the table is reused, no new calls
    with pytest.raises(ValueError):
        modexp(3, 2 ** 40)


def test_controlled_function_is_a_plain_F_on_the_joined_register():
    # SURVEY 8f row 3 (controlled-F; the reference has none): controls first, then reg1; f(x) only where every control is 1
    from oracle import oracle as orc
    from qip_b200.backend import tabulate
    f = lambda x: (5 * x + 3) % 8
    cf = controlled(f, 2, 3)
    t = tabulate(cf, 5)
    assert all(int(t[(c << 3) | x]) == (f(x) if c == 3 else 0) for c in range(4) for x in range(8))
    assert cf(0b11101) == f(0b101) and cf(0b10101) == 0
    tv = tabulate(controlled(modexp(7, 15), 1, 4), 5)                      # vectorised inner function
    assert all(int(tv[16 + x]) == pow(7, x, 15) for x in range(16)) and not tv[:16].any()
    # against the oracle's func_apply: control qubit 6 (not adjacent), reg1 = [0, 3, 1], reg2 = [2, 4, 5]
    n = 7
    rng = np.random.default_rng(3)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    reg1, reg2, ctrl = [0, 3, 1], [2, 4, 5], 6
    for make in (orc.OracleBackend.make_state, TileHostBackend.make_state):
        b = make(n, [list(range(n))], [psi])
        b.func_apply(np.array([ctrl] + reg1, dtype=np.int32), np.array(reg2, dtype=np.int32), controlled(f, 1, 3))
        got = np.asarray(b.get_state())
        plain = orc.OracleBackend.make_state(n, [list(range(n))], [psi])
        plain.func_apply(reg1, reg2, f)
        applied = plain.get_state()
        on = np.array([(i >> (n - 1 - ctrl)) & 1 for i in range(2 ** n)], dtype=bool)     # qubit q = index bit n-1-q
        assert np.allclose(got[on], applied[on], atol=1e-15, rtol=0) and np.allclose(got[~on], psi[~on], atol=1e-15, rtol=0)
    with pytest.raises(ValueError):
        controlled(f, -1, 3)


@needs_ref
def test_controlled_function_through_the_reference_front_end():
    # the same function object works in the UNMODIFIED reference (python-int calls, func_apply.pyx:66-69) and compiled here
    O, P, Q, QFFT = _ref()
    f = lambda x: (3 * x + 1) % 8
    rng = np.random.default_rng(8)
    psi = rng.normal(size=16) + 1j * rng.normal(size=16)
    psi /= np.linalg.norm(psi)
    joined = Q.Qubit(n=4, default=psi)                               # qubit 0 = control, qubits 1..3 = x
    phi = rng.normal(size=8) + 1j * rng.normal(size=8)
    phi /= np.linalg.norm(phi)
    reg2 = Q.Qubit(n=3, default=phi)
    u1, u2 = O.F(controlled(f, 1, 3), joined, reg2)
    want, _ = P.run(u1, u2)
    got, _ = compile_circuit(u1, u2).run(backend_constructor=TileHostBackend.make_state)
    assert np.allclose(got, want, atol=1e-15, rtol=0)
    # control = 0 half: the product state is untouched; control = 1 half: reg2 index q -> q xor f(x)
    amp = np.asarray(got).reshape(2, 8, 8)
    for x in range(8):
        assert np.allclose(amp[0, x], psi[x] * phi, atol=1e-15, rtol=0)
        assert np.allclose(amp[1, x, np.arange(8) ^ f(x)], psi[8 + x] * phi, atol=1e-15, rtol=0)


def test_distributed_front_door_conventions(monkeypatch):
    import qip_b200.distributed as D
    monkeypatch.setattr(D, "_engine_factory", lambda: TileHostBackend)
    psi = np.zeros(8)
    psi[3], psi[6] = 0.6, 0.8                                      # SURVEY 8c probe: 0.6|011> + 0.8|110>
    b = D.DistributedBackend.make_state(3, [[0, 1, 2]], [psi])
    idx, probs = b.measure_probabilities(np.array([1, 2], dtype=np.int32))           # top_k defaults to all outcomes
    assert idx[:2] == [2, 3] and np.allclose(probs, [0.64, 0.36, 0, 0])
    idx, probs = b.measure_probabilities(np.array([1, 2], dtype=np.int32), top_k=1)
    assert idx == [2] and np.allclose(probs, [0.64])
    b.kronselect_dot({0: np.array([[0, 1], [1, 0]])})             # everything else is the engine's
    assert abs(np.asarray(b.get_state())[7]) == pytest.approx(0.6)
    assert b.n == 3
