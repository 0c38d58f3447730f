"""CPU tier: pins the oracle (oracle/qip_oracle.c) against
  (a) the golden op streams recorded from the unmodified reference (tests/golden/),
  (b) the reference's own compiled Cython kernels (oracle/_ref) on seeded random inputs,
  (c) the kernel-level known-answer tests of the reference's tests/utiltest.py, restated here
      (they call qip.util.kronselect_dot directly rather than run(), so they are not in (a)).
"""
import random

import numpy as np
import pytest

import replay
from oracle import oracle as orc
from oracle.ref_loader import have_ref_ext
from qip_b200.mats import CMat, SwapMat

META, STREAMS, ARRAYS = replay.load_streams()


def test_golden_covers_reference_suite():
    labels = [s["label"] for s in STREAMS]
    assert META["n_reference_tests"] == 37
    for name in ("test_bell", "test_cswap_5bitcompare", "test_many_bits", "test_toffoli", "test_fop",
                 "test_measure_stochastic_top", "test_cosine_pipeline", "test_rop_tuplefeed"):
        assert any(l.endswith("::" + name) for l in labels), name
    assert sum(l.startswith("config/") for l in labels) >= 15
    assert sum(l.startswith("unpinned/") for l in labels) >= 20


@pytest.mark.parametrize("i", range(len(STREAMS)), ids=[s["label"] for s in STREAMS])
def test_oracle_replays_golden_stream(i):
    s = STREAMS[i]
    assert replay.replay(s, ARRAYS, orc.OracleBackend.make_state, tol=1e-13)


needs_ref = pytest.mark.skipif(not have_ref_ext(), reason="oracle/_ref not built")


def _rand_state(rng, n):
    v = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    return v / np.linalg.norm(v)


def _rand_mats(rng, n):
    """A random op in the boundary vocabulary: dense 1-3 qubit, nested CMat, SwapMat, multi-entry."""
    qs = list(rng.permutation(n))
    kind = int(rng.integers(0, 6))
    if kind == 0:
        return {int(qs[0]): rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))}
    if kind == 1:
        k = int(rng.integers(2, 4))
        return {tuple(int(q) for q in qs[:k]): rng.normal(size=(2 ** k, 2 ** k)) + 1j * rng.normal(size=(2 ** k, 2 ** k))}
    if kind == 2:
        return {tuple(int(q) for q in qs[:2]): CMat(rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)))}
    if kind == 3:
        return {tuple(int(q) for q in qs[:3]): CMat(CMat(rng.normal(size=(2, 2)) + 0j))}
    if kind == 4:
        w = int(rng.integers(1, 3))
        return {tuple(int(q) for q in qs[:2 * w + 1]): CMat(SwapMat(w))} if rng.random() < 0.5 else \
               {tuple(int(q) for q in qs[:2 * w]): SwapMat(w)}
    return {int(qs[0]): rng.normal(size=(2, 2)) + 0j, int(qs[1]): rng.normal(size=(2, 2)) + 0j}


@needs_ref
@pytest.mark.parametrize("seed", range(12))
def test_oracle_matches_compiled_reference_gates(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(5, 9))
    psi = _rand_state(rng, n)
    a = orc.OracleBackend.make_state(n, [list(range(n))], [psi])
    b = orc.RefBackend.make_state(n, [list(range(n))], [psi])
    for _ in range(6):
        mats = _rand_mats(rng, n)
        a.kronselect_dot(mats)
        b.kronselect_dot(mats)
    assert replay.close(a.get_state(), b.get_state(), 1e-13)


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_oracle_matches_compiled_reference_measurement(seed):
    rng = np.random.default_rng(100 + seed)
    n = 6
    psi = _rand_state(rng, n)
    a = orc.OracleBackend.make_state(n, [list(range(n))], [psi])
    b = orc.RefBackend.make_state(n, [list(range(n))], [psi])
    for idx in ([0], [5, 2], [1, 3, 4], [4, 1]):
        assert np.allclose(a.measure_probabilities(idx), b.measure_probabilities(np.array(idx, dtype=np.int32)), rtol=0, atol=1e-15)
        random.seed(seed)
        ma, pa = a.soft_measure(idx)
        random.seed(seed)
        mb, pb = b.soft_measure(idx)
        assert ma == mb and abs(pa - pb) < 1e-15
    random.seed(seed)
    ma, pa = a.measure([2, 4])
    random.seed(seed)
    mb, pb = b.measure(np.array([2, 4], dtype=np.int32))
    assert ma == mb and abs(pa - pb) < 1e-15
    assert replay.close(a.get_state(), b.get_state(), 1e-13)


@needs_ref
def test_oracle_matches_compiled_reference_func_apply():
    rng = np.random.default_rng(5)
    n = 6
    psi = _rand_state(rng, n)
    a = orc.OracleBackend.make_state(n, [list(range(n))], [psi])
    b = orc.RefBackend.make_state(n, [list(range(n))], [psi])
    f = lambda x: (3 * x + 1) % 8
    a.func_apply([4, 0, 2], [5, 1, 3], f)
    b.func_apply([4, 0, 2], [5, 1, 3], f)
    assert np.array_equal(a.get_state(), b.get_state())


# ---- restated KATs of the reference's tests/utiltest.py -------------------------------------
def test_kat_simple_and_nonsym_kron():            # tests/utiltest.py:9-43
    eye = np.eye(2)
    for h in (np.array([[1, 1], [1, -1]]), np.array([[1, 2], [3, 4]])):
        hh = np.kron(h, h)
        ref = np.kron(h, np.kron(eye, h))
        for i in range(8):
            v = np.zeros(8, dtype=np.complex128)
            v[i] = 1
            t1 = np.zeros(8, dtype=np.complex128)
            t2 = np.zeros(8, dtype=np.complex128)
            orc.cdot({0: h, 2: h}, v, 3, t1)
            orc.cdot({(0, 2): hh}, v, 3, t2)
            assert np.array_equal(ref[:, i], t1) and np.array_equal(ref[:, i], t2)


def test_kat_index_order():                        # tests/utiltest.py:45-66
    a = np.array([[1, 2], [3, 4]])
    b = np.array([[5, 6], [7, 8]])
    ref = np.kron(a, np.kron(np.eye(2), b))
    for i in range(8):
        v = np.zeros(8, dtype=np.complex128)
        v[i] = 1
        outs = [np.zeros(8, dtype=np.complex128) for _ in range(3)]
        orc.cdot({0: a, 2: b}, v, 3, outs[0])
        orc.cdot({(0, 2): np.kron(a, b)}, v, 3, outs[1])
        orc.cdot({(2, 0): np.kron(b, a)}, v, 3, outs[2])
        for o in outs:
            assert np.array_equal(ref[:, i], o)


def test_kat_output_and_input_offsets():           # tests/utiltest.py:68-125
    h = np.array([[1, 1], [1, -1]])
    hh = np.kron(h, h)
    ref = np.kron(h, np.kron(np.eye(2), h))
    for i in range(8):
        v = np.zeros(8, dtype=np.complex128)
        v[i] = 1
        for off in range(4):
            t1 = np.zeros(4, dtype=np.complex128)
            t2 = np.zeros(4, dtype=np.complex128)
            orc.cdot({0: h, 2: h}, v, 3, t1, output_offset=off)
            orc.cdot({(0, 2): hh}, v, 3, t2, output_offset=off)
            assert np.array_equal(ref[off:off + 4, i], t1) and np.array_equal(ref[off:off + 4, i], t2)
        acc = np.zeros(8, dtype=np.complex128)
        for off in range(0, 8, 2):
            tmp = np.zeros(8, dtype=np.complex128)
            orc.cdot({0: h, 2: h}, v[off:off + 2].copy(), 3, tmp, input_offset=off)
            acc += tmp
        assert np.array_equal(ref[:, i], acc)


def test_kat_func_apply():                         # tests/utiltest.py:127-153
    state = np.zeros(16, dtype=np.complex128)
    state[0], state[8], state[4], state[12] = 1, 2, 3, 4
    b = orc.OracleBackend(4, state)
    b.func_apply([0, 1], [2, 3], lambda x: (x + 1) % 4)
    out = b.get_state()
    assert out[1] == 1 and out[11] == 2 and out[6] == 3 and out[12] == 4


def test_validation_errors_match_reference():      # qip/util.py:29-58, kronprod.pyx:114-116, 402-407
    v = np.zeros(8, dtype=np.complex128)
    o = np.zeros(8, dtype=np.complex128)
    with pytest.raises(ValueError):
        orc.cdot({0: np.eye(2)}, np.zeros(16, dtype=np.complex128), 3, o)
    with pytest.raises(Exception, match="Type of indices"):
        orc.cdot({"a": np.eye(2)}, v, 3, o)
    with pytest.raises(Exception, match="Shape of square submatrix"):
        orc.cdot({(0, 1): np.eye(2)}, v, 3, o)
    with pytest.raises(ValueError, match="not numpy"):
        class Odd:
            shape = (2, 2)
        orc.cdot({0: Odd()}, v, 3, o)
    b = orc.OracleBackend(3, v.copy())
    with pytest.raises(ValueError):
        b.measure([0], measured=2)
    with pytest.raises(ValueError):
        b.measure([0], measured=0, measured_prob=0.0)
