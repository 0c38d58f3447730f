"""CPU tier: the fused pass's C++ lowering and sweep arithmetic, executed on the host.

tests/csrc/fused_emul.cu (test infrastructure, built here with nvcc) includes qip_b200/csrc/fused.cu and runs
qipb_apply_fused's real lowering (diagonal runs -> stage tables, structured 2-qubit block forms, stages riding on a
dense 1-qubit sweep, splitting into launches) and the real sweep functions on a host tile, "thread" by "thread".
The passes come from the product planner (qip_b200.ops.plan) and are packed by the product's pack_pass; the result
is compared with the numpy bit simulator (tests/bitsim.py), which applies the same BitGates one by one.

Not covered here (device only): TMA staging, mbarriers, the grid loop -- tests/test_gpu_parity.py."""
import ctypes
import os
import sys

import numpy as np
import pytest

import bitsim
from qip_b200 import lib as qlib
from qip_b200 import ops
from qip_b200.backend import pack_pass
from qip_b200.circuits import H2, X2, haar_unitary, layered_stream, qfft_stream, rm_mat
from qip_b200.mats import CMat, SwapMat
from qip_b200.ops import BitGate, Pass

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc"))
import build_emul  # noqa: E402


@pytest.fixture(scope="module")
def emul():
    L = ctypes.CDLL(build_emul.build())
    L.qipb_emul_fused.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                  ctypes.c_int, ctypes.POINTER(qlib.Gate), ctypes.POINTER(ctypes.c_int)]
    L.qipb_emul_fused.restype = ctypes.c_int
    L.qipb_emul_fused_fill.argtypes = L.qipb_emul_fused.argtypes
    L.qipb_emul_fused_fill.restype = ctypes.c_int
    L.qipb_emul_last_error.restype = ctypes.c_char_p
    return L


def logical_gates(stream, n):
    gates = []
    for mats in stream:
        for g in ops.decode_mats(mats, n):
            s = ops.simplify(g)
            if s is not None:
                gates.append(s)
    return gates


def run_emulated(L, state, passes, n, dtype):
    """Fused passes through the emulator, stand-alone passes through the bit simulator."""
    code = qlib.C128 if dtype == np.complex128 else qlib.C64
    info_total = np.zeros(16, dtype=np.int64)
    st = np.ascontiguousarray(state, dtype=dtype)
    for p in passes:
        if not p.fused:
            st = np.ascontiguousarray(bitsim.run_passes(st.astype(np.complex128), [p], n), dtype=dtype)
            continue
        arr, tbits = pack_pass(p)
        info = (ctypes.c_int * 16)()
        rc = L.qipb_emul_fused(st.ctypes.data_as(ctypes.c_void_p), n, code, len(p.tile_bits), tbits, len(p.gates), arr, info)
        assert rc == 0, L.qipb_emul_last_error()
        info_total += np.array(list(info))
    return st, info_total


def random_state(n, seed):
    rng = np.random.default_rng(seed)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    return psi / np.linalg.norm(psi)


def check(L, stream, n, seed=0, dtype=np.complex128, tile_bits=12, min_low_bits=7):
    gates = logical_gates(stream, n)
    passes, _ = ops.plan(gates, n, 16 if dtype == np.complex128 else 8, strategy="tile", tile_bits=tile_bits,
                         min_low_bits=min_low_bits)
    psi = random_state(n, seed)
    want = bitsim.run_passes(psi.copy(), passes, n)
    got, info = run_emulated(L, psi, passes, n, dtype)
    tol = 1e-12 if dtype == np.complex128 else 2e-5
    err = float(np.max(np.abs(got - want))) / float(np.max(np.abs(want)))
    assert err <= tol, err
    return info


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("seed", range(3))
def test_emulated_layered_passes_match_bit_simulator(emul, dtype, seed):
    # production-shaped launches (2^12 tiles, 2 KiB runs -> the specialised sweeps): structured block forms, lone
    # diagonal gates, clustered diagonal runs, low-bit (bank-conflict-free) sweeps
    info = check(emul, layered_stream(14, 3, seed), 14, seed, dtype)
    assert info[1] == info[0] >= 3 and info[4] >= 1, info
    assert info[9] == info[0] and info[8] >= 2, info             # WIDE launches; block pairs share sweeps


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("seed", range(3))
def test_emulated_block_pairs_match_unpaired_passes(emul, monkeypatch, dtype, seed):
    # two dense 2-qubit blocks per sweep (sweep_pair2) against the same passes run block by block, in every combination
    # of matrix forms (general / real / real x phases / monomial) and with partners hoisted over commuting ops
    n = 15
    rng = np.random.default_rng(100 + seed)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi = (psi / np.linalg.norm(psi)).astype(dtype)
    hh = np.kron(H2, H2)
    cx = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)
    sw = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)
    forms = [lambda: haar_unitary(rng, 4), lambda: hh @ cx, lambda: (hh @ sw) @ np.diag(np.exp(1j * rng.uniform(0, 6, 4))),
             lambda: cx @ np.diag([1, 1j, 1, np.exp(0.3j)]), lambda: sw]
    tile = list(range(7)) + [8, 10, 11, 13, 14]
    hi = [b for b in tile if b >= 3]
    gates = []
    for rep in range(14):
        bits = [int(b) for b in rng.choice(hi, size=4, replace=False)]
        a, b = forms[int(rng.integers(0, 5))](), forms[int(rng.integers(0, 5))]()
        gates.append(BitGate("matrix", (bits[0], bits[1]), 0, np.ascontiguousarray(a)))
        if rep % 3 == 0:                                       # ops in between that commute / do not commute with the partner
            gates.append(BitGate("matrix", (int(rng.integers(0, n)),), 0, np.diag(np.exp(1j * rng.uniform(0, 6, 2))), True))
        if rep % 4 == 1:
            gates.append(BitGate("matrix", (bits[2],), 1 << 12, np.ascontiguousarray(rm_mat(2)), True))
        if rep % 5 == 2:
            gates.append(BitGate("matrix", (int(rng.choice(tile)),), 0, H2.astype(np.complex128)))
        gates.append(BitGate("matrix", (bits[2], bits[3]), (1 << 9) if rep % 6 == 3 else 0, np.ascontiguousarray(b)))
    p = Pass(True, gates, tuple(tile))
    monkeypatch.setenv("QIPB_FUSED_PAIR", "1")
    paired, info = run_emulated(emul, psi.copy(), [p], n, dtype)
    assert info[8] >= 6 and info[9] == info[0], info
    monkeypatch.setenv("QIPB_FUSED_PAIR", "0")
    single, info0 = run_emulated(emul, psi.copy(), [p], n, dtype)
    assert info0[8] == 0, info0
    tol = 1e-13 if dtype == np.complex128 else 2e-5
    assert np.max(np.abs(paired - single)) <= tol * np.max(np.abs(single))
    want = bitsim.run_passes(psi.astype(np.complex128), [p], n)
    assert np.max(np.abs(paired - want)) <= (1e-12 if dtype == np.complex128 else 1e-5) * np.max(np.abs(want))


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_emulated_qfft_stages_ride_on_the_hadamard_sweeps(emul, dtype):
    info = check(emul, qfft_stream(15), 15, 1, dtype)
    assert info[2] >= 10 and info[3] + 2 * info[5] + 4 * info[12] >= 10, info    # (grouped steps carry post = 2 / 3 / 9 instead of 1)


def test_emulated_small_tiles_take_the_generic_sweeps(emul):
    # tiny states / short runs: the non-UNI kernel variant (generic loops, ExpandAny)
    for n, tb, low in ((6, 5, 2), (9, 8, 3), (12, 10, 4)):
        info = check(emul, list(layered_stream(n, 2, n)) + list(qfft_stream(n)), n, n, tile_bits=tb, min_low_bits=low)
        assert info[1] == 0, info


def test_emulated_controls_everywhere_and_many_in_tile_controls(emul):
    n = 14
    rng = np.random.default_rng(4)
    u2, u4 = haar_unitary(rng, 2), haar_unitary(rng, 4)
    stream = [{(13, 0): CMat(X2)}, {(0, 13): CMat(X2)}, {(5, 13, 2): CMat(CMat(u2))}, {(1, 2, 12, 11): CMat(CMat(u4))},
              {(13, 12, 11, 10, 9, 8): CMat(CMat(CMat(CMat(CMat(u2)))))}, {(12, 13, 10, 11): CMat(CMat(SwapMat(1)))},
              {3: rm_mat(2)}, {(0, 1): CMat(rm_mat(3))}, {(13, 0): CMat(rm_mat(5))}, {(2, 7): np.diag(np.exp(1j * rng.normal(size=4)))},
              {(6, 7): u4}, {(7, 6): u4}, {(12, 13): u4}, {(13, 11): haar_unitary(rng, 4)}, {13: u2}, {12: H2}, {11: H2}]
    check(emul, stream, n, 2)
    check(emul, stream, n, 3, np.complex64)


def test_emulated_structured_forms_are_exact(emul):
    # every structured form of a merged 2-qubit block (real, real x column phases, monomial, general) on high, low
    # and mixed tile bits; each segment is merged into ONE block, all blocks run in one fused pass
    n = 13
    rng = np.random.default_rng(8)
    for a, b in ((12, 11), (12, 0), (1, 0), (5, 9), (2, 12)):
        qa, qb = n - 1 - a, n - 1 - b
        segments = [[{qa: H2}, {qb: H2}, {(qa, qb): CMat(X2)}],                                     # real
                    [{qa: rm_mat(3)}, {qb: H2}, {(qb, qa): CMat(X2)}],                             # real x column phases
                    [{qa: rm_mat(2)}, {qb: rm_mat(5)}, {(qa, qb): CMat(X2)}, {(qa, qb): SwapMat(1)}],   # monomial
                    [{(qa, qb): haar_unitary(rng, 4)}]]                                             # general
        blocks = []
        for seg in segments:
            merged = ops.merge_blocks(logical_gates(seg, n), 2)
            assert len(merged) == 1 and merged[0].k == 2 and not merged[0].diagonal
            blocks.append(ops.lower(merged[0], n))
        need = set()
        for g in blocks:
            need |= set(g.bits)
        p = ops.Pass(True, blocks, ops.choose_tile(need, n, 12))
        psi = random_state(n, a * 13 + b)
        want = bitsim.run_passes(psi.copy(), [p], n)
        got, info = run_emulated(emul, psi, [p], n, np.complex128)
        assert float(np.max(np.abs(got - want))) <= 1e-13
        assert info[4] == 3, info                          # three structured blocks, one general


def test_emulator_reports_lowering_errors(emul):
    g = (qlib.Gate * 1)()
    g[0].k = 1
    g[0].bits[0] = 12                      # non-diagonal target that is not a tile bit
    g[0].mat[0] = 1.0
    st = np.zeros(2 ** 13, dtype=np.complex128)
    info = (ctypes.c_int * 16)()
    rc = emul.qipb_emul_fused(st.ctypes.data_as(ctypes.c_void_p), 13, qlib.C128, 12, qlib.int_array(range(12)), 1, g, info)
    assert rc != 0 and b"not a tile bit" in emul.qipb_emul_last_error()


# ---- the opt-in forms (QIPB_FUSED_EXT=1): real 1-qubit gates and paired QFT steps ----
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n", [13, 15, 16])
def test_emulated_ext_qfft_pairs_two_steps_per_sweep(emul, monkeypatch, dtype, n):
    monkeypatch.setenv("QIPB_FUSED_EXT", "1")
    monkeypatch.setenv("QIPB_FUSED_QFT4", "0")                 # (the radix-16 form has its own test)
    info = check(emul, qfft_stream(n), n, n, dtype)
    assert info[5] >= 3 and info[6] >= 2 * info[5] and info[7] >= 1, info
    monkeypatch.setenv("QIPB_FUSED_EXT", "0")
    info = check(emul, qfft_stream(n), n, n, dtype)
    assert info[5] == 0 and info[6] == 0, info


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_emulated_ext_real_gates_in_layered_and_mixed_passes(emul, monkeypatch, dtype):
    monkeypatch.setenv("QIPB_FUSED_EXT", "1")
    n = 14
    check(emul, layered_stream(n, 3, 5), n, 5, dtype)        # (most 1-qubit gates of a layer merge into 2-qubit blocks)
    # real gates that are not Hadamards, a Hadamard whose stage has an outside control (no pairing), controlled
    # real gates (general path), a QFT on a sub-register placed on high and on low tile bits
    rng = np.random.default_rng(3)
    ry = np.array([[np.cos(0.3), -np.sin(0.3)], [np.sin(0.3), np.cos(0.3)]])
    stream = [{0: ry}, {1: X2}, {(2, 3): CMat(ry)}, {13: ry}, {12: H2}, {5: H2}]
    stream += [{(q, 5): CMat(rm_mat(2 + q % 3))} for q in (0, 1, 2, 9)]
    stream += [{6: H2}] + [{(q, 6): CMat(rm_mat(3))} for q in (7, 8, 10)]
    stream += list(qfft_stream(6, first_qubit=1)) + list(qfft_stream(5, first_qubit=9)) + [{(0, 4): haar_unitary(rng, 4)}]
    info = check(emul, stream, n, 6, dtype)
    assert info[6] >= 3, info


# ---- randomised passes: every gate shape the C ABI accepts, any tile-bit set, both kernels' op limits ----
def _random_pass(rng, n, tb, lowrun, ngates):
    tile = list(range(lowrun)) + sorted(int(b) for b in rng.choice(np.arange(lowrun, n), size=tb - lowrun, replace=False))
    gates = []

    def ctrl(exclude, nctrl):
        others = [b for b in range(n) if b not in exclude]
        cm = 0
        for c in rng.choice(others, size=min(nctrl, len(others)), replace=False):
            cm |= 1 << int(c)
        return cm

    for _ in range(ngates):
        kind = int(rng.integers(0, 8))
        nctrl = int(rng.choice([0, 0, 0, 1, 1, 2, 3]))
        if kind == 7:                                          # swap of two tile bits, possibly controlled
            t = [int(x) for x in rng.choice(tile, size=2, replace=False)]
            gates.append(ops.BitGate("swap", tuple(t), ctrl(t, nctrl)))
            continue
        diag = kind >= 4
        if kind <= 1:                                          # dense 1-qubit (general / Hadamard)
            t = [int(rng.choice(tile))]
            mat = haar_unitary(rng, 2) if kind == 0 else H2.astype(np.complex128)
        elif kind <= 3:                                        # dense 2-qubit (general / real)
            t = [int(x) for x in rng.choice(tile, size=2, replace=False)]
            mat = haar_unitary(rng, 4)
            if kind == 3:
                mat = np.real(mat) + 0j
        elif kind == 4:                                        # controlled scalar phase (what Rm / C(Rm) simplify to)
            t, mat, nctrl = [], np.array([[np.exp(1j * rng.normal())]]), max(1, nctrl)
        elif kind == 5:                                        # diagonal gates on ANY bits, inside or outside the tile
            t, mat = [int(rng.integers(0, n))], np.diag(np.exp(1j * rng.normal(size=2)))
        else:
            t, mat = [int(x) for x in rng.choice(n, size=2, replace=False)], np.diag(np.exp(1j * rng.normal(size=4)))
        gates.append(ops.BitGate("matrix", tuple(t), ctrl(t, nctrl), np.ascontiguousarray(mat, dtype=np.complex128), diag))
    return ops.Pass(True, gates, tuple(tile))


@pytest.mark.parametrize("seed", range(48))
def test_emulated_random_passes(emul, monkeypatch, seed):
    rng = np.random.default_rng(1000 + seed)
    mode = seed % 4
    if mode == 0:                                              # production shape: 2^12 tiles, long runs
        n, tb, lowrun = int(rng.choice([13, 14])), 12, int(rng.choice([5, 6, 7, 8]))
    elif mode == 1:                                            # 2^11 tiles (128-thread CTAs)
        n, tb, lowrun = int(rng.choice([12, 13])), 11, int(rng.choice([5, 6, 7]))
    elif mode == 2:                                            # tiny states, any tile
        n = int(rng.integers(3, 11))
        tb = int(rng.integers(2, n + 1))
        lowrun = int(rng.integers(0, tb + 1))
    else:                                                      # short runs (no TMA staging), or one single tile
        n, tb, lowrun = (13, 12, int(rng.choice([0, 2, 4]))) if seed % 8 == 3 else (12, 12, 12)
    ngates = int(rng.choice([2, 5, 15, 60, 130]))              # 130 > FUSED_MAX_OPS: split into several launches
    dtype = np.complex128 if seed % 3 else np.complex64
    monkeypatch.setenv("QIPB_FUSED_EXT", str(seed % 2))
    p = _random_pass(rng, n, tb, lowrun, ngates)
    psi = random_state(n, seed)
    want = bitsim.run_passes(psi.copy(), [p], n)
    got, _ = run_emulated(emul, psi, [p], n, dtype)
    err = float(np.max(np.abs(got - want))) / float(np.max(np.abs(want)))
    assert err <= (1e-12 if dtype == np.complex128 else 5e-5), err


# ---- fill mode (qipb_apply_fused_fill): product-state init + first pass in one write-only sweep ----
def _fill_args(factors, p):
    """What B200Backend._fill_first_pass hands to the library (the product's own packing)."""
    from qip_b200.backend import pack_fill_pass
    arr, tbits = pack_fill_pass(factors, p)
    return arr, tbits, len(factors) + len(p.gates)


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("seed", range(4))
def test_emulated_fill_mode_builds_the_product_state_inside_the_first_pass(emul, dtype, seed):
    from qip_b200.backend import product_state_factors, split_feeds
    n = 14
    rng = np.random.default_rng(seed)
    # one-qubit vector feeds, a one-hot int feed on three qubits, two un-fed qubits (|0>)
    qubits = [int(q) for q in rng.permutation(n)]
    hot_group, unfed, vec = qubits[:3], qubits[3:5], qubits[5:]
    groups = [[q] for q in vec] + [hot_group]
    feeds = [rng.normal(size=2) + 1j * rng.normal(size=2) for _ in vec] + [5]
    vgroups, vfeeds, fmask, fval = split_feeds(groups, feeds, n, lambda q: n - 1 - q)
    factors = product_state_factors(vgroups, vfeeds, fmask, fval, n)
    assert factors is not None and len(factors) == n
    psi = np.ones(1, dtype=np.complex128)
    for b in range(n - 1, -1, -1):                         # index bit b: most significant first
        psi = np.kron(psi, np.array(factors[b]))
    want0 = orc_state(n, groups, feeds)
    assert float(np.max(np.abs(psi - want0))) <= 1e-13 * float(np.max(np.abs(want0)))
    stream = list(layered_stream(n, 1, seed)) if seed % 2 else list(qfft_stream(n))
    gates = logical_gates(stream, n)
    passes, _ = ops.plan(gates, n, 16 if dtype == np.complex128 else 8, strategy="tile")
    assert passes[0].fused
    want = bitsim.run_passes(psi.copy(), passes[:1], n)
    st = np.full(2 ** n, np.nan + 1j * np.nan, dtype=dtype)            # the buffer's content must never be read
    arr, tbits, ng = _fill_args(factors, passes[0])
    info = (ctypes.c_int * 16)()
    rc = emul.qipb_emul_fused_fill(st.ctypes.data_as(ctypes.c_void_p), n, qlib.C128 if dtype == np.complex128 else qlib.C64,
                                   len(passes[0].tile_bits), tbits, ng, arr, info)
    assert rc == 0, emul.qipb_emul_last_error()
    assert info[7] >= 1
    err = float(np.max(np.abs(st - want))) / float(np.max(np.abs(want)))
    assert err <= (1e-12 if dtype == np.complex128 else 2e-5), err


def orc_state(n, groups, feeds):
    from oracle import oracle as orc
    hot = np.zeros(2 ** 3)
    feeds = [f if not isinstance(f, int) else np.eye(2 ** 3)[f] for f in feeds]
    return orc.OracleBackend.make_state(n, groups, feeds).get_state()


def test_emulated_fill_mode_refuses_what_it_cannot_serve(emul):
    from qip_b200.backend import product_state_factors
    # multi-qubit vector groups and device-resident feeds are not product-of-one-qubit states
    assert product_state_factors([[0, 1]], [np.ones(4) / 2], 0b1100, 0, 4) is None
    n = 14
    factors = [(1 + 0j, 0j)] * n
    # a small tile (2^10): the library answers "unsupported" (3) and launches nothing
    g = ops.BitGate("matrix", (3,), 0, H2.astype(np.complex128), False)
    p = ops.Pass(True, [g, g], tuple(range(10)))
    arr, tbits, ng = _fill_args(factors, p)
    st = np.zeros(2 ** n, dtype=np.complex128)
    info = (ctypes.c_int * 16)()
    rc = emul.qipb_emul_fused_fill(st.ctypes.data_as(ctypes.c_void_p), n, qlib.C128, 10, tbits, ng, arr, info)
    assert rc == qlib.ERR_UNSUPPORTED and info[0] == 0 and not st.any()


def test_chunked_fused_pass_equals_the_whole_pass_and_leaves_other_chunks_alone():
    # qipb_apply_fused_chunk: the chunks partition the state, every chunk launch changes only its own amplitudes (the host
    # double asserts it) and all of them together equal qipb_apply_fused -- controls, diagonal targets and stage cells on
    # the chunk bits are read from the tile's base index like any other outside bit
    import hostlib
    L = hostlib.HostLib()
    n = 14
    rng = np.random.default_rng(1)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    for tile, fix in (((0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 12, 13), (7,)), (tuple(range(12)), (12, 13)),
                      ((0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 13), (8, 12))):
        hi = [b for b in tile if b >= 7]
        gates = [BitGate("matrix", (hi[0], hi[3]), 0, haar_unitary(rng, 4)), BitGate("matrix", (hi[1],), 1 << fix[0], H2.astype(complex)),
                 BitGate("matrix", (fix[0],), 0, np.diag([1, 1j]), True), BitGate("matrix", (tile[0],), 0, haar_unitary(rng, 2)),
                 BitGate("matrix", (hi[2],), 0, H2.astype(complex))]
        gates += [BitGate("matrix", (), (1 << hi[2]) | (1 << b), np.diag([np.exp(0.1j * (b + 1))]), True) for b in range(n) if b != hi[2]]
        p = Pass(True, gates, tile)
        arr, tb = pack_pass(p)
        a, b = psi.copy(), psi.copy()
        pa, pb = ctypes.c_void_p(a.ctypes.data), ctypes.c_void_p(b.ctypes.data)
        assert L.qipb_apply_fused(None, pa, n, qlib.C128, len(tile), tb, len(gates), arr) == 0, L.err
        for j in range(1 << len(fix)):
            fv = sum(((j >> t) & 1) << fb for t, fb in enumerate(fix))
            assert L.qipb_apply_fused_chunk(None, pb, n, qlib.C128, len(tile), tb, len(gates), arr, len(fix), qlib.int_array(fix), fv) == 0, L.err
        assert np.max(np.abs(a - b)) <= 1e-14 * np.max(np.abs(a))     # (a one-tile chunk takes the generic sweeps: other rounding)
        want = bitsim.run_passes(psi.copy(), [p], n)
        assert np.max(np.abs(a - want)) <= 1e-12 * np.max(np.abs(want))


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("seed", range(3))
def test_emulated_trios_match_separate_sweeps(emul, monkeypatch, dtype, seed):
    # a dense 2-qubit block and a lone dense 1-qubit gate share one sweep over 8-amplitude register groups (sweep_trio):
    # every block form x (general / real) 1-qubit gate, partner found across commuting ops, either one first, outside
    # controls that switch one of the two off on some tiles
    n = 15
    rng = np.random.default_rng(200 + seed)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi = (psi / np.linalg.norm(psi)).astype(dtype)
    hh = np.kron(H2, H2)
    cx = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)
    sw = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)
    forms = [lambda: haar_unitary(rng, 4), lambda: hh @ cx, lambda: (hh @ sw) @ np.diag(np.exp(1j * rng.uniform(0, 6, 4))),
             lambda: cx @ np.diag([1, 1j, 1, np.exp(0.3j)]), lambda: sw]
    ry = np.array([[np.cos(0.3), -np.sin(0.3)], [np.sin(0.3), np.cos(0.3)]], dtype=np.complex128)
    ones = [lambda: haar_unitary(rng, 2), lambda: H2.astype(np.complex128), lambda: ry]
    tile = list(range(7)) + [8, 10, 11, 13, 14]
    hi = [b for b in tile if b >= 3]
    gates = []
    for rep in range(12):
        bits = [int(b) for b in rng.choice(hi, size=3, replace=False)]
        blk = BitGate("matrix", (bits[0], bits[1]), (1 << 9) if rep % 5 == 4 else 0, np.ascontiguousarray(forms[rep % 5]()))
        one = BitGate("matrix", (bits[2],), (1 << 12) if rep % 6 == 5 else 0, np.ascontiguousarray(ones[rep % 3]()))
        first, second = (blk, one) if rep % 2 == 0 else (one, blk)
        gates.append(first)
        if rep % 3 == 1:                                       # something in between that commutes with the partner
            gates.append(BitGate("matrix", (int(rng.integers(0, n)),), 0, np.diag(np.exp(1j * rng.uniform(0, 6, 2))), True))
        if rep % 4 == 2:                                       # ... and something that does not (a control on its bit)
            gates.append(BitGate("matrix", (int(rng.choice([b for b in tile if b not in bits])),), 1 << bits[2], H2.astype(np.complex128)))
        gates.append(second)
    p = Pass(True, gates, tuple(tile))
    monkeypatch.setenv("QIPB_FUSED_PAIR", "0")
    monkeypatch.setenv("QIPB_FUSED_TRIO", "1")
    trio, info = run_emulated(emul, psi.copy(), [p], n, dtype)
    assert info[10] >= 4 and info[9] == info[0], info             # (a gate that carries a riding stage is not available for a trio)
    monkeypatch.setenv("QIPB_FUSED_TRIO", "0")
    single, info0 = run_emulated(emul, psi.copy(), [p], n, dtype)
    assert info0[10] == 0, info0
    tol = 1e-13 if dtype == np.complex128 else 2e-5
    assert np.max(np.abs(trio - single)) <= tol * np.max(np.abs(single))
    want = bitsim.run_passes(psi.astype(np.complex128), [p], n)
    assert np.max(np.abs(trio - want)) <= (1e-12 if dtype == np.complex128 else 1e-5) * np.max(np.abs(want))


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_emulated_stage_rides_on_a_dense_two_qubit_sweep(emul, monkeypatch, dtype):
    # sweep_dense2_stage: the table stage behind an un-controlled dense 2-qubit block is applied while the group is in
    # registers (every block form; stages with and without a common in-tile control; targets below and above the table split)
    n = 15
    rng = np.random.default_rng(7)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi = (psi / np.linalg.norm(psi)).astype(dtype)
    hh = np.kron(H2, H2)
    cx = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)
    tile = list(range(7)) + [8, 10, 11, 13, 14]
    gates = []
    for rep, (bits, mat) in enumerate([((10, 13), haar_unitary(rng, 4)), ((4, 11), hh @ cx), ((8, 5), (hh @ cx) @ np.diag(np.exp(1j * rng.uniform(0, 6, 4)))),
                                       ((14, 3), haar_unitary(rng, 4))]):
        gates.append(BitGate("matrix", bits, 0, np.ascontiguousarray(mat)))
        common = (1 << 6) if rep % 2 else 0                       # a stage needs >= 3 diagonal gates in a row
        for b in (0, 2, 7, 9, 12, bits[0]):
            if (1 << b) != common:
                gates.append(BitGate("matrix", (), (1 << b) | common, np.diag([np.exp(0.2j * (b + 1 + rep))]), True))
    p = Pass(True, gates, tuple(tile))
    monkeypatch.setenv("QIPB_FUSED_TRIO", "0")
    monkeypatch.setenv("QIPB_FUSED_PAIR", "0")
    monkeypatch.setenv("QIPB_FUSED_RIDE2", "1")
    ride, info = run_emulated(emul, psi.copy(), [p], n, dtype)
    assert info[11] >= 3, info                                    # (complex64: bit 3 is a bank-conflict bit, that block does not ride)
    monkeypatch.setenv("QIPB_FUSED_RIDE2", "0")
    plain, info0 = run_emulated(emul, psi.copy(), [p], n, dtype)
    assert info0[11] == 0, info0
    tol = 1e-13 if dtype == np.complex128 else 2e-5
    assert np.max(np.abs(ride - plain)) <= tol * np.max(np.abs(plain))
    want = bitsim.run_passes(psi.astype(np.complex128), [p], n)
    assert np.max(np.abs(ride - want)) <= (1e-12 if dtype == np.complex128 else 1e-5) * np.max(np.abs(want))


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n", [13, 15, 17])
def test_emulated_qfft_takes_four_steps_per_sweep(emul, monkeypatch, dtype, n):
    # sweep_qft4: a radix-16 butterfly with the member phases factorised out of the stage tables by the host; against the
    # radix-4 / single-step forms and the bit simulator, incl. a QFT on a sub-register and a stage that does NOT factorise
    monkeypatch.setenv("QIPB_FUSED_QFT4", "1")
    info = check(emul, qfft_stream(n), n, n, dtype)
    assert info[12] >= 1, info
    monkeypatch.setenv("QIPB_FUSED_QFT4", "0")
    info0 = check(emul, qfft_stream(n), n, n, dtype)
    assert info0[12] == 0 and info0[5] > info[5], (info0, info)
    monkeypatch.setenv("QIPB_FUSED_QFT4", "1")
    stream = list(qfft_stream(9, first_qubit=2)) + list(layered_stream(n, 1, 3)) + list(qfft_stream(n))
    check(emul, stream, n, 4, dtype)
    # a stage whose gates couple TWO other bits (a doubly-controlled phase) does not factorise: that group must fall back
    q = list(range(n))
    odd = []
    for k in range(4):
        odd.append({q[k]: H2})
        for i in range(k + 1, n):
            odd.append({(q[i], q[k]): CMat(rm_mat(1 + i - k))})
        if k == 1:
            odd.append({(q[5], q[6], q[k]): CMat(CMat(rm_mat(3)))})
            odd.append({(q[7], q[6], q[k]): CMat(CMat(rm_mat(2)))})
    check(emul, odd, n, 5, dtype)


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n", [13, 16])
def test_emulated_qft_steps_on_the_lowest_bits(emul, monkeypatch, dtype, n):
    # sweep_qft_low: the Hadamard + phase steps on tile bits 0..2 (one sweep for up to three of them; on the device the
    # butterflies cross lanes, the emulator runs the same arithmetic pair by pair) against the per-step sweeps
    monkeypatch.setenv("QIPB_FUSED_QFTLOW", "1")
    info = check(emul, qfft_stream(n), n, n + 1, dtype)
    assert info[13] >= 2, info                                    # (the last Hadamard of a QFT has no phases behind it)
    monkeypatch.setenv("QIPB_FUSED_QFTLOW", "0")
    info0 = check(emul, qfft_stream(n), n, n + 1, dtype)
    assert info0[13] == 0, info0
    monkeypatch.setenv("QIPB_FUSED_QFTLOW", "1")
    # a QFT whose register ends on bit 1 / bit 0 only, and one that is followed by other gates on the low bits
    check(emul, list(qfft_stream(n - 1, first_qubit=0)) + list(layered_stream(n, 1, 2)), n, 3, dtype)
    check(emul, list(layered_stream(n, 1, 5)) + list(qfft_stream(n - 2, first_qubit=2)), n, 4, dtype)
