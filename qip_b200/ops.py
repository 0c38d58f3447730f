"""Host-side decoding of the boundary's `mats` vocabulary into primitive gates, and the fusion planner.

Pure python / numpy -- no device code -- so it is covered by the CPU test tier.

decode_mats     : `{int | tuple : ndarray | list | CMat | SwapMat}` -> list[Gate], with the same
                  validation (exception types, order, messages) as qip/util.py:29-58 and
                  qip/ext/kronprod.pyx:114-116.
simplify        : exact structural rewrites that cut HBM traffic: identity removal, promotion of
                  "identity unless this qubit is 1" targets to control bits (what CMat means,
                  qip/ext/kronprod.pyx:215-225, discovered for plain ndarrays too -- e.g. the
                  R/Rm phase gate diag(1, e^{i phi}) of qip/operators.py:108-127 becomes a scalar
                  phase on a controlled sub-space), diagonal detection.
plan_passes     : greedy grouping of consecutive gates into fused tile passes (qip_b200/csrc/fused.cu)
                  under a bytes-moved cost model.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class Gate:
    """One primitive: `mat` (2^k x 2^k) on `targets`, enabled where all `controls` are 1.

    Qubit numbers are the reference's global qubit indices (0 = most significant bit of the
    state index, qip/ext/kronprod.pyx:168,187).  targets[0] is the most significant bit of the
    matrix index (kronprod.pyx:184-189).  kind == "swap": exchange of two qubits (SwapMat(1))."""
    kind: str                       # "matrix" | "swap"
    targets: Tuple[int, ...]
    controls: Tuple[int, ...] = ()
    mat: Optional[np.ndarray] = None
    diagonal: bool = False

    @property
    def k(self):
        return len(self.targets)

    def qubits(self):
        return tuple(self.targets) + tuple(self.controls)


# --------------------------------------------------------------------------------- decoding
def _is_ctrl(m):
    return getattr(m, "_kron_struct", None) == 2


def _is_swap(m):
    return getattr(m, "_kron_struct", None) == 3


def decode_mats(mats, n: int) -> List[Gate]:
    """Validate and decode one kronselect_dot call.  Entries are returned in dict order; they act
    on disjoint targets (shared qubits may only be controls) so applying them one after another
    equals the reference's single product-matrix op."""
    norm = []
    for key in mats:                                           # qip/util.py:35-58
        if type(key) != tuple and type(key) != int:
            raise Exception("Type of indices must be tuple: {}".format(key))
        m = mats[key]
        if type(m) == list:
            m = np.array(m)
        tkey = key if type(key) == tuple else (key,)
        if not hasattr(m, "shape"):
            raise ValueError("Cannot pass matrices which are not numpy, SwapMat, or CMat")
        if 2 ** len(tkey) != m.shape[0] or 2 ** len(tkey) != m.shape[1]:
            raise Exception("Shape of square submatrix must equal 2**(number of indices): "
                            "{}: {}".format(key, m))
        norm.append((tuple(int(i) for i in tkey), m))

    gates: List[Gate] = []
    target_owner = {}
    control_users = set()
    for key, m in norm:
        for q in key:
            if not (0 <= q < n):
                raise ValueError("qubit index {} out of range for {} qubits".format(q, n))
        if len(set(key)) != len(key):
            raise ValueError("repeated qubit index in {}".format(key))
        controls = []
        inner, rest = m, list(key)
        while _is_ctrl(inner):                                 # kronprod.pyx:215-225
            controls.append(rest.pop(0))
            inner = inner.m
            if type(inner) == list:
                inner = np.array(inner)
        entry_gates = []
        if _is_swap(inner):                                    # kronprod.pyx:227-231
            w = int(inner.n)
            if len(rest) != 2 * w:
                raise Exception("Shape of square submatrix must equal 2**(number of indices): "
                                "{}: {}".format(key, m))
            for a, b in zip(rest[:w], rest[w:]):
                entry_gates.append(Gate("swap", (a, b), tuple(controls)))
        elif isinstance(inner, np.ndarray):
            if inner.ndim != 2 or inner.shape[0] != 2 ** len(rest) or inner.shape[1] != 2 ** len(rest):
                raise Exception("Shape of square submatrix must equal 2**(number of indices): "
                                "{}: {}".format(key, m))
            mat = np.ascontiguousarray(inner, dtype=np.complex128)   # kronprod.pyx:99
            entry_gates.append(Gate("matrix", tuple(rest), tuple(controls), mat))
        else:                                                  # kronprod.pyx:114-116
            raise ValueError("Cannot pass matrices which are not numpy, SwapMat, or CMat")
        for q in rest:
            if q in target_owner or q in control_users:
                raise ValueError("qubit {} is acted on by more than one entry of one op; entries may "
                                 "share control qubits only".format(q))
            target_owner[q] = key
        for q in controls:
            if q in target_owner:
                raise ValueError("qubit {} is a control of one entry and a target of another".format(q))
        control_users.update(controls)
        gates.extend(entry_gates)
    return gates


# --------------------------------------------------------------------------------- simplification
_SWAP4 = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def _bit_split(mat, k, j):
    """Blocks of `mat` w.r.t. matrix-index bit for target j (0 = MSB): (M00, M01, M10, M11)."""
    d = 1 << k
    lo = 1 << (k - 1 - j)
    hi = d // (2 * lo)
    t = mat.reshape(hi, 2, lo, hi, 2, lo)          # (row hi, row bit, row lo, col hi, col bit, col lo)
    h = d // 2
    return (t[:, 0, :, :, 0, :].reshape(h, h), t[:, 0, :, :, 1, :].reshape(h, h),
            t[:, 1, :, :, 0, :].reshape(h, h), t[:, 1, :, :, 1, :].reshape(h, h))


_EYES = {d: np.eye(d) for d in (1, 2, 4, 8, 16)}


def _is_identity(m) -> bool:
    d = m.shape[0]
    eye = _EYES.get(d)
    return np.array_equal(m, eye if eye is not None else np.eye(d))


def simplify(g: Gate) -> Optional[Gate]:
    """Exact rewrites (only comparisons with exactly 0.0 / 1.0).  Returns None for the identity."""
    if g.kind != "matrix":
        return g
    mat, targets, controls = g.mat, list(g.targets), list(g.controls)
    changed = True
    while changed and targets:
        changed = False
        k = len(targets)
        for j in range(k):
            m00, m01, m10, m11 = _bit_split(mat, k, j)
            if not m01.any() and not m10.any() and _is_identity(m00):
                controls.append(targets.pop(j))
                mat = np.ascontiguousarray(m11)
                changed = True
                break
    d = mat.shape[0]
    if _is_identity(mat):
        return None
    if d == 4 and np.array_equal(mat, _SWAP4):
        return Gate("swap", tuple(targets), tuple(controls))
    diagonal = not (mat - np.diag(np.diag(mat))).any()
    return Gate("matrix", tuple(targets), tuple(controls), np.ascontiguousarray(mat, dtype=np.complex128), diagonal)


# --------------------------------------------------------------------------------- block merging
def gate_unitary(g: Gate):
    """(qubit list, dense matrix) of a gate over controls ++ targets (first qubit = MSB)."""
    if g.kind == "swap":
        inner = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)
    else:
        inner = g.mat
    qs = list(g.controls) + list(g.targets)
    d = 1 << len(qs)
    di = inner.shape[0]
    full = np.eye(d, dtype=np.complex128)
    full[d - di:, d - di:] = inner          # all controls = 1 is the last block of the index space
    return qs, full


def _apply_to_block(block_mat, block_qubits, u, u_qubits):
    """(u acting on u_qubits, embedded in block space) @ block_mat."""
    m = len(block_qubits)
    t = block_mat.reshape([2] * m + [1 << m])
    axes = [block_qubits.index(q) for q in u_qubits]
    ku = len(u_qubits)
    ut = u.reshape([2] * (2 * ku))
    r = np.tensordot(ut, t, axes=(list(range(ku, 2 * ku)), axes))
    r = np.moveaxis(r, list(range(ku)), axes)
    return np.ascontiguousarray(r.reshape(1 << m, 1 << m))


def tile_cost(g: Optional[Gate]) -> float:
    """Relative cost of a gate inside a fused tile pass, calibrated on B200 (profiles/): a dense gate
    costs one shared-memory sweep whose length hardly depends on k (1q ~ 4, 2q ~ 5 units, scaled by
    the fraction of the state its controls select); a diagonal gate is nearly free because runs of
    diagonal gates are folded into one table sweep by the kernel (fused.cu, "stage").  Used to refuse
    merges that would turn cheap diagonal gates into a dense block (H followed by a C-phase in a QFT)."""
    if g is None:
        return 0.0
    frac = 2.0 ** (-len(g.controls))
    if g.kind == "swap":
        return 1.0 * frac
    if g.diagonal or g.k == 0:
        return 0.3
    return (3.0 + g.k) * frac


def merge_blocks(gates: Sequence[Gate], max_k: int = 4, pack: bool = True, cost_aware: bool = False) -> List[Gate]:
    """Order-preserving merge of small gates into dense blocks on at most max_k qubits by multiplying
    their matrices on the host (changes rounding at the 1e-16 level only).  Three passes:
      1. backward: a gate joins the latest earlier block it shares a qubit with (it may slide back
         over blocks it shares no qubit with) when the union still fits;
      2. forward: a block joins the NEXT block it shares a qubit with when the union fits (nothing in
         between touches its qubits, so it may slide forward);
      3. pack: blocks on disjoint qubits are tensored together while nothing between them touches
         the later one's qubits.
    Every block is re-simplified, so a block made only of controlled / diagonal gates keeps its
    cheap form.  Gates wider than max_k stay as they are."""
    # a block is [qubit list, matrix]; an opaque (too wide) gate is [qubit list, None, gate]
    blocks: List[list] = []
    last = {}

    def cost_of(bq, bm):
        return tile_cost(simplify(Gate("matrix", tuple(bq), (), bm)))

    def worth(block, qs, u, first=None):
        """Would (u on qs) . block [or block . first] be cheaper than running the two apart?"""
        if not cost_aware:
            return True
        trial = [list(block[0]), block[1]]
        _absorb(trial, qs, u)
        return cost_of(trial[0], trial[1]) <= cost_of(block[0], block[1]) + cost_of(list(qs), u) + 0.5

    for g in gates:
        if len(g.qubits()) > max_k:
            blocks.append([list(g.qubits()), None, g])
            for q in g.qubits():
                last[q] = len(blocks) - 1
            continue
        qs, u = gate_unitary(g)
        b = max([last.get(q, -1) for q in qs] + [-1])
        if b >= 0 and blocks[b][1] is not None and len(set(blocks[b][0]) | set(qs)) <= max_k \
                and worth(blocks[b], qs, u):
            _absorb(blocks[b], qs, u)
            target = b
        else:
            blocks.append([list(qs), u])
            target = len(blocks) - 1
        for q in qs:
            last[q] = target

    # pass 2: forward merges
    alive = [True] * len(blocks)
    for i in range(len(blocks)):
        if blocks[i][1] is None:
            continue
        qi = set(blocks[i][0])
        for j in range(i + 1, len(blocks)):
            if not alive[j]:
                continue
            qj = set(blocks[j][0])
            if qi & qj:
                if blocks[j][1] is not None and len(qi | qj) <= max_k and worth(blocks[i], blocks[j][0], blocks[j][1]):
                    # i's content runs first, then j's: rebuild j as (j's matrix) . (i's matrix)
                    merged = [list(blocks[i][0]), blocks[i][1]]
                    _absorb(merged, blocks[j][0], blocks[j][1])
                    blocks[j] = merged
                    alive[i] = False
                break
    blocks = [b for b, a in zip(blocks, alive) if a]

    # pass 3: pack disjoint blocks
    if pack:
        alive = [True] * len(blocks)
        for i in range(len(blocks)):
            if not alive[i] or blocks[i][1] is None:
                continue
            touched = set()
            for j in range(i + 1, len(blocks)):
                if len(blocks[i][0]) >= max_k:
                    break
                if not alive[j]:
                    continue
                qj = set(blocks[j][0])
                if blocks[j][1] is not None and not (qj & touched) and not (qj & set(blocks[i][0])) \
                        and len(blocks[i][0]) + len(qj) <= max_k and not cost_aware:
                    _absorb(blocks[i], blocks[j][0], blocks[j][1])
                    alive[j] = False
                else:
                    touched |= qj
        blocks = [b for b, a in zip(blocks, alive) if a]

    out: List[Gate] = []
    for b in blocks:
        if b[1] is None:
            out.append(b[2])
            continue
        s = simplify(Gate("matrix", tuple(b[0]), (), b[1]))
        if s is not None:
            out.append(s)
    return out


def _absorb(block, qs, u):
    """block <- (u on qs) . block, growing the block's qubit list if needed."""
    bq, bm = block[0], block[1]
    new_q = bq + [q for q in qs if q not in bq]
    if len(new_q) != len(bq):
        bm = _apply_to_block(np.eye(1 << len(new_q), dtype=np.complex128), new_q, bm, bq)
    block[0] = new_q
    block[1] = _apply_to_block(bm, new_q, u, qs)


def sink_lone_diagonals(gates: Sequence[Gate], min_run: int = 3) -> List[Gate]:
    """Cluster isolated diagonal gates.  The fused kernel folds a run of >= 3 consecutive diagonal gates
    into ONE table sweep (fused.cu, "stage"), but a lone diagonal gate costs a sweep of its own.  A diagonal
    gate commutes with every other diagonal gate and with any gate whose non-diagonal targets avoid its qubits
    (controls are diagonal too), so diagonal gates that sit in runs shorter than `min_run` are moved as late
    as commutation allows; they meet at the end of the segment (or in front of the first gate that really
    needs one of their qubits) and fold there.  Runs that already fold -- the controlled phases behind each H of a
    QFT, which ride on that H's sweep -- stay where they are.  Exact: only the order of commuting gates changes."""
    n = len(gates)
    is_diag = [g.kind == "matrix" and (g.diagonal or g.k == 0) for g in gates]
    floater = [False] * n
    i = 0
    while i < n:
        if not is_diag[i]:
            i += 1
            continue
        j = i
        while j < n and is_diag[j]:
            j += 1
        if j - i < min_run:
            for t in range(i, j):
                floater[t] = True
        i = j
    if not any(floater):
        return list(gates)
    out: List[Gate] = []
    pending: List[Gate] = []
    for idx, g in enumerate(gates):
        if floater[idx]:
            pending.append(g)
            continue
        if not is_diag[idx] and pending:
            hit = set(g.targets)                     # a swap's targets are non-diagonal as well
            stay = []
            for d in pending:
                if hit & (set(d.targets) | set(d.controls)):
                    out.append(d)
                else:
                    stay.append(d)
            pending = stay
        out.append(g)
    out.extend(pending)
    return out


# --------------------------------------------------------------------------------- bit-level form
@dataclass
class BitGate:
    """A Gate lowered to local index bits.  ctrl_mask: bits that must be 1."""
    kind: str
    bits: Tuple[int, ...]           # bits[0] = most significant matrix-index bit
    ctrl_mask: int = 0
    mat: Optional[np.ndarray] = None
    diagonal: bool = False

    @property
    def k(self):
        return len(self.bits)

    def nctrl(self):
        return bin(self.ctrl_mask).count("1")


def lower(g: Gate, n: int) -> BitGate:
    bits = tuple(n - 1 - q for q in g.targets)
    cm = 0
    for q in g.controls:
        cm |= 1 << (n - 1 - q)
    return BitGate(g.kind, bits, cm, g.mat, g.diagonal)


def merge_bitgates(gates: Sequence["BitGate"], max_k: int = 2) -> List["BitGate"]:
    """merge_blocks on gates that are already lowered to index bits (bit positions act as labels)."""
    logical = []
    for g in gates:
        ctrl = tuple(b for b in range(64) if (g.ctrl_mask >> b) & 1)
        logical.append(Gate(g.kind, tuple(g.bits), ctrl, g.mat, g.diagonal))
    out = []
    merged = merge_blocks(logical, max_k, cost_aware=True)
    import os
    if os.environ.get("QIPB_SINK_DIAGONALS", "1") != "0":
        merged = sink_lone_diagonals(merged)
    for g in merged:
        cm = 0
        for b in g.controls:
            cm |= 1 << b
        out.append(BitGate(g.kind, tuple(g.targets), cm, g.mat, g.diagonal))
    return out


# --------------------------------------------------------------------------------- fusion planner
@dataclass
class Pass:
    """Either one stand-alone kernel (fused == False, exactly one gate) or a fused tile pass."""
    fused: bool
    gates: List[BitGate]
    tile_bits: Tuple[int, ...] = ()


def _fusable(g: BitGate) -> bool:
    if g.kind == "swap":
        return True                 # lowered to a dense 4x4 permutation inside a tile
    return g.k <= 2


def _needs(g: BitGate):
    """Bits that must be tile bits for g to run inside a fused pass."""
    if g.kind == "swap":
        return set(g.bits)
    if g.diagonal or g.k == 0:
        return set()
    return set(g.bits)


def gate_bytes(g: BitGate, nbits: int, amp_bytes: int) -> float:
    """HBM bytes of a stand-alone launch: read + write of every touched amplitude."""
    touched = 2.0 ** (nbits - g.nctrl())
    if g.kind == "swap":
        touched /= 2.0
    return 2.0 * amp_bytes * touched


def choose_tile(required, nbits: int, tile_bits: int):
    tb = min(tile_bits, nbits)
    tile = set(required)
    b = 0
    while len(tile) < tb and b < nbits:
        tile.add(b)
        b += 1
    return tuple(sorted(tile))


def _PACK_1Q() -> bool:
    import os
    return os.environ.get("QIPB_PACK_1Q", "1") != "0"          # tuning knob for profiling runs


def _ncoef(g: BitGate) -> int:
    if g.kind == "swap":
        return 16
    return (1 << g.k) if (g.diagonal or g.k == 0) else (1 << g.k) ** 2


def _is_diag(g: BitGate) -> bool:
    return g.kind == "matrix" and (g.diagonal or g.k == 0)


def pack_lone_1q(gates: Sequence[BitGate], min_run: int = 3) -> List[BitGate]:
    """Inside ONE fused pass: tensor pairs of un-controlled dense 1-qubit gates on different bits into one dense
    2-qubit block.  Every gate of a pass is a sweep over the tile in shared memory; two lone Hadamards cost two
    sweeps of 16 FP64 instructions per pair, their Kronecker product is a REAL 4x4 block -- one sweep of 32 per group
    of four (fused.cu MK_REAL) and half the shared-memory traffic and barriers.  Exact up to the rounding of the
    products of matrix entries.  Gate B joins an earlier gate A when nothing between them touches B's bit (B may
    slide back).  A 1-qubit gate that is directly followed by a run of >= `min_run` diagonal gates is left alone:
    the kernel applies that run on the gate's own sweep (a QFT step)."""
    out: List[BitGate] = []
    open_idx: List[int] = []                       # positions in `out` of lone gates still waiting for a partner
    n = len(gates)

    def rides(i: int) -> bool:
        j = i + 1
        while j < n and _is_diag(gates[j]):
            j += 1
        return j - (i + 1) >= min_run

    for i, g in enumerate(gates):
        lone = g.kind == "matrix" and g.k == 1 and not g.diagonal and g.ctrl_mask == 0 and not rides(i)
        if lone:
            bit = 1 << g.bits[0]
            partner = None
            for idx in reversed(open_idx):
                a = out[idx]
                if a.bits[0] == g.bits[0]:
                    continue
                blocked = False
                for h in out[idx + 1:]:
                    hm = h.ctrl_mask
                    for b in h.bits:
                        hm |= 1 << b
                    if hm & bit:
                        blocked = True
                        break
                if not blocked:
                    partner = idx
                    break
            if partner is not None:
                a = out[partner]
                out[partner] = BitGate("matrix", (a.bits[0], g.bits[0]), 0, np.ascontiguousarray(np.kron(a.mat, g.mat)), False)
                open_idx.remove(partner)
                continue
            open_idx.append(len(out))
        out.append(g)
    return out


def plan_passes(gates: Sequence[BitGate], nbits: int, amp_bytes: int = 16, tile_bits: int = 12,
                min_low_bits: int = 7, max_gates: int = 280, enable: bool = True, max_coefs: int = 1 << 30) -> List[Pass]:
    """Greedy, order-preserving fusion.  A group grows while the union of the bits its non-diagonal
    gates need, together with the `min_low_bits` lowest bits, still fits in a tile; a group is run
    fused only when that moves fewer bytes than launching its gates one by one."""
    passes: List[Pass] = []
    tb = min(tile_bits, nbits)
    low = set(range(min(min_low_bits, tb)))
    cur: List[BitGate] = []
    need = set()
    ncoefs = 0

    def flush():
        nonlocal cur, need, ncoefs
        if not cur:
            return
        solo = sum(gate_bytes(g, nbits, amp_bytes) for g in cur)
        fused_cost = 2.0 * amp_bytes * 2.0 ** nbits
        if len(cur) >= 2 and enable and fused_cost < solo:
            passes.append(Pass(True, pack_lone_1q(cur) if _PACK_1Q() else cur, choose_tile(need, nbits, tb)))
        else:
            passes.extend(Pass(False, [g]) for g in cur)
        cur, need, ncoefs = [], set(), 0

    for g in gates:
        if not enable or not _fusable(g):
            flush()
            passes.append(Pass(False, [g]))
            continue
        nn = need | _needs(g)
        if len(nn | low) > tb or len(cur) >= max_gates or ncoefs + _ncoef(g) > max_coefs:
            flush()
            nn = _needs(g)
            if len(nn | low) > tb:        # cannot happen for k <= 2 and tb >= min_low_bits + 2
                passes.append(Pass(False, [g]))
                continue
        cur.append(g)
        ncoefs += _ncoef(g)
        need = nn
    flush()
    return passes


# --------------------------------------------------------------------------------- strategy choice
# Cost of one dense 2-qubit gate inside a fused tile pass, as a fraction of one full HBM sweep (read +
# write of the whole state), calibrated on B200 (profiles/r01_probe_fused_final.txt: 11.3 ms floor at 31
# qubits, 15.4 ms with 4 dense gates, 27 ms with 8): the sweeps are FP64-issue bound and only the first
# two hide behind the tile traffic.  plan() uses it to choose between
#   A. blocks of <= 2 qubits -> fused tile passes (diagonal gates and controls anywhere), and
#   B. blocks of <= 4 qubits -> one register-blocked stand-alone launch per block (HBM roofline).
TILE_GATE_COST = 0.23


def pass_cost(p: Pass, nbits: int, amp_bytes: int) -> float:
    """Estimated cost of a pass in HBM-byte equivalents, calibrated at 31/33 qubits complex128 on B200
    (profiles/): a fused pass costs one full sweep of HBM traffic overlapped with ~0.23 sweep-times of
    shared-memory / FP64 work per dense 2-qubit gate (0.15 per dense 1-qubit gate); the register-blocked
    K=3/K=4 kernels are FP64-limited."""
    full = 2.0 * amp_bytes * 2.0 ** nbits
    if not p.fused:
        g = p.gates[0]
        slow = 1.5 if (g.kind == "matrix" and g.k >= 4 and not g.diagonal) else 1.1 if (g.kind == "matrix" and g.k == 3 and not g.diagonal) else 1.0
        return slow * gate_bytes(g, nbits, amp_bytes)
    work = 0.0
    for g in p.gates:
        frac = 2.0 ** (-g.nctrl())
        if g.kind == "swap":
            work += 0.5 * frac
        elif g.diagonal or g.k == 0:
            work += 0.1 * frac
        else:
            work += frac * (1.0 if g.k >= 2 else 0.65)
    return full * max(1.0, 0.5 + TILE_GATE_COST * work)


def plan(gates: Sequence[Gate], n: int, amp_bytes: int = 16, fuse: bool = True, tile_bits: int = 12,
         min_low_bits: int = 7, strategy: str = "auto"):
    """Logical gates -> (list[Pass], name of the chosen strategy)."""
    if not fuse:
        return [Pass(False, [lower(g, n)]) for g in gates], "unfused"
    out = {}
    if strategy in ("auto", "tile"):
        import os
        ca = os.environ.get("QIPB_COST_AWARE", "1") != "0"         # tuning knob for profiling runs
        merged = merge_blocks(gates, 2, cost_aware=ca)
        if os.environ.get("QIPB_SINK_DIAGONALS", "1") != "0":      # tuning knob for profiling runs
            merged = sink_lone_diagonals(merged)
        a = plan_passes([lower(g, n) for g in merged], n, amp_bytes,
                        tile_bits=tile_bits, min_low_bits=min_low_bits)
        out["tile"] = (sum(pass_cost(p, n, amp_bytes) for p in a), a)
    if strategy in ("auto", "dense4"):
        b = [Pass(False, [lower(g, n)]) for g in merge_blocks(gates, 4)]
        out["dense4"] = (sum(pass_cost(p, n, amp_bytes) for p in b), b)
    name = min(out, key=lambda k: out[k][0])
    return out[name][1], name
