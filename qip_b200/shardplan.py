"""Scheduling of a gate stream onto a state sharded over P = 2^G GPUs by its top G index bits.

Pure host logic (numpy only): it is exercised on the CPU tier by a virtual-shard numpy executor and
a world_size-2 gloo run (tests/test_sharding.py); qip_b200/sharded.py executes the same actions
with the CUDA kernels and NVLink peer memory.

Replaces qip/distributed's manager/worker decomposition (qip/distributed/manager.py:138-236,
worker/worker.py:57-169): there every gate is computed as a P x P grid of mat-vec blocks followed by
a reduce-to-diagonal and a re-broadcast.  Here the shard of rank r is the contiguous index range
whose top G bits equal r (the same slicing as manager.py:162-170, but P shards, not P^2 blocks) and
gate locality is exploited:
  * control bits on global (rank) positions switch a gate on or off per rank -- no communication;
  * diagonal gates with targets on global positions pick a sub-diagonal per rank -- none either;
  * un-controlled Swap is a relabelling of the logical->physical map -- free;
  * a dense 1-qubit gate on a global position whose qubit is not needed again soon runs as ONE
    fused compute+exchange kernel over peer memory (PeerGate1);
  * anything else non-diagonal on a global position first swaps that position with a local one
    (Exchange: each rank trades half of its shard with one partner), victims chosen Belady-style
    among the high local positions, whose halves are contiguous.
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .ops import BitGate, Gate

SWAP4 = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


class Layout(object):
    """Logical qubit -> physical index-bit position.  Positions >= nl are rank bits."""

    def __init__(self, n: int, gbits: int):
        self.n, self.G, self.nl = n, gbits, n - gbits
        self.pos = [n - 1 - q for q in range(n)]          # canonical: qubit 0 = most significant bit

    def copy(self):
        c = Layout(self.n, self.G)
        c.pos = list(self.pos)
        return c

    def qubit_at(self, p: int) -> int:
        return self.pos.index(p)

    def swap_qubits(self, a: int, b: int):
        self.pos[a], self.pos[b] = self.pos[b], self.pos[a]

    def is_global(self, q: int) -> bool:
        return self.pos[q] >= self.nl

    def canonical(self) -> bool:
        return all(self.pos[q] == self.n - 1 - q for q in range(self.n))


@dataclass
class PhysGate:
    """A gate in PHYSICAL positions over the whole (sharded) index; rank-independent."""
    bits: Tuple[int, ...]
    ctrl_mask: int
    mat: np.ndarray
    diagonal: bool


@dataclass
class Apply:
    gate: PhysGate


@dataclass
class Exchange:
    gpos: int        # physical rank-bit position (>= nl)
    lpos: int        # physical local position (< nl)


@dataclass
class MultiExchange:
    """Several (rank bit, local bit) swaps at once: one all-to-all-class kernel moves (1 - 2^-g) of a
    shard per direction instead of g/2 shards for g pairwise exchanges."""
    pairs: List[Tuple[int, int]]     # (gpos, lpos), all positions distinct


@dataclass
class PeerGate1:
    gpos: int
    mat: np.ndarray
    ctrl_mask: int   # physical mask (may contain rank bits)


@dataclass
class LocalSwap:
    a: int
    b: int


def to_phys(g: Gate, lay: Layout) -> PhysGate:
    cm = 0
    for q in g.controls:
        cm |= 1 << lay.pos[q]
    if g.kind == "swap":
        return PhysGate(tuple(lay.pos[q] for q in g.targets), cm, SWAP4, False)
    return PhysGate(tuple(lay.pos[q] for q in g.targets), cm, g.mat, g.diagonal)


def lower_for_rank(pg: PhysGate, nl: int, rank: int) -> Optional[BitGate]:
    """Rank-local form of a PhysGate whose non-diagonal targets are all local; None = no-op here."""
    lowmask = (1 << nl) - 1
    cg = pg.ctrl_mask >> nl
    if (rank & cg) != cg:
        return None
    k = len(pg.bits)
    glob = [j for j, b in enumerate(pg.bits) if b >= nl]
    if not glob:
        return BitGate("matrix", pg.bits, pg.ctrl_mask & lowmask, pg.mat, pg.diagonal)
    if not pg.diagonal:
        raise ValueError("non-diagonal target on a rank bit must be localised first")
    d = np.diag(pg.mat)
    keep = [j for j in range(k) if j not in glob]
    sub = np.zeros(1 << len(keep), dtype=np.complex128)
    for c in range(1 << len(keep)):
        idx = 0
        for t, j in enumerate(keep):
            if (c >> (len(keep) - 1 - t)) & 1:
                idx |= 1 << (k - 1 - j)
        for j in glob:
            if (rank >> (pg.bits[j] - nl)) & 1:
                idx |= 1 << (k - 1 - j)
        sub[c] = d[idx]
    if len(keep) == 0 and sub[0] == 1.0:
        return None
    return BitGate("matrix", tuple(pg.bits[j] for j in keep), pg.ctrl_mask & lowmask, np.diag(sub), True)


def choose_initial_layout(gates: Sequence[Gate], lay: Layout) -> None:
    """Before the state exists, put on the rank bits the qubits whose first NON-DIAGONAL use comes
    latest in the queued gates (diagonal gates and controls never need a qubit to be local).  A QFT then
    runs its first n-G Hadamards without any exchange.  `lay` is permuted in place."""
    if lay.G == 0:
        return
    first = {}
    for i, g in enumerate(gates):
        if g.kind == "swap":
            used = list(g.targets) if g.controls else []       # un-controlled swaps are relabels
        else:
            used = [] if (g.diagonal or g.k == 0) else list(g.targets)
        for q in used:
            first.setdefault(q, i)
    never = len(gates) + 1
    order = sorted(range(lay.n), key=lambda q: (-first.get(q, never), q))    # latest first use first
    want_global = order[:lay.G]
    # keep the canonical layout unless it is strictly worse
    canon_global = [q for q in range(lay.n) if lay.is_global(q)]
    if sorted(first.get(q, never) for q in canon_global) == sorted(first.get(q, never) for q in want_global):
        return
    for q in want_global:
        if not lay.is_global(q):
            victim = next(v for v in canon_global if v not in want_global and lay.is_global(v))
            lay.swap_qubits(q, victim)


def defer_global(gates: Sequence[Gate], lay: Layout) -> List[Gate]:
    """Order-preserving reorder that postpones everything which needs a qubit sitting on a rank bit.

    Gates on disjoint qubits commute, so a gate with a non-diagonal target on a rank bit -- and every later
    gate that shares a qubit with a postponed one -- can move behind all the gates that are purely local.
    A layer that touches every qubit (the layered benchmark: 1-qubit gates on all n qubits, then a perfect
    matching of 2-qubit gates) then needs ONE multi-bit exchange instead of one per first use: all local
    work first, one remap that brings every rank-bit qubit in (the victims are finished for this flush),
    then the postponed tail.  Relabelling swaps are tracked on a copy of the layout; `lay` is not changed."""
    return [g for part in split_global(gates, lay) for g in part]


def split_global(gates: Sequence[Gate], lay: Layout) -> Tuple[List[Gate], List[Gate]]:
    """(purely local gates, postponed gates) of defer_global, each in its original relative order."""
    if lay.G == 0 or len(gates) < 2:
        return list(gates), []
    sim = lay.copy()
    now: List[Gate] = []
    later: List[Gate] = []
    held = set()                                 # qubits touched by a postponed gate
    for g in gates:
        qs = set(g.targets) | set(g.controls)
        relabel = g.kind == "swap" and not g.controls
        if relabel:
            nondiag = []
        elif g.kind == "swap":
            nondiag = list(g.targets)
        else:
            nondiag = [] if (g.diagonal or g.k == 0) else list(g.targets)
        if (qs & held) or any(sim.is_global(q) for q in nondiag):
            later.append(g)
            held |= qs
            continue
        if relabel:
            sim.swap_qubits(g.targets[0], g.targets[1])
        now.append(g)
    return now, later


def schedule(gates: Sequence[Gate], lay: Layout, top_window: int = 8, peer_gates: bool = False,
             count_passes=None, tile_bits: int = 12, min_low_bits: int = 7) -> List[object]:
    """Turn logical gates into rank-independent actions; `lay` is updated in place.
    peer_gates: run a dense 1-qubit gate on a rank bit as one fused compute+exchange kernel when its
    qubit is never needed locally afterwards.  Off by default: the fused kernel moves one shard per
    direction, an Exchange followed by a local gate only half a shard (measured on B200).
    count_passes(list[BitGate]) -> int: the executor's real pass planner, used to confirm a hoisted exchange
    (see choose_prefetch) before it replaces the base schedule; without it only the estimate decides."""
    now, later = split_global(gates, lay)
    base = now + later
    at = choose_prefetch(base, len(now), lay, top_window, peer_gates, count_passes, tile_bits, min_low_bits)
    return _schedule_ordered(base, lay, top_window, peer_gates, prefetch_at=at)


# Cost of moving g rank bits in one remap, in units of one fused pass over a shard (measured on B200 at 33 local
# qubits: a pass 60-80 ms, a 1-bit exchange ~100 ms, a 3-bit remap 155 ms; profiles/r01_bench_*_n8.json).
def _move_cost(g: int) -> float:
    return 2.6 * (1.0 - 2.0 ** (-g))


def _estimate(actions: Sequence[object], nl: int, tile_bits: int = 12, min_low_bits: int = 7) -> float:
    """Cheap, rank-independent estimate of a schedule: fused passes (the bit-set grouping of ops.plan_passes on the
    non-diagonal target positions, no matrices touched) plus the moves."""
    cost = 0.0
    tb = min(tile_bits, nl)
    low = set(range(min(min_low_bits, tb)))
    need: set = set()
    open_pass = False
    for a in actions:
        if isinstance(a, Apply):
            bits = set() if a.gate.diagonal or len(a.gate.bits) == 0 else {b for b in a.gate.bits if b < nl}
        elif isinstance(a, LocalSwap):
            bits = {a.a, a.b}
        else:
            cost += 1.0 if open_pass else 0.0
            open_pass, need = False, set()
            cost += _move_cost(len(a.pairs)) if isinstance(a, MultiExchange) else _move_cost(1)
            continue
        if len(bits) > 2:                                      # wide gates run stand-alone: one sweep each
            cost += (1.0 if open_pass else 0.0) + 1.0
            open_pass, need = False, set()
            continue
        nn = need | bits
        if open_pass and len(nn | low) > tb:
            cost += 1.0
            nn = set(bits)
        need, open_pass = nn, True
    return cost + (1.0 if open_pass else 0.0)


def _real_cost(actions: Sequence[object], nl: int, rank: int, count_passes) -> float:
    cost = 0.0
    for step in compile_program(actions, nl, rank, lambda batch: [None] * count_passes(batch)):
        if isinstance(step, tuple):
            cost += len(step[1])
        else:
            cost += _move_cost(len(step.pairs)) if isinstance(step, MultiExchange) else _move_cost(1)
    return cost


def choose_prefetch(base: Sequence[Gate], n_now: int, lay: Layout, top_window: int = 8, peer_gates: bool = False,
                    count_passes=None, tile_bits: int = 12, min_low_bits: int = 7) -> Optional[int]:
    """Where (index into `base` = local gates + postponed tail of defer_global) the ONE multi-bit exchange of a flush
    should run, or None for "right before the first gate that needs it" (the base schedule).

    Hoisting: the exchange may run earlier, after a prefix of the local gates, when the qubits it evicts are finished
    by then; the postponed gates then share passes with the rest of the local work instead of needing passes of their
    own (QFT over 2^G shards: 6 passes + one remap instead of 6 + remap + 1).  Candidate cut points are scheduled on a
    copy of the layout and ranked with a cheap pass-count estimate; the best one replaces the base schedule only if it
    is strictly cheaper -- and, when the executor's planner is available, only if the real pass counts confirm it."""
    if lay.G == 0 or n_now == 0 or n_now == len(base) or peer_gates or len(base) > 20000:
        return None                                            # (very long queues: keep the planning cost linear)
    import os
    if os.environ.get("QIPB_SHARD_HOIST", "1") == "0":        # tuning knob for profiling runs
        return None
    nl = lay.nl
    base_actions = _schedule_ordered(base, lay.copy(), top_window, peer_gates)
    best_cost, best = _estimate(base_actions, nl, tile_bits, min_low_bits), None
    ncand = 8
    for cut in sorted({(n_now * k) // ncand for k in range(ncand)}):
        try:
            acts = _schedule_ordered(base, lay.copy(), top_window, peer_gates, prefetch_at=cut)
        except ValueError:
            continue
        c = _estimate(acts, nl, tile_bits, min_low_bits)
        if c < best_cost - 1e-9:
            best_cost, best = c, cut
    base_cost = _estimate(base_actions, nl, tile_bits, min_low_bits)
    if best is None or count_passes is None or base_cost - best_cost >= 1.0 - 1e-9:
        return best                                            # a whole pass saved: clear enough without re-planning
    rank = (1 << lay.G) - 1                                   # every control on a rank bit satisfied: no gate drops out
    hoisted = _schedule_ordered(base, lay.copy(), top_window, peer_gates, prefetch_at=best)
    if _real_cost(hoisted, nl, rank, count_passes) < _real_cost(base_actions, nl, rank, count_passes) - 1e-9:
        return best
    return None


def _schedule_ordered(gates: Sequence[Gate], lay: Layout, top_window: int = 8, peer_gates: bool = False,
                      prefetch_at: Optional[int] = None) -> List[object]:
    """The scheduling loop proper on a fixed gate order; `lay` is updated in place.  prefetch_at: index of the gate in
    front of which every rank-bit qubit with a later non-diagonal use is brought in (victims: qubits of the top local
    window that are not needed again, Belady-style like any other exchange)."""
    nl = lay.nl
    actions: List[object] = []

    def nondiag_targets(g: Gate):
        if g.kind == "swap":
            return list(g.targets) if g.controls else []
        return [] if (g.diagonal or g.k == 0) else list(g.targets)

    uses = [nondiag_targets(g) for g in gates]

    def next_use(q: int, start: int) -> int:
        for j in range(start, len(gates)):
            if q in uses[j]:
                return j
        return 1 << 30

    for i, g in enumerate(gates):
        if prefetch_at is not None and i == prefetch_at and lay.G > 0:
            wanted = sorted((q for q in range(lay.n) if lay.is_global(q) and next_use(q, i) < (1 << 30)),
                            key=lambda q: next_use(q, i))
            window = [lay.qubit_at(p) for p in range(nl - 1, max(-1, nl - 1 - top_window), -1)]
            taken = set()
            for q in wanted:
                cands = [v for v in window if v not in taken and next_use(v, i) >= (1 << 30)]    # finished qubits only
                if not cands:
                    break
                victim = cands[0]
                taken.add(victim)
                actions.append(Exchange(lay.pos[q], lay.pos[victim]))
                lay.swap_qubits(q, victim)
        if g.kind == "swap" and not g.controls:
            lay.swap_qubits(g.targets[0], g.targets[1])          # pure relabel
            continue
        need = uses[i]
        glob = [q for q in need if lay.is_global(q)]
        if glob and lay.G > 0:
            if (peer_gates and g.kind == "matrix" and g.k == 1 and next_use(glob[0], i + 1) >= (1 << 30)):
                cm = 0
                for q in g.controls:
                    cm |= 1 << lay.pos[q]
                actions.append(PeerGate1(lay.pos[glob[0]], g.mat, cm))
                continue
            # batch: other global qubits that are needed non-diagonally before any victim would be
            batch = list(glob)
            for j in range(i + 1, min(len(gates), i + 1 + 4 * lay.n)):
                for q in uses[j]:
                    if lay.is_global(q) and q not in batch and len(batch) < lay.G:
                        batch.append(q)
            window = [lay.qubit_at(p) for p in range(nl - 1, max(-1, nl - 1 - top_window), -1)]
            taken = set()
            for q in batch:
                cands = [v for v in window if v not in need and v not in taken and v not in batch]
                if not cands:
                    if q in glob:
                        raise ValueError("cannot localise qubit %d: no free local position" % q)
                    continue
                victim = max(cands, key=lambda v: next_use(v, i))
                if q not in glob and next_use(victim, i) <= next_use(q, i):
                    continue                                   # bringing q in early would evict something needed sooner
                taken.add(victim)
                actions.append(Exchange(lay.pos[q], lay.pos[victim]))
                lay.swap_qubits(q, victim)
        actions.append(Apply(to_phys(g, lay)))
    # qipb_peer_remap needs more local bits than exchanged pairs (nbits > g): tiny shards coalesce less or not at all
    return coalesce_exchanges(actions, max_pairs=max(1, min(3, (lay.n - lay.G) - 1)))


def coalesce_exchanges(actions: List[object], max_pairs: int = 3) -> List[object]:
    """Merge runs of consecutive Exchange actions on pairwise distinct positions into MultiExchange."""
    out: List[object] = []
    run: List[Exchange] = []

    def flush():
        if len(run) == 1:
            out.append(run[0])
        elif run:
            out.append(MultiExchange([(e.gpos, e.lpos) for e in run]))
        run.clear()

    for a in actions:
        if isinstance(a, Exchange) and len(run) < max_pairs and all(a.gpos != e.gpos and a.lpos != e.lpos for e in run):
            run.append(a)
            continue
        flush()
        if isinstance(a, Exchange):
            run.append(a)
        else:
            out.append(a)
    flush()
    return out


def annotate_chunks(program: Sequence[object], nl: int, want_bits: int, floor: int, window: int) -> None:
    """Exchange / compute overlap (ShardedB200Backend._run_overlapped): give every Exchange / MultiExchange of a
    program the local bit positions along which the state may be cut into chunks while the exchange is pipelined
    against the fused passes next to it -- attribute `chunk_bits` (ascending, possibly empty) on the action object.

    Every rank must cut along the SAME bits (chunk j of one shard trades with chunk j of its peers), but passes are
    planned per rank (a control on a rank bit drops a gate, a diagonal target there picks a sub-diagonal, and the
    planner merges and regroups accordingly), so the choice cannot depend on the caller's own passes: `program` is the
    compiled program of ONE agreed reference rank (the last one: all its rank bits are 1, no gate drops out), which
    every rank computes for itself.  For an exchange with passes P before and N after it, the window (a, b) = (last a
    passes of P, first b of N, each at most `window`) with the most passes that still leaves `want_bits` quiet local
    bits -- not exchanged, at or above `floor`, and not a tile bit of any pass of the window -- decides; the chunk bits
    are the highest quiet positions.  A rank then lets as many of ITS passes join the pipeline as avoid those bits
    (almost always the same ones)."""
    xtypes = (Exchange, MultiExchange)
    steps = list(program)
    for i, x in enumerate(steps):
        if not isinstance(x, xtypes):
            continue
        prev = steps[i - 1][1] if i > 0 and isinstance(steps[i - 1], tuple) else []
        nxt = steps[i + 1][1] if i + 1 < len(steps) and isinstance(steps[i + 1], tuple) else []
        victims = {l for _, l in (x.pairs if isinstance(x, MultiExchange) else [(x.gpos, x.lpos)])}
        cands = [p for p in range(nl - 1, floor - 1, -1) if p not in victims]
        best = None
        for a in range(min(window, len(prev)), -1, -1):
            for b in range(min(window, len(nxt)), -1, -1):
                if a + b == 0:
                    continue
                win = (prev[len(prev) - a:] if a else []) + nxt[:b]
                if any(not p.fused for p in win):
                    continue
                busy = set()
                for p in win:
                    busy.update(p.tile_bits)
                quiet = [p for p in cands if p not in busy]
                c = min(len(quiet), want_bits)
                if c < 1:
                    continue
                # more passes hide more of the exchange; more chunks shorten the pipeline's fill and drain
                score = (min(a + b, 4), min(c, 2), a + b, c)
                if best is None or score > best[0]:
                    best = (score, sorted(quiet[:c]))
        x.chunk_bits = best[1] if best else []


def compile_program(actions: Sequence[object], nl: int, rank: int, plan_local, emit=None) -> List[object]:
    """The rank-local program of a schedule: every maximal run of Apply / LocalSwap actions is resolved
    for `rank` (lower_for_rank) and handed to `plan_local(list[BitGate]) -> list[ops.Pass]`; the result is
    kept as ("local", passes).  Exchanges and peer gates stay as they are.  The program depends only on
    (actions, rank), so a compiled circuit caches it per flush (ShardedB200Backend.flush) and a replay
    skips scheduling, merging and planning.  `emit(step)` is called as soon as a step is known, so that the
    executor launches a batch while the host still plans the next one.  Pure host logic: tests/shardsim.py
    runs the same programs on virtual shards."""
    program: List[object] = []
    batch: List[BitGate] = []

    def push(step):
        program.append(step)
        if emit is not None:
            emit(step)

    def flush():
        if batch:
            passes = plan_local(list(batch))
            if passes:
                push(("local", passes))
            batch.clear()

    for a in actions:
        if isinstance(a, Apply):
            bg = lower_for_rank(a.gate, nl, rank)
            if bg is not None:
                batch.append(bg)
        elif isinstance(a, LocalSwap):
            batch.append(BitGate("swap", (a.a, a.b)))
        else:
            flush()
            push(a)
    flush()
    return program


def canonicalise(lay: Layout) -> List[object]:
    """Actions that bring the layout back to qubit q at bit n-1-q (needed before the state is
    read out in index order).  Rank-bit <-> rank-bit swaps go through a local position."""
    nl = lay.nl
    actions: List[object] = []

    def swap_positions(p1: int, p2: int):
        if p1 == p2:
            return
        a, b = lay.qubit_at(p1), lay.qubit_at(p2)
        hi, lo = max(p1, p2), min(p1, p2)
        if hi < nl:
            actions.append(LocalSwap(p1, p2))
        elif lo < nl:
            actions.append(Exchange(hi, lo))
        else:
            t = nl - 1                                          # via the top local position
            actions.append(Exchange(p1, t))
            actions.append(Exchange(p2, t))
            actions.append(Exchange(p1, t))
        lay.swap_qubits(a, b)

    for q in range(lay.n):
        want = lay.n - 1 - q
        if lay.pos[q] != want:
            swap_positions(lay.pos[q], want)
    # qipb_peer_remap needs more local bits than exchanged pairs (nbits > g): tiny shards coalesce less or not at all
    return coalesce_exchanges(actions, max_pairs=max(1, min(3, (lay.n - lay.G) - 1)))
