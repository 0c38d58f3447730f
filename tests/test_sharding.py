"""CPU tier: the multi-GPU path's host logic (qip_b200/shardplan.py).

  * virtual shards: P = 2, 4, 8 logical shards in one process, executed with numpy, against the
    single-shard result and the oracle -- every action kind (rank-resolved controls and diagonals,
    relabelled swaps, fused peer gates, bit exchanges, canonicalisation);
  * world_size-2 gloo run: two processes, one shard each, exchanging over torch.distributed.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import bitsim
import shardsim
from oracle import oracle as orc
from qip_b200 import ops
from qip_b200 import shardplan as sp
from qip_b200.circuits import H2, X2, haar_unitary, layered_stream, qfft_stream, rm_mat
from qip_b200.mats import CMat, SwapMat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def logical_gates(stream, n, merge=0):
    gates = []
    for mats in stream:
        for g in ops.decode_mats(mats, n):
            s = ops.simplify(g)
            if s is not None:
                gates.append(s)
    return ops.merge_blocks(gates, merge) if merge else gates


def run_sharded(psi, gates, n, gbits):
    lay = sp.Layout(n, gbits)
    actions = sp.schedule(gates, lay)
    vs = shardsim.VirtualShards(psi, gbits)
    vs.run(actions)
    vs.run(sp.canonicalise(lay))
    assert lay.canonical()
    return vs, actions


def reference_state(psi, gates, n):
    st = psi.copy()
    for g in gates:
        st = bitsim.apply_bitgate(st, ops.lower(g, n), n)
    return st


@pytest.mark.parametrize("gbits", [1, 2, 3])
@pytest.mark.parametrize("seed", range(4))
def test_layered_circuit_on_virtual_shards(gbits, seed):
    n = 9
    rng = np.random.default_rng(seed)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    gates = logical_gates(layered_stream(n, 3, seed), n)
    vs, actions = run_sharded(psi, gates, n, gbits)
    assert float(np.max(np.abs(vs.gather() - reference_state(psi, gates, n)))) <= 1e-13
    if gbits >= 2:
        assert any(isinstance(a, (sp.Exchange, sp.MultiExchange, sp.PeerGate1)) for a in actions)


def _permuted(psi, lay):
    """The state as it is laid out physically: bit lay.pos[q] of the physical index is qubit q."""
    n = lay.n
    idx = np.arange(2 ** n, dtype=np.int64)
    logical = np.zeros(2 ** n, dtype=np.int64)
    for q in range(n):
        logical |= ((idx >> lay.pos[q]) & 1) << (n - 1 - q)
    return psi[logical]


@pytest.mark.parametrize("gbits", [1, 2, 3])
@pytest.mark.parametrize("mode", ["lazy_layout", "peer_gates", "plain"])
def test_qft_on_virtual_shards_needs_few_exchanges(gbits, mode):
    n = 10
    rng = np.random.default_rng(gbits)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    gates = logical_gates(qfft_stream(n), n)
    lay = sp.Layout(n, gbits)
    if mode == "lazy_layout":
        sp.choose_initial_layout(gates, lay)        # the last qubits to be Hadamard-ed go on the rank bits
        assert sorted(q for q in range(n) if lay.is_global(q)) == list(range(n - gbits, n))
    start = _permuted(psi, lay)
    actions = sp.schedule(gates, lay, peer_gates=(mode == "peer_gates"))
    moves = []
    for a in actions:
        if isinstance(a, sp.MultiExchange):
            moves.extend(sp.Exchange(g, l) for g, l in a.pairs)
        elif isinstance(a, (sp.Exchange, sp.PeerGate1)):
            moves.append(a)
    # every C-phase is communication-free and the final bit reversal is a relabel
    if mode == "lazy_layout":
        assert len(moves) == gbits and all(isinstance(a, sp.Exchange) for a in moves)
    elif mode == "peer_gates":
        assert len(moves) == gbits and all(isinstance(a, sp.PeerGate1) for a in moves)
    else:
        assert len(moves) <= 2 * gbits and all(isinstance(a, sp.Exchange) for a in moves)
    vs = shardsim.VirtualShards(start, gbits)
    vs.run(actions)
    vs.run(sp.canonicalise(lay))
    want = np.fft.ifft(psi) * np.sqrt(2 ** n)
    assert float(np.max(np.abs(vs.gather() - want))) <= 1e-12


def test_rank_resolved_controls_diagonals_and_relabels_need_no_exchange():
    n, gbits = 8, 2
    rng = np.random.default_rng(0)
    stream = [{(0, 5): CMat(X2)},                      # control on a rank qubit, local target
              {(1, 0, 6): CMat(CMat(haar_unitary(rng, 2)))},
              {0: rm_mat(3)}, {(1, 0): CMat(rm_mat(2))},        # diagonal on rank qubits
              {(0, 4): np.diag(np.exp(1j * rng.normal(size=4)))},
              {(0, 7): SwapMat(1)}, {(1, 2): SwapMat(1)}]        # swaps with rank qubits: relabels
    gates = logical_gates(stream, n)
    lay = sp.Layout(n, gbits)
    actions = sp.schedule(gates, lay)
    assert not any(isinstance(a, (sp.Exchange, sp.MultiExchange, sp.PeerGate1)) for a in actions)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    vs = shardsim.VirtualShards(psi, gbits)
    vs.run(actions)
    fix = sp.canonicalise(lay)
    assert any(isinstance(a, (sp.Exchange, sp.MultiExchange)) for a in fix)   # the relabels are paid for at read-out only
    vs.run(fix)
    assert float(np.max(np.abs(vs.gather() - reference_state(psi, gates, n)))) <= 1e-13


def test_lower_for_rank_skips_and_selects():
    nl = 4
    pg = sp.PhysGate((1,), (1 << 5) | (1 << 2), X2.astype(np.complex128), False)
    assert sp.lower_for_rank(pg, nl, 0b01) is None                      # rank bit 5-4=1 is 0 on rank 1
    bg = sp.lower_for_rank(pg, nl, 0b10)
    assert bg.bits == (1,) and bg.ctrl_mask == (1 << 2)
    d = np.diag([1, 2, 3, 4]).astype(np.complex128)
    pg = sp.PhysGate((4, 0), 0, d, True)                                 # MSB target is rank bit 0
    assert np.array_equal(np.diag(sp.lower_for_rank(pg, nl, 0).mat), [1, 2])
    assert np.array_equal(np.diag(sp.lower_for_rank(pg, nl, 1).mat), [3, 4])
    pg = sp.PhysGate((5,), 0, np.diag([1, 1j]).astype(np.complex128), True)
    assert sp.lower_for_rank(pg, nl, 0) is None and sp.lower_for_rank(pg, nl, 2).mat[0, 0] == 1j
    with pytest.raises(ValueError):
        sp.lower_for_rank(sp.PhysGate((5,), 0, H2.astype(np.complex128), False), nl, 0)


def test_merged_blocks_schedule_on_shards():
    n, gbits = 9, 2
    rng = np.random.default_rng(3)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    gates = logical_gates(layered_stream(n, 2, 11), n)
    merged = ops.merge_blocks(gates, 4)
    vs, _ = run_sharded(psi, merged, n, gbits)
    assert float(np.max(np.abs(vs.gather() - reference_state(psi, gates, n)))) <= 1e-12


def test_two_process_gloo_shards_match_oracle(tmp_path):
    """world_size 2, gloo backend, one shard per process (tests/gloo_worker.py)."""
    out = tmp_path / "res.npy"
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "gloo_worker.py"), str(out)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = np.load(out)
    n = 8
    rng = np.random.default_rng(42)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    c = orc.OracleBackend.make_state(n, [list(range(n))], [psi])
    for mats in list(layered_stream(n, 2, 7)) + list(qfft_stream(n)):
        c.kronselect_dot(mats)
    assert float(np.max(np.abs(got - c.get_state()))) <= 1e-12


def test_consecutive_exchanges_coalesce_into_one_remap():
    acts = [sp.Exchange(9, 3), sp.Exchange(8, 5), sp.Exchange(10, 6), sp.Exchange(9, 2), sp.LocalSwap(0, 1), sp.Exchange(8, 4)]
    out = sp.coalesce_exchanges(acts)
    assert isinstance(out[0], sp.MultiExchange) and out[0].pairs == [(9, 3), (8, 5), (10, 6)]
    assert isinstance(out[1], sp.Exchange) and isinstance(out[2], sp.LocalSwap) and isinstance(out[3], sp.Exchange)
    # a batch of rank qubits needed at once (layered circuit on 8 shards) is ONE remap
    n, gbits = 10, 3
    gates = logical_gates(layered_stream(n, 2, 5), n)
    lay = sp.Layout(n, gbits)
    actions = sp.schedule(gates, lay)
    assert any(isinstance(a, sp.MultiExchange) and len(a.pairs) >= 2 for a in actions)


@pytest.mark.parametrize("gbits", [1, 2, 3])
def test_deferred_global_gates_need_one_exchange_per_layer(gbits):
    # shardplan.defer_global: gates on disjoint qubits commute, so everything that needs a rank-bit qubit moves
    # behind the local work and a layer that touches every qubit is local work -> ONE remap -> a short tail
    n = 14
    lay = sp.Layout(n, gbits)
    for seed in range(6):
        gates = logical_gates(layered_stream(n, 1, 40 + seed), n)
        reordered = sp.defer_global(gates, lay)
        assert sorted(map(id, reordered)) == sorted(map(id, gates))               # a permutation of the same gates
        # dependent gates (sharing a qubit) keep their relative order
        where = {id(g): i for i, g in enumerate(reordered)}
        for i, a in enumerate(gates):
            qa = set(a.targets) | set(a.controls)
            for b in gates[i + 1:]:
                if qa & (set(b.targets) | set(b.controls)):
                    assert where[id(a)] < where[id(b)]
        actions = sp.schedule(gates, lay)                                         # lay carries over: layer after layer
        moves = [a for a in actions if isinstance(a, (sp.Exchange, sp.MultiExchange, sp.PeerGate1))]
        assert len(moves) <= 1, (seed, moves)
    # and the result is still the circuit: virtual shards against the in-order simulator
    rng = np.random.default_rng(3)
    psi = rng.normal(size=2 ** 9) + 1j * rng.normal(size=2 ** 9)
    psi /= np.linalg.norm(psi)
    gates = logical_gates(list(layered_stream(9, 2, 5)) + [{(0, 8): SwapMat(1)}, {0: H2}, {(1, 0): CMat(X2)}, {(8, 0): SwapMat(1)}, {8: H2}], 9)
    vs, _ = run_sharded(psi, gates, 9, gbits)
    assert float(np.max(np.abs(vs.gather() - reference_state(psi, gates, 9)))) <= 1e-13


@pytest.mark.parametrize("gbits", [1, 2, 3])
@pytest.mark.parametrize("kind", ["layered", "qft", "mixed"])
def test_rank_local_merge_and_plan_on_virtual_shards(gbits, kind):
    # the rank-local pipeline of ShardedB200Backend._plan_local (merge_bitgates with lone-diagonal clustering ->
    # plan_passes -> fused / stand-alone passes), executed per virtual shard by the numpy pass executor
    n = 10
    rng = np.random.default_rng(17 + gbits)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    if kind == "layered":
        stream = list(layered_stream(n, 3, 21))
    elif kind == "qft":
        stream = list(qfft_stream(n))
    else:
        stream = [{0: rm_mat(3)}, {(1, 2): haar_unitary(rng, 4)}, {3: rm_mat(2)}, {(0, 4): CMat(X2)}, {9: rm_mat(4)},
                  {(3, 6): haar_unitary(rng, 4)}, {(2, 5): CMat(rm_mat(3))}, {5: H2}, {1: rm_mat(6)}, {(6, 1): CMat(X2)},
                  {0: H2}, {(9, 0): SwapMat(1)}, {(0, 7): haar_unitary(rng, 4)}, {8: rm_mat(2)}, {7: rm_mat(5)}]
    gates = logical_gates(stream, n)
    lay = sp.Layout(n, gbits)
    vs = shardsim.VirtualShards(_permuted(psi, lay), gbits)
    vs.run_planned(sp.schedule(gates, lay))
    vs.run_planned(sp.canonicalise(lay))
    assert lay.canonical()
    assert float(np.max(np.abs(vs.gather() - reference_state(psi, gates, n)))) <= 1e-12


def _host_only_backend(n, gbits, rank, tile_bits=5, min_low_bits=2):
    """A ShardedB200Backend with its device side cut off: the real apply_gates / kronselect_dot / flush
    run (queueing, schedule, rank-local planning, launching step by step, the compiled-circuit program
    cache); only the four launch helpers are replaced by recorders."""
    import types
    from qip_b200.sharded import ShardedB200Backend
    b = object.__new__(ShardedB200Backend)
    b.n, b.G, b.nl, b.rank, b.P = n, gbits, n - gbits, rank, 1 << gbits
    b.layout = sp.Layout(n, gbits)
    b.queue, b._seg_cache, b._seg_keys = [], None, []
    b.fuse, b.peer_gates, b.amp_bytes = True, False, 16
    b.eng = types.SimpleNamespace(tile_bits=tile_bits, min_low_bits=min_low_bits)
    b.stats = {"gates": 0}
    b._pending_init, b._virtual_init, b.lazy_init = None, None, False
    b.overlap = False                               # (the chunk pipeline needs the device side: tests/test_sharded_on_host.py)
    b.device = -1                                   # torch.cuda.device(-1) is a no-op context
    b._stream = lambda: None
    b.programs = []                                 # one list of launched steps per flush

    def launch(step):
        b.programs[-1].append(step)
    b._run_passes = lambda passes: launch(("local", passes))
    b._exchange = b._multi_exchange = b._peer_gate = launch
    real_flush = b.flush

    def flush():
        b.programs.append([])
        real_flush()
        if not b.programs[-1]:
            b.programs.pop()
    b.flush = flush
    return b


@pytest.mark.parametrize("gbits", [1, 2])
def test_sharded_flush_caches_the_rank_local_program_of_compiled_segments(gbits):
    # qip_b200.graph.CompiledCircuit.run hands every gate segment to apply_gates(gates, cache, key); the flush
    # that executes them must replay the cached program only for the same segments from the same layout, and
    # the replayed program must still be the circuit
    n = 9
    P = 1 << gbits
    rng = np.random.default_rng(5)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    seg_a = logical_gates(layered_stream(n, 2, 7), n)
    seg_b = logical_gates(list(qfft_stream(n)), n)
    caches = [dict() for _ in range(P)]            # one process (and one CompiledCircuit cache) per rank
    first_programs = None
    for replay in range(3):
        backends = [_host_only_backend(n, gbits, r) for r in range(P)]
        for r, b in enumerate(backends):
            b.apply_gates(seg_a, caches[r], ("seg", 0))
            b.apply_gates(seg_b, caches[r], ("seg", 1))
            b.flush()                               # both segments in one flush: one schedule, one cache entry
            b.apply_gates(seg_a, caches[r], ("seg", 0))
            b.flush()                               # same segment, different starting layout: its own entry
            assert b.queue == [] and len(b.programs) == 2
            assert b.stats.get("cached_flushes", 0) == (0 if replay == 0 else 2)
            assert len(caches[r]) == 2
        if replay == 0:
            first_programs = [b.programs for b in backends]
        else:
            for b, first in zip(backends, first_programs):
                for prog, first_prog in zip(b.programs, first):             # replays reuse the planned passes
                    assert len(prog) == len(first_prog)
                    assert all((x[1] is y[1]) if isinstance(x, tuple) else (x is y) for x, y in zip(prog, first_prog))
        vs = shardsim.VirtualShards(psi, gbits)
        for step in range(2):
            vs.run_programs([b.programs[step] for b in backends])
        lay = backends[0].layout
        assert all(b.layout.pos == lay.pos for b in backends)
        vs.run_planned(sp.canonicalise(lay))
        assert float(np.max(np.abs(vs.gather() - reference_state(psi, seg_a + seg_b + seg_a, n)))) <= 1e-12

    # un-keyed gates in the queue make the flush un-cacheable (and leave the cache alone)
    b = _host_only_backend(n, gbits, 0)
    b.apply_gates(seg_a, caches[0], ("seg", 0))
    b.kronselect_dot({0: H2})
    b.flush()
    assert b.stats.get("cached_flushes", 0) == 0 and len(caches[0]) == 2
    b.apply_gates(seg_a, caches[0], ("seg", 0))    # the next flush is cacheable again
    b.flush()
    assert len(caches[0]) == 3


@pytest.mark.parametrize("gbits", [1, 2, 3])
def test_hoisted_exchange_saves_the_tail_pass_of_a_sharded_qft(gbits, monkeypatch):
    # shardplan.choose_prefetch: the rank-bit qubits of a QFT come in right after the first passes (their victims are
    # finished by then), so their Hadamards share the remaining passes instead of needing one more sweep after the remap
    from qip_b200.ops import merge_bitgates, plan_passes
    n, tb, low = 13, 5, 2
    nl = n - gbits
    gates = logical_gates(list(qfft_stream(n)), n)

    def plan_local(batch):
        return plan_passes(merge_bitgates(batch, 2), nl, 16, tile_bits=min(tb, nl), min_low_bits=low)

    def program_cost(hoist):
        monkeypatch.setenv("QIPB_SHARD_HOIST", hoist)
        lay = sp.Layout(n, gbits)
        sp.choose_initial_layout(gates, lay)
        start = lay.copy()
        actions = sp.schedule(gates, lay, count_passes=lambda b: len(plan_local(b)), tile_bits=tb, min_low_bits=low)
        prog = sp.compile_program(actions, nl, (1 << gbits) - 1, plan_local)
        passes = sum(len(st[1]) for st in prog if isinstance(st, tuple))
        moves = sum(1 for st in prog if not isinstance(st, tuple))
        return passes, moves, actions, start, lay

    base_passes, base_moves, _, _, _ = program_cost("0")
    passes, moves, actions, start, lay = program_cost("1")
    assert moves == base_moves == 1 and passes <= base_passes, (passes, base_passes)
    if gbits == 1:
        assert passes < base_passes                    # (with more rank bits the last local pass of this small case has room)
    # and it is still the QFT: virtual shards, rank-local planned programs, against the in-order simulator
    rng = np.random.default_rng(gbits)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    vs = shardsim.VirtualShards(_permuted(psi, start), gbits)
    vs.run_planned(actions, tile_bits=tb, min_low_bits=low)
    vs.run_planned(sp.canonicalise(lay), tile_bits=tb, min_low_bits=low)
    assert lay.canonical()
    assert float(np.max(np.abs(vs.gather() - reference_state(psi, gates, n)))) <= 1e-12


@pytest.mark.parametrize("gbits", [1, 2, 3])
@pytest.mark.parametrize("seed", range(6))
def test_hoisted_schedules_of_layered_circuits_are_the_circuit(gbits, seed):
    # layer after layer on one evolving layout (what the benchmark does): whatever choose_prefetch decides, the
    # result is the circuit and a layer still needs at most one move
    n, tb, low = 10, 5, 2
    lay = sp.Layout(n, gbits)
    rng = np.random.default_rng(seed)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    vs = shardsim.VirtualShards(psi, gbits)
    want = psi.copy()
    hoisted = 0
    for layer in range(4):
        gates = logical_gates(list(layered_stream(n, 1, 10 * seed + layer)), n)
        actions = sp.schedule(gates, lay, tile_bits=tb, min_low_bits=low)
        moves = [i for i, a in enumerate(actions) if isinstance(a, (sp.Exchange, sp.MultiExchange))]
        assert len(moves) <= 1
        hoisted += bool(moves) and any(isinstance(a, sp.Apply) for a in actions[:moves[0]]) and moves[0] < len(actions) - 12
        vs.run_planned(actions, tile_bits=tb, min_low_bits=low)
        want = reference_state(want, gates, n)
    vs.run_planned(sp.canonicalise(lay), tile_bits=tb, min_low_bits=low)
    assert float(np.max(np.abs(vs.gather() - want))) <= 1e-12


def test_hoisted_exchange_at_benchmark_size_36_qubits_on_8_shards(monkeypatch):
    # planning only (no state): BASELINE configs[4] -- 6 fused passes + ONE 3-bit remap instead of 6 + remap + 1
    from qip_b200.ops import merge_bitgates, plan_passes
    n, gbits = 36, 3
    nl = n - gbits
    gates = logical_gates(list(qfft_stream(n)), n)

    def plan_local(batch):
        return plan_passes(merge_bitgates(batch, 2), nl, 16)

    shape = {}
    for hoist in ("0", "1"):
        monkeypatch.setenv("QIPB_SHARD_HOIST", hoist)
        lay = sp.Layout(n, gbits)
        sp.choose_initial_layout(gates, lay)
        prog = sp.compile_program(sp.schedule(gates, lay, count_passes=lambda b: len(plan_local(b))), nl, 7, plan_local)
        shape[hoist] = (sum(len(st[1]) for st in prog if isinstance(st, tuple)),
                        [len(st.pairs) for st in prog if isinstance(st, sp.MultiExchange)])
    assert shape["0"] == (7, [3]) and shape["1"] == (6, [3]), shape
