"""CPU tier: the real qip_b200.backend.B200Backend driven end to end on the host.

tests/hostlib.py swaps libqipb200 for a host double (fused passes -> the emulator built from the product's own
fused.cu; every other entry point restated in numpy from include/qip_b200.h) and torch.cuda for no-ops, so the
backend's own logic -- eager validation, lazy queue, swap relabelling + canonicalisation, planning, both measurement
bit orders, sampling scan, func_apply bit maps, range access, reduce_measure, the lazy product-state init -- replays
the golden op streams recorded from the unmodified reference without a GPU.  The GPU tier runs the same streams on
the device (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

import hostlib
import replay
from qip_b200.circuits import layered_stream, qfft_stream

META, STREAMS, ARRAYS = replay.load_streams()
NO_SAMPLING = [i for i, s in enumerate(STREAMS)
               if not any(op["op"] in ("measure", "soft_measure", "reduce_measure") for op in s["ops"])]


def _make(**kw):
    from qip_b200 import B200Backend

    def make(n, groups, feeds, statetype=np.complex128):
        return B200Backend.make_state(n, groups, feeds, statetype=statetype, **kw)
    return make


@pytest.mark.parametrize("i", range(len(STREAMS)), ids=[s["label"] for s in STREAMS])
def test_backend_replays_golden_stream_on_host(monkeypatch, i):
    hostlib.install(monkeypatch)
    assert replay.replay(STREAMS[i], ARRAYS, _make(), tol=1e-12)


@pytest.mark.parametrize("i", range(0, len(STREAMS), 3), ids=[STREAMS[i]["label"] for i in range(0, len(STREAMS), 3)])
def test_backend_replays_golden_stream_on_host_unfused_and_lazy_init(monkeypatch, i):
    hostlib.install(monkeypatch)
    assert replay.replay(STREAMS[i], ARRAYS, _make(fuse=False), tol=1e-12)
    assert replay.replay(STREAMS[i], ARRAYS, _make(lazy_init=True), tol=1e-12)


@pytest.mark.parametrize("i", NO_SAMPLING[::4], ids=[STREAMS[i]["label"] for i in NO_SAMPLING[::4]])
def test_backend_replays_golden_stream_on_host_complex64(monkeypatch, i):
    hostlib.install(monkeypatch)
    assert replay.replay(STREAMS[i], ARRAYS, _make(), tol=1e-5, statetype=np.complex64)


@pytest.mark.parametrize("workload", ["layered", "qft"])
def test_backend_lazy_init_takes_the_fill_path_at_production_tile_size(monkeypatch, workload):
    # 14 qubits: 2^12-amplitude tiles with 2 KiB runs -> the library serves qipb_apply_fused_fill; the result equals
    # the eagerly initialised run and no stand-alone init kernel is launched
    from oracle import oracle as orc
    L = hostlib.install(monkeypatch)
    n = 14
    rng = np.random.default_rng(2)
    groups = [[q] for q in range(n - 3)] + [[n - 2, n - 3]]          # qubit n-1 stays un-fed
    feeds = []
    for _ in range(n - 3):
        v = rng.normal(size=2) + 1j * rng.normal(size=2)
        feeds.append(v / np.linalg.norm(v))
    stream = list(layered_stream(n, 2, 1)) if workload == "layered" else list(qfft_stream(n))
    b = _make(lazy_init=True, strategy="tile")(n, groups, feeds + [2])
    c = orc.OracleBackend.make_state(n, groups, feeds + [np.eye(4)[2]])
    for mats in stream:
        b.kronselect_dot(mats)
        c.kronselect_dot(mats)
    got = np.asarray(b.get_state())
    assert b.stats.get("fill_passes") == 1 and "apply_fused_fill" in L.log and "init_kron" not in L.log
    assert float(np.max(np.abs(got - c.get_state()))) <= 1e-12
    probs = b.measure_probabilities(np.array([3, 0, 9], dtype=np.int32))
    assert np.allclose(probs, c.measure_probabilities([3, 0, 9]), rtol=0, atol=1e-13)
    b.close()


def test_compiled_circuit_replays_through_the_real_backend_with_cached_plans(monkeypatch):
    # qip_b200.graph.CompiledCircuit -> B200Backend.apply_gates (plan cache) / func_apply / measure / measure_probabilities,
    # three replays with different feed values; every replay equals the oracle run of the same ops
    import random
    from oracle import oracle as orc
    from qip_b200 import B200Backend
    from qip_b200.graph import CompiledCircuit
    from qip_b200.mats import SwapMat
    hostlib.install(monkeypatch)
    n = 13
    groups = [[q] for q in range(n)]
    f = lambda x: (5 * x + 3) % 8
    seg1 = list(layered_stream(n, 2, 4)) + [{(0, n - 1): SwapMat(1)}]
    seg2 = list(qfft_stream(6, first_qubit=2))
    ops_c = ([("k", m) for m in seg1] + [("f", [0, 1, 2], [7, 9, 11], f)] + [("k", m) for m in seg2] +
             [("p", [4, 0, 12]), ("m", [1, 6])] + [("k", m) for m in layered_stream(n, 1, 6)])
    rng = np.random.default_rng(0)

    def feeds():
        out = []
        for _ in range(n):
            v = rng.normal(size=2) + 1j * rng.normal(size=2)
            out.append(v / np.linalg.norm(v))
        return out
    first = feeds()
    circ = CompiledCircuit.from_ops(n, groups, first, ops_c)
    for replay_no, lazy in enumerate((False, True, True)):
        fl = first if replay_no == 0 else feeds()
        c = orc.OracleBackend.make_state(n, groups, fl)
        for m in seg1:
            c.kronselect_dot(m)
        c.func_apply([0, 1, 2], [7, 9, 11], f)
        for m in seg2:
            c.kronselect_dot(m)
        want_p = c.measure_probabilities([4, 0, 12])
        random.seed(7)
        want_m = c.measure([1, 6])
        for m in layered_stream(n, 1, 6):
            c.kronselect_dot(m)
        random.seed(7)
        state, classic = circ.run(feed={(q,): v for q, v in enumerate(fl)}, lazy_init=lazy)
        ip, im = len(seg1) + 1 + len(seg2), len(seg1) + 2 + len(seg2)
        assert np.allclose(classic[ip], want_p, rtol=0, atol=1e-13)
        assert classic[im][0] == want_m[0] and abs(classic[im][1] - want_m[1]) <= 1e-13
        assert float(np.max(np.abs(np.asarray(state) - c.get_state()))) <= 1e-12
        if lazy:
            assert circ.last_stats.get("fill_passes") == 1
    assert len(circ._plans) == 3                      # one cached plan per gate segment, shared by all replays


def test_whole_register_device_refeed_copies_by_default_and_adopts_on_request(monkeypatch):
    # a re-fed device state is copied (the reference never aliases a feed, qip/backend.py:86-104); with
    # adopt_feed=True the fed buffer BECOMES the state (no second 2^n allocation: the Grover loop at 33 qubits)
    from oracle import oracle as orc
    from qip_b200 import B200Backend, DeviceState
    from qip_b200.graph import CompiledCircuit
    hostlib.install(monkeypatch)
    n = 12
    ops_k = list(layered_stream(n, 1, 9))
    first = B200Backend.make_state(n, [], [], host_state_max_qubits=-1)
    for m in ops_k:
        first.kronselect_dot(m)
    handle = first.get_state()
    assert isinstance(handle, DeviceState)
    before = np.asarray(handle).copy()
    c = orc.OracleBackend.make_state(n, [], [])
    for m in ops_k + ops_k:
        c.kronselect_dot(m)
    want = c.get_state()
    b = B200Backend.make_state(n, [list(range(n))], [handle])                       # default: a copy
    assert b.state.data_ptr() != handle.tensor.data_ptr()
    for m in ops_k:
        b.kronselect_dot(m)
    assert float(np.max(np.abs(np.asarray(b.get_state()) - want))) <= 1e-12
    assert np.array_equal(np.asarray(handle), before)                                # the fed state is untouched
    a = B200Backend.make_state(n, [list(range(n))], [handle], adopt_feed=True)      # the fed buffer becomes the state
    assert a.state.data_ptr() == handle.tensor.data_ptr()
    for m in ops_k:
        a.kronselect_dot(m)
    assert float(np.max(np.abs(np.asarray(a.get_state()) - want))) <= 1e-12
    # and through CompiledCircuit.run (backend keyword arguments are forwarded to make_state)
    circ = CompiledCircuit.from_ops(n, [list(range(n))], [handle], [("k", m) for m in ops_k])
    state, _ = circ.run(feed={(0,): DeviceState(a.state)}, device_state=True, adopt_feed=True)
    assert state.tensor.data_ptr() == a.state.data_ptr()
    c3 = orc.OracleBackend.make_state(n, [list(range(n))], [want])
    for m in ops_k:
        c3.kronselect_dot(m)
    assert float(np.max(np.abs(np.asarray(state) - c3.get_state()))) <= 1e-12


@pytest.mark.parametrize("seed", range(16))
def test_random_sessions_through_the_real_backend_on_host(monkeypatch, seed):
    import fuzzlib
    from qip_b200 import B200Backend
    hostlib.install(monkeypatch)
    n = 5 + seed % 9
    fuzzlib.session(B200Backend.make_state, seed, n, lazy_init=bool(seed % 2), fuse=bool(seed % 5))


def test_function_tables_are_uploaded_once_per_function_object(monkeypatch):
    # Grover re-applies the same oracles every iteration: a function that carries its table keeps the device copy
    from qip_b200 import B200Backend
    from qip_b200 import backend as be
    from qip_b200.functions import equals, tabulated
    hostlib.install(monkeypatch)
    uploads = []
    real = be.device_table

    def counting(func, table, device, nbits_out=64):
        before = getattr(func, "_device_table", None)
        out = real(func, table, device, nbits_out)
        if getattr(func, "_device_table", None) is not before or not hasattr(func, "table"):
            uploads.append(1)
        return out
    monkeypatch.setattr(be, "device_table", counting)
    n = 8
    oracle_f = tabulated(equals(37), 7)
    b = B200Backend.make_state(n, [[q] for q in range(n)], [np.array([1.0, 1.0]) / np.sqrt(2)] * n)
    for _ in range(3):
        b.func_apply(list(range(7)), [7], oracle_f)
    assert len(uploads) == 1
    b.func_apply(list(range(7)), [7], lambda x: int(x == 37))            # a plain callable: tabulated and uploaded each time
    assert len(uploads) == 2
    st = np.asarray(b.get_state())
    assert abs(np.vdot(st, st) - 1.0) < 1e-12
    b.close()


def test_controlled_function_through_the_real_backend_on_host(monkeypatch):
    # qip_b200.functions.controlled: C(F) as a plain F on [controls ++ reg1]; byte table (3 output qubits), real B200Backend
    from oracle import oracle as orc
    from qip_b200 import B200Backend
    from qip_b200.functions import controlled, modexp
    hostlib.install(monkeypatch)
    n = 12
    rng = np.random.default_rng(21)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    ctrl, reg1, reg2 = [9, 2], [0, 7, 4, 1], [3, 11, 5, 8]
    cf = controlled(modexp(7, 15), 2, 4)
    g = B200Backend.make_state(n, [list(range(n))], [psi])
    c = orc.OracleBackend.make_state(n, [list(range(n))], [psi])
    g.func_apply(np.array(ctrl + reg1, dtype=np.int32), np.array(reg2, dtype=np.int32), cf)
    c.func_apply(ctrl + reg1, reg2, cf)
    got = np.asarray(g.get_state())
    assert float(np.max(np.abs(got - c.get_state()))) <= 1e-15
    off = np.array([not (((i >> (n - 1 - 9)) & 1) and ((i >> (n - 1 - 2)) & 1)) for i in range(2 ** n)])
    assert np.array_equal(got[off], psi[off])                                        # a control at 0: untouched


def test_graft_entry_smoke_runs_on_the_host_double(monkeypatch, capsys):
    # the driver's smoke() (one small hot-path invocation checked against the oracle) exercised end to end on the CPU tier
    import __graft_entry__ as entry
    hostlib.install(monkeypatch)
    entry.smoke()
    assert "smoke ok" in capsys.readouterr().out


@pytest.mark.parametrize("workload", ["layered", "qft"])
def test_bench_b200_arm_produces_the_contract_line_on_host_doubles(workload):
    # bench.py itself (not only the engine) must keep working between GPU runs: its B200 arm end to end on the host doubles
    import json
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "bench_on_host.py"), "--qubits", "14", "--steps", "2", "--warmup", "1",
                        "--workload", workload, "--no-micro", "--no-cpu"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "cpu_baseline"):
        assert key in d, key
    assert d["metric"] == "gate_apply_GBps" and d["n_gpus"] == 1 and d["steps"] == 2 and d["config"]["qubits"] == 14
    assert d["gpu_launches"] >= 1 and d["roofline"]["kernel"].startswith("fused_kernel") and d["roofline"]["bound"] == "hbm"
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0


def test_bench_reference_arm_prints_the_contract_line():
    # bench.py --impl reference: the reference's own CPU kernels (oracle/_ref, else the C port) on a bounded sample
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "qft",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert d["impl"] == "reference" and d["metric"] == "gate_apply_GBps" and d["unit"] == "GB/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 exit quietly under torchrun
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_wide_top_k_is_ranked_on_the_device_with_the_same_tie_order(monkeypatch):
    # measure_probabilities(top_k) over more than 16 qubits sorts the histogram where it is (stable, descending: equal
    # probabilities in ascending outcome order, like the host path) and brings only the winners back
    from qip_b200 import B200Backend
    from qip_b200 import backend as be
    hostlib.install(monkeypatch)
    n = 17
    feeds = [np.array([1.0, 1.0]) / np.sqrt(2)] * 14 + [np.array([0.6, 0.8]), np.array([1.0, 0.0]), np.array([0.8, 0.6])]
    b = B200Backend.make_state(n, [[q] for q in range(n)], feeds)
    idx = np.arange(n, dtype=np.int32)
    got_i, got_p = b.measure_probabilities(idx, top_k=9)
    monkeypatch.setattr(be, "_DEVICE_TOPK_MIN_QUBITS", 64)
    want_i, want_p = b.measure_probabilities(idx, top_k=9)
    assert got_i == want_i and got_p == want_p
    assert got_i == sorted(got_i) and len(set(got_p)) == 1          # nine tied winners, ascending outcomes
    b.close()
