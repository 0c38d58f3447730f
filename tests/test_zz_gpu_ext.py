"""GPU tier (-m gpu), run last: the opt-in forms of the fused pass (QIPB_FUSED_EXT=1 -- real 1-qubit sweeps and two QFT
steps per sweep, qip_b200/csrc/fused.cu "EXT sweeps").  Off by default until measured on B200; their lowering and
arithmetic are covered on the CPU tier by tests/test_fused_emul.py, this file runs the EXT kernel on the device."""
import numpy as np
import pytest

from oracle import oracle as orc
from qip_b200.circuits import H2, X2, haar_unitary, layered_stream, qfft_stream, rm_mat
from qip_b200.mats import CMat

# first run of these kernels on a device happens at the end of the round: bound every test (pytest-timeout, whole
# process on expiry -- this file sorts last, so nothing else is cut off)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]


def _rand_state(rng, n):
    v = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    return v / np.linalg.norm(v)


@pytest.mark.parametrize("statetype,tol", [(np.complex128, 1e-12), (np.complex64, 1e-5)])
@pytest.mark.parametrize("n", [16, 21, 24])
def test_ext_qfft_closed_form(monkeypatch, statetype, tol, n):
    from qip_b200 import B200Backend
    monkeypatch.setenv("QIPB_FUSED_EXT", "1")
    rng = np.random.default_rng(n)
    psi = _rand_state(rng, n)
    g = B200Backend.make_state(n, [list(range(n))], [psi], statetype=statetype)
    for mats in qfft_stream(n):
        g.kronselect_dot(mats)
    out = np.asarray(g.get_state())
    assert g.ext_launch_count() >= 1
    want = np.fft.ifft(psi) * np.sqrt(2 ** n)                 # QFFT == sqrt(N) * ifft
    assert float(np.max(np.abs(out - want))) / float(np.max(np.abs(want))) <= tol
    g.close()


@pytest.mark.parametrize("statetype,tol", [(np.complex128, 1e-12), (np.complex64, 1e-5)])
def test_ext_real_gates_and_unpaired_steps_match_oracle(monkeypatch, statetype, tol):
    from qip_b200 import B200Backend
    monkeypatch.setenv("QIPB_FUSED_EXT", "1")
    n = 16
    rng = np.random.default_rng(3)
    psi = _rand_state(rng, n)
    ry = np.array([[np.cos(0.3), -np.sin(0.3)], [np.sin(0.3), np.cos(0.3)]])
    stream = list(layered_stream(n, 2, 7))
    stream += [{0: ry}, {1: X2}, {(2, 3): CMat(ry)}, {15: ry}, {14: H2}, {5: H2}]
    stream += [{(q, 5): CMat(rm_mat(2 + q % 3))} for q in (0, 1, 2, 9)]
    stream += [{6: H2}] + [{(q, 6): CMat(rm_mat(3))} for q in (7, 8, 10)]
    stream += list(qfft_stream(6, first_qubit=1)) + list(qfft_stream(5, first_qubit=9)) + [{(0, 4): haar_unitary(rng, 4)}]
    g = B200Backend.make_state(n, [list(range(n))], [psi], statetype=statetype, strategy="tile")
    c = orc.OracleBackend.make_state(n, [list(range(n))], [psi])
    for mats in stream:
        g.kronselect_dot(mats)
        c.kronselect_dot(mats)
    a, b = np.asarray(g.get_state()), c.get_state()
    assert g.ext_launch_count() >= 1
    assert float(np.max(np.abs(a - b))) / float(np.max(np.abs(b))) <= tol
    g.close()


@pytest.mark.parametrize("statetype,tol", [(np.complex128, 1e-12), (np.complex64, 1e-5)])
@pytest.mark.parametrize("workload", ["layered", "qft"])
def test_lazy_product_state_init_is_fused_into_the_first_pass(monkeypatch, statetype, tol, workload):
    # QIPB_LAZY_INIT=1: a product of one-qubit feeds (+ one-hot and un-fed qubits) is never written by an init kernel;
    # the first fused pass builds its tiles from the per-bit factors (qipb_apply_fused_fill)
    from qip_b200 import B200Backend
    monkeypatch.setenv("QIPB_LAZY_INIT", "1")
    n = 16
    rng = np.random.default_rng(9)
    qubits = [int(q) for q in rng.permutation(n)]
    hot_group, vec = qubits[:3], qubits[5:]                    # qubits[3:5] stay un-fed
    groups = [[q] for q in vec] + [hot_group]
    vfeeds = []
    for _ in vec:
        v = rng.normal(size=2) + 1j * rng.normal(size=2)
        vfeeds.append(v / np.linalg.norm(v))
    stream = list(layered_stream(n, 2, 5)) if workload == "layered" else list(qfft_stream(n))
    g = B200Backend.make_state(n, groups, vfeeds + [5], statetype=statetype, strategy="tile")
    c = orc.OracleBackend.make_state(n, groups, vfeeds + [np.eye(8)[5]])
    for mats in stream:
        g.kronselect_dot(mats)
        c.kronselect_dot(mats)
    a, b = np.asarray(g.get_state()), c.get_state()
    assert g.stats.get("fill_passes", 0) == 1 and g.ext_launch_count() >= 1
    assert float(np.max(np.abs(a - b))) / float(np.max(np.abs(b))) <= tol
    g.close()
    # observed before any gate: the stand-alone init kernel builds it
    g = B200Backend.make_state(n, groups, vfeeds + [5], statetype=statetype)
    c = orc.OracleBackend.make_state(n, groups, vfeeds + [np.eye(8)[5]])
    assert float(np.max(np.abs(np.asarray(g.get_state()) - c.get_state()))) <= tol
    assert "fill_passes" not in g.stats
    g.close()
