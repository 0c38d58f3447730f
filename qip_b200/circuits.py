"""Op-stream generators for the BASELINE.json configs.

Each generator yields the `mats` dicts the reference front-end would hand to
StateType.kronselect_dot for that circuit, so that benchmarks and GPU tests can run without the
reference installed.  They are cross-checked against the real front-end in
tests/test_host_logic.py (CPU tier, where /root/reference exists) and against the golden streams
recorded by tests/golden/make_golden.py.
"""
import numpy as np

from .mats import CMat, SwapMat

H2 = (1 / np.sqrt(2)) * np.array([[1, 1], [1, -1]])          # qip/operators.py:97-99 (float64 2x2)
X2 = np.flip(np.eye(2), 0)                                   # qip/operators.py:41-43


def rm_mat(m: int, negate: bool = False):
    """Rm(m): diag(1, e^{+-2 pi i / 2^m}), qip/operators.py:108-127."""
    phi = (-2 if negate else 2) * np.pi / pow(2.0, m)
    return np.array([[1, 0], [0, np.exp(1.0j * phi)]])


def qfft_stream(n: int, rev: bool = True, first_qubit: int = 0):
    """QFFT over qubits first_qubit .. first_qubit+n-1, qip/qfft.py:8-43: for k ascending H(q_k), then
    C(Rm(1+i-k)) with key (q_i, q_k) for i > k; finally the Swap reversal.  The phase sign is +
    because the front-end drops negate=True."""
    q = [first_qubit + i for i in range(n)]
    for k in range(n):
        yield {q[k]: H2}
        for i in range(k + 1, n):
            yield {(q[i], q[k]): CMat(rm_mat(1 + i - k))}
    if rev:
        for i in range(n // 2):
            yield {(q[i], q[n - 1 - i]): SwapMat(1)}


def haar_unitary(rng, d: int):
    z = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    qm, r = np.linalg.qr(z)
    return qm * (np.diag(r) / np.abs(np.diag(r)))


def layered_stream(n: int, depth: int, seed: int):
    """Random layered circuit of SURVEY 8d config 4 (same draws as tests/golden/make_golden.py)."""
    rng = np.random.default_rng(seed)
    for _ in range(depth):
        for i in range(n):
            if rng.random() < 0.5:
                yield {i: H2}
            else:
                yield {i: rm_mat(int(rng.integers(1, 9)))}
        perm = rng.permutation(n)
        for a, b in zip(perm[0::2], perm[1::2]):
            a, b = int(a), int(b)
            kind = int(rng.integers(0, 3))
            if kind == 0:
                yield {(a, b): CMat(X2)}
            elif kind == 1:
                yield {(a, b): SwapMat(1)}
            else:
                yield {(a, b): haar_unitary(rng, 4)}


def layer_gate_count(n: int) -> int:
    return n + n // 2


def inverse_stream(stream):
    """Adjoint circuit: reversed order, each matrix conjugate-transposed (MatrixOp dagger,
    qip/operators.py:19-23)."""
    ops = list(stream)
    for mats in reversed(ops):
        yield {k: (v.conj().T if isinstance(v, np.ndarray) else v.conj().T) for k, v in mats.items()}


def grover_iteration(nsearch: int, x0: int):
    """One Grover iteration on qubits 0..nsearch-1 + ancilla nsearch, the structure of
    examples/grovers_iterative.py:20-39: F(x == x0), H(search), F(x == 0), H(search).
    Yields ("f", reg1, reg2, func) and ("m", mats) items."""
    search = list(range(nsearch))
    anc = [nsearch]
    yield ("f", search, anc, lambda x: (x == x0) * 1)
    yield ("m", {i: H2 for i in search})
    yield ("f", search, anc, lambda x: (x == 0) * 1)
    yield ("m", {i: H2 for i in search})
