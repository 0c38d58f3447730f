"""Generate tests/golden/ref_streams.{json,npz}: golden op streams recorded from the UNMODIFIED
reference (python front-end from /root/reference + its Cython kernels built into oracle/_ref).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Every `run()` the reference performs is recorded AT THE BACKEND BOUNDARY:
the make_state arguments, every kronselect_dot / func_apply / measure / measure_probabilities /
soft_measure / reduce_measure call with its arguments, the value each call returned, the uniform
draw `random.random()` handed to soft_measure, and the final state.  Replaying a stream through
any backend with the same surface must reproduce the recorded returns and final amplitudes.

Sources of streams:
  1. the reference's own test-suite (tests/qiptest.py, qubit_util_test.py, qfttest.py: every test);
  2. the BASELINE.json configs at oracle-sized n, built with the reference front-end (README CSwap
     11 q, examples/cswap_measure.py 7 q, QFFT 8/10 q, Grover 6+1 q, random layered circuit 8/10 q);
  3. the paths no reference test pins: measure_probabilities bit order, top-k,
     soft_measure with seeds, soft_measure(measured=), measure(measured=, measured_prob=),
     reduce_measure, total_prob.
tests/utiltest.py's kernel-level KATs (kronselect_dot vs numpy.kron incl. offsets; func_apply) are
re-stated directly in tests/test_oracle.py since they do not go through run().
"""
import json
import os
import random
import sys
import unittest

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_loader import import_reference_qip  # noqa: E402

qip = import_reference_qip()
from qip.backend import CythonBackend  # noqa: E402
import qip.pipeline  # noqa: E402
from qip.qip import Qubit, Measure, StochasticMeasure  # noqa: E402
from qip.operators import H, C, Not, Swap, Rm, R, F, X, Y, Z, MatrixOp, CMat, SwapMat  # noqa: E402
from qip.qubit_util import QubitOpWrapper  # noqa: E402
from qip.qfft import QFFT  # noqa: E402
from qip.pipeline import run  # noqa: E402


class Store:
    def __init__(self):
        self.arrays = {}
        self.streams = []

    def put(self, a):
        a = np.asarray(a)
        key = "a%d" % len(self.arrays)
        self.arrays[key] = a
        return "@" + key

    def put_state(self, a):
        a = np.asarray(a, dtype=np.complex128)
        nz = np.flatnonzero(a)
        if len(a) > 4096 and len(nz) * 8 <= len(a):
            return {"sparse": True, "len": int(len(a)), "idx": self.put(nz.astype(np.int64)), "val": self.put(a[nz])}
        return {"sparse": False, "val": self.put(a)}


STORE = Store()
_ORIG_MAKE_STATE = CythonBackend.make_state   # captured before run()'s default is patched
CURRENT_LABEL = ["?"]
DRAWS = []

_real_random = random.random


def _recording_random():
    v = _real_random()
    DRAWS.append(v)
    return v


random.random = _recording_random


def enc_mat(m):
    ks = getattr(m, "_kron_struct", None)
    if ks == 2:
        return {"type": "C", "m": enc_mat(m.m)}
    if ks == 3:
        return {"type": "swap", "n": int(m.n)}
    return {"type": "dense", "val": STORE.put(np.asarray(m, dtype=np.complex128))}


def enc_mats(mats):
    return [{"key": (list(k) if isinstance(k, tuple) else int(k)), "is_int_key": not isinstance(k, tuple),
             "mat": enc_mat(v)} for k, v in mats.items()]


class Recorder(CythonBackend):
    """CythonBackend that logs every boundary call (qip/backend.py:68-175)."""

    @staticmethod
    def make_state(n, index_groups, feed_list, statetype=np.complex128, **kw):
        inner = _ORIG_MAKE_STATE(n, index_groups, feed_list, statetype=statetype, **kw)
        rec = Recorder(inner.n, inner.state, inner.arena)
        rec.stream = {
            "label": CURRENT_LABEL[0],
            "n": int(n),
            "index_groups": [[int(i) for i in g] for g in index_groups],
            "feeds": [STORE.put(np.asarray(f, dtype=np.complex128)) for f in feed_list],
            "ops": [],
        }
        STORE.streams.append(rec.stream)
        return rec

    def get_state(self):
        s = super().get_state()
        self.stream["final_state"] = STORE.put_state(s)
        return s

    def kronselect_dot(self, mats, input_offset=0, output_offset=0):
        self.stream["ops"].append({"op": "kronselect_dot", "mats": enc_mats(mats)})
        return super().kronselect_dot(mats, input_offset, output_offset)

    def func_apply(self, reg1_indices, reg2_indices, func, input_offset=0, output_offset=0):
        r1 = [int(i) for i in reg1_indices]
        r2 = [int(i) for i in reg2_indices]
        table = [int(func(x)) for x in range(2 ** len(r1))]
        self.stream["ops"].append({"op": "func_apply", "reg1": r1, "reg2": r2,
                                   "table": STORE.put(np.array(table, dtype=np.int64))})
        return super().func_apply(reg1_indices, reg2_indices, func, input_offset, output_offset)

    def total_prob(self):
        # compiled prob_magnitude has an uninitialised accumulator: record the
        # intended value
        v = float(np.sum(np.abs(self.state) ** 2))
        self.stream["ops"].append({"op": "total_prob", "ret": v})
        return v

    def measure(self, indices, measured=None, measured_prob=None, input_offset=0, output_offset=0):
        DRAWS.clear()
        m, p = super().measure(indices, measured=measured, measured_prob=measured_prob)
        self.stream["ops"].append({"op": "measure", "indices": [int(i) for i in indices],
                                   "measured": measured, "measured_prob": measured_prob,
                                   "draws": list(DRAWS), "ret": [int(m), float(p)],
                                   "state_after": STORE.put_state(self.state)})
        return m, p

    def soft_measure(self, indices, measured=None, input_offset=0):
        DRAWS.clear()
        m, p = super().soft_measure(indices, measured=measured)
        self.stream["ops"].append({"op": "soft_measure", "indices": [int(i) for i in indices],
                                   "measured": measured, "draws": list(DRAWS), "ret": [int(m), float(p)]})
        return m, p

    def reduce_measure(self, indices, measured=None, measured_prob=None, input_offset=0, output_offset=0):
        DRAWS.clear()
        k = len(indices)
        m, p = super().reduce_measure(indices, measured=measured, measured_prob=measured_prob)
        # the reference keeps an N-sized arena; only the 2^(n-k) prefix is defined
        pre = np.array(self.state[: 2 ** (self.n - k)])
        self.stream["ops"].append({"op": "reduce_measure", "indices": [int(i) for i in indices],
                                   "measured": measured, "measured_prob": measured_prob,
                                   "draws": list(DRAWS), "ret": [int(m), float(p)],
                                   "state_after": STORE.put_state(pre)})
        self.stream["reduced"] = True
        return m, p

    def measure_probabilities(self, indices, top_k=0):
        ret = super().measure_probabilities(indices, top_k=top_k)
        op = {"op": "measure_probabilities", "indices": [int(i) for i in indices], "top_k": int(top_k)}
        if top_k:
            op["ret_idx"] = [int(i) for i in ret[0]]
            op["ret_p"] = [float(x) for x in ret[1]]
        else:
            op["ret"] = STORE.put(np.asarray(ret, dtype=np.float64))
        self.stream["ops"].append(op)
        return ret


def patch_reference():
    qip.pipeline.CythonBackend.make_state = Recorder.make_state   # default backend of run()


# ---------------------------------------------------------------- 1. reference test-suite
def record_reference_tests():
    sys.path.insert(0, os.path.join(os.environ.get("QIP_REFERENCE", "/root/reference"), "tests"))
    total = 0
    for modname in ("qiptest", "qubit_util_test", "qfttest"):
        suite = unittest.defaultTestLoader.loadTestsFromName(modname)

        def walk(s):
            for t in s:
                if isinstance(t, unittest.TestSuite):
                    yield from walk(t)
                else:
                    yield t
        for t in walk(suite):
            CURRENT_LABEL[0] = "reftest/%s::%s" % (modname, t._testMethodName)
            random.seed(len(t._testMethodName))
            res = unittest.TestResult()
            t.run(res)
            assert res.wasSuccessful(), (CURRENT_LABEL[0], res.failures, res.errors)
            total += 1
    return total


# ---------------------------------------------------------------- 2. BASELINE configs, oracle-sized
class Mat2Op(MatrixOp):
    """User MatrixOp subclass returning one dense 4x4 on (i, j) -- SURVEY 8d config 4."""

    def __init__(self, u, *inputs, **kw):
        super().__init__(*inputs, **kw)
        self.u = u

    def makemats(self, index_groups):
        from qip.util import flatten
        return {tuple(flatten(index_groups)): self.u}


def haar(rng, d):
    z = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    q, r = np.linalg.qr(z)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def layered_circuit(n, depth, seed):
    """The generator of SURVEY 8d config 4 (kept in sync with qip_b200/circuits.py)."""
    rng = np.random.default_rng(seed)
    qs = [Qubit(n=1) for _ in range(n)]
    for _ in range(depth):
        for i in range(n):
            if rng.random() < 0.5:
                qs[i] = H(qs[i])
            else:
                qs[i] = Rm(int(rng.integers(1, 9)), qs[i])
        perm = rng.permutation(n)
        for a, b in zip(perm[0::2], perm[1::2]):
            a, b = int(a), int(b)
            kind = int(rng.integers(0, 3))
            if kind == 0:
                qs[a], qs[b] = C(X)(qs[a], qs[b])
            elif kind == 1:
                qs[a], qs[b] = Swap(qs[a], qs[b])
            else:
                u = haar(rng, 4)
                qs[a], qs[b] = QubitOpWrapper(Mat2Op, u)(qs[a], qs[b])
    return qs


def record_configs():
    # config 1: README CSwap (README.md:8-39 == tests/qiptest.py:194-228), 11 qubits, + Measure, seeds 0..31
    for seed in range(32):
        CURRENT_LABEL[0] = "config/cswap11_measure_seed%d" % seed
        random.seed(seed)
        q1, q2, q3 = Qubit(n=1), Qubit(n=5), Qubit(n=5)
        h1 = H(q1)
        c1, c2, c3 = C(Swap)(h1, q2, q3)
        m1 = Measure(H(c1))
        s2 = np.zeros(32)
        s3 = np.zeros(32)
        s2[0] = 1.0
        s3[1] = 1.0
        run(m1, c2, c3, feed={q1: [1.0, 0.0], q2: s2, q3: s3})
    # examples/cswap_measure.py:5-23 (7 qubits, cos/sin feeds)
    CURRENT_LABEL[0] = "config/cswap7_example"
    q1, q2, q3 = Qubit(n=1), Qubit(n=3), Qubit(n=3)
    c1, c2, c3 = C(Swap)(H(q1), q2, q3)
    m1 = H(c1)
    st2 = np.cos(np.arange(0, 8) * np.pi / 8.0)
    st3 = np.sin(np.arange(0, 8) * np.pi / 8.0)
    run(m1, c2, c3, feed={q1: [1.0, 0.0], q2: st2 / np.linalg.norm(st2), q3: st3 / np.linalg.norm(st3)})
    # config 2: QFFT, random normalised feed
    for n in (8, 10):
        CURRENT_LABEL[0] = "config/qfft%d" % n
        rng = np.random.default_rng(n)
        psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
        psi /= np.linalg.norm(psi)
        q = Qubit(n=n)
        run(QFFT(q), feed={q: psi})
    # config 3: Grover, examples/grovers_iterative.py:20-39 structure, tuple-key re-feed
    n, x0 = 6, 42
    q = Qubit(n=n, default=np.ones(2 ** n) / np.sqrt(2 ** n))
    anc = Qubit(n=1, default=[1 / np.sqrt(2), -1 / np.sqrt(2)])
    os_, oa = F(lambda x: int(x == x0), q, anc)
    fs, fa = F(lambda x: int(x == 0), H(os_), oa)
    ds = H(fs)
    sm_ds, sm_da = StochasticMeasure(ds), StochasticMeasure(fa)
    CURRENT_LABEL[0] = "config/grover6_iter0"
    state, c = run(sm_ds, sm_da)
    for it in range(1, 4):
        CURRENT_LABEL[0] = "config/grover6_iter%d" % it
        state, c = run(sm_ds, sm_da, feed={(q, anc): np.array(state)})
    # config 4: random layered circuit
    for n, depth, seed in ((8, 4, 33), (10, 3, 34)):
        CURRENT_LABEL[0] = "config/layered_n%d_d%d_s%d" % (n, depth, seed)
        qs = layered_circuit(n, depth, seed)
        run(*qs)


# ---------------------------------------------------------------- 3. paths no reference test pins
def record_unpinned():
    psi3 = np.zeros(8, dtype=np.complex128)
    psi3[0b011] = 0.6
    psi3[0b110] = 0.8
    rng = np.random.default_rng(7)
    psi5 = rng.normal(size=32) + 1j * rng.normal(size=32)
    psi5 /= np.linalg.norm(psi5)

    def fresh(label, n, psi):
        CURRENT_LABEL[0] = label
        return Recorder.make_state(n, [list(range(n))], [psi])

    b = fresh("unpinned/probabilities_bit_order", 3, psi3)
    b.measure_probabilities(np.array([1, 2], dtype=np.int32))
    b.measure_probabilities(np.array([2, 1], dtype=np.int32))
    b.measure_probabilities(np.array([1, 2], dtype=np.int32), top_k=4)
    b.measure_probabilities(np.array([2, 1], dtype=np.int32), top_k=4)
    b.total_prob()
    b.get_state()

    b = fresh("unpinned/probabilities_random5", 5, psi5)
    for idx in ([0], [4], [3, 1], [1, 3], [4, 0, 2], [0, 1, 2, 3, 4], [4, 3, 2, 1, 0]):
        b.measure_probabilities(np.array(idx, dtype=np.int32))
    b.measure_probabilities(np.array([3, 0, 4], dtype=np.int32), top_k=8)
    b.total_prob()
    b.get_state()

    for seed in range(12):
        b = fresh("unpinned/soft_measure_seed%d" % seed, 5, psi5)
        random.seed(seed)
        b.soft_measure(np.array([3, 1], dtype=np.int32))
        b.soft_measure(np.array([1, 3], dtype=np.int32))
        b.soft_measure(np.array([0, 2, 4], dtype=np.int32))
        b.soft_measure(np.array([3, 1], dtype=np.int32), measured=2)
        b.get_state()

    for seed in range(6):
        b = fresh("unpinned/measure_seed%d" % seed, 5, psi5)
        random.seed(100 + seed)
        b.measure(np.array([2], dtype=np.int32))
        b.measure(np.array([4, 0], dtype=np.int32))
        b.get_state()

    b = fresh("unpinned/measure_given_outcome", 3, psi3)
    b.measure(np.array([0], dtype=np.int32), measured=1, measured_prob=0.64)
    b.get_state()
    b = fresh("unpinned/measure_given_outcome_only", 3, psi3)
    random.seed(5)
    b.measure(np.array([0, 2], dtype=np.int32), measured=0b01)
    b.get_state()

    for seed in range(4):
        b = fresh("unpinned/reduce_measure_seed%d" % seed, 5, psi5)
        random.seed(200 + seed)
        b.reduce_measure(np.array([1, 3], dtype=np.int32))


def main():
    patch_reference()
    ntests = record_reference_tests()
    record_configs()
    record_unpinned()
    random.random = _real_random
    meta = {"generator": "tests/golden/make_golden.py", "reference": "Renmusxd/QIP 0.5.1 (unmodified, Cython kernels via oracle/_ref)",
            "n_reference_tests": ntests, "n_streams": len(STORE.streams)}
    with open(os.path.join(HERE, "ref_streams.json"), "w") as f:
        json.dump({"meta": meta, "streams": STORE.streams}, f)
    np.savez_compressed(os.path.join(HERE, "ref_streams.npz"), **STORE.arrays)
    print(meta)


if __name__ == "__main__":
    main()
