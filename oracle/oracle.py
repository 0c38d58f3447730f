"""ctypes front for oracle/qip_oracle.c + two CPU backends with the reference's StateType surface.

TEST INFRASTRUCTURE ONLY.  Nothing under qip_b200/ imports this module.  Allowed users: tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

  OracleBackend  -- the C restatement (kind "port"): 64-bit indices, no -ffast-math.
  RefBackend     -- the reference's own compiled Cython kernels from oracle/_ref (kind
                    "reference"), driven by a restatement of the ~20 lines of python glue in
                    qip/util.py:15-67 and qip/backend.py:73-175 (the .py files do not travel to
                    the GPU box; the compiled kernels do).

Both accept the `mats` vocabulary of the boundary: keys int | tuple[int],
values ndarray/list, or any object carrying `_kron_struct` == 2 (`.m`, controlled) / == 3 (`.n`,
swap) -- i.e. the reference's qip.operators.CMat / SwapMat or qip_b200's own carriers.
"""
import ctypes
import os
import random
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libqip_oracle.so")
SRC_PATH = os.path.join(HERE, "qip_oracle.c")

ORC_DENSE, ORC_CTRL, ORC_SWAP = 1, 2, 3


class _OrcMat(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("dim", ctypes.c_int32), ("child", ctypes.c_int32),
                ("pad", ctypes.c_int32), ("data_off", ctypes.c_int64)]


def build(force: bool = False) -> str:
    """Compile the C restatement with the system gcc (no -ffast-math)."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(SRC_PATH):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-O2", "-fopenmp", "-shared", "-fPIC", "-o", LIB_PATH, SRC_PATH, "-lm"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        i32p = ctypes.POINTER(ctypes.c_int32)
        vp = ctypes.c_void_p
        i64 = ctypes.c_int64
        L.orc_cdot.argtypes = [ctypes.c_int, ctypes.c_int, i32p, i32p, ctypes.POINTER(_OrcMat), vp,
                               vp, i64, i64, vp, i64, i64]
        L.orc_cdot.restype = None
        L.orc_prob_magnitude.argtypes = [vp, i64]
        L.orc_prob_magnitude.restype = ctypes.c_double
        L.orc_measure_probabilities.argtypes = [ctypes.c_int, ctypes.c_int, i32p, vp, vp]
        L.orc_measure_probabilities.restype = None
        L.orc_outcome_probabilities.argtypes = [ctypes.c_int, ctypes.c_int, i32p, vp, i64, i64, vp]
        L.orc_outcome_probabilities.restype = None
        L.orc_soft_measure.argtypes = [ctypes.c_int, ctypes.c_int, i32p, vp, i64, i64, i64,
                                       ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
        L.orc_soft_measure.restype = i64
        L.orc_collapse.argtypes = [ctypes.c_int, ctypes.c_int, i32p, i64, ctypes.c_double, vp, vp, i64]
        L.orc_collapse.restype = None
        L.orc_reduce.argtypes = [ctypes.c_int, ctypes.c_int, i32p, i64, ctypes.c_double, vp, vp]
        L.orc_reduce.restype = None
        L.orc_func_apply.argtypes = [ctypes.c_int, ctypes.c_int, i32p, ctypes.c_int, i32p, vp, vp, vp]
        L.orc_func_apply.restype = None
        L.orc_make_state.argtypes = [ctypes.c_int, ctypes.c_int, i32p, i32p, vp, vp]
        L.orc_make_state.restype = None
        _lib = L
    return _lib


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def normalise_mats(mats):
    """Key/shape validation of qip/util.py:35-58 (same exception types and order)."""
    out = {}
    for key in mats:
        if type(key) != tuple and type(key) != int:
            raise Exception("Type of indices must be tuple: {}".format(key))
        m = mats[key]
        if type(m) == list:
            m = np.array(m)
        tkey = key if type(key) == tuple else (key,)
        if 2 ** len(tkey) != m.shape[0] or 2 ** len(tkey) != m.shape[1]:
            raise Exception("Shape of square submatrix must equal 2**(number of indices): "
                            "{}: {}".format(key, m))
        out[tkey] = m
    return out


def pack_mats(mats):
    """Flatten a normalised mats dict into (group_len, group_q, orc_mat[], pool)."""
    group_len, group_q, pool = [], [], []
    entries = []           # top-level carriers first (index g == group g), nested ones appended
    pending = []
    for key, m in mats.items():
        group_len.append(len(key))
        group_q.extend(key)
        entries.append(None)
        pending.append((len(entries) - 1, m))
    pool_len = 0
    while pending:
        slot, m = pending.pop(0)
        ks = getattr(m, "_kron_struct", None)
        if ks == 2:
            entries.append(None)
            child = len(entries) - 1
            entries[slot] = (ORC_CTRL, int(m.shape[0]), child, 0)
            pending.append((child, m.m))
        elif ks == 3:
            entries[slot] = (ORC_SWAP, int(m.n), 0, 0)
        elif isinstance(m, (np.ndarray, list)):
            a = np.ascontiguousarray(np.asarray(m), dtype=np.complex128)
            entries[slot] = (ORC_DENSE, int(a.shape[0]), 0, pool_len)
            pool.append(a.reshape(-1))
            pool_len += a.size
        else:
            raise ValueError("Cannot pass matrices which are not numpy, SwapMat, or CMat")
    arr = (_OrcMat * len(entries))()
    for i, (kind, dim, child, off) in enumerate(entries):
        arr[i].kind, arr[i].dim, arr[i].child, arr[i].data_off = kind, dim, child, off
    poolarr = np.concatenate(pool) if pool else np.zeros(1, dtype=np.complex128)
    return group_len, group_q, arr, poolarr


def cdot(mats, vec, n, out, input_offset=0, output_offset=0):
    """qip.util.kronselect_dot(..., dot_impl=cdot_loop) restated over the C oracle."""
    if len(vec) + input_offset > 2 ** n:
        raise ValueError("Input vector size plus offset may be no larger than the total number of qubit states (2^n)")
    if len(out) + output_offset > 2 ** n:
        raise ValueError("Output vector size plus offset may be no larger than the total number of qubit states (2^n)")
    nm = normalise_mats(mats)
    gl, gq, arr, pool = pack_mats(nm)
    gl, glp = _i32(gl)
    gq, gqp = _i32(gq)
    lib().orc_cdot(n, len(gl), glp, gqp, arr, pool.ctypes.data, vec.ctypes.data, len(vec), input_offset,
                   out.ctypes.data, len(out), output_offset)


def _check_measure_args(k, measured, measured_prob):
    # qip/ext/kronprod.pyx:402-407 / 453-458
    if measured is not None and not (0 <= measured < 2 ** k):
        raise ValueError("Measured value must be less than 2**len(indices)")
    if measured_prob is not None and not (0.0 < measured_prob <= 1.0):
        raise ValueError("measured_prob must be 0 < p <= 1")


def top_probabilities(probs_big_endian, top_k):
    """Intended semantics of measure_top_probabilities (qip/ext/kronprod.pyx:266-315): the top_k
    outcomes by probability, descending; ties broken by ascending outcome (the reference's ad-hoc
    heap reads one slot past its arrays once full and is only exact on its own test case; see
    DESIGN.md 'quirks')."""
    k = min(int(top_k), len(probs_big_endian))
    order = np.argsort(-np.asarray(probs_big_endian), kind="stable")[:k]
    return [int(i) for i in order], [float(probs_big_endian[i]) for i in order]


class OracleBackend(object):
    """The C restatement behind the reference's StateType method surface (qip/backend.py:14-65)."""
    kind = "port"

    def __init__(self, n, state):
        self.n = n
        self.state = state
        self.arena = np.empty_like(state)

    @staticmethod
    def make_state(n, index_groups, feed_list, statetype=np.complex128, **_):
        state = np.zeros(2 ** n, dtype=np.complex128)
        gl = [len(g) for g in index_groups]
        gq = [q for g in index_groups for q in g]
        feeds = [np.asarray(f, dtype=np.complex128).reshape(-1) for f in feed_list]
        for g, f in zip(index_groups, feeds):
            if len(f) != 2 ** len(g):
                raise ValueError("feed length must be 2**len(group)")
        fcat = np.concatenate(feeds) if feeds else np.zeros(1, dtype=np.complex128)
        gla, glp = _i32(gl if gl else [0])
        gqa, gqp = _i32(gq if gq else [0])
        lib().orc_make_state(n, len(gl), glp, gqp, fcat.ctypes.data, state.ctypes.data)
        return OracleBackend(n, state)

    def _swap(self):
        self.state, self.arena = self.arena, self.state

    def get_state(self):
        return self.state

    def kronselect_dot(self, mats, input_offset=0, output_offset=0):
        cdot(mats, self.state, self.n, self.arena, input_offset, output_offset)
        self._swap()

    def func_apply(self, reg1_indices, reg2_indices, func, input_offset=0, output_offset=0):
        r1, r1p = _i32(reg1_indices)
        r2, r2p = _i32(reg2_indices)
        ftab = np.array([int(func(x)) for x in range(2 ** len(r1))], dtype=np.int64)
        lib().orc_func_apply(self.n, len(r1), r1p, len(r2), r2p, ftab.ctypes.data,
                             self.state.ctypes.data, self.arena.ctypes.data)
        self._swap()

    def total_prob(self):
        return lib().orc_prob_magnitude(self.state.ctypes.data, len(self.state))

    def soft_measure(self, indices, measured=None, input_offset=0, r=None):
        q, qp = _i32(indices)
        if r is None:
            r = random.random()      # exactly one draw, like qip/ext/kronprod.pyx:362
        p = ctypes.c_double(0.0)
        m = lib().orc_soft_measure(self.n, len(q), qp, self.state.ctypes.data, len(self.state), input_offset,
                                   -1 if measured is None else int(measured), float(r), ctypes.byref(p))
        return int(m), p.value

    def measure(self, indices, measured=None, measured_prob=None, input_offset=0, output_offset=0, r=None):
        q, qp = _i32(indices)
        _check_measure_args(len(q), measured, measured_prob)
        if measured is None or measured_prob is None:
            m, p = self.soft_measure(indices, measured=measured, r=r)
        else:
            m, p = measured, measured_prob
        lib().orc_collapse(self.n, len(q), qp, m, p, self.state.ctypes.data, self.arena.ctypes.data,
                           len(self.state))
        self._swap()
        return m, p

    def reduce_measure(self, indices, measured=None, measured_prob=None, input_offset=0, output_offset=0, r=None):
        q, qp = _i32(indices)
        _check_measure_args(len(q), measured, measured_prob)
        if measured is None or measured_prob is None:
            m, p = self.soft_measure(indices, measured=measured, r=r)
        else:
            m, p = measured, measured_prob
        out = np.empty(2 ** (self.n - len(q)), dtype=np.complex128)
        lib().orc_reduce(self.n, len(q), qp, m, p, self.state.ctypes.data, out.ctypes.data)
        self.n -= len(q)
        self.state, self.arena = out, np.empty_like(out)
        return m, p

    def measure_probabilities(self, indices, top_k=0):
        q, qp = _i32(indices)
        probs = np.zeros(2 ** len(q), dtype=np.float64)
        if top_k:
            lib().orc_outcome_probabilities(self.n, len(q), qp, self.state.ctypes.data, len(self.state), 0,
                                            probs.ctypes.data)
            return top_probabilities(probs, top_k)
        lib().orc_measure_probabilities(self.n, len(q), qp, self.state.ctypes.data, probs.ctypes.data)
        return probs

    def get_state_size(self):
        return len(self.state)

    def get_relative_range(self, start, end):
        return self.state[start:end]

    def overwrite_relative_range(self, start, end, data):
        self.state[start:end] = data

    def addto_relative_range(self, start, end, data):
        self.state[start:end] += data

    def close(self):
        pass


class RefBackend(object):
    """The reference's compiled kernels (oracle/_ref) behind the same surface; complex128, n <= 30.

    Glue restated from qip/util.py:60-65 (index/matrix packing) and qip/backend.py:106-175
    (state/arena ping-pong).  make_state is restated with numpy instead of the reference's python
    loop over 2^(#fed qubits) entries (qip/backend.py:94-101) -- same values, seconds not minutes."""
    kind = "reference"

    def __init__(self, n, state):
        from oracle.ref_loader import load_ref_ext
        self._kp, self._fa = load_ref_ext()
        self.n = n
        self.state = state
        self.arena = np.empty_like(state)

    @staticmethod
    def make_state(n, index_groups, feed_list, statetype=np.complex128, **_):
        ob = OracleBackend.make_state(n, index_groups, feed_list)
        return RefBackend(n, ob.state)

    def _swap(self):
        self.state, self.arena = self.arena, self.state

    def get_state(self):
        return self.state

    def kronselect_dot(self, mats, input_offset=0, output_offset=0):
        nm = normalise_mats(mats)
        keys = list(nm.keys())
        cindices = np.array([np.array(k, dtype=np.int32) for k in keys])
        vals = [nm[k].astype(np.complex128) if type(nm[k]) == np.ndarray else nm[k] for k in keys]
        if all(type(v) == np.ndarray for v in vals):
            cmats = np.array(vals)
        else:
            cmats = np.empty(len(vals), dtype=object)
            for i, v in enumerate(vals):
                cmats[i] = v
        self._kp.cdot_loop(cindices, cmats, self.state, self.n, self.arena,
                           input_offset=input_offset, output_offset=output_offset)
        self._swap()

    def func_apply(self, reg1_indices, reg2_indices, func, input_offset=0, output_offset=0):
        self._fa.func_apply(np.asarray(reg1_indices, dtype=np.int32), np.asarray(reg2_indices, dtype=np.int32),
                            func, self.state, self.n, self.arena)
        self._swap()

    def total_prob(self):
        return float(np.sum(np.abs(self.state) ** 2))   # the compiled prob_magnitude has UB (8g-5)

    def soft_measure(self, indices, measured=None, input_offset=0):
        return self._kp.soft_measure(np.asarray(indices, dtype=np.int32), self.n, self.state,
                                     measured=measured, input_offset=input_offset)

    def measure(self, indices, measured=None, measured_prob=None, input_offset=0, output_offset=0):
        m, p = self._kp.measure(np.asarray(indices, dtype=np.int32), self.n, self.state, self.arena,
                                measured=measured, measured_prob=measured_prob,
                                input_offset=input_offset, output_offset=output_offset)
        self._swap()
        return m, p

    def measure_probabilities(self, indices, top_k=0):
        idx = np.asarray(indices, dtype=np.int32)
        if top_k:
            return self._kp.measure_top_probabilities(idx, self.n, top_k, self.state)
        return self._kp.measure_probabilities(idx, self.n, self.state)

    def get_state_size(self):
        return len(self.state)

    def close(self):
        pass
