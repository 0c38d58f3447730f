"""Replay of golden op streams (tests/golden/ref_streams.*, recorded from the unmodified reference
by tests/golden/make_golden.py) through any backend with the reference's StateType surface."""
import json
import os
import random

import numpy as np

from qip_b200.mats import CMat, SwapMat

HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def load_streams():
    if "s" not in _cache:
        with open(os.path.join(HERE, "golden", "ref_streams.json")) as f:
            doc = json.load(f)
        arrays = np.load(os.path.join(HERE, "golden", "ref_streams.npz"))
        _cache["s"] = (doc["meta"], doc["streams"], arrays)
    return _cache["s"]


def _arr(arrays, ref):
    return arrays[ref[1:]]


def dec_mat(arrays, m):
    if m["type"] == "C":
        return CMat(dec_mat(arrays, m["m"]))
    if m["type"] == "swap":
        return SwapMat(m["n"])
    return np.array(_arr(arrays, m["val"]))


def dec_mats(arrays, mats):
    out = {}
    for e in mats:
        key = int(e["key"]) if e["is_int_key"] else tuple(e["key"])
        out[key] = dec_mat(arrays, e["mat"])
    return out


def dec_state(arrays, s):
    if s["sparse"]:
        a = np.zeros(s["len"], dtype=np.complex128)
        a[_arr(arrays, s["idx"])] = _arr(arrays, s["val"])
        return a
    return np.array(_arr(arrays, s["val"]))


def close(a, b, tol):
    a = np.asarray(a)
    b = np.asarray(b)
    scale = max(1.0, float(np.max(np.abs(b))) if b.size else 1.0)
    return a.shape == b.shape and float(np.max(np.abs(a - b))) <= tol * scale if b.size else a.shape == b.shape


class _Draws(object):
    """Feeds the recorded uniform draws to random.random() during a replayed measurement."""

    def __init__(self, draws):
        self.draws = list(draws)
        self.saved = None

    def __enter__(self):
        self.saved = random.random
        it = iter(self.draws)
        random.random = lambda: next(it)
        return self

    def __exit__(self, *a):
        random.random = self.saved


def replay(stream, arrays, make_state, tol=1e-12, statetype=np.complex128, check_states=True):
    """Run one recorded stream; assert every recorded return value and state."""
    label = stream["label"]
    feeds = [np.array(_arr(arrays, f)) for f in stream["feeds"]]
    b = make_state(stream["n"], stream["index_groups"], feeds, statetype=statetype)
    for op in stream["ops"]:
        kind = op["op"]
        if kind == "kronselect_dot":
            b.kronselect_dot(dec_mats(arrays, op["mats"]))
        elif kind == "func_apply":
            table = _arr(arrays, op["table"])
            b.func_apply(np.array(op["reg1"], dtype=np.int32), np.array(op["reg2"], dtype=np.int32),
                         lambda x, t=table: int(t[x]), stream["n"])
        elif kind == "total_prob":
            assert abs(b.total_prob() - op["ret"]) <= max(tol, 1e-15) * max(1.0, op["ret"]), label
        elif kind in ("measure", "reduce_measure"):
            with _Draws(op["draws"]):
                m, p = getattr(b, kind)(np.array(op["indices"], dtype=np.int32), measured=op["measured"],
                                        measured_prob=op["measured_prob"])
            assert m == op["ret"][0], (label, kind, m, op["ret"])
            assert abs(p - op["ret"][1]) <= max(tol, 1e-15) * 4, (label, kind, p, op["ret"])
            if check_states:
                got = np.asarray(b.get_state())
                want = dec_state(arrays, op["state_after"])
                assert close(got[: len(want)], want, tol), (label, kind, "state_after")
        elif kind == "soft_measure":
            with _Draws(op["draws"]):
                m, p = b.soft_measure(np.array(op["indices"], dtype=np.int32), measured=op["measured"])
            assert m == op["ret"][0], (label, kind, m, op["ret"])
            assert abs(p - op["ret"][1]) <= max(tol, 1e-15) * 4, (label, kind, p, op["ret"])
        elif kind == "measure_probabilities":
            ret = b.measure_probabilities(np.array(op["indices"], dtype=np.int32), top_k=op["top_k"])
            if op["top_k"]:
                idx, ps = ret
                assert close(ps, op["ret_p"], max(tol, 1e-15) * 4), (label, kind, ps, op["ret_p"])
                want_p = op["ret_p"]
                for j, (i_got, i_want) in enumerate(zip(idx, op["ret_idx"])):
                    tied = any(abs(want_p[j] - want_p[t]) <= 1e-9 for t in range(len(want_p)) if t != j)
                    if not tied:
                        assert i_got == i_want, (label, kind, idx, op["ret_idx"])
            else:
                assert close(ret[:], _arr(arrays, op["ret"]), max(tol, 1e-15) * 4), (label, kind)
        else:
            raise AssertionError("unknown op " + kind)
    if "final_state" in stream and not stream.get("reduced"):
        got = np.asarray(b.get_state())
        want = dec_state(arrays, stream["final_state"])
        assert close(got, want, tol), (label, "final_state", float(np.max(np.abs(got - want))))
    b.close()
    return True
