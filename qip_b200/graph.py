"""Graph-level op-stream compiler: walk a QIP graph ONCE, replay it on the B200.

The reference executes a graph node by node (qip/pipeline.py:186-222 run_graph re-sorts the frontier
for every node, NodeFeeder.feed :235-240 calls into the backend per node) and rebuilds everything on
every `run()`.  Here the UNMODIFIED reference front-end is run once against a recording StateType
(through the reference's own hook, `backend_constructor=`, qip/pipeline.py:95-96,133), which captures
the whole boundary traffic of the circuit: the feed layout, every `kronselect_dot` / `func_apply` /
`measure` / `measure_probabilities` call and which graph node each classical result belongs to.

    circ = compile_circuit(out1, out2, feed={q: psi})          # traces once, no device work
    state, classic = circ.run(feed={q: psi2})                   # replays on B200Backend
    state, classic = circ.run(feed={q: state})                  # device-resident re-feed (Grover loop)

What the replay saves over `run(..., backend_constructor=B200Backend.make_state)`:
  * the graph walk and the per-node python of the front-end;
  * decoding / validation / simplification of every `mats` dict (done at trace time);
  * from the second replay on, merging and planning too: each gate segment between two observation
    points remembers its planned fused passes and their packed C structs (B200Backend.apply_gates);
  * function tables of `F(...)` nodes (tabulated once);
  * host round trips of the state: feeds may be DeviceState handles / torch tensors, one-hot `int`
    feeds and int `Qubit.default`s never become 2^k host vectors, and `device_state=True` returns the
    handle instead of a host copy.
Results are identical to the node-by-node run: the same gates reach the same kernels in the same order.
"""
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .ops import Gate, decode_mats, simplify


class _Slot(object):
    """Placeholder for a classical result produced during the trace (filled in at replay)."""

    def __init__(self, index: int, part: Optional[int] = None):
        self.index, self.part = index, part

    def __getitem__(self, item):              # StochasticMeasure slices the result with [:] (qip/qip.py:181)
        return self

    def __iter__(self):                       # Measure unpacks `bits, prob = state.measure(...)` (qip/qip.py:152)
        return iter((_Slot(self.index, 0), _Slot(self.index, 1)))


class _Recorder(object):
    """StateType (qip/backend.py:14-65) that records instead of computing."""

    def __init__(self, n, index_groups, feed_list, statetype):
        self.n = int(n)
        self.groups = [[int(q) for q in g] for g in index_groups]
        self.feeds = list(feed_list)
        self.statetype = statetype
        self.ops: List[tuple] = []            # ("gates", [Gate]) | ("func", reg1, reg2, func) | ("measure", idx, slot) | ("probs", idx, top_k, slot)
        self.nslots = 0

    def _slot(self):
        self.nslots += 1
        return self.nslots - 1

    def kronselect_dot(self, mats, input_offset=0, output_offset=0):
        if input_offset != 0 or output_offset != 0:
            raise ValueError("offset windows are not supported")
        gates = [s for s in (simplify(g) for g in decode_mats(mats, self.n)) if s is not None]
        if self.ops and self.ops[-1][0] == "gates":
            self.ops[-1][1].extend(gates)
        else:
            self.ops.append(("gates", gates))

    def func_apply(self, reg1_indices, reg2_indices, func, input_offset=0, output_offset=0):
        self.ops.append(("func", [int(i) for i in reg1_indices], [int(i) for i in reg2_indices], func))

    def measure(self, indices, measured=None, measured_prob=None, input_offset=0, output_offset=0):
        if measured is not None or measured_prob is not None:
            raise ValueError("compiled circuits do not support pre-selected measurement outcomes")
        s = self._slot()
        self.ops.append(("measure", [int(i) for i in indices], s))
        return _Slot(s)

    def measure_probabilities(self, indices, top_k=0):
        s = self._slot()
        self.ops.append(("probs", [int(i) for i in indices], int(top_k), s))
        return _Slot(s)

    def get_state(self):
        return None

    def close(self):
        pass


def _fill(template, results):
    if isinstance(template, _Slot):
        r = results[template.index]
        return r if template.part is None else r[template.part]
    if isinstance(template, tuple):
        return tuple(_fill(t, results) for t in template)
    if isinstance(template, list):
        return [_fill(t, results) for t in template]
    return template


class CompiledCircuit(object):
    """The recorded boundary traffic of one graph; `run(feed=...)` replays it on a B200 backend."""

    def __init__(self, n, statetype, feed_keys, groups, feeds, ops, classic_template):
        self.n, self.statetype = n, statetype
        self.feed_keys = feed_keys            # tuple-of-qubits keys in feed_list order (defaults appended by the front-end)
        self.groups, self.default_feeds = groups, feeds
        self.ops, self.classic_template = ops, classic_template
        self.ngates = sum(len(o[1]) for o in ops if o[0] == "gates")
        self._plans: Dict[Any, Any] = {}      # (backend kind, segment index) -> planned passes
        self._tables: Dict[int, np.ndarray] = {}

    @classmethod
    def from_ops(cls, n, index_groups, feed_list, ops, statetype=np.complex128) -> "CompiledCircuit":
        """Compile a plain boundary-level op list (no front-end needed): items ("k", mats),
        ("f", reg1, reg2, func), ("m", indices), ("p", indices, top_k).  Feed keys are the group
        positions 0, 1, ...; the classic map of `run` is keyed by the op's position in `ops`."""
        rec = _Recorder(n, index_groups, feed_list, statetype)
        template = {}
        for i, op in enumerate(ops):
            if op[0] == "k":
                rec.kronselect_dot(op[1])
            elif op[0] == "f":
                rec.func_apply(op[1], op[2], op[3])
            elif op[0] == "m":
                template[i] = tuple(rec.measure(op[1]))
            elif op[0] == "p":
                template[i] = rec.measure_probabilities(op[1], top_k=op[2] if len(op) > 2 else 0)
            else:
                raise ValueError("unknown op kind {!r}".format(op[0]))
        keys = [(j,) for j in range(len(rec.groups))]
        return cls(rec.n, statetype, keys, rec.groups, rec.feeds, rec.ops, template)

    def _feed_list(self, feed):
        feed = {} if feed is None else {(k if type(k) == tuple else (k,)): v for k, v in feed.items()}
        unknown = [k for k in feed if k not in self.feed_keys]
        if unknown:
            raise ValueError("feed keys {} were not part of the compiled circuit's feed".format(unknown))
        return [feed.get(k, d) for k, d in zip(self.feed_keys, self.default_feeds)]

    def run(self, feed=None, backend_constructor=None, device_state=False, **backend_kwargs):
        """Replay.  `feed` has the keys given at compile time (values may differ; missing keys keep the
        compile-time value).  Returns (state, classic_map) like qip/pipeline.py:71-135."""
        from .backend import B200Backend, tabulate
        make = backend_constructor or B200Backend.make_state
        if device_state and backend_constructor is None:
            backend_kwargs.setdefault("host_state_max_qubits", -1)
        b = make(self.n, self.groups, self._feed_list(feed), statetype=self.statetype, **backend_kwargs)
        kind = (type(b).__name__, getattr(b, "strategy", None), getattr(b, "fuse", None), getattr(b, "tile_bits", None))
        results: Dict[int, Any] = {}
        for i, op in enumerate(self.ops):
            if op[0] == "gates":
                b.apply_gates(op[1], self._plans, (kind, i))
            elif op[0] == "func":
                _, reg1, reg2, func = op
                if i not in self._tables:                      # tabulated once; the function object also keeps the device copy
                    self._tables[i] = _TableFunc(tabulate(func, len(reg1)))
                b.func_apply(np.array(reg1, dtype=np.int32), np.array(reg2, dtype=np.int32), self._tables[i])
            elif op[0] == "measure":
                results[op[2]] = b.measure(np.array(op[1], dtype=np.int32))
            else:
                results[op[3]] = b.measure_probabilities(np.array(op[1], dtype=np.int32), top_k=op[2])[:]
        state = b.get_state()
        self.last_stats = dict(getattr(b, "stats", {}))
        b.close()                             # a DeviceState keeps its tensor alive; nothing else of the run is retained
        return state, {node: _fill(t, results) for node, t in self.classic_template.items()}


class _TableFunc(object):
    """A function given by its table (what qip/ext/func_apply.pyx:66-69 builds on every call)."""

    def __init__(self, table):
        self.table = np.ascontiguousarray(table, dtype=np.int64)

    def __call__(self, x):
        return self.table[x]


def compile_circuit(*outputs, feed=None, statetype=np.complex128, strict=False) -> CompiledCircuit:
    """Trace `run(*outputs, feed=feed)` of the reference front-end (qip/pipeline.py:71-135) against a
    recording backend.  `feed` fixes the feed KEYS (and default values); nothing runs on the device."""
    from qip.pipeline import run as reference_run
    box = {}

    def constructor(n, index_groups, feed_list, statetype=np.complex128):
        box["rec"] = _Recorder(n, index_groups, feed_list, statetype)
        return box["rec"]

    norm_feed = None if feed is None else {(k if type(k) == tuple else (k,)): v for k, v in feed.items()}
    _, classic = reference_run(*outputs, feed=feed, strict=strict, backend_constructor=constructor, statetype=statetype)
    rec = box["rec"]
    # feed_list order = the front-end's feed dict order: the user's keys first, injected defaults after
    # (qip/pipeline.py:101-109, 124-125).  Recover the keys of the injected defaults from the groups.
    keys = list(norm_feed.keys()) if norm_feed else []
    feeds = list(rec.feeds)
    for j in range(len(keys), len(rec.groups)):
        keys.append(("default", tuple(rec.groups[j])))
        # an int `Qubit.default` reaches the backend as a one-hot host vector (qip/pipeline.py:103-106):
        # turn it back into the index so that replays never upload 2^k amplitudes for it
        v = feeds[j]
        if isinstance(v, np.ndarray) and v.ndim == 1 and np.count_nonzero(v) == 1:
            hot = int(np.flatnonzero(v)[0])
            if v[hot] == 1.0:
                feeds[j] = hot
    return CompiledCircuit(rec.n, statetype, keys, rec.groups, feeds, rec.ops, dict(classic))


def run(*outputs, feed=None, statetype=np.complex128, strict=False, backend_constructor=None, device_state=False,
        **backend_kwargs):
    """Drop-in for qip.pipeline.run on the B200: compile + one replay.  Unlike going through the
    reference's `run(..., backend_constructor=B200Backend.make_state)`, device-resident feeds stay on the
    device and nothing is materialised on the host unless asked for."""
    circ = compile_circuit(*outputs, feed=feed, statetype=statetype, strict=strict)
    return circ.run(feed=feed, backend_constructor=backend_constructor, device_state=device_state, **backend_kwargs)
