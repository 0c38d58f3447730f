"""DistributedBackend -- the front door of qip/distributed/backend.py:13-129 on the multi-GPU engine
(SURVEY 8f row 4).

User code written against the reference's distributed mode does

    from qip.distributed.backend import DistributedBackend
    out, classic = run(node, feed=..., backend_constructor=DistributedBackend.make_state)

and talks protobuf over TCP to a manager that farms 4^(n - worker_n) matrix blocks out to worker
processes (qip/distributed/manager.py:138-195).  Switching the import to `qip_b200.distributed` keeps the
call unchanged: the same `make_state(n, index_groups, feed_list, statetype)` signature
(qip/distributed/backend.py:31-33), one-hot `int` feeds (:42-45), and the distributed backend's return
conventions -- `measure_probabilities` ALWAYS answers `(indices, probabilities)` sorted by probability,
with `top_k` defaulting to all 2^k outcomes (:113-129).  Behind it sits ShardedB200Backend when the
process is one rank of a torch.distributed job (one rank per GPU, state sharded by its top qubits, NVLink
peer exchanges), or B200Backend on a single GPU.  There is no manager, no worker pool and no wire format:
inside one NVSwitch box they have nothing to do.  Unlike the reference's distributed backend,
`func_apply`, `soft_measure` and `get_state` work here (the reference raises NotImplemented, :74-76,
110-112).
"""
import numpy as np


def _engine_factory():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        from .sharded import ShardedB200Backend
        return ShardedB200Backend
    from .backend import B200Backend
    return B200Backend


class DistributedBackend(object):
    """StateType (qip/backend.py:14-65) with the distributed backend's conventions; every other method is
    the engine's."""

    def __init__(self, engine):
        self.engine = engine
        self.n = engine.n

    @staticmethod
    def make_state(n, index_groups, feed_list, statetype=np.complex128, **kwargs) -> "DistributedBackend":
        return DistributedBackend(_engine_factory().make_state(n, index_groups, feed_list, statetype=statetype, **kwargs))

    def measure_probabilities(self, indices, top_k: int = 0):
        """qip/distributed/backend.py:113-129: always the top-k form; top_k = 0 means every outcome."""
        if not top_k:
            top_k = pow(2, len(indices))
        return self.engine.measure_probabilities(indices, top_k=top_k)

    def __getattr__(self, name):
        return getattr(self.engine, name)
