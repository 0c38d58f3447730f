"""ShardedB200Backend -- the StateType surface over P = 2^G B200s, one process per GPU.

Replaces qip/distributed (DistributedBackend + manager + workers, qip/distributed/backend.py:31-129,
manager.py, worker/worker.py) for a single NVSwitch box: the state is sharded by its top G index
bits (rank r owns the contiguous range whose top bits are r, the slicing of manager.py:162-170),
gates are scheduled by qip_b200.shardplan, local work runs through the same fused sm_100a kernels
as the single-GPU backend, and the only data motion is
  * qipb_peer_swap_bit  -- global<->local qubit swap, in place over NVLink peer memory (CUDA IPC),
  * qipb_peer_gate1     -- fused compute+exchange for a 1-qubit gate on a rank bit,
  * a float64 all-reduce (NCCL, via torch.distributed) of probability histograms.
Every rank runs the same python program (SPMD) and must make the same calls in the same order.
torch.distributed supplies rendezvous, handle exchange, the stream-ordered barriers around peer
kernels and the histogram all-reduce; it never carries amplitudes.
"""
import ctypes
import os
import math
import random
from typing import List, Optional

import numpy as np

from . import lib as _lib
from . import shardplan as sp
from .backend import (B200Backend, _check_measure_args, scan_outcome, tabulate, top_probabilities)
from .ops import BitGate, Gate, Pass, decode_mats, merge_bitgates, plan_passes, simplify


def _torch():
    import torch
    return torch


class _RawCuda(object):
    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False),
                                         "version": 3, "strides": None}


# Shard allocations are pooled per process: cudaMalloc of a 128 GiB shard, its IPC export and the P-1 peer
# mappings cost ~0.3 s per run() at 8 GPUs.  Every rank runs the same program (SPMD), so the pools of all
# ranks hold the same sizes at the same time and a re-used shard keeps its peer mappings valid.  A request of
# a different size first releases what is pooled (collectively).
_SHARD_POOL = {}        # (device index, bytes, world size) -> (ptr, peers)
# Shards of states that were dropped without close() while the pool slot of their size was taken (the reference's
# run() never closes a state, qip/pipeline.py:248): freed collectively at the next construction / drain.
_ORPHANS = []           # [(ptr, peers)]

_HOST_STATE_MAX_QUBITS = 30     # get_state(): the whole state as a host ndarray on every rank up to here, a handle beyond


class ShardedState(object):
    """Array-like handle on a sharded state in canonical index order (returned by get_state() beyond 30 qubits, where
    the whole vector does not belong in host memory): len(), .shape, slicing / indexing (every rank receives the
    requested range; D2H of that range only), `local_shard` (this rank's device tensor).  Like every call of the sharded
    engine, reads are collective: all ranks must make them in the same order.  The handle keeps the engine alive."""

    def __init__(self, backend):
        self.backend = backend
        self.shape = (2 ** backend.n,)
        self.dtype = backend.np_dtype

    def __len__(self):
        return self.shape[0]

    @property
    def local_shard(self):
        return self.backend.eng.state

    @property
    def local_range(self):
        b = self.backend
        return (b.rank << b.nl, (b.rank + 1) << b.nl)

    def __getitem__(self, item):
        size = self.shape[0]
        if isinstance(item, slice):
            start, stop, step = item.indices(size)
            if step > 0:
                if stop <= start:
                    return np.zeros(0, dtype=self.dtype)
                return self.backend.get_relative_range(start, stop)[::step]
            raise IndexError("negative steps are not supported on a sharded state")
        i = int(item)
        if i < 0:
            i += size
        if not (0 <= i < size):
            raise IndexError("index out of range")
        return self.backend.get_relative_range(i, i + 1)[0]

    def __array__(self, dtype=None, copy=None):
        if self.backend.n > 34:
            raise ValueError("a %d-qubit state does not fit a host array; slice the handle instead" % self.backend.n)
        a = self.backend.get_relative_range(0, self.shape[0])
        return a.astype(dtype) if dtype is not None else a


class ShardedB200Backend(object):
    def __init__(self, n: int, dtype, fuse: bool = True, tile_bits: int = 12, min_low_bits: int = 7,
                 peer_gates: bool = False, lazy_layout: bool = True, lazy_init=None, overlap=None):
        torch = _torch()
        import torch.distributed as dist
        if not dist.is_initialized():
            raise _lib.QipbError("ShardedB200Backend needs torch.distributed initialised (one rank per GPU)")
        self.dist = dist
        self.rank, self.P = dist.get_rank(), dist.get_world_size()
        self.G = int(round(math.log2(self.P)))
        if (1 << self.G) != self.P:
            raise ValueError("number of ranks must be a power of two")
        self.n = int(n)
        self.nl = self.n - self.G
        if self.nl < 1:
            raise ValueError("need at least one local qubit per rank")
        self.device = torch.device("cuda", torch.cuda.current_device())
        # local engine: a single-GPU backend over the nl local bits, re-used for its launch helpers
        self.eng = B200Backend(self.nl, dtype, device=self.device, fuse=fuse, tile_bits=tile_bits, min_low_bits=min_low_bits)
        self.L = self.eng.L
        self.ctx = self.eng.ctx
        self.code, self.amp_bytes, self.np_dtype = self.eng.code, self.eng.amp_bytes, self.eng.np_dtype
        self.layout = sp.Layout(self.n, self.G)
        self.queue: List[Gate] = []
        self._seg_cache, self._seg_keys = None, []      # compiled-circuit segments queued since the last flush
        self.fuse = fuse
        self.peer_gates = peer_gates
        self.lazy_layout = lazy_layout
        self._pending_init = None       # (groups, feeds): the state is built at the first flush
        self._virtual_init = None       # (per-local-bit factors, kron arguments): product state not yet written (lazy_init)
        self.lazy_init = os.environ.get("QIPB_LAZY_INIT", "0") == "1" if lazy_init is None else bool(lazy_init)
        self.stats = {"gates": 0, "exchanges": 0, "peer_gates": 0, "nvlink_bytes_out": 0}
        # exchange / compute overlap (see _run_overlapped): chunk bits per pipeline, passes per side that may join it
        self.overlap = os.environ.get("QIPB_SHARD_OVERLAP", "1") != "0" if overlap is None else bool(overlap)
        self.overlap_chunk_bits = int(os.environ.get("QIPB_OVERLAP_CHUNK_BITS", "3"))
        self.overlap_window = int(os.environ.get("QIPB_OVERLAP_WINDOW", "3"))
        # shards below this size are exchanged in one piece: a chunk pass of a 16 GiB shard is a fraction of a millisecond
        # and the pipeline's barriers (each waits for a kernel boundary) cost more than the exchange they would hide
        # (measured on 8 B200s: QFFT of a 33-qubit state, 16 GiB per GPU, 0.098 s without the pipeline, 0.153 s with it)
        self.overlap_min_bytes = int(os.environ.get("QIPB_OVERLAP_MIN_BYTES", str(1 << 35)))
        self._xs = None                 # second CUDA stream: the exchange of chunk j runs under the passes of its neighbours
        # shard memory comes from cudaMalloc (qipb_dev_alloc) so that its IPC handle maps it exactly
        self._token = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.peers = {}
        self.ptr = None
        self._free_orphans()
        key = self._pool_key()
        pooled = _SHARD_POOL.pop(key, None)
        if pooled is not None:
            self._wrap_shard(pooled[0])
            self.peers = pooled[1]
            self._sync_all()
        else:
            self._drain_pool()
            ptr = ctypes.c_void_p()
            _lib.check(self.L.qipb_dev_alloc(self.ctx, (1 << self.nl) * self.amp_bytes, ctypes.byref(ptr)))
            self._adopt_shard(ptr)

    def _pool_key(self):
        return (self.device.index or 0, (1 << self.nl) * self.amp_bytes, self.P)

    def _free_orphans(self):
        """Free the shards of states that were garbage-collected un-closed (collective: every rank runs the same
        program, so every rank holds the same orphans at this point)."""
        torch = _torch()
        if not _ORPHANS:
            return
        torch.cuda.synchronize()
        self._sync_all()
        torch.cuda.synchronize()
        while _ORPHANS:
            ptr, peers = _ORPHANS.pop()
            for pp in peers.values():
                self.L.qipb_ipc_close(self.ctx, pp)
            self.dist.barrier()
            self.L.qipb_dev_free(self.ctx, ptr)

    def _drain_pool(self):
        """Release every pooled shard (all ranks call this at the same point of the program)."""
        torch = _torch()
        if not _SHARD_POOL:
            return
        torch.cuda.synchronize()
        for key in list(_SHARD_POOL.keys()):
            ptr, peers = _SHARD_POOL.pop(key)
            for pp in peers.values():
                self.L.qipb_ipc_close(self.ctx, pp)
            self.dist.barrier()
            self.L.qipb_dev_free(self.ctx, ptr)

    def _wrap_shard(self, ptr):
        torch = _torch()
        self.ptr = ptr
        self.eng.state = torch.as_tensor(_RawCuda(ptr.value, 1 << self.nl, "<c16" if self.amp_bytes == 16 else "<c8"),
                                         device=self.device)
        assert self.eng.state.data_ptr() == ptr.value

    def _adopt_shard(self, ptr):
        """Wrap a cudaMalloc'ed shard for torch, export its IPC handle and map every peer's shard."""
        self._wrap_shard(ptr)
        handle = ctypes.create_string_buffer(64)
        _lib.check(self.L.qipb_ipc_export(self.ctx, ptr, handle))
        handles = [None] * self.P
        self.dist.all_gather_object(handles, bytes(handle.raw))
        self.peers = {}
        for r in range(self.P):
            if r != self.rank:
                pp = ctypes.c_void_p()
                _lib.check(self.L.qipb_ipc_open(self.ctx, handles[r], ctypes.byref(pp)))
                self.peers[r] = pp
        self._sync_all()

    def _release_shard(self):
        torch = _torch()
        torch.cuda.synchronize()
        self._sync_all()
        torch.cuda.synchronize()
        for pp in self.peers.values():
            self.L.qipb_ipc_close(self.ctx, pp)
        self.peers = {}
        self.dist.barrier()
        self.eng.state = None
        self.L.qipb_dev_free(self.ctx, self.ptr)
        self.ptr = None

    # ------------------------------------------------------------------ plumbing
    def _sync_all(self):
        """Stream-ordered barrier: peers' earlier kernels are complete when later work starts."""
        self.dist.all_reduce(self._token)

    def _stream(self):
        self.eng._stream()

    @property
    def profile(self):
        return self.eng.profile

    @profile.setter
    def profile(self, v):
        self.eng.profile = v

    def launch_count(self):
        return self.eng.launch_count()

    @staticmethod
    def make_state(n, index_groups, feed_list, statetype=np.complex128, **kwargs) -> "ShardedB200Backend":
        b = ShardedB200Backend(n, statetype, **kwargs)
        b._init_state(index_groups, feed_list)
        return b

    def _init_state(self, index_groups, feed_list):
        n = self.n
        groups = [[int(q) for q in g] for g in index_groups]
        flat = [q for g in groups for q in g]
        if len(groups) != len(feed_list) or len(set(flat)) != len(flat) or any(not (0 <= q < n) for q in flat):
            raise ValueError("bad feed groups")
        for g, f in zip(groups, feed_list):
            if isinstance(f, (int, np.integer)):
                if not (0 <= int(f) < 2 ** len(g)):
                    raise ValueError("one-hot feed index out of range")
            elif (f.shape[0] if hasattr(f, "shape") and len(f.shape) == 1 else np.asarray(f).reshape(-1).shape[0]) != 2 ** len(g):
                raise ValueError("feed length does not match 2**len(group)")
        self._pending_init = (groups, list(feed_list))
        if not self.lazy_layout:
            self._materialise()

    def _materialise(self):
        """Build the initial state.  Deferred to the first flush so that the shard layout can be chosen
        from the queued gates (qubits needed non-diagonally last go on the rank bits)."""
        if self._pending_init is None:
            return
        torch = _torch()
        groups, feed_list = self._pending_init
        self._pending_init = None
        n, nl = self.n, self.nl
        flat = [q for g in groups for q in g]
        if self.lazy_layout:
            sp.choose_initial_layout(self.queue, self.layout)
        pos = self.layout.pos
        self._stream()
        if not groups:
            _lib.check(self.L.qipb_init_basis(self.ctx, self.ptr, nl, self.code, 0 if self.rank == 0 else -1))
            return
        from .backend import feeds_to_device, split_feeds
        vgroups, vfeeds, fixed_mask, fixed_value = split_feeds(groups, feed_list, n, lambda q: pos[q])
        if not vgroups:                                        # only one-hot feeds: a basis state
            mine = (fixed_value >> nl) == self.rank
            _lib.check(self.L.qipb_init_basis(self.ctx, self.ptr, nl, self.code,
                                              (fixed_value & ((1 << nl) - 1)) if mine else -1))
            return
        kron_args = (vgroups, vfeeds, fixed_mask, fixed_value, list(pos))
        if self.lazy_init and self.fuse:
            # product of one-qubit feeds: kept virtual; the first rank-local fused pass writes its tiles from the per-bit
            # factors (qipb_apply_fused_fill).  The factors of the rank bits are a scalar of this shard.
            from .backend import product_state_factors
            factors = product_state_factors(vgroups, vfeeds, fixed_mask, fixed_value, n, lambda q: pos[q])
            if factors is not None:
                scalar = 1.0 + 0j
                for b in range(nl, n):
                    scalar *= factors[b][(self.rank >> (b - nl)) & 1]
                local = list(factors[:nl])
                local[0] = (local[0][0] * scalar, local[0][1] * scalar)
                self._virtual_init = (local, kron_args)
                return
        self._launch_kron(*kron_args)

    def _launch_kron(self, vgroups, vfeeds, fixed_mask, fixed_value, pos):
        from .backend import feeds_to_device
        dev_feeds = feeds_to_device(vfeeds, self.device)
        _lib.check(self.L.qipb_init_kron(self.ctx, self.ptr, self.nl, self.code, len(vgroups),
                                         _lib.int_array([len(g) for g in vgroups]),
                                         _lib.int_array([pos[q] for g in vgroups for q in g]),
                                         ctypes.c_void_p(dev_feeds.data_ptr()), fixed_mask, fixed_value, self.rank))
        self._keep = dev_feeds

    def _materialise_virtual(self):
        """Build a still-virtual product state with the stand-alone kron kernel."""
        if self._virtual_init is None:
            return
        torch = _torch()
        _, kron_args = self._virtual_init
        self._virtual_init = None
        with torch.cuda.device(self.device):
            self._stream()
            self._launch_kron(*kron_args)

    # ------------------------------------------------------------------ gates
    def apply_gates(self, gates, cache=None, key=None) -> None:
        """Pre-decoded gates of a compiled circuit (qip_b200.graph).  The exchange plan depends on the whole
        queue, so segments are only queued here; with `cache`/`key` the flush that runs them remembers its
        rank-local program under (keys of all its segments, layout before) and a replay skips scheduling."""
        self.queue.extend(gates)
        self.stats["gates"] += len(gates)
        if cache is not None and self._seg_keys is not None:
            self._seg_cache = cache
            self._seg_keys.append(key)
        else:
            self._seg_keys = None                  # mixed with un-keyed gates: this flush is not cacheable

    def kronselect_dot(self, mats, input_offset: int = 0, output_offset: int = 0) -> None:
        if input_offset != 0 or output_offset != 0:
            raise ValueError("offset windows are not supported; the state is sharded by its top qubits")
        for g in decode_mats(mats, self.n):
            s = simplify(g)
            if s is not None:
                self.queue.append(s)
                self.stats["gates"] += 1
                self._seg_keys = None              # un-keyed gates in the queue: the next flush is not cacheable

    def _plan_local(self, batch: List[BitGate]):
        """Rank-local: every rank merges and plans its own resolved gate list."""
        if self.fuse:
            batch = merge_bitgates(batch, 2)
        return plan_passes(batch, self.nl, self.amp_bytes, tile_bits=self.eng.tile_bits,
                           min_low_bits=self.eng.min_low_bits, enable=self.fuse)

    def _fill_first_pass(self, passes):
        """Virtual product state + first rank-local fused pass in one write-only sweep (B200Backend._fill_first_pass)."""
        from .backend import pack_fill_pass
        torch = _torch()
        factors, _ = self._virtual_init
        p = passes[0]
        if self.nl + len(p.gates) <= _lib.MAX_FUSED_GATES:
            arr, tbits = pack_fill_pass(factors, p)
            if self.eng.profile is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            rc = self.L.qipb_apply_fused_fill(self.ctx, self.ptr, self.nl, self.code, len(p.tile_bits), tbits,
                                              self.nl + len(p.gates), arr)
            if rc == 0:
                if self.eng.profile is not None:
                    e1.record()
                    self.eng.profile.append(("fused_kernel[fill]", float(self.amp_bytes) * 2.0 ** self.nl, e0, e1))
                self._virtual_init = None
                self.stats["fill_passes"] = self.stats.get("fill_passes", 0) + 1
                return passes[1:]
            if rc != _lib.ERR_UNSUPPORTED:
                _lib.check(rc)
        self._materialise_virtual()
        return passes

    def _run_passes(self, passes):
        torch = _torch()
        for p in passes:
            if self.eng.profile is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if p.fused:
                self.eng._launch_fused(p)
            else:
                self.eng._launch_single(p.gates[0])
            if self.eng.profile is not None:
                from .backend import kernel_label
                e1.record()
                self.eng.profile.append(kernel_label(p, self.nl, self.amp_bytes) + (e0, e1))

    def _exchange(self, a: sp.Exchange):
        gb = a.gpos - self.nl
        my_g = (self.rank >> gb) & 1
        partner = self.rank ^ (1 << gb)
        half = 1 << (self.nl - 2) if self.nl >= 2 else 0
        total = 1 << (self.nl - 1)
        begin, count = (0, total - half) if my_g == 0 else (total - half, half)
        torch = _torch()
        self._sync_all()
        if self.eng.profile is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.check(self.L.qipb_peer_swap_bit(self.ctx, self.ptr, self.peers[partner], self.nl, self.code, a.lpos, my_g,
                                             begin, count))
        if self.eng.profile is not None:
            e1.record()
            # NVLink bytes leaving this GPU: half a shard (reads served to the partner + writes to it)
            self.eng.profile.append(("peer_swap_bit_kernel[nvlink]", float(self.amp_bytes * total), e0, e1))
        self._sync_all()
        self.stats["exchanges"] += 1
        self.stats["nvlink_bytes_out"] += self.amp_bytes * total

    def _remap_args(self, pairs):
        """(g, my value on the exchanged rank bits, peer pointer table indexed by value, local bits) of a remap."""
        g = len(pairs)
        value = 0
        for t, (gpos, _) in enumerate(pairs):
            value |= ((self.rank >> (gpos - self.nl)) & 1) << t
        peers = (ctypes.c_void_p * 8)()
        for b in range(1 << g):
            if b == value:
                continue
            r = self.rank
            for t, (gpos, _) in enumerate(pairs):
                bit = 1 << (gpos - self.nl)
                r = (r | bit) if (b >> t) & 1 else (r & ~bit)
            peers[b] = self.peers[r]
        return g, value, peers, _lib.int_array([l for _, l in pairs])

    def _multi_exchange(self, a: sp.MultiExchange):
        torch = _torch()
        g, value, peers, lbits = self._remap_args(a.pairs)
        self._sync_all()
        if self.eng.profile is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.check(self.L.qipb_peer_remap(self.ctx, self.ptr, peers, self.nl, self.code, g, lbits, value))
        nbytes = self.amp_bytes * ((1 << self.nl) - (1 << (self.nl - g)))
        if self.eng.profile is not None:
            e1.record()
            self.eng.profile.append(("peer_remap_kernel[nvlink]", float(nbytes), e0, e1))
        self._sync_all()
        self.stats["exchanges"] += 1
        self.stats["multi_exchanges"] = self.stats.get("multi_exchanges", 0) + 1
        self.stats["nvlink_bytes_out"] += nbytes

    # ------------------------------------------------------------------ exchange / compute overlap
    # An exchange is NVLink-bound (0.7 TB/s per direction against 6.5 TB/s of HBM) and between two barriers: run alone it
    # idles the SMs and most of HBM for 100 ms (one bit) to 155 ms (three bits) per layer at 33 local qubits -- a third
    # of an 8-GPU layer.  The state is therefore cut into 2^c CHUNKS along c local bits that are neither exchanged nor
    # tile bits of the passes next to the exchange ("quiet" bits: to those passes and to the remap they are ordinary
    # outside bits, so each of them acts on every chunk independently), and the window
    #       passes A (before)  ->  exchange X  ->  passes B (after)
    # runs as a pipeline over the chunks on two streams: A(j+1) and B(j-1) compute while X(j) crosses NVLink.  Chunked
    # launches are the same kernels (qipb_apply_fused_chunk / qipb_peer_remap_chunk: tile / pair enumeration skips the
    # chunk bits); the remap of a chunk is a persistent launch of one CTA per SM, which fits beside the three resident
    # CTAs of the fused kernel.  Cross-rank ordering: stream-ordered barriers on the exchange stream around every X(j)
    # (all ranks have finished A(j) before anyone touches chunk j of a peer; all peers are done with it before B(j)).
    def _xstream(self):
        """(compute stream, exchange stream) of the chunk pipeline.  The chunked passes run on a HIGH-priority stream, the
        chunked remaps on a normal one: at every kernel boundary the block scheduler then hands the freed SM slots to
        the next pass first (3 CTAs per SM), and the persistent remap CTAs -- 256 threads x 56 registers -- only ever
        fit one per SM beside them.  Measured on 2 B200s without priorities: remap CTAs packed four to an SM whenever a
        pass ended, the following pass ran on the SMs that were left, and the overlap bought nothing."""
        torch = _torch()
        if self._xs is None:
            self._xs = (torch.cuda.Stream(device=self.device, priority=-1), torch.cuda.Stream(device=self.device, priority=0))
        return self._xs

    def _plan_overlap(self, prev_passes, xstep, next_passes):
        """(a, b, chunk bits): how many of this rank's passes before / after the exchange join the pipeline, and the
        local bits along which the state is cut (chosen rank-independently by shardplan.annotate_chunks); None when
        there is nothing to overlap."""
        cbits = list(getattr(xstep, "chunk_bits", None) or [])
        if not self.overlap or not cbits:
            return None
        cset = set(cbits)

        def joinable(passes):
            k = 0
            for p in passes:
                if k >= self.overlap_window or not p.fused or (cset & set(p.tile_bits)):
                    break
                k += 1
            return k

        # (every rank runs the chunked exchange once chunk bits are set -- the number of barriers must agree -- even if
        # none of ITS passes can join)
        return joinable(list(reversed(prev_passes))), joinable(next_passes), cbits

    def _run_overlapped(self, prev_passes, xstep, next_passes, plan):
        """passes A -> exchange -> passes B as a pipeline over chunks (see above).  Returns nothing; every launch is
        asynchronous, the main stream ends ordered behind the last chunk's exchange."""
        torch = _torch()
        a, b, cbits = plan
        pairs = xstep.pairs if isinstance(xstep, sp.MultiExchange) else [(xstep.gpos, xstep.lpos)]
        g, value, peers, lbits = self._remap_args(pairs)
        K = 1 << len(cbits)
        fix = _lib.int_array(cbits)
        head, A = (prev_passes[:len(prev_passes) - a], prev_passes[len(prev_passes) - a:]) if a else (prev_passes, [])
        B = next_passes[:b]
        main = torch.cuda.current_stream(self.device)
        cs, xs = self._xstream()
        # CTAs of the persistent remap: half an SM count saturates NVLink (measured on 2 B200s: 13.3 ms per 1/8 chunk
        # with 74 CTAs as with 148) and leaves the fused passes beside it 10 % slower instead of 25 % -- more CTAs only
        # queue more remote stores, whose back-pressure stalls the local traffic of the passes
        sm = int(float(os.environ.get("QIPB_XCHG_CTAS_PER_SM", "0.5")) * self.eng.sm_count())
        max_ctas = max(1, sm // ((1 << g) - 1))
        prof = self.eng.profile
        chunk_bytes_pass = 2.0 * self.amp_bytes * 2.0 ** self.nl / K
        nbytes = self.amp_bytes * ((1 << self.nl) - (1 << (self.nl - g)))
        self._run_passes(head)
        start = torch.cuda.Event()
        start.record()                                        # everything queued so far on the caller's stream
        cs.wait_event(start)

        def chunk_passes(passes, fv):
            for p in passes:
                if prof is not None:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                self.eng._launch_fused_chunk(p, cbits, fv)
                if prof is not None:
                    e1.record()
                    prof.append(("fused_kernel", chunk_bytes_pass, e0, e1))

        # One barrier between consecutive chunk exchanges serves both purposes: behind it every rank has finished X(j-1)
        # (chunk j-1 is final everywhere: B(j-1) may start) and A(j) (chunk j of every peer may be touched).  The NCCL
        # barrier kernel does not fit beside three fused CTAs and a remap CTA, so it runs at a kernel boundary of the
        # compute stream: every barrier costs the exchange stream about half a chunk pass -- hence K + 1 of them, not 2 K.
        done_x = []
        fvs = []
        for j in range(K):
            fv = 0
            for t, pbit in enumerate(cbits):
                fv |= ((j >> t) & 1) << pbit
            fvs.append(fv)
            with torch.cuda.stream(cs):
                self._stream()
                chunk_passes(A, fv)
                ready = torch.cuda.Event()
                ready.record()                                # this rank's A(j) (and everything before) is done
            with torch.cuda.stream(xs):
                xs.wait_event(ready)
                self._sync_all()                              # every rank: A(j) done, X(j-1) done
                if j > 0:
                    ev = torch.cuda.Event()
                    ev.record()
                    done_x.append((fvs[j - 1], ev))
                self._stream()
                if prof is not None:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                _lib.check(self.L.qipb_peer_remap_chunk(self.ctx, self.ptr, peers, self.nl, self.code, g, lbits, value,
                                                        len(cbits), fix, fv, max_ctas))
                if prof is not None:
                    e1.record()
                    prof.append(("peer_remap_kernel[nvlink]", float(nbytes) / K, e0, e1))
        with torch.cuda.stream(xs):
            self._sync_all()                                  # every peer is done with the last chunk
            ev = torch.cuda.Event()
            ev.record()
            done_x.append((fvs[K - 1], ev))
        with torch.cuda.stream(cs):
            self._stream()
            for fv, ev in done_x:
                cs.wait_event(ev)
                chunk_passes(B, fv)
            end = torch.cuda.Event()
            end.record()
        main.wait_event(end)
        if not B:
            for _, ev in done_x:
                main.wait_event(ev)
        self._stream()                                        # the library is back on the caller's stream
        self.stats["exchanges"] += 1
        if isinstance(xstep, sp.MultiExchange):
            self.stats["multi_exchanges"] = self.stats.get("multi_exchanges", 0) + 1
        self.stats["nvlink_bytes_out"] += nbytes
        self.stats["overlapped_exchanges"] = self.stats.get("overlapped_exchanges", 0) + 1
        self.stats["overlap_passes"] = self.stats.get("overlap_passes", 0) + a + b
        self.stats["overlap_chunks"] = self.stats.get("overlap_chunks", 0) + K

    def _peer_gate(self, a: sp.PeerGate1):
        gb = a.gpos - self.nl
        my_g = (self.rank >> gb) & 1
        partner = self.rank ^ (1 << gb)
        cg = a.ctrl_mask >> self.nl
        total = 1 << self.nl
        half = total // 2
        off, count = (0, total - half) if my_g == 0 else (total - half, half)
        self._sync_all()
        if (self.rank & cg) == cg:
            _lib.check(self.L.qipb_peer_gate1(self.ctx, self.ptr, self.peers[partner], self.code, off, count,
                                              _lib.mat_array(a.mat), my_g, a.ctrl_mask & ((1 << self.nl) - 1)))
        self._sync_all()
        self.stats["peer_gates"] += 1
        self.stats["nvlink_bytes_out"] += self.amp_bytes * half

    def _run_step(self, a):
        if self._virtual_init is not None:
            if isinstance(a, tuple) and a[1] and a[1][0].fused:
                a = ("local", self._fill_first_pass(a[1]))
            else:
                self._materialise_virtual()
        if isinstance(a, tuple):
            self._run_passes(a[1])
        elif isinstance(a, sp.Exchange):
            self._exchange(a)
        elif isinstance(a, sp.MultiExchange):
            self._multi_exchange(a)
        else:
            self._peer_gate(a)

    def _run_program(self, program):
        """Execute a rank-local program.  Wherever an exchange has fused passes next to it and quiet local bits exist,
        the window runs as a chunk pipeline (_run_overlapped); everything else step by step."""
        torch = _torch()
        steps = list(program)
        xtypes = (sp.Exchange, sp.MultiExchange)
        with torch.cuda.device(self.device):
            self._stream()
            if self._virtual_init is not None:                 # lazy product state: the first fused pass writes it
                if steps and isinstance(steps[0], tuple) and steps[0][1] and steps[0][1][0].fused:
                    steps[0] = ("local", self._fill_first_pass(steps[0][1]))
                else:
                    self._materialise_virtual()
            i = 0
            while i < len(steps):
                st = steps[i]
                prev = x = None
                if isinstance(st, tuple) and i + 1 < len(steps) and isinstance(steps[i + 1], xtypes):
                    prev, x, xi = st[1], steps[i + 1], i + 1
                elif isinstance(st, xtypes):
                    prev, x, xi = [], st, i
                if x is not None:
                    after = steps[xi + 1][1] if (xi + 1 < len(steps) and isinstance(steps[xi + 1], tuple)) else []
                    plan = self._plan_overlap(prev, x, after)
                    if plan is not None:
                        self._run_overlapped(prev, x, after[:plan[1]], plan)
                        if after:
                            steps[xi + 1] = ("local", after[plan[1]:])     # what is left may lead the next window
                        i = xi + 1
                        continue
                self._run_step(st)
                i += 1

    def _compile_and_run(self, actions):
        """Plan the whole flush, then run it (the host plans flush k+1 while the device still executes flush k: every
        launch is asynchronous and the barriers are stream-ordered)."""
        torch = _torch()
        actions = list(actions)
        ref_rank = self.P - 1
        program = sp.compile_program(actions, self.nl, self.rank, self._plan_local)
        if self.overlap and self.fuse and self.P > 1 and (self.amp_bytes << self.nl) >= self.overlap_min_bytes and \
                any(isinstance(a, (sp.Exchange, sp.MultiExchange)) for a in actions):
            # the chunk bits of every exchange are read off the program of ONE agreed rank, so that all ranks cut alike
            ref = program if self.rank == ref_rank else sp.compile_program(actions, self.nl, ref_rank, self._plan_local)
            sp.annotate_chunks(ref, self.nl, self.overlap_chunk_bits, min(self.eng.min_low_bits, self.nl), self.overlap_window)
        self._run_program(program)
        return program

    def _execute(self, actions):
        self._compile_and_run(actions)

    def flush(self) -> None:
        self._materialise()
        cache, keys = self._seg_cache, self._seg_keys
        self._seg_cache, self._seg_keys = None, []
        if not self.queue:
            self._materialise_virtual()
            return
        gates = list(self.queue)
        self.queue = []
        ckey = None
        if cache is not None and keys:
            # what the schedule depends on: the gates (named by their segment keys), the layout they start
            # from, and the knobs of the rank-local planner; the program itself is per rank
            ckey = ("sharded-flush", tuple(keys), tuple(self.layout.pos), self.rank, self.P, self.fuse,
                    self.peer_gates, self.eng.tile_bits, self.eng.min_low_bits, self.amp_bytes)
            hit = cache.get(ckey)
            if hit is not None:
                program, pos_after = hit
                self.layout.pos = list(pos_after)
                self.stats["cached_flushes"] = self.stats.get("cached_flushes", 0) + 1
                self._run_program(program)
                self._materialise_virtual()
                return
        program = self._compile_and_run(sp.schedule(gates, self.layout, peer_gates=self.peer_gates,
                                                    count_passes=lambda batch: len(self._plan_local(batch)),
                                                    tile_bits=self.eng.tile_bits, min_low_bits=self.eng.min_low_bits))
        if ckey is not None:
            cache[ckey] = (program, tuple(self.layout.pos))
        self._materialise_virtual()                 # nothing ran (only relabelled swaps were queued)

    def func_apply(self, reg1_indices, reg2_indices, func, input_offset: int = 0, output_offset: int = 0) -> None:
        torch = _torch()
        reg1 = [int(i) for i in reg1_indices]
        reg2 = [int(i) for i in reg2_indices]
        allq = reg1 + reg2
        if len(set(allq)) != len(allq) or any(not (0 <= q < self.n) for q in allq):
            raise ValueError("func_apply registers must be disjoint qubit indices in [0, n)")
        table = tabulate(func, len(reg1))
        self.flush()
        # the written register must be local: swap any rank-held reg2 qubit with a free local one
        moves = []
        for q in reg2:
            if self.layout.is_global(q):
                for p in range(self.nl - 1, -1, -1):
                    v = self.layout.qubit_at(p)
                    if v not in reg2:
                        moves.append(sp.Exchange(self.layout.pos[q], p))
                        self.layout.swap_qubits(q, v)
                        break
                else:
                    raise ValueError("func_apply: output register does not fit in one shard")
        self._execute(moves)
        x_fixed = 0
        r1bits = []
        for j, q in enumerate(reg1):
            p = self.layout.pos[q]
            if p >= self.nl:
                r1bits.append(-1)
                if (self.rank >> (p - self.nl)) & 1:
                    x_fixed |= 1 << (len(reg1) - 1 - j)
            else:
                r1bits.append(p)
        with torch.cuda.device(self.device):
            self._stream()
            from .backend import device_table
            dev_table, small = device_table(func, table, self.device, len(reg2))
            entry = self.L.qipb_func_xor_u8 if small else self.L.qipb_func_xor
            _lib.check(entry(self.ctx, self.ptr, self.nl, self.code, len(reg1), _lib.int_array(r1bits),
                             len(reg2), _lib.int_array([self.layout.pos[q] for q in reg2]),
                             ctypes.c_void_p(dev_table.data_ptr()), x_fixed))
            self._keep = dev_table

    # ------------------------------------------------------------------ measurement
    def _split(self, pos_list):
        loc = [p for p in pos_list if p < self.nl]
        glob = [p for p in pos_list if p >= self.nl]
        return loc, glob

    def _probabilities(self, indices, order: str, filter_qubits=(), filter_value_bits=(), on_device: bool = False):
        """Histogram over `indices`; filter = (qubits, their required bit values)."""
        torch = _torch()
        n = self.n
        idx = [int(i) for i in indices]
        k = len(idx)
        if len(set(idx)) != k or any(not (0 <= q < n) for q in idx):
            raise ValueError("measured indices must be distinct qubit indices in [0, n)")
        self.flush()
        if order == "given-le":
            qorder, outb = idx, list(range(k))
        else:
            qorder, outb = sorted(idx), [k - 1 - j for j in range(k)]
        # split into local measured bits (histogrammed by the kernel) and rank-held ones (fixed per rank)
        loc = [(self.layout.pos[q], ob) for q, ob in zip(qorder, outb) if self.layout.pos[q] < self.nl]
        glo = [(self.layout.pos[q], ob) for q, ob in zip(qorder, outb) if self.layout.pos[q] >= self.nl]
        fmask = fval = 0
        active = True
        for q, v in zip(filter_qubits, filter_value_bits):
            p = self.layout.pos[q]
            if p < self.nl:
                fmask |= 1 << p
                fval |= (v & 1) << p
            elif ((self.rank >> (p - self.nl)) & 1) != (v & 1):
                active = False
        loc_sorted = sorted(loc, key=lambda t: t[1])           # keep relative output order
        kl = len(loc_sorted)
        with torch.cuda.device(self.device):
            self._stream()
            out = torch.zeros(2 ** k, dtype=torch.float64, device=self.device)
            if active:
                part = torch.empty(2 ** kl, dtype=torch.float64, device=self.device)
                _lib.check(self.L.qipb_probabilities(self.ctx, self.ptr, self.nl, self.code, kl,
                                                     _lib.int_array([p for p, _ in loc_sorted]),
                                                     _lib.int_array(list(range(kl))), fmask, fval,
                                                     ctypes.c_void_p(part.data_ptr())))
                base = 0
                for p, ob in glo:
                    if (self.rank >> (p - self.nl)) & 1:
                        base |= 1 << ob
                j = np.arange(2 ** kl, dtype=np.int64)
                tgt = np.full(2 ** kl, base, dtype=np.int64)
                for t, (_, ob) in enumerate(loc_sorted):
                    tgt |= ((j >> t) & 1) << ob
                out.index_copy_(0, torch.from_numpy(tgt).to(self.device), part)
            self.dist.all_reduce(out)
            return out if on_device else out.cpu().numpy()

    def total_prob(self) -> float:
        return float(self._probabilities([], "sorted-be")[0])

    def soft_measure(self, indices, measured: Optional[int] = None, input_offset: int = 0):
        k = len(indices)
        r = random.random()          # every rank draws the same value only if seeded identically
        rt = _torch().tensor([r], dtype=_torch().float64, device=self.device)
        self.dist.broadcast(rt, 0)   # rank 0's draw decides (one draw consumed per rank, like the reference)
        r = float(rt.item())
        if measured is not None:
            srt = sorted(int(i) for i in indices)
            bits = [(int(measured) >> (k - 1 - j)) & 1 for j in range(k)]
            p = float(self._probabilities([], "sorted-be", srt, bits)[0])
            return int(measured), p
        return scan_outcome(self._probabilities(indices, "sorted-be"), r)

    def measure(self, indices, measured: Optional[int] = None, measured_prob: Optional[float] = None,
                input_offset: int = 0, output_offset: int = 0):
        torch = _torch()
        k = len(indices)
        _check_measure_args(k, measured, measured_prob)
        if measured is None or measured_prob is None:
            m, p = self.soft_measure(indices, measured=measured)
        else:
            m, p = int(measured), float(measured_prob)
        self.flush()
        srt = sorted(int(i) for i in indices)
        mask = want = 0
        alive = True
        for j, q in enumerate(srt):
            v = (m >> (k - 1 - j)) & 1
            pos = self.layout.pos[q]
            if pos < self.nl:
                mask |= 1 << pos
                want |= v << pos
            elif ((self.rank >> (pos - self.nl)) & 1) != v:
                alive = False
        with torch.cuda.device(self.device):
            self._stream()
            if alive:
                _lib.check(self.L.qipb_collapse(self.ctx, self.ptr, self.nl, self.code, mask, want, math.sqrt(1.0 / p)))
            else:
                _lib.check(self.L.qipb_init_basis(self.ctx, self.ptr, self.nl, self.code, -1))
        return m, p

    def reduce_measure(self, indices, measured: Optional[int] = None, measured_prob: Optional[float] = None,
                       input_offset: int = 0, output_offset: int = 0):
        """Measured qubits are first made local (bit exchanges), then every rank compacts its shard;
        the state keeps P shards of 2^(n-k-G) amplitudes and n decreases by k."""
        torch = _torch()
        k = len(indices)
        _check_measure_args(k, measured, measured_prob)
        if self.nl - k < 1:
            raise ValueError("reduce_measure would leave less than one local qubit per rank")
        if measured is None or measured_prob is None:
            m, p = self.soft_measure(indices, measured=measured)
        else:
            m, p = int(measured), float(measured_prob)
        self.flush()
        srt = sorted(int(i) for i in indices)
        moves = []
        for q in srt:
            if self.layout.is_global(q):
                for pp in range(self.nl - 1, -1, -1):
                    v = self.layout.qubit_at(pp)
                    if v not in srt:
                        moves.append(sp.Exchange(self.layout.pos[q], pp))
                        self.layout.swap_qubits(q, v)
                        break
        self._execute(moves)
        mask = want = 0
        for j, q in enumerate(srt):
            pos = self.layout.pos[q]
            mask |= 1 << pos
            want |= ((m >> (k - 1 - j)) & 1) << pos
        new_nl = self.nl - k
        with torch.cuda.device(self.device):
            self._stream()
            torch.cuda.synchronize()
            self._sync_all()
            torch.cuda.synchronize()
            new_ptr = ctypes.c_void_p()
            _lib.check(self.L.qipb_dev_alloc(self.ctx, (1 << new_nl) * self.amp_bytes, ctypes.byref(new_ptr)))
            _lib.check(self.L.qipb_reduce(self.ctx, self.ptr, new_ptr, self.nl, self.code, mask, want, math.sqrt(1.0 / p)))
            torch.cuda.synchronize()
            # remaining logical qubits are renumbered in ascending order; local positions are compacted
            remaining = [q for q in range(self.n) if q not in srt]
            old_pos = [self.layout.pos[q] for q in remaining]
            new_layout = sp.Layout(self.n - k, self.G)
            for newq, op in enumerate(old_pos):
                if op >= self.nl:
                    new_layout.pos[newq] = op - k
                else:
                    new_layout.pos[newq] = op - bin(mask & ((1 << op) - 1)).count("1")
            self._release_shard()
            self.n -= k
            self.nl = new_nl
            self.layout = new_layout
            self.eng.n = new_nl
            self._adopt_shard(new_ptr)
        return m, p

    def measure_probabilities(self, indices, top_k: int = 0):
        if top_k:
            from .backend import _DEVICE_TOPK_MIN_QUBITS, top_probabilities_device
            if len(indices) > _DEVICE_TOPK_MIN_QUBITS:
                return top_probabilities_device(self._probabilities(indices, "sorted-be", on_device=True), top_k)
            return top_probabilities(self._probabilities(indices, "sorted-be"), top_k)
        return self._probabilities(indices, "given-le")

    # ------------------------------------------------------------------ state access
    def get_state(self):
        """qip/backend.py:106-107, called unconditionally at the end of every run() (qip/pipeline.py:248).  Up to 30
        qubits: canonical-order host copy of the WHOLE state on every rank.  Beyond: a ShardedState handle (the
        state stays sharded in HBM, canonicalised; slices are gathered on demand), like B200Backend's DeviceState."""
        torch = _torch()
        self.flush()
        self._execute(sp.canonicalise(self.layout))
        if self.n > _HOST_STATE_MAX_QUBITS:
            return ShardedState(self)
        parts = [torch.empty(1 << self.nl, dtype=self.eng.tdtype, device=self.device) for _ in range(self.P)]
        self.dist.all_gather(parts, self.eng.state)
        return torch.cat(parts).cpu().numpy()

    def get_state_size(self) -> int:
        return 2 ** self.n

    def _my_overlap(self, start: int, end: int):
        lo, hi = self.rank << self.nl, (self.rank + 1) << self.nl
        a, b = max(start, lo), min(end, hi)
        return (a, b, lo) if a < b else None

    def get_relative_range(self, start: int, end: int):
        """qip/backend.py:165-166 on the canonical (index-ordered) state; every rank gets the range."""
        self.flush()
        self._execute(sp.canonicalise(self.layout))
        ov = self._my_overlap(start, end)
        piece = self.eng.state[ov[0] - ov[2]:ov[1] - ov[2]].cpu().numpy() if ov else np.zeros(0, dtype=self.np_dtype)
        parts = [None] * self.P
        self.dist.all_gather_object(parts, piece)
        return np.concatenate(parts)

    def overwrite_relative_range(self, start: int, end: int, data):
        torch = _torch()
        self.flush()
        self._execute(sp.canonicalise(self.layout))
        ov = self._my_overlap(start, end)
        if ov:
            src = np.ascontiguousarray(np.asarray(data, dtype=self.np_dtype)[ov[0] - start:ov[1] - start])
            self.eng.state[ov[0] - ov[2]:ov[1] - ov[2]].copy_(torch.from_numpy(src))
        self._sync_all()

    def addto_relative_range(self, start: int, end: int, data):
        torch = _torch()
        self.flush()
        self._execute(sp.canonicalise(self.layout))
        ov = self._my_overlap(start, end)
        if ov:
            src = np.ascontiguousarray(np.asarray(data, dtype=self.np_dtype)[ov[0] - start:ov[1] - start])
            with torch.cuda.device(self.device):
                self._stream()
                dev = torch.from_numpy(src).to(self.device)
                _lib.check(self.L.qipb_add_range(self.ctx, self.ptr, self.code, ov[0] - ov[2], ov[1] - ov[0],
                                                 ctypes.c_void_p(dev.data_ptr())))
                torch.cuda.current_stream(self.device).synchronize()
        self._sync_all()

    def synchronize(self):
        self.flush()
        _torch().cuda.current_stream(self.device).synchronize()

    def close(self):
        """Return the shard (with its peer mappings) to the pool; nothing may touch it afterwards."""
        if getattr(self, "ptr", None) is None:
            return
        torch = _torch()
        torch.cuda.synchronize()
        self._sync_all()                        # no peer kernel is still reading or writing this shard
        torch.cuda.synchronize()
        key = self._pool_key()
        if key in _SHARD_POOL:                  # two live states of one size: keep one pooled, free the other
            self._release_shard()
        else:
            _SHARD_POOL[key] = (self.ptr, self.peers)
            self.peers = {}
            self.eng.state = None
            self.ptr = None
        self.eng.close()

    def __del__(self):
        """A state dropped without close() -- the reference's run() returns get_state() and forgets the backend
        (qip/pipeline.py:248) -- must not leak its shard (up to 128 GiB plus P-1 peer mappings).  No collective may run
        from a finaliser: the shard goes back to the pool as it is (the next user orders itself behind every rank's
        outstanding work with its first stream-ordered barrier) or, if the slot is taken, onto the orphan list that the
        next construction frees collectively."""
        try:
            if getattr(self, "ptr", None) is None:
                return
            key = self._pool_key()
            if key in _SHARD_POOL:
                _ORPHANS.append((self.ptr, self.peers))
            else:
                _SHARD_POOL[key] = (self.ptr, self.peers)
            self.ptr, self.peers = None, {}
            self.eng.state = None
            self.eng.close()
        except Exception:
            pass
