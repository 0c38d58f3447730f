"""Tests-only HOST double of libqipb200 for the single-GPU entry points (include/qip_b200.h), so that the real
qip_b200.backend.B200Backend -- queueing, relabelling, planning, measurement conventions, range access, the lazy
product-state init -- runs on the CPU tier.

  * qipb_apply_fused / qipb_apply_fused_fill go to the host emulator (tests/csrc/fused_emul.cu: the product's own
    lowering and sweep code);
  * every other entry point is restated in numpy from the contract written in the header.
"Device memory" is the memory of CPU torch tensors; `install(monkeypatch)` swaps the library loader and the torch.cuda
calls of qip_b200.backend for these doubles.  Nothing here is importable from the product."""
import contextlib
import ctypes
import os
import sys

import numpy as np

import bitsim
from qip_b200 import lib as qlib
from qip_b200.ops import BitGate

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc"))
import build_emul  # noqa: E402


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, int):
        return p
    if isinstance(p, ctypes.c_void_p):
        return p.value or 0
    return ctypes.cast(p, ctypes.c_void_p).value or 0


def _view(ptr, count, dtype):
    buf = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(_addr(ptr))
    return np.frombuffer(buf, dtype=dtype, count=count)


def _amps(ptr, nbits, code):
    return _view(ptr, 1 << nbits, np.complex128 if code == qlib.C128 else np.complex64)


def _ints(p, n):
    return [int(p[i]) for i in range(n)]


class HostLib(object):
    check_chunk_isolation = True        # (off for virtual ranks: there a peer thread exchanges the OTHER chunks meanwhile)

    def __init__(self):
        self.emul = ctypes.CDLL(build_emul.build())
        argt = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                ctypes.POINTER(qlib.Gate), ctypes.POINTER(ctypes.c_int)]
        self.emul.qipb_emul_fused.argtypes = argt
        self.emul.qipb_emul_fused_fill.argtypes = argt
        self.emul.qipb_emul_fused_chunk.argtypes = argt + [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_ulonglong]
        self.emul.qipb_emul_last_error.restype = ctypes.c_char_p
        self.launches = 0
        self.ext_launches = 0
        self.err = b""
        self.log = []

    # ---- context ----
    def qipb_version(self):
        return 100

    def qipb_last_error(self):
        return self.err

    def qipb_create(self, device, out):
        out._obj.value = 1
        return 0

    def qipb_destroy(self, ctx):
        return 0

    def qipb_set_stream(self, ctx, stream):
        return 0

    def qipb_sync(self, ctx):
        return 0

    def qipb_launch_count(self, ctx):
        return self.launches

    def qipb_ring_launch_count(self, ctx):
        return 0

    def qipb_ext_launch_count(self, ctx):
        return self.ext_launches

    # ---- state construction ----
    def qipb_init_basis(self, ctx, state, nbits, code, index):
        self.log.append("init_basis")
        a = _amps(state, nbits, code)
        a[:] = 0
        if 0 <= index < (1 << nbits):
            a[index] = 1
        self.launches += 1
        return 0

    def qipb_init_kron(self, ctx, state, nbits, code, ngroups, group_len, group_bits, feeds, fixed_mask, fixed_value, shard):
        self.log.append("init_kron")
        a = _amps(state, nbits, code)
        G = (int(shard) << nbits) | np.arange(1 << nbits, dtype=np.int64)
        val = np.where((G & fixed_mask) == fixed_value, 1.0 + 0j, 0j)
        lens = _ints(group_len, ngroups)
        bits = _ints(group_bits, sum(lens))
        fv = _view(feeds, sum(1 << L for L in lens), np.complex128)
        off = pos = 0
        for L in lens:
            sub = np.zeros_like(G)
            for t in range(L):                       # first listed bit = most significant sub-index bit
                sub |= ((G >> bits[pos + t]) & 1) << (L - 1 - t)
            val = val * fv[off + sub]
            off += 1 << L
            pos += L
        a[:] = val
        self.launches += 2
        return 0

    # ---- gates ----
    def qipb_apply_matrix(self, ctx, state, nbits, code, k, bits, mat, ctrl_mask, diagonal):
        self.log.append("apply_matrix")
        a = _amps(state, nbits, code)
        d = 1 << k
        m = np.array([complex(mat[2 * e], mat[2 * e + 1]) for e in range(d * d)]).reshape(d, d)
        if diagonal:
            m = np.diag(np.diag(m))
        g = BitGate("matrix", tuple(_ints(bits, k)), int(ctrl_mask), m, bool(diagonal) or k == 0)
        a[:] = bitsim.apply_bitgate(a.astype(np.complex128), g, nbits)
        self.launches += 1
        return 0

    def qipb_apply_swap(self, ctx, state, nbits, code, bit_a, bit_b, ctrl_mask):
        self.log.append("apply_swap")
        a = _amps(state, nbits, code)
        a[:] = bitsim.apply_bitgate(a.astype(np.complex128), BitGate("swap", (bit_a, bit_b), int(ctrl_mask)), nbits)
        self.launches += 1
        return 0

    def _fused(self, fn, name, state, nbits, code, ntile, tile_bits, ngates, gates):
        self.log.append(name)
        info = (ctypes.c_int * 16)()
        rc = fn(_addr(state), nbits, code, ntile, tile_bits, ngates, gates, info)
        if rc:
            self.err = self.emul.qipb_emul_last_error()
        self.launches += info[0]
        self.ext_launches += info[7]
        return rc

    def qipb_apply_fused(self, ctx, state, nbits, code, ntile, tile_bits, ngates, gates):
        return self._fused(self.emul.qipb_emul_fused, "apply_fused", state, nbits, code, ntile, tile_bits, ngates, gates)

    def qipb_apply_fused_chunk(self, ctx, state, nbits, code, ntile, tile_bits, ngates, gates, nfix, fix_bits, fix_value):
        self.log.append("apply_fused_chunk")
        info = (ctypes.c_int * 16)()
        before = _amps(state, nbits, code).copy() if self.check_chunk_isolation else None
        rc = self.emul.qipb_emul_fused_chunk(_addr(state), nbits, code, ntile, tile_bits, ngates, gates, info, nfix, fix_bits,
                                             ctypes.c_ulonglong(int(fix_value)))
        if rc:
            self.err = self.emul.qipb_emul_last_error()
            return rc
        if before is not None:
            fb = _ints(fix_bits, nfix)
            mask = sum(1 << b for b in fb)
            idx = np.arange(1 << nbits, dtype=np.int64)
            outside = (idx & mask) != int(fix_value)
            assert np.array_equal(_amps(state, nbits, code)[outside], before[outside]), "a chunked pass touched another chunk"
        self.launches += info[0]
        self.ext_launches += info[7]
        return rc

    def qipb_apply_fused_fill(self, ctx, state, nbits, code, ntile, tile_bits, ngates, gates):
        _amps(state, nbits, code)[:] = np.nan              # the previous content must never matter
        return self._fused(self.emul.qipb_emul_fused_fill, "apply_fused_fill", state, nbits, code, ntile, tile_bits, ngates, gates)

    # ---- func_apply ----
    def qipb_func_xor_u8(self, ctx, state, nbits, code, n1, reg1, n2, reg2, table, x_fixed):
        assert n2 <= 8
        return self.qipb_func_xor(ctx, state, nbits, code, n1, reg1, n2, reg2, table, x_fixed, tab_dtype=np.uint8)

    def qipb_func_xor(self, ctx, state, nbits, code, n1, reg1, n2, reg2, table, x_fixed, tab_dtype=np.int64):
        self.log.append("func_xor")
        a = _amps(state, nbits, code)
        i = np.arange(1 << nbits, dtype=np.int64)
        x = np.full_like(i, int(x_fixed))
        r1 = _ints(reg1, n1)
        for j, b in enumerate(r1):                   # most significant register bit first
            if b >= 0:
                x |= ((i >> b) & 1) << (n1 - 1 - j)
        t = _view(table, 1 << n1, tab_dtype).astype(np.int64)
        y = t[x] & ((1 << n2) - 1)
        flip = np.zeros_like(i)
        r2 = _ints(reg2, n2)
        for j, b in enumerate(r2):
            flip |= ((y >> (n2 - 1 - j)) & 1) << b
        out = a.copy()
        out[i ^ flip] = a[i]
        a[:] = out
        self.launches += 1
        return 0

    # ---- measurement ----
    def qipb_probabilities(self, ctx, state, nbits, code, k, bits, out_bits, fmask, fval, out):
        self.log.append("probabilities")
        a = _amps(state, nbits, code)
        i = np.arange(1 << nbits, dtype=np.int64)
        o = np.zeros_like(i)
        for j in range(k):
            o |= ((i >> int(bits[j])) & 1) << int(out_bits[j])
        keep = (i & int(fmask)) == int(fval)
        p = np.abs(a.astype(np.complex128)) ** 2
        res = np.bincount(o[keep], weights=p[keep], minlength=1 << k)
        _view(out, 1 << k, np.float64)[:] = res
        self.launches += 1
        return 0

    def qipb_collapse(self, ctx, state, nbits, code, mask, want, scale):
        self.log.append("collapse")
        a = _amps(state, nbits, code)
        i = np.arange(1 << nbits, dtype=np.int64)
        a[:] = np.where((i & int(mask)) == int(want), a * scale, 0)
        self.launches += 1
        return 0

    def qipb_reduce(self, ctx, src, dst, nbits, code, mask, want, scale):
        self.log.append("reduce")
        a = _amps(src, nbits, code)
        rest = [b for b in range(nbits) if not (int(mask) >> b) & 1]
        j = np.arange(1 << len(rest), dtype=np.int64)
        i = np.full_like(j, int(want))
        for t, b in enumerate(rest):
            i |= ((j >> t) & 1) << b
        _amps(dst, len(rest), code)[:] = a[i] * scale
        self.launches += 1
        return 0

    def qipb_add_range(self, ctx, state, code, start, count, data):
        self.log.append("add_range")
        dt = np.complex128 if code == qlib.C128 else np.complex64
        a = _view(_addr(state) + int(start) * np.dtype(dt).itemsize, int(count), dt)
        a += _view(data, int(count), dt)
        self.launches += 1
        return 0


class _FakeStream(object):
    cuda_stream = 0

    def synchronize(self):
        pass

    def wait_event(self, ev):
        pass


class _FakeEvent(object):
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def wait(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 0.0


class _FakeCuda(object):
    Event = _FakeEvent

    @staticmethod
    def is_available():
        return True

    @staticmethod
    def current_device():
        return 0

    @staticmethod
    def device(dev):
        return contextlib.nullcontext()

    @staticmethod
    def current_stream(dev=None):
        return _FakeStream()

    @staticmethod
    def Stream(device=None, priority=0):
        return _FakeStream()

    @staticmethod
    def stream(s):
        return contextlib.nullcontext()

    @staticmethod
    def synchronize():
        pass


class _FakeTorch(object):
    """torch with `cuda` replaced and every device mapped to the CPU."""

    def __init__(self, torch):
        self._t = torch
        self.cuda = _FakeCuda()

    def device(self, *a, **k):
        return self._t.device("cpu")

    def __getattr__(self, name):
        return getattr(self._t, name)


def install(monkeypatch):
    """Route qip_b200.backend to the host doubles; returns the HostLib (its .log lists the entry points called)."""
    import torch
    from qip_b200 import backend as be
    L = HostLib()
    monkeypatch.setattr(qlib, "load", lambda: L)
    monkeypatch.setattr(qlib, "check", lambda rc: (_ for _ in ()).throw(qlib.QipbError("libqipb200: " + L.err.decode())) if rc else None)
    fake = _FakeTorch(torch)
    monkeypatch.setattr(be, "_torch", lambda: fake)
    monkeypatch.setattr(be, "_CTX_POOL", {}, raising=False)
    return L


# ======================================================================================================
# Multi-"GPU" double: P virtual ranks = P threads of ONE process running the real ShardedB200Backend (SPMD).
# Shards are host buffers; "peer memory" is simply the other thread's buffer (same address space), so the
# qipb_peer_* kernels are restated in numpy on both buffers; torch.distributed's collectives are thread
# rendezvous.  The product brackets every peer kernel with a barrier (ShardedB200Backend._sync_all), which is
# what makes the concurrent halves of a pair race-free here exactly as on the device.
import threading  # noqa: E402


def _insert_zero(v, p):
    lo = v & ((1 << p) - 1)
    return ((v >> p) << (p + 1)) | lo


class _PeerMixin(object):
    _bufs = {}

    def qipb_dev_alloc(self, ctx, nbytes, out):
        buf = np.zeros(int(nbytes), dtype=np.uint8)
        addr = buf.ctypes.data
        _PeerMixin._bufs[addr] = buf
        out._obj.value = addr
        return 0

    def qipb_dev_free(self, ctx, ptr):
        _PeerMixin._bufs.pop(_addr(ptr), None)
        return 0

    def qipb_ipc_export(self, ctx, ptr, handle_out):
        raw = int(_addr(ptr)).to_bytes(8, "little") + bytes(56)
        ctypes.memmove(handle_out, raw, 64)
        return 0

    def qipb_ipc_open(self, ctx, handle, out):
        out._obj.value = int.from_bytes(bytes(handle)[:8], "little")
        return 0

    def qipb_ipc_close(self, ctx, ptr):
        return 0

    def qipb_peer_swap(self, ctx, local, peer, code, local_off, peer_off, count):
        dt = np.complex128 if code == qlib.C128 else np.complex64
        it = np.dtype(dt).itemsize
        a = _view(_addr(local) + int(local_off) * it, int(count), dt)
        b = _view(_addr(peer) + int(peer_off) * it, int(count), dt)
        tmp = a.copy()
        a[:] = b
        b[:] = tmp
        return 0

    def qipb_peer_swap_bit(self, ctx, local, peer, nbits, code, lbit, my_g, w_begin, count):
        self.log.append("peer_swap_bit")
        a, b = _amps(local, nbits, code), _amps(peer, nbits, code)
        w = np.arange(int(w_begin), int(w_begin) + int(count), dtype=np.int64)
        base = _insert_zero(w, lbit)
        li = base | ((1 - my_g) << lbit)
        pi = base | (my_g << lbit)
        tmp = a[li].copy()
        a[li] = b[pi]
        b[pi] = tmp
        return 0

    def qipb_peer_remap(self, ctx, local, peers, nbits, code, g, lbits, my_value):
        return self.qipb_peer_remap_chunk(ctx, local, peers, nbits, code, g, lbits, my_value, 0, None, 0, 0, name="peer_remap")

    def qipb_peer_remap_chunk(self, ctx, local, peers, nbits, code, g, lbits, my_value, nfix, fix_bits, fix_value, max_ctas,
                              name="peer_remap_chunk"):
        self.log.append(name)
        a = _amps(local, nbits, code)
        lb = _ints(lbits, g)
        fb = _ints(fix_bits, nfix) if nfix else []
        assert not (set(lb) & set(fb))
        half = 1 << (nbits - g - nfix - 1)
        for slot in range((1 << g) - 1):
            bval = my_value ^ (slot + 1)
            peer = _amps(peers[bval], nbits, code)
            lsel = sum(((bval >> t) & 1) << lb[t] for t in range(g)) | int(fix_value)
            psel = sum(((my_value >> t) & 1) << lb[t] for t in range(g)) | int(fix_value)
            base = np.arange(half, dtype=np.int64) + (0 if my_value < bval else half)
            for p in sorted(lb + fb):
                base = _insert_zero(base, p)
            tmp = a[base | lsel].copy()
            a[base | lsel] = peer[base | psel]
            peer[base | psel] = tmp
        return 0

    def qipb_peer_gate1(self, ctx, local, peer, code, off, count, mat, local_is_hi, ctrl_mask):
        self.log.append("peer_gate1")
        dt = np.complex128 if code == qlib.C128 else np.complex64
        it = np.dtype(dt).itemsize
        idx = np.arange(int(off), int(off) + int(count), dtype=np.int64)
        on = (idx & int(ctrl_mask)) == int(ctrl_mask)
        l = _view(_addr(local) + int(off) * it, int(count), dt)
        p = _view(_addr(peer) + int(off) * it, int(count), dt)
        m = [complex(mat[2 * e], mat[2 * e + 1]) for e in range(4)]
        lo, hi = (p, l) if local_is_hi else (l, p)
        r0 = m[0] * lo + m[1] * hi
        r1 = m[2] * lo + m[3] * hi
        lo[on], hi[on] = r0[on].astype(dt), r1[on].astype(dt)
        return 0


class ShardedHostLib(_PeerMixin, HostLib):
    check_chunk_isolation = False


class ThreadDist(object):
    """torch.distributed for P threads of one process (the subset ShardedB200Backend uses)."""

    def __init__(self, P):
        self.P = P
        self.bar = threading.Barrier(P, timeout=120)
        self.local = threading.local()
        self.slots = [None] * P

    def is_initialized(self):
        return True

    def get_rank(self):
        return self.local.rank

    def get_world_size(self):
        return self.P

    def barrier(self):
        self.bar.wait()

    def _exchange(self, obj):
        self.slots[self.local.rank] = obj
        self.bar.wait()
        got = list(self.slots)
        self.bar.wait()
        return got

    def all_gather_object(self, out, obj):
        import pickle
        out[:] = [pickle.loads(b) for b in self._exchange(pickle.dumps(obj))]     # objects travel by value, as over the wire

    def all_reduce(self, t, op=None):
        got = self._exchange(t.clone())
        total = got[0].clone()
        for x in got[1:]:
            total += x
        t.copy_(total)

    def broadcast(self, t, src):
        got = self._exchange(t.clone())
        t.copy_(got[src])

    def all_gather(self, parts, t):
        got = self._exchange(t.clone())
        for dst, src in zip(parts, got):
            dst.copy_(src)


class _PerThreadDict(object):
    """The module-level shard pool of qip_b200.sharded is per PROCESS; virtual ranks are threads, so every thread
    gets its own."""

    def __init__(self):
        self.local = threading.local()

    def _d(self):
        if not hasattr(self.local, "d"):
            self.local.d = {}
        return self.local.d

    def pop(self, *a):
        return self._d().pop(*a)

    def keys(self):
        return self._d().keys()

    def __setitem__(self, k, v):
        self._d()[k] = v

    def __contains__(self, k):
        return k in self._d()

    def __bool__(self):
        return bool(self._d())


class _PerThreadList(object):
    """qip_b200.sharded._ORPHANS (shards of un-closed states) per virtual rank."""

    def __init__(self):
        self.local = threading.local()

    def _l(self):
        if not hasattr(self.local, "l"):
            self.local.l = []
        return self.local.l

    def append(self, v):
        self._l().append(v)

    def pop(self, *a):
        return self._l().pop(*a)

    def __len__(self):
        return len(self._l())

    def __bool__(self):
        return bool(self._l())

    def __len__(self):
        return len(self._d())


def run_virtual_ranks(monkeypatch, P, body):
    """Run body(rank) on P virtual ranks with the real ShardedB200Backend available; re-raises the first failure."""
    import torch
    import torch.distributed as tdist
    from qip_b200 import backend as be
    from qip_b200 import sharded as sh
    L = ShardedHostLib()
    monkeypatch.setattr(qlib, "load", lambda: L)
    monkeypatch.setattr(qlib, "check", lambda rc: (_ for _ in ()).throw(qlib.QipbError("libqipb200: " + L.err.decode())) if rc else None)
    fake = _FakeTorch(torch)
    monkeypatch.setattr(be, "_torch", lambda: fake)
    monkeypatch.setattr(sh, "_torch", lambda: fake)
    monkeypatch.setattr(be, "_CTX_POOL", {}, raising=False)
    monkeypatch.setattr(sh, "_SHARD_POOL", _PerThreadDict())
    monkeypatch.setattr(sh, "_ORPHANS", _PerThreadList())
    td = ThreadDist(P)
    for name in ("is_initialized", "get_rank", "get_world_size", "barrier", "all_gather_object", "all_reduce",
                 "broadcast", "all_gather"):
        monkeypatch.setattr(tdist, name, getattr(td, name))

    def wrap_shard(self, ptr):
        self.ptr = ptr
        dt = np.complex128 if self.amp_bytes == 16 else np.complex64
        self.eng.state = torch.from_numpy(_view(ptr, 1 << self.nl, dt))
        assert self.eng.state.data_ptr() == ptr.value
    monkeypatch.setattr(sh.ShardedB200Backend, "_wrap_shard", wrap_shard)

    errors = [None] * P

    def runner(rank):
        td.local.rank = rank
        try:
            body(rank)
        except BaseException as e:                 # noqa: BLE001 -- reported to the test below
            errors[rank] = e
            td.bar.abort()
    threads = [threading.Thread(target=runner, args=(r,)) for r in range(P)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    real = [e for e in errors if e is not None and not isinstance(e, threading.BrokenBarrierError)]
    if real:
        raise real[0]
    if any(errors):
        raise [e for e in errors if e is not None][0]
    return L
