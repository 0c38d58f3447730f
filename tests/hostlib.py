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
    def __init__(self):
        self.emul = ctypes.CDLL(build_emul.build())
        argt = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                ctypes.POINTER(qlib.Gate), ctypes.POINTER(ctypes.c_int)]
        self.emul.qipb_emul_fused.argtypes = argt
        self.emul.qipb_emul_fused_fill.argtypes = argt
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
        info = (ctypes.c_int * 8)()
        rc = fn(_addr(state), nbits, code, ntile, tile_bits, ngates, gates, info)
        if rc:
            self.err = self.emul.qipb_emul_last_error()
        self.launches += info[0]
        self.ext_launches += info[7]
        return rc

    def qipb_apply_fused(self, ctx, state, nbits, code, ntile, tile_bits, ngates, gates):
        return self._fused(self.emul.qipb_emul_fused, "apply_fused", state, nbits, code, ntile, tile_bits, ngates, gates)

    def qipb_apply_fused_fill(self, ctx, state, nbits, code, ntile, tile_bits, ngates, gates):
        _amps(state, nbits, code)[:] = np.nan              # the previous content must never matter
        return self._fused(self.emul.qipb_emul_fused_fill, "apply_fused_fill", state, nbits, code, ntile, tile_bits, ngates, gates)

    # ---- func_apply ----
    def qipb_func_xor(self, ctx, state, nbits, code, n1, reg1, n2, reg2, table, x_fixed):
        self.log.append("func_xor")
        a = _amps(state, nbits, code)
        i = np.arange(1 << nbits, dtype=np.int64)
        x = np.full_like(i, int(x_fixed))
        r1 = _ints(reg1, n1)
        for j, b in enumerate(r1):                   # most significant register bit first
            if b >= 0:
                x |= ((i >> b) & 1) << (n1 - 1 - j)
        t = _view(table, 1 << n1, np.int64)
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


class _FakeEvent(object):
    def __init__(self, enable_timing=False):
        pass

    def record(self):
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
