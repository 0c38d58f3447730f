"""ctypes binding of libqipb200.so (the C ABI in include/qip_b200.h).

The product path has NO CPU fallback: if the CUDA library cannot be loaded, or no sm_100 device is
visible, every entry point raises.  The library is built in-tree (qip_b200/csrc/Makefile, nvcc
-gencode arch=compute_100a,code=sm_100a) so that the .so travels with the repository snapshot.
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.environ.get("QIPB_LIB") or os.path.join(CSRC, "libqipb200.so")     # (QIPB_LIB: an alternative build for A/B runs)

C128, C64 = 0, 1
ERR_UNSUPPORTED = 3          # a valid request this entry point cannot serve; nothing was launched (qipb_apply_fused_fill)
MAX_DENSE_K, MAX_BIG_K, MAX_TILE_BITS, MAX_FUSED_GATES = 4, 10, 12, 280

EXPORTS = [
    "qipb_version", "qipb_last_error", "qipb_create", "qipb_destroy", "qipb_set_stream", "qipb_sync",
    "qipb_launch_count", "qipb_ring_launch_count", "qipb_ext_launch_count", "qipb_dev_alloc", "qipb_dev_free", "qipb_memcpy_h2d", "qipb_memcpy_d2h",
    "qipb_init_basis", "qipb_init_kron", "qipb_apply_matrix", "qipb_apply_swap", "qipb_apply_fused", "qipb_apply_fused_chunk", "qipb_apply_fused_fill",
    "qipb_func_xor", "qipb_func_xor_u8", "qipb_probabilities", "qipb_collapse", "qipb_reduce", "qipb_add_range",
    "qipb_ipc_export", "qipb_ipc_open", "qipb_ipc_close", "qipb_peer_swap", "qipb_peer_swap_bit", "qipb_peer_remap", "qipb_peer_remap_chunk", "qipb_peer_gate1",
]


class Gate(ctypes.Structure):
    """struct qipb_gate"""
    _fields_ = [("k", ctypes.c_int32), ("diagonal", ctypes.c_int32), ("bits", ctypes.c_int32 * 2),
                ("ctrl_mask", ctypes.c_uint64), ("mat", ctypes.c_double * 32)]


class QipbError(RuntimeError):
    pass


def build(verbose=False):
    """Compile libqipb200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise QipbError("building libqipb200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


_lib = None


def load():
    """Load the library and declare every prototype.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QipbError("libqipb200.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, u64, i32p = ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_int)
    dblp = ctypes.POINTER(ctypes.c_double)
    ci = ctypes.c_int
    L.qipb_version.restype = ci
    L.qipb_last_error.restype = ctypes.c_char_p
    L.qipb_create.argtypes = [ci, ctypes.POINTER(vp)]
    L.qipb_destroy.argtypes = [vp]
    L.qipb_set_stream.argtypes = [vp, vp]
    L.qipb_sync.argtypes = [vp]
    L.qipb_launch_count.argtypes = [vp]
    L.qipb_launch_count.restype = ctypes.c_ulonglong
    L.qipb_ring_launch_count.argtypes = [vp]
    L.qipb_ring_launch_count.restype = ctypes.c_ulonglong
    L.qipb_ext_launch_count.argtypes = [vp]
    L.qipb_ext_launch_count.restype = ctypes.c_ulonglong
    L.qipb_dev_alloc.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(vp)]
    L.qipb_dev_free.argtypes = [vp, vp]
    L.qipb_memcpy_h2d.argtypes = [vp, vp, vp, ctypes.c_size_t]
    L.qipb_memcpy_d2h.argtypes = [vp, vp, vp, ctypes.c_size_t]
    L.qipb_init_basis.argtypes = [vp, vp, ci, ci, ctypes.c_longlong]
    L.qipb_init_kron.argtypes = [vp, vp, ci, ci, ci, i32p, i32p, vp, u64, u64, u64]
    L.qipb_apply_matrix.argtypes = [vp, vp, ci, ci, ci, i32p, dblp, u64, ci]
    L.qipb_apply_swap.argtypes = [vp, vp, ci, ci, ci, ci, u64]
    L.qipb_apply_fused.argtypes = [vp, vp, ci, ci, ci, i32p, ci, ctypes.POINTER(Gate)]
    L.qipb_apply_fused_fill.argtypes = [vp, vp, ci, ci, ci, i32p, ci, ctypes.POINTER(Gate)]
    L.qipb_apply_fused_chunk.argtypes = [vp, vp, ci, ci, ci, i32p, ci, ctypes.POINTER(Gate), ci, i32p, u64]
    L.qipb_func_xor.argtypes = [vp, vp, ci, ci, ci, i32p, ci, i32p, vp, u64]
    L.qipb_func_xor_u8.argtypes = [vp, vp, ci, ci, ci, i32p, ci, i32p, vp, u64]
    L.qipb_probabilities.argtypes = [vp, vp, ci, ci, ci, i32p, i32p, u64, u64, vp]
    L.qipb_collapse.argtypes = [vp, vp, ci, ci, u64, u64, ctypes.c_double]
    L.qipb_reduce.argtypes = [vp, vp, vp, ci, ci, u64, u64, ctypes.c_double]
    L.qipb_add_range.argtypes = [vp, vp, ci, u64, u64, vp]
    L.qipb_ipc_export.argtypes = [vp, vp, ctypes.c_char_p]
    L.qipb_ipc_open.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(vp)]
    L.qipb_ipc_close.argtypes = [vp, vp]
    L.qipb_peer_swap.argtypes = [vp, vp, vp, ci, u64, u64, u64]
    L.qipb_peer_swap_bit.argtypes = [vp, vp, vp, ci, ci, ci, ci, u64, u64]
    L.qipb_peer_remap.argtypes = [vp, vp, ctypes.POINTER(vp), ci, ci, ci, i32p, ci]
    L.qipb_peer_remap_chunk.argtypes = [vp, vp, ctypes.POINTER(vp), ci, ci, ci, i32p, ci, ci, i32p, u64, ci]
    L.qipb_peer_gate1.argtypes = [vp, vp, vp, ci, u64, u64, dblp, ci, u64]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("qipb_last_error", "qipb_launch_count", "qipb_ring_launch_count", "qipb_ext_launch_count", "qipb_version"):
            fn.restype = ci
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise QipbError("libqipb200: " + load().qipb_last_error().decode("utf-8", "replace"))


def int_array(values):
    values = [int(v) for v in values]
    return (ctypes.c_int * max(1, len(values)))(*values)


def mat_array(mat):
    import numpy
    a = numpy.ascontiguousarray(mat, dtype=numpy.complex128).reshape(-1)
    buf = (ctypes.c_double * (2 * a.size))()
    ctypes.memmove(buf, a.ctypes.data, 16 * a.size)
    return buf
