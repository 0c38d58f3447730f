"""Loader for oracle/_ref (the reference's own Cython kernels, built by oracle/build_ref.py).

TEST INFRASTRUCTURE ONLY -- never imported by qip_b200/.

Two entry points:
  load_ref_ext()          -> (kronprod, func_apply) compiled reference modules.  Works wherever
                             oracle/_ref/qip_ref_ext/*.so exists (this container AND the GPU box).
  import_reference_qip()  -> the reference's python package `qip`, wired to those modules.  Works
                             only where /root/reference exists (this container); used by
                             tests/golden/make_golden.py to record golden op streams and by the CPU
                             tests that cross-check the oracle.  Applies the three import-time shims
                             the reference needs on Python 3.12 / numpy 2.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
EXT_DIR = os.path.join(HERE, "_ref", "qip_ref_ext")
REF = os.environ.get("QIP_REFERENCE", "/root/reference")

_cache = {}


def have_ref_ext() -> bool:
    return os.path.isdir(EXT_DIR) and any(f.startswith("kronprod.") for f in os.listdir(EXT_DIR))


def have_reference_tree() -> bool:
    return os.path.isdir(os.path.join(REF, "qip"))


def _shims():
    import collections
    import collections.abc
    import numpy
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable      # qip/util.py:153, qip/qubit_util.py:23
    if not hasattr(numpy, "int"):
        numpy.int = numpy.int64                               # qip/ext/kronprod.pyx:139
    if not hasattr(numpy, "complex_"):
        numpy.complex_ = numpy.complex128                     # qip/distributed/proto/conversion.py:73


def load_ref_ext():
    if "ext" in _cache:
        return _cache["ext"]
    if not have_ref_ext():
        raise ImportError("oracle/_ref not built: run `python oracle/build_ref.py` where /root/reference exists")
    _shims()
    # The modules were compiled as top-level `util`, `kronprod`, `func_apply` (that is how
    # `from util cimport *` resolves at run time); import them with EXT_DIR first on sys.path
    # and make sure an unrelated top-level `util` is not picked up.
    saved = {k: sys.modules.pop(k) for k in ("util", "kronprod", "func_apply") if k in sys.modules}
    sys.path.insert(0, EXT_DIR)
    try:
        util = importlib.import_module("util")
        kronprod = importlib.import_module("kronprod")
        func_apply = importlib.import_module("func_apply")
    finally:
        sys.path.remove(EXT_DIR)
    for k, v in saved.items():
        if not getattr(v, "__file__", "").startswith(EXT_DIR):
            # restore the foreign module under its name; ours stay reachable through _cache
            sys.modules[k] = v
    _cache["ext"] = (kronprod, func_apply)
    _cache["util"] = util
    return _cache["ext"]


def import_reference_qip():
    """Import the unmodified reference python package from /root/reference (this container only)."""
    if "qip" in _cache:
        return _cache["qip"]
    if not have_reference_tree():
        raise ImportError("%s does not exist (the GPU box has no reference tree)" % REF)
    kronprod, func_apply = load_ref_ext()
    _shims()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import types
    # matplotlib is absent; tests/qfttest.py:7 and the examples import pyplot at module top.
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mpl = types.ModuleType("matplotlib")
            mpl.pyplot = types.ModuleType("matplotlib.pyplot")
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = mpl.pyplot
    import qip.ext  # the (read-only) package directory; holds no compiled modules
    sys.modules["qip.ext.kronprod"] = kronprod
    sys.modules["qip.ext.func_apply"] = func_apply
    sys.modules["qip.ext.util"] = _cache["util"]
    qip.ext.kronprod = kronprod
    qip.ext.func_apply = func_apply
    import qip
    import qip.backend  # noqa: F401
    import qip.pipeline  # noqa: F401
    import qip.operators  # noqa: F401
    import qip.qfft  # noqa: F401
    _cache["qip"] = qip
    return qip
