"""CPU tier: the C-ABI library loads and exports every symbol include/qip_b200.h declares.
No compute calls here (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

from qip_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "qip_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qipb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    lib.build()
    L = lib.load()
    assert L.qipb_version() == 100


def test_every_declared_symbol_is_exported():
    L = lib.load()
    names = header_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(L, name), name
    assert sorted(lib.EXPORTS) == names


def test_gate_struct_matches_header_layout():
    # struct qipb_gate { int32 k; int32 diagonal; int32 bits[2]; uint64 ctrl_mask; double mat[32]; }
    assert ctypes.sizeof(lib.Gate) == 4 + 4 + 8 + 8 + 32 * 8
    assert lib.Gate.ctrl_mask.offset == 16 and lib.Gate.mat.offset == 24


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from qip_b200 import B200Backend
    with pytest.raises(Exception):
        B200Backend.make_state(2, [], [], statetype=np.complex128)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "qip_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "qip_oracle" not in text, f
