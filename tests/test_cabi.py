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


def test_shipped_library_carries_the_blackwell_paths_it_claims():
    """SASS of the built library (cuobjdump, no GPU needed): the fused pass moves its tiles with TMA bulk copies on
    mbarriers (UBLKCP + SYNCS), dense K >= 5 gates multiply FP64 tensor tiles (DMMA) fed by cp.async (LDGSTS), and the
    code is sm_100a only (the mnemonics of /opt/skills/guides/B200_PROFILING.md)."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not installed")
    lib.build()
    arch = subprocess.run([exe, "-lelf", lib.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
    assert "sm_100a" in arch and not re.search(r"sm_(?!100a)\d+", arch), arch
    csrc = os.path.dirname(lib.LIB_PATH)
    sass = {}
    for obj in ("gates.o", "fused.o"):
        path = os.path.join(csrc, obj)
        assert os.path.exists(path), path
        sass[obj] = subprocess.run([exe, "-sass", path], capture_output=True, text=True, timeout=600).stdout
    assert sass["gates.o"].count("DMMA.8x8x4") >= 48 and "LDGSTS" in sass["gates.o"]
    assert "big_gate_mma_kernel" in sass["gates.o"]
    for mnemonic in ("UBLKCP.S.G", "UBLKCP.G.S", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT"):
        assert mnemonic in sass["fused.o"], mnemonic
