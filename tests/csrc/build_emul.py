"""Build recipe of the tests-only host emulator of the fused pass (tests/csrc/fused_emul.cu -> libfused_emul.so)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fused_emul.cu")
SO = os.path.join(HERE, "libfused_emul.so")
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "qip_b200", "csrc")


def build() -> str:
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("fused.cu", "fused_shared.cuh", "common.cuh")]
    deps.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "qip_b200.h"))
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    cmd = ["nvcc", "-O1", "-Xptxas", "-O0", "-Xcicc", "-O0", "-std=c++17", "-DQIPB_ENABLE_PAIRS=1", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
           "-o", SO, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the fused-pass emulator failed:\n" + r.stdout + r.stderr)
    return SO
