"""Build recipe for oracle/_ref: the reference's OWN Cython kernels, compiled unmodified.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.

What this does
--------------
Cythonizes the three reference kernel modules *where they lie* under /root/reference
(qip/ext/kronprod.pyx, func_apply.pyx, util.pyx -- the reference's only native code) and
links them with /usr/bin/gcc + OpenMP.  Outputs (generated .c, .o, .so) go ONLY into
oracle/_ref/ (git-ignored, NOT gpurun-ignored, so the built .so files travel to the GPU
box).  No reference source is copied into the repository.

Deviations from the reference's setup.py (none touch kernel code):
  * compiler_directives language_level=2, cpow=True, legacy_implicit_noexcept=True so that the
    Cython-0.29-era sources compile under Cython 3;
  * include_path points at qip/ext so `from util cimport *` (kronprod.pyx:14) resolves;
  * -march=x86-64-v3 instead of -march=native (setup.py:42) because the .so is built in this
    container and executed on the GPU box's host CPU.

Run:  python oracle/build_ref.py        (needs /root/reference; a no-op message otherwise)
"""
import os
import sys
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("QIP_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def built() -> bool:
    ext = os.path.join(OUT, "qip_ref_ext")
    if not os.path.isdir(ext):
        return False
    names = os.listdir(ext)
    return all(any(n.startswith(m + ".") and n.endswith(".so") for n in names)
               for m in ("kronprod", "func_apply", "util"))


def build(force: bool = False) -> bool:
    if built() and not force:
        return True
    src = os.path.join(REF, "qip", "ext")
    if not os.path.isdir(src):
        print("oracle/build_ref.py: %s not present; keeping prebuilt oracle/_ref if any" % src)
        return built()
    import numpy
    from setuptools import Extension
    from setuptools.dist import Distribution
    from Cython.Build import cythonize

    pkg = os.path.join(OUT, "qip_ref_ext")
    bld = os.path.join(OUT, "build")
    os.makedirs(pkg, exist_ok=True)
    os.makedirs(bld, exist_ok=True)
    os.environ.setdefault("CC", "/usr/bin/gcc")
    os.environ.setdefault("LDSHARED", "/usr/bin/gcc -shared")
    flags = ["-O3", "-ffast-math", "-march=x86-64-v3", "-fopenmp"]
    exts = [Extension(m, [os.path.join(src, m + ".pyx")],
                      include_dirs=[numpy.get_include()],
                      libraries=["m"],
                      extra_compile_args=flags, extra_link_args=["-fopenmp"])
            for m in ("util", "kronprod", "func_apply")]
    exts = cythonize(exts, include_path=[src], build_dir=bld, quiet=True,
                     compiler_directives={"language_level": 2, "cpow": True,
                                          "legacy_implicit_noexcept": True})
    dist = Distribution({"name": "qip_ref_ext", "ext_modules": exts})
    cmd = dist.get_command_obj("build_ext")
    cmd.build_lib = pkg
    cmd.build_temp = bld
    cmd.ensure_finalized()
    cmd.run()
    return built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built:", ok)
    sys.exit(0 if ok else 1)
