"""CPU tier, build container only (needs /root/reference): the UNMODIFIED reference front-end
(Qubit / H / C(...) / Swap / Rm / QFFT / F / Measure graphs, qip/pipeline.py run()) drives the product's
host pipeline through the reference's own plug-in hook `backend_constructor=` (qip/pipeline.py:95-96,133).

The device kernels cannot run here (no GPU), so the state container is the tests-only numpy executor
behind the PRODUCT's decode -> simplify -> merge -> plan code (tests/test_host_logic.py
PlannedHostBackend).  What this pins: the boundary -- every call, argument form and return value the
reference's graph nodes use (int32 ndarrays, CMat/SwapMat objects of the reference's own classes, `n` passed
as input_offset by FOp, `[:]` on the result of measure_probabilities ...) is accepted by the product's
host side, on the reference's own test-suite."""
import sys
import unittest

import pytest

from oracle.ref_loader import have_ref_ext, have_reference_tree

pytestmark = pytest.mark.skipif(not (have_reference_tree() and have_ref_ext()),
                                reason="needs /root/reference and oracle/_ref (build container only)")


def _suite():
    import os
    from oracle.ref_loader import REF, import_reference_qip
    import_reference_qip()
    tdir = os.path.join(REF, "tests")
    if tdir not in sys.path:
        sys.path.insert(0, tdir)
    names = []
    for mod in ("qiptest", "qubit_util_test", "qfttest"):
        s = unittest.defaultTestLoader.loadTestsFromName(mod)

        def walk(x):
            for t in x:
                if isinstance(t, unittest.TestSuite):
                    yield from walk(t)
                else:
                    yield t
        names += [(mod, t._testMethodName, t) for t in walk(s)]
    return names


def test_reference_suite_runs_on_the_product_host_pipeline():
    cases = _suite()
    import qip.pipeline
    from test_host_logic import Dense4HostBackend, TileHostBackend
    assert len(cases) == 37
    original = qip.pipeline.CythonBackend.make_state
    try:
        for backend in (TileHostBackend, Dense4HostBackend):
            qip.pipeline.CythonBackend.make_state = staticmethod(
                lambda n, groups, feeds, statetype=None, **kw: backend.make_state(n, groups, feeds))
            for mod, name, t in cases:
                res = unittest.TestResult()
                t.run(res)
                assert res.wasSuccessful(), (backend.__name__, mod, name, res.failures, res.errors)
    finally:
        qip.pipeline.CythonBackend.make_state = original


def test_backend_constructor_kwarg_and_carriers_of_the_reference():
    import numpy as np
    from oracle.ref_loader import import_reference_qip
    import_reference_qip()
    from qip.operators import C, H, Swap
    from qip.pipeline import run
    from qip.qip import Measure, Qubit
    from test_host_logic import TileHostBackend
    q1, q2, q3 = Qubit(n=1), Qubit(n=5), Qubit(n=5)        # README CSwap circuit (README.md:8-39)
    c1, c2, c3 = C(Swap)(H(q1), q2, q3)
    m = Measure(H(c1))
    s2, s3 = np.zeros(32), np.zeros(32)
    s2[0] = s3[1] = 1.0
    out_ref, cl_ref = run(m, c2, c3, feed={q1: [1.0, 0.0], q2: s2, q3: s3})
    out, cl = run(m, c2, c3, feed={q1: [1.0, 0.0], q2: s2, q3: s3}, backend_constructor=TileHostBackend.make_state)
    assert abs(cl[m][1] - 0.5) < 1e-12 and abs(cl_ref[m][1] - 0.5) < 1e-12
    assert abs(np.linalg.norm(out) - 1.0) < 1e-12
