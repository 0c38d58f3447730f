"""Tests-only launcher: bench.py's B200 arm on the host doubles (tests/hostlib.py), so that the bench's own control flow
-- workload generation, timed region, per-kernel accounting, roofline / e2e / cpu_baseline objects, the JSON line -- is
exercised on the CPU tier.  Timings are meaningless here (fake events); only the shape of the output is checked."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import hostlib  # noqa: E402


class _Event(object):
    def __init__(self, enable_timing=False):
        pass

    def record(self):
        pass

    def elapsed_time(self, other):
        return 1.0


def main():
    import torch
    mp = pytest.MonkeyPatch()
    hostlib.install(mp)
    mp.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    mp.setattr(torch.cuda, "mem_get_info", lambda *a, **k: (200 << 30, 200 << 30))
    mp.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    mp.setattr(torch.cuda, "empty_cache", lambda *a, **k: None)
    mp.setattr(torch.cuda, "Event", _Event)
    hostlib._FakeCuda.Event = _Event
    import bench
    sys.exit(bench.main())


if __name__ == "__main__":
    main()
