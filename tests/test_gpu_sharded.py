"""GPU tier, multi-GPU: ShardedB200Backend (NVLink peer kernels + NCCL) vs the CPU oracle, one rank
per GPU.  Skipped on single-GPU boxes; run with `gpurun --gpus 2 -- python -m pytest tests -m gpu -k sharded`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_backend_matches_oracle_on_all_gpus():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 1 << (ngpu.bit_length() - 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0 and "SHARDED PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
