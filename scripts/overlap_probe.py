#!/usr/bin/env python
"""Timeline of one sharded step (rank 0): every launch of the chunk pipeline with its start / end on the device clock.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/overlap_probe.py [--workload layered|qft]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="layered")
    ap.add_argument("--qubits", type=int, default=33)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from qip_b200.circuits import layered_stream, qfft_stream
    from qip_b200.sharded import ShardedB200Backend
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    rank, world = dist.get_rank(), dist.get_world_size()
    n = args.qubits + int(np.log2(world))
    b = ShardedB200Backend.make_state(n, [], [])
    steps = [list(qfft_stream(n)) if args.workload == "qft" else list(layered_stream(n, 1, 33 + s)) for s in range(3 + args.steps)]
    for s in range(3):
        for m in steps[s]:
            b.kronselect_dot(m)
        b.flush()
    torch.cuda.synchronize()
    dist.barrier()
    base = torch.cuda.Event(enable_timing=True)
    base.record()
    b.profile = []
    for s in range(3, 3 + args.steps):
        for m in steps[s]:
            b.kronselect_dot(m)
        b.flush()
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("%s n=%d world=%d: %.1f ms for %d steps; stats %s" % (args.workload, n, world, base.elapsed_time(end), args.steps, b.stats))
        for name, nbytes, e0, e1 in b.profile:
            t0, t1 = base.elapsed_time(e0), base.elapsed_time(e1)
            print("%8.2f -> %8.2f  (%6.2f ms)  %-28s %6.1f GB  %5.0f GB/s" % (t0, t1, t1 - t0, name, nbytes / 1e9, nbytes / (t1 - t0) / 1e6))
    b.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
