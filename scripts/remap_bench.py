"""Times repeated global<->local remaps on all GPUs (run under torchrun): first-call vs steady state."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qip_b200.sharded import ShardedB200Backend
from qip_b200 import shardplan as sp

lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rank, world = dist.get_rank(), dist.get_world_size()
G = int(np.log2(world))
nl = int(sys.argv[1]) if len(sys.argv) > 1 else 30
b = ShardedB200Backend.make_state(nl + G, [], [], lazy_layout=False)
b.flush()
torch.cuda.synchronize()
def timed(label, fn):
    dist.barrier(); torch.cuda.synchronize()
    b.profile = []
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    t1 = time.perf_counter(); dist.barrier()
    dt = time.perf_counter() - t0
    kern = sum(e0.elapsed_time(e1) for _, _, e0, e1 in b.profile)
    b.profile = None
    allk = [None] * world
    dist.all_gather_object(allk, (round(kern, 1), round((t1 - t0) * 1e3, 1)))
    if rank == 0: print("%-40s wall %.1f ms; per-rank (kernel ms, local wall ms): %s" % (label, dt * 1e3, allk), flush=True)
pairs = [(nl + t, nl - 1 - t) for t in range(G)]
for it in range(3):
    timed("multi-exchange %d bits, call %d" % (G, it), lambda: b._multi_exchange(sp.MultiExchange(pairs)))
for it in range(2):
    for t in range(G):
        timed("pairwise exchange bit %d, round %d" % (t, it), lambda t=t: b._exchange(sp.Exchange(nl + t, nl - 1 - t)))
b.close()
dist.destroy_process_group()
