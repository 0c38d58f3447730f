"""Worker of tests/test_sharding.py::test_two_process_gloo_shards_match_oracle: one shard per process,
gloo backend, the same shardplan actions the CUDA executor runs, exchanges via send/recv."""
import sys

import numpy as np
import torch
import torch.distributed as dist

import shardsim
from qip_b200 import ops
from qip_b200 import shardplan as sp
from qip_b200.circuits import layered_stream, qfft_stream


def main(out_path):
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    gbits = int(np.log2(world))
    n = 8
    nl = n - gbits
    rng = np.random.default_rng(42)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    shard = np.array(psi[rank << nl:(rank + 1) << nl])
    gates = []
    for mats in list(layered_stream(n, 2, 7)) + list(qfft_stream(n)):
        for g in ops.decode_mats(mats, n):
            s = ops.simplify(g)
            if s is not None:
                gates.append(s)
    lay = sp.Layout(n, gbits)
    actions = sp.schedule(gates, lay) + sp.canonicalise(lay)

    def trade(send_arr, partner):
        recv = torch.empty(len(send_arr), dtype=torch.complex128)
        snd = torch.from_numpy(np.ascontiguousarray(send_arr))
        if rank < partner:
            dist.send(snd, partner)
            dist.recv(recv, partner)
        else:
            dist.recv(recv, partner)
            dist.send(snd, partner)
        return recv.numpy()

    expanded = []
    for a in actions:
        if isinstance(a, sp.MultiExchange):
            expanded.extend(sp.Exchange(g, l) for g, l in a.pairs)
        else:
            expanded.append(a)
    for a in expanded:
        if isinstance(a, sp.Exchange):
            gb = a.gpos - nl
            my_g = (rank >> gb) & 1
            partner = rank ^ (1 << gb)
            mine, _ = shardsim._swap_sets(nl, a.lpos, my_g)
            shard[mine] = trade(shard[mine], partner)
        elif isinstance(a, sp.PeerGate1):
            gb = a.gpos - nl
            my_g = (rank >> gb) & 1
            partner = rank ^ (1 << gb)
            other = trade(shard, partner)
            cg = a.ctrl_mask >> nl
            if (rank & cg) == cg:
                lowmask = (1 << nl) - 1
                idx = np.arange(1 << nl, dtype=np.int64)
                on = (idx & (a.ctrl_mask & lowmask)) == (a.ctrl_mask & lowmask)
                lo, hi = (shard, other) if my_g == 0 else (other, shard)
                new = a.mat[my_g, 0] * lo + a.mat[my_g, 1] * hi
                shard = np.where(on, new, shard)
        else:
            shard = shardsim.apply_local(shard, rank, a, nl)
    parts = [torch.empty(1 << nl, dtype=torch.complex128) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(shard)))
    if rank == 0:
        np.save(out_path, np.concatenate([p.numpy() for p in parts]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
