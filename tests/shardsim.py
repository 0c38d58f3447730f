"""Tests-only numpy executors for qip_b200.shardplan actions.

VirtualShards : P logical shards in ONE process (checks the sharding logic without GPUs).
GlooShard     : one shard per process, exchanges over torch.distributed (gloo) send/recv -- the
                world_size-2 CPU test of the N>1 path.
Both apply exactly the actions the CUDA executor (qip_b200/sharded.py) applies, rank by rank.
"""
import numpy as np

import bitsim
from qip_b200 import shardplan as sp
from qip_b200.ops import BitGate


def _swap_sets(nl, lpos, my_g):
    idx = np.arange(1 << nl, dtype=np.int64)
    bit = (idx >> lpos) & 1
    mine = idx[bit == (1 - my_g)]      # my amplitudes whose local bit differs from my rank bit
    theirs = idx[bit == my_g]          # the partner's amplitudes (in ITS shard) at the mirrored bit
    return mine, theirs


def apply_local(shard, rank, action, nl):
    if isinstance(action, sp.Apply):
        bg = sp.lower_for_rank(action.gate, nl, rank)
        if bg is not None:
            return bitsim.apply_bitgate(shard, bg, nl)
        return shard
    if isinstance(action, sp.LocalSwap):
        return bitsim.apply_bitgate(shard, BitGate("swap", (action.a, action.b)), nl)
    raise TypeError(action)


class VirtualShards(object):
    def __init__(self, state, gbits):
        self.n = int(np.log2(len(state)))
        self.G = gbits
        self.nl = self.n - gbits
        self.P = 1 << gbits
        self.shards = [np.array(state[r << self.nl:(r + 1) << self.nl]) for r in range(self.P)]
        self.exchanges = 0
        self.bytes_out = 0

    def run(self, actions):
        nl = self.nl
        expanded = []
        for a in actions:
            if isinstance(a, sp.MultiExchange):       # same permutation as its pairwise exchanges in sequence
                expanded.extend(sp.Exchange(g, l) for g, l in a.pairs)
                self.multi_exchanges = getattr(self, "multi_exchanges", 0) + 1
            else:
                expanded.append(a)
        for a in expanded:
            if isinstance(a, sp.Exchange):
                gb = a.gpos - nl
                for r in range(self.P):
                    if (r >> gb) & 1:
                        continue
                    p = r | (1 << gb)
                    mine, theirs = _swap_sets(nl, a.lpos, 0)     # rank r has rank bit 0
                    # r's amplitudes with local bit 1  <->  p's amplitudes with local bit 0
                    tmp = self.shards[r][mine].copy()
                    self.shards[r][mine] = self.shards[p][theirs]
                    self.shards[p][theirs] = tmp
                self.exchanges += 1
                self.bytes_out += 16 * (1 << (nl - 1))
            elif isinstance(a, sp.PeerGate1):
                gb = a.gpos - nl
                lowmask = (1 << nl) - 1
                for r in range(self.P):
                    if (r >> gb) & 1:
                        continue
                    p = r | (1 << gb)
                    cg = a.ctrl_mask >> nl
                    if (r & cg) != cg:          # controls on other rank bits (gb itself is never a control)
                        continue
                    idx = np.arange(1 << nl, dtype=np.int64)
                    on = (idx & (a.ctrl_mask & lowmask)) == (a.ctrl_mask & lowmask)
                    lo, hi = self.shards[r], self.shards[p]
                    nlo = np.where(on, a.mat[0, 0] * lo + a.mat[0, 1] * hi, lo)
                    nhi = np.where(on, a.mat[1, 0] * lo + a.mat[1, 1] * hi, hi)
                    self.shards[r], self.shards[p] = nlo, nhi
                self.exchanges += 1
                self.bytes_out += 16 * (1 << (nl - 1))
            else:
                for r in range(self.P):
                    self.shards[r] = apply_local(self.shards[r], r, a, nl)

    def run_planned(self, actions, tile_bits=5, min_low_bits=2):
        """Like run(), but rank-local gates go through the product's rank-local pipeline exactly as
        ShardedB200Backend._execute does it: shardplan.compile_program resolves the batches between
        exchanges per rank, merges them (ops.merge_bitgates, incl. the clustering of lone diagonal gates)
        and plans them into passes (ops.plan_passes); the programs are then executed in lockstep."""
        from qip_b200.ops import merge_bitgates, plan_passes
        nl = self.nl

        def plan_local(batch):
            return plan_passes(merge_bitgates(batch, 2), nl, 16, tile_bits=min(tile_bits, nl),
                               min_low_bits=min(min_low_bits, nl))

        self.run_programs([sp.compile_program(actions, nl, r, plan_local) for r in range(self.P)])

    def run_programs(self, programs):
        """Execute one rank-local program per virtual shard (what ShardedB200Backend._run_program does on
        each GPU).  Every program holds the same sequence of moves (exchanges / peer gates are
        rank-independent); ("local", passes) steps in between may be missing on ranks with nothing to do."""
        nl = self.nl
        cursors = [0] * self.P
        while True:
            for r in range(self.P):
                prog = programs[r]
                while cursors[r] < len(prog) and isinstance(prog[cursors[r]], tuple):
                    self.shards[r] = bitsim.run_passes(self.shards[r], prog[cursors[r]][1], nl)
                    cursors[r] += 1
            done = [cursors[r] >= len(programs[r]) for r in range(self.P)]
            if all(done):
                return
            assert not any(done), "programs disagree on the number of moves"
            move = programs[0][cursors[0]]
            for r in range(self.P):
                other = programs[r][cursors[r]]
                assert type(other) is type(move) and vars(other).keys() == vars(move).keys()
                cursors[r] += 1
            self.run([move])

    def gather(self):
        return np.concatenate(self.shards)
