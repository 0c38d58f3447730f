"""Tests-only numpy executor for the planner's BitGate / Pass form (checks HOST logic on the CPU
tier; the CUDA kernels are checked on the GPU tier)."""
import numpy as np


def apply_bitgate(state, g, nbits):
    N = state.shape[0]
    idx = np.arange(N, dtype=np.int64)
    on = (idx & g.ctrl_mask) == g.ctrl_mask
    if g.kind == "swap":
        a, b = g.bits
        ba = (idx >> a) & 1
        bb = (idx >> b) & 1
        src = idx ^ np.where(ba != bb, (1 << a) | (1 << b), 0)
        out = state.copy()
        out[on] = state[src[on]]
        return out
    k = g.k
    if k == 0:
        out = state.copy()
        out[on] = state[on] * g.mat[0, 0]
        return out
    sub = np.zeros(N, dtype=np.int64)
    for j, b in enumerate(g.bits):
        sub |= ((idx >> b) & 1) << (k - 1 - j)
    tmask = 0
    for b in g.bits:
        tmask |= 1 << b
    base = idx & ~tmask
    out = state.copy()
    acc = np.zeros(N, dtype=np.complex128)
    for c in range(1 << k):
        off = 0
        for j, b in enumerate(g.bits):
            if (c >> (k - 1 - j)) & 1:
                off |= 1 << b
        acc += g.mat[sub, c] * state[base | off]
    out[on] = acc[on]
    return out


def run_passes(state, passes, nbits):
    for p in passes:
        if p.fused:
            tile = set(p.tile_bits)
            assert len(p.tile_bits) == len(tile) and list(p.tile_bits) == sorted(tile)
            for g in p.gates:
                if g.kind == "swap" or not (g.diagonal or g.k == 0):
                    assert set(g.bits) <= tile, "non-diagonal target outside the tile"
                assert g.k <= 2
        else:
            assert len(p.gates) == 1
        for g in p.gates:
            state = apply_bitgate(state, g, nbits)
    return state
