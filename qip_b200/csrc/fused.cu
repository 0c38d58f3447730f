// qip_b200/csrc/fused.cu -- fused multi-gate pass: one HBM round trip for a whole list of gates.
//
// The reference sweeps the full state once per op (qip/ext/kronprod.pyx:157-197; QFFT on n qubits
// is n(n+1)/2 + n/2 sweeps, qip/qfft.py:33-39).  Here the state is cut into tiles of 2^TB
// amplitudes spanned by an arbitrary set of TB index bits (the lowest L of them contiguous, so
// every global access is a run of 2^L amplitudes = 2^L * 16 B for complex128).  A CTA stages one
// tile in shared memory (64 KiB at TB = 12, complex128), runs the gate list on it and writes it
// back in place.  Legal in one pass:
//   * dense 1- and 2-qubit gates whose target bits are tile bits,
//   * diagonal gates and phase gates on ANY bits (bits outside the tile are constant per tile and
//     only select which diagonal entry applies),
//   * control bits anywhere (outside the tile they switch the gate on or off per tile).
// Roofline: HBM-bound until the gate list is long enough for shared-memory bandwidth / FP64 to
// take over (about 5 dense gates per pass at 3 CTAs/SM); 2 * sizeof(amp) * 2^nbits bytes per pass.
#include "common.cuh"
#include "../../include/qip_b200.h"

namespace qipb {

#define FUSED_THREADS 256
#define FUSED_MAX_INS 16

struct DevGate {
    unsigned char k;        // target bits in total (0..2)
    unsigned char kin;      // how many of them are tile bits
    unsigned char diag;
    unsigned char nins;     // fixed tile-local positions (in-tile targets + in-tile controls)
    unsigned char ins[FUSED_MAX_INS];   // ascending tile-local positions
    unsigned char tl[2];    // target j (matrix order, 0 = MSB): tile-local position, 0xFF if outside
    unsigned char tg[2];    // target j: position in the state index
    unsigned char pad[8];
    u64 in_or;              // tile-local mask of in-tile control bits
    u64 out_ctrl;           // state-index mask of controls outside the tile
    double2 m[16];
};

struct FusedArgs {
    int nbits, tb, ngates, lowrun;      // lowrun = number of contiguous low tile bits (0..L-1)
    u64 ntiles;
    unsigned char tbit[16];             // tile-local bit -> state bit, ascending
    DevGate g[QIPB_MAX_FUSED_GATES];
};

template <typename A>
__global__ void __launch_bounds__(FUSED_THREADS) fused_kernel(A *__restrict__ state, const __grid_constant__ FusedArgs f) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    A *tile = reinterpret_cast<A *>(smem_raw);
    const int tid = threadIdx.x;
    const u32 tsize = 1u << f.tb;
    const u32 lowmask = (1u << f.lowrun) - 1u;

    for (u64 t = blockIdx.x; t < f.ntiles; t += gridDim.x) {
        u64 base = t;
        for (int j = 0; j < f.tb; ++j) base = insert_zero(base, f.tbit[j]);

        // ---- stage the tile: runs of 2^lowrun consecutive amplitudes ----
        for (u32 e = tid; e < tsize; e += FUSED_THREADS) {
            u64 off = e & lowmask;
            for (int j = f.lowrun; j < f.tb; ++j) off |= (u64)((e >> j) & 1u) << f.tbit[j];
            tile[e] = state[base + off];
        }
        __syncthreads();

        // ---- run the gate list on the tile ----
        for (int gi = 0; gi < f.ngates; ++gi) {
            const DevGate &g = f.g[gi];
            if ((base & g.out_ctrl) == g.out_ctrl) {         // uniform per tile
                const u32 ngroups = tsize >> g.nins;
                if (g.diag) {
                    // effective diagonal over the in-tile targets; outside targets are fixed by base
                    u32 sel_out = 0;
                    for (int j = 0; j < g.k; ++j)
                        if (g.tl[j] == 0xFF) sel_out |= (u32)((base >> g.tg[j]) & 1ull) << (g.k - 1 - j);
                    const int D = 1 << g.k;
                    if (g.kin == 0) {
                        const double2 d = g.m[sel_out * D + sel_out];
                        if (!(d.x == 1.0 && d.y == 0.0))
                            for (u32 w = tid; w < ngroups; w += FUSED_THREADS) {
                                u32 e = w;
                                for (int q = 0; q < g.nins; ++q) e = (u32)insert_zero(e, g.ins[q]);
                                e |= (u32)g.in_or;
                                tile[e] = cmul<A>(d, tile[e]);
                            }
                    } else {
                        for (u32 w = tid; w < ngroups; w += FUSED_THREADS) {
                            u32 e = w;
                            for (int q = 0; q < g.nins; ++q) e = (u32)insert_zero(e, g.ins[q]);
                            e |= (u32)g.in_or;
                            const int nin = 1 << g.kin;
                            for (int c = 0; c < nin; ++c) {
                                // spread c over the in-tile targets (matrix order)
                                u32 sel = sel_out, eo = e;
                                int bitpos = g.kin - 1;
                                for (int j = 0; j < g.k; ++j)
                                    if (g.tl[j] != 0xFF) {
                                        const u32 b = (c >> bitpos) & 1u;
                                        sel |= b << (g.k - 1 - j);
                                        eo |= b << g.tl[j];
                                        --bitpos;
                                    }
                                const double2 d = g.m[sel * D + sel];
                                if (!(d.x == 1.0 && d.y == 0.0)) tile[eo] = cmul<A>(d, tile[eo]);
                            }
                        }
                    }
                } else if (g.k == 1) {
                    const u32 o1 = 1u << g.tl[0];
                    for (u32 w = tid; w < ngroups; w += FUSED_THREADS) {
                        u32 e = w;
                        for (int q = 0; q < g.nins; ++q) e = (u32)insert_zero(e, g.ins[q]);
                        e |= (u32)g.in_or;
                        const A a0 = tile[e], a1 = tile[e | o1];
                        A r0 = cmul<A>(g.m[0], a0);
                        cfma<A>(r0, g.m[1], a1);
                        A r1 = cmul<A>(g.m[2], a0);
                        cfma<A>(r1, g.m[3], a1);
                        tile[e] = r0;
                        tile[e | o1] = r1;
                    }
                } else {   // dense k == 2
                    const u32 oh = 1u << g.tl[0], ol = 1u << g.tl[1];
                    for (u32 w = tid; w < ngroups; w += FUSED_THREADS) {
                        u32 e = w;
                        for (int q = 0; q < g.nins; ++q) e = (u32)insert_zero(e, g.ins[q]);
                        e |= (u32)g.in_or;
                        const u32 idx[4] = {e, e | ol, e | oh, e | oh | ol};
                        A a[4], r[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) a[j] = tile[idx[j]];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            r[i] = cmul<A>(g.m[i * 4], a[0]);
#pragma unroll
                            for (int j = 1; j < 4; ++j) cfma<A>(r[i], g.m[i * 4 + j], a[j]);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) tile[idx[i]] = r[i];
                    }
                }
            }
            __syncthreads();
        }

        // ---- write the tile back ----
        for (u32 e = tid; e < tsize; e += FUSED_THREADS) {
            u64 off = e & lowmask;
            for (int j = f.lowrun; j < f.tb; ++j) off |= (u64)((e >> j) & 1u) << f.tbit[j];
            state[base + off] = tile[e];
        }
        __syncthreads();
    }
}

template <typename A>
static int launch_fused(qipb_ctx *ctx, A *state, const FusedArgs &f) {
    const size_t smem = sizeof(A) << f.tb;
    QIPB_CUDA(cudaFuncSetAttribute(fused_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((220u * 1024u) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    u64 grid = (u64)ctx->sm_count * per_sm;
    if (grid > f.ntiles) grid = f.ntiles;
    fused_kernel<A><<<(unsigned)grid, FUSED_THREADS, smem, ctx->stream>>>(state, f);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}

}  // namespace qipb

using namespace qipb;

extern "C" int qipb_apply_fused(qipb_ctx *ctx, void *state, int nbits, int dtype, int ntile_bits, const int *tile_bits,
                                int ngates, const qipb_gate *gates) {
    QIPB_REQUIRE(ctx && state && gates && tile_bits, "null argument");
    QIPB_REQUIRE(nbits >= 0 && nbits <= 40, "nbits %d unsupported", nbits);
    QIPB_REQUIRE(ntile_bits >= 0 && ntile_bits <= QIPB_MAX_TILE_BITS && ntile_bits <= nbits, "tile bits %d unsupported", ntile_bits);
    QIPB_REQUIRE(ngates >= 1 && ngates <= QIPB_MAX_FUSED_GATES, "ngates %d unsupported (1..%d)", ngates, QIPB_MAX_FUSED_GATES);
    QIPB_CUDA(cudaSetDevice(ctx->device));
    static thread_local FusedArgs f;    // ~29 KiB: keep it off the stack
    memset(&f, 0, sizeof(f));
    f.nbits = nbits;
    f.tb = ntile_bits;
    f.ngates = ngates;
    f.ntiles = 1ull << (nbits - ntile_bits);
    int local_of[64];
    for (int b = 0; b < 64; ++b) local_of[b] = -1;
    u64 tmask = 0;
    for (int j = 0; j < ntile_bits; ++j) {
        QIPB_REQUIRE(tile_bits[j] >= 0 && tile_bits[j] < nbits, "tile bit %d out of range", tile_bits[j]);
        QIPB_REQUIRE(j == 0 || tile_bits[j] > tile_bits[j - 1], "tile bits must be ascending and distinct");
        f.tbit[j] = (unsigned char)tile_bits[j];
        local_of[tile_bits[j]] = j;
        tmask |= 1ull << tile_bits[j];
    }
    f.lowrun = 0;
    while (f.lowrun < ntile_bits && tile_bits[f.lowrun] == f.lowrun) f.lowrun++;
    for (int gi = 0; gi < ngates; ++gi) {
        const qipb_gate &s = gates[gi];
        DevGate &d = f.g[gi];
        QIPB_REQUIRE(s.k >= 0 && s.k <= 2, "fused gate %d: k=%d unsupported", gi, s.k);
        d.k = (unsigned char)s.k;
        d.diag = (unsigned char)(s.diagonal != 0 || s.k == 0);
        u64 tgt = 0, fixed_local = 0;
        d.kin = 0;
        for (int j = 0; j < s.k; ++j) {
            const int b = s.bits[j];
            QIPB_REQUIRE(b >= 0 && b < nbits && !((tgt >> b) & 1ull), "fused gate %d: bad target bit %d", gi, b);
            tgt |= 1ull << b;
            d.tg[j] = (unsigned char)b;
            if (local_of[b] >= 0) {
                d.tl[j] = (unsigned char)local_of[b];
                fixed_local |= 1ull << local_of[b];
                d.kin++;
            } else {
                QIPB_REQUIRE(d.diag, "fused gate %d: non-diagonal target bit %d is not a tile bit", gi, b);
                d.tl[j] = 0xFF;
            }
        }
        QIPB_REQUIRE((s.ctrl_mask & tgt) == 0, "fused gate %d: control mask overlaps targets", gi);
        QIPB_REQUIRE(nbits == 64 || (s.ctrl_mask >> nbits) == 0, "fused gate %d: control outside local bits", gi);
        d.out_ctrl = s.ctrl_mask & ~tmask;
        d.in_or = 0;
        for (int b = 0; b < nbits; ++b)
            if (((s.ctrl_mask & tmask) >> b) & 1ull) {
                d.in_or |= 1ull << local_of[b];
                fixed_local |= 1ull << local_of[b];
            }
        d.nins = 0;
        for (int j = 0; j < ntile_bits; ++j)
            if ((fixed_local >> j) & 1ull) d.ins[d.nins++] = (unsigned char)j;
        const int D = 1 << s.k;
        for (int e = 0; e < D * D; ++e) d.m[e] = make_double2(s.mat[2 * e], s.mat[2 * e + 1]);
    }
    if (dtype == QIPB_C128) return launch_fused<double2>(ctx, (double2 *)state, f);
    if (dtype == QIPB_C64) return launch_fused<float2>(ctx, (float2 *)state, f);
    QIPB_REQUIRE(false, "unknown dtype %d", dtype);
}
