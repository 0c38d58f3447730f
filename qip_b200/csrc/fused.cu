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
#define FUSED_MAX_INS 12

struct DevGate {
    unsigned char k;        // target bits in total (0..2)
    unsigned char kin;      // how many of them are tile bits
    unsigned char diag;
    unsigned char nins;     // fixed tile-local positions (in-tile targets + in-tile controls)
    unsigned char tl[2];    // target j (matrix order, 0 = MSB): tile-local position, 0xFF if outside
    unsigned char tg[2];    // target j: position in the state index
    u32 nmask[FUSED_MAX_INS];   // ~((1 << p) - 1) for the fixed positions p, ascending
    u32 in_or;              // tile-local mask of in-tile control bits
    u32 pad;
    u64 out_ctrl;           // state-index mask of controls outside the tile
    double2 m[16];
};

struct FusedArgs {
    int nbits, tb, ngates, lowrun;      // lowrun = number of contiguous low tile bits (0..L-1)
    u64 ntiles;
    unsigned char tbit[16];             // tile-local bit -> state bit, ascending
    DevGate g[QIPB_MAX_FUSED_GATES];
};
static_assert(sizeof(FusedArgs) <= 32764, "FusedArgs must fit in the kernel parameter space");

// ---- mbarrier / bulk-copy (TMA 1-D) primitives ------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, u32 bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, u32 bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");
}

__device__ __forceinline__ u32 expand_local(u32 w, const DevGate &g) {
    for (int q = 0; q < g.nins; ++q) w += (w & g.nmask[q]);      // insert a zero bit at each fixed position
    return w | g.in_or;
}

template <typename A>
__device__ __forceinline__ void run_gate(A *tile, const DevGate &g, u64 base, u32 tsize, int tid) {
    if ((base & g.out_ctrl) != g.out_ctrl) return;             // uniform per tile
    const u32 ngroups = tsize >> g.nins;
    if (g.diag) {
        // effective diagonal over the in-tile targets; targets outside the tile are fixed by `base`
        u32 sel_out = 0;
        for (int j = 0; j < g.k; ++j)
            if (g.tl[j] == 0xFF) sel_out |= (u32)((base >> g.tg[j]) & 1ull) << (g.k - 1 - j);
        const int D = 1 << g.k;
        if (g.kin == 0) {
            const double2 d = g.m[sel_out * D + sel_out];
            if (d.x == 1.0 && d.y == 0.0) return;
            for (u32 w = tid; w < ngroups; w += FUSED_THREADS) {
                const u32 e = expand_local(w, g);
                tile[e] = cmul<A>(d, tile[e]);
            }
        } else if (g.kin == 1) {
            const int j = (g.tl[0] != 0xFF) ? 0 : 1;
            const u32 o1 = 1u << g.tl[j];
            const u32 s1 = 1u << (g.k - 1 - j);
            const double2 d0 = g.m[sel_out * D + sel_out], d1 = g.m[(sel_out | s1) * D + (sel_out | s1)];
            for (u32 w = tid; w < ngroups; w += FUSED_THREADS) {
                const u32 e = expand_local(w, g);
                tile[e] = cmul<A>(d0, tile[e]);
                tile[e | o1] = cmul<A>(d1, tile[e | o1]);
            }
        } else {
            const u32 oh = 1u << g.tl[0], ol = 1u << g.tl[1];
            for (u32 w = tid; w < ngroups; w += FUSED_THREADS) {
                const u32 e = expand_local(w, g);
                tile[e] = cmul<A>(g.m[0], tile[e]);
                tile[e | ol] = cmul<A>(g.m[5], tile[e | ol]);
                tile[e | oh] = cmul<A>(g.m[10], tile[e | oh]);
                tile[e | oh | ol] = cmul<A>(g.m[15], tile[e | oh | ol]);
            }
        }
    } else if (g.k == 1) {
        const u32 o1 = 1u << g.tl[0];
        for (u32 w = tid; w < ngroups; w += FUSED_THREADS) {
            const u32 e = expand_local(w, g);
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
            const u32 e = expand_local(w, g);
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

// BULK: tile staging with cp.async.bulk (TMA 1-D bulk copies, one per contiguous run, completion on
// an mbarrier) instead of LDG/STS through registers.  Requires runs of >= 16 bytes.
template <typename A, bool BULK>
__global__ void __launch_bounds__(FUSED_THREADS) fused_kernel(A *__restrict__ state, const __grid_constant__ FusedArgs f) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    A *tile = reinterpret_cast<A *>(smem_raw);
    const int tid = threadIdx.x;
    const u32 tsize = 1u << f.tb;
    const u32 lowmask = (1u << f.lowrun) - 1u;
    const u32 run_amps = 1u << f.lowrun;
    const u32 nruns = tsize >> f.lowrun;
    const u32 run_bytes = run_amps * (u32)sizeof(A);
    if (BULK) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            fence_proxy_async();
        }
        __syncthreads();
    }
    u32 parity = 0;

    for (u64 t = blockIdx.x; t < f.ntiles; t += gridDim.x) {
        u64 base = t;
        for (int j = 0; j < f.tb; ++j) base = insert_zero(base, f.tbit[j]);

        // ---- stage the tile: runs of 2^lowrun consecutive amplitudes ----
        if (BULK) {
            if (tid < 32) {
                if (tid == 0) mbar_expect_tx(&bar, tsize * (u32)sizeof(A));
                __syncwarp();
                for (u32 r = tid; r < nruns; r += 32) {
                    u64 off = 0;
                    for (int j = f.lowrun; j < f.tb; ++j) off |= (u64)((r >> (j - f.lowrun)) & 1u) << f.tbit[j];
                    bulk_g2s(tile + (size_t)r * run_amps, state + base + off, run_bytes, &bar);
                }
            }
            mbar_wait(&bar, parity);
            parity ^= 1u;
        } else {
            for (u32 e = tid; e < tsize; e += FUSED_THREADS) {
                u64 off = e & lowmask;
                for (int j = f.lowrun; j < f.tb; ++j) off |= (u64)((e >> j) & 1u) << f.tbit[j];
                tile[e] = state[base + off];
            }
            __syncthreads();
        }

        // ---- run the gate list on the tile ----
        for (int gi = 0; gi < f.ngates; ++gi) {
            run_gate<A>(tile, f.g[gi], base, tsize, tid);
            __syncthreads();
        }

        // ---- write the tile back ----
        if (BULK) {
            fence_proxy_async();          // generic-proxy writes -> visible to the async proxy
            __syncthreads();
            if (tid < 32) {
                for (u32 r = tid; r < nruns; r += 32) {
                    u64 off = 0;
                    for (int j = f.lowrun; j < f.tb; ++j) off |= (u64)((r >> (j - f.lowrun)) & 1u) << f.tbit[j];
                    bulk_s2g(state + base + off, tile + (size_t)r * run_amps, run_bytes);
                }
                bulk_commit_wait_read();  // smem may be overwritten once the copies have read it
            }
            __syncthreads();
        } else {
            for (u32 e = tid; e < tsize; e += FUSED_THREADS) {
                u64 off = e & lowmask;
                for (int j = f.lowrun; j < f.tb; ++j) off |= (u64)((e >> j) & 1u) << f.tbit[j];
                state[base + off] = tile[e];
            }
            __syncthreads();
        }
    }
}

template <typename A>
static int launch_fused(qipb_ctx *ctx, A *state, const FusedArgs &f) {
    const size_t smem = sizeof(A) << f.tb;
    const bool bulk = (sizeof(A) << f.lowrun) >= 512 && f.ntiles >= 2;
    int per_sm = (int)((220u * 1024u) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    u64 grid = (u64)ctx->sm_count * per_sm;
    if (grid > f.ntiles) grid = f.ntiles;
    if (bulk) {
        QIPB_CUDA(cudaFuncSetAttribute(fused_kernel<A, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fused_kernel<A, true><<<(unsigned)grid, FUSED_THREADS, smem, ctx->stream>>>(state, f);
    } else {
        QIPB_CUDA(cudaFuncSetAttribute(fused_kernel<A, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fused_kernel<A, false><<<(unsigned)grid, FUSED_THREADS, smem, ctx->stream>>>(state, f);
    }
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
                d.in_or |= 1u << local_of[b];
                fixed_local |= 1ull << local_of[b];
            }
        d.nins = 0;
        for (int j = 0; j < ntile_bits; ++j)
            if ((fixed_local >> j) & 1ull) d.nmask[d.nins++] = ~((1u << j) - 1u);
        const int D = 1 << s.k;
        for (int e = 0; e < D * D; ++e) d.m[e] = make_double2(s.mat[2 * e], s.mat[2 * e + 1]);
    }
    if (dtype == QIPB_C128) return launch_fused<double2>(ctx, (double2 *)state, f);
    if (dtype == QIPB_C64) return launch_fused<float2>(ctx, (float2 *)state, f);
    QIPB_REQUIRE(false, "unknown dtype %d", dtype);
}
