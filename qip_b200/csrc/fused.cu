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
#include <stdlib.h>
#include <complex>
#include <vector>
#include "common.cuh"
#include "../../include/qip_b200.h"

namespace qipb {

#define FUSED_THREADS 256
#define FUSED_MAX_INS 12
#define FUSED_OUT_CELLS 4
#define FUSED_LO_BITS 6          // tile-local bits 0..5 form the 'lo' table cell, the rest the 'hi' cell

struct DevGate {
    unsigned char k;        // target bits in total (0..2)
    unsigned char kin;      // how many of them are tile bits
    unsigned char diag;
    unsigned char nins;     // fixed tile-local positions (dense: in-tile targets + controls; diagonal: controls below sb)
    unsigned char tl[2];    // target j (matrix order, 0 = MSB): tile-local position, 0xFF if outside
    unsigned char tg[2];    // target j: position in the state index
    unsigned char ins[FUSED_MAX_INS];   // the fixed positions, ascending
    u32 in_or;              // tile-local mask of in-tile control bits (diagonal gates: those below the warp-slice bits)
    u32 hi_need;            // diagonal gates: control bits at or above the warp-slice bits, as a mask over the warp id
    u32 coef;               // offset of this gate's coefficients in FusedArgs::pool (stage: in the table buffer)
    u64 out_ctrl;           // state-index mask of controls outside the tile
    // diag == 2 ("stage"): a run of diagonal gates folded into per-cell phase tables
    unsigned char nout;     // number of cells made of bits outside the tile
    unsigned char cn[FUSED_OUT_CELLS];        // bits per outside cell
    unsigned char cb[FUSED_OUT_CELLS][7];     // their state-index positions, ascending
    unsigned char blockwide; // diagonal op that is alone between two block syncs: spread over all 256 threads
    unsigned char pad2[2];
};

#define FUSED_MAX_GATES QIPB_MAX_FUSED_GATES
#define FUSED_POOL QIPB_MAX_FUSED_COEFS      // complex coefficients: dense 2q = 16, dense 1q = 4, diagonal k = 2^k

struct FusedArgs {
    int nbits, tb, ngates, lowrun;      // lowrun = number of contiguous low tile bits (0..L-1)
    int sb, wb;                         // diagonal runs: a warp owns 2^sb consecutive tile elements, 2^wb warps work
    int npool, pad;
    u64 ntiles;
    unsigned char tbit[16];             // tile-local bit -> state bit, ascending
    const double2 *tables;              // stage tables (global memory)
    DevGate g[FUSED_MAX_GATES];
    double2 pool[FUSED_POOL];
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

// Pin a coefficient in a register: the value comes out of an asm volatile, so ptxas cannot
// re-materialise it with another constant-bank load inside the inner loop.
__device__ __forceinline__ double2 pin(const double2 v) {
    double2 r;
    asm volatile("mov.f64 %0, %2;\n\tmov.f64 %1, %3;" : "=d"(r.x), "=d"(r.y) : "d"(v.x), "d"(v.y));
    return r;
}
// Same with the first four masks pre-loaded into registers (identity masks = 0 beyond nins).
struct LocalIns { u32 m0, m1, m2, m3, orv; int more; };
__device__ __forceinline__ LocalIns load_ins(const DevGate &g) {
    LocalIns l;
    l.m0 = g.nins > 0 ? ~((1u << g.ins[0]) - 1u) : 0u;     // w += w & mask inserts a zero bit at that position
    l.m1 = g.nins > 1 ? ~((1u << g.ins[1]) - 1u) : 0u;
    l.m2 = g.nins > 2 ? ~((1u << g.ins[2]) - 1u) : 0u;
    l.m3 = g.nins > 3 ? ~((1u << g.ins[3]) - 1u) : 0u;
    l.orv = g.in_or;
    l.more = g.nins > 4;
    return l;
}
// Block-wide variant for diagonal ops: the control bits at or above the warp-slice bits are fixed
// positions too (they are masks over the warp id in the warp-sliced mode).
__device__ __forceinline__ LocalIns load_ins_block(const DevGate &g, int sb) {
    u32 m[8];
    int n = 0;
    for (int q = 0; q < g.nins && n < 8; ++q) m[n++] = ~((1u << g.ins[q]) - 1u);
    for (int b = 0; b < 3; ++b)
        if ((g.hi_need >> b) & 1u) m[n++] = ~((1u << (sb + b)) - 1u);
    LocalIns l;
    l.m0 = n > 0 ? m[0] : 0u;
    l.m1 = n > 1 ? m[1] : 0u;
    l.m2 = n > 2 ? m[2] : 0u;
    l.m3 = n > 3 ? m[3] : 0u;
    l.orv = g.in_or | (g.hi_need << sb);
    l.more = 0;
    return l;
}
__device__ __forceinline__ int popc3(u32 v) { return __popc(v & 7u); }

__device__ __forceinline__ u32 expand_fast(u32 w, const LocalIns &l, const DevGate &g) {
    w += (w & l.m0);
    w += (w & l.m1);
    w += (w & l.m2);
    w += (w & l.m3);
    if (l.more)
        for (int q = 4; q < g.nins; ++q) w += (w & ~((1u << g.ins[q]) - 1u));
    return w | l.orv;
}

template <typename A, bool NOPIN>
__device__ __forceinline__ void run_gate(A *tile, const DevGate &g, const double2 *M, u64 base, u32 tsize, int tid) {
    if ((base & g.out_ctrl) != g.out_ctrl) return;             // uniform per tile
    const u32 ngroups = tsize >> g.nins;
    const LocalIns li = load_ins(g);
    if (g.diag) {
        return;   // diagonal gates run warp-sliced (run_diag)
    } else if (g.k == 1) {
        const u32 o1 = 1u << g.tl[0];
        // coefficients hoisted into registers once per gate (the gate index is dynamic, so reading
        // g.m inside the loop would be an LDC per use and every DFMA would wait on it)
        const double2 m0 = NOPIN ? M[0] : pin(M[0]), m1 = NOPIN ? M[1] : pin(M[1]), m2 = NOPIN ? M[2] : pin(M[2]), m3 = NOPIN ? M[3] : pin(M[3]);
        for (u32 w = tid; w < ngroups; w += 2 * FUSED_THREADS) {
            const u32 w2 = w + FUSED_THREADS;
            const bool two = w2 < ngroups;
            const u32 e = expand_fast(w, li, g), e2 = expand_fast(two ? w2 : w, li, g);
            const A a0 = tile[e], a1 = tile[e | o1], b0 = tile[e2], b1 = tile[e2 | o1];
            A r0 = cmul<A>(m0, a0), r1 = cmul<A>(m2, a0), s0 = cmul<A>(m0, b0), s1 = cmul<A>(m2, b0);
            cfma<A>(r0, m1, a1);
            cfma<A>(r1, m3, a1);
            cfma<A>(s0, m1, b1);
            cfma<A>(s1, m3, b1);
            tile[e] = r0;
            tile[e | o1] = r1;
            if (two) {
                tile[e2] = s0;
                tile[e2 | o1] = s1;
            }
        }
    } else {   // dense k == 2
        const u32 oh = 1u << g.tl[0], ol = 1u << g.tl[1];
        double2 m[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) m[i] = NOPIN ? M[i] : pin(M[i]);   // NOPIN: ptxas may re-load them inside the loop
        // software pipeline: the next group's four LDS are issued before this group's 64 DFMA, so the
        // shared-memory latency hides behind the FP64 work (the compiler cannot hoist them itself:
        // the stores of this iteration may alias the loads of the next as far as it can tell)
        u32 w = tid;
        A a[4];
        u32 idx[4];
        if (w < ngroups) {
            const u32 e = expand_fast(w, li, g);
            idx[0] = e; idx[1] = e | ol; idx[2] = e | oh; idx[3] = e | oh | ol;
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] = tile[idx[j]];
        }
        while (w < ngroups) {
            const u32 wn = w + FUSED_THREADS;
            A an[4];
            u32 idn[4];
            if (wn < ngroups) {
                const u32 e = expand_fast(wn, li, g);
                idn[0] = e; idn[1] = e | ol; idn[2] = e | oh; idn[3] = e | oh | ol;
#pragma unroll
                for (int j = 0; j < 4; ++j) an[j] = tile[idn[j]];
            }
            A r[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                r[i] = cmul<A>(m[i * 4], a[0]);
#pragma unroll
                for (int j = 1; j < 4; ++j) cfma<A>(r[i], m[i * 4 + j], a[j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) tile[idx[i]] = r[i];
#pragma unroll
            for (int j = 0; j < 4; ++j) { a[j] = an[j]; idx[j] = idn[j]; }
            w = wn;
        }
    }
}

// Diagonal / phase gate, warp-sliced: warp `wid` owns tile elements [wid << sb, (wid+1) << sb), so a
// run of consecutive diagonal gates needs only __syncwarp() between gates -- every element is always
// touched by the same warp.  Elements satisfying the in-tile controls are enumerated (control bits
// inserted as ones); the diagonal entry is selected per element from the target bits (bits outside
// the tile are constant per tile and come from `base`).
template <typename A>
__device__ __forceinline__ void run_diag(A *tile, const DevGate &g, const double2 *M, u64 base, int sb, int wb, int wid, int lane) {
    if ((base & g.out_ctrl) != g.out_ctrl) return;             // uniform per tile
    if (wid >= (1 << wb) || ((u32)wid & g.hi_need) != g.hi_need) return;   // uniform per warp
    const LocalIns li = load_ins(g);
    const u32 n = (1u << sb) >> g.nins;
    const u32 wbase = (u32)wid << sb;
    u32 sel_out = 0;
    for (int j = 0; j < g.k; ++j)
        if (g.tl[j] == 0xFF) sel_out |= (u32)((base >> g.tg[j]) & 1ull) << (g.k - 1 - j);
    if (g.kin == 0) {
        const double2 d = pin(M[sel_out]);
        if (d.x == 1.0 && d.y == 0.0) return;
        u32 x = lane;
        for (; x + 96 < n; x += 128) {          // four independent elements per iteration
            u32 e[4];
            A v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { e[q] = wbase | expand_fast(x + 32 * q, li, g); v[q] = tile[e[q]]; }
#pragma unroll
            for (int q = 0; q < 4; ++q) tile[e[q]] = cmul<A>(d, v[q]);
        }
        for (; x < n; x += 32) {
            const u32 e = wbase | expand_fast(x, li, g);
            tile[e] = cmul<A>(d, tile[e]);
        }
    } else if (g.kin == 1) {
        const int j = (g.tl[0] != 0xFF) ? 0 : 1;
        const int tb1 = g.tl[j];
        const u32 s1 = 1u << (g.k - 1 - j);
        const double2 d0 = pin(M[sel_out]), d1 = pin(M[sel_out | s1]);
        for (u32 x = lane; x < n; x += 32) {
            const u32 e = wbase | expand_fast(x, li, g);
            const bool hi = (e >> tb1) & 1u;
            const double2 d = make_double2(hi ? d1.x : d0.x, hi ? d1.y : d0.y);
            tile[e] = cmul<A>(d, tile[e]);
        }
    } else {
        const int t0 = g.tl[0], t1 = g.tl[1];
        const double2 q0 = pin(M[0]), q1 = pin(M[1]), q2 = pin(M[2]), q3 = pin(M[3]);
        for (u32 x = lane; x < n; x += 32) {
            const u32 e = wbase | expand_fast(x, li, g);
            const bool b1 = (e >> t0) & 1u, b0 = (e >> t1) & 1u;
            const double2 lo = make_double2(b0 ? q1.x : q0.x, b0 ? q1.y : q0.y);
            const double2 hi = make_double2(b0 ? q3.x : q2.x, b0 ? q3.y : q2.y);
            const double2 d = make_double2(b1 ? hi.x : lo.x, b1 ? hi.y : lo.y);
            tile[e] = cmul<A>(d, tile[e]);
        }
    }
}

// A "stage": a run of diagonal gates that share the control bits in_or/hi_need/out_ctrl and whose
// remaining bits each fall into ONE cell of the index (tile-lo, tile-hi, or a group of <= 7 bits
// outside the tile).  Their product is a phase  S * T_hi[e >> 6] * T_lo[e & 63]  per element, S being
// the product of the outside-cell tables at this tile's base index: one shared-memory sweep and three
// complex multiplies replace the whole run (a QFT stage is ~n controlled phases).
template <typename A>
__device__ __forceinline__ void run_stage(A *tile, const DevGate &g, const double2 *T, u64 base, int tb, int sb, int wb,
                                          int wid, int lane) {
    if ((base & g.out_ctrl) != g.out_ctrl) return;
    const bool bw = g.blockwide && g.nins + popc3(g.hi_need) <= 4;
    if (!bw && (wid >= (1 << wb) || ((u32)wid & g.hi_need) != g.hi_need)) return;
    const int lo = tb < FUSED_LO_BITS ? tb : FUSED_LO_BITS;
    const u32 nlo = 1u << lo, nhi = 1u << (tb - lo);
    double2 S = make_double2(1.0, 0.0);
    const double2 *To = T + nlo + nhi;
    for (int c = 0; c < g.nout; ++c) {
        u32 idx = 0;
        for (int j = 0; j < g.cn[c]; ++j) idx |= (u32)((base >> g.cb[c][j]) & 1ull) << j;
        S = cmul<double2>(S, To[idx]);
        To += 1u << g.cn[c];
    }
    const LocalIns li = bw ? load_ins_block(g, sb) : load_ins(g);
    const u32 n = bw ? ((1u << tb) >> (g.nins + popc3(g.hi_need))) : ((1u << sb) >> g.nins);
    const u32 wbase = bw ? 0u : ((u32)wid << sb);
    const u32 step = bw ? FUSED_THREADS : 32u;
    u32 x = bw ? (u32)(wid * 32 + lane) : (u32)lane;
    for (; x + 3 * step < n; x += 4 * step) {   // four independent elements per iteration
        u32 e[4];
        A v[4];
        double2 th[4], tl[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            e[q] = wbase | expand_fast(x + step * q, li, g);
            v[q] = tile[e[q]];
            th[q] = T[nlo + (e[q] >> lo)];
            tl[q] = T[e[q] & (nlo - 1u)];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double2 ph = cmul<double2>(S, th[q]);
            ph = cmul<double2>(ph, tl[q]);
            tile[e[q]] = cmul<A>(ph, v[q]);
        }
    }
    for (; x < n; x += step) {
        const u32 e = wbase | expand_fast(x, li, g);
        double2 ph = cmul<double2>(S, T[nlo + (e >> lo)]);
        ph = cmul<double2>(ph, T[e & (nlo - 1u)]);
        tile[e] = cmul<A>(ph, tile[e]);
    }
}

// BULK: tile staging with cp.async.bulk (TMA 1-D bulk copies, one per contiguous run, completion on
// an mbarrier) instead of LDG/STS through registers.  Requires runs of >= 16 bytes.
// VAR 0: gate descriptors read from the kernel-parameter (constant) bank, 3 CTAs/SM.
// VAR 1: descriptors + coefficients copied once per CTA into shared memory after the tile, 3 CTAs/SM.
// VAR 2: as 1, dense coefficients pinned in registers (more registers: 2 CTAs/SM).
template <typename A, bool BULK, int VAR>
__global__ void __launch_bounds__(FUSED_THREADS, (VAR == 2 ? 2 : 3)) fused_kernel(A *__restrict__ state, const __grid_constant__ FusedArgs f) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    A *tile = reinterpret_cast<A *>(smem_raw);
    const int tid = threadIdx.x;
    const u32 tsize = 1u << f.tb;
    const DevGate *gates = f.g;
    const double2 *pool = f.pool;
    if (VAR >= 1) {
        DevGate *sg = reinterpret_cast<DevGate *>(smem_raw + ((size_t)sizeof(A) << f.tb));
        double2 *sp = reinterpret_cast<double2 *>(sg + ((f.ngates + 1) & ~1));
        const u32 *src = reinterpret_cast<const u32 *>(f.g);
        u32 *dst = reinterpret_cast<u32 *>(sg);
        for (u32 i = tid; i < (u32)f.ngates * (sizeof(DevGate) / 4); i += FUSED_THREADS) dst[i] = src[i];
        const u32 *psrc = reinterpret_cast<const u32 *>(f.pool);
        u32 *pdst = reinterpret_cast<u32 *>(sp);
        for (u32 i = tid; i < (u32)f.npool * 4; i += FUSED_THREADS) pdst[i] = psrc[i];
        gates = sg;
        pool = sp;
        __syncthreads();
    }
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
        bool prev_diag = false;
        for (int gi = 0; gi < f.ngates; ++gi) {
            const DevGate &g = gates[gi];
            if (g.diag == 2) {
                if (g.blockwide && prev_diag) __syncthreads();
                run_stage<A>(tile, g, f.tables + g.coef, base, f.tb, f.sb, f.wb, tid >> 5, tid & 31);
                if (g.blockwide) {
                    __syncthreads();
                    prev_diag = false;
                } else {
                    __syncwarp();
                    prev_diag = true;
                }
            } else if (g.diag) {
                run_diag<A>(tile, g, pool + g.coef, base, f.sb, f.wb, tid >> 5, tid & 31);
                __syncwarp();
                prev_diag = true;
            } else {
                if (prev_diag) __syncthreads();
                run_gate<A, VAR != 2>(tile, g, pool + g.coef, base, tsize, tid);
                __syncthreads();
                prev_diag = false;
            }
        }
        if (prev_diag) __syncthreads();

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

template <typename A, bool BULK, int VAR>
static int launch_fused_v(qipb_ctx *ctx, A *state, const FusedArgs &f) {
    const size_t tile_bytes = sizeof(A) << f.tb;
    const size_t desc_bytes = VAR >= 1 ? (size_t)((f.ngates + 1) & ~1) * sizeof(DevGate) + (size_t)f.npool * sizeof(double2) : 0;
    const size_t smem = tile_bytes + desc_bytes;
    int per_sm = (int)((220u * 1024u) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    u64 grid = (u64)ctx->sm_count * per_sm;
    if (grid > f.ntiles) grid = f.ntiles;
    QIPB_CUDA(cudaFuncSetAttribute(fused_kernel<A, BULK, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fused_kernel<A, BULK, VAR><<<(unsigned)grid, FUSED_THREADS, smem, ctx->stream>>>(state, f);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}

static int fused_variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("QIPB_FUSED_VARIANT");     // tuning knob for profiling runs
        v = e ? atoi(e) : 1;
        if (v < 0 || v > 2) v = 1;
    }
    return v;
}

template <typename A>
static int launch_fused(qipb_ctx *ctx, A *state, const FusedArgs &f) {
    const bool bulk = (sizeof(A) << f.lowrun) >= 512 && f.ntiles >= 2;
    const int v = fused_variant();
    if (bulk) {
        if (v == 0) return launch_fused_v<A, true, 0>(ctx, state, f);
        if (v == 2) return launch_fused_v<A, true, 2>(ctx, state, f);
        return launch_fused_v<A, true, 1>(ctx, state, f);
    }
    if (v == 0) return launch_fused_v<A, false, 0>(ctx, state, f);
    if (v == 2) return launch_fused_v<A, false, 2>(ctx, state, f);
    return launch_fused_v<A, false, 1>(ctx, state, f);
}


// ---- host side: folding runs of diagonal gates into stages -------------------------------------
typedef std::complex<double> cplx;

struct Op {
    bool stage;
    int gate;                 // !stage: index into the caller's gate list
    u64 common;               // stage: control bits shared by every gate of the stage
    u32 tab_off;              // stage: offset of its tables in the table buffer
    int nout;
    std::vector<int> cells[FUSED_OUT_CELLS];
};

static bool stages_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("QIPB_FUSED_STAGES");          // tuning knob for profiling runs
        v = e ? atoi(e) : 1;
    }
    return v != 0;
}

// cell of a state bit: 0 = tile-lo, 1 = tile-hi, 2.. = groups of 7 outside bits (ascending)
struct CellMap {
    int cell_of[64];
    int idx_in_cell[64];
    std::vector<int> bits[2 + FUSED_OUT_CELLS];
    int ncells;
};

static int residual_cell(const qipb_gate &g, u64 common, const CellMap &cm) {
    // -1: no residual bits; -2: spans several cells; else the cell id
    int cell = -1;
    u64 rest = g.ctrl_mask & ~common;
    for (int j = 0; j < g.k; ++j) rest |= 1ull << g.bits[j];
    for (int b = 0; b < 64; ++b)
        if ((rest >> b) & 1ull) {
            const int c = cm.cell_of[b];
            if (c < 0) return -2;
            if (cell == -1) cell = c;
            else if (cell != c) return -2;
        }
    return cell;
}

static void build_stages(const qipb_gate *gates, const std::vector<int> &run, int nbits, int tb, const int *local_of,
                         u64 tmask, std::vector<Op> &ops, std::vector<cplx> &tables) {
    CellMap cm;
    for (int b = 0; b < 64; ++b) { cm.cell_of[b] = -1; cm.idx_in_cell[b] = 0; }
    const int lo = tb < FUSED_LO_BITS ? tb : FUSED_LO_BITS;
    int nout_bits = 0;
    for (int b = 0; b < nbits; ++b) {
        int c;
        if (local_of[b] >= 0) c = local_of[b] < lo ? 0 : 1;
        else c = 2 + nout_bits++ / 7;
        if (c >= 2 + FUSED_OUT_CELLS) continue;               // beyond the supported outside cells: not table-izable
        cm.cell_of[b] = c;
        cm.idx_in_cell[b] = (int)cm.bits[c].size();
        cm.bits[c].push_back(b);
    }
    // tile cells are indexed by tile-local position (lo: e & 63, hi: e >> 6)
    for (int b = 0; b < nbits; ++b)
        if (local_of[b] >= 0) cm.idx_in_cell[b] = local_of[b] < lo ? local_of[b] : local_of[b] - lo;
    cm.ncells = 2 + (nout_bits + 6) / 7;
    if (cm.ncells > 2 + FUSED_OUT_CELLS) cm.ncells = 2 + FUSED_OUT_CELLS;

    size_t pos = 0;
    while (pos < run.size()) {
        // grow a stage greedily: the common control set may only shrink, every member must stay single-cell
        u64 common = gates[run[pos]].ctrl_mask;
        size_t end = pos + 1;
        {
            // a single gate is "single-cell" w.r.t. its own controls iff its targets are
            if (residual_cell(gates[run[pos]], common, cm) == -2) common = 0;
        }
        while (end < run.size()) {
            const u64 nc = common & gates[run[end]].ctrl_mask;
            bool ok = true;
            for (size_t t = pos; t <= end && ok; ++t) ok = residual_cell(gates[run[t]], nc, cm) != -2;
            if (!ok) break;
            common = nc;
            ++end;
        }
        if (end - pos < 3 || residual_cell(gates[run[pos]], common, cm) == -2) {
            // too short to pay for tables (or not table-izable at all): keep the gate as it is
            Op o;
            o.stage = false;
            o.gate = run[pos];
            ops.push_back(o);
            ++pos;
            continue;
        }
        // tables: T_lo, T_hi, then the outside cells that are actually used
        std::vector<std::vector<cplx>> T(cm.ncells);
        std::vector<bool> used(cm.ncells, false);
        used[0] = used[1] = true;
        for (int c = 0; c < cm.ncells; ++c) T[c].assign((size_t)1 << (c == 0 ? lo : c == 1 ? tb - lo : (int)cm.bits[c].size()), cplx(1.0, 0.0));
        for (size_t t = pos; t < end; ++t) {
            const qipb_gate &g = gates[run[t]];
            int c = residual_cell(g, common, cm);
            if (c == -1) c = 0;
            used[c] = true;
            const int D = 1 << g.k;
            const u64 rc = g.ctrl_mask & ~common;
            for (size_t v = 0; v < T[c].size(); ++v) {
                bool on = true;
                for (int b = 0; b < nbits && on; ++b)
                    if ((rc >> b) & 1ull) on = (v >> cm.idx_in_cell[b]) & 1u;
                if (!on) continue;
                int sel = 0;
                for (int j = 0; j < g.k; ++j)
                    if ((v >> cm.idx_in_cell[g.bits[j]]) & 1u) sel |= 1 << (g.k - 1 - j);
                T[c][v] *= cplx(g.mat[2 * (sel * D + sel)], g.mat[2 * (sel * D + sel) + 1]);
            }
        }
        Op o;
        o.stage = true;
        o.gate = -1;
        o.common = common;
        o.tab_off = (u32)tables.size();
        o.nout = 0;
        tables.insert(tables.end(), T[0].begin(), T[0].end());
        tables.insert(tables.end(), T[1].begin(), T[1].end());
        for (int c = 2; c < cm.ncells; ++c)
            if (used[c]) {
                o.cells[o.nout] = cm.bits[c];
                tables.insert(tables.end(), T[c].begin(), T[c].end());
                o.nout++;
            }
        ops.push_back(o);
        pos = end;
    }
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
    static thread_local FusedArgs f;    // ~30 KiB: keep it off the stack
    memset(&f, 0, sizeof(f));
    u32 pool_used = 0;
    f.nbits = nbits;
    f.tb = ntile_bits;
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
    f.wb = ntile_bits - 5 < 0 ? 0 : (ntile_bits - 5 > 3 ? 3 : ntile_bits - 5);   // 2^wb of the 8 warps, >= 32 elements each
    f.sb = ntile_bits - f.wb;
    // ---- pass 1: validate, and fold runs of diagonal gates into stages ----
    std::vector<Op> ops;
    std::vector<cplx> tables;
    {
        std::vector<int> run;
        auto flush_run = [&]() {
            if (!run.empty()) build_stages(gates, run, nbits, ntile_bits, local_of, tmask, ops, tables);
            run.clear();
        };
        for (int gi = 0; gi < ngates; ++gi) {
            const qipb_gate &s = gates[gi];
            QIPB_REQUIRE(s.k >= 0 && s.k <= 2, "fused gate %d: k=%d unsupported", gi, s.k);
            u64 tgt = 0;
            const bool diag = s.diagonal != 0 || s.k == 0;
            for (int j = 0; j < s.k; ++j) {
                const int b = s.bits[j];
                QIPB_REQUIRE(b >= 0 && b < nbits && !((tgt >> b) & 1ull), "fused gate %d: bad target bit %d", gi, b);
                tgt |= 1ull << b;
                QIPB_REQUIRE(diag || local_of[b] >= 0, "fused gate %d: non-diagonal target bit %d is not a tile bit", gi, b);
            }
            QIPB_REQUIRE((s.ctrl_mask & tgt) == 0, "fused gate %d: control mask overlaps targets", gi);
            QIPB_REQUIRE(nbits == 64 || (s.ctrl_mask >> nbits) == 0, "fused gate %d: control outside local bits", gi);
            if (diag && stages_enabled()) {
                run.push_back(gi);
            } else {
                flush_run();
                Op o;
                o.gate = gi;
                o.stage = false;
                ops.push_back(o);
            }
        }
        flush_run();
    }
    QIPB_REQUIRE((int)ops.size() <= FUSED_MAX_GATES, "fused pass needs %d device ops (max %d)", (int)ops.size(), FUSED_MAX_GATES);

    // ---- pass 2: device descriptors ----
    f.ngates = (int)ops.size();
    for (size_t oi = 0; oi < ops.size(); ++oi) {
        const Op &o = ops[oi];
        DevGate &d = f.g[oi];
        const u64 ctrl_mask = o.stage ? o.common : gates[o.gate].ctrl_mask;
        u64 fixed_local = 0;
        if (o.stage) {
            d.k = 0;
            d.kin = 0;
            d.diag = 2;
            d.coef = o.tab_off;
            d.nout = (unsigned char)o.nout;
            for (int c = 0; c < o.nout; ++c) {
                d.cn[c] = (unsigned char)o.cells[c].size();
                for (size_t j = 0; j < o.cells[c].size(); ++j) d.cb[c][j] = (unsigned char)o.cells[c][j];
            }
        } else {
            const qipb_gate &s = gates[o.gate];
            d.k = (unsigned char)s.k;
            d.diag = (unsigned char)(s.diagonal != 0 || s.k == 0);
            d.kin = 0;
            for (int j = 0; j < s.k; ++j) {
                const int b = s.bits[j];
                d.tg[j] = (unsigned char)b;
                if (local_of[b] >= 0) {
                    d.tl[j] = (unsigned char)local_of[b];
                    fixed_local |= 1ull << local_of[b];
                    d.kin++;
                } else {
                    d.tl[j] = 0xFF;
                }
            }
            const int D = 1 << s.k;
            const int ncoef = d.diag ? D : D * D;
            QIPB_REQUIRE(pool_used + ncoef <= FUSED_POOL, "fused pass needs more than %d matrix coefficients", FUSED_POOL);
            d.coef = pool_used;
            for (int e = 0; e < ncoef; ++e) {
                const int src = d.diag ? e * D + e : e;
                f.pool[pool_used + e] = make_double2(s.mat[2 * src], s.mat[2 * src + 1]);
            }
            pool_used += ncoef;
        }
        d.out_ctrl = ctrl_mask & ~tmask;
        d.in_or = 0;
        d.hi_need = 0;
        if (d.diag) fixed_local = 0;      // diagonal gates enumerate elements: only controls are fixed
        for (int b = 0; b < nbits; ++b)
            if (((ctrl_mask & tmask) >> b) & 1ull) {
                const int lb = local_of[b];
                if (d.diag && lb >= f.sb) {
                    d.hi_need |= 1u << (lb - f.sb);
                } else {
                    d.in_or |= 1u << lb;
                    fixed_local |= 1ull << lb;
                }
            }
        d.nins = 0;
        for (int j = 0; j < ntile_bits; ++j)
            if ((fixed_local >> j) & 1ull) d.ins[d.nins++] = (unsigned char)j;
    }
    f.npool = (int)pool_used;
    for (int oi = 0; oi < f.ngates; ++oi)
        if (f.g[oi].diag == 2) {
            const bool prev_dense = oi == 0 || f.g[oi - 1].diag == 0;
            const bool next_dense = oi + 1 == f.ngates || f.g[oi + 1].diag == 0;
            f.g[oi].blockwide = (prev_dense && next_dense) ? 1 : 0;
        }

    // ---- stage tables: pinned staging ring -> device buffer, stream ordered ----
    f.tables = nullptr;
    if (!tables.empty()) {
        const size_t need = tables.size();
        if (ctx->tab_cap < need) {
            QIPB_CUDA(cudaStreamSynchronize(ctx->stream));
            size_t cap = 1u << 16;
            while (cap < need) cap <<= 1;
            if (ctx->tab_dev) QIPB_CUDA(cudaFree(ctx->tab_dev));
            ctx->tab_dev = nullptr;
            QIPB_CUDA(cudaMalloc(&ctx->tab_dev, cap * sizeof(double2)));
            for (int i = 0; i < 4; ++i) {
                if (ctx->tab_host[i]) QIPB_CUDA(cudaFreeHost(ctx->tab_host[i]));
                ctx->tab_host[i] = nullptr;
                QIPB_CUDA(cudaMallocHost(&ctx->tab_host[i], cap * sizeof(double2)));
                if (!ctx->tab_ev[i]) QIPB_CUDA(cudaEventCreateWithFlags(&ctx->tab_ev[i], cudaEventDisableTiming));
            }
            ctx->tab_cap = cap;
        }
        const int slot = ctx->tab_slot;
        ctx->tab_slot = (slot + 1) & 3;
        QIPB_CUDA(cudaEventSynchronize(ctx->tab_ev[slot]));      // the copy that last used this slot is done
        memcpy(ctx->tab_host[slot], tables.data(), need * sizeof(double2));
        QIPB_CUDA(cudaMemcpyAsync(ctx->tab_dev, ctx->tab_host[slot], need * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
        QIPB_CUDA(cudaEventRecord(ctx->tab_ev[slot], ctx->stream));
        f.tables = ctx->tab_dev;
    }
    if (dtype == QIPB_C128) return launch_fused<double2>(ctx, (double2 *)state, f);
    if (dtype == QIPB_C64) return launch_fused<float2>(ctx, (float2 *)state, f);
    QIPB_REQUIRE(false, "unknown dtype %d", dtype);
}
