// qip_b200/csrc/fused3.cu -- register-resident fused pass (the "v3" tile kernel).
//
// Same contract as fused.cu's kernel (one HBM round trip for a whole gate list on 2^12-amplitude
// tiles), different execution model.  fused.cu makes one shared-memory sweep PER GATE; profiles/
// shows it issue / shared-memory bound at ~0.47 of the HBM roofline.  Here a thread keeps 16
// amplitudes of the tile in REGISTERS: the tile's 12 local bits are split into 4 "register bits"
// (the 16 slots of a thread) and 8 "thread bits" (the 256 threads).  Every gate whose non-diagonal
// targets are register bits runs on registers only -- no shared-memory traffic, no barrier, and the
// matrix coefficients are loaded once per thread per gate and reused for all groups the thread holds.
// When the next gates need other bits, the tile is transposed through shared memory once (a "phase
// change": store, barrier, load with the new register bits).  Diagonal gates, stages and controls
// work on any bit in any phase (a control on a register bit selects slots, on a thread bit threads,
// outside the tile whole tiles).  Un-controlled swaps of two register bits are register renames.
//
// One persistent CTA per SM (256 threads, up to ~170 registers) with a ring of three 64 KiB tile
// buffers: TMA bulk loads of tile k+2 and bulk stores of tile k-1 overlap the compute of tile k.
#include <stdlib.h>
#include <algorithm>
#include "fused_shared.cuh"

namespace qipb {

#define V3_THREADS 256
#define V3_TB 12
#define V3_NBUF 3
#define V3_MAX_OPS 200
#define V3_MAX_STAGES 56
#define V3_POOL 900

enum { V3_PHASE = 0, V3_DENSE1 = 1, V3_DENSE2 = 2, V3_SCALE = 3, V3_STAGE = 4, V3_REGSWAP = 5 };

struct V3Desc {
    unsigned char type;
    unsigned char ra, rb;       // register-bit indices (DENSE2 / REGSWAP: ra > rb; ra = matrix MSB for DENSE2)
    unsigned char stage;        // V3_STAGE: index into V3Args::stages
    unsigned short jmask;       // slots (0..15) that satisfy the controls sitting on register bits
    unsigned short pad;
    u32 tmask;                  // controls on thread bits, as a mask over the tile-local index
    u32 coef;                   // offset into V3Args::pool (V3_STAGE: into the table buffer)
    u64 out_ctrl;               // controls outside the tile (state-index mask)
    unsigned char rpos[4];      // V3_PHASE: tile-local positions of the four register bits, ascending
    unsigned char pad2[4];
};

struct V3Args {
    int nbits, nops, lowrun, pad;
    u64 ntiles;
    const double2 *tables;
    unsigned char tbit[16];
    V3Desc ops[V3_MAX_OPS];
    StageInfo stages[V3_MAX_STAGES];
    double2 pool[V3_POOL];
};
static_assert(sizeof(V3Args) <= 32764, "V3Args must fit in the kernel parameter space");

// ---- register-level gate bodies (all slot indices are compile-time) ------------------------------
template <typename A, int R>
__device__ __forceinline__ void v3_dense1(A (&a)[16], const double2 *__restrict__ M, u32 jmask) {
    const double2 m0 = M[0], m1 = M[1], m2 = M[2], m3 = M[3];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (j & (1 << R)) continue;
        if ((jmask >> j) & 1u) {                       // uniform
            const A x0 = a[j], x1 = a[j | (1 << R)];
            A r0 = cmul<A>(m0, x0);
            cfma<A>(r0, m1, x1);
            A r1 = cmul<A>(m2, x0);
            cfma<A>(r1, m3, x1);
            a[j] = r0;
            a[j | (1 << R)] = r1;
        }
    }
}

template <typename A, int RA, int RB>      // RA > RB; matrix index = (bit RA, bit RB)
__device__ __forceinline__ void v3_dense2(A (&a)[16], const double2 *__restrict__ M, u32 jmask) {
    double2 m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = M[i];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (j & ((1 << RA) | (1 << RB))) continue;
        if ((jmask >> j) & 1u) {
            const int s[4] = {j, j | (1 << RB), j | (1 << RA), j | (1 << RA) | (1 << RB)};
            A x[4], r[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) x[c] = a[s[c]];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                r[i] = cmul<A>(m[i * 4], x[0]);
#pragma unroll
                for (int c = 1; c < 4; ++c) cfma<A>(r[i], m[i * 4 + c], x[c]);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) a[s[c]] = r[c];
        }
    }
}

template <typename A, int RA, int RB>
__device__ __forceinline__ void v3_regswap(A (&a)[16], u32 jmask) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (j & ((1 << RA) | (1 << RB))) continue;
        if ((jmask >> j) & 1u) {
            const A t = a[j | (1 << RA)];
            a[j | (1 << RA)] = a[j | (1 << RB)];
            a[j | (1 << RB)] = t;
        }
    }
}

template <typename A>
__device__ __forceinline__ void v3_dispatch1(A (&a)[16], int r, const double2 *M, u32 jmask) {
    switch (r) {
        case 0: v3_dense1<A, 0>(a, M, jmask); break;
        case 1: v3_dense1<A, 1>(a, M, jmask); break;
        case 2: v3_dense1<A, 2>(a, M, jmask); break;
        default: v3_dense1<A, 3>(a, M, jmask); break;
    }
}
template <typename A>
__device__ __forceinline__ void v3_dispatch2(A (&a)[16], int ra, int rb, const double2 *M, u32 jmask) {
    switch (ra * 4 + rb) {
        case 4: v3_dense2<A, 1, 0>(a, M, jmask); break;
        case 8: v3_dense2<A, 2, 0>(a, M, jmask); break;
        case 9: v3_dense2<A, 2, 1>(a, M, jmask); break;
        case 12: v3_dense2<A, 3, 0>(a, M, jmask); break;
        case 13: v3_dense2<A, 3, 1>(a, M, jmask); break;
        default: v3_dense2<A, 3, 2>(a, M, jmask); break;
    }
}
template <typename A>
__device__ __forceinline__ void v3_dispatch_swap(A (&a)[16], int ra, int rb, u32 jmask) {
    switch (ra * 4 + rb) {
        case 4: v3_regswap<A, 1, 0>(a, jmask); break;
        case 8: v3_regswap<A, 2, 0>(a, jmask); break;
        case 9: v3_regswap<A, 2, 1>(a, jmask); break;
        case 12: v3_regswap<A, 3, 0>(a, jmask); break;
        case 13: v3_regswap<A, 3, 1>(a, jmask); break;
        default: v3_regswap<A, 3, 2>(a, jmask); break;
    }
}

// tile-local index of slot j for this thread: ebase has zeros at the register-bit positions
__device__ __forceinline__ u32 v3_slot(u32 ebase, int j, u32 o0, u32 o1, u32 o2, u32 o3) {
    return ebase | ((j & 1) ? o0 : 0u) | ((j & 2) ? o1 : 0u) | ((j & 4) ? o2 : 0u) | ((j & 8) ? o3 : 0u);
}

template <typename A>
__global__ void __launch_bounds__(V3_THREADS, 1) fused3_kernel(A *__restrict__ state, const __grid_constant__ V3Args f) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full_bar[V3_NBUF];
    const int tid = threadIdx.x;
    constexpr u32 tsize = 1u << V3_TB;
    constexpr u32 tile_bytes = tsize * (u32)sizeof(A);
    const u32 run_amps = 1u << f.lowrun;
    const u32 nruns = tsize >> f.lowrun;
    const u32 run_bytes = run_amps * (u32)sizeof(A);

    if (tid == 0) {
        for (int b = 0; b < V3_NBUF; ++b) mbar_init(&full_bar[b], 1);
        fence_proxy_async();
    }
    __syncthreads();

    auto tile_base = [&](u64 t) {
        u64 base = t;
        for (int j = 0; j < V3_TB; ++j) base = insert_zero(base, f.tbit[j]);
        return base;
    };
    auto issue_load = [&](u64 t, int buf) {          // warp 0 only
        A *dst = reinterpret_cast<A *>(smem_raw + (size_t)buf * tile_bytes);
        const u64 base = tile_base(t);
        if (tid == 0) mbar_expect_tx(&full_bar[buf], tile_bytes);
        __syncwarp();
        for (u32 r = tid; r < nruns; r += 32) {
            u64 off = 0;
            for (int j = f.lowrun; j < V3_TB; ++j) off |= (u64)((r >> (j - f.lowrun)) & 1u) << f.tbit[j];
            bulk_g2s(dst + (size_t)r * run_amps, state + base + off, run_bytes, &full_bar[buf]);
        }
    };

    // prologue: the first two tiles of this CTA
    const u64 t0 = blockIdx.x, stride = gridDim.x;
    if (tid < 32) {
        if (t0 < f.ntiles) issue_load(t0, 0);
        if (t0 + stride < f.ntiles) issue_load(t0 + stride, 1);
    }

    u32 k = 0;
    for (u64 t = t0; t < f.ntiles; t += stride, ++k) {
        const int buf = k % V3_NBUF;
        A *tile = reinterpret_cast<A *>(smem_raw + (size_t)buf * tile_bytes);
        const u64 base = tile_base(t);
        mbar_wait(&full_bar[buf], (k / V3_NBUF) & 1u);

        // ---- run the op list with the tile in registers ----
        A a[16];
        u32 ebase = 0, o0 = 0, o1 = 0, o2 = 0, o3 = 0;
        bool loaded = false;
        for (int oi = 0; oi < f.nops; ++oi) {
            const V3Desc &d = f.ops[oi];
            if (d.type == V3_PHASE) {
                if (loaded) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) tile[v3_slot(ebase, j, o0, o1, o2, o3)] = a[j];
                    __syncthreads();
                }
                o0 = 1u << d.rpos[0]; o1 = 1u << d.rpos[1]; o2 = 1u << d.rpos[2]; o3 = 1u << d.rpos[3];
                u32 e = tid;                                   // spread the 8 thread bits around the register bits
                e += e & ~(o0 - 1u);
                e += e & ~(o1 - 1u);
                e += e & ~(o2 - 1u);
                e += e & ~(o3 - 1u);
                ebase = e;
#pragma unroll
                for (int j = 0; j < 16; ++j) a[j] = tile[v3_slot(ebase, j, o0, o1, o2, o3)];
                loaded = true;
                continue;
            }
            if ((base & d.out_ctrl) != d.out_ctrl) continue;       // uniform per tile
            const bool mine = (ebase & d.tmask) == d.tmask;         // controls on thread bits
            if (d.type == V3_STAGE) {
                const StageInfo &si = f.stages[d.stage];
                const double2 *T = f.tables + si.tab_off;
                constexpr u32 nlo = 1u << FUSED_LO_BITS, nhi = 1u << (V3_TB - FUSED_LO_BITS);
                const double2 S = stage_scalar(si, T, base, nlo, nhi);
                if (mine) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if ((d.jmask >> j) & 1u) {
                            const u32 e = v3_slot(ebase, j, o0, o1, o2, o3);
                            double2 ph = cmul<double2>(S, T[nlo + (e >> FUSED_LO_BITS)]);
                            ph = cmul<double2>(ph, T[e & (nlo - 1u)]);
                            a[j] = cmul<A>(ph, a[j]);
                        }
                }
                continue;
            }
            if (!mine) continue;
            const double2 *M = f.pool + d.coef;
            switch (d.type) {
                case V3_DENSE1: v3_dispatch1<A>(a, d.ra, M, d.jmask); break;
                case V3_DENSE2: v3_dispatch2<A>(a, d.ra, d.rb, M, d.jmask); break;
                case V3_REGSWAP: v3_dispatch_swap<A>(a, d.ra, d.rb, d.jmask); break;
                default: {   // V3_SCALE
                    const double2 ph = M[0];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if ((d.jmask >> j) & 1u) a[j] = cmul<A>(ph, a[j]);
                }
            }
        }
        if (loaded) {
#pragma unroll
            for (int j = 0; j < 16; ++j) tile[v3_slot(ebase, j, o0, o1, o2, o3)] = a[j];
        }

        // ---- write the tile back, prefetch the tile after next ----
        fence_proxy_async();
        __syncthreads();
        if (tid < 32) {
            for (u32 r = tid; r < nruns; r += 32) {
                u64 off = 0;
                for (int j = f.lowrun; j < V3_TB; ++j) off |= (u64)((r >> (j - f.lowrun)) & 1u) << f.tbit[j];
                bulk_s2g(state + base + off, tile + (size_t)r * run_amps, run_bytes);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // buffer (k+2)%3 held tile k-1: its stores (the group before the one just committed) must
            // have finished READING shared memory before the next load overwrites it
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            const u64 tn = t + 2 * stride;
            if (tn < f.ntiles) issue_load(tn, (k + 2) % V3_NBUF);
        }
    }
    if (tid < 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- host side --------------------------------------------------------------------------------
static bool is_swap4(const double *m) {
    static const double want[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
    for (int e = 0; e < 16; ++e)
        if (m[2 * e] != want[e] || m[2 * e + 1] != 0.0) return false;
    return true;
}

bool fused3_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("QIPB_FUSED_V3");
        v = e ? atoi(e) : 0;
    }
    return v != 0;
}

template <typename A>
static int launch3(qipb_ctx *ctx, A *state, const V3Args &f) {
    const size_t smem = (size_t)V3_NBUF * (sizeof(A) << V3_TB);
    QIPB_CUDA(cudaFuncSetAttribute(fused3_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    u64 grid = (u64)ctx->sm_count;
    if (grid > f.ntiles) grid = f.ntiles;
    fused3_kernel<A><<<(unsigned)grid, V3_THREADS, smem, ctx->stream>>>(state, f);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}

// Returns 0 on success, -1 when this pass is not eligible for the register-resident kernel (the
// caller then falls back to fused.cu's kernel), > 0 on error.
int fused3_apply(qipb_ctx *ctx, void *state, int nbits, int dtype, int ntile_bits, const int *tile_bits,
                 const int *local_of, u64 tmask, int ngates, const qipb_gate *gates) {
    if (ntile_bits != V3_TB) return -1;
    int lowrun = 0;
    while (lowrun < ntile_bits && tile_bits[lowrun] == lowrun) lowrun++;
    if ((16u << lowrun) < 512u || nbits - ntile_bits < 1) return -1;

    // ---- fold diagonal runs into stages (single diagonal gates with targets too) ----
    std::vector<Op> ops;
    std::vector<cplx> tables;
    {
        std::vector<int> run;
        auto flush_run = [&]() {
            if (run.empty()) return;
            // k = 0 phases in short runs stay scalar ops; anything with diagonal targets becomes a stage
            bool has_targets = false;
            for (int gi : run) has_targets |= gates[gi].k > 0;
            build_stages(gates, run, nbits, ntile_bits, local_of, tmask, ops, tables, has_targets ? 1 : 3);
            run.clear();
        };
        for (int gi = 0; gi < ngates; ++gi) {
            const qipb_gate &s = gates[gi];
            if (s.k < 0 || s.k > 2) return -1;
            if (s.diagonal != 0 || s.k == 0) {
                run.push_back(gi);
            } else {
                flush_run();
                Op o;
                o.stage = false;
                o.gate = gi;
                ops.push_back(o);
            }
        }
        flush_run();
    }
    for (const Op &o : ops)
        if (!o.stage && (gates[o.gate].diagonal != 0) && gates[o.gate].k > 0) return -1;   // not table-izable

    // ---- phases: consecutive ops whose non-diagonal targets fit in four register bits ----
    struct Phase { std::vector<int> bits; size_t first, last; };
    std::vector<Phase> phases;
    {
        Phase cur;
        cur.first = 0;
        for (size_t oi = 0; oi < ops.size(); ++oi) {
            std::vector<int> need;
            if (!ops[oi].stage) {
                const qipb_gate &s = gates[ops[oi].gate];
                if (!(s.diagonal != 0 || s.k == 0))
                    for (int j = 0; j < s.k; ++j) {
                        if (local_of[s.bits[j]] < 0) return -1;
                        need.push_back(local_of[s.bits[j]]);
                    }
            }
            std::vector<int> uni = cur.bits;
            for (int b : need)
                if (std::find(uni.begin(), uni.end(), b) == uni.end()) uni.push_back(b);
            if (uni.size() > 4) {
                cur.last = oi;
                phases.push_back(cur);
                cur = Phase();
                cur.first = oi;
                cur.bits = need;
            } else {
                cur.bits = uni;
            }
        }
        cur.last = ops.size();
        phases.push_back(cur);
    }

    static thread_local V3Args f;
    memset(&f, 0, sizeof(f));
    f.nbits = nbits;
    f.lowrun = lowrun;
    f.ntiles = 1ull << (nbits - ntile_bits);
    for (int j = 0; j < ntile_bits; ++j) f.tbit[j] = (unsigned char)tile_bits[j];
    int nops = 0, nstages = 0;
    u32 pool_used = 0;
    for (Phase &ph : phases) {
        // fill up to four register bits with the highest free tile-local positions (lanes then sit on
        // the low bits: conflict-free 128-bit shared-memory accesses)
        for (int b = V3_TB - 1; b >= 0 && ph.bits.size() < 4; --b)
            if (std::find(ph.bits.begin(), ph.bits.end(), b) == ph.bits.end()) ph.bits.push_back(b);
        std::sort(ph.bits.begin(), ph.bits.end());
        if (nops >= V3_MAX_OPS) return -1;
        V3Desc &pd = f.ops[nops++];
        pd.type = V3_PHASE;
        u32 regmask = 0;
        for (int r = 0; r < 4; ++r) {
            pd.rpos[r] = (unsigned char)ph.bits[r];
            regmask |= 1u << ph.bits[r];
        }
        auto reg_index = [&](int local) { return (int)(std::find(ph.bits.begin(), ph.bits.end(), local) - ph.bits.begin()); };
        for (size_t oi = ph.first; oi < ph.last; ++oi) {
            if (nops >= V3_MAX_OPS) return -1;
            const Op &o = ops[oi];
            V3Desc &d = f.ops[nops++];
            const u64 ctrl = o.stage ? o.common : gates[o.gate].ctrl_mask;
            d.out_ctrl = ctrl & ~tmask;
            // controls inside the tile: on register bits -> slot mask, on thread bits -> thread mask
            u32 cl = 0;
            for (int b = 0; b < nbits; ++b)
                if (((ctrl & tmask) >> b) & 1ull) cl |= 1u << local_of[b];
            d.tmask = cl & ~regmask;
            unsigned short jm = 0;
            for (int j = 0; j < 16; ++j) {
                u32 e = 0;
                for (int r = 0; r < 4; ++r)
                    if ((j >> r) & 1) e |= 1u << ph.bits[r];
                if ((e & cl & regmask) == (cl & regmask)) jm |= (unsigned short)(1u << j);
            }
            d.jmask = jm;
            if (o.stage) {
                if (nstages >= V3_MAX_STAGES) return -1;
                d.type = V3_STAGE;
                d.stage = (unsigned char)nstages;
                StageInfo &si = f.stages[nstages++];
                si.tab_off = o.tab_off;
                si.nout = (unsigned char)o.nout;
                for (int c = 0; c < o.nout; ++c) {
                    si.cn[c] = (unsigned char)o.cells[c].size();
                    for (size_t j = 0; j < o.cells[c].size(); ++j) si.cb[c][j] = (unsigned char)o.cells[c][j];
                }
                continue;
            }
            const qipb_gate &s = gates[o.gate];
            if (s.k == 0) {
                d.type = V3_SCALE;
                if (pool_used + 1 > V3_POOL) return -1;
                d.coef = pool_used;
                f.pool[pool_used++] = make_double2(s.mat[0], s.mat[1]);
            } else if (s.k == 1) {
                d.type = V3_DENSE1;
                d.ra = (unsigned char)reg_index(local_of[s.bits[0]]);
                if (pool_used + 4 > V3_POOL) return -1;
                d.coef = pool_used;
                for (int e = 0; e < 4; ++e) f.pool[pool_used++] = make_double2(s.mat[2 * e], s.mat[2 * e + 1]);
            } else {
                const int r0 = reg_index(local_of[s.bits[0]]), r1 = reg_index(local_of[s.bits[1]]);   // bits[0] = matrix MSB
                d.ra = (unsigned char)(r0 > r1 ? r0 : r1);
                d.rb = (unsigned char)(r0 > r1 ? r1 : r0);
                if (is_swap4(s.mat)) {
                    d.type = V3_REGSWAP;
                    continue;
                }
                d.type = V3_DENSE2;
                if (pool_used + 16 > V3_POOL) return -1;
                d.coef = pool_used;
                // the kernel indexes the matrix by (bit ra, bit rb) with ra the MSB: permute if bits[0] is the lower one
                for (int i = 0; i < 4; ++i)
                    for (int c = 0; c < 4; ++c) {
                        const int si = r0 > r1 ? i : ((i & 1) << 1) | (i >> 1);
                        const int sc = r0 > r1 ? c : ((c & 1) << 1) | (c >> 1);
                        f.pool[pool_used + i * 4 + c] = make_double2(s.mat[2 * (si * 4 + sc)], s.mat[2 * (si * 4 + sc) + 1]);
                    }
                pool_used += 16;
            }
        }
    }
    f.nops = nops;
    int rc = upload_tables(ctx, tables, &f.tables);
    if (rc) return rc;
    if (getenv("QIPB_DEBUG"))
        fprintf(stderr, "[qipb] fused3 launch: %d input gates -> %d ops in %d phases (%d stages), pool %u\n", ngates, nops,
                (int)phases.size(), nstages, pool_used);
    if (dtype == QIPB_C128) return launch3<double2>(ctx, (double2 *)state, f);
    return launch3<float2>(ctx, (float2 *)state, f);
}

}  // namespace qipb
