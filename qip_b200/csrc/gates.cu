// qip_b200/csrc/gates.cu -- in-place k-qubit gate application for sm_100a.
//
// Replaces the reference's cdot_loop (qip/ext/kronprod.pyx:157-197), which computes every output
// row as a gather over 2^K columns into a second buffer.  Here one thread owns whole 2^k-amplitude
// groups: it loads the group (one 128-bit transaction per complex128 amplitude), multiplies by the
// 2^k x 2^k matrix held in the kernel's constant bank, and stores back IN PLACE, so a pass moves
// exactly 2 * sizeof(amp) * (#touched amplitudes) bytes of HBM and needs no arena.
//
// Roofline: HBM-bound (14-30 flop per 32 B moved for 1-2 qubit gates); see DESIGN.md.
//
// Work decomposition: a group is named by a work index w in [0, 2^(nbits - nins)); zeros are
// inserted at the `nins` fixed positions (targets + controls, ascending) to get the base index,
// control bits are ORed in, and the 2^k members sit at base + off[j].  Consecutive threads get
// consecutive w, so as long as the lowest index bits are not fixed, a warp's j-th load covers
// 32 consecutive amplitudes = 512 contiguous bytes.  Each thread handles U groups and issues all
// of its loads before the first FMA to keep >= 8 independent 128-bit loads in flight.
#include "common.cuh"
#include "../../include/qip_b200.h"

namespace qipb {

#define QIPB_MAX_INS 48

template <int K>
struct GateArgs {
    u64 nwork;                    // number of groups
    u64 fixed_or;                 // control bits (all must be 1)
    int nins;
    unsigned char ins[QIPB_MAX_INS];   // ascending positions where a zero bit is inserted
    u64 off[1 << K];              // member offsets, matrix-index order
    double2 m[(1 << K) * (1 << K)];
};

__device__ __forceinline__ u64 expand_index(u64 w, const unsigned char *ins, int nins, u64 fixed_or) {
    for (int j = 0; j < nins; ++j) w = insert_zero(w, ins[j]);
    return w | fixed_or;
}

// 256-bit (two complex128) accesses for the case where the two members of a group are adjacent
// in memory (target bit 0): one LDG.256 / STG.256 per group instead of two half-used sectors.
__device__ __forceinline__ void ld256(const double2 *p, double2 &a, double2 &b) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}
__device__ __forceinline__ void st256(double2 *p, const double2 a, const double2 b) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}

template <typename A, int K, int U, bool DIAG>
__global__ void __launch_bounds__(256) gate_kernel(A *__restrict__ state, const __grid_constant__ GateArgs<K> g) {
    constexpr int D = 1 << K;
    A a[U][D];
    u64 base[U];
    bool live[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const u64 w = ((u64)blockIdx.x * U + u) * blockDim.x + threadIdx.x;
        live[u] = w < g.nwork;
        base[u] = expand_index(w, g.ins, g.nins, g.fixed_or);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (live[u]) {
#pragma unroll
            for (int j = 0; j < D; ++j) a[u][j] = state[base[u] + g.off[j]];
        }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (live[u]) {
            if (DIAG) {
#pragma unroll
                for (int i = 0; i < D; ++i) state[base[u] + g.off[i]] = cmul<A>(g.m[i * D + i], a[u][i]);
            } else {
                A r[D];
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    r[i] = cmul<A>(g.m[i * D], a[u][0]);
#pragma unroll
                    for (int j = 1; j < D; ++j) cfma<A>(r[i], g.m[i * D + j], a[u][j]);
                }
#pragma unroll
                for (int i = 0; i < D; ++i) state[base[u] + g.off[i]] = r[i];
            }
        }
}

// K = 1, complex128, target bit 0: the pair is one aligned 32-byte object.
template <int U>
__global__ void __launch_bounds__(256) gate1_bit0_kernel(double2 *__restrict__ state, const __grid_constant__ GateArgs<1> g) {
    double2 a0[U], a1[U];
    u64 base[U];
    bool live[U];
    const bool rev = g.off[0] != 0;     // matrix index 0 sits at the odd address
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const u64 w = ((u64)blockIdx.x * U + u) * blockDim.x + threadIdx.x;
        live[u] = w < g.nwork;
        base[u] = expand_index(w, g.ins, g.nins, g.fixed_or);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (live[u]) ld256(state + base[u], a0[u], a1[u]);
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (live[u]) {
            const double2 x0 = rev ? a1[u] : a0[u], x1 = rev ? a0[u] : a1[u];
            double2 r0 = cmul<double2>(g.m[0], x0);
            cfma<double2>(r0, g.m[1], x1);
            double2 r1 = cmul<double2>(g.m[2], x0);
            cfma<double2>(r1, g.m[3], x1);
            st256(state + base[u], rev ? r1 : r0, rev ? r0 : r1);
        }
}

// Dense gate on many target bits (QIPB_MAX_DENSE_K < k <= QIPB_MAX_BIG_K): a batched 2^k x 2^k matrix-vector product
// (kronprod.pyx:157-197 with a wide K: 2^K column probes per output row there).
struct BigArgs {
    u64 nwork;
    u64 fixed_or;
    int nins;
    int k;
    int gb;                               // groups per batch (a power of two; X[D][gb] fits the shared memory of an SM)
    int nbuf;                             // tensor path: 1 or 2 batch buffers
    unsigned char ins[QIPB_MAX_INS];
    unsigned char tbit[QIPB_MAX_BIG_K];   // bit position of matrix-index bit j (j = 0 least significant)
};
// A batch of GB groups (columns) is staged in shared memory as X[D][GB]; thread (gc, rb) = (tid % GB, tid / GB) owns
// column gc and the rows rb, rb + NRB, ... (NRB = 256 / GB): out[i][gc] = sum_j M[i][j] X[j][gc], RPT rows at a time in
// registers.  With GB = 32 a warp shares its rows, so the matrix entries are warp-uniform loads served by L1 / L2 (16 KiB
// at K = 5, 16 MiB at K = 10), the X reads are conflict-free (32 consecutive amplitudes) and for every X value loaded a
// thread issues 4 * RPT FP64 FMAs: the kernel is FP64 bound (4 * 2^K FMAs per amplitude against 32 bytes moved).  Every
// input of a batch is in shared memory before the first output is written, so the update is in place.
template <typename A, int RPT>
__global__ void __launch_bounds__(256) big_gate_kernel(A *__restrict__ state, const double2 *__restrict__ mat,
                                                        const __grid_constant__ BigArgs g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    A *X = reinterpret_cast<A *>(smem_raw);
    const int D = 1 << g.k;
    const int GB = g.gb, NRB = 256 / GB;
    const int gc = threadIdx.x % GB, rb = threadIdx.x / GB;
    for (u64 w0 = (u64)blockIdx.x * GB; w0 < g.nwork; w0 += (u64)gridDim.x * GB) {
        for (int idx = threadIdx.x; idx < D * GB; idx += 256) {
            const int j = idx / GB, c = idx % GB;
            if (w0 + c < g.nwork) {
                u64 off = 0;
                for (int b = 0; b < g.k; ++b) off |= (u64)((j >> b) & 1) << g.tbit[b];
                X[idx] = state[expand_index(w0 + c, g.ins, g.nins, g.fixed_or) + off];
            }
        }
        __syncthreads();
        if (w0 + gc < g.nwork) {
            const u64 base = expand_index(w0 + gc, g.ins, g.nins, g.fixed_or);
            for (int i0 = rb; i0 < D; i0 += NRB * RPT) {
                A acc[RPT];
#pragma unroll
                for (int r = 0; r < RPT; ++r) acc[r] = make_amp<A>(0, 0);
                for (int j = 0; j < D; ++j) {
                    const A x = X[j * GB + gc];
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        const int i = i0 + NRB * r;
                        if (i < D) cfma<A>(acc[r], __ldg(mat + (size_t)i * D + j), x);
                    }
                }
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const int i = i0 + NRB * r;
                    if (i < D) {
                        u64 off = 0;
                        for (int b = 0; b < g.k; ++b) off |= (u64)((i >> b) & 1) << g.tbit[b];
                        state[base + off] = acc[r];
                    }
                }
            }
        }
        __syncthreads();
    }
}

// Batches on the FP64 tensor path: Y[D][GB] = M[D][D] X[D][GB] as 8x8x4 `mma.sync` tiles (DMMA), four real
// products per complex one.  The scalar kernel above issues one shared-memory load and RPT matrix loads per 4 * RPT
// FMAs and stays at a fifth of the FP64 rate; here a warp task (S row strips of 8 x CB column blocks of 8, S * CB = 4)
// issues S matrix-fragment loads (global, L1 / L2 resident) and CB shared-memory loads per 16 tile products = 4096 FMAs.
// Fragment layout of mma.m8n8k4.f64 (lane = 4 * g + t): A[g][t], B[t][g], C[g][2t], C[g][2t + 1].
// X rows are padded to GB + 2 amplitudes so that the B-fragment loads (4 rows x 2 columns per quarter warp) fall into
// distinct 16-byte bank groups.  The update stays in place: a batch is complete in shared memory before its first store.
__device__ __forceinline__ void dmma884(double &c0, double &c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// The gather of a batch is asynchronous (16-byte cp.async, every copy of a thread in flight at once); with g.nbuf == 2 the
// next batch is gathered while the tiles of the current one are multiplied.
// complex64 states use the same tiles: amplitudes are widened to double on the way into shared memory (register-staged
// loads instead of cp.async) and rounded once on the way out, so a K-qubit gate costs one rounding per amplitude.
template <typename A, int S, int CB>
__global__ void __launch_bounds__(256, 3) big_gate_mma_kernel(A *__restrict__ state, const double2 *__restrict__ mat,
                                                            const __grid_constant__ BigArgs g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = 1 << g.k;
    const int GB = g.gb, LD = GB + 2, gshift = 31 - __clz(GB);
    double2 *Xall = reinterpret_cast<double2 *>(smem_raw);
    u64 *rowoff = reinterpret_cast<u64 *>(Xall + (size_t)g.nbuf * D * LD);
    u64 *colall = rowoff + D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fg = lane >> 2, ft = lane & 3;
    for (int j = threadIdx.x; j < D; j += 256) {
        u64 off = 0;
        for (int b = 0; b < g.k; ++b) off |= (u64)((j >> b) & 1) << g.tbit[b];
        rowoff[j] = off;
    }
    // start the gather of the batch at w into buffer b (the caller has made sure nobody still reads that buffer)
    auto gather = [&](const u64 w, const int b) {
        double2 *X = Xall + (size_t)b * D * LD;
        u64 *colbase = colall + b * GB;
        if (threadIdx.x < GB && w + threadIdx.x < g.nwork)
            colbase[threadIdx.x] = expand_index(w + threadIdx.x, g.ins, g.nins, g.fixed_or);
        __syncthreads();                                  // colbase (and, the first time, rowoff)
        if constexpr (sizeof(A) == sizeof(double2)) {
            for (int idx = threadIdx.x; idx < D * GB; idx += 256) {
                const int j = idx >> gshift, c = idx & (GB - 1);
                if (w + c < g.nwork) cp_async16(X + j * LD + c, state + colbase[c] + rowoff[j]);
                else X[j * LD + c] = make_double2(0.0, 0.0);
            }
        } else {
            for (int i0 = threadIdx.x; i0 < D * GB; i0 += 256 * 4) {          // four loads in flight per thread
                A v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = i0 + 256 * u, j = idx >> gshift, c = idx & (GB - 1);
                    v[u] = (idx < D * GB && w + c < g.nwork) ? state[colbase[c] + rowoff[j]] : make_amp<A>(0, 0);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = i0 + 256 * u, j = idx >> gshift, c = idx & (GB - 1);
                    if (idx < D * GB) X[j * LD + c] = make_double2((double)v[u].x, (double)v[u].y);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int ncg = GB / (8 * CB), ntask = (D / (8 * S)) * ncg;
    const u64 stride = (u64)gridDim.x * GB;
    u64 w0 = (u64)blockIdx.x * GB;
    if (w0 < g.nwork) gather(w0, 0);
    for (int buf = 0; w0 < g.nwork; w0 += stride) {
        const bool more = w0 + stride < g.nwork;
        if (g.nbuf == 2 && more) {
            gather(w0 + stride, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double2 *X = Xall + (size_t)buf * D * LD;
        const u64 *colbase = colall + buf * GB;
        for (int t = warp; t < ntask; t += 8) {
            const int sg = t / ncg, c0 = (t % ncg) * 8 * CB;
            if (w0 + c0 >= g.nwork) continue;             // warp-uniform: a column group past the end of the state
            double yr[S][CB][2], yi[S][CB][2];
#pragma unroll
            for (int s = 0; s < S; ++s)
#pragma unroll
                for (int cb = 0; cb < CB; ++cb) yr[s][cb][0] = yr[s][cb][1] = yi[s][cb][0] = yi[s][cb][1] = 0.0;
            const double2 *mrow = mat + (size_t)(sg * 8 * S + fg) * D + ft;
            const double2 *xcol = X + ft * LD + c0 + fg;
#pragma unroll 2
            for (int j0 = 0; j0 < D; j0 += 4) {
                double2 a[S], x[CB];
#pragma unroll
                for (int s = 0; s < S; ++s) a[s] = __ldg(mrow + (size_t)s * 8 * D + j0);
#pragma unroll
                for (int cb = 0; cb < CB; ++cb) x[cb] = xcol[j0 * LD + cb * 8];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const double nai = -a[s].y;
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb) dmma884(yr[s][cb][0], yr[s][cb][1], a[s].x, x[cb].x);
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb) dmma884(yi[s][cb][0], yi[s][cb][1], a[s].x, x[cb].y);
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb) dmma884(yr[s][cb][0], yr[s][cb][1], nai, x[cb].y);
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb) dmma884(yi[s][cb][0], yi[s][cb][1], a[s].y, x[cb].x);
                }
            }
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const u64 ro = rowoff[(sg * S + s) * 8 + fg];
#pragma unroll
                for (int cb = 0; cb < CB; ++cb)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = c0 + cb * 8 + 2 * ft + e;
                        if (w0 + c < g.nwork) state[colbase[c] + ro] = make_amp<A>(yr[s][cb][e], yi[s][cb][e]);
                    }
            }
        }
        __syncthreads();                                  // every warp is done with this buffer
        if (g.nbuf == 2) buf ^= 1;
        else if (more) gather(w0 + stride, 0);
    }
}

// ---- host side ------------------------------------------------------------------------------
template <int K>
static int fill_args(GateArgs<K> &g, int nbits, const int *bits, const double *mat, u64 ctrl_mask) {
    u64 tmask = 0;
    for (int j = 0; j < K; ++j) {
        QIPB_REQUIRE(bits[j] >= 0 && bits[j] < nbits, "target bit %d out of range [0,%d)", bits[j], nbits);
        QIPB_REQUIRE(!((tmask >> bits[j]) & 1ull), "repeated target bit %d", bits[j]);
        tmask |= 1ull << bits[j];
    }
    QIPB_REQUIRE((ctrl_mask & tmask) == 0, "control mask overlaps target bits");
    QIPB_REQUIRE(nbits == 64 || (ctrl_mask >> nbits) == 0, "control mask outside the %d local bits", nbits);
    const u64 fixed = tmask | ctrl_mask;
    g.nins = 0;
    for (int b = 0; b < nbits; ++b)
        if ((fixed >> b) & 1ull) {
            QIPB_REQUIRE(g.nins < QIPB_MAX_INS, "too many fixed bits");
            g.ins[g.nins++] = (unsigned char)b;
        }
    g.fixed_or = ctrl_mask;
    g.nwork = 1ull << (nbits - g.nins);
    for (int j = 0; j < (1 << K); ++j) {
        u64 off = 0;
        for (int t = 0; t < K; ++t)
            if ((j >> (K - 1 - t)) & 1) off |= 1ull << bits[t];   // bits[0] = MSB of the matrix index
        g.off[j] = off;
    }
    for (int e = 0; e < (1 << K) * (1 << K); ++e) g.m[e] = make_double2(mat[2 * e], mat[2 * e + 1]);
    return QIPB_OK;
}

template <typename A, int K, int U, bool DIAG>
static int launch_gate(qipb_ctx *ctx, A *state, const GateArgs<K> &g) {
    const u64 per_block = 256ull * U;
    const u64 blocks = (g.nwork + per_block - 1) / per_block;
    QIPB_REQUIRE(blocks <= 0x7fffffffull, "grid too large");
    gate_kernel<A, K, U, DIAG><<<(unsigned)blocks, 256, 0, ctx->stream>>>(state, g);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}

template <typename A, int K>
static int apply_k(qipb_ctx *ctx, A *state, int nbits, const int *bits, const double *mat, u64 ctrl_mask, int diagonal) {
    GateArgs<K> g;
    memset(&g, 0, sizeof(g));
    int rc = fill_args<K>(g, nbits, bits, mat, ctrl_mask);
    if (rc) return rc;
    // groups per thread chosen so that every thread has 8 amplitudes (128 B of c128) in flight
    constexpr int U = (K == 0) ? 8 : (K == 1) ? 4 : (K == 2) ? 2 : 1;
    if (diagonal) return launch_gate<A, K, U, true>(ctx, state, g);
    return launch_gate<A, K, U, false>(ctx, state, g);
}

// Launch of the FP64-tensor kernel for one dense gate on g.k >= 5 bits (dmat: the matrix on the device).
template <typename A>
static int launch_big_mma(qipb_ctx *ctx, A *state, const double2 *dmat, BigArgs &g) {
    const size_t D = (size_t)1 << g.k;
    // batch width and buffers, measured on B200 (scripts/big_gate_probe.py, profiles/r02_big_gate_probe.txt): three CTAs per SM
    // (registers allow no more) matter more than a wide batch, and with three CTAs overlapping their phases a second buffer
    // (gather of batch i + 1 behind the tiles of batch i) costs more in barriers than it hides
    int gb = g.k == 5 ? 64 : g.k <= 7 ? 32 : 8;
    int nbuf = 1;
    if (const char *e = getenv("QIPB_BIG_GB")) {           // tuning knobs for profiling runs
        const int v = atoi(e);
        if (v == 8 || v == 16 || v == 32 || v == 64 || v == 128) gb = v;
    }
    if (const char *e = getenv("QIPB_BIG_NBUF")) nbuf = atoi(e) == 2 ? 2 : 1;
    while (gb > 8 && (u64)gb > 2 * g.nwork) gb >>= 1;      // a state with fewer groups than a batch is wide
    auto smem_of = [&](int w, int nb) { return (size_t)nb * D * (w + 2) * sizeof(double2) + (D + 2 * w) * sizeof(u64); };
    if (smem_of(gb, nbuf) > 220u * 1024u) nbuf = 1;
    while (gb > 8 && smem_of(gb, nbuf) > 220u * 1024u) gb >>= 1;
    g.gb = gb;
    g.nbuf = nbuf;
    const size_t smem = smem_of(gb, nbuf);
    const u64 nbatch = (g.nwork + gb - 1) / gb;
    const int per_sm = (int)((226u * 1024u) / (smem + 1024));
    u64 blocks = (u64)ctx->sm_count * (per_sm > 3 ? 3 : per_sm < 1 ? 1 : per_sm);
    if (blocks > nbatch) blocks = nbatch;
    if (gb >= 32) {
        QIPB_CUDA(cudaFuncSetAttribute(big_gate_mma_kernel<A, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        big_gate_mma_kernel<A, 1, 4><<<(unsigned)blocks, 256, smem, ctx->stream>>>(state, dmat, g);
    } else if (gb == 16) {
        QIPB_CUDA(cudaFuncSetAttribute(big_gate_mma_kernel<A, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        big_gate_mma_kernel<A, 2, 2><<<(unsigned)blocks, 256, smem, ctx->stream>>>(state, dmat, g);
    } else {
        QIPB_CUDA(cudaFuncSetAttribute(big_gate_mma_kernel<A, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        big_gate_mma_kernel<A, 4, 1><<<(unsigned)blocks, 256, smem, ctx->stream>>>(state, dmat, g);
    }
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}

template <typename A>
static int apply_big(qipb_ctx *ctx, A *state, int nbits, int k, const int *bits, const double *mat, u64 ctrl_mask) {
    BigArgs g;
    memset(&g, 0, sizeof(g));
    u64 tmask = 0;
    for (int j = 0; j < k; ++j) {
        QIPB_REQUIRE(bits[j] >= 0 && bits[j] < nbits, "target bit out of range");
        QIPB_REQUIRE(!((tmask >> bits[j]) & 1ull), "repeated target bit");
        tmask |= 1ull << bits[j];
        g.tbit[k - 1 - j] = (unsigned char)bits[j];
    }
    QIPB_REQUIRE((ctrl_mask & tmask) == 0, "control mask overlaps target bits");
    QIPB_REQUIRE(nbits == 64 || (ctrl_mask >> nbits) == 0, "control mask outside the %d local bits", nbits);
    const u64 fixed = tmask | ctrl_mask;
    for (int b = 0; b < nbits; ++b)
        if ((fixed >> b) & 1ull) g.ins[g.nins++] = (unsigned char)b;
    g.fixed_or = ctrl_mask;
    g.k = k;
    g.nwork = 1ull << (nbits - g.nins);
    const size_t D = (size_t)1 << k;
    double2 *dmat = nullptr;
    QIPB_CUDA(cudaMallocAsync(&dmat, D * D * sizeof(double2), ctx->stream));
    QIPB_CUDA(cudaMemcpyAsync(dmat, mat, D * D * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
    static const int use_mma = [] { const char *e = getenv("QIPB_BIG_MMA"); return e ? atoi(e) : 1; }();
    if (use_mma) {
        const int rc = launch_big_mma<A>(ctx, state, dmat, g);
        const cudaError_t fe = cudaFreeAsync(dmat, ctx->stream);
        if (rc) return rc;
        QIPB_CUDA(fe);
        return QIPB_OK;
    }
    g.gb = 32;
    while ((size_t)g.gb * D * sizeof(A) > 128u * 1024u) g.gb >>= 1;
    const size_t smem = (size_t)g.gb * D * sizeof(A);
    const int rows_per_thread = (int)(D / (256 / g.gb));
    const u64 nbatch = (g.nwork + g.gb - 1) / g.gb;
    const int per_sm = smem > 100u * 1024u ? 1 : (int)((200u * 1024u) / (smem + 1024));
    u64 blocks = (u64)ctx->sm_count * (per_sm > 8 ? 8 : per_sm);
    if (blocks > nbatch) blocks = nbatch;
    if (rows_per_thread <= 4) {
        QIPB_CUDA(cudaFuncSetAttribute(big_gate_kernel<A, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        big_gate_kernel<A, 4><<<(unsigned)blocks, 256, smem, ctx->stream>>>(state, dmat, g);
    } else if (rows_per_thread <= 8) {
        QIPB_CUDA(cudaFuncSetAttribute(big_gate_kernel<A, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        big_gate_kernel<A, 8><<<(unsigned)blocks, 256, smem, ctx->stream>>>(state, dmat, g);
    } else {
        QIPB_CUDA(cudaFuncSetAttribute(big_gate_kernel<A, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        big_gate_kernel<A, 16><<<(unsigned)blocks, 256, smem, ctx->stream>>>(state, dmat, g);
    }
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    QIPB_CUDA(cudaFreeAsync(dmat, ctx->stream));
    return QIPB_OK;
}

template <typename A>
static int apply_matrix_t(qipb_ctx *ctx, A *state, int nbits, int k, const int *bits, const double *mat,
                          u64 ctrl_mask, int diagonal) {
    switch (k) {
        case 0: return apply_k<A, 0>(ctx, state, nbits, bits, mat, ctrl_mask, 1);
        case 1: return apply_k<A, 1>(ctx, state, nbits, bits, mat, ctrl_mask, diagonal);
        case 2: return apply_k<A, 2>(ctx, state, nbits, bits, mat, ctrl_mask, diagonal);
        case 3: return apply_k<A, 3>(ctx, state, nbits, bits, mat, ctrl_mask, diagonal);
        case 4: return apply_k<A, 4>(ctx, state, nbits, bits, mat, ctrl_mask, diagonal);
        default: return apply_big<A>(ctx, state, nbits, k, bits, mat, ctrl_mask);
    }
}

}  // namespace qipb

using namespace qipb;

extern "C" int qipb_apply_matrix(qipb_ctx *ctx, void *state, int nbits, int dtype, int k, const int *bits,
                                 const double *mat, uint64_t ctrl_mask, int diagonal) {
    QIPB_REQUIRE(ctx && state && mat, "null argument");
    QIPB_REQUIRE(nbits >= 0 && nbits <= 40, "nbits %d unsupported", nbits);
    QIPB_REQUIRE(k >= 0 && k <= QIPB_MAX_BIG_K && k <= nbits, "k=%d unsupported (0..%d, <= nbits)", k, QIPB_MAX_BIG_K);
    QIPB_REQUIRE(k == 0 || bits, "null bits");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    // complex128, k == 1, target bit 0, dense: 256-bit path
    if (dtype == QIPB_C128 && k == 1 && !diagonal && bits[0] == 0 && nbits >= 1) {
        GateArgs<1> g;
        memset(&g, 0, sizeof(g));
        int rc = fill_args<1>(g, nbits, bits, mat, (u64)ctrl_mask);
        if (rc) return rc;
        const u64 per_block = 256ull * 4;
        const u64 blocks = (g.nwork + per_block - 1) / per_block;
        gate1_bit0_kernel<4><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double2 *)state, g);
        ctx->launches++;
        QIPB_CUDA(cudaGetLastError());
        return QIPB_OK;
    }
    if (dtype == QIPB_C128) return apply_matrix_t<double2>(ctx, (double2 *)state, nbits, k, bits, mat, ctrl_mask, diagonal);
    if (dtype == QIPB_C64) return apply_matrix_t<float2>(ctx, (float2 *)state, nbits, k, bits, mat, ctrl_mask, diagonal);
    QIPB_REQUIRE(false, "unknown dtype %d", dtype);
}

// Swap of two index bits = exchange of the members 01 <-> 10 of every group; expressed as a
// one-"qubit" X gate whose two members sit at base + 2^a and base + 2^b, so only the amplitudes
// that actually move are touched (half of the state).
extern "C" int qipb_apply_swap(qipb_ctx *ctx, void *state, int nbits, int dtype, int bit_a, int bit_b,
                               uint64_t ctrl_mask) {
    QIPB_REQUIRE(ctx && state, "null argument");
    QIPB_REQUIRE(bit_a >= 0 && bit_a < nbits && bit_b >= 0 && bit_b < nbits && bit_a != bit_b, "bad swap bits %d,%d", bit_a, bit_b);
    QIPB_REQUIRE(!(ctrl_mask & ((1ull << bit_a) | (1ull << bit_b))), "control mask overlaps swap bits");
    QIPB_REQUIRE(nbits >= 2 && nbits <= 40 && (ctrl_mask >> nbits) == 0, "control mask outside the %d local bits", nbits);
    QIPB_CUDA(cudaSetDevice(ctx->device));
    GateArgs<1> g;
    memset(&g, 0, sizeof(g));
    const u64 fixed = (1ull << bit_a) | (1ull << bit_b) | ctrl_mask;
    for (int b = 0; b < nbits; ++b)
        if ((fixed >> b) & 1ull) g.ins[g.nins++] = (unsigned char)b;
    g.fixed_or = ctrl_mask;
    g.nwork = 1ull << (nbits - g.nins);
    g.off[0] = 1ull << bit_a;
    g.off[1] = 1ull << bit_b;
    g.m[0] = make_double2(0, 0); g.m[1] = make_double2(1, 0);
    g.m[2] = make_double2(1, 0); g.m[3] = make_double2(0, 0);
    if (dtype == QIPB_C128) return launch_gate<double2, 1, 4, false>(ctx, (double2 *)state, g);
    if (dtype == QIPB_C64) return launch_gate<float2, 1, 4, false>(ctx, (float2 *)state, g);
    QIPB_REQUIRE(false, "unknown dtype %d", dtype);
}
