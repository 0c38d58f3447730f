// qip_b200/csrc/init.cu -- initial-state builders and the func_apply permutation for sm_100a.
//
// init_kron replaces CythonBackend.make_state's python loop over 2^(#fed qubits) entries
// (qip/backend.py:88-101, qip/util.py:108-124) with one write-only sweep: every thread evaluates
// the kron-product entry of its own amplitude from the fed vectors (device resident, L2-cached).
// func_xor replaces func_apply (qip/ext/func_apply.pyx:39-112) with an in-place pairwise swap:
// for fixed x the map q -> f(x) xor q is an involution, so no second buffer and no zero-fill.
// Roofline: HBM-bound; init writes sizeof(amp)*2^nbits, func_xor moves only the amplitudes whose
// f(x) != 0 (plus one table read per amplitude).
#include "common.cuh"
#include "../../include/qip_b200.h"

namespace qipb {

#define QIPB_MAX_GROUPS 40
#define KRON_LOW_BITS 12          // the product over groups that live on index bits 0..11 is tabulated once

enum { KRON_LOW = 0, KRON_HIGH = 1, KRON_MIXED = 2 };

struct KronArgs {
    u64 n;
    u64 shard_base;       // shard_index << nbits
    u64 fixed_mask;       // global bits with a prescribed value (un-fed qubits: 0; one-hot feeds: the index)
    u64 fixed_value;
    int ngroups;
    int lb;               // min(KRON_LOW_BITS, nbits)
    int run_begin[QIPB_MAX_GROUPS + 1];
    u64 feed_off[QIPB_MAX_GROUPS];
    unsigned char cls[QIPB_MAX_GROUPS];      // KRON_LOW / KRON_HIGH / KRON_MIXED (straddles bit lb)
    BitRuns runs;         // all groups' gathers, concatenated
};

__device__ __forceinline__ double2 kron_factor(const KronArgs &k, const double2 *__restrict__ feeds, int g, u64 G) {
    u64 sub = 0;
    for (int r = k.run_begin[g]; r < k.run_begin[g + 1]; ++r)
        sub |= ((G >> k.runs.src[r]) & ((1ull << k.runs.len[r]) - 1ull)) << k.runs.dst[r];
    return feeds[k.feed_off[g] + sub];
}

__device__ __forceinline__ double2 zmul(const double2 a, const double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// The kron product factorises over index bits: amp(G) = HIGH(G >> lb) * LOW(G & (2^lb - 1)) * MIXED(G).
// Step 1: LOW for all 2^lb low patterns (groups entirely below bit lb, and the prescribed low bits).
__global__ void __launch_bounds__(256) kron_low_table_kernel(double2 *__restrict__ table, const double2 *__restrict__ feeds,
                                                             const __grid_constant__ KronArgs k) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1ull << k.lb)) return;
    const u64 lowmask = (1ull << k.lb) - 1ull;
    double2 v = make_double2(0.0, 0.0);
    if ((i & k.fixed_mask & lowmask) == (k.fixed_value & lowmask)) {
        v = make_double2(1.0, 0.0);
        for (int g = 0; g < k.ngroups; ++g)                   // left to right like qip/backend.py:98-101
            if (k.cls[g] == KRON_LOW) v = zmul(v, kron_factor(k, feeds, g, i));
    }
    table[i] = v;
}

// Step 2: one block per run of 2^lb consecutive amplitudes.  HIGH is the same for the whole run: warp 0
// gathers one factor per lane, lane 0 multiplies them in group order.  Every amplitude is then one table
// read, one complex multiply (plus the straddling groups, if any) and one coalesced 128-bit store --
// a write-only sweep at HBM speed however many groups the feed has (33 one-qubit feeds used to cost 33
// gathers per amplitude: 159 ms at 33 qubits, against 32 ms for three wide groups).
template <typename A>
__global__ void __launch_bounds__(256) init_kron_kernel(A *__restrict__ state, const double2 *__restrict__ feeds,
                                                        const double2 *__restrict__ low_table, const __grid_constant__ KronArgs k) {
    typedef typename amp_traits<A>::real R;
    __shared__ double2 fac[QIPB_MAX_GROUPS];
    __shared__ double2 high;
    const u64 run = (u64)blockIdx.x;
    const u64 i0 = run << k.lb;
    const u64 G0 = k.shard_base | i0;
    const u64 himask = ~((1ull << k.lb) - 1ull);
    const int t = threadIdx.x;
    if (t < k.ngroups && k.cls[t] == KRON_HIGH) fac[t] = kron_factor(k, feeds, t, G0);
    __syncthreads();
    if (t == 0) {
        double2 v = make_double2(0.0, 0.0);
        if ((G0 & k.fixed_mask & himask) == (k.fixed_value & himask)) {
            v = make_double2(1.0, 0.0);
            for (int g = 0; g < k.ngroups; ++g)
                if (k.cls[g] == KRON_HIGH) v = zmul(v, fac[g]);
        }
        high = v;
    }
    __syncthreads();
    const double2 H = high;
    const u32 runlen = 1u << k.lb;
    bool mixed = false;
    for (int g = 0; g < k.ngroups; ++g) mixed |= k.cls[g] == KRON_MIXED;
    for (u32 j = t; j < runlen; j += 256) {
        double2 v = zmul(H, low_table[j]);
        if (mixed) {
            const u64 G = G0 | j;
            for (int g = 0; g < k.ngroups; ++g)
                if (k.cls[g] == KRON_MIXED) v = zmul(v, kron_factor(k, feeds, g, G));
        }
        state[i0 + j] = make_amp<A>((R)v.x, (R)v.y);
    }
}

template <typename A>
__global__ void __launch_bounds__(256) init_basis_kernel(A *__restrict__ state, u64 n, long long index) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    state[i] = ((long long)i == index) ? make_amp<A>(1, 0) : make_amp<A>(0, 0);
}

struct FuncArgs {
    u64 n;
    u64 x_fixed;
    u64 ymask;            // (1 << n2) - 1
    BitRuns xruns;        // state index -> x
    BitRuns yruns;        // y -> state index bits of reg2
};

// TAB: long long (the general table) or unsigned char (|reg2| <= 8: the only bits of f(x) that matter, func_apply.pyx:97,
// fit a byte -- an eighth of the table traffic and of the upload; a Grover oracle on 27 search qubits is 128 MiB, not 1 GiB)
template <typename A, typename TAB>
__global__ void __launch_bounds__(256) func_xor_kernel(A *__restrict__ state, const TAB *__restrict__ table,
                                                       const __grid_constant__ FuncArgs f) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f.n) return;
    const u64 x = runs_gather(f.xruns, i) | f.x_fixed;
    const u64 y = (u64)table[x] & f.ymask;
    if (y == 0) return;
    const u64 partner = i ^ runs_gather(f.yruns, y);
    if (i < partner) {
        const A a = state[i], b = state[partner];
        state[i] = b;
        state[partner] = a;
    }
}

}  // namespace qipb

using namespace qipb;

extern "C" int qipb_init_basis(qipb_ctx *ctx, void *state, int nbits, int dtype, long long index) {
    QIPB_REQUIRE(ctx && state, "null argument");
    QIPB_REQUIRE(nbits >= 0 && nbits <= 40, "nbits %d unsupported", nbits);
    QIPB_CUDA(cudaSetDevice(ctx->device));
    const u64 n = 1ull << nbits;
    const u64 blocks = (n + 255) / 256;
    if (dtype == QIPB_C128) init_basis_kernel<double2><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double2 *)state, n, index);
    else if (dtype == QIPB_C64) init_basis_kernel<float2><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float2 *)state, n, index);
    else QIPB_REQUIRE(false, "unknown dtype %d", dtype);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}

extern "C" int qipb_init_kron(qipb_ctx *ctx, void *state, int nbits, int dtype, int ngroups, const int *group_len,
                              const int *group_bits, const void *feeds_dev, uint64_t fixed_mask, uint64_t fixed_value,
                              uint64_t shard_index) {
    QIPB_REQUIRE(ctx && state && (ngroups == 0 || (feeds_dev && group_len && group_bits)), "null argument");
    QIPB_REQUIRE(nbits >= 0 && nbits <= 40, "nbits %d unsupported", nbits);
    QIPB_REQUIRE(ngroups >= 0 && ngroups <= QIPB_MAX_GROUPS, "ngroups %d unsupported (0..%d)", ngroups, QIPB_MAX_GROUPS);
    QIPB_REQUIRE((fixed_value & ~fixed_mask) == 0, "fixed_value has bits outside fixed_mask");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    KronArgs k;
    memset(&k, 0, sizeof(k));
    k.n = 1ull << nbits;
    k.shard_base = (u64)shard_index << nbits;
    k.fixed_mask = fixed_mask;
    k.fixed_value = fixed_value;
    k.ngroups = ngroups;
    k.lb = nbits < KRON_LOW_BITS ? nbits : KRON_LOW_BITS;
    u64 foff = 0, seen = 0;
    const int *gb = group_bits;
    for (int g = 0; g < ngroups; ++g) {
        const int L = group_len[g];
        QIPB_REQUIRE(L >= 1 && L <= 40, "group length %d unsupported", L);
        int src[64], dst[64];
        int nlow = 0;
        for (int t = 0; t < L; ++t) {
            QIPB_REQUIRE(gb[t] >= 0 && gb[t] < 63, "group bit out of range");
            QIPB_REQUIRE(!((seen >> gb[t]) & 1ull), "bit %d fed twice", gb[t]);
            seen |= 1ull << gb[t];
            src[t] = gb[t];
            dst[t] = L - 1 - t;      // first listed qubit = most significant sub-index bit
            nlow += gb[t] < k.lb;
        }
        k.cls[g] = nlow == L ? KRON_LOW : (nlow == 0 ? KRON_HIGH : KRON_MIXED);
        BitRuns one;
        memset(&one, 0, sizeof(one));
        int rc = build_runs(one, L, src, dst);
        if (rc) return rc;
        k.run_begin[g] = k.runs.nruns;
        QIPB_REQUIRE(k.runs.nruns + one.nruns <= QIPB_MAX_RUNS, "feed layout needs more than %d bit runs", QIPB_MAX_RUNS);
        for (int r = 0; r < one.nruns; ++r) {
            k.runs.src[k.runs.nruns] = one.src[r];
            k.runs.dst[k.runs.nruns] = one.dst[r];
            k.runs.len[k.runs.nruns] = one.len[r];
            k.runs.nruns++;
        }
        k.feed_off[g] = foff;
        foff += 1ull << L;
        gb += L;
    }
    k.run_begin[ngroups] = k.runs.nruns;
    QIPB_REQUIRE((seen & fixed_mask) == 0, "fixed_mask overlaps fed bits");
    QIPB_REQUIRE(dtype == QIPB_C128 || dtype == QIPB_C64, "unknown dtype %d", dtype);
    if (!ctx->kron_table) QIPB_CUDA(cudaMalloc(&ctx->kron_table, sizeof(double2) << KRON_LOW_BITS));
    const u64 tblocks = ((1ull << k.lb) + 255) / 256;
    kron_low_table_kernel<<<(unsigned)tblocks, 256, 0, ctx->stream>>>(ctx->kron_table, (const double2 *)feeds_dev, k);
    const u64 blocks = k.n >> k.lb;
    if (dtype == QIPB_C128) init_kron_kernel<double2><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double2 *)state, (const double2 *)feeds_dev, ctx->kron_table, k);
    else init_kron_kernel<float2><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float2 *)state, (const double2 *)feeds_dev, ctx->kron_table, k);
    ctx->launches += 2;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}

template <typename TAB>
static int func_xor_impl(qipb_ctx *ctx, void *state, int nbits, int dtype, int n1, const int *reg1_bits, int n2,
                         const int *reg2_bits, const TAB *table_dev, uint64_t x_fixed) {
    QIPB_REQUIRE(ctx && state && table_dev, "null argument");
    QIPB_REQUIRE(nbits >= 0 && nbits <= 40 && n1 >= 0 && n2 >= 0 && n1 <= 62 && n2 <= nbits && n2 <= 62, "bad register sizes");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    FuncArgs f;
    memset(&f, 0, sizeof(f));
    f.n = 1ull << nbits;
    f.x_fixed = x_fixed;
    f.ymask = (1ull << n2) - 1ull;
    int src[64], dst[64];
    u64 seen = 0;
    int nloc = 0;
    for (int j = 0; j < n1; ++j) {
        if (reg1_bits[j] < 0) continue;   // register bit held by the rank index: its value is in x_fixed
        QIPB_REQUIRE(reg1_bits[j] < nbits && !((seen >> reg1_bits[j]) & 1ull), "bad reg1 bit");
        seen |= 1ull << reg1_bits[j];
        src[nloc] = reg1_bits[j];
        dst[nloc] = n1 - 1 - j;
        nloc++;
    }
    int rc = build_runs(f.xruns, nloc, src, dst);
    if (rc) return rc;
    for (int j = 0; j < n2; ++j) {
        QIPB_REQUIRE(reg2_bits[j] >= 0 && reg2_bits[j] < nbits && !((seen >> reg2_bits[j]) & 1ull), "bad reg2 bit");
        seen |= 1ull << reg2_bits[j];
        src[j] = n2 - 1 - j;
        dst[j] = reg2_bits[j];
    }
    rc = build_runs(f.yruns, n2, src, dst);
    if (rc) return rc;
    const u64 blocks = (f.n + 255) / 256;
    if (dtype == QIPB_C128) func_xor_kernel<double2, TAB><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double2 *)state, table_dev, f);
    else if (dtype == QIPB_C64) func_xor_kernel<float2, TAB><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float2 *)state, table_dev, f);
    else QIPB_REQUIRE(false, "unknown dtype %d", dtype);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}

extern "C" int qipb_func_xor(qipb_ctx *ctx, void *state, int nbits, int dtype, int n1, const int *reg1_bits, int n2,
                             const int *reg2_bits, const long long *table_dev, uint64_t x_fixed) {
    return func_xor_impl<long long>(ctx, state, nbits, dtype, n1, reg1_bits, n2, reg2_bits, table_dev, x_fixed);
}

extern "C" int qipb_func_xor_u8(qipb_ctx *ctx, void *state, int nbits, int dtype, int n1, const int *reg1_bits, int n2,
                                const int *reg2_bits, const unsigned char *table_dev, uint64_t x_fixed) {
    QIPB_REQUIRE(n2 <= 8, "a byte table holds at most 8 output bits (n2 = %d)", n2);
    return func_xor_impl<unsigned char>(ctx, state, nbits, dtype, n1, reg1_bits, n2, reg2_bits, table_dev, x_fixed);
}
