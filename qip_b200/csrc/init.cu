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

struct KronArgs {
    u64 n;
    u64 shard_base;       // shard_index << nbits
    u64 zero_mask;        // global bits that must be 0 (un-fed qubits)
    int ngroups;
    int vb;               // a thread writes 2^vb amplitudes: index bits 8 .. 8+vb-1 vary inside a thread
    int run_begin[QIPB_MAX_GROUPS + 1];
    u64 feed_off[QIPB_MAX_GROUPS];
    unsigned char varies[QIPB_MAX_GROUPS];   // group reads one of the bits that vary inside a thread
    BitRuns runs;         // all groups' gathers, concatenated
};

__device__ __forceinline__ double2 kron_factor(const KronArgs &k, const double2 *__restrict__ feeds, int g, u64 G) {
    u64 sub = 0;
    for (int r = k.run_begin[g]; r < k.run_begin[g + 1]; ++r)
        sub |= ((G >> k.runs.src[r]) & ((1ull << k.runs.len[r]) - 1ull)) << k.runs.dst[r];
    return feeds[k.feed_off[g] + sub];
}

// A block of 256 threads writes 256 * 2^vb consecutive amplitudes; thread t owns t, t+256, t+512, ...
// (every store instruction of a warp is 512 contiguous bytes).  The factors of the groups that do
// not depend on the bits varying inside a thread are multiplied once per thread (left to right, like
// qip/backend.py:98-101), the others once per amplitude -- for n one-qubit feeds that is n/2^vb + vb
// complex multiplies per amplitude instead of n.
template <typename A>
__global__ void __launch_bounds__(256) init_kron_kernel(A *__restrict__ state, const double2 *__restrict__ feeds,
                                                        const __grid_constant__ KronArgs k) {
    typedef typename amp_traits<A>::real R;
    const int vec = 1 << k.vb;
    const u64 i0 = ((u64)blockIdx.x << (8 + k.vb)) + threadIdx.x;
    if (i0 >= k.n) return;
    const u64 G0 = k.shard_base | i0;
    double2 P = make_double2(1.0, 0.0);
    for (int g = 0; g < k.ngroups; ++g)
        if (!k.varies[g]) {
            const double2 f = kron_factor(k, feeds, g, G0);
            const double2 t = P;
            P.x = t.x * f.x - t.y * f.y;
            P.y = t.x * f.y + t.y * f.x;
        }
    for (int j = 0; j < vec; ++j) {
        const u64 i = i0 + ((u64)j << 8);
        if (i >= k.n) break;
        const u64 G = k.shard_base | i;
        double2 v = make_double2(0.0, 0.0);
        if ((G & k.zero_mask) == 0) {
            v = P;
            for (int g = 0; g < k.ngroups; ++g)
                if (k.varies[g]) {
                    const double2 f = kron_factor(k, feeds, g, G);
                    const double2 t = v;
                    v.x = t.x * f.x - t.y * f.y;
                    v.y = t.x * f.y + t.y * f.x;
                }
        }
        state[i] = make_amp<A>((R)v.x, (R)v.y);
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

template <typename A>
__global__ void __launch_bounds__(256) func_xor_kernel(A *__restrict__ state, const long long *__restrict__ table,
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
                              const int *group_bits, const void *feeds_dev, uint64_t zero_mask, uint64_t shard_index) {
    QIPB_REQUIRE(ctx && state && feeds_dev && group_len && group_bits, "null argument");
    QIPB_REQUIRE(nbits >= 0 && nbits <= 40, "nbits %d unsupported", nbits);
    QIPB_REQUIRE(ngroups >= 1 && ngroups <= QIPB_MAX_GROUPS, "ngroups %d unsupported (1..%d)", ngroups, QIPB_MAX_GROUPS);
    QIPB_CUDA(cudaSetDevice(ctx->device));
    KronArgs k;
    memset(&k, 0, sizeof(k));
    k.n = 1ull << nbits;
    k.shard_base = (u64)shard_index << nbits;
    k.zero_mask = zero_mask;
    k.ngroups = ngroups;
    k.vb = nbits >= 11 ? 3 : (nbits > 8 ? nbits - 8 : 0);
    const u64 vmask = ((1ull << k.vb) - 1ull) << 8;
    u64 foff = 0, seen = 0;
    const int *gb = group_bits;
    for (int g = 0; g < ngroups; ++g) {
        const int L = group_len[g];
        QIPB_REQUIRE(L >= 1 && L <= 40, "group length %d unsupported", L);
        int src[64], dst[64];
        for (int t = 0; t < L; ++t) {
            QIPB_REQUIRE(gb[t] >= 0 && gb[t] < 63, "group bit out of range");
            QIPB_REQUIRE(!((seen >> gb[t]) & 1ull), "bit %d fed twice", gb[t]);
            seen |= 1ull << gb[t];
            src[t] = gb[t];
            dst[t] = L - 1 - t;      // first listed qubit = most significant sub-index bit
            if ((vmask >> gb[t]) & 1ull) k.varies[g] = 1;
        }
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
    QIPB_REQUIRE((seen & zero_mask) == 0, "zero_mask overlaps fed bits");
    const u64 per_block = 256ull << k.vb;
    const u64 blocks = (k.n + per_block - 1) / per_block;
    if (dtype == QIPB_C128) init_kron_kernel<double2><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double2 *)state, (const double2 *)feeds_dev, k);
    else if (dtype == QIPB_C64) init_kron_kernel<float2><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float2 *)state, (const double2 *)feeds_dev, k);
    else QIPB_REQUIRE(false, "unknown dtype %d", dtype);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}

extern "C" int qipb_func_xor(qipb_ctx *ctx, void *state, int nbits, int dtype, int n1, const int *reg1_bits, int n2,
                             const int *reg2_bits, const long long *table_dev, uint64_t x_fixed) {
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
    if (dtype == QIPB_C128) func_xor_kernel<double2><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double2 *)state, table_dev, f);
    else if (dtype == QIPB_C64) func_xor_kernel<float2><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float2 *)state, table_dev, f);
    else QIPB_REQUIRE(false, "unknown dtype %d", dtype);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}
