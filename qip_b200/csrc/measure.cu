// qip_b200/csrc/measure.cu -- probability reductions, collapse and compaction for sm_100a.
//
// Replaces the serial sweeps of qip/ext/kronprod.pyx: prob_magnitude (:320-326),
// measure_probabilities (:240-261), the per-outcome sums inside soft_measure (:371-376) and
// measure_top_probabilities (:290-296), the collapse sweep of measure (:439-444) and the gather of
// reduce_measure (:478-489).
//
// Probabilities.  The index is split at bit CB = 8: a CTA of 256 threads reads 256 consecutive
// amplitudes per step (4 KiB coalesced for complex128).  Measured bits >= CB and the slice of the
// un-measured high bits a CTA walks are fixed per work item, so every thread accumulates exactly
// ONE outcome bin in a register while it streams; measured bits < CB are folded at the end with a
// fixed-order shared-memory sum.  Per-slice partials are combined by a second fixed-order kernel:
// no floating-point atomics anywhere, results are bit-reproducible run to run.
// Roofline: HBM-bound, sizeof(amp) * 2^nbits bytes read once (less under a filter).
#include "common.cuh"
#include "../../include/qip_b200.h"

namespace qipb {

#define PROB_CB 8
#define PROB_THREADS 256

struct ProbArgs {
    int nbits;
    int k_lo;             // measured bits below CB
    u64 lo_meas_mask;     // measured positions < CB
    u64 lo_rest_mask;     // un-measured, un-filtered positions < min(CB, nbits)
    u64 hi_meas_mask;     // measured positions >= CB
    u64 hi_rest_mask;     // un-measured, un-filtered positions in [CB, nbits)
    u64 filter_mask, filter_value;
    u64 n_hi_meas;        // 2^popcount(hi_meas_mask)
    u64 slices;           // S: how many CTAs share one hi-measured value
    u64 slice_len;        // steps per slice
    u64 nbins;            // 2^k
    BitRuns outmap;       // state index -> output bin
};

template <typename A>
__global__ void __launch_bounds__(PROB_THREADS) prob_kernel(const A *__restrict__ state, double *__restrict__ dst,
                                                            const __grid_constant__ ProbArgs p) {
    __shared__ double red[PROB_THREADS];
    const u64 tid = threadIdx.x;
    const u64 lowspan = (p.nbits >= PROB_CB) ? (u64)PROB_THREADS : (1ull << p.nbits);
    const u64 lowmask = lowspan - 1ull;
    const bool lane_ok = tid < lowspan && ((tid & p.filter_mask & lowmask) == (p.filter_value & lowmask));
    const u64 items = p.n_hi_meas * p.slices;
    for (u64 item = blockIdx.x; item < items; item += gridDim.x) {
        const u64 mh = item / p.slices, s = item - mh * p.slices;
        const u64 base = deposit_bits(mh, p.hi_meas_mask) | (p.filter_value & ~lowmask) | tid;
        u64 r = deposit_bits(s * p.slice_len, p.hi_rest_mask);
        double acc = 0.0;
        if (lane_ok) {
            u64 it = 0;
            for (; it + 4 <= p.slice_len; it += 4) {
                const u64 r0 = r, r1 = masked_inc(r0, p.hi_rest_mask), r2 = masked_inc(r1, p.hi_rest_mask),
                          r3 = masked_inc(r2, p.hi_rest_mask);
                r = masked_inc(r3, p.hi_rest_mask);
                const A a0 = state[base | r0], a1 = state[base | r1], a2 = state[base | r2], a3 = state[base | r3];
                acc += norm2(a0);
                acc += norm2(a1);
                acc += norm2(a2);
                acc += norm2(a3);
            }
            for (; it < p.slice_len; ++it) {
                acc += norm2(state[base | r]);
                r = masked_inc(r, p.hi_rest_mask);
            }
        }
        red[tid] = acc;
        __syncthreads();
        if (tid < (1ull << p.k_lo)) {
            const u64 mine = deposit_bits(tid, p.lo_meas_mask);
            const u64 nrest = 1ull << __popcll(p.lo_rest_mask);
            double sum = 0.0;
            u64 rr = 0;
            for (u64 j = 0; j < nrest; ++j) {
                sum += red[mine | rr | (p.filter_value & lowmask)];
                rr = masked_inc(rr, p.lo_rest_mask);
            }
            const u64 o = runs_gather(p.outmap, (base & ~lowmask) | mine);
            dst[s * p.nbins + o] = sum;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) prob_finish_kernel(const double *__restrict__ partial, double *__restrict__ out,
                                                          u64 nbins, u64 slices) {
    const u64 o = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nbins) return;
    double sum = 0.0;
    for (u64 s = 0; s < slices; ++s) sum += partial[s * nbins + o];
    out[o] = sum;
}

// Single-bin finish (k = 0, many slices): fixed-order block tree.
__global__ void __launch_bounds__(256) prob_finish_scalar_kernel(const double *__restrict__ partial, double *__restrict__ out,
                                                                 u64 slices) {
    __shared__ double red[256];
    double sum = 0.0;
    for (u64 s = threadIdx.x; s < slices; s += 256) sum += partial[s];
    red[threadIdx.x] = sum;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

template <typename A, int U>
__global__ void __launch_bounds__(256) collapse_kernel(A *__restrict__ state, u64 n, u64 mask, u64 want, double scale) {
    typedef typename amp_traits<A>::real R;
    const R sc = (R)scale;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const u64 i = ((u64)blockIdx.x * U + u) * blockDim.x + threadIdx.x;
        if (i < n) {
            if ((i & mask) == want) {
                A a = state[i];
                a.x *= sc;
                a.y *= sc;
                state[i] = a;
            } else {
                state[i] = make_amp<A>(0, 0);
            }
        }
    }
}

template <typename A>
__global__ void __launch_bounds__(256) reduce_kernel(const A *__restrict__ src, A *__restrict__ dst, u64 nout, u64 want,
                                                     double scale, const __grid_constant__ BitRuns spread) {
    typedef typename amp_traits<A>::real R;
    const R sc = (R)scale;
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nout) return;
    A a = src[runs_gather(spread, j) | want];
    a.x *= sc;
    a.y *= sc;
    dst[j] = a;
}

template <typename A>
__global__ void __launch_bounds__(256) add_range_kernel(A *__restrict__ state, const A *__restrict__ data, u64 count) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    A a = state[i];
    const A b = data[i];
    a.x += b.x;
    a.y += b.y;
    state[i] = a;
}

static int ensure_scratch(qipb_ctx *ctx, size_t bytes) {
    if (ctx->scratch_bytes >= bytes) return QIPB_OK;
    if (ctx->scratch) QIPB_CUDA(cudaFree(ctx->scratch));
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    QIPB_CUDA(cudaMalloc(&ctx->scratch, bytes));
    ctx->scratch_bytes = bytes;
    return QIPB_OK;
}

template <typename A>
static int probabilities_t(qipb_ctx *ctx, const A *state, int nbits, int k, const int *bits, const int *out_bits,
                           u64 fmask, u64 fval, double *out_dev) {
    ProbArgs p;
    memset(&p, 0, sizeof(p));
    p.nbits = nbits;
    const u64 all = (nbits >= 64) ? ~0ull : ((1ull << nbits) - 1ull);
    const int cb = nbits < PROB_CB ? nbits : PROB_CB;
    const u64 lowmask = (1ull << cb) - 1ull;
    u64 meas = 0, outseen = 0;
    for (int j = 0; j < k; ++j) {
        QIPB_REQUIRE(bits[j] >= 0 && bits[j] < nbits, "measured bit %d out of range", bits[j]);
        QIPB_REQUIRE(out_bits[j] >= 0 && out_bits[j] < k, "output bit %d out of range", out_bits[j]);
        QIPB_REQUIRE(!((meas >> bits[j]) & 1ull) && !((outseen >> out_bits[j]) & 1ull), "repeated measured/output bit");
        meas |= 1ull << bits[j];
        outseen |= 1ull << out_bits[j];
    }
    QIPB_REQUIRE((fmask & meas) == 0 && (fmask & ~all) == 0 && (fval & ~fmask) == 0, "bad filter");
    p.filter_mask = fmask;
    p.filter_value = fval;
    p.lo_meas_mask = meas & lowmask;
    p.hi_meas_mask = meas & ~lowmask;
    p.lo_rest_mask = lowmask & ~meas & ~fmask;
    p.hi_rest_mask = all & ~lowmask & ~meas & ~fmask;
    p.k_lo = __builtin_popcountll(p.lo_meas_mask);
    p.n_hi_meas = 1ull << __builtin_popcountll(p.hi_meas_mask);
    p.nbins = 1ull << k;
    int rc = build_runs(p.outmap, k, bits, out_bits);
    if (rc) return rc;
    const u64 R = 1ull << __builtin_popcountll(p.hi_rest_mask);
    const u64 target = (u64)ctx->sm_count * 16;
    u64 S = 1;
    while (p.n_hi_meas * S < target && S * 2 * 8 <= R) S <<= 1;
    p.slices = S;
    p.slice_len = R / S;
    const u64 items = p.n_hi_meas * S;
    const u64 grid = items < target * 2 ? items : target * 2;
    double *dst = out_dev;
    if (S > 1) {
        rc = ensure_scratch(ctx, sizeof(double) * S * p.nbins);
        if (rc) return rc;
        dst = ctx->scratch;
    }
    prob_kernel<A><<<(unsigned)grid, PROB_THREADS, 0, ctx->stream>>>(state, dst, p);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    if (S > 1) {
        if (p.nbins == 1) {
            prob_finish_scalar_kernel<<<1, 256, 0, ctx->stream>>>(dst, out_dev, S);
        } else {
            prob_finish_kernel<<<(unsigned)((p.nbins + 255) / 256), 256, 0, ctx->stream>>>(dst, out_dev, p.nbins, S);
        }
        ctx->launches++;
        QIPB_CUDA(cudaGetLastError());
    }
    return QIPB_OK;
}

}  // namespace qipb

using namespace qipb;

extern "C" int qipb_probabilities(qipb_ctx *ctx, const void *state, int nbits, int dtype, int k, const int *bits,
                                  const int *out_bits, uint64_t filter_mask, uint64_t filter_value, double *out_dev) {
    QIPB_REQUIRE(ctx && state && out_dev, "null argument");
    QIPB_REQUIRE(nbits >= 0 && nbits <= 40 && k >= 0 && k <= nbits, "bad nbits/k");
    QIPB_REQUIRE(k == 0 || (bits && out_bits), "null bits");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    if (dtype == QIPB_C128) return probabilities_t<double2>(ctx, (const double2 *)state, nbits, k, bits, out_bits, filter_mask, filter_value, out_dev);
    if (dtype == QIPB_C64) return probabilities_t<float2>(ctx, (const float2 *)state, nbits, k, bits, out_bits, filter_mask, filter_value, out_dev);
    QIPB_REQUIRE(false, "unknown dtype %d", dtype);
}

extern "C" int qipb_collapse(qipb_ctx *ctx, void *state, int nbits, int dtype, uint64_t mask, uint64_t want, double scale) {
    QIPB_REQUIRE(ctx && state, "null argument");
    QIPB_REQUIRE(nbits >= 0 && nbits <= 40 && (want & ~mask) == 0, "bad collapse arguments");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    const u64 n = 1ull << nbits;
    const u64 blocks = (n + 256ull * 4 - 1) / (256ull * 4);
    QIPB_REQUIRE(blocks <= 0x7fffffffull, "grid too large");
    if (dtype == QIPB_C128) collapse_kernel<double2, 4><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double2 *)state, n, mask, want, scale);
    else if (dtype == QIPB_C64) collapse_kernel<float2, 4><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float2 *)state, n, mask, want, scale);
    else QIPB_REQUIRE(false, "unknown dtype %d", dtype);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}

extern "C" int qipb_reduce(qipb_ctx *ctx, const void *src, void *dst, int nbits, int dtype, uint64_t mask, uint64_t want, double scale) {
    QIPB_REQUIRE(ctx && src && dst && src != dst, "null or aliased argument");
    QIPB_REQUIRE(nbits >= 0 && nbits <= 40 && (want & ~mask) == 0, "bad reduce arguments");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    int srcb[64], dstb[64], nrest = 0;
    for (int b = 0; b < nbits; ++b)
        if (!((mask >> b) & 1ull)) { srcb[nrest] = nrest; dstb[nrest] = b; nrest++; }
    BitRuns spread;
    memset(&spread, 0, sizeof(spread));
    int rc = build_runs(spread, nrest, srcb, dstb);
    if (rc) return rc;
    const u64 nout = 1ull << nrest;
    const u64 blocks = (nout + 255) / 256;
    QIPB_REQUIRE(blocks <= 0x7fffffffull, "grid too large");
    if (dtype == QIPB_C128) reduce_kernel<double2><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const double2 *)src, (double2 *)dst, nout, want, scale, spread);
    else if (dtype == QIPB_C64) reduce_kernel<float2><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const float2 *)src, (float2 *)dst, nout, want, scale, spread);
    else QIPB_REQUIRE(false, "unknown dtype %d", dtype);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}

extern "C" int qipb_add_range(qipb_ctx *ctx, void *state, int dtype, uint64_t start, uint64_t count, const void *data_dev) {
    QIPB_REQUIRE(ctx && state && data_dev, "null argument");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    if (count == 0) return QIPB_OK;
    const u64 blocks = (count + 255) / 256;
    QIPB_REQUIRE(blocks <= 0x7fffffffull, "grid too large");
    if (dtype == QIPB_C128) add_range_kernel<double2><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double2 *)state + start, (const double2 *)data_dev, count);
    else if (dtype == QIPB_C64) add_range_kernel<float2><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float2 *)state + start, (const float2 *)data_dev, count);
    else QIPB_REQUIRE(false, "unknown dtype %d", dtype);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}
