// qip_b200/csrc/common.cuh -- shared device/host helpers for the sm_100a state-vector kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace qipb {

typedef unsigned long long u64;
typedef unsigned int u32;

// Sweep arithmetic and index helpers are host+device so that tests/csrc/fused_emul.cu can run the very same
// code on a CPU tile (the product never executes them on the host).
#ifndef QIPB_HD      /* (tests/csrc/fused_emul.cu overrides it: forced inlining of every sweep makes the HOST compile take minutes) */
#define QIPB_HD __host__ __device__ __forceinline__
#endif

// ---- amplitude types -------------------------------------------------------------------
// complex128 amplitude = double2 (one 128-bit transaction), complex64 = float2 (64-bit).
template <typename A> struct amp_traits;
template <> struct amp_traits<double2> { typedef double real; };
template <> struct amp_traits<float2>  { typedef float  real; };

template <typename A> QIPB_HD A make_amp(typename amp_traits<A>::real x, typename amp_traits<A>::real y);
template <> QIPB_HD double2 make_amp<double2>(double x, double y) { return make_double2(x, y); }
template <> QIPB_HD float2  make_amp<float2>(float x, float y)    { return make_float2(x, y); }

// acc += m * a, with m a (double) matrix coefficient converted to the amplitude precision.
template <typename A>
__device__ __forceinline__ void cfma(A &acc, const double2 m, const A a) {
    typedef typename amp_traits<A>::real R;
    const R mr = (R)m.x, mi = (R)m.y;
    acc.x = fma(mr, a.x, acc.x);
    acc.x = fma(-mi, a.y, acc.x);
    acc.y = fma(mr, a.y, acc.y);
    acc.y = fma(mi, a.x, acc.y);
}
template <typename A>
QIPB_HD A cmul(const double2 m, const A a) {
    typedef typename amp_traits<A>::real R;
    const R mr = (R)m.x, mi = (R)m.y;
    A r;
    r.x = mr * a.x - mi * a.y;
    r.y = mr * a.y + mi * a.x;
    return r;
}
template <typename A>
__device__ __forceinline__ double norm2(const A a) {
    return (double)a.x * (double)a.x + (double)a.y * (double)a.y;
}

// ---- bit helpers -----------------------------------------------------------------------
// Insert a zero bit at position p (bits >= p move up by one).
QIPB_HD u64 insert_zero(u64 v, int p) {
    const u64 lo = v & ((1ull << p) - 1ull);
    return ((v >> p) << (p + 1)) | lo;
}

// A gather/scatter of bit-fields described as runs of consecutive bits:
//   gather : out |= ((v >> src) & mask(len)) << dst      for every run
// Registers of consecutive qubits collapse to one run, so the common case is 1-2 shifts.
#define QIPB_MAX_RUNS 40
struct BitRuns {
    int nruns;
    unsigned char src[QIPB_MAX_RUNS];
    unsigned char dst[QIPB_MAX_RUNS];
    unsigned char len[QIPB_MAX_RUNS];
};
__device__ __forceinline__ u64 runs_gather(const BitRuns &r, u64 v) {
    u64 out = 0;
    for (int i = 0; i < r.nruns; ++i)
        out |= ((v >> r.src[i]) & ((1ull << r.len[i]) - 1ull)) << r.dst[i];
    return out;
}

// Next value of a counter that only lives on the bits of `mask` (all other bits stay 0).
__device__ __forceinline__ u64 masked_inc(u64 v, u64 mask) { return ((v | ~mask) + 1ull) & mask; }

// Deposit the low bits of v onto the set bits of mask (software pdep), low to high.
__host__ __device__ inline u64 deposit_bits(u64 v, u64 mask) {
    u64 out = 0;
    for (u64 m = mask; m; m &= m - 1) {
        const u64 low = m & (~m + 1ull);
        if (v & 1ull) out |= low;
        v >>= 1;
    }
    return out;
}

// ---- error plumbing ----------------------------------------------------------------------
void set_error(const char *fmt, ...);
#define QIPB_OK 0
#define QIPB_ERR_ARG 1
#define QIPB_ERR_CUDA 2
#define QIPB_ERR_UNSUPPORTED 3      // the request is valid but this entry point cannot serve it (caller falls back); nothing was launched
#define QIPB_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            qipb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return QIPB_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)
#define QIPB_REQUIRE(cond, ...)                                                          \
    do {                                                                                 \
        if (!(cond)) { qipb::set_error(__VA_ARGS__); return QIPB_ERR_ARG; }              \
    } while (0)

// Merge single-bit (src -> dst) moves into runs of consecutive bits.
static inline int build_runs(BitRuns &r, int n, const int *src, const int *dst) {
    // sort by src ascending (insertion sort; n <= 64)
    int order[64];
    for (int i = 0; i < n; ++i) order[i] = i;
    for (int i = 1; i < n; ++i)
        for (int j = i; j > 0 && src[order[j]] < src[order[j - 1]]; --j) { int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
    r.nruns = 0;
    for (int t = 0; t < n; ++t) {
        const int s = src[order[t]], d = dst[order[t]];
        if (r.nruns > 0) {
            const int q = r.nruns - 1;
            if (r.src[q] + r.len[q] == s && r.dst[q] + r.len[q] == d) { r.len[q]++; continue; }
        }
        QIPB_REQUIRE(r.nruns < QIPB_MAX_RUNS, "bit map needs more than %d runs", QIPB_MAX_RUNS);
        r.src[r.nruns] = (unsigned char)s;
        r.dst[r.nruns] = (unsigned char)d;
        r.len[r.nruns] = 1;
        r.nruns++;
    }
    return QIPB_OK;
}

}  // namespace qipb

// Context shared by all entry points (opaque in the C header).
struct qipb_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;
    double *scratch;        // device scratch for reduction partials
    size_t scratch_bytes;
    unsigned long long launches;   // kernels launched through this context
    unsigned long long ring_launches;  // of which: persistent ring kernel of the fused pass (fused.cu)
    // diagonal-stage tables of the fused pass (fused.cu): device buffer + pinned staging ring
    double2 *tab_dev;
    size_t tab_cap;                // capacity in double2 elements (device buffer and every ring slot)
    double2 *tab_host[4];
    cudaEvent_t tab_ev[4];
    int tab_slot;
    double2 *kron_table;           // init.cu: product of the low-bit feed groups, 2^12 entries
    unsigned long long ext_launches;   // fused launches that took the EXT kernel (opt-in forms, fused.cu)
    // fused.cu: tile counters of the dynamically scheduled launches, a ring of {next tile, CTAs done} pairs (device
    // memory, zero when idle: the last CTA of a launch resets its pair)
    unsigned int *sched_ring;
    unsigned int sched_slot;
};
#define QIPB_SCHED_SLOTS 256
