// qip_b200/csrc/exchange.cu -- multi-GPU state exchange over NVLink peer memory for sm_100a.
//
// Replaces the reference's worker<->worker socket exchange (qip/distributed/worker/worker.py:
// 302-357: protobuf chunks of 2048 amplitudes over TCP) and the per-gate reduce-to-diagonal +
// re-broadcast of qip/distributed/manager.py:224-236.  The state is sharded by its top qubits,
// one process per GPU; each process maps its peers' shards with CUDA IPC and the kernels below
// load/store the partner's HBM directly through NVLink 5 / NVSwitch.
//   peer_swap_kernel  : in-place exchange of an amplitude range with the partner (global<->local
//                       qubit swap); one kernel, no staging buffer, traffic in both directions.
//   peer_gate1_kernel : FUSED compute + exchange -- a 1-qubit gate on a global (rank) bit is
//                       applied while the halves cross the link: (lo,hi) <- M (lo,hi).
// Each rank of a pair processes a disjoint half of the range, so both directions of the link and
// both GPUs' SMs are used.  Roofline: NVLink-bound, sizeof(amp)*count/2 bytes out and in per rank.
#include "common.cuh"
#include "../../include/qip_b200.h"

namespace qipb {

template <typename A, int U>
__global__ void __launch_bounds__(256) peer_swap_kernel(A *__restrict__ local, A *__restrict__ peer, u64 count) {
    A a[U], b[U];
    u64 idx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        idx[u] = ((u64)blockIdx.x * U + u) * blockDim.x + threadIdx.x;
        if (idx[u] < count) { a[u] = local[idx[u]]; b[u] = peer[idx[u]]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (idx[u] < count) { local[idx[u]] = b[u]; peer[idx[u]] = a[u]; }
}

// Swap of a rank (global) index bit with local bit `lbit`: the amplitudes whose local bit differs
// from this rank's global-bit value trade places with the partner's amplitudes whose local bit
// equals it.  w enumerates the 2^(nbits-1) indices with bit `lbit` removed.
template <typename A, int U>
__global__ void __launch_bounds__(256) peer_swap_bit_kernel(A *__restrict__ local, A *__restrict__ peer, int lbit, int my_g,
                                                            u64 w_begin, u64 count) {
    A a[U], b[U];
    u64 li[U], pi[U];
    bool live[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const u64 t = ((u64)blockIdx.x * U + u) * blockDim.x + threadIdx.x;
        live[u] = t < count;
        const u64 base = insert_zero(w_begin + t, lbit);
        li[u] = base | ((u64)(1 - my_g) << lbit);
        pi[u] = base | ((u64)my_g << lbit);
        if (live[u]) { a[u] = local[li[u]]; b[u] = peer[pi[u]]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (live[u]) { local[li[u]] = b[u]; peer[pi[u]] = a[u]; }
}

// g rank bits <-> g local bits in ONE kernel.  `a` = this rank's value on the g rank bits; for every
// other value b the sub-block of local indices whose g local bits equal b trades places with the
// sub-block "local bits == a" of the rank whose rank bits equal b.  The 2^(nbits-g) pairs of a
// sub-block pair are indexed by w; the rank with the smaller value handles the first half of w, the
// other one the second half, so every link direction and every GPU carries the same load:
// (1 - 2^-g) of a shard per direction in total, against g/2 shards for g pairwise swaps.
// Chunks: the remap may be restricted to the amplitudes whose `fix` local bits (at most 4, none of them exchanged) have
// a given value -- the sharded engine exchanges a state chunk by chunk so that the NVLink traffic of one chunk runs
// under the fused passes of its neighbours (qipb_peer_remap_chunk).  A chunked launch is PERSISTENT: a bounded number of
// CTAs strides over the work, so that it can share the SMs with a fused pass instead of queueing millions of CTAs.
struct RemapArgs {
    void *peers[8];               // indexed by b
    u64 half;                     // 2^(nbits - g - nfix - 1): pairs per (sub-block pair, rank of the pair)
    u64 nblocks;                  // ceil(half / (256 * U))
    u64 fix_value;                // index bits of the chunk
    int g, a, nins;
    unsigned char lbit[4];        // the local bit positions, lbit[t] pairs with value bit t
    unsigned char ins[8];         // exchanged and fixed positions, ascending (for zero insertion)
};

template <typename A, int U>
__global__ void __launch_bounds__(256) peer_remap_kernel(A *__restrict__ local, const __grid_constant__ RemapArgs r) {
    // partner order is an XOR schedule: CTAs are dispatched slot by slot, and at every step the ranks
    // form disjoint pairs (a <-> a ^ (slot+1)) -- no peer is ever the target of several ranks at once
    // (with "slot -> b" every rank would hit peer 0 first: measured 3.7x slower on 8 GPUs)
    const int b = r.a ^ ((int)blockIdx.y + 1);
    A *__restrict__ peer = reinterpret_cast<A *>(r.peers[b]);
    u64 lsel = 0, psel = 0;
    for (int t = 0; t < r.g; ++t) {
        lsel |= (u64)((b >> t) & 1) << r.lbit[t];
        psel |= (u64)((r.a >> t) & 1) << r.lbit[t];
    }
    const u64 w0 = r.a < b ? 0 : r.half;
    lsel |= r.fix_value;
    psel |= r.fix_value;
    for (u64 blk = blockIdx.x; blk < r.nblocks; blk += gridDim.x) {
        A x[U], y[U];
        u64 li[U], pi[U];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const u64 t = (blk * U + u) * blockDim.x + threadIdx.x;
            live[u] = t < r.half;
            u64 base = w0 + t;
            for (int q = 0; q < r.nins; ++q) base = insert_zero(base, r.ins[q]);
            li[u] = base | lsel;
            pi[u] = base | psel;
            if (live[u]) { x[u] = local[li[u]]; y[u] = peer[pi[u]]; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (live[u]) { local[li[u]] = y[u]; peer[pi[u]] = x[u]; }
    }
}

struct PeerGateArgs {
    u64 count;
    u64 ctrl_mask;        // local-index bits that must be 1
    int local_is_hi;
    double2 m[4];
};

template <typename A, int U>
__global__ void __launch_bounds__(256) peer_gate1_kernel(A *__restrict__ local, A *__restrict__ peer, u64 off,
                                                         const __grid_constant__ PeerGateArgs g) {
    A l[U], p[U];
    u64 idx[U];
    bool live[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const u64 w = ((u64)blockIdx.x * U + u) * blockDim.x + threadIdx.x;
        idx[u] = off + w;
        live[u] = w < g.count && ((idx[u] & g.ctrl_mask) == g.ctrl_mask);
        if (live[u]) { l[u] = local[idx[u]]; p[u] = peer[idx[u]]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (live[u]) {
            const A lo = g.local_is_hi ? p[u] : l[u], hi = g.local_is_hi ? l[u] : p[u];
            A r0 = cmul<A>(g.m[0], lo);
            cfma<A>(r0, g.m[1], hi);
            A r1 = cmul<A>(g.m[2], lo);
            cfma<A>(r1, g.m[3], hi);
            local[idx[u]] = g.local_is_hi ? r1 : r0;
            peer[idx[u]] = g.local_is_hi ? r0 : r1;
        }
}

}  // namespace qipb

using namespace qipb;

extern "C" int qipb_ipc_export(qipb_ctx *ctx, void *dev_ptr, unsigned char handle_out[64]) {
    QIPB_REQUIRE(ctx && dev_ptr && handle_out, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    QIPB_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle_out, &h, 64);
    return QIPB_OK;
}

extern "C" int qipb_ipc_open(qipb_ctx *ctx, const unsigned char handle[64], void **peer_ptr_out) {
    QIPB_REQUIRE(ctx && handle && peer_ptr_out, "null argument");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    QIPB_CUDA(cudaIpcOpenMemHandle(peer_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return QIPB_OK;
}

extern "C" int qipb_ipc_close(qipb_ctx *ctx, void *peer_ptr) {
    QIPB_REQUIRE(ctx && peer_ptr, "null argument");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    QIPB_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return QIPB_OK;
}

extern "C" int qipb_peer_swap(qipb_ctx *ctx, void *local, void *peer, int dtype, uint64_t local_off, uint64_t peer_off,
                              uint64_t count) {
    QIPB_REQUIRE(ctx && local && peer, "null argument");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    if (count == 0) return QIPB_OK;
    const u64 blocks = (count + 256ull * 4 - 1) / (256ull * 4);
    QIPB_REQUIRE(blocks <= 0x7fffffffull, "grid too large");
    if (dtype == QIPB_C128) peer_swap_kernel<double2, 4><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double2 *)local + local_off, (double2 *)peer + peer_off, count);
    else if (dtype == QIPB_C64) peer_swap_kernel<float2, 4><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float2 *)local + local_off, (float2 *)peer + peer_off, count);
    else QIPB_REQUIRE(false, "unknown dtype %d", dtype);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}

extern "C" int qipb_peer_swap_bit(qipb_ctx *ctx, void *local, void *peer, int nbits, int dtype, int lbit, int my_gbit,
                                  uint64_t w_begin, uint64_t count) {
    QIPB_REQUIRE(ctx && local && peer, "null argument");
    QIPB_REQUIRE(nbits >= 1 && nbits <= 40 && lbit >= 0 && lbit < nbits && (my_gbit == 0 || my_gbit == 1), "bad peer_swap_bit arguments");
    QIPB_REQUIRE(w_begin + count <= (1ull << (nbits - 1)), "peer_swap_bit range exceeds 2^(nbits-1)");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    if (count == 0) return QIPB_OK;
    const u64 blocks = (count + 256ull * 4 - 1) / (256ull * 4);
    QIPB_REQUIRE(blocks <= 0x7fffffffull, "grid too large");
    if (dtype == QIPB_C128) peer_swap_bit_kernel<double2, 4><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double2 *)local, (double2 *)peer, lbit, my_gbit, w_begin, count);
    else if (dtype == QIPB_C64) peer_swap_bit_kernel<float2, 4><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float2 *)local, (float2 *)peer, lbit, my_gbit, w_begin, count);
    else QIPB_REQUIRE(false, "unknown dtype %d", dtype);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}

static int peer_remap_impl(qipb_ctx *ctx, void *local, void *const *peers, int nbits, int dtype, int g, const int *lbits,
                           int my_value, int nfix, const int *fix_bits, uint64_t fix_value, int max_ctas) {
    QIPB_REQUIRE(ctx && local && peers && lbits, "null argument");
    QIPB_REQUIRE(g >= 1 && g <= 3 && nfix >= 0 && nfix <= 4 && (nfix == 0 || fix_bits) && nbits > g + nfix && nbits <= 40 &&
                 my_value >= 0 && my_value < (1 << g), "bad peer_remap arguments");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    RemapArgs r;
    memset(&r, 0, sizeof(r));
    r.g = g;
    r.a = my_value;
    r.half = 1ull << (nbits - g - nfix - 1);
    u64 seen = 0, fmask = 0;
    for (int t = 0; t < g; ++t) {
        QIPB_REQUIRE(lbits[t] >= 0 && lbits[t] < nbits && !((seen >> lbits[t]) & 1ull), "bad local bit %d", lbits[t]);
        seen |= 1ull << lbits[t];
        r.lbit[t] = (unsigned char)lbits[t];
    }
    for (int t = 0; t < nfix; ++t) {
        QIPB_REQUIRE(fix_bits[t] >= 0 && fix_bits[t] < nbits && !((seen >> fix_bits[t]) & 1ull), "bad fixed bit %d", fix_bits[t]);
        seen |= 1ull << fix_bits[t];
        fmask |= 1ull << fix_bits[t];
    }
    QIPB_REQUIRE((fix_value & ~fmask) == 0, "chunk value has bits outside the fixed bits");
    r.fix_value = fix_value;
    for (int b = 0; b < nbits; ++b)
        if ((seen >> b) & 1ull) r.ins[r.nins++] = (unsigned char)b;
    for (int b = 0; b < (1 << g); ++b) {
        QIPB_REQUIRE(b == my_value || peers[b], "missing peer pointer for value %d", b);
        r.peers[b] = peers[b];
    }
    r.nblocks = (r.half + 256ull * 4 - 1) / (256ull * 4);
    u64 bx = r.nblocks;
    if (max_ctas > 0 && bx > (u64)max_ctas) bx = (u64)max_ctas;        // persistent: blockIdx.x strides over nblocks
    QIPB_REQUIRE(bx <= 0x7fffffffull, "grid too large");
    dim3 grid((unsigned)bx, (unsigned)((1 << g) - 1));
    if (dtype == QIPB_C128) peer_remap_kernel<double2, 4><<<grid, 256, 0, ctx->stream>>>((double2 *)local, r);
    else if (dtype == QIPB_C64) peer_remap_kernel<float2, 4><<<grid, 256, 0, ctx->stream>>>((float2 *)local, r);
    else QIPB_REQUIRE(false, "unknown dtype %d", dtype);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}

extern "C" int qipb_peer_remap(qipb_ctx *ctx, void *local, void *const *peers, int nbits, int dtype, int g,
                               const int *lbits, int my_value) {
    return peer_remap_impl(ctx, local, peers, nbits, dtype, g, lbits, my_value, 0, nullptr, 0, 0);
}

extern "C" int qipb_peer_remap_chunk(qipb_ctx *ctx, void *local, void *const *peers, int nbits, int dtype, int g,
                                     const int *lbits, int my_value, int nfix, const int *fix_bits, uint64_t fix_value,
                                     int max_ctas) {
    return peer_remap_impl(ctx, local, peers, nbits, dtype, g, lbits, my_value, nfix, fix_bits, fix_value, max_ctas);
}

extern "C" int qipb_peer_gate1(qipb_ctx *ctx, void *local, void *peer, int dtype, uint64_t off, uint64_t count,
                               const double *mat, int local_is_hi, uint64_t ctrl_mask) {
    QIPB_REQUIRE(ctx && local && peer && mat, "null argument");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    if (count == 0) return QIPB_OK;
    PeerGateArgs g;
    memset(&g, 0, sizeof(g));
    g.count = count;
    g.ctrl_mask = ctrl_mask;
    g.local_is_hi = local_is_hi;
    for (int e = 0; e < 4; ++e) g.m[e] = make_double2(mat[2 * e], mat[2 * e + 1]);
    const u64 blocks = (count + 256ull * 4 - 1) / (256ull * 4);
    QIPB_REQUIRE(blocks <= 0x7fffffffull, "grid too large");
    if (dtype == QIPB_C128) peer_gate1_kernel<double2, 4><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double2 *)local, (double2 *)peer, off, g);
    else if (dtype == QIPB_C64) peer_gate1_kernel<float2, 4><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float2 *)local, (float2 *)peer, off, g);
    else QIPB_REQUIRE(false, "unknown dtype %d", dtype);
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}
