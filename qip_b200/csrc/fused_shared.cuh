// qip_b200/csrc/fused_shared.cuh -- mbarrier / bulk-copy primitives and host-side stage types of the fused pass (fused.cu).
#pragma once
#include <complex>
#include <vector>
#include "common.cuh"
#include "../../include/qip_b200.h"

namespace qipb {

#define FUSED_OUT_CELLS 4
#define FUSED_LO_BITS 6           // tile-local bits 0..5 form the 'lo' table cell, the rest the 'hi' cell

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
__device__ __forceinline__ void mbar_arrive(void *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src_gmem, u32 bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");
}


// diag == 2: a run of diagonal gates folded into per-cell phase tables (stored over DevGate::m)
struct StageInfo {
    u32 tab_off;                              // offset of T_lo in the table buffer (double2 units)
    unsigned char nout;                       // number of cells made of bits outside the tile
    unsigned char cn[FUSED_OUT_CELLS];        // bits per outside cell
    unsigned char cb[FUSED_OUT_CELLS][7];     // their state-index positions, ascending
};

// Product of a stage's outside-cell tables at this tile's base index (uniform per tile).
QIPB_HD double2 stage_scalar(const StageInfo &si, const double2 *__restrict__ T, u64 base, u32 nlo, u32 nhi) {
    double2 S = make_double2(1.0, 0.0);
    const double2 *To = T + nlo + nhi;
    for (int c = 0; c < si.nout; ++c) {
        u32 idx = 0;
        for (int j = 0; j < si.cn[c]; ++j) idx |= (u32)((base >> si.cb[c][j]) & 1ull) << j;
        S = cmul<double2>(S, To[idx]);
        To += 1u << si.cn[c];
    }
    return S;
}

// ---- host side ------------------------------------------------------------------------------
typedef std::complex<double> cplx;

struct Op {
    bool stage;
    int gate;                 // !stage: index into the caller's gate list
    u64 common;               // stage: control bits shared by every gate of the stage
    u64 support;              // stage: every bit one of its gates reads (targets and controls); pairing / commutation checks
    u32 tab_off;              // stage: offset of its tables in the table buffer
    int nout;
    std::vector<int> cells[FUSED_OUT_CELLS];
};


bool stages_enabled();
void build_stages(const qipb_gate *gates, const std::vector<int> &run, int nbits, int tb, const int *local_of,
                  u64 tmask, std::vector<Op> &ops, std::vector<cplx> &tables, int min_run);
int upload_tables(qipb_ctx *ctx, const std::vector<cplx> &tables, const double2 **out);

}  // namespace qipb
