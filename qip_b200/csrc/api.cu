// qip_b200/csrc/api.cu -- context, memory helpers and error plumbing of the C ABI (include/qip_b200.h).
#include <stdarg.h>
#include "common.cuh"
#include "../../include/qip_b200.h"

namespace qipb {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace qipb

using namespace qipb;

extern "C" int qipb_version(void) { return 100; }

extern "C" const char *qipb_last_error(void) { return g_err; }

extern "C" int qipb_create(int device, qipb_ctx **out) {
    QIPB_REQUIRE(out, "null argument");
    int ndev = 0;
    QIPB_CUDA(cudaGetDeviceCount(&ndev));
    QIPB_REQUIRE(device >= 0 && device < ndev, "device %d not available (%d visible)", device, ndev);
    QIPB_CUDA(cudaSetDevice(device));
    // single attributes, not cudaGetDeviceProperties (tens of milliseconds: a context is made per run())
    int major = 0, minor = 0, sms = 0;
    QIPB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    QIPB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    QIPB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    QIPB_REQUIRE(major >= 10, "libqipb200 is built for sm_100a only; device %d is sm_%d%d", device, major, minor);
    qipb_ctx *c = new qipb_ctx();
    c->device = device;
    c->sm_count = sms;
    c->stream = 0;
    c->scratch = nullptr;
    c->scratch_bytes = 0;
    c->launches = 0;
    c->ring_launches = 0;
    c->ext_launches = 0;
    c->tab_dev = nullptr;
    c->tab_cap = 0;
    c->tab_slot = 0;
    c->kron_table = nullptr;
    c->sched_ring = nullptr;
    c->sched_slot = 0;
    for (int i = 0; i < 4; ++i) { c->tab_host[i] = nullptr; c->tab_ev[i] = nullptr; }
    *out = c;
    return QIPB_OK;
}

extern "C" int qipb_destroy(qipb_ctx *ctx) {
    if (!ctx) return QIPB_OK;
    cudaSetDevice(ctx->device);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->tab_dev) cudaFree(ctx->tab_dev);
    if (ctx->kron_table) cudaFree(ctx->kron_table);
    if (ctx->sched_ring) cudaFree(ctx->sched_ring);
    for (int i = 0; i < 4; ++i) {
        if (ctx->tab_host[i]) cudaFreeHost(ctx->tab_host[i]);
        if (ctx->tab_ev[i]) cudaEventDestroy(ctx->tab_ev[i]);
    }
    delete ctx;
    return QIPB_OK;
}

extern "C" int qipb_set_stream(qipb_ctx *ctx, void *cuda_stream) {
    QIPB_REQUIRE(ctx, "null context");
    ctx->stream = (cudaStream_t)cuda_stream;
    return QIPB_OK;
}

extern "C" int qipb_sync(qipb_ctx *ctx) {
    QIPB_REQUIRE(ctx, "null context");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    QIPB_CUDA(cudaStreamSynchronize(ctx->stream));
    return QIPB_OK;
}

extern "C" unsigned long long qipb_launch_count(qipb_ctx *ctx) { return ctx ? ctx->launches : 0ull; }

extern "C" unsigned long long qipb_ring_launch_count(qipb_ctx *ctx) { return ctx ? ctx->ring_launches : 0ull; }

extern "C" unsigned long long qipb_ext_launch_count(qipb_ctx *ctx) { return ctx ? ctx->ext_launches : 0ull; }

extern "C" int qipb_dev_alloc(qipb_ctx *ctx, size_t bytes, void **out) {
    QIPB_REQUIRE(ctx && out, "null argument");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    QIPB_CUDA(cudaMalloc(out, bytes));
    return QIPB_OK;
}

extern "C" int qipb_dev_free(qipb_ctx *ctx, void *ptr) {
    QIPB_REQUIRE(ctx, "null context");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    QIPB_CUDA(cudaFree(ptr));
    return QIPB_OK;
}

extern "C" int qipb_memcpy_h2d(qipb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
    QIPB_REQUIRE(ctx && dst_dev && src_host, "null argument");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    QIPB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    QIPB_CUDA(cudaStreamSynchronize(ctx->stream));
    return QIPB_OK;
}

extern "C" int qipb_memcpy_d2h(qipb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes) {
    QIPB_REQUIRE(ctx && dst_host && src_dev, "null argument");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    QIPB_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    QIPB_CUDA(cudaStreamSynchronize(ctx->stream));
    return QIPB_OK;
}
