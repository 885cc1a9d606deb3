// ta_ctx.cu — context lifetime, thread-local error string, scratch workspace.
#include <string.h>
#include "ta_internal.h"

static thread_local char g_err[512] = "";

char* ta_err_buf() { return g_err; }

int ta_set_err(int code, const char* fmt, const char* a, long long b) {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

int ta_check_launch(ta_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s launch failed: %s", what, cudaGetErrorString(e));
        return TA_ERR_CUDA;
    }
    ctx->launches++;
    return TA_OK;
}

int ta_workspace(ta_ctx* ctx, cudaStream_t st, size_t bytes, void** out, int slot) {
    void*& buf = slot ? ctx->ws2 : ctx->ws;
    size_t& have = slot ? ctx->ws2_bytes : ctx->ws_bytes;
    if (bytes > have) {
        // rare (first call / a larger problem): plain cudaFree + cudaMalloc, which also order
        // themselves after any kernel still using the old buffer
        (void)st;
        if (buf) TA_CUDA(cudaFree(buf));
        buf = nullptr;
        have = 0;
        size_t want = bytes + bytes / 4 + 4096;
        TA_CUDA(cudaMalloc(&buf, want));
        have = want;
    }
    *out = buf;
    return TA_OK;
}

extern "C" int ta_abi_version(void) { return TA_ABI_VERSION; }
extern "C" const char* ta_last_error(void) { return g_err; }

extern "C" int ta_ctx_create(int device, ta_ctx** out) {
    if (!out) return ta_set_err(TA_ERR_INVALID, "ta_ctx_create: out is NULL");
    int n = 0;
    TA_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n)
        return ta_set_err(TA_ERR_INVALID, "ta_ctx_create: bad device %s%lld", "", device);
    TA_CUDA(cudaSetDevice(device));
    ta_ctx* c = new ta_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device;
    TA_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    TA_CUDA(cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    TA_CUDA(cudaMalloc(&c->d_flags, 4 * sizeof(int)));
    TA_CUDA(cudaMemset(c->d_flags, 0, 4 * sizeof(int)));
    TA_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    // keep freed stream-ordered allocations in the pool: ta_eval_plan_host allocates a few GB
    // per call and would otherwise hand them back to the driver at every synchronisation
    cudaMemPool_t pool;
    TA_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = UINT64_MAX;
    TA_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    *out = c;
    return TA_OK;
}

extern "C" int ta_ctx_destroy(ta_ctx* c) {
    if (!c) return TA_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->own_stream);
    if (c->ws) cudaFree(c->ws);
    if (c->ws2) cudaFree(c->ws2);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    cudaFree(c->d_flags);
    cudaStreamDestroy(c->own_stream);
    delete c;
    return TA_OK;
}

extern "C" int ta_ctx_sm_count(const ta_ctx* c) { return c ? c->sm_count : 0; }
extern "C" int64_t ta_ctx_launch_count(const ta_ctx* c) { return c ? c->launches : 0; }
