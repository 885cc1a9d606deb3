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
    if (ctx->timing && ctx->n_ev < TA_MAX_EVENTS) {
        cudaEventRecord(ctx->ev[ctx->n_ev], ctx->cur_stream);
        ctx->ev_name[ctx->n_ev++] = what;
    }
    return TA_OK;
}

void ta_begin(ta_ctx* ctx, cudaStream_t st) {
    ctx->cur_stream = st;
    if (ctx->timing && ctx->n_ev < TA_MAX_EVENTS) {
        cudaEventRecord(ctx->ev[ctx->n_ev], st);
        ctx->ev_name[ctx->n_ev++] = nullptr;
    }
}

extern "C" int ta_ctx_timing(ta_ctx* c, int enable) {
    if (!c) return ta_set_err(TA_ERR_INVALID, "ta_ctx_timing: ctx is NULL");
    TA_CUDA(cudaSetDevice(c->device));
    if (enable && !c->ev) {
        c->ev = new cudaEvent_t[TA_MAX_EVENTS];
        c->ev_name = new const char*[TA_MAX_EVENTS];
        for (int i = 0; i < TA_MAX_EVENTS; ++i) TA_CUDA(cudaEventCreate(&c->ev[i]));
    }
    c->timing = enable ? 1 : 0;
    c->n_ev = 0;
    return TA_OK;
}

extern "C" int ta_ctx_timing_read(ta_ctx* c, char* names, int names_cap, double* total_ms,
                                  int* launches, int cap) {
    if (!c || !names || !total_ms || !launches)
        return ta_set_err(TA_ERR_INVALID, "ta_ctx_timing_read: NULL argument");
    TA_CUDA(cudaSetDevice(c->device));
    int n = 0, used = 0;
    const char* seen[256];
    if (names_cap > 0) names[0] = 0;
    for (int i = 1; i < c->n_ev; ++i) {
        if (!c->ev_name[i]) continue;
        TA_CUDA(cudaEventSynchronize(c->ev[i]));
        float ms = 0.f;
        TA_CUDA(cudaEventElapsedTime(&ms, c->ev[i - 1], c->ev[i]));
        int k = 0;
        for (; k < n; ++k) if (seen[k] == c->ev_name[i] || !strcmp(seen[k], c->ev_name[i])) break;
        if (k == n) {
            if (n >= cap || n >= 256) continue;
            const int len = (int)strlen(c->ev_name[i]);
            if (used + len + 2 > names_cap) continue;
            memcpy(names + used, c->ev_name[i], len);
            names[used + len] = '\n';
            names[used + len + 1] = 0;
            used += len + 1;
            seen[n] = c->ev_name[i];
            total_ms[n] = 0.0;
            launches[n] = 0;
            ++n;
        }
        total_ms[k] += ms;
        launches[k] += 1;
    }
    c->n_ev = 0;
    return n;
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
    if (c->ev) {
        for (int i = 0; i < TA_MAX_EVENTS; ++i) cudaEventDestroy(c->ev[i]);
        delete[] c->ev;
        delete[] c->ev_name;
    }
    if (c->h_stage) cudaFreeHost(c->h_stage);
    cudaFree(c->d_flags);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->own_stream);
    delete c;
    return TA_OK;
}

// Number of track pairs whose intersection exceeded their union since the last call (the
// reference raises AssertionError there, eval.py:95); waits for `stream` and resets the counter.
extern "C" int ta_ctx_take_assert_count(ta_ctx* c, void* stream, int32_t* count) {
    if (!c || !count) return ta_set_err(TA_ERR_INVALID, "ta_ctx_take_assert_count: NULL argument");
    TA_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    int v = 0;
    TA_CUDA(cudaMemcpyAsync(&v, c->d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    TA_CUDA(cudaMemsetAsync(c->d_flags, 0, sizeof(int), st));
    TA_CUDA(cudaStreamSynchronize(st));
    *count = v;
    return TA_OK;
}

// Diagnostics: how many groups the most recent ta_match_greedy / ta_frame_eval evaluation call
// on this context handed to the general matcher (waits for `stream`).
extern "C" int ta_ctx_debug_list_count(ta_ctx* c, void* stream, int32_t* count) {
    if (!c || !count) return ta_set_err(TA_ERR_INVALID, "ta_ctx_debug_list_count: NULL argument");
    *count = -1;
    if (!c->last_list_count) return TA_OK;
    TA_CUDA(cudaSetDevice(c->device));
    TA_CUDA(cudaMemcpyAsync(count, c->last_list_count, sizeof(int32_t), cudaMemcpyDeviceToHost,
                            (cudaStream_t)stream));
    TA_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return TA_OK;
}

// cudaMemsetAsync(p, 0, bytes) on `stream`: callers zero num_gt with it (the counts are
// accumulated) instead of launching a framework fill kernel.
extern "C" int ta_zero(ta_ctx* c, void* stream, void* p, int64_t bytes) {
    if (!c) return ta_set_err(TA_ERR_INVALID, "ta_zero: ctx is NULL");
    if (bytes < 0 || (bytes > 0 && !p)) return ta_set_err(TA_ERR_INVALID, "ta_zero: bad arguments");
    if (bytes == 0) return TA_OK;
    TA_CUDA(cudaSetDevice(c->device));
    TA_CUDA(cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream));
    return TA_OK;
}

extern "C" int ta_ctx_sm_count(const ta_ctx* c) { return c ? c->sm_count : 0; }
extern "C" int64_t ta_ctx_launch_count(const ta_ctx* c) { return c ? c->launches : 0; }
