// ta_internal.h — context, error plumbing and launch helpers shared by the .cu files of
// libta_eval.so.  Not part of the public C ABI (include/ta_eval.h).
#ifndef TA_INTERNAL_H
#define TA_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ta_eval.h"

struct ta_ctx {
    int device;
    int sm_count;
    int smem_optin;        // max dynamic shared memory per block (opt-in)
    int64_t launches;      // kernels launched through this context
    int* d_flags;          // [0]: "intersection > union" counter (eval.py:95)
    cudaStream_t own_stream;
    cudaStream_t copy_stream;   // uploads of ta_eval_plans_host (created on first use)
    // stream-ordered scratch of ta_pr_accumulate, grown on demand
    void* ws;
    size_t ws_bytes;
    void* ws2;             // second scratch slot (frame path: detection -> group map, lists)
    size_t ws2_bytes;
    // pinned staging for ta_eval_plan_host results
    void* h_stage;
    size_t h_stage_bytes;
    // optional per-kernel timing (ta_ctx_timing): one event after every launch, one marker at
    // every API entry; a kernel's time is the interval since the previous event on the stream
    cudaStream_t cur_stream;
    int timing;
    int n_ev;
    cudaEvent_t* ev;
    const char** ev_name;      // NULL = entry marker
    // diagnostics: device counter of the groups the most recent flat-kernel call handed to the
    // general matcher (ta_ctx_debug_list_count)
    const int32_t* last_list_count;
};

#define TA_MAX_EVENTS 16384
// Call at every API entry that launches kernels: remembers the stream for ta_check_launch.
void ta_begin(ta_ctx* ctx, cudaStream_t st);

char* ta_err_buf();
int ta_set_err(int code, const char* fmt, const char* a = "", long long b = 0);

#define TA_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return ta_set_err(TA_ERR_CUDA, "CUDA error %s at " __FILE__ ":%lld",              \
                              cudaGetErrorString(e_), (long long)__LINE__);                  \
    } while (0)

// cudaGetLastError() after a launch; counts the launch on success.
int ta_check_launch(ta_ctx* ctx, const char* what);
// Returns a device scratch buffer of at least `bytes` (stream-ordered; contents undefined).
int ta_workspace(ta_ctx* ctx, cudaStream_t st, size_t bytes, void** out, int slot = 0);

#endif  // TA_INTERNAL_H
