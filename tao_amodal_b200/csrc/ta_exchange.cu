// ta_exchange.cu — the cross-GPU step of the evaluation, inside the C ABI (include/ta_eval.h).
//
// Videos (and their images) shard across ranks; IoU and greedy matching need no communication.
// accumulate (tao_amodal/evaluation/tao_amodal/eval.py:498-518, lvis_amodal/eval.py:340-361),
// however, orders ALL detections of a category by score across all videos, so one record per
// detection travels to the rank that owns its category (contiguous category blocks, balanced by
// detection count).  Per evaluation that is
//     k_xchg_gather      records packed in send order (grouped by destination) by one kernel
//     ncclSend/ncclRecv  one grouped all-to-all of the packed bytes over NVLink / NVSwitch
//     k_xchg_scatter     (frame path) full rows of the few detections the general matcher handled
//     ncclAllReduce      sum of the non-ignored GT counts  int32 [n_cat][n_cfg]
// followed by ta_pr_accumulate on the owner's categories; the owners keep (and copy out) their
// own slices of precision / recall — nothing is gathered on one GPU.
//
// NCCL is bound at run time (dlopen of libnccl.so.2; an already loaded copy — e.g. the one
// torch.distributed brought in — is reused), so the library still loads on machines without it
// and single-GPU use never touches it.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "ta_internal.h"

namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.handle) return TA_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return ta_set_err(TA_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define TA_SYM(field, name)                                                              \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                           \
    if (!g_nccl.field) return ta_set_err(TA_ERR_NCCL, "libnccl lacks symbol %s", name)
    TA_SYM(GetUniqueId, "ncclGetUniqueId");
    TA_SYM(CommInitRank, "ncclCommInitRank");
    TA_SYM(CommDestroy, "ncclCommDestroy");
    TA_SYM(GroupStart, "ncclGroupStart");
    TA_SYM(GroupEnd, "ncclGroupEnd");
    TA_SYM(Send, "ncclSend");
    TA_SYM(Recv, "ncclRecv");
    TA_SYM(AllReduce, "ncclAllReduce");
    TA_SYM(GetErrorString, "ncclGetErrorString");
#undef TA_SYM
    g_nccl.handle = h;
    return TA_OK;
}
}  // namespace

#define TA_NCCL(call)                                                                        \
    do {                                                                                     \
        ncclResult_t r_ = (call);                                                            \
        if (r_ != ncclSuccess)                                                               \
            return ta_set_err(TA_ERR_NCCL, "NCCL error %s at " __FILE__ ":%lld",              \
                              g_nccl.GetErrorString(r_), (long long)__LINE__);               \
    } while (0)

struct ta_exchange {
    ta_ctx* ctx;
    int rank, world;
    ncclComm_t comm;
};

extern "C" int ta_exchange_unique_id(void* id, int32_t id_bytes) {
    if (!id || id_bytes < (int32_t)sizeof(ncclUniqueId))
        return ta_set_err(TA_ERR_INVALID, "ta_exchange_unique_id: buffer must hold %s%lld bytes", "",
                          (long long)sizeof(ncclUniqueId));
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId u;
    TA_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return TA_OK;
}

extern "C" int ta_exchange_create(ta_ctx* ctx, int32_t rank, int32_t world, const void* id,
                                  ta_exchange** out) {
    if (!ctx || !id || !out) return ta_set_err(TA_ERR_INVALID, "ta_exchange_create: NULL argument");
    if (world < 1 || rank < 0 || rank >= world)
        return ta_set_err(TA_ERR_INVALID, "ta_exchange_create: bad rank / world");
    int rc = nccl_load();
    if (rc) return rc;
    TA_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t comm;
    TA_NCCL(g_nccl.CommInitRank(&comm, world, u, rank));
    ta_exchange* x = new ta_exchange();
    x->ctx = ctx;
    x->rank = rank;
    x->world = world;
    x->comm = comm;
    *out = x;
    return TA_OK;
}

extern "C" int ta_exchange_destroy(ta_exchange* x) {
    if (!x) return TA_OK;
    if (g_nccl.handle && x->comm) g_nccl.CommDestroy(x->comm);
    delete x;
    return TA_OK;
}

// out[i] = src[index[i]] for records of `words` 32-bit words (send-side packing: `index` lists
// the local records grouped by destination rank; also the owner-side gather of sparse rows)
__global__ void __launch_bounds__(256)
k_xchg_gather(int64_t n, int words, const int32_t* __restrict__ index,
              const uint32_t* __restrict__ src, uint32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (words == 1) {
        if (i < n) out[i] = src[index[i]];
        return;
    }
    // one thread per output word: consecutive threads write consecutive words of a record
    const int64_t rec = i / words;
    if (rec >= n) return;
    const int w = (int)(i - rec * words);
    out[i] = src[(int64_t)index[rec] * words + w];
}

// dst[index[i]] = src[i]: received full rows land at their record's position in the dense row table
__global__ void __launch_bounds__(256)
k_xchg_scatter(int64_t n, int words, const int32_t* __restrict__ index,
               const uint32_t* __restrict__ src, uint32_t* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t rec = i / words;
    if (rec >= n) return;
    const int w = (int)(i - rec * words);
    dst[(int64_t)index[rec] * words + w] = src[i];
}

extern "C" int ta_exchange_gather(ta_ctx* ctx, void* stream, int64_t n, int32_t words,
                                  const int32_t* index, const uint32_t* src, uint32_t* out) {
    if (!ctx) return ta_set_err(TA_ERR_INVALID, "ta_exchange_gather: ctx is NULL");
    if (n < 0 || words < 1) return ta_set_err(TA_ERR_INVALID, "ta_exchange_gather: bad sizes");
    if (n == 0) return TA_OK;
    TA_CUDA(cudaSetDevice(ctx->device));
    ta_begin(ctx, (cudaStream_t)stream);
    const int64_t threads = n * words;
    k_xchg_gather<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, words, index, src, out);
    return ta_check_launch(ctx, "k_xchg_gather");
}

extern "C" int ta_exchange_scatter(ta_ctx* ctx, void* stream, int64_t n, int32_t words,
                                   const int32_t* index, const uint32_t* src, uint32_t* dst) {
    if (!ctx) return ta_set_err(TA_ERR_INVALID, "ta_exchange_scatter: ctx is NULL");
    if (n < 0 || words < 1) return ta_set_err(TA_ERR_INVALID, "ta_exchange_scatter: bad sizes");
    if (n == 0) return TA_OK;
    TA_CUDA(cudaSetDevice(ctx->device));
    ta_begin(ctx, (cudaStream_t)stream);
    const int64_t threads = n * words;
    k_xchg_scatter<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, words, index, src, dst);
    return ta_check_launch(ctx, "k_xchg_scatter");
}

// All-to-all of byte ranges: this rank sends send[send_off[r] .. send_off[r+1]) to rank r and
// receives recv[recv_off[r] .. recv_off[r+1]) from it (offsets in BYTES, host arrays of
// world + 1 entries).  One NCCL group; the part addressed to this rank itself is a device copy.
extern "C" int ta_exchange_alltoallv(ta_exchange* x, void* stream, const void* send,
                                     const int64_t* send_off, void* recv, const int64_t* recv_off) {
    if (!x || !send_off || !recv_off) return ta_set_err(TA_ERR_INVALID, "ta_exchange_alltoallv: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    TA_CUDA(cudaSetDevice(x->ctx->device));
    const char* s = static_cast<const char*>(send);
    char* r = static_cast<char*>(recv);
    const int me = x->rank;
    if (send_off[me + 1] - send_off[me] != recv_off[me + 1] - recv_off[me])
        return ta_set_err(TA_ERR_INVALID, "ta_exchange_alltoallv: self segment sizes differ");
    if (send_off[me + 1] > send_off[me])
        TA_CUDA(cudaMemcpyAsync(r + recv_off[me], s + send_off[me], (size_t)(send_off[me + 1] - send_off[me]),
                                cudaMemcpyDeviceToDevice, st));
    if (x->world == 1) return TA_OK;
    TA_NCCL(g_nccl.GroupStart());
    for (int p = 0; p < x->world; ++p) {
        if (p == me) continue;
        const int64_t ns = send_off[p + 1] - send_off[p], nr = recv_off[p + 1] - recv_off[p];
        if (ns > 0) TA_NCCL(g_nccl.Send(s + send_off[p], (size_t)ns, ncclInt8, p, x->comm, st));
        if (nr > 0) TA_NCCL(g_nccl.Recv(r + recv_off[p], (size_t)nr, ncclInt8, p, x->comm, st));
    }
    TA_NCCL(g_nccl.GroupEnd());
    return TA_OK;
}

// Several exchange calls issued between group_begin / group_end become ONE NCCL launch.
extern "C" int ta_exchange_group_begin(ta_exchange* x) {
    if (!x) return ta_set_err(TA_ERR_INVALID, "ta_exchange_group_begin: NULL argument");
    if (x->world == 1) return TA_OK;
    TA_NCCL(g_nccl.GroupStart());
    return TA_OK;
}
extern "C" int ta_exchange_group_end(ta_exchange* x) {
    if (!x) return ta_set_err(TA_ERR_INVALID, "ta_exchange_group_end: NULL argument");
    if (x->world == 1) return TA_OK;
    TA_NCCL(g_nccl.GroupEnd());
    return TA_OK;
}

// In-place sum over ranks.  dtype: 0 = int32, 1 = int64, 2 = float64.
extern "C" int ta_exchange_allreduce_sum(ta_exchange* x, void* stream, void* buf, int64_t count,
                                         int32_t dtype) {
    if (!x || !buf) return ta_set_err(TA_ERR_INVALID, "ta_exchange_allreduce_sum: NULL argument");
    if (count < 0 || dtype < 0 || dtype > 2) return ta_set_err(TA_ERR_INVALID, "ta_exchange_allreduce_sum: bad arguments");
    if (count == 0 || x->world == 1) return TA_OK;
    TA_CUDA(cudaSetDevice(x->ctx->device));
    const ncclDataType_t dt = dtype == 0 ? ncclInt32 : (dtype == 1 ? ncclInt64 : ncclFloat64);
    TA_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, dt, ncclSum, x->comm, (cudaStream_t)stream));
    return TA_OK;
}

extern "C" int ta_exchange_rank(const ta_exchange* x) { return x ? x->rank : -1; }
extern "C" int ta_exchange_world(const ta_exchange* x) { return x ? x->world : 0; }
