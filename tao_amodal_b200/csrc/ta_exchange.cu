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
// own slices of precision / recall — nothing is gathered on one GPU.  The default route does the
// all-to-all and the all-reduce with ONE kernel of this library over NVLink peer memory instead
// (peer windows, second half of this file); the NCCL calls stay as the alternative route and
// carry the setup traffic.
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
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
    TA_SYM(AllGather, "ncclAllGather");
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

// ---- peer windows: the same exchange with this library's own kernels over NVLink ----------------
// A window is one cudaMalloc'ed buffer per rank whose CUDA IPC handle every other rank has opened
// (one process per GPU; NVLink / NVSwitch peer access).  An exchange is then
//     acquire    wait until every peer has finished reading the window's previous contents
//     (caller)   write the records into the window (ta_peer_window_put, or kernels writing there)
//     exchange   ONE kernel: publish the window (epoch flag), wait for each peer's flag, pull this
//                owner's slices out of the peers' windows with 16-byte loads over NVLink, sum the
//                GT counts of all ranks, and tell the peers the window has been read
// i.e. no NCCL launch, no proxy thread and no send-side packing on the step path: the transfer is
// the owner's loads.  Flags live in the first TA_PW_HDR bytes of the window: [0] ready epoch
// (written by the owner of the window), [16 + r] the epoch rank r has finished pulling.
#define TA_PW_HDR 1024
#define TA_PW_SPIN_LIMIT (1ll << 37)     // clock ticks (about a minute): a dead peer must not hang the GPU for good

struct ta_peer_window {
    ta_exchange* x;
    int64_t bytes;
    char* base;                      // this rank's window
    char* peer[64];                  // every rank's window as mapped here (peer[rank] == base)
    int64_t peer_bytes[64];          // and its size
    uint32_t epoch;
    int n_seg, n_words, n_sum;
    int64_t total_vec;
    void* d_seg;                     // PwSeg[n_seg]
    void* d_words;                   // PwWord[n_words]
    ta_peer_sum* d_sum;
    unsigned int* d_counter;         // blocks that finished pulling (reset by the last one)
    int* d_err;                      // spin limit hit
    char** d_peer;
};

__device__ __forceinline__ uint32_t pw_load_flag(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void pw_store_flag(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ bool pw_wait(const uint32_t* p, uint32_t epoch, int* err) {
    const long long t0 = clock64();
    while ((int32_t)(pw_load_flag(p) - epoch) < 0) {
        if (clock64() - t0 > TA_PW_SPIN_LIMIT) { atomicExch(err, 1); return false; }
        __nanosleep(200);
    }
    return true;
}

// wait until every peer has pulled epoch `prev` from this rank's window
__global__ void k_peer_acquire(const uint32_t* __restrict__ hdr, int world, int me, uint32_t prev, int* err) {
    const int r = threadIdx.x;
    if (r < world && r != me) pw_wait(hdr + 16 + r, prev, err);
}

// What one exchange pulls, flattened on the host (ta_peer_window_set_plan): every copy is cut into
// the words before the first 16-byte boundary of its SOURCE, a run of 16-byte vectors, and the
// words after it.  All vectors of all copies form ONE index space (vec_begin = prefix sums), so a
// thread keeps PW_UNROLL loads from several peers in flight at once instead of finishing one
// peer's slice (a few NVLink round trips each) before touching the next.
struct PwSeg { const uint4* s128; uint32_t* d32; int64_t vec_begin; };
struct PwWord { const uint32_t* s; uint32_t* d; };
#define PW_MAX_SEG 160
#define PW_UNROLL 8

__global__ void __launch_bounds__(256)
k_peer_exchange(char* const* __restrict__ peer, int world, int me, uint32_t epoch,
                const PwSeg* __restrict__ segs, int n_seg, int64_t total_vec,
                const PwWord* __restrict__ words, int n_words,
                const ta_peer_sum* __restrict__ sums, int n_sum,
                unsigned int* counter, int* err) {
    __shared__ int64_t begin_s[PW_MAX_SEG + 1];
    __shared__ const uint4* src_s[PW_MAX_SEG];
    __shared__ uint32_t* dst_s[PW_MAX_SEG];
    __shared__ int ok_s;
    // publish: everything written into this window earlier on the stream is visible before the flag
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        __threadfence_system();
        pw_store_flag(reinterpret_cast<uint32_t*>(peer[me]), epoch);
    }
    for (int i = threadIdx.x; i < n_seg; i += blockDim.x) {
        const PwSeg sg = segs[i];
        begin_s[i] = sg.vec_begin;
        src_s[i] = sg.s128;
        dst_s[i] = sg.d32;
    }
    if (threadIdx.x == 0) {
        begin_s[n_seg] = total_vec;
        ok_s = 1;
    }
    __syncthreads();
    // every peer has published this epoch (one poller per peer and block)
    if (threadIdx.x < world && threadIdx.x != me)
        if (!pw_wait(reinterpret_cast<const uint32_t*>(peer[threadIdx.x]), epoch, err)) ok_s = 0;
    __syncthreads();
    const bool ok = ok_s != 0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    if (ok) {
        for (int64_t i = tid; i < n_words; i += nthreads) *words[i].d = __ldcv(words[i].s);
        for (int64_t base = tid; base < total_vec; base += PW_UNROLL * nthreads) {
            uint4 v[PW_UNROLL];
            uint32_t* o[PW_UNROLL];
#pragma unroll
            for (int u = 0; u < PW_UNROLL; ++u) {
                const int64_t idx = base + u * nthreads;
                o[u] = nullptr;
                if (idx < total_vec) {
                    int lo = 0, hi = n_seg;              // begin_s[lo] <= idx < begin_s[hi]
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (begin_s[mid] <= idx) lo = mid; else hi = mid;
                    }
                    const int64_t k = idx - begin_s[lo];
                    v[u] = __ldcv(src_s[lo] + k);
                    o[u] = dst_s[lo] + 4 * k;
                }
            }
#pragma unroll
            for (int u = 0; u < PW_UNROLL; ++u) {
                if (!o[u]) continue;
                if ((((uintptr_t)o[u]) & 15) == 0) *reinterpret_cast<uint4*>(o[u]) = v[u];
                else { o[u][0] = v[u].x; o[u][1] = v[u].y; o[u][2] = v[u].z; o[u][3] = v[u].w; }
            }
        }
        // sums over all ranks (GT counts): every rank computes the same totals
        for (int k = 0; k < n_sum; ++k) {
            const ta_peer_sum q = sums[k];
            for (int64_t i = tid; i < q.count; i += nthreads) {
                int32_t acc = 0;
                for (int p = 0; p < world; ++p)
                    acc += (int32_t)__ldcv(reinterpret_cast<const uint32_t*>(peer[p] + q.off) + i);
                q.dst[i] = acc;
            }
        }
    }
    // the last block to finish tells every peer that this rank has read its window
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(counter, 1u) == gridDim.x - 1) {
            *counter = 0;
            __threadfence_system();
            for (int p = 0; p < world; ++p)
                if (p != me) pw_store_flag(reinterpret_cast<uint32_t*>(peer[p]) + 16 + me, epoch);
        }
    }
}

extern "C" int ta_peer_window_create(ta_exchange* x, int64_t bytes, ta_peer_window** out) {
    if (!x || !out || bytes < 0) return ta_set_err(TA_ERR_INVALID, "ta_peer_window_create: bad arguments");
    if (x->world > 64) return ta_set_err(TA_ERR_TOO_LARGE, "ta_peer_window_create: more than 64 ranks");
    TA_CUDA(cudaSetDevice(x->ctx->device));
    ta_peer_window* w = new ta_peer_window();
    memset(w, 0, sizeof(*w));
    w->x = x;
    w->bytes = TA_PW_HDR + ((bytes + 255) & ~(int64_t)255);
    TA_CUDA(cudaMalloc(&w->base, (size_t)w->bytes));
    TA_CUDA(cudaMemset(w->base, 0, (size_t)w->bytes));
    TA_CUDA(cudaMalloc(&w->d_counter, 2 * sizeof(int)));
    TA_CUDA(cudaMemset(w->d_counter, 0, 2 * sizeof(int)));
    w->d_err = reinterpret_cast<int*>(w->d_counter) + 1;
    TA_CUDA(cudaDeviceSynchronize());
    w->peer[x->rank] = w->base;
    w->peer_bytes[x->rank] = w->bytes;
    if (x->world > 1) {
        // every rank's IPC handle and window size through the communicator that already exists
        struct Card { cudaIpcMemHandle_t h; int64_t bytes; };
        Card mine;
        TA_CUDA(cudaIpcGetMemHandle(&mine.h, w->base));
        mine.bytes = w->bytes;
        char* d_all = nullptr;
        const size_t hb = sizeof(Card);
        TA_CUDA(cudaMalloc(&d_all, hb * (size_t)x->world));
        TA_CUDA(cudaMemcpy(d_all + hb * (size_t)x->rank, &mine, hb, cudaMemcpyHostToDevice));
        cudaStream_t st = x->ctx->own_stream;
        TA_NCCL(g_nccl.AllGather(d_all + hb * (size_t)x->rank, d_all, hb, ncclInt8, x->comm, st));
        TA_CUDA(cudaStreamSynchronize(st));
        Card* all = new Card[x->world];
        TA_CUDA(cudaMemcpy(all, d_all, hb * (size_t)x->world, cudaMemcpyDeviceToHost));
        cudaFree(d_all);
        for (int p = 0; p < x->world; ++p) {
            if (p == x->rank) continue;
            void* q = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&q, all[p].h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                delete[] all;
                return ta_set_err(TA_ERR_CUDA, "ta_peer_window_create: cannot map a peer's window (%s)",
                                  cudaGetErrorString(e));
            }
            w->peer[p] = static_cast<char*>(q);
            w->peer_bytes[p] = all[p].bytes;
        }
        delete[] all;
    }
    TA_CUDA(cudaMalloc(&w->d_peer, sizeof(char*) * 64));
    TA_CUDA(cudaMemcpy(w->d_peer, w->peer, sizeof(char*) * 64, cudaMemcpyHostToDevice));
    w->epoch = 0;
    *out = w;
    return TA_OK;
}

extern "C" void* ta_peer_window_ptr(ta_peer_window* w) { return w ? w->base + TA_PW_HDR : nullptr; }

extern "C" int ta_peer_window_destroy(ta_peer_window* w) {
    if (!w) return TA_OK;
    cudaSetDevice(w->x->ctx->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < w->x->world; ++p)
        if (p != w->x->rank && w->peer[p]) cudaIpcCloseMemHandle(w->peer[p]);
    if (w->d_seg) cudaFree(w->d_seg);
    if (w->d_words) cudaFree(w->d_words);
    if (w->d_sum) cudaFree(w->d_sum);
    if (w->d_peer) cudaFree(w->d_peer);
    if (w->d_counter) cudaFree(w->d_counter);
    if (w->base) cudaFree(w->base);
    delete w;
    return TA_OK;
}

extern "C" int ta_peer_window_set_plan(ta_peer_window* w, int32_t n_copy, const ta_peer_copy* copies,
                                       int32_t n_sum, const ta_peer_sum* sums) {
    if (!w || n_copy < 0 || n_sum < 0 || (n_copy && !copies) || (n_sum && !sums))
        return ta_set_err(TA_ERR_INVALID, "ta_peer_window_set_plan: bad arguments");
    if (n_copy > PW_MAX_SEG) return ta_set_err(TA_ERR_TOO_LARGE, "ta_peer_window_set_plan: more than %s%lld copies", "", PW_MAX_SEG);
    TA_CUDA(cudaSetDevice(w->x->ctx->device));
    const int world = w->x->world, me = w->x->rank;
    const int64_t room = w->bytes - TA_PW_HDR;
    for (int i = 0; i < n_copy; ++i) {
        const ta_peer_copy& c = copies[i];
        if (c.peer < 0 || c.peer >= world || c.src_off < 0 || c.bytes < 0 || (c.src_off & 3) ||
            (c.bytes & 3) || c.src_off + c.bytes > w->peer_bytes[c.peer] - TA_PW_HDR ||
            (c.bytes && (!c.dst || ((uintptr_t)c.dst & 3))))
            return ta_set_err(TA_ERR_INVALID, "ta_peer_window_set_plan: copy %s%lld is out of the window or misaligned", "", i);
    }
    for (int i = 0; i < n_sum; ++i)
        if (sums[i].off < 0 || (sums[i].off & 3) || sums[i].count < 0 || sums[i].off + 4 * sums[i].count > room ||
            (sums[i].count && !sums[i].dst))
            return ta_set_err(TA_ERR_INVALID, "ta_peer_window_set_plan: sum %s%lld is out of the window", "", i);
    // segments in the order (peer - me - 1) mod world: at any moment the ranks read from different
    // peers, so the links fill evenly
    PwSeg* hs = new PwSeg[n_copy + 1];
    PwWord* hw = new PwWord[6 * (size_t)n_copy + 1];
    int n_seg = 0, n_words = 0;
    int64_t total = 0;
    for (int step = 0; step < world; ++step) {
        const int p = (me + 1 + step) % world;
        for (int i = 0; i < n_copy; ++i) {
            const ta_peer_copy& c = copies[i];
            if (c.peer != p || c.bytes == 0) continue;
            const char* src = w->peer[p] + TA_PW_HDR + c.src_off;
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
            uint32_t* d32 = static_cast<uint32_t*>(c.dst);
            const int64_t nw = c.bytes >> 2;
            int64_t head = (int64_t)((16 - ((uintptr_t)src & 15)) & 15) >> 2;
            if (head > nw) head = nw;
            const int64_t n_vec = (nw - head) >> 2;
            for (int64_t k = 0; k < head; ++k) hw[n_words++] = {s32 + k, d32 + k};
            for (int64_t k = head + 4 * n_vec; k < nw; ++k) hw[n_words++] = {s32 + k, d32 + k};
            if (n_vec) {
                hs[n_seg++] = {reinterpret_cast<const uint4*>(s32 + head), d32 + head, total};
                total += n_vec;
            }
        }
    }
    ta_peer_sum* hq = n_sum ? new ta_peer_sum[n_sum] : nullptr;
    for (int i = 0; i < n_sum; ++i) {
        hq[i] = sums[i];
        hq[i].off += TA_PW_HDR;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (w->d_seg) cudaFree(w->d_seg);
    if (w->d_words) cudaFree(w->d_words);
    if (w->d_sum) cudaFree(w->d_sum);
    w->d_seg = w->d_words = nullptr;
    w->d_sum = nullptr;
    if (e == cudaSuccess && n_seg) e = cudaMalloc(&w->d_seg, sizeof(PwSeg) * n_seg);
    if (e == cudaSuccess && n_seg) e = cudaMemcpy(w->d_seg, hs, sizeof(PwSeg) * n_seg, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_words) e = cudaMalloc(&w->d_words, sizeof(PwWord) * n_words);
    if (e == cudaSuccess && n_words) e = cudaMemcpy(w->d_words, hw, sizeof(PwWord) * n_words, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_sum) e = cudaMalloc(&w->d_sum, sizeof(ta_peer_sum) * n_sum);
    if (e == cudaSuccess && n_sum) e = cudaMemcpy(w->d_sum, hq, sizeof(ta_peer_sum) * n_sum, cudaMemcpyHostToDevice);
    delete[] hs;
    delete[] hw;
    delete[] hq;
    TA_CUDA(e);
    w->n_seg = n_seg;
    w->n_words = n_words;
    w->total_vec = total;
    w->n_sum = n_sum;
    return TA_OK;
}

extern "C" int ta_peer_window_acquire(ta_peer_window* w, void* stream) {
    if (!w) return ta_set_err(TA_ERR_INVALID, "ta_peer_window_acquire: NULL argument");
    ta_ctx* ctx = w->x->ctx;
    TA_CUDA(cudaSetDevice(ctx->device));
    if (w->x->world == 1 || w->epoch == 0) return TA_OK;
    ta_begin(ctx, (cudaStream_t)stream);
    k_peer_acquire<<<1, 64, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t*>(w->base), w->x->world,
                                                      w->x->rank, w->epoch, w->d_err);
    return ta_check_launch(ctx, "k_peer_acquire");
}

extern "C" int ta_peer_window_put(ta_peer_window* w, void* stream, int64_t off, const void* src, int64_t bytes) {
    if (!w || off < 0 || bytes < 0 || off + bytes > w->bytes - TA_PW_HDR || (bytes && !src))
        return ta_set_err(TA_ERR_INVALID, "ta_peer_window_put: out of the window");
    TA_CUDA(cudaSetDevice(w->x->ctx->device));
    if (bytes)
        TA_CUDA(cudaMemcpyAsync(w->base + TA_PW_HDR + off, src, (size_t)bytes, cudaMemcpyDeviceToDevice,
                                (cudaStream_t)stream));
    return TA_OK;
}

extern "C" int ta_peer_window_exchange(ta_peer_window* w, void* stream) {
    if (!w) return ta_set_err(TA_ERR_INVALID, "ta_peer_window_exchange: NULL argument");
    ta_ctx* ctx = w->x->ctx;
    TA_CUDA(cudaSetDevice(ctx->device));
    ta_begin(ctx, (cudaStream_t)stream);
    w->epoch += 1;
    // every block must be resident (blocks wait for remote flags): two per SM
    const int blocks = ctx->sm_count * 2;
    k_peer_exchange<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        w->d_peer, w->x->world, w->x->rank, w->epoch, static_cast<const PwSeg*>(w->d_seg), w->n_seg,
        w->total_vec, static_cast<const PwWord*>(w->d_words), w->n_words, w->d_sum, w->n_sum,
        w->d_counter, w->d_err);
    return ta_check_launch(ctx, "k_peer_exchange");
}

// 1 when a wait on a peer's flag ran into the spin limit since the last call (a peer died or never
// entered the exchange: the received data is then incomplete); waits for `stream`.
extern "C" int ta_peer_window_check(ta_peer_window* w, void* stream, int32_t* timed_out) {
    if (!w || !timed_out) return ta_set_err(TA_ERR_INVALID, "ta_peer_window_check: NULL argument");
    TA_CUDA(cudaSetDevice(w->x->ctx->device));
    int v = 0;
    TA_CUDA(cudaMemcpyAsync(&v, w->d_err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    TA_CUDA(cudaMemsetAsync(w->d_err, 0, sizeof(int), (cudaStream_t)stream));
    TA_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *timed_out = v;
    return TA_OK;
}
