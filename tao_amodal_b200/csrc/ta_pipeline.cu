// ta_pipeline.cu — ta_eval_plans_host / ta_eval_plan_host: H2D -> IoU -> match -> accumulate ->
// D2H.  Every plan runs on its own context's stream; ALL host-to-device copies go through one
// upload stream in a fixed order (the box pool, then plan by plan: what the matcher needs, then
// what only accumulate needs), and each plan's stream waits for the events of its own inputs.
// The first plan's kernels and its download therefore overlap the upload of the next plan, and
// the IoU / matching kernels overlap the upload of the accumulate order.  Device memory comes
// from the stream-ordered pool (cudaMallocAsync; the context sets the pool's release threshold
// so repeated calls reuse the same pages).
#include <initializer_list>
#include <stdlib.h>

#include "ta_internal.h"

// uint16 -> int32 (frame slots with TA_PLAN_SLOT_U16, category indices with TA_PLAN_GRP_U16)
__global__ void k_widen_slots(const uint16_t* __restrict__ src, int32_t* __restrict__ dst, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (int32_t)src[i];
}

// lossless fp32 -> fp64 widening of box coordinates shipped as float (TA_PLAN_BOX_F32)
__global__ void k_widen_boxes(const float4* __restrict__ src, double* __restrict__ dst, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = src[i];
    double2* o = reinterpret_cast<double2*>(dst + 4 * i);
    o[0] = make_double2((double)v.x, (double)v.y);
    o[1] = make_double2((double)v.z, (double)v.w);
}

// dst[i] = pool[idx[i]]: the detection boxes of a plan that shares another plan's box upload
__global__ void k_gather_boxes(const double2* __restrict__ pool, const int32_t* __restrict__ idx,
                               double2* __restrict__ dst, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t s = idx[i];
    dst[2 * i] = pool[2 * s];
    dst[2 * i + 1] = pool[2 * s + 1];
}

// Exclusive prefix sum of uint16 counts -> int64 offsets [n+1] (TA_PLAN_GRP_U16): block sums,
// a one-block scan of the sums, then the per-block pass that writes the offsets.
constexpr int SC_T = 256, SC_E = 8, SC_B = SC_T * SC_E;

__device__ __forceinline__ uint32_t sc_block_excl(uint32_t v, uint32_t* total) {
    // exclusive scan of one value per thread over a 256-thread block
    __shared__ uint32_t wsum[SC_T / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < SC_T / 32; ++k) {
        const uint32_t s = wsum[k];
        if (k < w) base += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(SC_T) k_cnt_sums(const uint16_t* __restrict__ c, int64_t n,
                                                   int64_t* __restrict__ part) {
    const int64_t i0 = (int64_t)blockIdx.x * SC_B + (int64_t)threadIdx.x * SC_E;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SC_E; ++k)
        if (i0 + k < n) s += c[i0 + k];
    uint32_t tot;
    sc_block_excl(s, &tot);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_cnt_scan_part(int64_t* __restrict__ part, int nb) {
    // in place: part[b] <- sum of part[0..b)
    __shared__ int64_t sh[1024];
    const int per = (nb + 1023) / 1024;
    const int b0 = threadIdx.x * per, b1 = min(nb, b0 + per);
    int64_t s = 0;
    for (int b = b0; b < b1; ++b) s += part[b];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const int64_t o = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
        __syncthreads();
        sh[threadIdx.x] += o;
        __syncthreads();
    }
    int64_t run = sh[threadIdx.x] - s;
    for (int b = b0; b < b1; ++b) {
        const int64_t v = part[b];
        part[b] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(SC_T) k_cnt_offsets(const uint16_t* __restrict__ c, int64_t n,
                                                      const int64_t* __restrict__ part,
                                                      int64_t* __restrict__ off) {
    const int64_t i0 = (int64_t)blockIdx.x * SC_B + (int64_t)threadIdx.x * SC_E;
    uint32_t v[SC_E], s = 0;
#pragma unroll
    for (int k = 0; k < SC_E; ++k) {
        v[k] = (i0 + k < n) ? c[i0 + k] : 0u;
        s += v[k];
    }
    uint32_t tot;
    int64_t run = part[blockIdx.x] + sc_block_excl(s, &tot);
    if (blockIdx.x == 0 && threadIdx.x == 0) off[0] = 0;
#pragma unroll
    for (int k = 0; k < SC_E; ++k) {
        run += v[k];
        if (i0 + k < n) off[i0 + k + 1] = run;
    }
}

// Lossless transport of box coordinates: float [n,4] -> double [n,4] on the device (the fp64
// arithmetic then starts from the identical values; TA_PLAN_BOX_F32 of ta_eval_plan_host does
// the same internally).  For callers that keep plans resident and refresh them from host memory.
extern "C" int ta_widen_boxes(ta_ctx* ctx, void* stream, int64_t n, const float* src, double* dst) {
    if (!ctx) return ta_set_err(TA_ERR_INVALID, "ta_widen_boxes: ctx is NULL");
    if (n < 0) return ta_set_err(TA_ERR_INVALID, "ta_widen_boxes: negative size");
    if (n == 0) return TA_OK;
    TA_CUDA(cudaSetDevice(ctx->device));
    ta_begin(ctx, (cudaStream_t)stream);
    k_widen_boxes<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(src), dst, n);
    return ta_check_launch(ctx, "k_widen_boxes");
}

// The device-side halves of the compact transport forms, for callers that keep plans resident
// and refresh them from host memory (ta_eval_plans_host does the same internally).
extern "C" int ta_widen_u16(ta_ctx* ctx, void* stream, int64_t n, const uint16_t* src, int32_t* dst) {
    if (!ctx) return ta_set_err(TA_ERR_INVALID, "ta_widen_u16: ctx is NULL");
    if (n < 0) return ta_set_err(TA_ERR_INVALID, "ta_widen_u16: negative size");
    if (n == 0) return TA_OK;
    TA_CUDA(cudaSetDevice(ctx->device));
    ta_begin(ctx, (cudaStream_t)stream);
    k_widen_slots<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n);
    return ta_check_launch(ctx, "k_widen_slots");
}

extern "C" int ta_gather_boxes(ta_ctx* ctx, void* stream, int64_t n, const double* pool,
                               const int32_t* idx, double* dst) {
    if (!ctx) return ta_set_err(TA_ERR_INVALID, "ta_gather_boxes: ctx is NULL");
    if (n < 0) return ta_set_err(TA_ERR_INVALID, "ta_gather_boxes: negative size");
    if (n == 0) return TA_OK;
    TA_CUDA(cudaSetDevice(ctx->device));
    ta_begin(ctx, (cudaStream_t)stream);
    k_gather_boxes<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const double2*)pool, idx, (double2*)dst, n);
    return ta_check_launch(ctx, "k_gather_boxes");
}

static int offsets_from_counts(ta_ctx* ctx, cudaStream_t st, int64_t n, const uint16_t* cnt,
                               int64_t* off, int64_t* part) {
    int rc;
    if (n == 0) {
        TA_CUDA(cudaMemsetAsync(off, 0, 8, st));
        return TA_OK;
    }
    const unsigned nb = (unsigned)((n + SC_B - 1) / SC_B);
    k_cnt_sums<<<nb, SC_T, 0, st>>>(cnt, n, part);
    if ((rc = ta_check_launch(ctx, "k_cnt_sums"))) return rc;
    k_cnt_scan_part<<<1, 1024, 0, st>>>(part, (int)nb);
    if ((rc = ta_check_launch(ctx, "k_cnt_scan_part"))) return rc;
    k_cnt_offsets<<<nb, SC_T, 0, st>>>(cnt, n, part, off);
    return ta_check_launch(ctx, "k_cnt_offsets");
}

// off[0] = 0, off[i+1] = off[i] + counts[i]: int64 [n+1] from uint16 [n] (TA_PLAN_GRP_U16)
extern "C" int ta_offsets_from_counts(ta_ctx* ctx, void* stream, int64_t n, const uint16_t* counts,
                                      int64_t* off) {
    if (!ctx) return ta_set_err(TA_ERR_INVALID, "ta_offsets_from_counts: ctx is NULL");
    if (n < 0) return ta_set_err(TA_ERR_INVALID, "ta_offsets_from_counts: negative size");
    TA_CUDA(cudaSetDevice(ctx->device));
    ta_begin(ctx, (cudaStream_t)stream);
    void* part = nullptr;
    int rc = ta_workspace(ctx, (cudaStream_t)stream, (size_t)((n + SC_B - 1) / SC_B + 1) * 8, &part, 1);
    if (rc != TA_OK) return rc;
    return offsets_from_counts(ctx, (cudaStream_t)stream, n, counts, off, (int64_t*)part);
}

// Pinned host memory for plans and results (what makes the copies of ta_eval_plan(s)_host
// asynchronous) without any other CUDA binding on the caller's side.
extern "C" void* ta_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        ta_set_err(TA_ERR_CUDA, "ta_host_alloc: cudaHostAlloc of %s%lld bytes failed", "", (long long)bytes);
        return nullptr;
    }
    return p;
}
extern "C" void ta_host_free(void* p) { if (p) cudaFreeHost(p); }

namespace {
constexpr int TA_MAX_PLANS = 8;
enum { ST_BOX = 0, ST_IN = 1, ST_ACC = 2 };

struct PendingCopy { void* dst; const void* src; size_t bytes; int stage; };

// Device buffers of one plan: allocated on the plan's stream, filled on the upload stream
// (queue() only records the copy), freed in stream order when the object goes away.
struct DevArena {
    cudaStream_t st;
    void* ptrs[80];
    int n = 0;
    PendingCopy cp[48];
    int n_cp = 0;
    int64_t h2d = 0;
    explicit DevArena(cudaStream_t s) : st(s) {}
    ~DevArena() { for (int i = 0; i < n; ++i) cudaFreeAsync(ptrs[i], st); }
    cudaError_t alloc(void** p, size_t bytes) {
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMallocAsync(p, bytes, st);
        if (e == cudaSuccess) ptrs[n++] = *p;
        return e;
    }
    template <typename T>
    cudaError_t queue(const T** dev, const T* host, size_t count, int stage) {
        void* p = nullptr;
        cudaError_t e = alloc(&p, count * sizeof(T));
        if (e != cudaSuccess) return e;
        if (count && host) cp[n_cp++] = {p, host, count * sizeof(T), stage};
        *dev = static_cast<const T*>(p);
        return e;
    }
    cudaError_t flush(int stage, cudaStream_t up) {
        for (int i = 0; i < n_cp; ++i) {
            if (cp[i].stage != stage) continue;
            cudaError_t e = cudaMemcpyAsync(cp[i].dst, cp[i].src, cp[i].bytes, cudaMemcpyHostToDevice, up);
            if (e != cudaSuccess) return e;
            h2d += (int64_t)cp[i].bytes;
        }
        return cudaSuccess;
    }
};

struct EventSet {
    cudaEvent_t ev[TA_MAX_PLANS * 8 + 1];
    int n = 0;
    ~EventSet() { for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]); }
    bool timed = false;
    cudaError_t make(cudaEvent_t* e) {
        cudaError_t r = cudaEventCreateWithFlags(e, timed ? cudaEventDefault : cudaEventDisableTiming);
        if (r == cudaSuccess) ev[n++] = *e;
        return r;
    }
};

struct PlanRun {
    ta_ctx* ctx = nullptr;
    const ta_plan_host* pl = nullptr;
    cudaStream_t st = nullptr;
    DevArena* ar = nullptr;
    bool track = false, is_pool = false;
    const int64_t *grp_dt_off = nullptr, *grp_gt_off = nullptr, *iou_off = nullptr, *cat_dt_off = nullptr,
                  *dt_trk_off = nullptr, *gt_trk_off = nullptr;
    const int32_t *grp_cat = nullptr, *acc_perm = nullptr, *big_list = nullptr, *dt_slot = nullptr,
                  *gt_slot = nullptr, *gt_hp = nullptr, *box_idx = nullptr;
    const double *dt_box = nullptr, *gt_box = nullptr, *dt_a = nullptr, *dt_b = nullptr, *gt_a = nullptr,
                 *gt_b = nullptr, *thrs = nullptr, *recs = nullptr;
    const uint8_t *dt_flag = nullptr, *gt_flag = nullptr;
    const ta_range_cfg* cfgs = nullptr;
    // transport forms, widened on the device
    const float *s_dt32 = nullptr, *s_gt32 = nullptr;
    const uint16_t *s_dslot = nullptr, *s_gslot = nullptr, *s_cnt_dt = nullptr, *s_cnt_gt = nullptr,
                   *s_cat16 = nullptr;
    void *w_dt = nullptr, *w_gt = nullptr, *w_dslot = nullptr, *w_gslot = nullptr, *w_dt_off = nullptr,
         *w_gt_off = nullptr, *w_cat = nullptr, *w_part = nullptr;
    void *d_iou = nullptr, *d_tpfp = nullptr, *d_numgt = nullptr, *d_prec = nullptr, *d_rec = nullptr,
         *d_tp = nullptr, *d_fp = nullptr, *d_word = nullptr;
    cudaEvent_t ev_alloc, ev_box, ev_in, ev_acc, ev_wide, ev_gath, ev_kern, ev_done;
    int64_t d2h = 0;
    int bad = 0;
};

int plan_alloc(PlanRun& r) {
    const ta_plan_host* pl = r.pl;
    DevArena& ar = *r.ar;
    const int64_t G = pl->n_groups;
    const bool pooled = pl->dt_box_idx != nullptr;
    if (pl->flags & TA_PLAN_GRP_U16) {
        TA_CUDA(ar.queue(&r.s_cnt_dt, (const uint16_t*)pl->grp_dt_off, (size_t)G, ST_IN));
        TA_CUDA(ar.queue(&r.s_cnt_gt, (const uint16_t*)pl->grp_gt_off, (size_t)G, ST_IN));
        TA_CUDA(ar.queue(&r.s_cat16, (const uint16_t*)pl->grp_cat, (size_t)G, ST_IN));
        TA_CUDA(ar.alloc(&r.w_dt_off, (size_t)(G + 1) * 8));
        TA_CUDA(ar.alloc(&r.w_gt_off, (size_t)(G + 1) * 8));
        TA_CUDA(ar.alloc(&r.w_cat, (size_t)G * 4));
        TA_CUDA(ar.alloc(&r.w_part, (size_t)((G + SC_B - 1) / SC_B + 1) * 8));
        r.grp_dt_off = (const int64_t*)r.w_dt_off;
        r.grp_gt_off = (const int64_t*)r.w_gt_off;
        r.grp_cat = (const int32_t*)r.w_cat;
    } else {
        TA_CUDA(ar.queue(&r.grp_dt_off, (const int64_t*)pl->grp_dt_off, (size_t)G + 1, ST_IN));
        TA_CUDA(ar.queue(&r.grp_gt_off, (const int64_t*)pl->grp_gt_off, (size_t)G + 1, ST_IN));
        TA_CUDA(ar.queue(&r.grp_cat, (const int32_t*)pl->grp_cat, (size_t)G, ST_IN));
    }
    // the frame path keeps its IoU tiles on chip: iou_off is only read for oversize groups
    if (r.track || pl->n_big > 0) TA_CUDA(ar.queue(&r.iou_off, pl->iou_off, (size_t)G + 1, ST_IN));
    const int box_stage = r.is_pool ? ST_BOX : ST_IN;
    if (pooled) {
        TA_CUDA(ar.queue(&r.box_idx, pl->dt_box_idx, (size_t)pl->n_dt_boxes, ST_IN));
        TA_CUDA(ar.alloc(&r.w_dt, (size_t)pl->n_dt_boxes * 32));
        r.dt_box = (const double*)r.w_dt;
    } else if (pl->flags & TA_PLAN_BOX_F32) {
        TA_CUDA(ar.queue(&r.s_dt32, (const float*)pl->dt_box, (size_t)pl->n_dt_boxes * 4, box_stage));
        TA_CUDA(ar.alloc(&r.w_dt, (size_t)pl->n_dt_boxes * 32));
        r.dt_box = (const double*)r.w_dt;
    } else {
        TA_CUDA(ar.queue(&r.dt_box, (const double*)pl->dt_box, (size_t)pl->n_dt_boxes * 4, box_stage));
    }
    if (pl->flags & TA_PLAN_BOX_F32) {
        TA_CUDA(ar.queue(&r.s_gt32, (const float*)pl->gt_box, (size_t)pl->n_gt_boxes * 4, ST_IN));
        TA_CUDA(ar.alloc(&r.w_gt, (size_t)pl->n_gt_boxes * 32));
        r.gt_box = (const double*)r.w_gt;
    } else {
        TA_CUDA(ar.queue(&r.gt_box, (const double*)pl->gt_box, (size_t)pl->n_gt_boxes * 4, ST_IN));
    }
    TA_CUDA(ar.queue(&r.gt_a, pl->gt_attr_a, (size_t)pl->n_gt, ST_IN));
    TA_CUDA(ar.queue(&r.dt_flag, pl->dt_flag, (size_t)pl->n_dt, ST_IN));
    TA_CUDA(ar.queue(&r.gt_flag, pl->gt_flag, (size_t)pl->n_gt, ST_IN));
    if (r.track) {
        TA_CUDA(ar.queue(&r.dt_trk_off, pl->dt_trk_off, (size_t)pl->n_dt + 1, ST_IN));
        TA_CUDA(ar.queue(&r.gt_trk_off, pl->gt_trk_off, (size_t)pl->n_gt + 1, ST_IN));
        if (pl->flags & TA_PLAN_SLOT_U16) {
            // slots < 65536 travel as uint16 (half the bytes) and are widened on the device
            TA_CUDA(ar.queue(&r.s_dslot, (const uint16_t*)pl->dt_slot, (size_t)pl->n_dt_boxes, ST_IN));
            TA_CUDA(ar.queue(&r.s_gslot, (const uint16_t*)pl->gt_slot, (size_t)pl->n_gt_boxes, ST_IN));
            TA_CUDA(ar.alloc(&r.w_dslot, (size_t)pl->n_dt_boxes * 4));
            TA_CUDA(ar.alloc(&r.w_gslot, (size_t)pl->n_gt_boxes * 4));
            r.dt_slot = (const int32_t*)r.w_dslot;
            r.gt_slot = (const int32_t*)r.w_gslot;
        } else {
            TA_CUDA(ar.queue(&r.dt_slot, (const int32_t*)pl->dt_slot, (size_t)pl->n_dt_boxes, ST_IN));
            TA_CUDA(ar.queue(&r.gt_slot, (const int32_t*)pl->gt_slot, (size_t)pl->n_gt_boxes, ST_IN));
        }
        TA_CUDA(ar.queue(&r.dt_a, pl->dt_attr_a, (size_t)pl->n_dt, ST_IN));
        TA_CUDA(ar.queue(&r.dt_b, pl->dt_attr_b, (size_t)pl->n_dt, ST_IN));
        TA_CUDA(ar.queue(&r.gt_b, pl->gt_attr_b, (size_t)pl->n_gt, ST_IN));
        TA_CUDA(ar.queue(&r.gt_hp, pl->gt_hp, (size_t)pl->n_gt, ST_IN));
    } else if (pl->n_big > 0) {
        TA_CUDA(ar.queue(&r.big_list, pl->big_list, (size_t)pl->n_big, ST_IN));
    }
    TA_CUDA(ar.queue(&r.thrs, pl->iou_thrs, (size_t)pl->n_thr, ST_IN));
    TA_CUDA(ar.queue(&r.cfgs, pl->cfgs, (size_t)pl->n_cfg, ST_IN));
    // read by accumulate only: they travel while the IoU / matching kernels run
    TA_CUDA(ar.queue(&r.cat_dt_off, pl->cat_dt_off, (size_t)pl->n_cat + 1, ST_ACC));
    TA_CUDA(ar.queue(&r.acc_perm, pl->acc_perm, (size_t)pl->n_dt, ST_ACC));
    TA_CUDA(ar.queue(&r.recs, pl->rec_thrs, (size_t)pl->n_rec, ST_ACC));

    const int64_t n_iou = pl->iou_off ? pl->iou_off[G] : 0;
    const size_t n_cell = (size_t)pl->n_thr * pl->n_cat * pl->n_cfg;
    // the frame path keeps IoU tiles on chip; only oversize groups use the global buffer
    TA_CUDA(ar.alloc(&r.d_iou, (r.track || pl->n_big > 0) ? (size_t)n_iou * 8 : 16));
    TA_CUDA(ar.alloc(&r.d_tpfp, (size_t)pl->n_cfg * pl->n_dt * 4));
    if (!r.track) TA_CUDA(ar.alloc(&r.d_word, (size_t)pl->n_dt * 4));
    TA_CUDA(ar.alloc(&r.d_numgt, (size_t)pl->n_cat * pl->n_cfg * 4));
    TA_CUDA(ar.alloc(&r.d_prec, n_cell * pl->n_rec * 8));
    TA_CUDA(ar.alloc(&r.d_rec, n_cell * 8));
    TA_CUDA(ar.alloc(&r.d_tp, n_cell * 8));
    TA_CUDA(ar.alloc(&r.d_fp, n_cell * 8));
    TA_CUDA(cudaMemsetAsync(r.d_numgt, 0, (size_t)pl->n_cat * pl->n_cfg * 4, r.st));
    TA_CUDA(cudaEventRecord(r.ev_alloc, r.st));
    return TA_OK;
}

inline unsigned blocks_of(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

// the pool owner's detection boxes in fp64 (before anybody gathers from them)
int plan_widen_dt(PlanRun& r) {
    const ta_plan_host* pl = r.pl;
    if (r.s_dt32 && pl->n_dt_boxes) {
        k_widen_boxes<<<blocks_of(pl->n_dt_boxes, 256), 256, 0, r.st>>>(
            (const float4*)r.s_dt32, (double*)r.w_dt, pl->n_dt_boxes);
        return ta_check_launch(r.ctx, "k_widen_boxes");
    }
    return TA_OK;
}

int plan_offsets(PlanRun& r, const uint16_t* cnt, void* off) {
    return offsets_from_counts(r.ctx, r.st, r.pl->n_groups, cnt, (int64_t*)off, (int64_t*)r.w_part);
}

int plan_compute(PlanRun& r, PlanRun* pool) {
    const ta_plan_host* pl = r.pl;
    cudaStream_t st = r.st;
    const int64_t G = pl->n_groups;
    int rc;
    TA_CUDA(cudaStreamWaitEvent(st, r.ev_in, 0));
    if (pool) {
        TA_CUDA(cudaStreamWaitEvent(st, pool->ev_wide, 0));
        if (pl->n_dt_boxes) {
            k_gather_boxes<<<blocks_of(pl->n_dt_boxes, 256), 256, 0, st>>>(
                (const double2*)pool->dt_box, r.box_idx, (double2*)r.w_dt, pl->n_dt_boxes);
            if ((rc = ta_check_launch(r.ctx, "k_gather_boxes"))) return rc;
        }
        TA_CUDA(cudaEventRecord(r.ev_gath, st));
    } else if (!r.is_pool) {
        if ((rc = plan_widen_dt(r))) return rc;
    }
    if (r.s_gt32 && pl->n_gt_boxes) {
        k_widen_boxes<<<blocks_of(pl->n_gt_boxes, 256), 256, 0, st>>>(
            (const float4*)r.s_gt32, (double*)r.w_gt, pl->n_gt_boxes);
        if ((rc = ta_check_launch(r.ctx, "k_widen_boxes"))) return rc;
    }
    if (r.s_dslot) {
        if (pl->n_dt_boxes) {
            k_widen_slots<<<blocks_of(pl->n_dt_boxes, 256), 256, 0, st>>>(r.s_dslot, (int32_t*)r.w_dslot, pl->n_dt_boxes);
            if ((rc = ta_check_launch(r.ctx, "k_widen_slots"))) return rc;
        }
        if (pl->n_gt_boxes) {
            k_widen_slots<<<blocks_of(pl->n_gt_boxes, 256), 256, 0, st>>>(r.s_gslot, (int32_t*)r.w_gslot, pl->n_gt_boxes);
            if ((rc = ta_check_launch(r.ctx, "k_widen_slots"))) return rc;
        }
    }
    if (r.s_cnt_dt) {
        if ((rc = plan_offsets(r, r.s_cnt_dt, r.w_dt_off))) return rc;
        if ((rc = plan_offsets(r, r.s_cnt_gt, r.w_gt_off))) return rc;
        if (G) {
            k_widen_slots<<<blocks_of(G, 256), 256, 0, st>>>(r.s_cat16, (int32_t*)r.w_cat, G);
            if ((rc = ta_check_launch(r.ctx, "k_widen_slots"))) return rc;
        }
    }
    if (r.track) {
        rc = ta_track_iou(r.ctx, st, pl->iou_mode, G, r.grp_dt_off, r.grp_gt_off, r.dt_trk_off, r.dt_box,
                          r.dt_slot, r.gt_trk_off, r.gt_box, r.gt_slot, pl->n_slots_max, r.iou_off,
                          (double*)r.d_iou);
        if (rc == TA_OK)
            rc = ta_match_greedy(r.ctx, st, G, nullptr, 0, r.grp_dt_off, r.grp_gt_off, r.grp_cat, r.iou_off,
                                 (const double*)r.d_iou, pl->n_thr, r.thrs, pl->n_cfg, r.cfgs,
                                 pl->n_dt, r.dt_a, r.dt_b, r.dt_flag, pl->n_gt, r.gt_a, r.gt_b, r.gt_hp,
                                 r.gt_flag, pl->g_max, (uint32_t*)r.d_tpfp, (int32_t*)r.d_numgt,
                                 nullptr, nullptr);
    } else {
        rc = ta_frame_eval(r.ctx, st, G, r.grp_dt_off, r.grp_gt_off, r.grp_cat, r.dt_box, r.gt_box,
                           pl->n_thr, r.thrs, pl->n_cfg, r.cfgs, pl->n_dt, r.dt_flag, pl->n_gt, r.gt_a,
                           r.gt_flag, pl->n_big, r.big_list, pl->g_max, r.iou_off, (double*)r.d_iou, 0,
                           nullptr, (uint32_t*)r.d_word, (uint32_t*)r.d_tpfp, (int32_t*)r.d_numgt,
                           nullptr, nullptr);
    }
    if (rc != TA_OK) return rc;
    TA_CUDA(cudaStreamWaitEvent(st, r.ev_acc, 0));
    return ta_pr_accumulate(r.ctx, st, pl->n_cat, r.cat_dt_off, r.acc_perm, pl->n_dt,
                            (const uint32_t*)r.d_tpfp, (const uint32_t*)r.d_word,
                            (const int32_t*)r.d_numgt, pl->n_thr, pl->n_cfg, pl->n_rec, r.recs,
                            (double*)r.d_prec, (double*)r.d_rec, (int64_t*)r.d_tp, (int64_t*)r.d_fp);
}

int plan_download(PlanRun& r, const ta_host_out& o) {
    const ta_plan_host* pl = r.pl;
    cudaStream_t st = r.st;
    const size_t n_cell = (size_t)pl->n_thr * pl->n_cat * pl->n_cfg, n_prec = n_cell * pl->n_rec;
    TA_CUDA(cudaMemcpyAsync(o.precision, r.d_prec, n_prec * 8, cudaMemcpyDeviceToHost, st));
    TA_CUDA(cudaMemcpyAsync(o.recall, r.d_rec, n_cell * 8, cudaMemcpyDeviceToHost, st));
    r.d2h += (int64_t)((n_prec + n_cell) * 8);
    if (o.tp_cnt) { TA_CUDA(cudaMemcpyAsync(o.tp_cnt, r.d_tp, n_cell * 8, cudaMemcpyDeviceToHost, st)); r.d2h += n_cell * 8; }
    if (o.fp_cnt) { TA_CUDA(cudaMemcpyAsync(o.fp_cnt, r.d_fp, n_cell * 8, cudaMemcpyDeviceToHost, st)); r.d2h += n_cell * 8; }
    if (o.num_gt) {
        TA_CUDA(cudaMemcpyAsync(o.num_gt, r.d_numgt, (size_t)pl->n_cat * pl->n_cfg * 4, cudaMemcpyDeviceToHost, st));
        r.d2h += (int64_t)pl->n_cat * pl->n_cfg * 4;
    }
    return TA_OK;
}

// the tiled IoU kernel counted the pairs with intersection > union (eval.py:95).  The copy
// lands in pageable memory, i.e. it blocks the host: issued after everything else is enqueued.
int plan_read_asserts(PlanRun& r) {
    if (r.track && r.pl->iou_mode == TA_IOU_3D) {
        TA_CUDA(cudaMemcpyAsync(&r.bad, r.ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, r.st));
        TA_CUDA(cudaMemsetAsync(r.ctx->d_flags, 0, sizeof(int), r.st));
    }
    return TA_OK;
}
}  // namespace

extern "C" int ta_eval_plans_host(int32_t n_plans, ta_ctx* const* ctxs,
                                  const ta_plan_host* const* plans, const ta_host_out* outs,
                                  int64_t* h2d_bytes, int64_t* d2h_bytes) {
    if (n_plans < 1 || n_plans > TA_MAX_PLANS || !ctxs || !plans || !outs)
        return ta_set_err(TA_ERR_INVALID, "ta_eval_plans_host: bad arguments (1..%s%lld plans)", "", TA_MAX_PLANS);
    PlanRun run[TA_MAX_PLANS];
    for (int i = 0; i < n_plans; ++i) {
        if (!ctxs[i] || !plans[i] || !outs[i].precision || !outs[i].recall)
            return ta_set_err(TA_ERR_INVALID, "ta_eval_plans_host: NULL argument");
        for (int j = 0; j < i; ++j)
            if (ctxs[j] == ctxs[i])
                return ta_set_err(TA_ERR_INVALID, "ta_eval_plans_host: every plan needs its own context");
        if (ctxs[i]->device != ctxs[0]->device)
            return ta_set_err(TA_ERR_INVALID, "ta_eval_plans_host: contexts on different devices");
        const ta_plan_host* pl = plans[i];
        if (pl->dt_box_idx) {
            const int p = pl->dt_box_pool;
            if (p < 0 || p >= n_plans || p == i || plans[p]->dt_box_idx || !plans[p]->dt_box)
                return ta_set_err(TA_ERR_INVALID, "ta_eval_plans_host: dt_box_pool of plan %s%lld is not a plan with boxes", "", i);
        } else if (!pl->dt_box && pl->n_dt_boxes) {
            return ta_set_err(TA_ERR_INVALID, "ta_eval_plans_host: plan %s%lld has no detection boxes", "", i);
        }
        run[i].ctx = ctxs[i];
        run[i].pl = pl;
        run[i].st = ctxs[i]->own_stream;
        run[i].track = pl->dt_trk_off != nullptr;
    }
    for (int i = 0; i < n_plans; ++i)
        if (plans[i]->dt_box_idx) run[plans[i]->dt_box_pool].is_pool = true;
    TA_CUDA(cudaSetDevice(ctxs[0]->device));
    ta_ctx* c0 = ctxs[0];
    if (!c0->copy_stream) TA_CUDA(cudaStreamCreateWithFlags(&c0->copy_stream, cudaStreamNonBlocking));
    cudaStream_t up = c0->copy_stream;
    int rc = TA_OK;
    {
        EventSet evs;
        // TA_PIPE_TRACE=1: timeline of the call on stderr (ms since the first upload)
        const bool trace = getenv("TA_PIPE_TRACE") != nullptr;
        evs.timed = trace;
        cudaEvent_t ev_start;
        TA_CUDA(evs.make(&ev_start));
        // arenas are declared after the events: they free (stream-ordered) before the events go
        DevArena* arenas[TA_MAX_PLANS] = {nullptr};
        struct ArenaGuard {
            DevArena** a; int n;
            ~ArenaGuard() { for (int i = 0; i < n; ++i) delete a[i]; }
        } guard{arenas, n_plans};
        for (int i = 0; i < n_plans && rc == TA_OK; ++i) {
            PlanRun& r = run[i];
            ta_begin(r.ctx, r.st);
            arenas[i] = r.ar = new DevArena(r.st);
            for (cudaEvent_t* e : {&r.ev_alloc, &r.ev_box, &r.ev_in, &r.ev_acc, &r.ev_wide, &r.ev_gath,
                                   &r.ev_kern, &r.ev_done})
                TA_CUDA(evs.make(e));
            rc = plan_alloc(r);
            if (rc == TA_OK) TA_CUDA(cudaStreamWaitEvent(up, r.ev_alloc, 0));
        }
        // uploads, in the order the plans will consume them
        if (rc == TA_OK) TA_CUDA(cudaEventRecord(ev_start, up));
        for (int i = 0; i < n_plans && rc == TA_OK; ++i)
            if (run[i].is_pool) {
                TA_CUDA(run[i].ar->flush(ST_BOX, up));
                TA_CUDA(cudaEventRecord(run[i].ev_box, up));
            }
        for (int i = 0; i < n_plans && rc == TA_OK; ++i) {
            TA_CUDA(run[i].ar->flush(ST_IN, up));
            TA_CUDA(cudaEventRecord(run[i].ev_in, up));
            TA_CUDA(run[i].ar->flush(ST_ACC, up));
            TA_CUDA(cudaEventRecord(run[i].ev_acc, up));
        }
        // pool owners widen their boxes first, so that the gathers of the other plans can wait
        // on an event that has been recorded
        for (int i = 0; i < n_plans && rc == TA_OK; ++i)
            if (run[i].is_pool) {
                TA_CUDA(cudaStreamWaitEvent(run[i].st, run[i].ev_box, 0));
                rc = plan_widen_dt(run[i]);
                if (rc == TA_OK) TA_CUDA(cudaEventRecord(run[i].ev_wide, run[i].st));
            }
        for (int i = 0; i < n_plans && rc == TA_OK; ++i) {
            PlanRun* pool = plans[i]->dt_box_idx ? &run[plans[i]->dt_box_pool] : nullptr;
            rc = plan_compute(run[i], pool);
            if (rc == TA_OK && trace) TA_CUDA(cudaEventRecord(run[i].ev_kern, run[i].st));
            if (rc == TA_OK) rc = plan_download(run[i], outs[i]);
            if (rc == TA_OK && trace) TA_CUDA(cudaEventRecord(run[i].ev_done, run[i].st));
        }
        // a pool's boxes stay until every gather from them has run
        for (int i = 0; i < n_plans && rc == TA_OK; ++i)
            if (plans[i]->dt_box_idx)
                TA_CUDA(cudaStreamWaitEvent(run[plans[i]->dt_box_pool].st, run[i].ev_gath, 0));
        if (rc != TA_OK) {
            // copies may still be in flight into buffers about to be freed
            cudaStreamSynchronize(up);
            for (int i = 0; i < n_plans; ++i) cudaStreamSynchronize(run[i].st);
        }
        for (int i = 0; i < n_plans; ++i) {
            if (h2d_bytes) h2d_bytes[i] = run[i].ar ? run[i].ar->h2d : 0;
            if (d2h_bytes) d2h_bytes[i] = run[i].d2h;
        }
        if (trace && rc == TA_OK) {
            for (int i = 0; i < n_plans; ++i) cudaStreamSynchronize(run[i].st);
            for (int i = 0; i < n_plans; ++i) {
                float t_box = 0, t_in = 0, t_acc = 0, t_k = 0, t_d = 0;
                if (run[i].is_pool) cudaEventElapsedTime(&t_box, ev_start, run[i].ev_box);
                cudaEventElapsedTime(&t_in, ev_start, run[i].ev_in);
                cudaEventElapsedTime(&t_acc, ev_start, run[i].ev_acc);
                cudaEventElapsedTime(&t_k, ev_start, run[i].ev_kern);
                cudaEventElapsedTime(&t_d, ev_start, run[i].ev_done);
                fprintf(stderr, "ta_pipe plan %d (%s): pool boxes up %.2f | inputs up %.2f | accumulate inputs up %.2f | "
                        "kernels done %.2f | results on host %.2f ms  (h2d %.1f MB, d2h %.1f MB)\n",
                        i, run[i].track ? "track" : "frame", t_box, t_in, t_acc, t_k, t_d,
                        run[i].ar->h2d / 1e6, run[i].d2h / 1e6);
            }
        }
    }   // arena frees are stream-ordered after the kernels
    for (int i = 0; i < n_plans && rc == TA_OK; ++i) rc = plan_read_asserts(run[i]);
    for (int i = 0; i < n_plans; ++i) {
        cudaError_t e = cudaStreamSynchronize(run[i].st);
        if (e != cudaSuccess && rc == TA_OK)
            rc = ta_set_err(TA_ERR_CUDA, "CUDA error %s in ta_eval_plans_host", cudaGetErrorString(e));
    }
    if (rc != TA_OK) return rc;
    for (int i = 0; i < n_plans; ++i)
        if (run[i].bad)
            return ta_set_err(TA_ERR_ASSERT, "track IoU: intersection exceeds union in %s%lld pairs", "", run[i].bad);
    return TA_OK;
}

extern "C" int ta_eval_plan_host(ta_ctx* ctx, const ta_plan_host* pl,
                                 double* precision, double* recall,
                                 int64_t* tp_cnt, int64_t* fp_cnt, int32_t* num_gt,
                                 int64_t* h2d_bytes, int64_t* d2h_bytes) {
    if (!ctx || !pl || !precision || !recall)
        return ta_set_err(TA_ERR_INVALID, "ta_eval_plan_host: NULL argument");
    if (pl->dt_box_idx)
        return ta_set_err(TA_ERR_INVALID, "ta_eval_plan_host: a plan that shares boxes needs ta_eval_plans_host");
    const ta_host_out out = {precision, recall, tp_cnt, fp_cnt, num_gt};
    return ta_eval_plans_host(1, &ctx, &pl, &out, h2d_bytes, d2h_bytes);
}
