// ta_pipeline.cu — ta_eval_plan_host: H2D -> IoU -> match -> accumulate -> D2H on the context's
// stream.  Device memory comes from the stream-ordered pool (cudaMallocAsync; the context sets
// the pool's release threshold so repeated calls reuse the same pages).
#include "ta_internal.h"

// frame slots shipped as uint16 (TA_PLAN_SLOT_U16) -> int32
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

namespace {
struct DevArena {
    cudaStream_t st;
    void* ptrs[64];
    int n = 0;
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
    cudaError_t upload(const T** dev, const T* host, size_t count) {
        void* p = nullptr;
        cudaError_t e = alloc(&p, count * sizeof(T));
        if (e != cudaSuccess) return e;
        if (count && host) {
            e = cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, st);
            h2d += (int64_t)(count * sizeof(T));
        }
        *dev = static_cast<const T*>(p);
        return e;
    }
};
}  // namespace

extern "C" int ta_eval_plan_host(ta_ctx* ctx, const ta_plan_host* pl,
                                 double* precision, double* recall,
                                 int64_t* tp_cnt, int64_t* fp_cnt, int32_t* num_gt,
                                 int64_t* h2d_bytes, int64_t* d2h_bytes) {
    if (!ctx || !pl || !precision || !recall)
        return ta_set_err(TA_ERR_INVALID, "ta_eval_plan_host: NULL argument");
    TA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->own_stream;
    ta_begin(ctx, st);
    const bool track = pl->dt_trk_off != nullptr;
    const int64_t G = pl->n_groups;
    int rc = TA_OK;
    {
        DevArena ar(st);
        const int64_t *grp_dt_off, *grp_gt_off, *iou_off, *cat_dt_off, *dt_trk_off = nullptr,
                      *gt_trk_off = nullptr;
        const int32_t *grp_cat, *acc_perm, *big_list = nullptr, *dt_slot = nullptr,
                      *gt_slot = nullptr, *gt_hp = nullptr;
        const double *dt_box, *gt_box, *dt_a = nullptr, *dt_b = nullptr, *gt_a, *gt_b = nullptr,
                     *thrs, *recs;
        const uint8_t *dt_flag, *gt_flag;
        const ta_range_cfg* cfgs;
        TA_CUDA(ar.upload(&grp_dt_off, pl->grp_dt_off, G + 1));
        TA_CUDA(ar.upload(&grp_gt_off, pl->grp_gt_off, G + 1));
        // the frame path keeps its IoU tiles on chip: iou_off is only read for oversize groups
        if (track || pl->n_big > 0) TA_CUDA(ar.upload(&iou_off, pl->iou_off, G + 1));
        else iou_off = nullptr;
        TA_CUDA(ar.upload(&cat_dt_off, pl->cat_dt_off, (size_t)pl->n_cat + 1));
        TA_CUDA(ar.upload(&grp_cat, pl->grp_cat, G));
        TA_CUDA(ar.upload(&acc_perm, pl->acc_perm, pl->n_dt));
        if (pl->flags & TA_PLAN_BOX_F32) {
            const float *s_dt, *s_gt;
            void *w_dt, *w_gt;
            TA_CUDA(ar.upload(&s_dt, (const float*)pl->dt_box, (size_t)pl->n_dt_boxes * 4));
            TA_CUDA(ar.upload(&s_gt, (const float*)pl->gt_box, (size_t)pl->n_gt_boxes * 4));
            TA_CUDA(ar.alloc(&w_dt, (size_t)pl->n_dt_boxes * 32));
            TA_CUDA(ar.alloc(&w_gt, (size_t)pl->n_gt_boxes * 32));
            if (pl->n_dt_boxes) {
                k_widen_boxes<<<(unsigned)((pl->n_dt_boxes + 255) / 256), 256, 0, st>>>(
                    (const float4*)s_dt, (double*)w_dt, pl->n_dt_boxes);
                if ((rc = ta_check_launch(ctx, "k_widen_boxes"))) return rc;
            }
            if (pl->n_gt_boxes) {
                k_widen_boxes<<<(unsigned)((pl->n_gt_boxes + 255) / 256), 256, 0, st>>>(
                    (const float4*)s_gt, (double*)w_gt, pl->n_gt_boxes);
                if ((rc = ta_check_launch(ctx, "k_widen_boxes"))) return rc;
            }
            dt_box = (const double*)w_dt;
            gt_box = (const double*)w_gt;
        } else {
            TA_CUDA(ar.upload(&dt_box, (const double*)pl->dt_box, (size_t)pl->n_dt_boxes * 4));
            TA_CUDA(ar.upload(&gt_box, (const double*)pl->gt_box, (size_t)pl->n_gt_boxes * 4));
        }
        TA_CUDA(ar.upload(&gt_a, pl->gt_attr_a, pl->n_gt));
        TA_CUDA(ar.upload(&dt_flag, pl->dt_flag, pl->n_dt));
        TA_CUDA(ar.upload(&gt_flag, pl->gt_flag, pl->n_gt));
        if (track) {
            TA_CUDA(ar.upload(&dt_trk_off, pl->dt_trk_off, pl->n_dt + 1));
            TA_CUDA(ar.upload(&gt_trk_off, pl->gt_trk_off, pl->n_gt + 1));
            if (pl->flags & TA_PLAN_SLOT_U16) {
                // slots < 65536 travel as uint16 (half the bytes) and are widened on the device
                const uint16_t *s_dt, *s_gt;
                void *w_dt, *w_gt;
                TA_CUDA(ar.upload(&s_dt, (const uint16_t*)pl->dt_slot, (size_t)pl->n_dt_boxes));
                TA_CUDA(ar.upload(&s_gt, (const uint16_t*)pl->gt_slot, (size_t)pl->n_gt_boxes));
                TA_CUDA(ar.alloc(&w_dt, (size_t)pl->n_dt_boxes * 4));
                TA_CUDA(ar.alloc(&w_gt, (size_t)pl->n_gt_boxes * 4));
                if (pl->n_dt_boxes) {
                    k_widen_slots<<<(unsigned)((pl->n_dt_boxes + 255) / 256), 256, 0, st>>>(s_dt, (int32_t*)w_dt, pl->n_dt_boxes);
                    if ((rc = ta_check_launch(ctx, "k_widen_slots"))) return rc;
                }
                if (pl->n_gt_boxes) {
                    k_widen_slots<<<(unsigned)((pl->n_gt_boxes + 255) / 256), 256, 0, st>>>(s_gt, (int32_t*)w_gt, pl->n_gt_boxes);
                    if ((rc = ta_check_launch(ctx, "k_widen_slots"))) return rc;
                }
                dt_slot = (const int32_t*)w_dt;
                gt_slot = (const int32_t*)w_gt;
            } else {
                TA_CUDA(ar.upload(&dt_slot, (const int32_t*)pl->dt_slot, pl->n_dt_boxes));
                TA_CUDA(ar.upload(&gt_slot, (const int32_t*)pl->gt_slot, pl->n_gt_boxes));
            }
            TA_CUDA(ar.upload(&dt_a, pl->dt_attr_a, pl->n_dt));
            TA_CUDA(ar.upload(&dt_b, pl->dt_attr_b, pl->n_dt));
            TA_CUDA(ar.upload(&gt_b, pl->gt_attr_b, pl->n_gt));
            TA_CUDA(ar.upload(&gt_hp, pl->gt_hp, pl->n_gt));
        } else if (pl->n_big > 0) {
            TA_CUDA(ar.upload(&big_list, pl->big_list, pl->n_big));
        }
        TA_CUDA(ar.upload(&thrs, pl->iou_thrs, pl->n_thr));
        TA_CUDA(ar.upload(&recs, pl->rec_thrs, pl->n_rec));
        TA_CUDA(ar.upload(&cfgs, pl->cfgs, pl->n_cfg));

        const int64_t n_iou = pl->iou_off ? pl->iou_off[G] : 0;
        const size_t n_cell = (size_t)pl->n_thr * pl->n_cat * pl->n_cfg;
        const size_t n_prec = n_cell * pl->n_rec;
        void *d_iou, *d_tpfp, *d_numgt, *d_prec, *d_rec, *d_tp, *d_fp, *d_word = nullptr;
        // the frame path keeps IoU tiles on chip; only oversize groups use the global buffer
        TA_CUDA(ar.alloc(&d_iou, (track || pl->n_big > 0) ? (size_t)n_iou * 8 : 16));
        TA_CUDA(ar.alloc(&d_tpfp, (size_t)pl->n_cfg * pl->n_dt * 4));
        if (!track) TA_CUDA(ar.alloc(&d_word, (size_t)pl->n_dt * 4));
        TA_CUDA(ar.alloc(&d_numgt, (size_t)pl->n_cat * pl->n_cfg * 4));
        TA_CUDA(ar.alloc(&d_prec, n_prec * 8));
        TA_CUDA(ar.alloc(&d_rec, n_cell * 8));
        TA_CUDA(ar.alloc(&d_tp, n_cell * 8));
        TA_CUDA(ar.alloc(&d_fp, n_cell * 8));
        TA_CUDA(cudaMemsetAsync(d_numgt, 0, (size_t)pl->n_cat * pl->n_cfg * 4, st));

        if (track) {
            rc = ta_track_iou(ctx, st, pl->iou_mode, G, grp_dt_off, grp_gt_off, dt_trk_off, dt_box,
                              dt_slot, gt_trk_off, gt_box, gt_slot, pl->n_slots_max, iou_off,
                              (double*)d_iou);
            if (rc == TA_OK)
                rc = ta_match_greedy(ctx, st, G, nullptr, 0, grp_dt_off, grp_gt_off, grp_cat, iou_off,
                                     (const double*)d_iou, pl->n_thr, thrs, pl->n_cfg, cfgs,
                                     pl->n_dt, dt_a, dt_b, dt_flag, pl->n_gt, gt_a, gt_b, gt_hp,
                                     gt_flag, pl->g_max, (uint32_t*)d_tpfp, (int32_t*)d_numgt,
                                     nullptr, nullptr);
        } else {
            rc = ta_frame_eval(ctx, st, G, grp_dt_off, grp_gt_off, grp_cat, dt_box, gt_box,
                               pl->n_thr, thrs, pl->n_cfg, cfgs, pl->n_dt, dt_flag, pl->n_gt, gt_a,
                               gt_flag, pl->n_big, big_list, pl->g_max, iou_off, (double*)d_iou, 0,
                               nullptr, (uint32_t*)d_word, (uint32_t*)d_tpfp, (int32_t*)d_numgt,
                               nullptr, nullptr);
        }
        if (rc == TA_OK)
            rc = ta_pr_accumulate(ctx, st, pl->n_cat, cat_dt_off, acc_perm, pl->n_dt,
                                  (const uint32_t*)d_tpfp, (const uint32_t*)d_word,
                                  (const int32_t*)d_numgt, pl->n_thr,
                                  pl->n_cfg, pl->n_rec, recs, (double*)d_prec, (double*)d_rec,
                                  (int64_t*)d_tp, (int64_t*)d_fp);
        int64_t d2h = 0;
        if (rc == TA_OK) {
            TA_CUDA(cudaMemcpyAsync(precision, d_prec, n_prec * 8, cudaMemcpyDeviceToHost, st));
            TA_CUDA(cudaMemcpyAsync(recall, d_rec, n_cell * 8, cudaMemcpyDeviceToHost, st));
            d2h += (int64_t)((n_prec + n_cell) * 8);
            if (tp_cnt) { TA_CUDA(cudaMemcpyAsync(tp_cnt, d_tp, n_cell * 8, cudaMemcpyDeviceToHost, st)); d2h += n_cell * 8; }
            if (fp_cnt) { TA_CUDA(cudaMemcpyAsync(fp_cnt, d_fp, n_cell * 8, cudaMemcpyDeviceToHost, st)); d2h += n_cell * 8; }
            if (num_gt) { TA_CUDA(cudaMemcpyAsync(num_gt, d_numgt, (size_t)pl->n_cat * pl->n_cfg * 4, cudaMemcpyDeviceToHost, st)); d2h += (int64_t)pl->n_cat * pl->n_cfg * 4; }
        }
        if (h2d_bytes) *h2d_bytes = ar.h2d;
        if (d2h_bytes) *d2h_bytes = d2h;
    }   // arena frees are stream-ordered after the kernels
    int bad = 0;
    if (rc == TA_OK && track && pl->iou_mode == TA_IOU_3D) {
        // the tiled IoU kernel counted the pairs with intersection > union (eval.py:95)
        TA_CUDA(cudaMemcpyAsync(&bad, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
        TA_CUDA(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), st));
    }
    TA_CUDA(cudaStreamSynchronize(st));
    if (bad)
        return ta_set_err(TA_ERR_ASSERT, "track IoU: intersection exceeds union in %s%lld pairs", "", bad);
    return rc;
}
