// ta_rle.cu — mask IoU of the segm evaluation path (sm_100a).
//
// Reference: LVISEval.compute_iou with iou_type = "segm" (lvis_amodal/eval.py:168-192) ->
// pycocotools.mask.iou -> rleIou (in-tree copy maskApi.c:78-96).  One CTA per (image, category)
// group, one thread per (detection, GT) pair: the pair's two run lists are consumed in lock
// step (ta_rle_pair_iou).  Pairs whose run-derived boxes do not overlap never touch their runs,
// which is the common case; the rest is a latency-bound integer walk, so the grid is sized by
// groups and the block by the typical pair count of a group.
#include "ta_internal.h"
#include "ta_device_fns.cuh"

struct RleIouArgs {
    const int32_t* grp_list;
    const int64_t* grp_dt_off;
    const int64_t* grp_gt_off;
    const int64_t* dt_rle_off;
    const uint32_t* dt_counts;
    const uint32_t* dt_hw;
    const double* dt_bb;
    const int64_t* gt_rle_off;
    const uint32_t* gt_counts;
    const uint32_t* gt_hw;
    const double* gt_bb;
    const int64_t* iou_off;
    double* iou;
};

#define RLE_THREADS 64

__global__ void __launch_bounds__(RLE_THREADS)
k_rle_iou(RleIouArgs a) {
    const int64_t grp = a.grp_list ? (int64_t)a.grp_list[blockIdx.x] : (int64_t)blockIdx.x;
    const int64_t d0 = a.grp_dt_off[grp], g0 = a.grp_gt_off[grp];
    const int64_t D = a.grp_dt_off[grp + 1] - d0, G = a.grp_gt_off[grp + 1] - g0;
    if (D == 0 || G == 0) return;
    double* out = a.iou + a.iou_off[grp];
    for (int64_t e = threadIdx.x; e < D * G; e += RLE_THREADS) {
        const int64_t d = d0 + e / G, g = g0 + e % G;
        const int64_t od = a.dt_rle_off[d], og = a.gt_rle_off[g];
        out[e] = ta_rle_pair_iou(a.dt_counts + od, a.dt_rle_off[d + 1] - od,
                                 a.gt_counts + og, a.gt_rle_off[g + 1] - og,
                                 a.dt_bb + 4 * d, a.gt_bb + 4 * g,
                                 a.dt_hw[2 * d], a.dt_hw[2 * d + 1], a.gt_hw[2 * g], a.gt_hw[2 * g + 1]);
    }
}

extern "C" int ta_rle_iou(ta_ctx* ctx, void* stream, int64_t n_groups,
                          const int32_t* grp_list, int64_t n_list,
                          const int64_t* grp_dt_off, const int64_t* grp_gt_off,
                          const int64_t* dt_rle_off, const uint32_t* dt_counts,
                          const uint32_t* dt_hw, const double* dt_bbox,
                          const int64_t* gt_rle_off, const uint32_t* gt_counts,
                          const uint32_t* gt_hw, const double* gt_bbox,
                          const int64_t* iou_off, double* iou_out) {
    if (!ctx) return ta_set_err(TA_ERR_INVALID, "ta_rle_iou: ctx is NULL");
    const int64_t n = grp_list ? n_list : n_groups;
    if (n < 0 || n > 0x7fffffffLL) return ta_set_err(TA_ERR_INVALID, "ta_rle_iou: bad group count");
    if (n == 0) return TA_OK;
    if (!grp_dt_off || !grp_gt_off || !dt_rle_off || !gt_rle_off || !dt_counts || !gt_counts ||
        !dt_hw || !gt_hw || !dt_bbox || !gt_bbox || !iou_off || !iou_out)
        return ta_set_err(TA_ERR_INVALID, "ta_rle_iou: NULL argument");
    TA_CUDA(cudaSetDevice(ctx->device));
    ta_begin(ctx, (cudaStream_t)stream);
    RleIouArgs a{grp_list, grp_dt_off, grp_gt_off, dt_rle_off, dt_counts, dt_hw, dt_bbox,
                 gt_rle_off, gt_counts, gt_hw, gt_bbox, iou_off, iou_out};
    k_rle_iou<<<(unsigned)n, RLE_THREADS, 0, (cudaStream_t)stream>>>(a);
    return ta_check_launch(ctx, "k_rle_iou");
}
