/*
 * ta_mask.h — C ABI of the run-length mask codec (host side, in libta_ingest.so, built from
 * tao_amodal_b200/csrc/ta_mask.cpp with g++; no CUDA) behind LVISEval(iou_type="segm").  It only
 * CONVERTS annotations to run lengths; the mask IoU itself runs on the device (ta_rle_iou,
 * include/ta_eval.h).
 *
 * The reference gets these from third-party pycocotools (lvis_amodal/lvis.py:155-192
 * ann_to_rle, results.py:58-66, eval.py:54-57,180-191); the in-tree copy of that code is
 * visualization/tao/third_party/pysot/training_dataset/coco/pycocotools/common/maskApi.c,
 * cited per function.  Masks are column-major run lengths: counts[0] zeros, counts[1] ones, ...
 *
 * A pool is an append-only list of masks owned by the library; every accessor copies into
 * caller-owned buffers.  Functions return 0 / a count on success, a negative value on error
 * (ta_mask_error() has the message).  Not thread-safe per pool.
 */
#ifndef TA_MASK_H
#define TA_MASK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ta_rle_pool ta_rle_pool;

ta_rle_pool* ta_rle_pool_create(void);
void         ta_rle_pool_destroy(ta_rle_pool* p);
const char*  ta_mask_error(void);

/* One mask from the union of n_parts polygons (part i = xy[part_off[i] .. part_off[i+1]) as
 * x0,y0,x1,y1,...): frPyObjects + merge of LVIS.ann_to_rle (lvis.py:168-173) — rleFrPoly
 * maskApi.c:164-216 per part, rleMerge :50-71 (union) over the parts.  Returns the mask index. */
int64_t ta_rle_pool_add_polygons(ta_rle_pool* p, int64_t n_parts, const int64_t* part_off,
                                 const double* xy, int64_t h, int64_t w);
/* n masks from boxes [x, y, w, h] (rleFrBbox maskApi.c:153-160: the 4-corner polygon that
 * lvis_amodal/results.py:50-52 synthesises for bbox results).  h / w per box.  Returns the
 * index of the first mask added. */
int64_t ta_rle_pool_add_boxes(ta_rle_pool* p, int64_t n, const double* boxes,
                              const int64_t* h, const int64_t* w);
/* One mask from uncompressed counts (frUncompressedRLE, _mask.pyx:270-286). */
int64_t ta_rle_pool_add_counts(ta_rle_pool* p, int64_t m, const uint32_t* counts, int64_t h, int64_t w);
/* One mask from a compressed counts string (rleFrString maskApi.c:233-246); len bytes. */
int64_t ta_rle_pool_add_string(ta_rle_pool* p, const char* s, int64_t len, int64_t h, int64_t w);

int64_t ta_rle_pool_size(const ta_rle_pool* p);           /* number of masks */
int64_t ta_rle_pool_total_counts(const ta_rle_pool* p);   /* sum of run counts over all masks */
/* Flat export of the whole pool: off int64 [n+1] (run offsets), counts uint32 [total],
 * hw uint32 [n][2], bbox double [n][4] (rleToBbox maskApi.c:133-151), area uint32 [n]
 * (rleArea :73-76).  Any pointer may be NULL. */
int ta_rle_pool_export(const ta_rle_pool* p, int64_t* off, uint32_t* counts, uint32_t* hw,
                       double* bbox, uint32_t* area);
/* Compressed string of mask i (rleToString maskApi.c:218-231) into buf (cap bytes incl. NUL);
 * returns its length, or the needed capacity negated when cap is too small. */
int64_t ta_rle_pool_to_string(const ta_rle_pool* p, int64_t i, char* buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* TA_MASK_H */
