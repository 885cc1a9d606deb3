/*
 * ta_eval.h — C ABI of the B200-native TAO-Amodal evaluation hot path.
 *
 * The reference (WesleyHsieh0806/TAO-Amodal) has no FFI for this path: its hot
 * loops are Python methods.  Each entry point below replaces the Python
 * function(s) named in its comment (paths relative to the reference root); a
 * maintainer binds them with ctypes from those methods.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++ / torch types.
 *   - every function returns 0 on success or a negative ta_status; the message
 *     of the last failure on the calling thread is ta_last_error().
 *   - the caller owns every buffer; the library never frees or retains a
 *     pointer after the call returns (device calls are asynchronous on
 *     `stream`, so buffers must outlive the stream work).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - all offsets arrays are exclusive prefix sums with n+1 entries (CSR).
 *   - boxes are fp64 [x, y, w, h]; scores fp64; ids int64; arithmetic is IEEE
 *     fp64 without FMA contraction, bit-compatible with the reference's
 *     bb_intersect_union (tao_amodal/evaluation/tao_amodal/eval.py:15-48).
 *   - pointers are DEVICE pointers unless the function name ends in _host.
 */
#ifndef TA_EVAL_H
#define TA_EVAL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TA_ABI_VERSION 4
#define TA_MAX_THRS 16   /* IoU thresholds packed as 16 TP bits + 16 FP bits per detection */

typedef enum ta_status {
    TA_OK = 0,
    TA_ERR_INVALID = -1,     /* bad argument */
    TA_ERR_CUDA = -2,        /* CUDA runtime error (message has the detail) */
    TA_ERR_TOO_LARGE = -3,   /* a group exceeds the on-chip capacity of the kernel */
    TA_ERR_ASSERT = -4,      /* the reference would have raised AssertionError (eval.py:95) */
    TA_ERR_NCCL = -5
} ta_status;

/* 3-D IoU flavours of TaoEval (Params.iou_3d_type, eval.py:753-757) */
typedef enum ta_iou_mode {
    TA_IOU_3D = 0,           /* sum_t I / sum_t U, tiled kernel (eval.py:73-96) */
    TA_IOU_AVG = 1,          /* mean_t (I/U)                     (eval.py:99-117) */
    TA_IOU_IMAGENETVID = 2,  /* frac of frames with I > 0.5 U    (eval.py:51-70) */
    TA_IOU_3D_SEQ = 3        /* as TA_IOU_3D, one thread per track pair, terms added
                                sequentially in ascending frame order */
} ta_iou_mode;

/* One (area, duration) cell of TaoEval.Params (eval.py:735-744) or one visibility range
 * of LVISEval.Params (lvis_amodal/eval.py:567-575).  A ground-truth entity is ignored
 * when  flag&1  ||  a<gt_a_lo || a>gt_a_hi  ||  b<gt_b_lo || b>gt_b_hi
 *       ||  hp < gt_hp_min  ||  (gt_need_oof && !(flag&2)).
 * An UNMATCHED detection is ignored when
 *       a<dt_a_lo || a>dt_a_hi || b<dt_b_lo || b>dt_b_hi || flag&1.               */
typedef struct ta_range_cfg {
    double gt_a_lo, gt_a_hi, gt_b_lo, gt_b_hi;
    double dt_a_lo, dt_a_hi, dt_b_lo, dt_b_hi;
    int32_t gt_hp_min;      /* INT32_MIN disables the rule */
    int32_t gt_need_oof;
} ta_range_cfg;

typedef struct ta_ctx ta_ctx;

int         ta_abi_version(void);
const char* ta_last_error(void);
int         ta_ctx_create(int device, ta_ctx** out);
int         ta_ctx_destroy(ta_ctx* ctx);
int         ta_ctx_sm_count(const ta_ctx* ctx);
/* cudaMemsetAsync(p, 0, bytes) on `stream` (num_gt is ACCUMULATED by the matchers: zero it first) */
int         ta_zero(ta_ctx* ctx, void* stream, void* p, int64_t bytes);
/* TA_IOU_3D (tiled kernel) counts the track pairs whose summed intersection exceeds their summed
 * union — the reference asserts i <= u there (eval.py:95).  ta_eval_plan_host returns
 * TA_ERR_ASSERT for them; callers of the staged API read (and reset) the counter here, which
 * waits for `stream`. */
int         ta_ctx_take_assert_count(ta_ctx* ctx, void* stream, int32_t* count);

/* Spatio-temporal IoU of every (predicted track, GT track) pair of every (video, category)
 * group.  Replaces TaoEval.compute_iou + compute_track_box_iou / compute_avg_track_iou /
 * compute_imagenetvid_iou + bb_intersect_union (tao_amodal/evaluation/tao_amodal/
 * eval.py:306-335, :51-117, :15-48).
 * Tracks of group g are [grp_dt_off[g], grp_dt_off[g+1]) (descending score) and
 * [grp_gt_off[g], grp_gt_off[g+1]); boxes of track t are [trk_off[t], trk_off[t+1]) sorted
 * by `slot` (dense frame index inside the video, one box per slot).  Writes the row-major
 * [D,G] matrix of group g at iou_out + iou_off[g].  n_slots_max bounds slot+1.       */
int ta_track_iou(ta_ctx* ctx, void* stream, int mode, int64_t n_groups,
                 const int64_t* grp_dt_off, const int64_t* grp_gt_off,
                 const int64_t* dt_trk_off, const double* dt_box, const int32_t* dt_slot,
                 const int64_t* gt_trk_off, const double* gt_box, const int32_t* gt_slot,
                 int32_t n_slots_max, const int64_t* iou_off, double* iou_out);

/* Per-(image, category) box IoU.  Replaces LVISEval.compute_iou -> pycocotools.mask.iou
 * -> bbIou (lvis_amodal/eval.py:168-192; maskApi.c:109-120 in-tree copy), iscrowd = 0.
 * grp_list (optional, n_list entries) restricts the call to those groups.            */
int ta_box_iou(ta_ctx* ctx, void* stream, int64_t n_groups,
               const int32_t* grp_list, int64_t n_list,
               const int64_t* grp_dt_off, const int64_t* grp_gt_off,
               const double* dt_box, const double* gt_box,
               const int64_t* iou_off, double* iou_out);

/* Per-(image, category) MASK IoU, iou_type = "segm".  Replaces LVISEval.compute_iou ->
 * pycocotools.mask.iou -> rleIou (lvis_amodal/eval.py:168-192; maskApi.c:78-96 in-tree copy),
 * iscrowd = 0.  Entity e's mask is the column-major run-length list
 * counts[rle_off[e] .. rle_off[e+1]) (zeros first) on an hw[e] = (height, width) canvas, with
 * bbox[e] = [x, y, w, h] as rleToBbox derives it from the runs (maskApi.c:133-151; pairs whose
 * boxes do not overlap are 0 without a walk, :80-82; masks of different size give -1, :85).
 * The run lists come from the host codec of include/ta_mask.h.  Output layout, grp_list as in
 * ta_box_iou; the matrices feed ta_match_greedy.                                        */
int ta_rle_iou(ta_ctx* ctx, void* stream, int64_t n_groups,
               const int32_t* grp_list, int64_t n_list,
               const int64_t* grp_dt_off, const int64_t* grp_gt_off,
               const int64_t* dt_rle_off, const uint32_t* dt_counts,
               const uint32_t* dt_hw, const double* dt_bbox,
               const int64_t* gt_rle_off, const uint32_t* gt_counts,
               const uint32_t* gt_hw, const double* gt_bbox,
               const int64_t* iou_off, double* iou_out);

/* COCO-style sequential greedy assignment for every group x range cfg x IoU threshold.
 * Replaces TaoEval.evaluate_vid (eval.py:337-457) and LVISEval.evaluate_img
 * (lvis_amodal/eval.py:194-303).  The reference's id tests are carried by flag bits so the
 * kernels never read 64-bit ids:
 *   dt_flag bit0  the detection's category is not exhaustively annotated in its video/image
 *                 (eval.py:437-439 / lvis :284-286)
 *           bit1  the detection LOCKS the GT it matches: the reference marks a GT taken with
 *                 `gt_m > 0` on the stored detection id (eval.py:407, lvis :248), so only
 *                 ids > 0 lock
 *   gt_flag bit0  "ignore" key of the GT, bit1 out_of_frame, bit2 the GT's id EQUALS the
 *                 evaluator's "unmatched" value (-1 TaoEval eval.py:390, 0 LVISEval :239):
 *                 a detection matched to it still counts as unmatched (eval.py:527-528)
 * grp_list (optional, n_list entries) restricts the call to those groups; NULL = all.
 * g_max bounds the number of GT entities of any processed group.
 * Outputs:
 *   dt_tpfp  uint32 [n_dt][n_cfg]  bit t = TP at threshold t, bit 16+t = FP (neither: ignored)
 *   num_gt   int32  [n_cat][n_cfg] non-ignored GT count, ACCUMULATED (caller zeroes)
 *   dt_match_gt (optional, may be NULL) int32 [n_cfg][n_thr][n_dt] matched GT position
 *            inside its group (original order) or -1
 *   gt_ignore_out (optional) uint8 [n_cfg][n_gt]                                       */
int ta_match_greedy(ta_ctx* ctx, void* stream, int64_t n_groups,
                    const int32_t* grp_list, int64_t n_list,
                    const int64_t* grp_dt_off, const int64_t* grp_gt_off,
                    const int32_t* grp_cat, const int64_t* iou_off, const double* iou,
                    int32_t n_thr, const double* iou_thrs,
                    int32_t n_cfg, const ta_range_cfg* cfgs,
                    int64_t n_dt, const double* dt_attr_a, const double* dt_attr_b,
                    const uint8_t* dt_flag,
                    int64_t n_gt, const double* gt_attr_a, const double* gt_attr_b,
                    const int32_t* gt_hp, const uint8_t* gt_flag, int32_t g_max,
                    uint32_t* dt_tpfp, int32_t* num_gt,
                    int32_t* dt_match_gt, uint8_t* gt_ignore_out);

/* Fused frame path: box IoU + greedy assignment of every (image, category) group —
 * LVISEval.compute_iou + evaluate_img (lvis_amodal/eval.py:168-303).  Detection area
 * (the unmatched-ignore test of :281-283) is w*h of the box, as lvis_amodal/results.py:56
 * defines it.  gt_attr_a is the GT visibility.
 *
 * Evaluation route (no per-cell outputs requested): the (category, image)-sorted detections
 * are cut into warp tasks at group boundaries by a SCHEDULE (ta_frame_sched_build, once per
 * plan).  It holds everything that does not depend on the boxes: the task table and the
 * per-detection descriptors (from the CSR offsets and dt_flag) and the GT side of the range
 * cfgs — per-GT ignore words and the non-ignored GT counts, from gt_attr_a / gt_flag / cfgs —
 * so a call with a schedule makes no pass over the GT attributes; the cfgs / GT attributes
 * given to ta_frame_eval must then be the ones the schedule was built with.  sched = NULL
 * derives all of it into context scratch on every call.  Each task's GT boxes are staged in shared memory by one
 * bulk-async copy and every detection is evaluated by one lane; groups in which a detection
 * reaches the lowest threshold with several GTs are redone by the general matcher.
 * Groups with GT and more than ta_frame_eval_max_gt() GT boxes, more than
 * ta_frame_eval_max_dt() detections or more than ta_frame_eval_max_pairs() box pairs must
 * be listed in big_list: they are routed through ta_box_iou + ta_match_greedy using `iou`
 * (sized by iou_off) as their IoU storage.  With write_iou != 0 every group's IoU matrix is
 * also written to `iou` (detail route, warp-per-group kernel).
 *
 * Outputs as in ta_match_greedy, plus the optional COMPACT form: with dt_word != NULL,
 * n_thr + 3 n_cfg <= 31 and no per-cell outputs, the result of detection d is the single word
 *     dt_word[d] = M | A << n_thr | B << (n_thr + n_cfg) | U << (n_thr + 2 n_cfg)
 * (M: thresholds at which d is matched; per cfg c: bit c of A = matched thresholds are TP,
 * of B = matched thresholds are FP, of U = unmatched thresholds are FP), i.e. row entry c is
 * (A_c ? M : 0) | ((B_c ? M : 0) | (U_c ? ~M : 0)) << 16; bit 31 set means "read the row
 * dt_tpfp[d][*]" (groups that went through the general matcher).  dt_tpfp rows of all other
 * detections are then NOT written.  ta_pr_accumulate takes the same pair.               */
int64_t ta_frame_sched_bytes(int64_t n_groups, int64_t n_dt, int64_t n_gt, int32_t n_cat,
                             int32_t n_cfg);
int ta_frame_sched_build(ta_ctx* ctx, void* stream, int64_t n_groups,
                         const int64_t* grp_dt_off, const int64_t* grp_gt_off,
                         const int32_t* grp_cat, int64_t n_dt, const uint8_t* dt_flag,
                         int64_t n_gt, const double* gt_attr_a, const uint8_t* gt_flag,
                         int32_t n_cat, int32_t n_cfg, const ta_range_cfg* cfgs, void* sched);
int ta_frame_eval(ta_ctx* ctx, void* stream, int64_t n_groups,
                  const int64_t* grp_dt_off, const int64_t* grp_gt_off, const int32_t* grp_cat,
                  const double* dt_box, const double* gt_box,
                  int32_t n_thr, const double* iou_thrs, int32_t n_cfg, const ta_range_cfg* cfgs,
                  int64_t n_dt, const uint8_t* dt_flag,
                  int64_t n_gt, const double* gt_attr_a, const uint8_t* gt_flag,
                  int64_t n_big, const int32_t* big_list, int32_t g_max_big,
                  const int64_t* iou_off, double* iou, int32_t write_iou,
                  const void* sched, uint32_t* dt_word,
                  uint32_t* dt_tpfp, int32_t* num_gt,
                  int32_t* dt_match_gt, uint8_t* gt_ignore_out);
int ta_frame_eval_max_gt(void);
int ta_frame_eval_max_dt(void);
int ta_frame_eval_max_pairs(void);

/* Precision / recall accumulation.  Replaces TaoEval.accumulate (eval.py:459-584) and
 * LVISEval.accumulate (lvis_amodal/eval.py:305-426).  acc_perm lists, category by
 * category (cat_dt_off), the detection indices in stable descending-score order.
 * dt_tpfp is the [n_dt][n_cfg] output of the matchers; dt_word (optional, may be NULL) the
 * compact per-detection words of ta_frame_eval, which take precedence over the rows unless
 * their bit 31 is set.
 * Outputs (reference tensor layouts, -1 where the reference leaves -1):
 *   precision f64 [n_thr][n_rec][n_cat][n_cfg], recall f64 [n_thr][n_cat][n_cfg],
 *   tp_cnt / fp_cnt int64 [n_thr][n_cat][n_cfg] (may be NULL).
 * Scratch (chunk counters) lives in the context and grows on demand.                   */
int ta_pr_accumulate(ta_ctx* ctx, void* stream, int32_t n_cat, const int64_t* cat_dt_off,
                     const int32_t* acc_perm, int64_t n_dt, const uint32_t* dt_tpfp,
                     const uint32_t* dt_word, const int32_t* num_gt, int32_t n_thr, int32_t n_cfg,
                     int32_t n_rec, const double* rec_thrs,
                     double* precision, double* recall, int64_t* tp_cnt, int64_t* fp_cnt);

/* Whole evaluation of one prepared plan from HOST buffers: copies the plan to the device,
 * runs IoU -> match -> accumulate on ctx's stream, copies precision / recall / counts
 * back, and returns when they are valid.  This is the call the reference-side
 * TaoEval.run()/LVISEval.run() replacement makes (evaluate + accumulate, eval.py:662-665).
 * Track path when the *_trk_off pointers are non-NULL, frame path (fused) otherwise.   */
#define TA_PLAN_BOX_F32 1
#define TA_PLAN_SLOT_U16 2
#define TA_PLAN_GRP_U16 4
typedef struct ta_plan_host {
    int64_t n_groups, n_dt, n_gt, n_dt_boxes, n_gt_boxes, n_big;
    int32_t n_cat, n_cfg, n_thr, n_rec, n_slots_max, g_max, iou_mode;
    int32_t flags;                            /* TA_PLAN_BOX_F32: dt_box / gt_box point to float
                                                 [N,4] arrays that hold the fp64 coordinates
                                                 exactly (lossless transport: half the PCIe
                                                 bytes); they are widened to fp64 on the device
                                                 before any arithmetic.
                                                 TA_PLAN_GRP_U16: grp_dt_off / grp_gt_off point to
                                                 uint16 COUNTS [n_groups] (detections / GT of
                                                 each group, all < 65536) and grp_cat to uint16
                                                 [n_groups]; the int64 offsets are rebuilt on the
                                                 device by a prefix sum (6 instead of 20 bytes
                                                 per group) */
    const void    *grp_dt_off, *grp_gt_off;   /* int64 [n_groups+1] (see TA_PLAN_GRP_U16) */
    const int64_t *iou_off, *cat_dt_off;
    const void    *grp_cat;                   /* int32 [n_groups] (see TA_PLAN_GRP_U16) */
    const int32_t *acc_perm, *big_list;
    const void    *dt_box, *gt_box;           /* double[N,4], or float[N,4] with TA_PLAN_BOX_F32 */
    const int64_t *dt_trk_off, *gt_trk_off;   /* NULL on the frame path */
    const void    *dt_slot, *gt_slot;         /* int32[N] (uint16[N] with TA_PLAN_SLOT_U16: slots
                                                 below 65536 shipped in half the bytes); NULL on the
                                                 frame path */
    const double  *dt_attr_a, *dt_attr_b, *gt_attr_a, *gt_attr_b;
    const uint8_t *dt_flag, *gt_flag;
    const int32_t *gt_hp;
    const double  *iou_thrs, *rec_thrs;
    const ta_range_cfg* cfgs;
    /* ta_eval_plans_host only: when dt_box_idx is non-NULL this plan's detection boxes are
     * rows of the dt_box array of plan number dt_box_pool of the same call (the track and the
     * frame evaluation of one result file hold the same boxes in two orders): dt_box is
     * ignored, box i of this plan = pool box dt_box_idx[i].  int32 [n_dt_boxes].  A frame-path
     * pool plan may carry n_dt_boxes > n_dt: its rows from n_dt on are boxes that only the
     * sharing plan uses.                                                                    */
    const int32_t *dt_box_idx;
    int32_t dt_box_pool;
    int32_t reserved_;
} ta_plan_host;

int ta_eval_plan_host(ta_ctx* ctx, const ta_plan_host* plan,
                      double* precision, double* recall,
                      int64_t* tp_cnt, int64_t* fp_cnt, int32_t* num_gt,
                      int64_t* h2d_bytes, int64_t* d2h_bytes);

/* Several plans of one result set in one call — what tools/eval_on_tao_amodal.py:118-151 does
 * with TaoEval and LVISEval on the same prediction file.  Plan i runs on ctxs[i] (distinct
 * contexts of one device) and writes outs[i]; results are valid on return.  All host-to-device
 * copies go through one upload stream in consumption order (shared box pool; then per plan
 * the matcher's inputs, then accumulate's), so plan i's kernels and download overlap the upload
 * of plan i+1; put the plan with the larger result tensors first.  h2d_bytes / d2h_bytes:
 * optional int64 [n_plans].  tp_cnt / fp_cnt / num_gt of an out may be NULL.              */
typedef struct ta_host_out {
    double  *precision, *recall;
    int64_t *tp_cnt, *fp_cnt;
    int32_t *num_gt;
} ta_host_out;
int ta_eval_plans_host(int32_t n_plans, ta_ctx* const* ctxs, const ta_plan_host* const* plans,
                       const ta_host_out* outs, int64_t* h2d_bytes, int64_t* d2h_bytes);

/* Page-locked host memory for plans and result tensors: with it the copies of
 * ta_eval_plan(s)_host run asynchronously at full PCIe rate (pageable memory works too, staged
 * by the driver).  NULL (and ta_last_error) on failure.                                   */
void* ta_host_alloc(size_t bytes);
void  ta_host_free(void* p);

/* float [n,4] -> double [n,4] on the device: lossless transport of box coordinates that are
 * exactly representable in float (what TA_PLAN_BOX_F32 does inside ta_eval_plan_host), for
 * callers that refresh resident plans from host memory. */
int ta_widen_boxes(ta_ctx* ctx, void* stream, int64_t n, const float* src, double* dst);

/* The device halves of the other compact transport forms of ta_plan_host, for the same callers:
 * uint16 -> int32 (TA_PLAN_SLOT_U16, the category column of TA_PLAN_GRP_U16); int64 offsets
 * [n+1] from uint16 counts [n] (TA_PLAN_GRP_U16; scratch comes from the context); dst[i] =
 * pool[idx[i]] for [.,4] double boxes (dt_box_idx).                                          */
int ta_widen_u16(ta_ctx* ctx, void* stream, int64_t n, const uint16_t* src, int32_t* dst);
int ta_offsets_from_counts(ta_ctx* ctx, void* stream, int64_t n, const uint16_t* counts,
                           int64_t* off);
int ta_gather_boxes(ta_ctx* ctx, void* stream, int64_t n, const double* pool, const int32_t* idx,
                    double* dst);

/* ---- multi-GPU exchange (one process per GPU; NCCL over NVLink / NVSwitch) ---------------
 * Videos (and their images) shard across ranks: ta_track_iou / ta_match_greedy / ta_frame_eval
 * run on each rank's own groups with no communication.  accumulate, however, orders all
 * detections of a category by score across ALL videos (tao_amodal/evaluation/tao_amodal/
 * eval.py:498-518, lvis_amodal/eval.py:340-361), so before ta_pr_accumulate one record per
 * detection (its compact word, or its full TP/FP row) travels to the rank that owns the
 * category, and the non-ignored GT counts are summed.  The reference is single-process; these
 * entry points are what a multi-GPU maintainer-side driver calls between evaluate and
 * accumulate (tao_amodal_b200/parallel.py is that driver here).
 *
 *   ta_exchange_unique_id   rank 0 draws an NCCL id (128 bytes) and hands it to the other ranks
 *                           by any host channel; every rank then calls ta_exchange_create.
 *   ta_exchange_gather      out[i] = src[index[i]]   records of `words` uint32 (send-side pack:
 *                           index = local records grouped by destination rank)
 *   ta_exchange_alltoallv   byte ranges send[send_off[r]..send_off[r+1]) -> rank r, received
 *                           into recv[recv_off[r]..); offsets are HOST arrays of world+1 entries;
 *                           one grouped ncclSend/ncclRecv, asynchronous on `stream`
 *   ta_exchange_scatter     dst[index[i]] = src[i]   (received full rows -> dense row table)
 *   ta_exchange_allreduce_sum   in-place sum over ranks; dtype 0 = int32, 1 = int64, 2 = float64
 *   ta_exchange_group_begin / _end   calls issued in between are fused into ONE NCCL launch
 * All pointers except the offset arrays are device pointers.  NCCL is loaded at run time
 * (libnccl.so.2; a copy already in the process is reused): TA_ERR_NCCL when absent or failing. */
typedef struct ta_exchange ta_exchange;
int ta_exchange_unique_id(void* id, int32_t id_bytes);
int ta_exchange_create(ta_ctx* ctx, int32_t rank, int32_t world, const void* id, ta_exchange** out);
int ta_exchange_destroy(ta_exchange* x);
int ta_exchange_rank(const ta_exchange* x);
int ta_exchange_world(const ta_exchange* x);

/* ---- peer windows: the exchange with the library's own kernel over NVLink peer memory --------
 * A window is a device buffer of this rank that every other rank has mapped (CUDA IPC; one
 * process per GPU on one NVLink / NVSwitch node).  Per evaluation and plan:
 *     ta_peer_window_acquire    (stream) wait until every peer has read the previous contents
 *     ta_peer_window_put ...    (stream) records -> window: TP/FP words or rows in local detection
 *                               order (an owner's share is one contiguous slice), GT counts
 *     ta_peer_window_exchange   (stream) ONE kernel: publish the window, wait for each peer's
 *                               flag, pull this owner's slices from the peers' windows with 16-byte
 *                               loads, sum the GT counts over all ranks, release the peers
 * followed by ta_pr_accumulate on what was pulled — the same data movement as
 * ta_exchange_alltoallv + ta_exchange_allreduce_sum, without an NCCL launch on the step path.
 * What to pull is fixed per plan (ta_peer_window_set_plan): copies {peer, byte offset in that
 * peer's window, bytes, local destination} (offsets / sizes multiples of 4) and sums {byte offset,
 * count, local int32 destination}: dst[i] = sum over all ranks of the int32 at off + 4 i of the
 * rank's window.  create / destroy / set_plan are collective or synchronising calls outside
 * the step; NCCL (the communicator of `x`) only carries the IPC handles at creation.
 * A peer that never arrives makes the waits give up after about a minute instead of hanging
 * the GPU; ta_peer_window_check reports that.                                             */
typedef struct ta_peer_window ta_peer_window;
typedef struct ta_peer_copy { int32_t peer, reserved_; int64_t src_off, bytes; void* dst; } ta_peer_copy;
typedef struct ta_peer_sum { int64_t off, count; int32_t* dst; } ta_peer_sum;
int   ta_peer_window_create(ta_exchange* x, int64_t bytes, ta_peer_window** out);
int   ta_peer_window_destroy(ta_peer_window* w);
void* ta_peer_window_ptr(ta_peer_window* w);
int   ta_peer_window_set_plan(ta_peer_window* w, int32_t n_copy, const ta_peer_copy* copies,
                              int32_t n_sum, const ta_peer_sum* sums);
int   ta_peer_window_acquire(ta_peer_window* w, void* stream);
int   ta_peer_window_put(ta_peer_window* w, void* stream, int64_t off, const void* src, int64_t bytes);
int   ta_peer_window_exchange(ta_peer_window* w, void* stream);
int   ta_peer_window_check(ta_peer_window* w, void* stream, int32_t* timed_out);
int ta_exchange_gather(ta_ctx* ctx, void* stream, int64_t n, int32_t words, const int32_t* index,
                       const uint32_t* src, uint32_t* out);
int ta_exchange_scatter(ta_ctx* ctx, void* stream, int64_t n, int32_t words, const int32_t* index,
                        const uint32_t* src, uint32_t* dst);
int ta_exchange_alltoallv(ta_exchange* x, void* stream, const void* send, const int64_t* send_off,
                          void* recv, const int64_t* recv_off);
int ta_exchange_allreduce_sum(ta_exchange* x, void* stream, void* buf, int64_t count, int32_t dtype);
int ta_exchange_group_begin(ta_exchange* x);
int ta_exchange_group_end(ta_exchange* x);

/* Optional per-kernel timing for benchmarks.  While enabled the context records a CUDA event
 * after every kernel launch (and a marker at every API entry); ta_ctx_timing_read waits for
 * them, aggregates the intervals by kernel name and resets the log.  names receives the
 * distinct kernel names separated by '\n'; total_ms[i] / launches[i] belong to the i-th name.
 * Returns the number of distinct names (<= cap) or a negative ta_status.                */
int ta_ctx_timing(ta_ctx* ctx, int enable);
int ta_ctx_timing_read(ta_ctx* ctx, char* names, int names_cap, double* total_ms, int* launches,
                       int cap);

/* Diagnostics: number of groups the most recent evaluation call of ta_match_greedy /
 * ta_frame_eval on this context handed to the general matcher (-1: none yet). */
int ta_ctx_debug_list_count(ta_ctx* ctx, void* stream, int32_t* count);

/* Number of kernel launches issued through this context since creation. */
int64_t ta_ctx_launch_count(const ta_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* TA_EVAL_H */
