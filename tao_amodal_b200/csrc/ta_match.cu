// ta_match.cu — greedy assignment kernels (sm_100a).
//
//   k_match_greedy   generic: one CTA per group, IoU matrix read from global memory, any
//                    group size (taken-bitmaps in shared memory)
//   k_frame_eval     fused frame path: one WARP per (image, category) group; GT boxes staged
//                    in shared memory, IoU tile computed into shared memory and consumed by
//                    the (range cfg x threshold) matcher lanes without touching HBM
//
// Both implement the reference loop of tao_amodal/evaluation/tao_amodal/eval.py:396-443 /
// lvis_amodal/eval.py:244-290 (see ta_match_one in ta_device_fns.cuh for the equivalence of
// the "ignored GTs last + break" walk with a two-class search in original GT order).
#include <limits.h>
#include "ta_internal.h"
#include "ta_device_fns.cuh"

// ------------------------------------------------------------------------------------------
// generic matcher
// ------------------------------------------------------------------------------------------
struct MatchArgs {
    const int32_t* grp_list;
    const int64_t* grp_dt_off;
    const int64_t* grp_gt_off;
    const int32_t* grp_cat;
    const int64_t* iou_off;
    const double* iou;
    int n_thr;
    const double* thrs;
    int n_cfg;
    const ta_range_cfg* cfgs;
    int64_t n_dt, n_gt;
    const double* dt_a;
    const double* dt_b;
    const uint8_t* dt_flag;
    const double* gt_a;
    const double* gt_b;
    const int32_t* gt_hp;
    const uint8_t* gt_flag;
    uint32_t* dt_tpfp;
    int32_t* num_gt;
    int32_t* dt_match_gt;
    uint8_t* gt_ignore_out;
    int cfgs_per_warp;
    // device-side group list (the flat kernels' leftovers): CTAs loop over dev_list[0 .. *dev_count)
    const int32_t* dev_list;
    const int32_t* dev_count;
    int skip_num_gt;              // the GT counts of these groups were taken by k_frame_prep
    int iou_cap;                  // doubles of shared memory behind the taken / ignore tables: the
                                  // group's IoU matrix is staged there when it fits
};

__global__ void k_match_greedy(MatchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nthreads = blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cpw = a.cfgs_per_warp;
    const int cw = lane / a.n_thr;
    const int t = lane - cw * a.n_thr;
    const int cfg = warp * cpw + cw;
    const bool active = (cw < cpw) && (cfg < a.n_cfg);
    const int n_items = a.dev_list ? *a.dev_count : (int)gridDim.x;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int grp = a.dev_list ? a.dev_list[item] : (a.grp_list ? a.grp_list[item] : item);
        const int64_t d0 = a.grp_dt_off[grp], g0 = a.grp_gt_off[grp];
        const int D = (int)(a.grp_dt_off[grp + 1] - d0), G = (int)(a.grp_gt_off[grp + 1] - g0);
        if (D == 0 && G == 0) continue;
        __syncthreads();          // shared state of the previous group fully consumed
        const int words = (G + 31) >> 5;
        // shared layout: taken[words][nthreads] u32, then gt_ig[n_cfg][G] u8
        uint32_t* taken = reinterpret_cast<uint32_t*>(smem_raw);
        uint8_t* gt_ig = smem_raw + (size_t)words * nthreads * sizeof(uint32_t);

        // ---- GT ignore flags for every cfg (eval.py:348-368 / lvis eval.py:201-217)
        for (int idx = threadIdx.x; idx < a.n_cfg * G; idx += nthreads) {
            const int c = idx / G, g = idx - c * G;
            const uint8_t ig = ta_gt_ignored(a.cfgs[c], a.gt_a[g0 + g], a.gt_b ? a.gt_b[g0 + g] : 0.0,
                                             a.gt_hp ? a.gt_hp[g0 + g] : 0, a.gt_flag[g0 + g]);
            gt_ig[idx] = ig;
            if (a.gt_ignore_out) a.gt_ignore_out[(int64_t)c * a.n_gt + g0 + g] = ig;
        }
        for (int w = 0; w < words; ++w) taken[w * nthreads + threadIdx.x] = 0u;
        __syncthreads();
        // ---- number of non-ignored GT per (category, cfg) (eval.py:520-522)
        if (!a.skip_num_gt && threadIdx.x < a.n_cfg && G > 0) {
            int cnt = 0;
            for (int g = 0; g < G; ++g) cnt += gt_ig[threadIdx.x * G + g] == 0;
            if (cnt) atomicAdd(&a.num_gt[(int64_t)a.grp_cat[grp] * a.n_cfg + threadIdx.x], cnt);
        }
        if (D == 0) continue;

        const double thr = active ? a.thrs[t] : 2.0;
        const double* iou = a.iou + a.iou_off[grp];
        if (D * G <= a.iou_cap) {
            // every matcher lane walks the same rows one after the other: stage the matrix once
            // (coalesced) instead of chasing it through L2 row by row
            double* iou_s = reinterpret_cast<double*>(
                smem_raw + (((size_t)words * nthreads * sizeof(uint32_t) + (size_t)a.n_cfg * G + 7) & ~(size_t)7));
            for (int i = threadIdx.x; i < D * G; i += nthreads) iou_s[i] = iou[i];
            __syncthreads();
            iou = iou_s;
        }
        const uint8_t* my_ig = gt_ig + (active ? cfg : 0) * G;
        ta_range_cfg rc;
        if (active) rc = a.cfgs[cfg];
        uint32_t* my_taken = taken + threadIdx.x;

        const int sh = cw * a.n_thr;
        const uint32_t thr_all = (1u << a.n_thr) - 1u;
        for (int d = 0; d < D; ++d) {
            bool tp = false, fp = false;
            if (active) {
                int m = -1;
                if (G > 0) m = ta_match_one(iou + (int64_t)d * G, G, my_ig, my_taken, nthreads, thr);
                const uint8_t dflag = a.dt_flag[d0 + d];
                bool unmatched = true, ig = false;
                if (m >= 0) {
                    if (dflag & 2) my_taken[(m >> 5) * nthreads] |= 1u << (m & 31);   // eval.py:407,428
                    unmatched = (a.gt_flag[g0 + m] & 4) != 0;                         // eval.py:427,443
                    ig = my_ig[m] != 0;                                               // eval.py:425
                }
                if (unmatched && !ig)
                    ig = ta_dt_unmatched_ignored(rc, a.dt_a[d0 + d], a.dt_b ? a.dt_b[d0 + d] : 0.0, dflag);
                tp = !ig && !unmatched;
                fp = !ig && unmatched;
                if (a.dt_match_gt)
                    a.dt_match_gt[((int64_t)cfg * a.n_thr + t) * a.n_dt + d0 + d] = m;
            }
            // the n_thr lanes of a cfg are consecutive: its TP / FP words are slices of two ballots
            const uint32_t bt = __ballot_sync(0xffffffffu, tp);
            const uint32_t bf = __ballot_sync(0xffffffffu, fp);
            if (active && t == 0)
                a.dt_tpfp[(d0 + d) * a.n_cfg + cfg] = ((bt >> sh) & thr_all) | (((bf >> sh) & thr_all) << 16);
        }
    }
}

// ------------------------------------------------------------------------------------------
// fused frame path
//
// One warp per (image, category) group; a warp owns FE_RUN consecutive groups at a time so the
// non-ignored-GT counts of a category are accumulated in registers and flushed with one atomic
// per (run, cfg).  Three routes per group, all bit-equivalent to lvis_amodal/eval.py:244-290:
//   A  no GT:   every detection is unmatched at every threshold.
//   B  "simple" group — no detection reaches the lowest IoU threshold with more than one GT.
//      Then a detection's only possible match g* does not depend on the range cfg (the
//      regular-before-ignored preference only arbitrates between several candidates), so one
//      lane per GT walks the detections in score order with the thresholds packed in a bit
//      mask: matched = ge(d) & ~taken(g*), taken |= ge(d) if d locks.  All cfg x threshold
//      cells of the group come out of one pass.
//   C  general: lane = (cfg, threshold) sequential matcher over the shared-memory IoU tile,
//      with a per-detection shortcut when that detection has at most one candidate.
// ------------------------------------------------------------------------------------------
#define FE_WARPS 4
#define FE_RUN 16             // consecutive groups per warp task
#define FE_MAX_GT 32          // GT boxes of a group handled on chip (taken / ignore masks = 1 word)
#define FE_MAX_DT 128         // detections of a group handled on chip (when it has GT)
#define FE_MAX_PAIRS 512      // IoU tile doubles per warp
#define FE_MAX_CFG 16
#define TA_WORD_FULL 0x80000000u   // compact result word: the full row in dt_tpfp is authoritative

struct FrameRules;
struct FrameArgs {
    int64_t n_groups;
    const int64_t* grp_dt_off;
    const int64_t* grp_gt_off;
    const int32_t* grp_cat;
    const double* dt_box;
    const double* gt_box;
    int n_thr;
    const double* thrs;
    int n_cfg;
    const ta_range_cfg* cfgs;
    int64_t n_dt, n_gt;
    const uint8_t* dt_flag;
    const double* gt_vis;
    const uint8_t* gt_flag;
    const int64_t* iou_off;
    double* iou;
    int write_iou;
    uint32_t* dt_tpfp;
    int32_t* num_gt;
    int32_t* dt_match_gt;
    uint8_t* gt_ignore_out;
    // track path (NULL on the frame path: area = w*h of the box, b = 0, hp = 0)
    const double* dt_a;           // detection attribute a (mean track area)
    const double* dt_b;           // detection attribute b (number of annotations)
    const double* gt_b;
    const int32_t* gt_hp;
    int exclude_big;              // k_frame_prep: leave oversize frame groups to the big_list route
    // flat path
    const int32_t* dt_grp;        // group of every detection
    int32_t* grp_flag;            // per group: already on the complex list
    int32_t* complex_list;        // groups that need the general matcher (route C)
    int32_t* complex_count;
    const FrameRules* rules_g;    // range-test tables built once per call by k_frame_rules
    // streamed flat path (k_frame_flat): per-plan schedule + per-call GT words
    int64_t n_tasks;
    const int64_t* task_dt;       // [n_tasks + 1] first detection of every task
    const int64_t* task_gt;       // [n_tasks + 1] first GT box of every task
    const uint32_t* dt_desc;      // [n_dt] (GT offset inside the task's 64-box block) << 8 | G << 2 | flag
    const uint32_t* gt_word;      // [n_gt] bit c = ignored by cfg c, bit 16 = id equals the unmatched value
    uint32_t* gt_word_out;        // k_gt_words output (same buffer)
    uint32_t* dt_word;            // [n_dt] compact result words (NULL: full rows in dt_tpfp)
};

// per-detection word: bits 0..15 "ignored when unmatched" per cfg, bit 16 locks its GT
// candidate word:     bits 0..15 thresholds reached by the best GT (route B: later the matched
//                     thresholds), bits 16..20 that GT, bits 21..22 min(#candidates, 2)
struct FrameSmem {
    double iou[FE_WARPS][FE_MAX_PAIRS];
    double gtb[FE_WARPS][FE_MAX_GT][4];
    uint32_t dmask[FE_WARPS][FE_MAX_DT];
    uint32_t cand[FE_WARPS][FE_MAX_DT];
    uint32_t gig[FE_WARPS][FE_MAX_CFG];       // per cfg: bit g = GT g ignored
    uint32_t csum[FE_WARPS][FE_MAX_DT];       // route C: per detection, "same outcome in every cfg"
};

// Range tests evaluated once per DISTINCT interval instead of once per cfg: the distinct
// [lo, hi] intervals of the four interval tests of ta_range_cfg (detection a / b, GT a / b), the
// distinct gt_hp_min values and the cfgs that need out_of_frame, each with the bit mask of the
// cfgs using it.  Intervals (-inf, +inf) and disabled hp rules are dropped.  LVISEval: one
// detection interval, five GT visibility intervals; TaoEval: 4 area x 4 duration intervals
// for 20 cfgs.
#define RR_MAX 32
struct FrameRules {
    int n_da, n_db, n_ga, n_gb, n_hp;
    uint32_t g_oof;
    double da_lo[RR_MAX], da_hi[RR_MAX], db_lo[RR_MAX], db_hi[RR_MAX];
    double ga_lo[RR_MAX], ga_hi[RR_MAX], gb_lo[RR_MAX], gb_hi[RR_MAX];
    uint32_t da_mask[RR_MAX], db_mask[RR_MAX], ga_mask[RR_MAX], gb_mask[RR_MAX], hp_mask[RR_MAX];
    int32_t hp_min[RR_MAX];
};

// cfg mask "this detection is ignored when unmatched" (eval.py:432-439, lvis eval.py:281-286)
__device__ __forceinline__ uint32_t fe_dt_unmatched_mask(const FrameRules& r, double a, double b,
                                                         uint8_t fl, uint32_t cfg_all) {
    uint32_t m = (fl & 1) ? cfg_all : 0u;
    for (int k = 0; k < r.n_da; ++k)
        if (a < r.da_lo[k] || a > r.da_hi[k]) m |= r.da_mask[k];
    for (int k = 0; k < r.n_db; ++k)
        if (b < r.db_lo[k] || b > r.db_hi[k]) m |= r.db_mask[k];
    return m;
}
// cfg mask "this GT is ignored" (eval.py:349-368, lvis eval.py:202-217)
__device__ __forceinline__ uint32_t fe_gt_ignore_mask(const FrameRules& r, double a, double b,
                                                      int32_t hp, uint8_t fl, uint32_t cfg_all) {
    uint32_t m = (fl & 1) ? cfg_all : 0u;
    if (!(fl & 2)) m |= r.g_oof;
    for (int k = 0; k < r.n_ga; ++k)
        if (a < r.ga_lo[k] || a > r.ga_hi[k]) m |= r.ga_mask[k];
    for (int k = 0; k < r.n_gb; ++k)
        if (b < r.gb_lo[k] || b > r.gb_hi[k]) m |= r.gb_mask[k];
    for (int k = 0; k < r.n_hp; ++k)
        if (hp < r.hp_min[k]) m |= r.hp_mask[k];
    return m;
}

__device__ __forceinline__ void fe_add_interval(double lo, double hi, int c, int& n, double* los,
                                                double* his, uint32_t* masks) {
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    if (lo == -INF && hi == INF) return;            // never excludes anything
    int k = 0;
    for (; k < n; ++k)
        if (los[k] == lo && his[k] == hi) break;
    if (k == n) { los[k] = lo; his[k] = hi; masks[k] = 0; ++n; }
    masks[k] |= 1u << c;
}

__device__ __forceinline__ void fe_build_rules(const ta_range_cfg* cfg_s, int n_cfg, FrameRules& rules) {
    rules.n_da = rules.n_db = rules.n_ga = rules.n_gb = rules.n_hp = 0;
    rules.g_oof = 0;
    for (int c = 0; c < n_cfg; ++c) {
        const ta_range_cfg& r = cfg_s[c];
        fe_add_interval(r.dt_a_lo, r.dt_a_hi, c, rules.n_da, rules.da_lo, rules.da_hi, rules.da_mask);
        fe_add_interval(r.dt_b_lo, r.dt_b_hi, c, rules.n_db, rules.db_lo, rules.db_hi, rules.db_mask);
        fe_add_interval(r.gt_a_lo, r.gt_a_hi, c, rules.n_ga, rules.ga_lo, rules.ga_hi, rules.ga_mask);
        fe_add_interval(r.gt_b_lo, r.gt_b_hi, c, rules.n_gb, rules.gb_lo, rules.gb_hi, rules.gb_mask);
        if (r.gt_need_oof) rules.g_oof |= 1u << c;
        if (r.gt_hp_min != INT_MIN) {
            int k = 0;
            for (; k < rules.n_hp; ++k) if (rules.hp_min[k] == r.gt_hp_min) break;
            if (k == rules.n_hp) { rules.hp_min[k] = r.gt_hp_min; rules.hp_mask[k] = 0; ++rules.n_hp; }
            rules.hp_mask[k] |= 1u << c;
        }
    }
}

// one thread, once per API call: the tables every other kernel of the call copies into shared
// memory (building them per CTA costs tens of microseconds of serial work with 20 cfgs)
__global__ void k_frame_rules(const ta_range_cfg* cfgs, int n_cfg, FrameRules* out) {
    // the serial table build runs on shared-memory copies (a single thread chasing global
    // memory took ~25 us); the warp copies in and out
    __shared__ ta_range_cfg cfg_s[RR_MAX];
    __shared__ FrameRules rules_s;
    for (int i = threadIdx.x; i < n_cfg && i < RR_MAX; i += blockDim.x) cfg_s[i] = cfgs[i];
    __syncthreads();
    if (threadIdx.x == 0) fe_build_rules(cfg_s, n_cfg < RR_MAX ? n_cfg : RR_MAX, rules_s);
    __syncthreads();
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&rules_s);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out);
    for (int i = threadIdx.x; i < (int)(sizeof(FrameRules) / 4); i += blockDim.x) dst[i] = src[i];
}

__device__ __forceinline__ void fe_setup(const FrameArgs& a, int n_thr, int n_cfg,
                                         ta_range_cfg* cfg_s, double* thr_s, FrameRules& rules) {
    for (int i = threadIdx.x; i < n_cfg; i += blockDim.x) cfg_s[i] = a.cfgs[i];
    if (threadIdx.x < n_thr) {
        const double th = a.thrs[threadIdx.x];
        thr_s[threadIdx.x] = (th < 1.0 - 1e-10) ? th : 1.0 - 1e-10;
    }
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.rules_g);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&rules);
    for (int i = threadIdx.x; i < (int)(sizeof(FrameRules) / 4); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

// DETAIL = the optional per-cell outputs (IoU matrices, matched GT, GT ignore flags); the
// evaluation path instantiates the lean variant
// NT / NC > 0 fix the number of thresholds / range cfgs at compile time (the evaluators'
// defaults, 10 x 6) so the per-cfg and per-threshold loops unroll.
// LIST: process the groups of a.complex_list one per task (the flat kernel's leftovers) and
// leave num_gt alone (k_frame_num_gt has counted them already).
template <bool DETAIL, int NT, int NC, bool LIST>
__global__ void __launch_bounds__(FE_WARPS * 32, 8)
k_frame_eval(FrameArgs a) {
    __shared__ FrameSmem sm;
    __shared__ ta_range_cfg cfg_s[FE_MAX_CFG];
    __shared__ double thr_s[TA_MAX_THRS];
    __shared__ FrameRules rules;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_thr = NT ? NT : a.n_thr, n_cfg = NC ? NC : a.n_cfg;
    fe_setup(a, n_thr, n_cfg, cfg_s, thr_s, rules);
    const uint32_t cfg_all = (n_cfg == 32) ? 0xffffffffu : ((1u << n_cfg) - 1u);
    double thr_min = thr_s[0];
    for (int i = 1; i < n_thr; ++i) thr_min = (thr_s[i] < thr_min) ? thr_s[i] : thr_min;
    const int cpw = 32 / n_thr;
    const int cw = lane / n_thr;
    const int t = lane - cw * n_thr;
    const double thr_c = (cw < cpw) ? thr_s[t] : 2.0;
    const uint32_t thr_all = (1u << n_thr) - 1u;
    double* iou_s = sm.iou[warp];
    uint32_t* dmask_s = sm.dmask[warp];
    uint32_t* cand_s = sm.cand[warp];
    uint32_t* gig_s = sm.gig[warp];

    const int64_t n_tasks = LIST ? (int64_t)*a.complex_count : (a.n_groups + FE_RUN - 1) / FE_RUN;
    const int64_t w0 = (int64_t)blockIdx.x * FE_WARPS + warp;
    const int64_t wstride = (int64_t)gridDim.x * FE_WARPS;
    for (int64_t task = w0; task < n_tasks; task += wstride) {
        const int64_t grp0 = LIST ? (int64_t)a.complex_list[task] : task * FE_RUN;
        const int n_in_task = LIST ? 1 : (int)((grp0 + FE_RUN < a.n_groups) ? FE_RUN : a.n_groups - grp0);
        // group table of the whole task in one coalesced load: lane i holds entry grp0 + i
        int64_t dt_off_r = 0, gt_off_r = 0;
        int cat_r = 0;
        if (lane <= n_in_task) {
            dt_off_r = a.grp_dt_off[grp0 + lane];
            gt_off_r = a.grp_gt_off[grp0 + lane];
        }
        if (lane < n_in_task) cat_r = a.grp_cat[grp0 + lane];
        // ---- route A for the whole task in one flat pass over its contiguous detections:
        // every detection of a group without GT is unmatched at every threshold
        {
            const int64_t d_begin = __shfl_sync(0xffffffffu, dt_off_r, 0);
            const int64_t d_end = __shfl_sync(0xffffffffu, dt_off_r, n_in_task);
            for (int64_t base = d_begin; base < d_end; base += 32) {
                const int64_t dd = base + lane;
                int gi = 0;                   // last group of the task starting at or before dd
#pragma unroll
                for (int step = FE_RUN / 2; step >= 1; step >>= 1) {
                    const int cnd = gi + step;
                    const int64_t v = __shfl_sync(0xffffffffu, dt_off_r, cnd & 31);
                    if (cnd < n_in_task && v <= dd) gi = cnd;
                }
                const int64_t ga = __shfl_sync(0xffffffffu, gt_off_r, gi);
                const int64_t gb = __shfl_sync(0xffffffffu, gt_off_r, gi + 1);
                if (dd < d_end && ga == gb) {
                    const double2 q = *reinterpret_cast<const double2*>(a.dt_box + 4 * dd + 2);
                    const uint32_t m = fe_dt_unmatched_mask(rules, q.x * q.y, 0.0, a.dt_flag[dd], cfg_all);
                    uint32_t* o = a.dt_tpfp + dd * n_cfg;
                    for (int c = 0; c < n_cfg; ++c) o[c] = ((m >> c) & 1u) ? 0u : (thr_all << 16);
                    if (DETAIL && a.dt_match_gt)
                        for (int ct = 0; ct < n_cfg * n_thr; ++ct)
                            a.dt_match_gt[(int64_t)ct * a.n_dt + dd] = -1;
                }
            }
        }
        int acc = 0, acc_cat = -1;          // lane c < n_cfg: non-ignored GT of (acc_cat, cfg c)
        for (int gi = 0; gi < n_in_task; ++gi) {
            const int64_t grp = grp0 + gi;
            const int64_t g0 = __shfl_sync(0xffffffffu, gt_off_r, gi);
            const int G = (int)(__shfl_sync(0xffffffffu, gt_off_r, gi + 1) - g0);
            if (G == 0) continue;           // route A (or an empty group)
            const int64_t d0 = __shfl_sync(0xffffffffu, dt_off_r, gi);
            const int D = (int)(__shfl_sync(0xffffffffu, dt_off_r, gi + 1) - d0);
            const int cat = __shfl_sync(0xffffffffu, cat_r, gi);
            if (G > FE_MAX_GT || D > FE_MAX_DT || D * G > FE_MAX_PAIRS) continue;   // big_list route
            __syncwarp();
            // ---- GT side: boxes to shared memory, ignore masks per cfg, non-ignored counts.
            // All global loads of the group (GT lane data, first detection per lane) are issued
            // before anything consumes them.
            double vis = 0.0;
            uint8_t gfl = 0;
            double2 gp = make_double2(0, 0), gq = make_double2(0, 0);
            if (lane < G) {
                gp = *reinterpret_cast<const double2*>(a.gt_box + 4 * (g0 + lane));
                gq = *reinterpret_cast<const double2*>(a.gt_box + 4 * (g0 + lane) + 2);
                vis = a.gt_vis[g0 + lane];
                gfl = a.gt_flag[g0 + lane];
            }
            double2 dp = make_double2(0, 0), dq = make_double2(0, 0);
            uint8_t dfl = 0;
            if (lane < D) {
                dp = *reinterpret_cast<const double2*>(a.dt_box + 4 * (d0 + lane));
                dq = *reinterpret_cast<const double2*>(a.dt_box + 4 * (d0 + lane) + 2);
                dfl = a.dt_flag[d0 + lane];
            }
            if (lane < G) {
                sm.gtb[warp][lane][0] = gp.x; sm.gtb[warp][lane][1] = gp.y;
                sm.gtb[warp][lane][2] = gq.x; sm.gtb[warp][lane][3] = gq.y;
            }
            const uint32_t gsent = __ballot_sync(0xffffffffu, (gfl & 4) != 0);   // id == "unmatched" value
            const uint32_t gmask = (lane < G) ? fe_gt_ignore_mask(rules, vis, 0.0, 0, gfl, cfg_all) : 0u;
            uint32_t my_gig = 0;
            for (int c = 0; c < n_cfg; ++c) {
                const uint32_t m = __ballot_sync(0xffffffffu, (gmask >> c) & 1u);
                if (lane == c) my_gig = m;
                if (DETAIL && a.gt_ignore_out && lane < G)
                    a.gt_ignore_out[(int64_t)c * a.n_gt + g0 + lane] = (gmask >> c) & 1u;
            }
            if (!LIST && cat != acc_cat) {
                if (acc > 0) atomicAdd(&a.num_gt[(int64_t)acc_cat * n_cfg + lane], acc);
                acc = 0;
                acc_cat = cat;
            }
            if (lane < n_cfg) {
                gig_s[lane] = my_gig;
                if (!LIST) acc += G - __popc(my_gig);
            }
            __syncwarp();
            // ---- detection side, lane = detection: IoU row (maskApi.c:109-120) into the tile
            // (layout [g][d]: conflict-free stores, broadcast reads), unmatched-ignore mask, lock
            // bit and candidate summary straight from registers
            double* iou_g = (DETAIL && a.write_iou) ? a.iou + a.iou_off[grp] : nullptr;
            bool multi = false;
            for (int d = lane; d < D; d += 32) {
                if (d >= 32) {
                    dp = *reinterpret_cast<const double2*>(a.dt_box + 4 * (d0 + d));
                    dq = *reinterpret_cast<const double2*>(a.dt_box + 4 * (d0 + d) + 2);
                    dfl = a.dt_flag[d0 + d];
                }
                dmask_s[d] = fe_dt_unmatched_mask(rules, dq.x * dq.y, 0.0, dfl, cfg_all) |
                             ((dfl & 2) ? (1u << 16) : 0u);
                int cnt = 0, gs = 0;
                double vs = 0.0;
                for (int g = 0; g < G; ++g) {
                    const double* gb = sm.gtb[warp][g];
                    const double v = ta_bb_iou(dp.x, dp.y, dq.x, dq.y, gb[0], gb[1], gb[2], gb[3]);
                    iou_s[g * D + d] = v;
                    if (DETAIL && iou_g) iou_g[d * G + g] = v;
                    if (!(v < thr_min)) { ++cnt; gs = g; vs = v; }
                }
                uint32_t ge = 0;
                if (cnt == 1)
                    for (int k = 0; k < n_thr; ++k) ge |= (!(vs < thr_s[k])) ? (1u << k) : 0u;
                cand_s[d] = ge | ((uint32_t)gs << 16) | ((uint32_t)(cnt > 2 ? 2 : cnt) << 21);
                multi |= cnt > 1;
            }
            const bool general = __any_sync(0xffffffffu, multi);
            __syncwarp();
            if (!general) {
                // ---- route B
                if (D <= 32) {
                    // lane = detection.  Threshold k of GT g is taken before detection d iff an
                    // earlier locking detection with the same single candidate g reaches k:
                    // one ballot per threshold, intersected with "same GT" and "earlier lane"
                    const uint32_t cd = (lane < D) ? cand_s[lane] : 0u;
                    const bool single = ((cd >> 21) & 3u) == 1u;
                    const uint32_t gs = (cd >> 16) & 31u, ge = cd & 0xffffu;
                    const bool locks = single && ((lane < D ? dmask_s[lane] : 0u) & (1u << 16));
                    uint32_t same = 0;
                    for (int g = 0; g < G; ++g) {
                        const uint32_t bg = __ballot_sync(0xffffffffu, single && gs == (uint32_t)g);
                        if (gs == (uint32_t)g) same = bg;
                    }
                    const uint32_t earlier = same & ((1u << lane) - 1u);
                    uint32_t taken = 0;
                    for (int k = 0; k < n_thr; ++k) {
                        const uint32_t bk = __ballot_sync(0xffffffffu, locks && ((ge >> k) & 1u));
                        if (bk & earlier) taken |= 1u << k;
                    }
                    if (lane < D && single) cand_s[lane] = (cd & ~0xffffu) | (ge & ~taken);
                } else if (lane < G) {
                    // lane g walks the detections whose only candidate is g
                    uint32_t taken = 0;
                    for (int d = 0; d < D; ++d) {
                        const uint32_t cd = cand_s[d];
                        if (((cd >> 21) & 3u) == 1u && ((cd >> 16) & 31u) == (uint32_t)lane) {
                            const uint32_t ge = cd & 0xffffu;
                            cand_s[d] = (cd & ~0xffffu) | (ge & ~taken);
                            if (dmask_s[d] & (1u << 16)) taken |= ge;     // eval.py:248, :270
                        }
                    }
                }
                __syncwarp();
                for (int d = lane; d < D; d += 32) {
                    const uint32_t cd = cand_s[d], dm = dmask_s[d];
                    const uint32_t M = (((cd >> 21) & 3u) == 1u) ? (cd & 0xffffu) : 0u;
                    const int gs = (cd >> 16) & 31;
                    const bool sent = (gsent >> gs) & 1u;
                    uint32_t* o = a.dt_tpfp + (d0 + d) * n_cfg;
                    uint32_t nig = 0;                  // cfgs in which the matched GT is not ignored
                    for (int c = 0; c < n_cfg; ++c) {
                        const bool gi2 = (gig_s[c] >> gs) & 1u;
                        const bool dc = (dm >> c) & 1u;
                        const uint32_t tp = (!sent && !gi2) ? M : 0u;
                        const uint32_t fp = ((sent && !gi2 && !dc) ? M : 0u) | (dc ? 0u : (thr_all & ~M));
                        o[c] = tp | (fp << 16);
                        nig |= gi2 ? 0u : (1u << c);
                    }
                    if (LIST && a.dt_word) {
                        // a group the flat kernel passed on that turns out to be simple: its
                        // detections keep compact words (layout: k_frame_flat)
                        const uint32_t ndc = ~dm & cfg_all;
                        const uint32_t A = (M && !sent) ? nig : 0u;
                        const uint32_t B = (M && sent) ? (nig & ndc) : 0u;
                        a.dt_word[d0 + d] = M | (A << n_thr) | (B << (n_thr + n_cfg)) | (ndc << (n_thr + 2 * n_cfg));
                    }
                    if (DETAIL && a.dt_match_gt)
                        for (int c = 0; c < n_cfg; ++c)
                            for (int k = 0; k < n_thr; ++k)
                                a.dt_match_gt[((int64_t)c * n_thr + k) * a.n_dt + d0 + d] =
                                    ((M >> k) & 1u) ? gs : -1;
                }
                continue;
            }
            // ---- route C: lane = (cfg within round, threshold)
            // With compact result words (LIST): a detection keeps a compact word when it is matched
            // at the same thresholds, to the same GT, under every range cfg — true for most
            // detections even of these groups — and points to its full row otherwise.
            const bool want_word = LIST && a.dt_word != nullptr;
            uint32_t* csum_s = sm.csum[warp];
            const uint32_t gall = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
            for (int c0 = 0; c0 < n_cfg; c0 += cpw) {
                const int cfg = c0 + cw;
                const bool active = (cw < cpw) && (cfg < n_cfg);
                const uint32_t gig = active ? gig_s[cfg] : 0u;
                const int sh = cw * n_thr;
                uint32_t taken = 0u;
                for (int d = 0; d < D; ++d) {
                    const uint32_t cd = cand_s[d], dm = dmask_s[d];
                    const uint32_t cnt = (cd >> 21) & 3u;
                    int m = -1;
                    if (cnt == 1u) {
                        const int gs = (cd >> 16) & 31;
                        if (!(iou_s[gs * D + d] < thr_c) && !((taken >> gs) & 1u)) m = gs;
                    } else if (cnt > 1u) {
                        const uint32_t free0 = ~taken & ~gig & gall;   // regular GTs still free
                        const uint32_t free1 = ~taken & gig & gall;    // ignored GTs still free
                        double best0 = thr_c, best1 = thr_c;
                        int m0 = -1, m1 = -1;
                        for (int g = 0; g < G; ++g) {
                            const double v = iou_s[g * D + d];
                            if (((free0 >> g) & 1u) && !(v < best0)) { best0 = v; m0 = g; }
                            if (((free1 >> g) & 1u) && !(v < best1)) { best1 = v; m1 = g; }
                        }
                        m = (m0 >= 0) ? m0 : m1;
                    }
                    bool unmatched = true, ig = false;
                    if (m >= 0) {
                        if (dm & (1u << 16)) taken |= 1u << m;
                        unmatched = (gsent >> m) & 1u;
                        ig = (gig >> m) & 1u;
                    }
                    if (unmatched && !ig) ig = (dm >> cfg) & 1u;
                    const bool cnts = active && !ig;
                    const uint32_t bt = __ballot_sync(0xffffffffu, cnts && !unmatched);
                    const uint32_t bf = __ballot_sync(0xffffffffu, cnts && unmatched);
                    if (active && t == 0)
                        a.dt_tpfp[(d0 + d) * n_cfg + cfg] =
                            ((bt >> sh) & thr_all) | (((bf >> sh) & thr_all) << 16);
                    if (DETAIL && a.dt_match_gt && active)
                        a.dt_match_gt[((int64_t)cfg * n_thr + t) * a.n_dt + d0 + d] = m;
                    if (want_word) {
                        const uint32_t bm = __ballot_sync(0xffffffffu, active && m >= 0);
                        const int g1 = bm ? __shfl_sync(0xffffffffu, m, __ffs(bm) - 1) : -1;
                        const bool same_g = __ballot_sync(0xffffffffu, active && m >= 0 && m != g1) == 0u;
                        const uint32_t M0 = bm & thr_all;
                        uint32_t rep = 0;
                        for (int q = 0; q < cpw && c0 + q < n_cfg; ++q) rep |= M0 << (q * n_thr);
                        const bool ok = same_g && bm == rep;
                        if (lane == 0) {
                            const uint32_t mine = M0 | ((uint32_t)(g1 + 1) << 16);
                            if (c0 == 0) csum_s[d] = ok ? mine : 0xffffffffu;
                            else if (csum_s[d] != 0xffffffffu &&
                                     (!ok || (csum_s[d] & 0xffffu) != M0 || (M0 && csum_s[d] != mine)))
                                csum_s[d] = 0xffffffffu;
                        }
                    }
                }
            }
            if (want_word) {
                __syncwarp();
                for (int d = lane; d < D; d += 32) {
                    const uint32_t cs = csum_s[d];
                    uint32_t word = TA_WORD_FULL;
                    if (cs != 0xffffffffu) {
                        const uint32_t M = cs & 0xffffu, dm = dmask_s[d];
                        const int gs = M ? (int)(cs >> 16) - 1 : 0;
                        const bool sent = (gsent >> gs) & 1u;
                        uint32_t nig = 0;
                        for (int c = 0; c < n_cfg; ++c) nig |= ((gig_s[c] >> gs) & 1u) ? 0u : (1u << c);
                        const uint32_t ndc = ~dm & cfg_all;
                        const uint32_t A = (M && !sent) ? nig : 0u;
                        const uint32_t B = (M && sent) ? (nig & ndc) : 0u;
                        word = M | (A << n_thr) | (B << (n_thr + n_cfg)) | (ndc << (n_thr + 2 * n_cfg));
                    }
                    a.dt_word[d0 + d] = word;
                }
            }
        }
        if (!LIST && acc > 0) atomicAdd(&a.num_gt[(int64_t)acc_cat * n_cfg + lane], acc);
    }
}

// ------------------------------------------------------------------------------------------
// flat frame path: one LANE per detection
//
// A warp owns 32 consecutive detections of the (category, image)-sorted array, whatever groups
// they belong to, so lanes stay busy however small the groups are.  Each lane computes the IoUs
// of its detection with the GTs of its group straight from global memory (neighbouring lanes
// share the GT boxes through L1) and summarises them as in route B of k_frame_eval: number of
// GTs reaching the lowest threshold, the last such GT g*, the mask ge of thresholds it reaches.
// For a detection with a single candidate the greedy loop of lvis eval.py:244-290 reduces to
//     matched(d) = ge(d) & ~OR{ ge(d') : d' earlier in the group, same g*, d' locks }
// "same group and same g*" is a __match_any_sync on (group, g*), "reaches threshold k and locks"
// one ballot per threshold.  The detections of the window's first group that lie before the
// window are replayed (candidates only) to seed the taken masks, so windows are independent.
// Groups in which some detection has several candidates are appended to complex_list and
// redone by k_frame_eval<LIST> (general matcher), which overwrites their rows.
// ------------------------------------------------------------------------------------------
struct FlatCand {
    int cnt, gs;
    uint32_t ge;
};

__device__ __forceinline__ FlatCand fe_candidate(const double* __restrict__ gt_box, int64_t g0, int G,
                                                 double2 dp, double2 dq, double thr_min,
                                                 const double* thr_s, int n_thr) {
    FlatCand c{0, 0, 0u};
    double vs = 0.0;
    // two GT boxes per iteration: both loads are in flight before either IoU is computed
    int g = 0;
    for (; g + 1 < G; g += 2) {
        const double2 p0 = *reinterpret_cast<const double2*>(gt_box + 4 * (g0 + g));
        const double2 q0 = *reinterpret_cast<const double2*>(gt_box + 4 * (g0 + g) + 2);
        const double2 p1 = *reinterpret_cast<const double2*>(gt_box + 4 * (g0 + g) + 4);
        const double2 q1 = *reinterpret_cast<const double2*>(gt_box + 4 * (g0 + g) + 6);
        const double v0 = ta_bb_iou(dp.x, dp.y, dq.x, dq.y, p0.x, p0.y, q0.x, q0.y);
        const double v1 = ta_bb_iou(dp.x, dp.y, dq.x, dq.y, p1.x, p1.y, q1.x, q1.y);
        if (!(v0 < thr_min)) { ++c.cnt; c.gs = g; vs = v0; }
        if (!(v1 < thr_min)) { ++c.cnt; c.gs = g + 1; vs = v1; }
    }
    if (g < G) {
        const double2 gp = *reinterpret_cast<const double2*>(gt_box + 4 * (g0 + g));
        const double2 gq = *reinterpret_cast<const double2*>(gt_box + 4 * (g0 + g) + 2);
        const double v = ta_bb_iou(dp.x, dp.y, dq.x, dq.y, gp.x, gp.y, gq.x, gq.y);
        if (!(v < thr_min)) { ++c.cnt; c.gs = g; vs = v; }
    }
    if (c.cnt == 1)
        for (int k = 0; k < n_thr; ++k) c.ge |= (!(vs < thr_s[k])) ? (1u << k) : 0u;
    return c;
}

#define FF_WARPS 4

// candidate summary from a row of a precomputed IoU matrix (track path)
__device__ __forceinline__ FlatCand fe_candidate_row(const double* __restrict__ row, int G,
                                                     double thr_min, const double* thr_s, int n_thr) {
    FlatCand c{0, 0, 0u};
    double vs = 0.0;
    for (int g = 0; g < G; ++g) {
        const double v = row[g];
        if (!(v < thr_min)) { ++c.cnt; c.gs = g; vs = v; }
    }
    if (c.cnt == 1)
        for (int k = 0; k < n_thr; ++k) c.ge |= (!(vs < thr_s[k])) ? (1u << k) : 0u;
    return c;
}

// TRACK: detections are tracks; IoUs come from the matrix ta_track_iou wrote (a.iou / a.iou_off),
// attributes from dt_a / dt_b / gt_b / gt_hp, and groups with more than 32 GT go to the list.
template <int NT, int NC, bool TRACK>
__global__ void __launch_bounds__(FF_WARPS * 32, 8)
k_track_flat(FrameArgs a) {
    __shared__ ta_range_cfg cfg_s[RR_MAX];
    __shared__ double thr_s[TA_MAX_THRS];
    __shared__ FrameRules rules;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_thr = NT ? NT : a.n_thr, n_cfg = NC ? NC : a.n_cfg;
    fe_setup(a, n_thr, n_cfg, cfg_s, thr_s, rules);
    const uint32_t cfg_all = (n_cfg == 32) ? 0xffffffffu : ((1u << n_cfg) - 1u);
    double thr_min = thr_s[0];
    for (int i = 1; i < n_thr; ++i) thr_min = (thr_s[i] < thr_min) ? thr_s[i] : thr_min;
    const uint32_t thr_all = (1u << n_thr) - 1u;
    const uint32_t lanes_lt = (1u << lane) - 1u;

    const int64_t n_win = (a.n_dt + 31) / 32;
    const int64_t w0 = (int64_t)blockIdx.x * FF_WARPS + warp;
    const int64_t wstride = (int64_t)gridDim.x * FF_WARPS;
    for (int64_t win = w0; win < n_win; win += wstride) {
        const int64_t base = win * 32;
        const int64_t d = base + lane;
        const bool valid = d < a.n_dt;
        int grp = -1, G = 0;
        int64_t d0 = 0, g0 = 0;
        double2 dp = make_double2(0, 0), dq = make_double2(0, 0);
        uint8_t dfl = 0;
        if (valid) {
            grp = a.dt_grp[d];
            if (!TRACK) {
                dp = *reinterpret_cast<const double2*>(a.dt_box + 4 * d);
                dq = *reinterpret_cast<const double2*>(a.dt_box + 4 * d + 2);
            }
            dfl = a.dt_flag[d];
            d0 = a.grp_dt_off[grp];
            g0 = a.grp_gt_off[grp];
            G = (int)(a.grp_gt_off[grp + 1] - g0);
        }
        // g* has 5 bits: larger groups need the general matcher (frame path: big_list route;
        // track path: appended to the complex list here)
        const bool skip = G > FE_MAX_GT;
        if (TRACK && skip && atomicExch(&a.grp_flag[grp], 1) == 0)
            a.complex_list[atomicAdd(a.complex_count, 1)] = grp;
        // ---- replay: detections of lane 0's group that precede the window seed its taken masks
        const int grp_f = __shfl_sync(0xffffffffu, grp, 0);
        const int G_f = __shfl_sync(0xffffffffu, G, 0);
        const int64_t d0_f = __shfl_sync(0xffffffffu, d0, 0);
        const int64_t g0_f = __shfl_sync(0xffffffffu, g0, 0);
        uint32_t carry = 0;                        // lane g: thresholds of GT g taken before the window
        if (G_f > 0 && G_f <= FE_MAX_GT && d0_f < base) {
            for (int64_t rb = d0_f; rb < base; rb += 32) {
                const int64_t rd = rb + lane;
                uint32_t x = 0;
                int rgs = -1;
                if (rd < base) {
                    FlatCand rc;
                    if (TRACK) {
                        rc = fe_candidate_row(a.iou + a.iou_off[grp_f] + (rd - d0_f) * G_f, G_f,
                                              thr_min, thr_s, n_thr);
                    } else {
                        const double2 rp = *reinterpret_cast<const double2*>(a.dt_box + 4 * rd);
                        const double2 rq = *reinterpret_cast<const double2*>(a.dt_box + 4 * rd + 2);
                        rc = fe_candidate(a.gt_box, g0_f, G_f, rp, rq, thr_min, thr_s, n_thr);
                    }
                    if (rc.cnt == 1 && (a.dt_flag[rd] & 2)) { x = rc.ge; rgs = rc.gs; }
                }
                for (int g = 0; g < G_f; ++g) {
                    const uint32_t m = __reduce_or_sync(0xffffffffu, rgs == g ? x : 0u);
                    if (lane == g) carry |= m;
                }
            }
        }
        // ---- own candidate
        FlatCand c{0, 0, 0u};
        const bool work = valid && G > 0 && !skip;
        if (work) {
            if (TRACK)
                c = fe_candidate_row(a.iou + a.iou_off[grp] + (d - d0) * G, G, thr_min, thr_s, n_thr);
            else
                c = fe_candidate(a.gt_box, g0, G, dp, dq, thr_min, thr_s, n_thr);
            if (c.cnt > 1 && atomicExch(&a.grp_flag[grp], 1) == 0)
                a.complex_list[atomicAdd(a.complex_count, 1)] = grp;
        }
        const bool single = work && c.cnt == 1;
        const bool locks = single && (dfl & 2);
        const uint32_t key = single ? (((uint32_t)grp << 5) | (uint32_t)c.gs) : (0x80000000u | lane);
        // (group << 5 | g*) is exact while the group index stays below 2^26; beyond that two
        // different groups could alias, so compare the group separately
        uint32_t same = __match_any_sync(0xffffffffu, key);
        if (a.n_groups >= (1 << 26)) same &= __match_any_sync(0xffffffffu, grp);
        const uint32_t earlier = same & lanes_lt;
        uint32_t taken = 0;
        for (int k = 0; k < n_thr; ++k) {
            const uint32_t bk = __ballot_sync(0xffffffffu, locks && ((c.ge >> k) & 1u));
            if (bk & earlier) taken |= 1u << k;
        }
        const uint32_t cr = __shfl_sync(0xffffffffu, carry, c.gs & 31);
        if (single && grp == grp_f) taken |= cr;
        const uint32_t M = single ? (c.ge & ~taken) : 0u;
        // ---- TP / FP words of every cfg
        if (valid && !skip) {
            const uint32_t dm = TRACK
                ? fe_dt_unmatched_mask(rules, a.dt_a[d], a.dt_b ? a.dt_b[d] : 0.0, dfl, cfg_all)
                : fe_dt_unmatched_mask(rules, dq.x * dq.y, 0.0, dfl, cfg_all);
            uint32_t gmask = 0;
            bool sent = false;
            if (M) {
                const int64_t gg = g0 + c.gs;
                const uint8_t gfl = a.gt_flag[gg];
                gmask = TRACK
                    ? fe_gt_ignore_mask(rules, a.gt_vis[gg], a.gt_b ? a.gt_b[gg] : 0.0,
                                        a.gt_hp ? a.gt_hp[gg] : 0, gfl, cfg_all)
                    : fe_gt_ignore_mask(rules, a.gt_vis[gg], 0.0, 0, gfl, cfg_all);
                sent = (gfl & 4) != 0;
            }
            uint32_t* o = a.dt_tpfp + d * n_cfg;
            for (int cf = 0; cf < n_cfg; ++cf) {
                const bool gi = (gmask >> cf) & 1u, dc = (dm >> cf) & 1u;
                const uint32_t tp = (!sent && !gi) ? M : 0u;
                const uint32_t fp = ((sent && !gi && !dc) ? M : 0u) | (dc ? 0u : (thr_all & ~M));
                o[cf] = tp | (fp << 16);
            }
        }
    }
}

// Per 32 consecutive groups (one warp): (1) detection -> group map, coalesced writes over the
// groups' contiguous detections; (2) non-ignored GT count per (category, cfg)
// (lvis eval.py:363-365) over the groups' contiguous GT boxes — lanes that share the category
// are summed with REDUX before the atomic.  The group of a detection / GT index is found by a
// 5-step search over the 33 offsets held one per lane.
template <int NC>
__global__ void __launch_bounds__(256)
k_frame_prep(FrameArgs a, int32_t* __restrict__ dt_grp, int gpw) {
    // gpw = groups per warp (<= 32): 32 for large plans, fewer for small ones so that the work
    // spreads over more warps (a track plan has a few thousand groups)
    __shared__ ta_range_cfg cfg_s[RR_MAX];
    __shared__ double thr_s[TA_MAX_THRS];
    __shared__ FrameRules rules;
    const int n_cfg = NC ? NC : a.n_cfg;
    fe_setup(a, a.n_thr, n_cfg, cfg_s, thr_s, rules);
    const uint32_t cfg_all = (n_cfg == 32) ? 0xffffffffu : ((1u << n_cfg) - 1u);
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t grp0 = wid * gpw;
    if (grp0 >= a.n_groups) return;
    const int n_in = (int)((grp0 + gpw < a.n_groups) ? gpw : a.n_groups - grp0);
    const int li = lane < n_in ? lane : n_in;
    const int64_t doff = a.grp_dt_off[grp0 + li], goff = a.grp_gt_off[grp0 + li];
    const int64_t d_end = a.grp_dt_off[grp0 + n_in], g_end = a.grp_gt_off[grp0 + n_in];
    const int cat_l = a.grp_cat[grp0 + (lane < n_in ? lane : n_in - 1)];
    // oversize groups are evaluated AND counted by the generic matcher (big_list route)
    const int64_t dn = __shfl_down_sync(0xffffffffu, doff, 1), gn = __shfl_down_sync(0xffffffffu, goff, 1);
    const int64_t D_l = (lane + 1 < n_in ? dn : d_end) - doff, G_l = (lane + 1 < n_in ? gn : g_end) - goff;
    const bool big_l = a.exclude_big && lane < n_in &&
                       (G_l > FE_MAX_GT || D_l > FE_MAX_DT || D_l * G_l > FE_MAX_PAIRS);
    const uint32_t big_mask = __ballot_sync(0xffffffffu, big_l);
    // ---- (1)
    if (dt_grp) {
        const int64_t d_begin = __shfl_sync(0xffffffffu, doff, 0);
        for (int64_t base = d_begin; base < d_end; base += 32) {
            const int64_t dd = base + lane;
            int gi = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const int cnd = gi + step;
                const int64_t v = __shfl_sync(0xffffffffu, doff, cnd & 31);
                if (cnd < n_in && v <= dd) gi = cnd;
            }
            if (dd < d_end) dt_grp[dd] = (int32_t)(grp0 + gi);
        }
    }
    // ---- (2)
    const int64_t g_begin = __shfl_sync(0xffffffffu, goff, 0);
    for (int64_t base = g_begin; base < g_end; base += 32) {
        const int64_t gg = base + lane;
        int gi = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const int cnd = gi + step;
            const int64_t v = __shfl_sync(0xffffffffu, goff, cnd & 31);
            if (cnd < n_in && v <= gg) gi = cnd;
        }
        int cat = __shfl_sync(0xffffffffu, cat_l, gi);
        uint32_t m = cfg_all;                                  // "ignored everywhere" = counts nothing
        if (gg < g_end && !((big_mask >> gi) & 1u)) {
            const uint8_t gfl = a.gt_flag[gg];
            m = fe_gt_ignore_mask(rules, a.gt_vis[gg], a.gt_b ? a.gt_b[gg] : 0.0,
                                  a.gt_hp ? a.gt_hp[gg] : 0, gfl, cfg_all);
            // per-GT word of the streamed flat kernel: ignore mask + "id equals the unmatched value"
            if (a.gt_word_out) a.gt_word_out[gg] = m | ((gfl & 4) ? (1u << 16) : 0u);
        } else {
            cat = -1 - lane;
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, cat);
        const bool leader = (__ffs(peers) - 1) == lane;
        for (int c = 0; c < n_cfg; ++c) {
            const int tot = __reduce_add_sync(peers, (int)(!((m >> c) & 1u)));
            if (leader && cat >= 0 && tot) atomicAdd(&a.num_gt[(int64_t)cat * n_cfg + c], tot);
        }
    }
}

// ------------------------------------------------------------------------------------------
// streamed flat frame path: k_frame_sched (once per plan) + k_frame_flat (every evaluation)
//
// Same lane-per-detection scheme as k_track_flat, re-cut so that a warp never needs anything
// another warp computed and never waits on a global load of GT data:
//   * the (category, image)-sorted detections are cut into TASKS at group boundaries: a new task
//     starts at the first group whose detection offset enters a new block of FS_TASK_DT
//     detections or whose GT offset enters a new block of FS_TASK_GT boxes.  A task therefore
//     owns whole groups (no replay of a neighbour's detections), about FS_TASK_DT detections,
//     and GT boxes that all lie in one window of FS_TASK_GT + 32 consecutive boxes.
//     task index of a group = dt_off / FS_TASK_DT + gt_off / FS_TASK_GT (closed form, no scan).
//   * one warp per task.  Lane 0 fetches the task's GT window with ONE 1-D bulk copy
//     (cp.async.bulk global -> shared, completion on an mbarrier) into a double buffer: the
//     copy of task t+1 is in flight while task t is evaluated.
//   * the warp walks the task's detections 32 at a time (next window's boxes prefetched into
//     registers).  Per lane: overlap test and i, u per GT from shared memory; a pair is a
//     candidate only if i >= (thr_min - margin) * u, which needs no division — the single exact
//     i / u (bit-identical to pycocotools bbIou, maskApi.c:109-120) is taken once per detection
//     for its last candidate.  "GT g* is taken at threshold k" lives in a per-task shared table
//     (taken_s[GT]) across windows and in ballots inside a window, exactly the rule of
//     k_track_flat: matched(d) = ge(d) & ~OR{ge(d') : d' earlier, same GT, d' locks}.
//   * result: ONE 32-bit word per detection when n_thr + 3 n_cfg <= 31 (COMPACT):
//       bits [0, T)            thresholds at which the detection is matched (to its only candidate)
//       bits [T, T+C)          cfg c: matched thresholds count as TP   (GT regular, id != sentinel)
//       bits [T+C, T+2C)       cfg c: matched thresholds count as FP   (GT id == sentinel value)
//       bits [T+2C, T+3C)      cfg c: unmatched thresholds count as FP (else ignored)
//       bit 31                 the detection's full row in dt_tpfp is authoritative (general matcher)
//     which ta_pr_accumulate expands on the fly; otherwise the full [n_cfg] row as before.
// ------------------------------------------------------------------------------------------
#define FS_TASK_DT 256
#define FS_TASK_GT 64
#define FS_GT_SPAN (FS_TASK_GT + FE_MAX_GT)
#define FS_WARPS 4
#define FS_SKIP_G 63u

struct SchedLayout {
    int64_t n_tasks;
    size_t o_task_dt, o_task_gt, o_desc, o_grp, o_gtw, o_numgt, total;
};
// n_cells = n_cat * n_cfg of the plan's range cfgs (0: no GT-side section, internal schedules)
static SchedLayout fs_layout(int64_t n_dt, int64_t n_gt, int64_t n_cells = 0) {
    SchedLayout L;
    L.n_tasks = n_dt / FS_TASK_DT + n_gt / FS_TASK_GT + 1;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    L.o_task_dt = take((size_t)(L.n_tasks + 1) * 8);
    L.o_task_gt = take((size_t)(L.n_tasks + 1) * 8);
    L.o_desc = take((size_t)(n_dt > 0 ? n_dt : 1) * 4);
    L.o_grp = take((size_t)(n_dt > 0 ? n_dt : 1) * 4);
    // GT side (depends on the range cfgs): per-GT words and the non-ignored GT counts
    L.o_gtw = take((size_t)(n_gt > 0 ? n_gt : 1) * 4);
    L.o_numgt = take(16 + (size_t)n_cells * 4);        // int32 n, then int32 [n_cat][n_cfg]
    L.total = off;
    return L;
}

// Per 32 consecutive groups (one warp): task table entries of the groups that open a task,
// and per detection the descriptor word + group index (coalesced over the groups' contiguous
// detections; the group of a detection is found by a 5-step search over the 33 offsets held
// one per lane).
__global__ void __launch_bounds__(256)
k_frame_sched(int64_t n_groups, const int64_t* __restrict__ grp_dt_off,
              const int64_t* __restrict__ grp_gt_off, const uint8_t* __restrict__ dt_flag,
              int64_t n_tasks, int64_t* __restrict__ task_dt, int64_t* __restrict__ task_gt,
              uint32_t* __restrict__ dt_desc, int32_t* __restrict__ dt_grp) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t grp0 = wid * 32;
    if (grp0 >= n_groups) return;
    const int n_in = (int)((grp0 + 32 < n_groups) ? 32 : n_groups - grp0);
    const int li = lane < n_in ? lane : n_in;
    const int64_t doff = grp_dt_off[grp0 + li], goff = grp_gt_off[grp0 + li];
    const int64_t d_end = grp_dt_off[grp0 + n_in], g_end = grp_gt_off[grp0 + n_in];
    if (lane < n_in) {
        const int64_t g = grp0 + lane;
        const int64_t idx = doff / FS_TASK_DT + goff / FS_TASK_GT;
        int64_t prev = -1;
        if (g > 0) prev = grp_dt_off[g - 1] / FS_TASK_DT + grp_gt_off[g - 1] / FS_TASK_GT;
        for (int64_t i = prev + 1; i <= idx; ++i) { task_dt[i] = doff; task_gt[i] = goff; }
        if (g == n_groups - 1) {
            const int64_t de = grp_dt_off[n_groups], ge = grp_gt_off[n_groups];
            for (int64_t i = idx + 1; i <= n_tasks; ++i) { task_dt[i] = de; task_gt[i] = ge; }
        }
    }
    const int64_t gn = __shfl_down_sync(0xffffffffu, goff, 1);
    const int64_t G_l = (lane + 1 < n_in ? gn : g_end) - goff;
    const int64_t d_begin = __shfl_sync(0xffffffffu, doff, 0);
    for (int64_t base = d_begin; base < d_end; base += 32) {
        const int64_t dd = base + lane;
        int gi = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const int cnd = gi + step;
            const int64_t v = __shfl_sync(0xffffffffu, doff, cnd & 31);
            if (cnd < n_in && v <= dd) gi = cnd;
        }
        const int64_t go = __shfl_sync(0xffffffffu, goff, gi);
        const int64_t Gg = __shfl_sync(0xffffffffu, G_l, gi);
        if (dd < d_end) {
            const uint32_t Gf = Gg > FE_MAX_GT ? FS_SKIP_G : (uint32_t)Gg;
            dt_desc[dd] = ((uint32_t)(go & (FS_TASK_GT - 1)) << 8) | (Gf << 2) | (dt_flag[dd] & 3u);
            dt_grp[dd] = (int32_t)(grp0 + gi);
        }
    }
}

// ---- mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP) --------------------------
__device__ __forceinline__ uint32_t fs_smem(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void fs_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fs_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void fs_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fs_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fs_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

// the detection-side rule tables the flat kernel needs (area intervals of the unmatched-ignore
// test; the b attribute of a frame detection is the constant 0)
struct FlatRules {
    int n_da;
    uint32_t const_mask;          // cfgs whose dt_b interval excludes 0
    double lo[RR_MAX], hi[RR_MAX];
    uint32_t mask[RR_MAX];
};

template <int NT, int NC, bool COMPACT>
__global__ void __launch_bounds__(FS_WARPS * 32, 8)
k_frame_flat(FrameArgs a) {
    __shared__ __align__(128) double gt_s[FS_WARPS][2][FS_GT_SPAN * 4];
    __shared__ uint32_t taken_s[FS_WARPS][FS_GT_SPAN];
    __shared__ __align__(8) unsigned long long bar_s[FS_WARPS][2];
    __shared__ double thr_s[TA_MAX_THRS];
    __shared__ FlatRules fr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_thr = NT ? NT : a.n_thr, n_cfg = NC ? NC : a.n_cfg;
    if (threadIdx.x < n_thr) {
        const double th = a.thrs[threadIdx.x];
        thr_s[threadIdx.x] = (th < 1.0 - 1e-10) ? th : 1.0 - 1e-10;
    }
    if (threadIdx.x == 32) {
        const FrameRules* r = a.rules_g;
        fr.n_da = r->n_da;
        uint32_t cm = 0;
        for (int k = 0; k < r->n_db; ++k)
            if (0.0 < r->db_lo[k] || 0.0 > r->db_hi[k]) cm |= r->db_mask[k];
        fr.const_mask = cm;
        for (int k = 0; k < r->n_da; ++k) { fr.lo[k] = r->da_lo[k]; fr.hi[k] = r->da_hi[k]; fr.mask[k] = r->da_mask[k]; }
    }
    const uint32_t bar0 = fs_smem(&bar_s[warp][0]), bar1 = fs_smem(&bar_s[warp][1]);
    if (lane == 0) {
        fs_mbar_init(bar0, 1);
        fs_mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t cfg_all = (n_cfg == 32) ? 0xffffffffu : ((1u << n_cfg) - 1u);
    const uint32_t thr_all = (1u << n_thr) - 1u;
    const uint32_t lanes_lt = (1u << lane) - 1u;
    double thr_min = thr_s[0];
    for (int i = 1; i < n_thr; ++i) thr_min = (thr_s[i] < thr_min) ? thr_s[i] : thr_min;
    // candidate filter: i < lo * u  =>  fl(i / u) < thr_min  (lo sits 2^-20 relative below thr_min,
    // the rounding errors of lo * u and i / u are 2^-53); thresholds <= 0 (or NaN) make every
    // pair a candidate, as in the reference's `iou < best` test
    const bool always = !(thr_min > 0.0);
    const double lo = thr_min * (1.0 - 1.0 / 1048576.0);
    uint32_t* taken_w = taken_s[warp];

    const int64_t n_tasks = a.n_tasks;
    const int64_t w0 = (int64_t)blockIdx.x * FS_WARPS + warp;
    const int64_t wstride = (int64_t)gridDim.x * FS_WARPS;

    // GT window of a task: boxes [task_gt[t] & ~63, task_gt[t + 1]), at most FS_GT_SPAN of them
    // (beyond that only oversize groups, which this kernel skips)
    auto issue = [&](int64_t t, int buf) {
        const int64_t g_lo = a.task_gt[t] & ~(int64_t)(FS_TASK_GT - 1);
        int64_t n = a.task_gt[t + 1] - g_lo;
        if (n > FS_GT_SPAN) n = FS_GT_SPAN;
        const uint32_t bar = buf ? bar1 : bar0;
        if (n > 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            fs_mbar_expect_tx(bar, (uint32_t)n * 32u);
            fs_bulk_g2s(fs_smem(&gt_s[warp][buf][0]), a.gt_box + 4 * g_lo, (uint32_t)n * 32u, bar);
        } else {
            fs_mbar_arrive(bar);
        }
    };

    if (w0 < n_tasks && lane == 0) issue(w0, 0);
    int it = 0;
    for (int64_t t = w0; t < n_tasks; t += wstride, ++it) {
        const int cur = it & 1;
        if (t + wstride < n_tasks && lane == 0) issue(t + wstride, cur ^ 1);
        const int64_t d_begin = a.task_dt[t], d_end = a.task_dt[t + 1];
        const int64_t g_base = a.task_gt[t] & ~(int64_t)(FS_TASK_GT - 1);
        for (int i = lane; i < FS_GT_SPAN; i += 32) taken_w[i] = 0u;
        // first window's detections
        double2 dp = make_double2(0, 0), dq = make_double2(0, 0);
        uint32_t desc = 0;
        if (d_begin + lane < d_end) {
            dp = *reinterpret_cast<const double2*>(a.dt_box + 4 * (d_begin + lane));
            dq = *reinterpret_cast<const double2*>(a.dt_box + 4 * (d_begin + lane) + 2);
            desc = a.dt_desc[d_begin + lane];
        }
        fs_mbar_wait(cur ? bar1 : bar0, (uint32_t)(it >> 1) & 1u);
        const double* gt_cur = gt_s[warp][cur];
        for (int64_t base = d_begin; base < d_end; base += 32) {
            const int64_t d = base + lane;
            const bool valid = d < d_end;
            // next window's loads are in flight while this one is evaluated
            double2 pn = make_double2(0, 0), qn = make_double2(0, 0);
            uint32_t descn = 0;
            if (d + 32 < d_end) {
                pn = *reinterpret_cast<const double2*>(a.dt_box + 4 * (d + 32));
                qn = *reinterpret_cast<const double2*>(a.dt_box + 4 * (d + 32) + 2);
                descn = a.dt_desc[d + 32];
            }
            const uint32_t G = (desc >> 2) & 63u, rel = desc >> 8, dfl = desc & 3u;
            const bool skip = G == FS_SKIP_G;
            const bool work = valid && !skip && G > 0u;
            const double da = dq.x * dq.y;
            const double r0 = dq.x + dp.x, b0 = dq.y + dp.y;
            int cnt = 0;
            uint32_t gs = 0;
            double ci = 0.0, cu = 1.0;
            const int Gm = __reduce_max_sync(0xffffffffu, work ? (int)G : 0);
            uint32_t M = 0;
            __syncwarp();                      // taken_w updates of the previous window are visible
            if (Gm > 0) {
                const double* gb = gt_cur + 4 * rel;
                for (int j = 0; j < Gm; ++j) {
                    if (work && j < (int)G) {
                        const double2 p = *reinterpret_cast<const double2*>(gb + 4 * j);
                        const double2 q = *reinterpret_cast<const double2*>(gb + 4 * j + 2);
                        // maskApi.c:109-120 without the division
                        const double ga = q.x * q.y;
                        const double r1 = q.x + p.x, b1 = q.y + p.y;
                        const double w = ((r0 < r1) ? r0 : r1) - ((dp.x > p.x) ? dp.x : p.x);
                        const double h = ((b0 < b1) ? b0 : b1) - ((dp.y > p.y) ? dp.y : p.y);
                        const bool ov = !(w <= 0.0) && !(h <= 0.0);
                        const double ii = w * h;
                        const double uu = da + ga - ii;
                        const uint32_t ex = ((uint32_t)__double2hiint(uu) >> 20) & 0x7ffu;
                        const bool mb = always || (ov && (!(ii < lo * uu) || (ex - 200u) > 1600u));
                        if (mb) { ++cnt; gs = (uint32_t)j; ci = ov ? ii : 0.0; cu = ov ? uu : 1.0; }
                    }
                }
                uint32_t ge = 0;
                if (cnt == 1) {
                    const double v = ci / cu;
                    if (v < thr_min) cnt = 0;
                    else
                        for (int k = 0; k < n_thr; ++k) ge |= (!(v < thr_s[k])) ? (1u << k) : 0u;
                }
                if (cnt > 1) {
                    const int grp = a.dt_grp[d];
                    if (atomicExch(&a.grp_flag[grp], 1) == 0)
                        a.complex_list[atomicAdd(a.complex_count, 1)] = grp;
                }
                const bool single = cnt == 1;
                const bool locks = single && (dfl & 2u);
                const uint32_t gidx = rel + gs;                // GT position inside the task window
                const uint32_t key = single ? gidx : (0x100u | (uint32_t)lane);
                if (__any_sync(0xffffffffu, single)) {
                    const uint32_t same = __match_any_sync(0xffffffffu, key);
                    const uint32_t earlier = same & lanes_lt;
                    uint32_t taken = single ? taken_w[gidx] : 0u;
                    for (int k = 0; k < n_thr; ++k) {
                        const uint32_t bk = __ballot_sync(0xffffffffu, locks && ((ge >> k) & 1u));
                        if (bk & earlier) taken |= 1u << k;
                    }
                    __syncwarp();              // every lane has read taken_w
                    if (locks) atomicOr(&taken_w[gidx], ge);
                    M = single ? (ge & ~taken) : 0u;
                }
            }
            // ---- result
            if (valid) {
                uint32_t dm = ((dfl & 1u) ? cfg_all : 0u) | fr.const_mask;
                for (int k = 0; k < fr.n_da; ++k)
                    if (da < fr.lo[k] || da > fr.hi[k]) dm |= fr.mask[k];
                uint32_t gmask = 0;
                bool sent = false;
                if (M) {
                    const uint32_t gw = a.gt_word[g_base + rel + gs];
                    gmask = gw & 0xffffu;
                    sent = (gw >> 16) & 1u;
                }
                if (COMPACT) {
                    uint32_t word;
                    if (skip || cnt > 1) {
                        word = TA_WORD_FULL;
                    } else {
                        const uint32_t nig = ~gmask & cfg_all, ndc = ~dm & cfg_all;
                        const uint32_t A = (M && !sent) ? nig : 0u;
                        const uint32_t B = (M && sent) ? (nig & ndc) : 0u;
                        word = M | (A << n_thr) | (B << (n_thr + n_cfg)) | (ndc << (n_thr + 2 * n_cfg));
                    }
                    a.dt_word[d] = word;
                } else if (!skip) {
                    uint32_t* o = a.dt_tpfp + d * n_cfg;
                    for (int cf = 0; cf < n_cfg; ++cf) {
                        const bool gi = (gmask >> cf) & 1u, dc = (dm >> cf) & 1u;
                        const uint32_t tp = (!sent && !gi) ? M : 0u;
                        const uint32_t fp = ((sent && !gi && !dc) ? M : 0u) | (dc ? 0u : (thr_all & ~M));
                        o[cf] = tp | (fp << 16);
                    }
                }
            }
            dp = pn; dq = qn; desc = descn;
        }
        __syncwarp();                          // buffer `cur` and taken_w fully consumed
    }
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" int ta_frame_eval_max_gt(void) { return FE_MAX_GT; }
extern "C" int ta_frame_eval_max_dt(void) { return FE_MAX_DT; }
extern "C" int ta_frame_eval_max_pairs(void) { return FE_MAX_PAIRS; }

extern "C" int ta_match_greedy(ta_ctx* ctx, void* stream, int64_t n_groups,
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
                               int32_t* dt_match_gt, uint8_t* gt_ignore_out) {
    if (!ctx) return ta_set_err(TA_ERR_INVALID, "ta_match_greedy: ctx is NULL");
    if (n_thr < 1 || n_thr > TA_MAX_THRS)
        return ta_set_err(TA_ERR_INVALID, "ta_match_greedy: n_thr must be in [1,16], got %s%lld", "", n_thr);
    if (n_cfg < 1 || n_groups < 0 || g_max < 0)
        return ta_set_err(TA_ERR_INVALID, "ta_match_greedy: bad sizes");
    if (grp_list) n_groups = n_list;
    if (n_groups <= 0) return TA_OK;
    if (n_groups > INT_MAX) return ta_set_err(TA_ERR_TOO_LARGE, "ta_match_greedy: too many groups");
    TA_CUDA(cudaSetDevice(ctx->device));
    ta_begin(ctx, (cudaStream_t)stream);
    const int cpw = 32 / n_thr;
    const int warps = (n_cfg + cpw - 1) / cpw;
    if (warps * 32 > 1024) return ta_set_err(TA_ERR_TOO_LARGE, "ta_match_greedy: too many range cfgs");
    const int threads = warps * 32;
    const int words = (g_max + 31) / 32;
    const size_t smem_tab = (size_t)words * threads * sizeof(uint32_t) + (size_t)n_cfg * g_max;
    // room for the IoU matrix of a group behind the tables (up to 32 KB)
    size_t smem = smem_tab;
    int iou_cap = 0;
    if (smem_tab + 8 + 4096 * 8 <= (size_t)ctx->smem_optin) {
        iou_cap = 4096;
        smem = ((smem_tab + 7) & ~(size_t)7) + (size_t)iou_cap * 8;
    }
    if (smem > (size_t)ctx->smem_optin)
        return ta_set_err(TA_ERR_TOO_LARGE,
                          "ta_match_greedy: a group with %s%lld ground-truth entities does not fit in shared memory",
                          "", (long long)g_max);
    MatchArgs a{grp_list, grp_dt_off, grp_gt_off, grp_cat, iou_off, iou, n_thr, iou_thrs, n_cfg, cfgs,
                n_dt, n_gt, dt_attr_a, dt_attr_b, dt_flag, gt_attr_a, gt_attr_b, gt_hp,
                gt_flag, dt_tpfp, num_gt, dt_match_gt, gt_ignore_out, cpw, nullptr, nullptr, 0, iou_cap};
    if (smem > 48 * 1024)
        TA_CUDA(cudaFuncSetAttribute(k_match_greedy, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    cudaStream_t st = (cudaStream_t)stream;
    if (!grp_list && !dt_match_gt && !gt_ignore_out && n_cfg <= 32 && n_dt > 0) {
        // evaluation route: GT counts + lane-per-detection matcher on the IoU matrix; the general
        // matcher below only sees the groups whose detections have several candidate GTs
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
        const size_t o_rules = take(sizeof(FrameRules));
        const size_t o_grp = take((size_t)n_dt * 4);
        const size_t o_flag = take((size_t)n_groups * 4 + 4);
        const size_t o_list = take((size_t)n_groups * 4);
        void* ws2 = nullptr;
        int rc = ta_workspace(ctx, st, off, &ws2, 1);
        if (rc) return rc;
        char* base = static_cast<char*>(ws2);
        FrameRules* rules_g = reinterpret_cast<FrameRules*>(base + o_rules);
        k_frame_rules<<<1, 32, 0, st>>>(cfgs, n_cfg, rules_g);
        if ((rc = ta_check_launch(ctx, "k_frame_rules"))) return rc;
        FrameArgs f{n_groups, grp_dt_off, grp_gt_off, grp_cat, nullptr, nullptr, n_thr, iou_thrs, n_cfg,
                    cfgs, n_dt, n_gt, dt_flag, gt_attr_a, gt_flag, iou_off, const_cast<double*>(iou), 0,
                    dt_tpfp, num_gt, nullptr, nullptr,
                    dt_attr_a, dt_attr_b, gt_attr_b, gt_hp, 0,
                    reinterpret_cast<int32_t*>(base + o_grp),
                    reinterpret_cast<int32_t*>(base + o_flag), reinterpret_cast<int32_t*>(base + o_list),
                    reinterpret_cast<int32_t*>(base + o_flag) + n_groups, rules_g};
        TA_CUDA(cudaMemsetAsync(f.grp_flag, 0, (size_t)n_groups * 4 + 4, st));
        const int gpw = n_groups >= 200000 ? 32 : 4;
        const int64_t prep_warps = (n_groups + gpw - 1) / gpw;
        k_frame_prep<0><<<(unsigned)((prep_warps + 7) / 8), 256, 0, st>>>(f, const_cast<int32_t*>(f.dt_grp), gpw);
        if ((rc = ta_check_launch(ctx, "k_frame_prep"))) return rc;
        int64_t blocks = ((n_dt + 31) / 32 + FF_WARPS - 1) / FF_WARPS;
        const int64_t fcap = (int64_t)ctx->sm_count * 16;
        if (blocks > fcap) blocks = fcap;
        k_track_flat<0, 0, true><<<(unsigned)blocks, FF_WARPS * 32, 0, st>>>(f);
        if ((rc = ta_check_launch(ctx, "k_track_flat"))) return rc;
        a.dev_list = f.complex_list;
        a.dev_count = f.complex_count;
        ctx->last_list_count = f.complex_count;
        a.skip_num_gt = 1;
        int64_t lblocks = (int64_t)ctx->sm_count * 4;
        if (lblocks > n_groups) lblocks = n_groups;
        k_match_greedy<<<(unsigned)lblocks, threads, smem, st>>>(a);
        return ta_check_launch(ctx, "k_match_greedy_list");
    }
    k_match_greedy<<<(unsigned)n_groups, threads, smem, st>>>(a);
    return ta_check_launch(ctx, "k_match_greedy");
}

// dt area for the big-group route of the frame path: w*h of the box (lvis results.py:56);
// with compact result words the detections of these groups point to their full rows
__global__ void k_box_area_list(int64_t n_list, const int32_t* __restrict__ grp_list,
                                const int64_t* __restrict__ grp_dt_off,
                                const double* __restrict__ dt_box, double* __restrict__ area,
                                uint32_t* __restrict__ dt_word) {
    const int64_t grp = grp_list[blockIdx.x];
    const int64_t d0 = grp_dt_off[grp], d1 = grp_dt_off[grp + 1];
    for (int64_t d = d0 + threadIdx.x; d < d1; d += blockDim.x) {
        area[d] = dt_box[4 * d + 2] * dt_box[4 * d + 3];
        if (dt_word) dt_word[d] = TA_WORD_FULL;
    }
}

extern "C" int64_t ta_frame_sched_bytes(int64_t n_groups, int64_t n_dt, int64_t n_gt,
                                        int32_t n_cat, int32_t n_cfg) {
    (void)n_groups;
    if (n_dt < 0 || n_gt < 0 || n_cat < 0 || n_cfg < 0) return 0;
    return (int64_t)fs_layout(n_dt, n_gt, (int64_t)n_cat * n_cfg).total;
}

static int fs_build(ta_ctx* ctx, cudaStream_t st, int64_t n_groups, const int64_t* grp_dt_off,
                    const int64_t* grp_gt_off, int64_t n_dt, const uint8_t* dt_flag, int64_t n_gt,
                    void* sched, const SchedLayout& L) {
    char* b = static_cast<char*>(sched);
    const int64_t warps = (n_groups + 31) / 32;
    k_frame_sched<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(
        n_groups, grp_dt_off, grp_gt_off, dt_flag, L.n_tasks,
        reinterpret_cast<int64_t*>(b + L.o_task_dt), reinterpret_cast<int64_t*>(b + L.o_task_gt),
        reinterpret_cast<uint32_t*>(b + L.o_desc), reinterpret_cast<int32_t*>(b + L.o_grp));
    return ta_check_launch(ctx, "k_frame_sched");
}

// num_gt[i] += part[i], i < part_hdr[0] (the schedule's precomputed non-ignored GT counts)
__global__ void k_add_counts(const int32_t* __restrict__ hdr, int32_t* __restrict__ num_gt) {
    const int n = hdr[0];
    const int32_t* part = hdr + 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (part[i]) num_gt[i] += part[i];
}
__global__ void k_set_header(int32_t* hdr, int n) { hdr[0] = n; hdr[1] = hdr[2] = hdr[3] = 0; }

extern "C" int ta_frame_sched_build(ta_ctx* ctx, void* stream, int64_t n_groups,
                                    const int64_t* grp_dt_off, const int64_t* grp_gt_off,
                                    const int32_t* grp_cat, int64_t n_dt, const uint8_t* dt_flag,
                                    int64_t n_gt, const double* gt_attr_a, const uint8_t* gt_flag,
                                    int32_t n_cat, int32_t n_cfg, const ta_range_cfg* cfgs,
                                    void* sched) {
    if (!ctx || !sched) return ta_set_err(TA_ERR_INVALID, "ta_frame_sched_build: NULL argument");
    if (n_groups < 0 || n_dt < 0 || n_gt < 0 || n_cat < 0)
        return ta_set_err(TA_ERR_INVALID, "ta_frame_sched_build: negative size");
    if (n_cfg < 1 || n_cfg > FE_MAX_CFG) return ta_set_err(TA_ERR_INVALID, "ta_frame_sched_build: bad n_cfg");
    if (n_groups > INT_MAX) return ta_set_err(TA_ERR_TOO_LARGE, "ta_frame_sched_build: too many groups");
    if (n_groups == 0) return TA_OK;
    TA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    ta_begin(ctx, st);
    const SchedLayout L = fs_layout(n_dt, n_gt, (int64_t)n_cat * n_cfg);
    int rc = fs_build(ctx, st, n_groups, grp_dt_off, grp_gt_off, n_dt, dt_flag, n_gt, sched, L);
    if (rc) return rc;
    // GT side: per-GT words + non-ignored GT counts of the groups the streamed kernel evaluates
    char* b = static_cast<char*>(sched);
    int32_t* hdr = reinterpret_cast<int32_t*>(b + L.o_numgt);
    TA_CUDA(cudaMemsetAsync(hdr, 0, 16 + (size_t)n_cat * n_cfg * 4, st));
    k_set_header<<<1, 1, 0, st>>>(hdr, n_cat * n_cfg);
    if ((rc = ta_check_launch(ctx, "k_set_header"))) return rc;
    void* ws2 = nullptr;
    if ((rc = ta_workspace(ctx, st, sizeof(FrameRules) + 256, &ws2, 1))) return rc;
    FrameRules* rules_g = reinterpret_cast<FrameRules*>(ws2);
    k_frame_rules<<<1, 32, 0, st>>>(cfgs, n_cfg, rules_g);
    if ((rc = ta_check_launch(ctx, "k_frame_rules"))) return rc;
    FrameArgs a{n_groups, grp_dt_off, grp_gt_off, grp_cat, nullptr, nullptr, 0, nullptr, n_cfg,
                cfgs, n_dt, n_gt, dt_flag, gt_attr_a, gt_flag, nullptr, nullptr, 0,
                nullptr, hdr + 4, nullptr, nullptr,
                nullptr, nullptr, nullptr, nullptr, 1, nullptr, nullptr, nullptr, nullptr, rules_g};
    a.gt_word_out = reinterpret_cast<uint32_t*>(b + L.o_gtw);
    const int gpw = n_groups >= 200000 ? 32 : 4;
    const int64_t prep_warps = (n_groups + gpw - 1) / gpw;
    k_frame_prep<0><<<(unsigned)((prep_warps + 7) / 8), 256, 0, st>>>(a, nullptr, gpw);
    return ta_check_launch(ctx, "k_frame_prep");
}

extern "C" int ta_frame_eval(ta_ctx* ctx, void* stream, int64_t n_groups,
                             const int64_t* grp_dt_off, const int64_t* grp_gt_off,
                             const int32_t* grp_cat, const double* dt_box, const double* gt_box,
                             int32_t n_thr, const double* iou_thrs, int32_t n_cfg,
                             const ta_range_cfg* cfgs, int64_t n_dt, const uint8_t* dt_flag,
                             int64_t n_gt, const double* gt_attr_a, const uint8_t* gt_flag,
                             int64_t n_big, const int32_t* big_list, int32_t g_max_big,
                             const int64_t* iou_off, double* iou, int32_t write_iou,
                             const void* sched, uint32_t* dt_word,
                             uint32_t* dt_tpfp, int32_t* num_gt,
                             int32_t* dt_match_gt, uint8_t* gt_ignore_out) {
    if (!ctx) return ta_set_err(TA_ERR_INVALID, "ta_frame_eval: ctx is NULL");
    if (n_thr < 1 || n_thr > TA_MAX_THRS)
        return ta_set_err(TA_ERR_INVALID, "ta_frame_eval: n_thr must be in [1,16], got %s%lld", "", n_thr);
    if (n_cfg < 1 || n_cfg > FE_MAX_CFG || n_groups < 0 || n_big < 0)
        return ta_set_err(TA_ERR_INVALID, "ta_frame_eval: bad sizes");
    if ((write_iou || n_big > 0) && (!iou || !iou_off))
        return ta_set_err(TA_ERR_INVALID, "ta_frame_eval: iou storage required");
    if (n_groups == 0) return TA_OK;
    if (n_groups > INT_MAX) return ta_set_err(TA_ERR_TOO_LARGE, "ta_frame_eval: too many groups");
    TA_CUDA(cudaSetDevice(ctx->device));
    ta_begin(ctx, (cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    FrameArgs a{n_groups, grp_dt_off, grp_gt_off, grp_cat, dt_box, gt_box, n_thr, iou_thrs, n_cfg,
                cfgs, n_dt, n_gt, dt_flag, gt_attr_a, gt_flag, iou_off, iou, write_iou,
                dt_tpfp, num_gt, dt_match_gt, gt_ignore_out,
                nullptr, nullptr, nullptr, nullptr, 1, nullptr, nullptr, nullptr, nullptr, nullptr};
    const int64_t cap = (int64_t)ctx->sm_count * 8;      // persistent: 8 CTAs per SM
    const bool detail = write_iou || dt_match_gt || gt_ignore_out;
    // compact result words: only the evaluation route, and only when a word holds T + 3 C bits
    const bool compact = !detail && dt_word != nullptr && n_thr + 3 * n_cfg <= 31;
    int rc;
    // scratch (slot 1): rule tables, per-GT words, complex-group flags / list, and the schedule
    // when the caller did not build one (ta_frame_sched_build)
    const SchedLayout L = fs_layout(n_dt, n_gt);       // the offsets used here do not depend on n_cells
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_rules = take(sizeof(FrameRules));
    const size_t o_gtw = take((size_t)(n_gt > 0 ? n_gt : 1) * 4);
    const size_t o_flag = take((size_t)n_groups * 4 + 4);      // flags + the list counter
    const size_t o_list = take((size_t)n_groups * 4);
    const size_t o_sched = take((detail || sched) ? 0 : L.total);
    void* ws2 = nullptr;
    if ((rc = ta_workspace(ctx, st, off, &ws2, 1))) return rc;
    char* base = static_cast<char*>(ws2);
    FrameRules* rules_g = reinterpret_cast<FrameRules*>(base + o_rules);
    k_frame_rules<<<1, 32, 0, st>>>(cfgs, n_cfg, rules_g);
    if ((rc = ta_check_launch(ctx, "k_frame_rules"))) return rc;
    a.rules_g = rules_g;
    if (detail) {
        // detail outputs: the warp-per-group kernel does everything
        const int64_t n_tasks = (n_groups + FE_RUN - 1) / FE_RUN;
        int64_t blocks = (n_tasks + FE_WARPS - 1) / FE_WARPS;
        if (blocks > cap) blocks = cap;
        k_frame_eval<true, 0, 0, false><<<(unsigned)blocks, FE_WARPS * 32, 0, st>>>(a);
        rc = ta_check_launch(ctx, "k_frame_eval");
    } else {
        // evaluation path: GT counts + per-GT words, the streamed lane-per-detection kernel, then
        // the general matcher on the (few) groups whose detections have several candidate GTs
        const bool own_sched = sched != nullptr;     // built by ta_frame_sched_build: GT side included
        if (!sched && n_dt > 0) {
            if ((rc = fs_build(ctx, st, n_groups, grp_dt_off, grp_gt_off, n_dt, dt_flag, n_gt,
                               base + o_sched, L)))
                return rc;
            sched = base + o_sched;
        }
        const char* sb = static_cast<const char*>(sched);
        a.n_tasks = L.n_tasks;
        if (sb) {
            a.task_dt = reinterpret_cast<const int64_t*>(sb + L.o_task_dt);
            a.task_gt = reinterpret_cast<const int64_t*>(sb + L.o_task_gt);
            a.dt_desc = reinterpret_cast<const uint32_t*>(sb + L.o_desc);
            a.dt_grp = reinterpret_cast<const int32_t*>(sb + L.o_grp);
        }
        if (own_sched) {
            a.gt_word = reinterpret_cast<const uint32_t*>(sb + L.o_gtw);
        } else {
            a.gt_word_out = reinterpret_cast<uint32_t*>(base + o_gtw);
            a.gt_word = a.gt_word_out;
        }
        a.dt_word = compact ? dt_word : nullptr;
        a.grp_flag = reinterpret_cast<int32_t*>(base + o_flag);
        a.complex_count = a.grp_flag + n_groups;
        a.complex_list = reinterpret_cast<int32_t*>(base + o_list);
        TA_CUDA(cudaMemsetAsync(a.grp_flag, 0, (size_t)n_groups * 4 + 4, st));
        ctx->last_list_count = a.complex_count;
        const bool spec = (n_thr == 10 && n_cfg == 6);
        if (own_sched) {
            // the schedule holds the GT side: add its counts, no per-call pass over the GT
            k_add_counts<<<64, 256, 0, st>>>(reinterpret_cast<const int32_t*>(sb + L.o_numgt), num_gt);
            if ((rc = ta_check_launch(ctx, "k_add_counts"))) return rc;
        } else {
            const int gpw = n_groups >= 200000 ? 32 : 4;
            const int64_t prep_warps = (n_groups + gpw - 1) / gpw;
            const unsigned pb = (unsigned)((prep_warps + 7) / 8);
            if (spec) k_frame_prep<6><<<pb, 256, 0, st>>>(a, nullptr, gpw);
            else k_frame_prep<0><<<pb, 256, 0, st>>>(a, nullptr, gpw);
            if ((rc = ta_check_launch(ctx, "k_frame_prep"))) return rc;
        }
        if (n_dt > 0) {
            int64_t blocks = (L.n_tasks + FS_WARPS - 1) / FS_WARPS;
            const int64_t fcap = (int64_t)ctx->sm_count * 8;
            if (blocks > fcap) blocks = fcap;
            const unsigned nb = (unsigned)blocks, nt = FS_WARPS * 32;
            if (spec && compact) k_frame_flat<10, 6, true><<<nb, nt, 0, st>>>(a);
            else if (spec) k_frame_flat<10, 6, false><<<nb, nt, 0, st>>>(a);
            else if (compact) k_frame_flat<0, 0, true><<<nb, nt, 0, st>>>(a);
            else k_frame_flat<0, 0, false><<<nb, nt, 0, st>>>(a);
            if ((rc = ta_check_launch(ctx, "k_frame_flat"))) return rc;
        }
        if (n_dt > 0) {
            const int64_t blocks = ctx->sm_count * 8;       // list length is only known on the device
            if (spec) k_frame_eval<false, 10, 6, true><<<(unsigned)blocks, FE_WARPS * 32, 0, st>>>(a);
            else k_frame_eval<false, 0, 0, true><<<(unsigned)blocks, FE_WARPS * 32, 0, st>>>(a);
            rc = ta_check_launch(ctx, "k_frame_eval_list");
        }
    }
    if (rc || n_big == 0) return rc;
    // oversize groups: generic kernels, IoU through `iou`; their detection areas go to scratch
    void* ws = nullptr;
    rc = ta_workspace(ctx, st, (size_t)n_dt * sizeof(double), &ws);
    if (rc) return rc;
    k_box_area_list<<<(unsigned)n_big, 128, 0, st>>>(n_big, big_list, grp_dt_off, dt_box, (double*)ws,
                                                       compact ? dt_word : nullptr);
    rc = ta_check_launch(ctx, "k_box_area_list");
    if (rc) return rc;
    rc = ta_box_iou(ctx, stream, n_groups, big_list, n_big, grp_dt_off, grp_gt_off, dt_box, gt_box,
                    iou_off, iou);
    if (rc) return rc;
    return ta_match_greedy(ctx, stream, n_groups, big_list, n_big, grp_dt_off, grp_gt_off, grp_cat,
                           iou_off, iou, n_thr, iou_thrs, n_cfg, cfgs, n_dt, (const double*)ws,
                           nullptr, dt_flag, n_gt, gt_attr_a, nullptr, nullptr, gt_flag, g_max_big,
                           dt_tpfp, num_gt, dt_match_gt, gt_ignore_out);
}
