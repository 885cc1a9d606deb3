// ta_device_fns.cuh — per-thread arithmetic shared by the kernels of ta_eval.cu.
//
// Everything here is scalar IEEE fp64 / integer code, written so that it also compiles as
// plain C++ (TA_HD expands to nothing under g++): tests/hostsim builds these exact functions
// for the host to check the kernel logic against the oracle without a GPU.  That build is a
// test artefact only; the product library has no CPU path.
#ifndef TA_DEVICE_FNS_CUH
#define TA_DEVICE_FNS_CUH

#include <stdint.h>
#include "ta_eval.h"

#ifdef __CUDACC__
#define TA_HD __host__ __device__ __forceinline__
#else
#define TA_HD inline
#endif

// Intersection area of a detection and a GT box given as corners.
// tao_amodal/evaluation/tao_amodal/eval.py:38-46 with Python's max/min tie rules:
//   max(dx, gx) keeps dx unless gx > dx; min(a, b) keeps a unless b < a;
//   max(w, 0) yields 0 only when 0 > w.
TA_HD double ta_inter_corners(double dx, double dy, double dx2, double dy2,
                              double gx, double gy, double gx2, double gy2) {
    const double left = (gx > dx) ? gx : dx;
    const double right = (gx2 < dx2) ? gx2 : dx2;
    const double top = (gy > dy) ? gy : dy;
    const double bottom = (gy2 < dy2) ? gy2 : dy2;
    const double w = right - left;
    const double h = bottom - top;
    const double ii = w * h;
    return ((0.0 > w) || (0.0 > h)) ? 0.0 : ii;
}

// pycocotools bbIou, iscrowd = 0 (in-tree copy: visualization/tao/third_party/pysot/
// training_dataset/coco/pycocotools/common/maskApi.c:109-120).
TA_HD double ta_bb_iou(double dx, double dy, double dw, double dh,
                       double gx, double gy, double gw, double gh) {
    const double da = dw * dh, ga = gw * gh;
    const double r0 = dw + dx, r1 = gw + gx;
    const double w = ((r0 < r1) ? r0 : r1) - ((dx > gx) ? dx : gx);
    if (w <= 0.0) return 0.0;
    const double b0 = dh + dy, b1 = gh + gy;
    const double h = ((b0 < b1) ? b0 : b1) - ((dy > gy) ? dy : gy);
    if (h <= 0.0) return 0.0;
    const double i = w * h;
    const double u = da + ga - i;
    return i / u;
}

// One (predicted track, GT track) pair by a sequential merge of the two slot-sorted box
// lists — the term structure of eval.py:83-96 (3D), :99-117 (avg), :51-70 (imagenetvid),
// terms taken in ascending frame order.
TA_HD double ta_pair_iou_merge(const double* db, const int32_t* ds, int nd,
                               const double* gb, const int32_t* gs, int ng,
                               int mode, int* assert_failed) {
    int a = 0, b = 0;
    double i = 0.0, u = 0.0, ratio_sum = 0.0;
    long long total = 0, matched = 0;
    while (a < nd || b < ng) {
        const int sa = (a < nd) ? ds[a] : 0x7fffffff;
        const int sb = (b < ng) ? gs[b] : 0x7fffffff;
        if (sa == sb) {
            const double* d = db + 4 * a;
            const double* g = gb + 4 * b;
            const double da = d[2] * d[3], ga = g[2] * g[3];
            const double ii = ta_inter_corners(d[0], d[1], d[0] + d[2], d[1] + d[3],
                                               g[0], g[1], g[0] + g[2], g[1] + g[3]);
            const double uu = da + ga - ii;
            i += ii;
            u += uu;
            ratio_sum += (uu > 0.0) ? ii / uu : 0.0;
            if (ii > 0.5 * uu) matched++;
            ++a; ++b;
        } else if (sa < sb) {
            u += db[4 * a + 2] * db[4 * a + 3];
            ++a;
        } else {
            u += gb[4 * b + 2] * gb[4 * b + 3];
            ++b;
        }
        ++total;
    }
    if (mode == TA_IOU_AVG) return total ? ratio_sum / (double)total : 0.0;
    if (mode == TA_IOU_IMAGENETVID) return total ? (double)matched / (double)total : 0.0;
    if (!(i <= u)) *assert_failed = 1;            // eval.py:95
    return (u > 0.0) ? i / u : 0.0;
}

// eval.py:349-368 (TaoEval) / lvis_amodal/eval.py:202-217: GT `_ignore` flag of one range cfg.
TA_HD uint8_t ta_gt_ignored(const ta_range_cfg& c, double a, double b, int32_t hp, uint8_t flag) {
    const bool ig = (flag & 1) || (a < c.gt_a_lo) || (a > c.gt_a_hi) || (b < c.gt_b_lo) ||
                    (b > c.gt_b_hi) || (hp < c.gt_hp_min) || (c.gt_need_oof && !(flag & 2));
    return ig ? 1 : 0;
}

// eval.py:432-439 / lvis_amodal/eval.py:281-286: mask applied to UNMATCHED detections.
TA_HD bool ta_dt_unmatched_ignored(const ta_range_cfg& c, double a, double b, uint8_t flag) {
    return (a < c.dt_a_lo) || (a > c.dt_a_hi) || (b < c.dt_b_lo) || (b > c.dt_b_hi) || (flag & 1);
}

// Best GT for one detection under one matcher state (eval.py:400-417).  The reference walks
// GTs sorted "ignored last" (stable) and stops at the first free ignored GT once a regular
// GT is held; that is two passes over the original order: regular GTs, then — only if none
// matched — ignored GTs.  Later GTs win IoU ties (`<` at :413).
TA_HD int ta_match_one(const double* iou_row, int G, const uint8_t* gt_ig,
                       const uint32_t* taken, int taken_stride, double thr) {
    double best = (thr < 1.0 - 1e-10) ? thr : 1.0 - 1e-10;   // min([iou_thr, 1 - 1e-10])
    int m = -1;
    for (int pass = 0; pass < 2; ++pass) {
        for (int g = 0; g < G; ++g) {
            if (gt_ig[g] != pass) continue;
            if (taken[(g >> 5) * taken_stride] & (1u << (g & 31))) continue;
            const double v = iou_row[g];
            if (v < best) continue;
            best = v;
            m = g;
        }
        if (m >= 0) break;
    }
    return m;
}

// Smallest TP count t in [0, ngt] with t / ngt >= r (fp64 division as in eval.py:543,561);
// ngt + 1 when no count reaches r.
TA_HD int64_t ta_min_tp_for_recall(double r, int32_t ngt) {
    // 32-bit counts (ngt + 1 <= 2^31 fits): one-instruction int -> double conversions on the GPU
    const uint32_t top = (uint32_t)ngt + 1u;
    const double n = (double)ngt;
    const double g = r * n;
    uint32_t t = (g <= 0.0) ? 0u : ((g >= n + 1.0) ? top : (uint32_t)g);
    if (t > top) t = top;
    while (t > 0u && ((double)(t - 1u) / n) >= r) --t;
    while (t < top && ((double)t / n) < r) ++t;
    return (int64_t)t;
}

// eval.py:550: tp / (fp + tp + np.spacing(1))
TA_HD double ta_precision_at(int64_t tp, int64_t fp) {
    const double t = (double)tp, f = (double)fp;
    return t / ((f + t) + 2.220446049250313e-16);
}

// ---- precision / recall accumulation helpers (ta_pr.cu; host build: tests/hostsim) ----------

#ifdef __CUDA_ARCH__
#define TA_CLZ(x) __clz((int)(x))
#define TA_POPC(x) __popc(x)
#else
#define TA_CLZ(x) ((x) ? __builtin_clz(x) : 32)
#define TA_POPC(x) __builtin_popcount(x)
#endif

// exact comparison of precisions tp/(fp + tp + eps) given as (t, n = tp + fp): cross products in
// 64 bits; equal ratios prefer the larger n, which only matters for (1, 1): its rounded value
// 1/(1 + 2^-52) is below k/k = 1.0 (for n >= 2 the eps term vanishes in fp64, eval.py:550)
TA_HD bool pr_better(uint32_t t1, uint32_t n1, uint32_t t2, uint32_t n2) {
    const unsigned long long x = (unsigned long long)t1 * n2, y = (unsigned long long)t2 * n1;
    return x > y || (x == y && n1 > n2);
}
// packed candidate: t (24 bits) | n (24 bits) | chunk index inside the category (16 bits);
// ta_pr_accumulate rejects categories with 2^24 or more detections
TA_HD unsigned long long pr_pack(uint32_t t, uint32_t n, uint32_t ch) {
    return ((unsigned long long)t << 40) | ((unsigned long long)n << 16) | ch;
}
TA_HD void pr_unpack(unsigned long long q, uint32_t& t, uint32_t& n, uint32_t& ch) {
    t = (uint32_t)(q >> 40); n = (uint32_t)(q >> 16) & 0xffffffu; ch = (uint32_t)q & 0xffffu;
}

// One stage of the 32 x 32 bit-matrix transpose across a warp (rows = lanes, columns = bits):
// `y` is the word of lane ^ j.  After the stages j = 16, 8, 4, 2, 1 lane b holds bit b of
// every lane's input word (bit l of the result <-> input lane l), i.e. 32 ballots at once.
TA_HD void pr_transpose_consts(int lane, int j, uint32_t& keep, uint32_t& rot) {
    const uint32_t m = (j == 16) ? 0x0000ffffu : (j == 8) ? 0x00ff00ffu : (j == 4) ? 0x0f0f0f0fu
                     : (j == 2) ? 0x33333333u : 0x55555555u;
    // lower half of a 2j block keeps its columns m and receives the partner's columns m moved
    // up by j; the upper half keeps ~m and receives the partner's columns ~m moved down by j.
    // Both moves are one rotation (the wrapped-around bits fall outside the receiving mask), so
    // a stage is SHFL + SHF.W + LOP3 with `keep` / `rot` fixed per lane.
    const bool hi = (lane & j) != 0;
    keep = hi ? ~m : m;
    rot = hi ? (uint32_t)(32 - j) : (uint32_t)j;
}
TA_HD uint32_t pr_transpose_apply(uint32_t x, uint32_t y, uint32_t keep, uint32_t rot) {
    const uint32_t yr = (y << rot) | (y >> (32u - rot));
    return yr ^ ((x ^ yr) & keep);            // keep ? x : yr, bitwise
}
TA_HD uint32_t pr_transpose_stage(uint32_t x, uint32_t y, int lane, int j) {
    uint32_t keep, rot;
    pr_transpose_consts(lane, j, keep, rot);
    return pr_transpose_apply(x, y, keep, rot);
}

#define TA_PR_WORDS 8          // 32-position words per chunk of 256 detections

// Compact per-detection result word of ta_frame_eval (include/ta_eval.h):
//   w = M | A << T | B << (T + C) | U << (T + 2 C),  bit 31 = "use the full row instead".
// TP/FP row entry of range cfg `cfg`: (A_cfg ? M : 0) | ((B_cfg ? M : 0) | (U_cfg ? ~M : 0)) << 16.
TA_HD uint32_t pr_expand(uint32_t w, int cfg, int n_thr, int n_cfg) {
    const uint32_t thr_all = (1u << n_thr) - 1u;
    const uint32_t M = w & thr_all;
    const uint32_t s = w >> (n_thr + cfg);
    const uint32_t tp = (s & 1u) ? M : 0u;
    const uint32_t fp = (((s >> n_cfg) & 1u) ? M : 0u) | (((s >> (2 * n_cfg)) & 1u) ? (thr_all & ~M) : 0u);
    return tp | (fp << 16);
}
// The same per (cfg, threshold) CELL over 32 detections at once: Mk / A / B / U are the ballots
// of the word bits k / T + cfg / T + C + cfg / T + 2 C + cfg (one 32 x 32 bit transpose of the
// compact words gives all of them).  Lanes whose word is 0 contribute nothing.
TA_HD void pr_cell_planes(uint32_t Mk, uint32_t A, uint32_t B, uint32_t U, uint32_t& tp, uint32_t& fp) {
    tp = A & Mk;
    fp = (B & Mk) | (U & ~Mk);
}

// Envelope walk of one (range cfg, threshold) cell over the TRUE POSITIVES of its chunks, last
// chunk first (accumulate, eval.py:527-573).  The state carries the running TP / FP counts, the
// suffix-maximum precision as the exact rational (bt, bn) and the next recall level to answer.
//   ta_pr_state_init   counts at the END of the last chunk to be walked; finds the last recall
//                      level whose tk is already reached (tk[0 .. n_rec) non-decreasing:
//                      ta_min_tp_for_recall of ascending recall thresholds)
//   ta_pr_walk_chunk   w[0..8) / w[8..16): the cell's TP / FP flags of the chunk's positions
//                      32 j .. 32 j + 31 in descending-score order (bit = position inside the
//                      word).  Stores the suffix maximum (tagged with the chunk) for every recall
//                      level k whose tk-th true positive lies in the chunk: q[k].
struct ta_pr_state {
    uint32_t tc, fc, bt, bn, next_tk;
    uint32_t after_tk;      // the level after next_tk, loaded one event ahead (breaks the chain
                            // store -> dependent load -> compare of consecutive recall levels)
    int kq;
    uint32_t nf;            // the next counted detection ABOVE the walk position is a false positive
                            // (or the walk started there): a true positive met now ends a TP run
};
TA_HD uint32_t ta_pr_tk_at(const int32_t* tk, int k) {
    if (k < 0) return 0u;
    const int32_t v = tk[k];
    return (uint32_t)(v > 1 ? v : 1);
}
TA_HD void ta_pr_state_init(ta_pr_state& s, uint32_t tc, uint32_t fc, const int32_t* tk, int n_rec) {
    s.tc = tc; s.fc = fc;
    s.bt = 0; s.bn = 1;                  // precision 0: the first true positive always beats it
    s.nf = 1;
    int lo = -1, hi = n_rec;             // invariant: tk'[lo] <= tc < tk'[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        const int32_t v = tk[mid];
        if ((uint32_t)(v > 1 ? v : 1) <= tc) lo = mid; else hi = mid;
    }
    s.kq = lo;
    s.next_tk = ta_pr_tk_at(tk, lo);
    s.after_tk = ta_pr_tk_at(tk, lo - 1);
}
// A chunk without true positives of the cell: only the false-positive count moves.
TA_HD void ta_pr_skip_chunk(ta_pr_state& s, uint32_t fc_before) {
    if (s.fc != fc_before) s.nf = 1;
    s.fc = fc_before;
}
// The walk visits only the true positives that END a run of true positives (the next counted
// detection after them is a false positive, or the walk's start): with no false positive in
// between, precision tc / (tc + fc) only grows from one true positive to the next, so the suffix
// maximum — and, with the strict comparison below, the very pair (bt, bn) a walk over all true
// positives keeps — is decided at run ends.  Inside a word the run ends are found at once: adding
// t << 1 to the complement of m = t | f carries every true positive to the next counted
// detection, so ((~m + (t << 1)) & m) & f are the false positives that follow a true positive.
// A recall level is answered lazily: before a candidate with fewer true positives than the level
// needs is merged (it lies before the level's tk-th true positive), and at the chunk's end.
#define TA_PR_CANDIDATE(c_, n_)                                                              \
    do {                                                                                     \
        const uint32_t cc_ = (c_), nn_ = (n_);                                               \
        if (next_tk > cc_) {                                                                 \
            const unsigned long long packed = pr_pack(bt, bn, ch_rel);                       \
            do {                                                                             \
                q[kq * q_stride] = packed;                                                   \
                --kq;                                                                        \
                next_tk = after_tk;                                                          \
                after_tk = ta_pr_tk_at(tk, kq - 1);                                          \
            } while (next_tk > cc_);                                                         \
        }                                                                                    \
        /* strict ">" is enough: a tie keeps the later detection's pair, and the only       \
           value-changing tie, (1,1) vs (k,k), has (1,1) as the candidate (first TP overall) */ \
        if ((unsigned long long)cc_ * bn > (unsigned long long)bt * nn_) { bt = cc_; bn = nn_; } \
    } while (0)
TA_HD void ta_pr_walk_chunk(ta_pr_state& s, const uint32_t* w, const int32_t* tk, uint32_t ch_rel,
                            unsigned long long* q, int64_t q_stride) {
    uint32_t tc = s.tc, fc = s.fc, bt = s.bt, bn = s.bn, next_tk = s.next_tk, after_tk = s.after_tk;
    uint32_t nf = s.nf;
    int kq = s.kq;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = TA_PR_WORDS - 1; j >= 0; --j) {
        const uint32_t t = w[j];
        const uint32_t f = w[TA_PR_WORDS + j];
        const uint32_t m = t | f;
        if (m == 0) continue;
        // the word's last counted detection is a true positive and what follows it is not
        if (nf && (t >> (31 - TA_CLZ(m)))) TA_PR_CANDIDATE(tc, tc + fc);
        uint32_t cand = ((~m + (t << 1)) & m) & f;
        while (cand) {
            const int p = 31 - TA_CLZ(cand);
            cand ^= 1u << p;
            // counts just below the false positive at p = the run end's own counts
            const uint32_t c = tc - (uint32_t)TA_POPC(t >> p);
            TA_PR_CANDIDATE(c, c + (fc - (uint32_t)TA_POPC(f >> p)));
        }
        tc -= (uint32_t)TA_POPC(t);
        fc -= (uint32_t)TA_POPC(f);
        nf = (f & (m & (0u - m))) != 0u;
    }
    // recall levels whose tk-th true positive lies in this chunk and before its first run end
    if (next_tk > tc) {
        const unsigned long long packed = pr_pack(bt, bn, ch_rel);
        do {
            q[kq * q_stride] = packed;
            --kq;
            next_tk = after_tk;
            after_tk = ta_pr_tk_at(tk, kq - 1);
        } while (next_tk > tc);
    }
    s.tc = tc; s.fc = fc; s.bt = bt; s.bn = bn; s.next_tk = next_tk; s.after_tk = after_tk; s.kq = kq;
    s.nf = nf;
}
// One chunk on its own (tests/hostsim; the kernel carries the state over several chunks)
TA_HD unsigned long long ta_pr_walk_bits(const uint32_t* T, const uint32_t* F, int64_t stride,
                                         uint32_t tc, uint32_t fc, const int32_t* tk, int n_rec,
                                         uint32_t ch_rel, unsigned long long* q, int64_t q_stride) {
    uint32_t w[2 * TA_PR_WORDS];
    for (int j = 0; j < TA_PR_WORDS; ++j) { w[j] = T[j * stride]; w[TA_PR_WORDS + j] = F[j * stride]; }
    ta_pr_state s;
    ta_pr_state_init(s, tc, fc, tk, n_rec);
    ta_pr_walk_chunk(s, w, tk, ch_rel, q, q_stride);
    return pr_pack(s.bt, s.bn, 0);
}

// ---- mask IoU of the segm path (ta_rle.cu; host build: tests/hostsim) -----------------------

// IoU of two column-major run-length masks, iscrowd = 0: pycocotools rleIou for one pair
// (in-tree copy visualization/tao/third_party/pysot/training_dataset/coco/pycocotools/common/
// maskApi.c:78-96).  db / gb are the masks' boxes as rleToBbox derives them (:133-151): the
// pair is only walked when their box IoU is positive (:80-82), masks of different size give -1
// (:85).  Both run lists are consumed in lock step; zero-length runs toggle like the C loop.
TA_HD double ta_rle_pair_iou(const uint32_t* cd, int64_t md, const uint32_t* cg, int64_t mg,
                             const double* db, const double* gb,
                             uint32_t dh, uint32_t dw, uint32_t gh, uint32_t gw) {
    const double o = ta_bb_iou(db[0], db[1], db[2], db[3], gb[0], gb[1], gb[2], gb[3]);
    if (!(o > 0.0)) return o;
    if (dh != gh || dw != gw) return -1.0;
    uint32_t ca = md > 0 ? cd[0] : 0u, cb = mg > 0 ? cg[0] : 0u;
    int64_t a = 1, b = 1;
    bool va = false, vb = false;
    uint32_t i = 0, u = 0, ct = 1;
    while (ct > 0) {
        const uint32_t c = ca < cb ? ca : cb;
        if (va || vb) { u += c; if (va && vb) i += c; }
        ct = 0;
        ca -= c; if (!ca && a < md) { ca = cd[a++]; va = !va; } ct += ca;
        cb -= c; if (!cb && b < mg) { cb = cg[b++]; vb = !vb; } ct += cb;
    }
    if (i == 0) u = 1;
    return (double)i / (double)u;
}

#endif  // TA_DEVICE_FNS_CUH
