// TEST ARTEFACT — host build of the per-thread kernel arithmetic (csrc/ta_device_fns.cuh).
//
// Serial loops that mirror the work decomposition of the kernels in csrc/ta_eval.cu so the
// algorithms (two-pass greedy matching, union re-association, bucketed PR interpolation) can
// be checked against the oracle and the reference goldens on a machine without a GPU.
// Never linked into, loaded by, or shipped with the product library.
#include <stdint.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "ta_device_fns.cuh"

extern "C" {

int hs_track_iou(int mode, int64_t n_groups, const int64_t* grp_dt_off, const int64_t* grp_gt_off,
                 const int64_t* dt_off, const double* dt_box, const int32_t* dt_slot,
                 const int64_t* gt_off, const double* gt_box, const int32_t* gt_slot,
                 const int64_t* iou_off, double* iou) {
    int bad_total = 0;
    for (int64_t grp = 0; grp < n_groups; ++grp) {
        const int64_t d0 = grp_dt_off[grp], g0 = grp_gt_off[grp];
        const int D = (int)(grp_dt_off[grp + 1] - d0), G = (int)(grp_gt_off[grp + 1] - g0);
        double* out = iou + iou_off[grp];
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < G; ++j) {
                const int64_t db0 = dt_off[d0 + i], db1 = dt_off[d0 + i + 1];
                const int64_t gb0 = gt_off[g0 + j], gb1 = gt_off[g0 + j + 1];
                if (mode == TA_IOU_3D) {
                    // k_track_iou_tiled: I over common slots, U = DA + GA - I
                    double da = 0, ga = 0, inter = 0;
                    for (int64_t k = db0; k < db1; ++k) da += dt_box[4 * k + 2] * dt_box[4 * k + 3];
                    for (int64_t k = gb0; k < gb1; ++k) ga += gt_box[4 * k + 2] * gt_box[4 * k + 3];
                    int64_t a = db0, b = gb0;
                    while (a < db1 && b < gb1) {
                        if (dt_slot[a] == gt_slot[b]) {
                            const double* d = dt_box + 4 * a; const double* g = gt_box + 4 * b;
                            inter += ta_inter_corners(d[0], d[1], d[0] + d[2], d[1] + d[3],
                                                      g[0], g[1], g[0] + g[2], g[1] + g[3]);
                            ++a; ++b;
                        } else if (dt_slot[a] < gt_slot[b]) ++a; else ++b;
                    }
                    const double uni = (da + ga) - inter;
                    out[(int64_t)i * G + j] = uni > 0.0 ? inter / uni : 0.0;
                } else {
                    int bad = 0;
                    out[(int64_t)i * G + j] = ta_pair_iou_merge(
                        dt_box + 4 * db0, dt_slot + db0, (int)(db1 - db0),
                        gt_box + 4 * gb0, gt_slot + gb0, (int)(gb1 - gb0), mode, &bad);
                    bad_total += bad;
                }
            }
    }
    return bad_total;
}

int hs_box_iou(int64_t n_groups, const int64_t* grp_dt_off, const int64_t* grp_gt_off,
               const double* dt_box, const double* gt_box, const int64_t* iou_off, double* iou) {
    for (int64_t grp = 0; grp < n_groups; ++grp) {
        const int64_t d0 = grp_dt_off[grp], g0 = grp_gt_off[grp];
        const int D = (int)(grp_dt_off[grp + 1] - d0), G = (int)(grp_gt_off[grp + 1] - g0);
        double* out = iou + iou_off[grp];
        for (int d = 0; d < D; ++d)
            for (int g = 0; g < G; ++g) {
                const double* a = dt_box + 4 * (d0 + d); const double* b = gt_box + 4 * (g0 + g);
                out[(int64_t)d * G + g] = ta_bb_iou(a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3]);
            }
    }
    return 0;
}

int hs_match_greedy(int64_t n_groups, const int64_t* grp_dt_off, const int64_t* grp_gt_off,
                    const int32_t* grp_cat, const int64_t* iou_off, const double* iou,
                    int32_t n_thr, const double* thrs, int32_t n_cfg, const ta_range_cfg* cfgs,
                    int64_t n_dt, const double* dt_a, const double* dt_b, const uint8_t* dt_flag,
                    int64_t n_gt, const double* gt_a, const double* gt_b,
                    const int32_t* gt_hp, const uint8_t* gt_flag,
                    int32_t g_max, uint32_t* dt_tpfp, int32_t* num_gt,
                    int32_t* dt_match_gt, uint8_t* gt_ignore_out) {
    (void)g_max;
    for (int64_t grp = 0; grp < n_groups; ++grp) {
        const int64_t d0 = grp_dt_off[grp], g0 = grp_gt_off[grp];
        const int D = (int)(grp_dt_off[grp + 1] - d0), G = (int)(grp_gt_off[grp + 1] - g0);
        if (D == 0 && G == 0) continue;
        const int words = (G + 31) / 32;
        for (int c = 0; c < n_cfg; ++c) {
            std::vector<uint8_t> ig(G);
            int cnt = 0;
            for (int g = 0; g < G; ++g) {
                ig[g] = ta_gt_ignored(cfgs[c], gt_a[g0 + g], gt_b[g0 + g], gt_hp[g0 + g], gt_flag[g0 + g]);
                cnt += ig[g] == 0;
                if (gt_ignore_out) gt_ignore_out[(int64_t)c * n_gt + g0 + g] = ig[g];
            }
            num_gt[(int64_t)grp_cat[grp] * n_cfg + c] += cnt;
            for (int d = 0; d < D; ++d) dt_tpfp[(d0 + d) * n_cfg + c] = 0;
            for (int t = 0; t < n_thr; ++t) {
                std::vector<uint32_t> taken(words ? words : 1, 0u);
                for (int d = 0; d < D; ++d) {
                    int m = -1;
                    if (G > 0) m = ta_match_one(iou + iou_off[grp] + (int64_t)d * G, G, ig.data(),
                                                taken.data(), 1, thrs[t]);
                    const uint8_t dfl = dt_flag[d0 + d];
                    bool unmatched = true, ign = false;
                    if (m >= 0) {
                        if (dfl & 2) taken[m >> 5] |= 1u << (m & 31);
                        unmatched = (gt_flag[g0 + m] & 4) != 0;
                        ign = ig[m] != 0;
                    }
                    if (unmatched && !ign)
                        ign = ta_dt_unmatched_ignored(cfgs[c], dt_a[d0 + d], dt_b[d0 + d], dt_flag[d0 + d]);
                    if (!ign) dt_tpfp[(d0 + d) * n_cfg + c] |= unmatched ? (1u << (16 + t)) : (1u << t);
                    if (dt_match_gt) dt_match_gt[((int64_t)c * n_thr + t) * n_dt + d0 + d] = m;
                }
            }
        }
    }
    return 0;
}

int hs_pr_accumulate(int32_t n_cat, const int64_t* cat_dt_off, const int32_t* acc_perm, int64_t n_dt,
                     const uint32_t* dt_tpfp, const int32_t* num_gt, int32_t n_thr, int32_t n_cfg,
                     int32_t n_rec, const double* rec_thrs, double* precision, double* recall,
                     int64_t* tp_cnt, int64_t* fp_cnt) {
    std::vector<int64_t> tk(n_rec);
    std::vector<double> bucket(n_rec);
    for (int c = 0; c < n_cat; ++c)
        for (int cfg = 0; cfg < n_cfg; ++cfg)
            for (int t = 0; t < n_thr; ++t) {
                const int ngt = num_gt[(int64_t)c * n_cfg + cfg];
                const int64_t cell = ((int64_t)t * n_cat + c) * n_cfg + cfg;
                if (ngt == 0) {
                    for (int k = 0; k < n_rec; ++k)
                        precision[(((int64_t)t * n_rec + k) * n_cat + c) * n_cfg + cfg] = -1.0;
                    recall[cell] = -1.0;
                    if (tp_cnt) tp_cnt[cell] = 0;
                    if (fp_cnt) fp_cnt[cell] = 0;
                    continue;
                }
                for (int k = 0; k < n_rec; ++k) { tk[k] = ta_min_tp_for_recall(rec_thrs[k], ngt); bucket[k] = 0.0; }
                int64_t tp = 0, fp = 0;
                for (int64_t p = cat_dt_off[c]; p < cat_dt_off[c + 1]; ++p) {
                    const uint32_t w = dt_tpfp[(int64_t)acc_perm[p] * n_cfg + cfg];
                    if ((w >> t) & 1u) {
                        ++tp;
                        const double pr = ta_precision_at(tp, fp);
                        const int lo = (int)(std::upper_bound(tk.begin(), tk.end(), tp) - tk.begin());
                        if (lo > 0 && pr > bucket[lo - 1]) bucket[lo - 1] = pr;
                    } else if ((w >> (16 + t)) & 1u) {
                        ++fp;
                    }
                }
                double best = 0.0;
                for (int k = n_rec - 1; k >= 0; --k) {
                    if (bucket[k] > best) best = bucket[k];
                    precision[(((int64_t)t * n_rec + k) * n_cat + c) * n_cfg + cfg] = best;
                }
                recall[cell] = (cat_dt_off[c + 1] > cat_dt_off[c]) ? (double)tp / (double)ngt : 0.0;
                if (tp_cnt) tp_cnt[cell] = tp;
                if (fp_cnt) fp_cnt[cell] = fp;
            }
    return 0;
}

// Serial emulation of the bit-plane PR pipeline of csrc/ta_pr.cu (k_pr_plan -> k_pr_bits ->
// k_pr_scan -> k_pr_envelope_bits -> k_pr_suffix -> k_pr_finalize) with the SAME per-thread
// functions (pr_transpose_stage on an emulated 32-lane warp, ta_pr_walk_bits, pr_better, ...).
static int pr_bits_impl(int32_t n_cat, const int64_t* cat_dt_off, const int32_t* acc_perm, int64_t n_dt,
                          const uint32_t* dt_tpfp, const int32_t* num_gt, int32_t n_thr, int32_t n_cfg,
                          int32_t n_rec, const double* rec_thrs, double* precision, double* recall,
                          int64_t* tp_cnt, int64_t* fp_cnt, int mode) {
    (void)n_dt;
    const bool rows = mode >= 1;          // cell-major answers (the shipped layout)
    const int CH = 32 * TA_PR_WORDS;
    const int n_cells = n_cfg * n_thr;
    std::vector<int> chunk_start(n_cat + 1, 0);
    for (int c = 0; c < n_cat; ++c)
        chunk_start[c + 1] = chunk_start[c] + (int)((cat_dt_off[c + 1] - cat_dt_off[c] + CH - 1) / CH);
    const int n_chunks = chunk_start[n_cat];
    std::vector<int> chunk_cat(n_chunks);
    std::vector<uint32_t> bits((size_t)n_chunks * 2 * TA_PR_WORDS * n_cells, 0u);
    std::vector<uint32_t> chunk_cnt((size_t)n_chunks * n_cfg * 32, 0u), cat_tot((size_t)n_cat * n_cfg * 32, 0u);
    std::vector<int32_t> tk((size_t)n_cat * n_cfg * n_rec);
    std::vector<unsigned long long> chunk_best((size_t)n_chunks * n_cells, 0ull);
    const int64_t per_t = (int64_t)n_cat * n_cfg;
    std::vector<unsigned long long> prec_bits((size_t)n_thr * n_rec * per_t, 0xdeadbeefdeadbeefull);
    // k_pr_bits
    for (int cat = 0; cat < n_cat; ++cat)
        for (int chunk = chunk_start[cat]; chunk < chunk_start[cat + 1]; ++chunk) {
            chunk_cat[chunk] = cat;
            const int64_t p0 = cat_dt_off[cat] + (int64_t)(chunk - chunk_start[cat]) * CH;
            const int n_pos = (int)std::min<int64_t>(CH, cat_dt_off[cat + 1] - p0);
            uint32_t* out = bits.data() + (size_t)chunk * 2 * TA_PR_WORDS * n_cells;
            for (int warp = 0; warp < TA_PR_WORDS; ++warp)
                for (int cfg = 0; cfg < n_cfg; ++cfg) {
                    uint32_t x[32], y[32];
                    for (int lane = 0; lane < 32; ++lane) {
                        const int p = warp * 32 + lane;
                        x[lane] = p < n_pos ? dt_tpfp[(int64_t)acc_perm[p0 + p] * n_cfg + cfg] : 0u;
                    }
                    for (int j = 16; j >= 1; j >>= 1) {
                        for (int lane = 0; lane < 32; ++lane) y[lane] = x[lane ^ j];     // shfl_xor
                        for (int lane = 0; lane < 32; ++lane) x[lane] = pr_transpose_stage(x[lane], y[lane], lane, j);
                    }
                    for (int lane = 0; lane < 32; ++lane) {
                        const int b = lane & 15;
                        if (b < n_thr) out[((lane >> 4) * TA_PR_WORDS + warp) * n_cells + cfg * n_thr + b] = x[lane];
                    }
                }
            for (int j = 0; j < n_cfg * 32; ++j) {
                const int cfg = j >> 5, bit = j & 31, t = bit & 15;
                uint32_t cnt = 0;
                if (t < n_thr)
                    for (int u = 0; u < TA_PR_WORDS; ++u)
                        cnt += (uint32_t)__builtin_popcount(out[((bit >> 4) * TA_PR_WORDS + u) * n_cells + cfg * n_thr + t]);
                chunk_cnt[((size_t)chunk * n_cfg + cfg) * 32 + bit] = cnt;
            }
        }
    // k_pr_scan
    for (int cat = 0; cat < n_cat; ++cat) {
        const bool has_dt = cat_dt_off[cat + 1] > cat_dt_off[cat];
        for (int j = 0; j < n_cfg * 32; ++j) {
            const int cfg = j >> 5, bit = j & 31, t = bit & 15;
            uint32_t run = 0;
            for (int ch = chunk_start[cat]; ch < chunk_start[cat + 1]; ++ch) {
                uint32_t& q = chunk_cnt[((size_t)ch * n_cfg) * 32 + j];
                const uint32_t v = q; q = run; run += v;
            }
            cat_tot[((size_t)cat * n_cfg) * 32 + j] = run;
            const int ngt = num_gt[(int64_t)cat * n_cfg + cfg];
            if (t < n_thr) {
                const int64_t cell = ((int64_t)t * n_cat + cat) * n_cfg + cfg;
                if (bit < 16) {
                    if (tp_cnt) tp_cnt[cell] = ngt ? (int64_t)run : 0;
                    recall[cell] = ngt == 0 ? -1.0 : (has_dt ? (double)run / (double)ngt : 0.0);
                } else if (fp_cnt) fp_cnt[cell] = ngt ? (int64_t)run : 0;
            }
        }
        for (int cfg = 0; cfg < n_cfg; ++cfg)
            for (int k = 0; k < n_rec; ++k) {
                const int ngt = num_gt[(int64_t)cat * n_cfg + cfg];
                tk[((size_t)cat * n_cfg + cfg) * n_rec + k] = ngt ? (int32_t)ta_min_tp_for_recall(rec_thrs[k], ngt) : 0x7fffffff;
            }
    }
    // k_pr_envelope_bits: one walker per (run of `seg` consecutive chunks, cell), last chunk first,
    // the state carried from chunk to chunk inside a category (mode 3: seg = 4, as the kernel
    // does on large inputs)
    const int seg = mode == 3 ? 4 : 1;
    for (int c_lo = 0; c_lo < n_chunks; c_lo += seg)
        for (int cell = 0; cell < n_cells; ++cell) {
            const int cfg = cell / n_thr, b = cell % n_thr;
            const int c_hi = std::min(c_lo + seg, n_chunks) - 1;
            int cat = -1, ch0 = 0;
            bool live = false;
            ta_pr_state st;
            const int32_t* tkp = nullptr;
            unsigned long long* q = nullptr;
            for (int c = c_hi; c >= c_lo; --c) {
                if (chunk_cat[c] != cat) {
                    cat = chunk_cat[c];
                    const int64_t cc = (int64_t)cat * n_cfg + cfg;
                    live = num_gt[cc] != 0;
                    if (live) {
                        ch0 = chunk_start[cat];
                        const int ch1 = chunk_start[cat + 1];
                        const uint32_t* nxt = (c + 1 < ch1) ? &chunk_cnt[((size_t)(c + 1) * n_cfg + cfg) * 32]
                                                            : &cat_tot[((size_t)cat * n_cfg + cfg) * 32];
                        tkp = &tk[((size_t)cat * n_cfg + cfg) * n_rec];
                        // rows: cell-major answers (k_pr_envelope_bits with a.ans), else the precision layout
                        q = rows ? prec_bits.data() + ((int64_t)b * per_t + cc) * n_rec
                                 : prec_bits.data() + (int64_t)b * n_rec * per_t + cc;
                        ta_pr_state_init(st, nxt[b], nxt[16 + b], tkp, n_rec);
                    }
                }
                if (!live) continue;
                const uint32_t* cnt = &chunk_cnt[((size_t)c * n_cfg + cfg) * 32];
                if (st.tc == cnt[b]) {
                    ta_pr_skip_chunk(st, cnt[16 + b]);
                } else {
                    const uint32_t* planes = bits.data() + (size_t)c * 2 * TA_PR_WORDS * n_cells + cell;
                    uint32_t w[2 * TA_PR_WORDS];
                    for (int j = 0; j < 2 * TA_PR_WORDS; ++j) w[j] = planes[(size_t)j * n_cells];
                    ta_pr_walk_chunk(st, w, tkp, (uint32_t)(c - ch0), q, rows ? 1 : per_t);
                }
                chunk_best[(size_t)c * n_cells + cell] = st.bt ? pr_pack(st.bt, st.bn, 0) : 0ull;
            }
        }
    // k_pr_suffix
    for (int cat = 0; cat < n_cat; ++cat)
        for (int j = 0; j < n_cells; ++j) {
            if (num_gt[(int64_t)cat * n_cfg + j / n_thr] == 0) continue;
            uint32_t bt = 0, bn = 0, d;
            for (int ch = chunk_start[cat + 1] - 1; ch >= chunk_start[cat]; --ch) {
                unsigned long long& q = chunk_best[(size_t)ch * n_cells + j];
                uint32_t ct, cn;
                pr_unpack(q, ct, cn, d);
                q = pr_pack(bt, bn, 0);
                if (pr_better(ct, cn, bt, bn)) { bt = ct; bn = cn; }
            }
        }
    // k_pr_finalize
    for (int64_t idx = 0; idx < (int64_t)n_thr * n_rec * per_t; ++idx) {
        const int64_t tk_idx = idx / per_t, cc = idx - tk_idx * per_t;
        const int ngt = num_gt[cc];
        if (ngt == 0) { precision[idx] = -1.0; continue; }
        const int t = (int)tk_idx / n_rec, k = (int)tk_idx - t * n_rec;
        const int cat = (int)(cc / n_cfg), cfg = (int)(cc - (int64_t)cat * n_cfg);
        const int32_t tkv = tk[cc * n_rec + k];
        const uint32_t need = (uint32_t)(tkv > 1 ? tkv : 1);
        if (need > cat_tot[cc * 32 + t]) { precision[idx] = 0.0; continue; }
        uint32_t qt, qn, ch, bt, bn, d;
        // mode 2 = k_pr_finalize_tile: per-entry finalize reading the cell-major answers
        pr_unpack(rows ? prec_bits[((int64_t)t * per_t + cc) * n_rec + k] : prec_bits[idx], qt, qn, ch);
        pr_unpack(chunk_best[((size_t)(chunk_start[cat] + ch) * n_cfg + cfg) * n_thr + t], bt, bn, d);
        if (pr_better(bt, bn, qt, qn)) { qt = bt; qn = bn; }
        precision[idx] = ta_precision_at((int64_t)qt, (int64_t)(qn - qt));
    }
    return 0;
}

int hs_pr_accumulate_bits(int32_t n_cat, const int64_t* cat_dt_off, const int32_t* acc_perm, int64_t n_dt,
                          const uint32_t* dt_tpfp, const int32_t* num_gt, int32_t n_thr, int32_t n_cfg,
                          int32_t n_rec, const double* rec_thrs, double* precision, double* recall,
                          int64_t* tp_cnt, int64_t* fp_cnt) {
    return pr_bits_impl(n_cat, cat_dt_off, acc_perm, n_dt, dt_tpfp, num_gt, n_thr, n_cfg, n_rec, rec_thrs,
                        precision, recall, tp_cnt, fp_cnt, 0);
}
int hs_pr_accumulate_bits_tile(int32_t n_cat, const int64_t* cat_dt_off, const int32_t* acc_perm, int64_t n_dt,
                               const uint32_t* dt_tpfp, const int32_t* num_gt, int32_t n_thr, int32_t n_cfg,
                               int32_t n_rec, const double* rec_thrs, double* precision, double* recall,
                               int64_t* tp_cnt, int64_t* fp_cnt) {
    return pr_bits_impl(n_cat, cat_dt_off, acc_perm, n_dt, dt_tpfp, num_gt, n_thr, n_cfg, n_rec, rec_thrs,
                        precision, recall, tp_cnt, fp_cnt, 2);
}

int hs_pr_accumulate_bits_seg(int32_t n_cat, const int64_t* cat_dt_off, const int32_t* acc_perm, int64_t n_dt,
                              const uint32_t* dt_tpfp, const int32_t* num_gt, int32_t n_thr, int32_t n_cfg,
                              int32_t n_rec, const double* rec_thrs, double* precision, double* recall,
                              int64_t* tp_cnt, int64_t* fp_cnt) {
    return pr_bits_impl(n_cat, cat_dt_off, acc_perm, n_dt, dt_tpfp, num_gt, n_thr, n_cfg, n_rec, rec_thrs,
                        precision, recall, tp_cnt, fp_cnt, 3);
}

void hs_transpose32(const uint32_t* in, uint32_t* out) {
    uint32_t x[32], y[32];
    for (int l = 0; l < 32; ++l) x[l] = in[l];
    for (int j = 16; j >= 1; j >>= 1) {
        for (int l = 0; l < 32; ++l) y[l] = x[l ^ j];
        for (int l = 0; l < 32; ++l) x[l] = pr_transpose_stage(x[l], y[l], l, j);
    }
    for (int l = 0; l < 32; ++l) out[l] = x[l];
}

int hs_rle_iou(int64_t n_groups, const int64_t* grp_dt_off, const int64_t* grp_gt_off,
               const int64_t* dt_off, const uint32_t* dt_cnt, const uint32_t* dt_hw, const double* dt_bb,
               const int64_t* gt_off, const uint32_t* gt_cnt, const uint32_t* gt_hw, const double* gt_bb,
               const int64_t* iou_off, double* iou) {
    for (int64_t grp = 0; grp < n_groups; ++grp) {
        const int64_t d0 = grp_dt_off[grp], g0 = grp_gt_off[grp];
        const int64_t D = grp_dt_off[grp + 1] - d0, G = grp_gt_off[grp + 1] - g0;
        for (int64_t e = 0; e < D * G; ++e) {
            const int64_t d = d0 + e / G, g = g0 + e % G;
            iou[iou_off[grp] + e] = ta_rle_pair_iou(
                dt_cnt + dt_off[d], dt_off[d + 1] - dt_off[d], gt_cnt + gt_off[g], gt_off[g + 1] - gt_off[g],
                dt_bb + 4 * d, gt_bb + 4 * g, dt_hw[2 * d], dt_hw[2 * d + 1], gt_hw[2 * g], gt_hw[2 * g + 1]);
        }
    }
    return 0;
}

// Plane words of one warp (32 detections) from COMPACT result words, the way k_pr_bits builds
// them (one 32 x 32 bit transpose + pr_cell_planes), next to the per-cfg expansion + transpose
// of the full-row path.  tp / fp: [n_cfg * n_thr] words each.  Returns 0 when both agree.
int hs_pr_compact_planes(const uint32_t* words32, int n_thr, int n_cfg, uint32_t* tp, uint32_t* fp) {
    uint32_t x[32], y[32];
    for (int l = 0; l < 32; ++l) x[l] = (words32[l] >> 31) ? 0u : words32[l];
    for (int j = 16; j >= 1; j >>= 1) {
        for (int l = 0; l < 32; ++l) y[l] = x[l ^ j];
        for (int l = 0; l < 32; ++l) x[l] = pr_transpose_stage(x[l], y[l], l, j);
    }
    int bad = 0;
    for (int cfg = 0; cfg < n_cfg; ++cfg) {
        // reference: expand every lane's word for this cfg and transpose
        uint32_t r[32], q[32];
        for (int l = 0; l < 32; ++l) r[l] = (words32[l] >> 31) ? 0u : pr_expand(words32[l], cfg, n_thr, n_cfg);
        for (int j = 16; j >= 1; j >>= 1) {
            for (int l = 0; l < 32; ++l) q[l] = r[l ^ j];
            for (int l = 0; l < 32; ++l) r[l] = pr_transpose_stage(r[l], q[l], l, j);
        }
        for (int k = 0; k < n_thr; ++k) {
            uint32_t a, b;
            pr_cell_planes(x[k], x[n_thr + cfg], x[n_thr + n_cfg + cfg], x[n_thr + 2 * n_cfg + cfg], a, b);
            tp[cfg * n_thr + k] = a;
            fp[cfg * n_thr + k] = b;
            if (a != r[k] || b != r[16 + k]) ++bad;
        }
    }
    return bad;
}

}  // extern "C"
