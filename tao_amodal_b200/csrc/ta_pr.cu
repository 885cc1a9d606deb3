// ta_pr.cu — precision / recall accumulation (sm_100a).
//
// Reference: TaoEval.accumulate (tao_amodal/evaluation/tao_amodal/eval.py:459-584) and
// LVISEval.accumulate (lvis_amodal/eval.py:305-426).  Per (category, range cfg, threshold) the
// reference walks the category's detections in descending-score order, forms running TP / FP
// counts, precision = tp / (fp + tp + eps), takes the suffix maximum ("envelope", :557-559) and
// samples it at the first position whose recall reaches each of the 101 recall thresholds
// (:561-571).  Equivalent formulation used here: the value at recall threshold k is the maximum
// precision over all TRUE POSITIVES whose running TP count is >= tk[k], where tk[k] is the
// smallest count with count / num_gt >= rec_thrs[k] (a false positive never exceeds the
// precision of the true positive before it).
//
// The category lists are cut into chunks of PR_CHUNK detections so that long categories do not
// serialise.  Precisions are compared as exact rationals (tp, tp + fp) — rounding is monotone, so
// the maximum of the rounded quotients is the rounded quotient of the rational maximum — and
// only the <= 101 winners per cell are divided (ta_precision_at).
//
// Shipped kernel set (TA_PR_IMPL=1, default):
//   k_pr_plan      chunk table: first chunk of every category
//   k_pr_bits      the chunk's TP/FP words are transposed ONCE (32 x 32 bit-matrix transpose
//                  across the warp) into one 256-bit TP and FP plane per (cfg, threshold) cell;
//                  chunk totals are popcounts
//   k_pr_scan_live   per-category exclusive scan of the 2 n_thr live counters per cfg; recall
//                  and TP / FP totals; the category's tk table
//   k_pr_envelope_bits   one thread per (chunk, cell) visits only the TRUE POSITIVES of its plane
//                  (clz / popc) instead of all 256 positions: running counts, suffix-maximum
//                  precision, and the answer of every recall level whose tk-th true positive
//                  lies inside the chunk, stored cell-major (sequential per thread)
//   k_pr_suffix    per cell: best precision of all later chunks, for every chunk
//   k_pr_finalize_tile   merge the in-chunk answer with the later chunks' best, divide, write
//                  (-1 without GT, 0 for recall levels nobody reaches): coalesced cell-major
//                  reads, shared-memory transpose, coalesced writes
// Baseline kernel set (TA_PR_IMPL=0; the A/B reference, tools/ab_variants.py):
//   k_pr_count     chunk totals with bit-sliced carry-save counters per lane + REDUX
//   k_pr_scan      scan of all 32 counters per cfg + the tk table
//   k_pr_envelope  one thread per cell walking ALL positions of the chunk from shared memory,
//                  answers scattered straight into the precision tensor
//   k_pr_finalize  one thread per precision entry
#include <limits.h>
#include <stdlib.h>
#include "ta_internal.h"
#include "ta_device_fns.cuh"

#define PR_CHUNK 256          // detections per chunk == threads per block

struct PrArgs {
    int n_cat, n_thr, n_cfg, n_rec;
    int64_t n_dt;
    int n_chunks_ub;
    const int64_t* cat_dt_off;
    const int32_t* acc_perm;
    const uint32_t* dt_tpfp;     // [n_dt][n_cfg]
    const uint32_t* dt_word;     // [n_dt] compact result words of ta_frame_eval (or NULL): bit 31
                                 // set = use the detection's row in dt_tpfp
    const int32_t* num_gt;       // [n_cat][n_cfg]
    const double* rec_thrs;
    // scratch
    int32_t* chunk_start;        // [n_cat + 1]
    uint32_t* chunk_cnt;         // [n_chunks_ub][n_cfg][32]: bit t -> TP count, bit 16+t -> FP count
    uint32_t* cat_tot;           // [n_cat][n_cfg][32] category totals (same bit layout)
    int32_t* tk;                 // [n_cat][n_cfg][n_rec]
    unsigned long long* chunk_best;  // [n_chunks_ub][n_cfg][n_thr] packed candidates
    int best_cell_major;             // (alternative layout [cell][chunk]; measured slower, unused)
    unsigned long long* ans;     // cell-major answers [n_thr][n_cat][n_cfg][n_rec] (bit-plane path): what
                                 // prec_bits holds, laid out so that one cell's recall levels are
                                 // contiguous (sequential stores in the envelope, row reads in finalize)
    int32_t* chunk_cat;          // [n_chunks_ub] category of a chunk            (bit-plane path)
    int64_t* chunk_p0;           // [n_chunks_ub] first position (accumulate order) of a chunk
    int32_t* chunk_np;           // [n_chunks_ub] positions in the chunk (<= PR_CHUNK)
    uint32_t* bits;              // [n_cfg * n_thr][n_chunks_ub][2 * TA_PR_WORDS] (bit-plane path):
                                 // 64 B per (cell, chunk): word j < 8 = TP flags of positions
                                 // 32 j .. 32 j + 31 of the chunk, word 8 + j = FP flags
    // outputs
    unsigned long long* prec_bits;   // precision buffer viewed as u64: packed (t << 32 | n) until
                                     // k_pr_finalize turns it into doubles
    double* precision;
    double* recall;
    int64_t* tp_cnt;
    int64_t* fp_cnt;
    int* flags;                  // [1]: a category exceeds PR_MAX_CAT_DT
};

#define PR_MAX_CAT_DT (1 << 24)  // detections per category (packed candidates: 24-bit counts)

// TP/FP word of (detection idx, cfg): the row entry, or the expansion of the compact word
// (pr_expand, ta_device_fns.cuh)
__device__ __forceinline__ uint32_t pr_word(const PrArgs& a, int64_t idx, int cfg) {
    if (a.dt_word) {
        const uint32_t w = a.dt_word[idx];
        if (!(w >> 31)) return pr_expand(w, cfg, a.n_thr, a.n_cfg);
    }
    return a.dt_tpfp[idx * a.n_cfg + cfg];
}

__global__ void __launch_bounds__(1024)
k_pr_plan(PrArgs a) {
    // one block: exclusive scan of ceil(len / PR_CHUNK) over the categories
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < a.n_cat; base += 1024) {
        const int c = base + threadIdx.x;
        int n = 0;
        if (c < a.n_cat) {
            const int64_t len = a.cat_dt_off[c + 1] - a.cat_dt_off[c];
            if (len >= PR_MAX_CAT_DT) atomicExch(&a.flags[1], 1);
            n = (int)((len + PR_CHUNK - 1) / PR_CHUNK);
        }
        int incl = n;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int v = wsum[lane], w = v;
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            wsum[lane] = w - v;          // exclusive prefix of the warp sums
        }
        __syncthreads();
        const int excl = carry_s + wsum[warp] + incl - n;
        if (c < a.n_cat) a.chunk_start[c] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + n;
        __syncthreads();
    }
    if (threadIdx.x == 0) a.chunk_start[a.n_cat] = carry_s;
}

// category of a chunk: last c with chunk_start[c] <= chunk
__device__ __forceinline__ int pr_find_cat(const int32_t* chunk_start, int n_cat, int chunk) {
    int lo = 0, hi = n_cat;   // invariant: chunk_start[lo] <= chunk < chunk_start[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_start[mid] <= chunk) lo = mid; else hi = mid;
    }
    return lo;
}

// Chunk table of the bit-plane path, one thread per chunk: category, first position, length —
// so that no later kernel has to search or chase offsets.
__global__ void __launch_bounds__(256)
k_pr_chunks(PrArgs a) {
    const int n_chunks = a.chunk_start[a.n_cat];
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= n_chunks) return;
    const int cat = pr_find_cat(a.chunk_start, a.n_cat, ch);
    const int rel = ch - a.chunk_start[cat];
    const int64_t p0 = a.cat_dt_off[cat] + (int64_t)rel * PR_CHUNK;
    a.chunk_cat[ch] = cat;
    a.chunk_p0[ch] = p0;
    a.chunk_np[ch] = (int)min((int64_t)PR_CHUNK, a.cat_dt_off[cat + 1] - p0);
}

// One warp per range cfg, lane = 8 consecutive positions of the chunk.  Each lane adds its 8
// TP/FP words into bit-sliced counters (carry-save: all 32 bit positions at once), expands
// them to one count per bit, and the warp sums the lanes with REDUX.
#define PR_COUNT_MAX_CFG 8    // cfgs (warps) per k_pr_count block

__global__ void __launch_bounds__(PR_COUNT_MAX_CFG * 32)
k_pr_count(PrArgs a) {
    const int chunk = blockIdx.x;
    if (chunk >= a.chunk_start[a.n_cat]) return;
    const int cat = pr_find_cat(a.chunk_start, a.n_cat, chunk);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t p0 = a.cat_dt_off[cat] + (int64_t)(chunk - a.chunk_start[cat]) * PR_CHUNK;
    const int n_pos = (int)min((int64_t)PR_CHUNK, a.cat_dt_off[cat + 1] - p0);
    for (int cfg = blockIdx.y * PR_COUNT_MAX_CFG + warp; cfg < a.n_cfg;
         cfg += gridDim.y * PR_COUNT_MAX_CFG) {
        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
        for (int u = 0; u < PR_CHUNK / 32; ++u) {
            const int p = u * 32 + lane;               // coalesced permutation reads
            uint32_t w = 0;
            if (p < n_pos) w = pr_word(a, (int64_t)a.acc_perm[p0 + p], cfg);
            const uint32_t k0 = c0 & w;  c0 ^= w;
            const uint32_t k1 = c1 & k0; c1 ^= k0;
            const uint32_t k2 = c2 & k1; c2 ^= k1;
            c3 ^= k2;
        }
        uint32_t mine = 0;
        for (int b = 0; b < 32; ++b) {
            if ((b & 15) >= a.n_thr) continue;
            const uint32_t v = ((c0 >> b) & 1u) + (((c1 >> b) & 1u) << 1) + (((c2 >> b) & 1u) << 2) +
                               (((c3 >> b) & 1u) << 3);
            const uint32_t tot = __reduce_add_sync(0xffffffffu, v);
            if (lane == b) mine = tot;
        }
        a.chunk_cnt[((int64_t)chunk * a.n_cfg + cfg) * 32 + lane] = mine;
    }
}

__global__ void k_pr_scan(PrArgs a) {
    // one block per category; thread j <-> counter (cfg, bit)
    const int cat = blockIdx.x;
    const int ch0 = a.chunk_start[cat], ch1 = a.chunk_start[cat + 1];
    const int n_ctr = a.n_cfg * 32;
    const bool has_dt = a.cat_dt_off[cat + 1] > a.cat_dt_off[cat];
    for (int j = threadIdx.x; j < n_ctr; j += blockDim.x) {
        const int cfg = j >> 5, bit = j & 31;
        uint32_t run = 0;
        int ch = ch0;
        for (; ch + 8 <= ch1; ch += 8) {          // 8 independent loads in flight
            uint32_t v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = a.chunk_cnt[((int64_t)(ch + u) * a.n_cfg) * 32 + j];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a.chunk_cnt[((int64_t)(ch + u) * a.n_cfg) * 32 + j] = run;
                run += v[u];
            }
        }
        for (; ch < ch1; ++ch) {
            uint32_t* q = a.chunk_cnt + ((int64_t)ch * a.n_cfg) * 32 + j;
            const uint32_t v = *q;
            *q = run;
            run += v;
        }
        a.cat_tot[((int64_t)cat * a.n_cfg) * 32 + j] = run;
        const int ngt = a.num_gt[(int64_t)cat * a.n_cfg + cfg];
        const int t = bit & 15;
        if (t < a.n_thr) {
            const int64_t cell = ((int64_t)t * a.n_cat + cat) * a.n_cfg + cfg;
            if (bit < 16) {
                if (a.tp_cnt) a.tp_cnt[cell] = ngt ? (int64_t)run : 0;
                // eval.py:522-525 (cell keeps -1 without GT), :543-547 (rc[-1] or 0)
                a.recall[cell] = ngt == 0 ? -1.0 : (has_dt ? (double)run / (double)ngt : 0.0);
            } else if (a.fp_cnt) {
                a.fp_cnt[cell] = ngt ? (int64_t)run : 0;
            }
        }
    }
    for (int j = threadIdx.x; j < a.n_cfg * a.n_rec; j += blockDim.x) {
        const int cfg = j / a.n_rec, k = j - cfg * a.n_rec;
        const int ngt = a.num_gt[(int64_t)cat * a.n_cfg + cfg];
        a.tk[((int64_t)cat * a.n_cfg + cfg) * a.n_rec + k] =
            ngt ? (int32_t)ta_min_tp_for_recall(a.rec_thrs[k], ngt) : INT_MAX;
    }
}

// pr_better / pr_pack / pr_unpack: ta_device_fns.cuh

__device__ __forceinline__ int64_t pr_best_idx(const PrArgs& a, int chunk, int cell) {
    return a.best_cell_major ? (int64_t)cell * a.n_chunks_ub + chunk
                             : (int64_t)chunk * (a.n_cfg * a.n_thr) + cell;
}

#define PR_ENV_MAX_CELLS 64   // (cfg, threshold) cells per k_pr_envelope block

__global__ void __launch_bounds__(PR_ENV_MAX_CELLS)
k_pr_envelope(PrArgs a, int cfgs_per_block) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: words u32 [PR_CHUNK][ncf] | tk i32 [ncf][n_rec]
    uint32_t* words = reinterpret_cast<uint32_t*>(smem_raw);
    int32_t* tk_s = reinterpret_cast<int32_t*>(words + (size_t)PR_CHUNK * cfgs_per_block);
    const int chunk = blockIdx.x;
    if (chunk >= a.chunk_start[a.n_cat]) return;
    const int cat = pr_find_cat(a.chunk_start, a.n_cat, chunk);
    const int cfg0 = blockIdx.y * cfgs_per_block;
    const int ncf = min(cfgs_per_block, a.n_cfg - cfg0);
    const int ch0 = a.chunk_start[cat], ch1 = a.chunk_start[cat + 1];
    const int64_t p0 = a.cat_dt_off[cat] + (int64_t)(chunk - ch0) * PR_CHUNK;
    const int n_pos = (int)min((int64_t)PR_CHUNK, a.cat_dt_off[cat + 1] - p0);
    // ---- stage the chunk's TP/FP words (rows gathered through the score permutation) and tk
    for (int i = threadIdx.x; i < n_pos * ncf; i += blockDim.x) {
        const int p = i / ncf, c = i - p * ncf;
        words[p * ncf + c] = pr_word(a, (int64_t)a.acc_perm[p0 + p], cfg0 + c);
    }
    for (int i = threadIdx.x; i < ncf * a.n_rec; i += blockDim.x)
        tk_s[i] = a.tk[((int64_t)cat * a.n_cfg + cfg0) * a.n_rec + i];
    __syncthreads();
    const int cell = threadIdx.x;
    if (cell >= ncf * a.n_thr) return;
    const int c = cell / a.n_thr, b = cell - c * a.n_thr;
    const int cfg = cfg0 + c;
    if (a.num_gt[(int64_t)cat * a.n_cfg + cfg] == 0) return;
    // counts at the END of this chunk = exclusive prefix of the next chunk (category totals
    // for the last chunk)
    const uint32_t* nxt = (chunk + 1 < ch1) ? a.chunk_cnt + ((int64_t)(chunk + 1) * a.n_cfg + cfg) * 32
                                            : a.cat_tot + ((int64_t)cat * a.n_cfg + cfg) * 32;
    uint32_t tc = nxt[b], fc = nxt[16 + b];
    const uint32_t t_begin = a.chunk_cnt[((int64_t)chunk * a.n_cfg + cfg) * 32 + b];
    unsigned long long* best_out = a.chunk_best + pr_best_idx(a, chunk, cfg * a.n_thr + b);
    if (tc == t_begin) { *best_out = 0ull; return; }          // no TP of this cell in the chunk
    const int32_t* tkc = tk_s + c * a.n_rec;
    // last recall threshold whose (clamped) tk is <= tc
    int kq = a.n_rec - 1;
    while (kq >= 0 && (uint32_t)max(tkc[kq], 1) > tc) --kq;
    uint32_t next_tk = kq >= 0 ? (uint32_t)max(tkc[kq], 1) : 0u;
    uint32_t bt = 0, bn = 1;       // precision 0: the first true positive always beats it
    const uint32_t ch_rel = (uint32_t)(chunk - ch0);
    const int64_t per_t = (int64_t)a.n_cat * a.n_cfg;
    // answers of this cell: prec_bits[(b * n_rec + k) * per_t + cat * n_cfg + cfg]
    unsigned long long* qp = a.prec_bits + ((int64_t)b * a.n_rec + kq) * per_t + (int64_t)cat * a.n_cfg + cfg;
    const uint32_t tp_bit = 1u << b, fp_bit = 1u << (16 + b);
    const uint32_t* wp = words + (n_pos - 1) * ncf + c;
    for (int p = n_pos; p > 0; --p, wp -= ncf) {
        const uint32_t w = *wp;
        if (w & tp_bit) {
            const uint32_t n = tc + fc;
            // strict ">" is enough here: a tie keeps the later detection's pair, and the only
            // value-changing tie, (1,1) vs (k,k), has (1,1) as the candidate (first TP overall)
            if ((unsigned long long)tc * bn > (unsigned long long)bt * n) { bt = tc; bn = n; }
            while (next_tk == tc) {
                *qp = pr_pack(bt, bn, ch_rel);
                qp -= per_t;
                --kq;
                next_tk = kq >= 0 ? (uint32_t)max(tkc[kq], 1) : 0u;
            }
            --tc;
        } else if (w & fp_bit) {
            --fc;
        }
    }
    *best_out = pr_pack(bt, bn, 0);
}


// ---- bit-plane variant -------------------------------------------------------------------
static_assert(PR_CHUNK == 32 * TA_PR_WORDS, "chunk = 8 words of 32 positions");
#define PR_PS 17              // shared-memory stride of one cell's 16 plane words

// One block per chunk, thread = position (warp w = positions 32 w .. 32 w + 31).
//   compact words (ta_frame_eval): ONE 32 x 32 bit transpose of the warp's words turns lane b
//     into the ballot of word bit b; the (cfg, threshold) cells then are
//     TP = A_cfg & M_k, FP = (B_cfg & M_k) | (U_cfg & ~M_k) (pr_cell_planes), two or three cells
//     per lane, their four operands fetched with shuffles.  Detections whose word points to a
//     full row (general matcher) are patched in one by one.
//   full rows (track path, exchanged rows): per cfg the warp transposes its 32 TP/FP words:
//     lane t then holds the TP plane word of threshold t, lane 16 + t the FP plane word.
// Planes are staged in shared memory cell-major ([cell][16 words]) and written as one 64-byte
// piece per cell; chunk totals are popcounts.
__global__ void __launch_bounds__(PR_CHUNK, 4)
k_pr_bits(PrArgs a) {
    extern __shared__ uint32_t plane_s[];    // [n_cells][PR_PS]: 16 plane words + 1 pad (bank-conflict-free)
    __shared__ uint32_t rowstage[PR_CHUNK / 32][32];
    const int n_chunks = a.chunk_start[a.n_cat];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_thr = a.n_thr, n_cfg = a.n_cfg;
    const int n_cells = n_cfg * n_thr;
    const int p = threadIdx.x;
    const int g = gridDim.x;
    uint32_t keep[5], rot[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) pr_transpose_consts(lane, 16 >> s, keep[s], rot[s]);
    // compact path: the cells this lane assembles (<= 3 rounds of 32: n_thr + 3 n_cfg <= 31 means
    // n_cells <= 80) — source lanes of its four operands and its shared-memory slot, fixed for
    // the whole kernel (the integer divisions stay out of the chunk loop)
    int src_k[3], src_a[3], soff[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int cell = min(r * 32 + lane, n_cells - 1);
        const int cfg = cell / n_thr;
        src_k[r] = cell - cfg * n_thr;
        src_a[r] = n_thr + cfg;
        soff[r] = (r * 32 + lane < n_cells) ? cell * PR_PS + warp : -1;
    }
    // Persistent blocks with a three-deep software pipeline over their chunks: the chunk table
    // entry of chunk c + 3g, the permutation entries of c + 2g and the result words of c + g
    // are in flight while chunk c is transposed (each is a dependent load of the previous one).
    auto desc_p0 = [&](int ch) { return ch < n_chunks ? a.chunk_p0[ch] : (int64_t)0; };
    auto desc_np = [&](int ch) { return ch < n_chunks ? a.chunk_np[ch] : 0; };
    auto perm_of = [&](int64_t p0, int np) { return p < np ? a.acc_perm[p0 + p] : -1; };
    auto word_of = [&](int pm) { return (a.dt_word && pm >= 0) ? a.dt_word[pm] : 0u; };
    int c = blockIdx.x;
    int64_t p0_1 = desc_p0(c + g), p0_2 = desc_p0(c + 2 * g);
    int np_1 = desc_np(c + g), np_2 = desc_np(c + 2 * g);
    int perm0 = perm_of(desc_p0(c), desc_np(c)), perm1 = perm_of(p0_1, np_1);
    uint32_t w0 = word_of(perm0);
    for (; c < n_chunks; c += g) {
        const uint32_t w1 = word_of(perm1);
        const int perm2 = perm_of(p0_2, np_2);
        const int64_t p0_3 = desc_p0(c + 3 * g);
        const int np_3 = desc_np(c + 3 * g);
        const bool live = perm0 >= 0;
        const uint32_t* row = a.dt_tpfp + (int64_t)(live ? perm0 : 0) * n_cfg;
        if (a.dt_word) {
            const uint32_t cw = w0;
            const bool full = (cw >> 31) != 0;
            uint32_t x = full ? 0u : cw;
#pragma unroll
            for (int s = 0; s < 5; ++s)
                x = pr_transpose_apply(x, __shfl_xor_sync(0xffffffffu, x, 16 >> s), keep[s], rot[s]);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                if (r * 32 < n_cells) {              // warp-uniform
                    const uint32_t Mk = __shfl_sync(0xffffffffu, x, src_k[r]);
                    const uint32_t A = __shfl_sync(0xffffffffu, x, src_a[r]);
                    const uint32_t B = __shfl_sync(0xffffffffu, x, src_a[r] + n_cfg);
                    const uint32_t U = __shfl_sync(0xffffffffu, x, src_a[r] + 2 * n_cfg);
                    uint32_t tp, fp;
                    pr_cell_planes(Mk, A, B, U, tp, fp);
                    if (soff[r] >= 0) {
                        plane_s[soff[r]] = tp;
                        plane_s[soff[r] + TA_PR_WORDS] = fp;
                    }
                }
            }
            uint32_t fm = __ballot_sync(0xffffffffu, full);
            while (fm) {                              // warp-uniform: detections with a full row
                const int src = __ffs(fm) - 1;
                fm &= fm - 1;
                __syncwarp();
                if (lane == src)
                    for (int q = 0; q < n_cfg; ++q) rowstage[warp][q] = row[q];
                __syncwarp();
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    if (soff[r] >= 0) {
                        const uint32_t rw = rowstage[warp][src_a[r] - n_thr];
                        plane_s[soff[r]] |= ((rw >> src_k[r]) & 1u) << src;
                        plane_s[soff[r] + TA_PR_WORDS] |= ((rw >> (16 + src_k[r])) & 1u) << src;
                    }
                }
            }
        } else {
            const int b = lane & 15;
            uint32_t* dst = plane_s + (lane >> 4) * TA_PR_WORDS + warp;
#pragma unroll 2
            for (int cfg = 0; cfg < n_cfg; ++cfg) {
                uint32_t x = live ? row[cfg] : 0u;
#pragma unroll
                for (int s = 0; s < 5; ++s)
                    x = pr_transpose_apply(x, __shfl_xor_sync(0xffffffffu, x, 16 >> s), keep[s], rot[s]);
                if (b < n_thr) dst[(cfg * n_thr + b) * PR_PS] = x;
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n_cells * 16; i += PR_CHUNK) {
            const int cell = i >> 4, w = i & 15;          // 16 consecutive threads write one 64-byte piece
            a.bits[((int64_t)cell * a.n_chunks_ub + c) * 16 + w] = plane_s[cell * PR_PS + w];
        }
        for (int j = threadIdx.x; j < n_cfg * 32; j += PR_CHUNK) {
            const int cfg = j >> 5, bit = j & 31, t = bit & 15;
            uint32_t cnt = 0;
            if (t < n_thr) {
                const uint32_t* src = plane_s + (cfg * n_thr + t) * PR_PS + (bit >> 4) * TA_PR_WORDS;
#pragma unroll
                for (int u = 0; u < TA_PR_WORDS; ++u) cnt += __popc(src[u]);
            }
            a.chunk_cnt[((int64_t)c * n_cfg + cfg) * 32 + bit] = cnt;
        }
        __syncthreads();                 // planes consumed before the next chunk overwrites them
        w0 = w1; perm0 = perm1; perm1 = perm2;
        p0_2 = p0_3; np_2 = np_3;
    }
}

// One THREAD per (group of PR_SEG consecutive chunks, cell).  It walks the group's chunks last to
// first and CARRIES its state (running counts, suffix-maximum precision, next recall level) from
// chunk to chunk while they belong to one category, so the per-thread set-up — counter loads and
// the search for the first recall level — is paid once per run instead of once per chunk.  The
// best it stores for a chunk is the best of that chunk AND the later chunks of the run: a valid
// input of k_pr_suffix (every stored value is a maximum over true positives at or after the
// chunk).  A warp takes the n_thr cells of ONE range cfg for 32 / n_thr consecutive groups, so
// lanes share tk rows and counters (L1 hits) and walk similar numbers of true positives.
// Measured alternatives at the bench size (one chunk per thread): lanes = 32 consecutive cells
// of one chunk 0.47 ms; lanes = consecutive chunks of one cell 0.73 ms; lanes = chunks of equal
// rank inside their categories 0.71 ms.
#define PR_SEG 4
template <int NT, int NC>
__global__ void __launch_bounds__(128)
k_pr_envelope_bits(PrArgs a, int seg) {
    const int n_thr = NT ? NT : a.n_thr, n_cfg = NC ? NC : a.n_cfg;
    const int n_chunks = a.chunk_start[a.n_cat];
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = (int)(gid & 31);
    const int cpw = 32 / n_thr;
    const int64_t wg = gid >> 5;
    const int cfg = (int)(wg % n_cfg);
    const int64_t trip = wg / n_cfg;
    if (lane >= cpw * n_thr) return;
    const int64_t grp = trip * cpw + lane / n_thr;
    const int b = lane % n_thr;
    const int64_t c_lo64 = grp * seg;
    if (c_lo64 >= n_chunks) return;
    const int c_lo = (int)c_lo64;
    const int c_hi = min(c_lo + seg, n_chunks) - 1;
    const int cell = cfg * n_thr + b;
    const int64_t per_t = (int64_t)a.n_cat * n_cfg;
    int cat = -1, ch0 = 0;
    bool live = false;
    ta_pr_state s;
    const int32_t* tkp = nullptr;
    unsigned long long* q = nullptr;
    for (int c = c_hi; c >= c_lo; --c) {
        const int c_cat = a.chunk_cat[c];
        if (c_cat != cat) {                          // a new run: set the state up
            cat = c_cat;
            const int64_t cc = (int64_t)cat * n_cfg + cfg;
            live = a.num_gt[cc] != 0;
            if (live) {
                ch0 = a.chunk_start[cat];
                const int ch1 = a.chunk_start[cat + 1];
                // counts at the END of this chunk = exclusive prefix of the next chunk (category
                // totals for the last chunk)
                const uint32_t* nxt = (c + 1 < ch1) ? a.chunk_cnt + ((int64_t)(c + 1) * n_cfg + cfg) * 32
                                                    : a.cat_tot + cc * 32;
                tkp = a.tk + cc * a.n_rec;
                q = a.ans + ((int64_t)b * per_t + cc) * a.n_rec;
                ta_pr_state_init(s, nxt[b], nxt[16 + b], tkp, a.n_rec);
            }
        }
        if (!live) continue;
        unsigned long long* best_out = a.chunk_best + pr_best_idx(a, c, cell);
        const uint32_t* cnt = a.chunk_cnt + ((int64_t)c * n_cfg + cfg) * 32;
        if (s.tc == cnt[b]) {                        // no TP of this cell in the chunk
            ta_pr_skip_chunk(s, cnt[16 + b]);
        } else {
            const uint4* pl = reinterpret_cast<const uint4*>(a.bits + ((int64_t)cell * a.n_chunks_ub + c) * 16);
            uint32_t w[16];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint4 v = pl[u];
                w[4 * u] = v.x; w[4 * u + 1] = v.y; w[4 * u + 2] = v.z; w[4 * u + 3] = v.w;
            }
            ta_pr_walk_chunk(s, w, tkp, (uint32_t)(c - ch0), q, (int64_t)1);
        }
        *best_out = s.bt ? pr_pack(s.bt, s.bn, 0) : 0ull;
    }
}

// Exclusive scan of the chunk totals + the tk table, bit-plane path: one block per category,
// thread <-> counter (cfg, bit); only the 2 n_thr live counters of a cfg are touched; 16
// independent loads in flight per thread (predicated, no serial tail).
#define PR_SCAN_DEPTH 16
__global__ void __launch_bounds__(128)
k_pr_scan_live(PrArgs a) {
    const int cat = blockIdx.x;
    const int ch0 = a.chunk_start[cat], ch1 = a.chunk_start[cat + 1];
    const int n_ctr = a.n_cfg * 32;
    const bool has_dt = a.cat_dt_off[cat + 1] > a.cat_dt_off[cat];
    for (int j = threadIdx.x; j < n_ctr; j += blockDim.x) {
        const int cfg = j >> 5, bit = j & 31, t = bit & 15;
        uint32_t run = 0;
        if (t < a.n_thr) {
            uint32_t* col = a.chunk_cnt + j;
            const int64_t rs = (int64_t)a.n_cfg * 32;          // row stride (one chunk)
            for (int ch = ch0; ch < ch1; ch += PR_SCAN_DEPTH) {
                uint32_t v[PR_SCAN_DEPTH];
#pragma unroll
                for (int u = 0; u < PR_SCAN_DEPTH; ++u)
                    v[u] = (ch + u < ch1) ? __ldcg(col + (int64_t)(ch + u) * rs) : 0u;
#pragma unroll
                for (int u = 0; u < PR_SCAN_DEPTH; ++u) {
                    if (ch + u < ch1) col[(int64_t)(ch + u) * rs] = run;
                    run += v[u];
                }
            }
            const int ngt = a.num_gt[(int64_t)cat * a.n_cfg + cfg];
            const int64_t cell = ((int64_t)t * a.n_cat + cat) * a.n_cfg + cfg;
            if (bit < 16) {
                if (a.tp_cnt) a.tp_cnt[cell] = ngt ? (int64_t)run : 0;
                // eval.py:522-525 (cell keeps -1 without GT), :543-547 (rc[-1] or 0)
                a.recall[cell] = ngt == 0 ? -1.0 : (has_dt ? (double)run / (double)ngt : 0.0);
            } else if (a.fp_cnt) {
                a.fp_cnt[cell] = ngt ? (int64_t)run : 0;
            }
        }
        a.cat_tot[((int64_t)cat * a.n_cfg) * 32 + j] = run;
    }
    // tk table of the category: smallest TP count that reaches each recall level
    for (int j = threadIdx.x; j < a.n_cfg * a.n_rec; j += blockDim.x) {
        const int cfg = j / a.n_rec, k = j - cfg * a.n_rec;
        const int ngt = a.num_gt[(int64_t)cat * a.n_cfg + cfg];
        a.tk[((int64_t)cat * a.n_cfg + cfg) * a.n_rec + k] =
            ngt ? (int32_t)ta_min_tp_for_recall(a.rec_thrs[k], ngt) : INT_MAX;
    }
}

// k_pr_finalize for cell-major answers, tiled: a block owns one threshold, PR_FT_CELLS
// consecutive (category, cfg) cells and PR_FT_K consecutive recall levels.
//   phase 1  thread = cell: reads its PR_FT_K answers the way the envelope wrote them (128
//            contiguous bytes), merges each with the best of the later chunks, divides, and
//            parks the values in shared memory ([level][cell]: conflict-free);
//   phase 2  every recall level of the tile is ONE contiguous run of PR_FT_CELLS doubles (2 KB) in
//            precision[T][R][C][cfg]: written with 16-byte stores, a row per warp.
// Long contiguous writes instead of 256-byte pieces scattered over 101 DRAM pages; every entry
// is independent.  Same values as k_pr_finalize.
#define PR_FT_CELLS 256
#define PR_FT_K 16
__global__ void __launch_bounds__(PR_FT_CELLS, 3)
k_pr_finalize_tile(PrArgs a) {
    __shared__ __align__(16) double tile_s[PR_FT_K][PR_FT_CELLS];
    const uint32_t per_t = (uint32_t)a.n_cat * (uint32_t)a.n_cfg;
    const uint32_t cc0 = blockIdx.x * PR_FT_CELLS;
    const int k0 = blockIdx.y * PR_FT_K;
    const int t = blockIdx.z;
    const int nk = min(PR_FT_K, a.n_rec - k0);
    const uint32_t cc = cc0 + threadIdx.x;
    if (cc < per_t) {
        const int ngt = a.num_gt[cc];
        if (ngt == 0) {
#pragma unroll
            for (int k = 0; k < PR_FT_K; ++k) tile_s[k][threadIdx.x] = -1.0;      // eval.py:522-525
        } else {
            const uint32_t tot = a.cat_tot[(int64_t)cc * 32 + t];
            const uint32_t cat = cc / (uint32_t)a.n_cfg, cfg = cc - cat * (uint32_t)a.n_cfg;
            const unsigned long long* best =
                a.chunk_best + pr_best_idx(a, a.chunk_start[cat], (int)(cfg * a.n_thr + t));
            const int64_t best_stride = a.best_cell_major ? 1 : (int64_t)a.n_cfg * a.n_thr;
            const int32_t* tkp = a.tk + (int64_t)cc * a.n_rec + k0;
            const unsigned long long* ansp = a.ans + ((int64_t)t * per_t + cc) * a.n_rec + k0;
            unsigned long long q[PR_FT_K];
            int32_t need[PR_FT_K];
#pragma unroll
            for (int k = 0; k < PR_FT_K; ++k) {
                need[k] = k < nk ? max(tkp[k], 1) : INT_MAX;
                q[k] = ((uint32_t)need[k] <= tot) ? ansp[k] : 0ull;
            }
#pragma unroll
            for (int k = 0; k < PR_FT_K; ++k) {
                double v = 0.0;                                        // eval.py:565-573
                if ((uint32_t)need[k] <= tot) {
                    uint32_t qt, qn, ch, bt, bn, dummy;
                    pr_unpack(q[k], qt, qn, ch);
                    pr_unpack(best[(int64_t)ch * best_stride], bt, bn, dummy);
                    if (pr_better(bt, bn, qt, qn)) { qt = bt; qn = bn; }
                    // ta_precision_at(qt, qn - qt): fp + tp = qn exactly (counts below 2^24)
                    v = __uint2double_rn(qt) / (__uint2double_rn(qn) + 2.220446049250313e-16);
                }
                tile_s[k][threadIdx.x] = v;
            }
        }
    }
    __syncthreads();
    // rows of the tile: PR_FT_CELLS (or fewer at the end) doubles each
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_c = min((uint32_t)PR_FT_CELLS, per_t - cc0);
    for (int k = warp; k < nk; k += PR_FT_CELLS / 32) {
        double* out = a.precision + ((int64_t)t * a.n_rec + k0 + k) * per_t + cc0;
        if ((((uintptr_t)out) & 15) == 0) {
            for (uint32_t c = 2 * lane; c + 1 < n_c; c += 64)
                *reinterpret_cast<double2*>(out + c) = *reinterpret_cast<const double2*>(&tile_s[k][c]);
            if ((n_c & 1) && lane == 0) out[n_c - 1] = tile_s[k][n_c - 1];
        } else {
            for (uint32_t c = lane; c < n_c; c += 32) out[c] = tile_s[k][c];
        }
    }
}

// per (category, cfg, threshold): chunk_best[ch] <- best precision of all LATER chunks
__global__ void k_pr_suffix(PrArgs a) {
    const int cat = blockIdx.x;
    const int ch0 = a.chunk_start[cat], ch1 = a.chunk_start[cat + 1];
    const int n_cell = a.n_cfg * a.n_thr;
    for (int j = threadIdx.x; j < n_cell; j += blockDim.x) {
        if (a.num_gt[(int64_t)cat * a.n_cfg + j / a.n_thr] == 0) continue;
        uint32_t bt = 0, bn = 0, dummy;
        int ch = ch1 - 1;
        for (; ch - 7 >= ch0; ch -= 8) {              // 8 independent loads in flight
            unsigned long long v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = a.chunk_best[pr_best_idx(a, ch - u, j)];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                uint32_t ct, cn;
                pr_unpack(v[u], ct, cn, dummy);
                a.chunk_best[pr_best_idx(a, ch - u, j)] = pr_pack(bt, bn, 0);
                if (pr_better(ct, cn, bt, bn)) { bt = ct; bn = cn; }
            }
        }
        for (; ch >= ch0; --ch) {
            unsigned long long* q = a.chunk_best + pr_best_idx(a, ch, j);
            uint32_t ct, cn;
            pr_unpack(*q, ct, cn, dummy);
            *q = pr_pack(bt, bn, 0);
            if (pr_better(ct, cn, bt, bn)) { bt = ct; bn = cn; }
        }
    }
}

__global__ void k_pr_finalize(PrArgs a) {
    // thread <-> one precision entry (threshold, recall k, category, cfg), cfg fastest: the
    // tensor [T][R][C][cfg] is written once, fully coalesced
    const int64_t per_t = (int64_t)a.n_cat * a.n_cfg;
    const int64_t n = (int64_t)a.n_thr * a.n_rec * per_t;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    int64_t cc, tk_idx;
    if (n < (1ll << 31)) {                       // 32-bit division is several times cheaper
        const uint32_t q = (uint32_t)idx / (uint32_t)per_t;
        tk_idx = q;
        cc = (uint32_t)idx - q * (uint32_t)per_t;
    } else {
        tk_idx = idx / per_t;
        cc = idx - tk_idx * per_t;
    }
    const int ngt = a.num_gt[cc];
    if (ngt == 0) { a.precision[idx] = -1.0; return; }        // eval.py:522-525
    const int t = (int)tk_idx / a.n_rec, k = (int)tk_idx - t * a.n_rec;
    const int cat = (int)(cc / a.n_cfg), cfg = (int)(cc - (int64_t)cat * a.n_cfg);
    const uint32_t need = (uint32_t)max(a.tk[cc * a.n_rec + k], 1);
    if (need > a.cat_tot[cc * 32 + t]) { a.precision[idx] = 0.0; return; }   // eval.py:565-573
    uint32_t qt, qn, ch, bt, bn, dummy;
    pr_unpack(a.prec_bits[idx], qt, qn, ch);
    pr_unpack(a.chunk_best[pr_best_idx(a, a.chunk_start[cat] + (int)ch, cfg * a.n_thr + t)],
              bt, bn, dummy);
    if (pr_better(bt, bn, qt, qn)) { qt = bt; qn = bn; }
    a.precision[idx] = ta_precision_at((int64_t)qt, (int64_t)(qn - qt));
}

// TA_PR_IMPL: 0 = position walk (k_pr_count, k_pr_scan, k_pr_envelope, k_pr_finalize);
//             1 = bit planes (k_pr_bits, k_pr_tk, k_pr_scan_live, k_pr_envelope_bits with
//                 cell-major answers, k_pr_finalize_tile) — the default.
// Both produce identical tensors (tests/test_gpu_parity.py runs both; tools/ab_variants.py
// compares them at the bench size: profiles/r1b_ab_variants.md).
#ifndef TA_PR_IMPL_DEFAULT
#define TA_PR_IMPL_DEFAULT 1
#endif
static int ta_pr_impl() {
    const char* e = getenv("TA_PR_IMPL");
    return (e && *e) ? (e[0] != '0') : TA_PR_IMPL_DEFAULT;
}

extern "C" int ta_pr_accumulate(ta_ctx* ctx, void* stream, int32_t n_cat, const int64_t* cat_dt_off,
                                const int32_t* acc_perm, int64_t n_dt, const uint32_t* dt_tpfp,
                                const uint32_t* dt_word, const int32_t* num_gt, int32_t n_thr, int32_t n_cfg,
                                int32_t n_rec, const double* rec_thrs,
                                double* precision, double* recall, int64_t* tp_cnt, int64_t* fp_cnt) {
    if (!ctx) return ta_set_err(TA_ERR_INVALID, "ta_pr_accumulate: ctx is NULL");
    if (n_thr < 1 || n_thr > TA_MAX_THRS || n_cfg < 1 || n_rec < 1 || n_cat < 0 || n_dt < 0)
        return ta_set_err(TA_ERR_INVALID, "ta_pr_accumulate: bad sizes");
    if (n_cat == 0) return TA_OK;
    if (n_dt / PR_CHUNK + n_cat + 1 > INT_MAX)
        return ta_set_err(TA_ERR_TOO_LARGE, "ta_pr_accumulate: too many detections");

    TA_CUDA(cudaSetDevice(ctx->device));
    ta_begin(ctx, (cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    const int n_chunks_ub = (int)(n_dt / PR_CHUNK) + n_cat;
    // scratch layout
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_start = take(((size_t)n_cat + 1) * 4);
    const size_t o_cnt = take((size_t)n_chunks_ub * n_cfg * 32 * 4);
    const size_t o_tot = take((size_t)n_cat * n_cfg * 32 * 4);
    const size_t o_tk = take((size_t)n_cat * n_cfg * n_rec * 4);
    const size_t o_best = take((size_t)n_chunks_ub * n_cfg * n_thr * 8);
    const size_t n_cells = (size_t)n_cfg * n_thr;
    // the staged planes of one chunk (64 B per cell) and the finalize tile ([n_rec][33] doubles)
    // must fit the default 48 KB of shared memory
    const int impl = (n_cells * PR_PS * 4 <= 48 * 1024) ? ta_pr_impl() : 0;
    const size_t o_ccat = take(impl ? (size_t)n_chunks_ub * 4 : 0);
    const size_t o_cp0 = take(impl ? (size_t)n_chunks_ub * 8 : 0);
    const size_t o_cnp = take(impl ? (size_t)n_chunks_ub * 4 : 0);

    const size_t o_bits = take(impl ? (size_t)n_chunks_ub * 2 * TA_PR_WORDS * n_cells * 4 : 0);
    const size_t o_ans = take(impl ? (size_t)n_thr * n_rec * n_cat * n_cfg * 8 : 0);
    void* ws = nullptr;
    int rc = ta_workspace(ctx, st, off, &ws);
    if (rc) return rc;
    char* base = static_cast<char*>(ws);
    PrArgs a;
    a.n_cat = n_cat; a.n_thr = n_thr; a.n_cfg = n_cfg; a.n_rec = n_rec; a.n_dt = n_dt;
    a.n_chunks_ub = n_chunks_ub;
    a.cat_dt_off = cat_dt_off; a.acc_perm = acc_perm; a.dt_tpfp = dt_tpfp; a.num_gt = num_gt;
    a.dt_word = (dt_word && n_thr + 3 * n_cfg <= 31) ? dt_word : nullptr;
    a.rec_thrs = rec_thrs;
    a.chunk_start = reinterpret_cast<int32_t*>(base + o_start);
    a.chunk_cnt = reinterpret_cast<uint32_t*>(base + o_cnt);
    a.cat_tot = reinterpret_cast<uint32_t*>(base + o_tot);
    a.tk = reinterpret_cast<int32_t*>(base + o_tk);
    a.chunk_best = reinterpret_cast<unsigned long long*>(base + o_best);
    a.chunk_cat = impl ? reinterpret_cast<int32_t*>(base + o_ccat) : nullptr;
    a.best_cell_major = 0;
    a.chunk_p0 = reinterpret_cast<int64_t*>(base + o_cp0);
    a.chunk_np = reinterpret_cast<int32_t*>(base + o_cnp);

    a.bits = reinterpret_cast<uint32_t*>(base + o_bits);
    a.ans = impl ? reinterpret_cast<unsigned long long*>(base + o_ans) : nullptr;
    a.prec_bits = reinterpret_cast<unsigned long long*>(precision);
    a.precision = precision; a.recall = recall; a.tp_cnt = tp_cnt; a.fp_cnt = fp_cnt;

    a.flags = ctx->d_flags;
    k_pr_plan<<<1, 1024, 0, st>>>(a);
    if ((rc = ta_check_launch(ctx, "k_pr_plan"))) return rc;
    if (n_dt >= PR_MAX_CAT_DT) {
        // only then can a single category be too long for the packed (24-bit) counts
        int big = 0;
        TA_CUDA(cudaMemcpyAsync(&big, ctx->d_flags + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        TA_CUDA(cudaStreamSynchronize(st));
        if (big) {
            TA_CUDA(cudaMemsetAsync(ctx->d_flags + 1, 0, sizeof(int), st));
            return ta_set_err(TA_ERR_TOO_LARGE,
                              "ta_pr_accumulate: a category has 2^24 or more detections");
        }
    }
    if (n_chunks_ub > 0 && impl) {
        k_pr_chunks<<<(n_chunks_ub + 255) / 256, 256, 0, st>>>(a);
        if ((rc = ta_check_launch(ctx, "k_pr_chunks"))) return rc;
        int bits_blocks = ctx->sm_count * 4;           // persistent: 4 CTAs of 256 threads per SM
        if (bits_blocks > n_chunks_ub) bits_blocks = n_chunks_ub;
        k_pr_bits<<<bits_blocks, PR_CHUNK, (size_t)PR_PS * n_cells * 4, st>>>(a);
        if ((rc = ta_check_launch(ctx, "k_pr_bits"))) return rc;
    } else if (n_chunks_ub > 0) {
        const int warps = n_cfg < PR_COUNT_MAX_CFG ? n_cfg : PR_COUNT_MAX_CFG;
        dim3 grid(n_chunks_ub, (n_cfg + PR_COUNT_MAX_CFG - 1) / PR_COUNT_MAX_CFG);
        k_pr_count<<<grid, warps * 32, 0, st>>>(a);
        if ((rc = ta_check_launch(ctx, "k_pr_count"))) return rc;
    }
    if (impl) {
        k_pr_scan_live<<<n_cat, 128, 0, st>>>(a);
        if ((rc = ta_check_launch(ctx, "k_pr_scan_live"))) return rc;
    } else {
        k_pr_scan<<<n_cat, 128, 0, st>>>(a);
        if ((rc = ta_check_launch(ctx, "k_pr_scan"))) return rc;
    }
    if (n_chunks_ub > 0 && impl) {
        // grid over the upper bound of chunks: the kernel reads the real count on the device
        const int cpw = 32 / n_thr;
        // chunks per thread: runs only pay off when categories span several chunks
        const int seg = (n_chunks_ub >= 8 * (int64_t)n_cat) ? PR_SEG : 1;
        const int64_t n_grp = ((int64_t)n_chunks_ub + seg - 1) / seg;
        const int64_t threads = ((n_grp + cpw - 1) / cpw) * n_cfg * 32;
        const unsigned eb = (unsigned)((threads + 127) / 128);
        if (n_thr == 10 && n_cfg == 6) k_pr_envelope_bits<10, 6><<<eb, 128, 0, st>>>(a, seg);
        else k_pr_envelope_bits<0, 0><<<eb, 128, 0, st>>>(a, seg);
        if ((rc = ta_check_launch(ctx, "k_pr_envelope_bits"))) return rc;
    } else if (n_chunks_ub > 0) {
        int cpb = PR_ENV_MAX_CELLS / n_thr;
        if (cpb > n_cfg) cpb = n_cfg;
        const size_t smem = (size_t)PR_CHUNK * cpb * 4 + (size_t)cpb * n_rec * 4;
        if (smem > (size_t)ctx->smem_optin)
            return ta_set_err(TA_ERR_TOO_LARGE, "ta_pr_accumulate: too many recall thresholds");
        if (smem > 48 * 1024)
            TA_CUDA(cudaFuncSetAttribute(k_pr_envelope, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(n_chunks_ub, (n_cfg + cpb - 1) / cpb);
        k_pr_envelope<<<grid, PR_ENV_MAX_CELLS, smem, st>>>(a, cpb);
        if ((rc = ta_check_launch(ctx, "k_pr_envelope"))) return rc;
    }
    k_pr_suffix<<<n_cat, 128, 0, st>>>(a);
    if ((rc = ta_check_launch(ctx, "k_pr_suffix"))) return rc;
    const size_t n_prec = (size_t)n_thr * n_rec * n_cat * n_cfg;
    if (impl) {
        dim3 grid((unsigned)(((size_t)n_cat * n_cfg + PR_FT_CELLS - 1) / PR_FT_CELLS),
                  (unsigned)((n_rec + PR_FT_K - 1) / PR_FT_K), (unsigned)n_thr);
        k_pr_finalize_tile<<<grid, PR_FT_CELLS, 0, st>>>(a);
        return ta_check_launch(ctx, "k_pr_finalize_tile");
    }
    k_pr_finalize<<<(unsigned)((n_prec + 255) / 256), 256, 0, st>>>(a);
    return ta_check_launch(ctx, "k_pr_finalize");
}
