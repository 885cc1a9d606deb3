// ta_pr.cu — precision / recall accumulation (sm_100a).
//
// Reference: TaoEval.accumulate (tao_amodal/evaluation/tao_amodal/eval.py:459-584) and
// LVISEval.accumulate (lvis_amodal/eval.py:305-426).  Per (category, range cfg, threshold) the
// reference walks the category's detections in descending-score order, forms running TP / FP
// counts, precision = tp / (fp + tp + eps), takes the suffix maximum ("envelope", :557-559) and
// samples it at the first position whose recall reaches each of the 101 recall thresholds
// (:561-571).  Equivalent formulation used here: the value at recall threshold k is the maximum
// precision over all TRUE POSITIVES whose running TP count is >= tk[k], where tk[k] is the
// smallest count with count / num_gt >= rec_thrs[k].  So every TP raises exactly one bucket
// (the last k with tk[k] <= its count) and a suffix maximum over k finishes the row.
//
// The category lists are cut into chunks of PR_CHUNK detections so that long categories do not
// serialise:
//   k_pr_plan      chunk table: first chunk of every category
//   k_pr_count     per chunk: TP / FP totals of every (cfg, threshold)  (ballot + popc)
//   k_pr_scan      per category: exclusive scan of the chunk totals, tk tables, recall, counts
//   k_pr_bucket    per chunk: running counts -> precision of every TP -> bucket max
//                  (shared-memory atomicMax, then one global atomicMax per touched bucket)
//   k_pr_finalize  per (threshold, category, cfg): suffix max over the recall axis, -1 fill
#include <limits.h>
#include "ta_internal.h"
#include "ta_device_fns.cuh"

#define PR_CHUNK 256          // detections per chunk == threads per block
#define PR_WARPS (PR_CHUNK / 32)
#define PR_CPB 6              // range cfgs handled by one k_pr_bucket block

struct PrArgs {
    int n_cat, n_thr, n_cfg, n_rec;
    int64_t n_dt;
    int n_chunks_ub;
    const int64_t* cat_dt_off;
    const int32_t* acc_perm;
    const uint32_t* dt_tpfp;     // [n_dt][n_cfg]
    const int32_t* num_gt;       // [n_cat][n_cfg]
    const double* rec_thrs;
    // scratch
    int32_t* chunk_start;        // [n_cat + 1]
    uint32_t* chunk_cnt;         // [n_chunks_ub][n_cfg][32]: bit t -> TP count, bit 16+t -> FP count
    int32_t* tk;                 // [n_cat][n_cfg][n_rec]
    // outputs
    unsigned long long* prec_bits;   // precision buffer viewed as u64 (max of fp64 bit patterns)
    double* precision;
    double* recall;
    int64_t* tp_cnt;
    int64_t* fp_cnt;
};

__global__ void k_pr_plan(PrArgs a) {
    // single warp: exclusive scan of ceil(len / PR_CHUNK) over the categories
    const int lane = threadIdx.x;
    int carry = 0;
    for (int base = 0; base < a.n_cat; base += 32) {
        const int c = base + lane;
        int n = 0;
        if (c < a.n_cat) n = (int)((a.cat_dt_off[c + 1] - a.cat_dt_off[c] + PR_CHUNK - 1) / PR_CHUNK);
        int incl = n;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (c < a.n_cat) a.chunk_start[c] = carry + incl - n;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) a.chunk_start[a.n_cat] = carry;
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

__global__ void __launch_bounds__(PR_CHUNK)
k_pr_count(PrArgs a) {
    __shared__ uint32_t wcnt[PR_WARPS][32];
    __shared__ int s_cat;
    const int chunk = blockIdx.x;
    if (chunk >= a.chunk_start[a.n_cat]) return;
    if (threadIdx.x == 0) s_cat = pr_find_cat(a.chunk_start, a.n_cat, chunk);
    __syncthreads();
    const int cat = s_cat;
    const int64_t p = a.cat_dt_off[cat] + (int64_t)(chunk - a.chunk_start[cat]) * PR_CHUNK + threadIdx.x;
    const bool live = p < a.cat_dt_off[cat + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t* row = live ? a.dt_tpfp + (int64_t)a.acc_perm[p] * a.n_cfg : nullptr;
    for (int c = 0; c < a.n_cfg; ++c) {
        const uint32_t w = live ? row[c] : 0u;
        uint32_t mine = 0;
        for (int b = 0; b < a.n_thr; ++b) {
            const uint32_t mt = __ballot_sync(0xffffffffu, (w >> b) & 1u);
            const uint32_t mf = __ballot_sync(0xffffffffu, (w >> (16 + b)) & 1u);
            if (lane == b) mine = __popc(mt);
            if (lane == 16 + b) mine = __popc(mf);
        }
        wcnt[warp][lane] = mine;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t s = 0;
#pragma unroll
            for (int w2 = 0; w2 < PR_WARPS; ++w2) s += wcnt[w2][threadIdx.x];
            a.chunk_cnt[((int64_t)chunk * a.n_cfg + c) * 32 + threadIdx.x] = s;
        }
        __syncthreads();
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
        for (int ch = ch0; ch < ch1; ++ch) {
            uint32_t* q = a.chunk_cnt + ((int64_t)ch * a.n_cfg) * 32 + j;
            const uint32_t v = *q;
            *q = run;
            run += v;
        }
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

__global__ void __launch_bounds__(PR_CHUNK)
k_pr_bucket(PrArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: bucket u64 [PR_CPB][n_thr][n_rec] | tk i32 [PR_CPB][n_rec] | wpre u32 [PR_WARPS][PR_CPB][32]
    unsigned long long* bucket = reinterpret_cast<unsigned long long*>(smem_raw);
    int32_t* tk_s = reinterpret_cast<int32_t*>(bucket + (size_t)PR_CPB * a.n_thr * a.n_rec);
    uint32_t* wpre = reinterpret_cast<uint32_t*>(tk_s + PR_CPB * a.n_rec);
    __shared__ int s_cat;
    const int chunk = blockIdx.x;
    if (chunk >= a.chunk_start[a.n_cat]) return;
    if (threadIdx.x == 0) s_cat = pr_find_cat(a.chunk_start, a.n_cat, chunk);
    __syncthreads();
    const int cat = s_cat;
    const int cfg0 = blockIdx.y * PR_CPB;
    const int ncf = min(PR_CPB, a.n_cfg - cfg0);
    // any cfg of this block with GT?  (cells without GT keep -1, nothing to accumulate)
    bool any = false;
    for (int c = 0; c < ncf; ++c) any |= a.num_gt[(int64_t)cat * a.n_cfg + cfg0 + c] != 0;
    if (!any) return;
    const int n_b = ncf * a.n_thr * a.n_rec;
    for (int i = threadIdx.x; i < n_b; i += PR_CHUNK) bucket[i] = 0ull;
    for (int i = threadIdx.x; i < ncf * a.n_rec; i += PR_CHUNK)
        tk_s[i] = a.tk[((int64_t)cat * a.n_cfg + cfg0) * a.n_rec + i];
    const int64_t p = a.cat_dt_off[cat] + (int64_t)(chunk - a.chunk_start[cat]) * PR_CHUNK + threadIdx.x;
    const bool live = p < a.cat_dt_off[cat + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t* row = live ? a.dt_tpfp + (int64_t)a.acc_perm[p] * a.n_cfg + cfg0 : nullptr;
    uint32_t w[PR_CPB];
#pragma unroll
    for (int c = 0; c < PR_CPB; ++c) w[c] = (live && c < ncf) ? row[c] : 0u;
    // ---- per-warp totals of every (cfg, bit)
#pragma unroll
    for (int c = 0; c < PR_CPB; ++c) {
        if (c < ncf) {
            uint32_t mine = 0;
            for (int b = 0; b < a.n_thr; ++b) {
                const uint32_t mt = __ballot_sync(0xffffffffu, (w[c] >> b) & 1u);
                const uint32_t mf = __ballot_sync(0xffffffffu, (w[c] >> (16 + b)) & 1u);
                if (lane == b) mine = __popc(mt);
                if (lane == 16 + b) mine = __popc(mf);
            }
            wpre[(warp * PR_CPB + c) * 32 + lane] = mine;
        }
    }
    __syncthreads();
    // ---- exclusive prefix over the warps + chunk offset (thread j <-> counter (cfg, bit))
    if (threadIdx.x < ncf * 32) {
        const int c = threadIdx.x >> 5, bit = threadIdx.x & 31;
        uint32_t run = a.chunk_cnt[((int64_t)chunk * a.n_cfg + cfg0 + c) * 32 + bit];
#pragma unroll
        for (int w2 = 0; w2 < PR_WARPS; ++w2) {
            uint32_t* q = &wpre[(w2 * PR_CPB + c) * 32 + bit];
            const uint32_t v = *q;
            *q = run;
            run += v;
        }
    }
    __syncthreads();
    // ---- every TP raises its bucket
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int c = 0; c < PR_CPB; ++c) {
        if (c < ncf && a.num_gt[(int64_t)cat * a.n_cfg + cfg0 + c] != 0) {
            const int32_t* tkc = tk_s + c * a.n_rec;
            for (int b = 0; b < a.n_thr; ++b) {
                const bool tp = (w[c] >> b) & 1u;
                const uint32_t mt = __ballot_sync(0xffffffffu, tp);
                const uint32_t mf = __ballot_sync(0xffffffffu, (w[c] >> (16 + b)) & 1u);
                if (tp) {
                    const int64_t tc = (int64_t)wpre[(warp * PR_CPB + c) * 32 + b] + __popc(mt & lt) + 1;
                    const int64_t fc = (int64_t)wpre[(warp * PR_CPB + c) * 32 + 16 + b] + __popc(mf & lt);
                    int lo = 0, hi = a.n_rec;   // first k with tk[k] > tc
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if ((int64_t)tkc[mid] <= tc) lo = mid + 1; else hi = mid;
                    }
                    if (lo > 0) {
                        const double pr = ta_precision_at(tc, fc);
                        atomicMax(&bucket[((size_t)c * a.n_thr + b) * a.n_rec + lo - 1],
                                  (unsigned long long)__double_as_longlong(pr));
                    }
                }
            }
        }
    }
    __syncthreads();
    // ---- flush touched buckets: precision is [n_thr][n_rec][n_cat][n_cfg]
    for (int i = threadIdx.x; i < n_b; i += PR_CHUNK) {
        const unsigned long long v = bucket[i];
        if (v) {
            const int c = i / (a.n_thr * a.n_rec);
            const int r = i - c * (a.n_thr * a.n_rec);
            const int b = r / a.n_rec, k = r - b * a.n_rec;
            atomicMax(&a.prec_bits[(((int64_t)b * a.n_rec + k) * a.n_cat + cat) * a.n_cfg + cfg0 + c], v);
        }
    }
}

__global__ void k_pr_finalize(PrArgs a) {
    // thread <-> (threshold, category, cfg), cfg fastest: coalesced along the innermost axes
    const int64_t n_cell = (int64_t)a.n_thr * a.n_cat * a.n_cfg;
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cell) return;
    const int64_t per_t = (int64_t)a.n_cat * a.n_cfg;
    const int t = (int)(cell / per_t);
    const int64_t cc = cell - (int64_t)t * per_t;          // cat * n_cfg + cfg
    const int ngt = a.num_gt[cc];
    if (ngt == 0) {
        for (int k = 0; k < a.n_rec; ++k) a.precision[((int64_t)t * a.n_rec + k) * per_t + cc] = -1.0;
        return;
    }
    unsigned long long best = 0ull;   // bit pattern of +0.0; precisions are positive
    for (int k = a.n_rec - 1; k >= 0; --k) {
        const int64_t idx = ((int64_t)t * a.n_rec + k) * per_t + cc;
        const unsigned long long v = a.prec_bits[idx];
        best = v > best ? v : best;
        a.prec_bits[idx] = best;
    }
}

extern "C" int ta_pr_accumulate(ta_ctx* ctx, void* stream, int32_t n_cat, const int64_t* cat_dt_off,
                                const int32_t* acc_perm, int64_t n_dt, const uint32_t* dt_tpfp,
                                const int32_t* num_gt, int32_t n_thr, int32_t n_cfg,
                                int32_t n_rec, const double* rec_thrs,
                                double* precision, double* recall, int64_t* tp_cnt, int64_t* fp_cnt) {
    if (!ctx) return ta_set_err(TA_ERR_INVALID, "ta_pr_accumulate: ctx is NULL");
    if (n_thr < 1 || n_thr > TA_MAX_THRS || n_cfg < 1 || n_rec < 1 || n_cat < 0 || n_dt < 0)
        return ta_set_err(TA_ERR_INVALID, "ta_pr_accumulate: bad sizes");
    if (n_cat == 0) return TA_OK;
    if (n_dt / PR_CHUNK + n_cat + 1 > INT_MAX)
        return ta_set_err(TA_ERR_TOO_LARGE, "ta_pr_accumulate: too many detections");
    TA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n_chunks_ub = (int)(n_dt / PR_CHUNK) + n_cat;
    // scratch layout
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_start = take(((size_t)n_cat + 1) * 4);
    const size_t o_cnt = take((size_t)n_chunks_ub * n_cfg * 32 * 4);
    const size_t o_tk = take((size_t)n_cat * n_cfg * n_rec * 4);
    void* ws = nullptr;
    int rc = ta_workspace(ctx, st, off, &ws);
    if (rc) return rc;
    char* base = static_cast<char*>(ws);
    PrArgs a;
    a.n_cat = n_cat; a.n_thr = n_thr; a.n_cfg = n_cfg; a.n_rec = n_rec; a.n_dt = n_dt;
    a.n_chunks_ub = n_chunks_ub;
    a.cat_dt_off = cat_dt_off; a.acc_perm = acc_perm; a.dt_tpfp = dt_tpfp; a.num_gt = num_gt;
    a.rec_thrs = rec_thrs;
    a.chunk_start = reinterpret_cast<int32_t*>(base + o_start);
    a.chunk_cnt = reinterpret_cast<uint32_t*>(base + o_cnt);
    a.tk = reinterpret_cast<int32_t*>(base + o_tk);
    a.prec_bits = reinterpret_cast<unsigned long long*>(precision);
    a.precision = precision; a.recall = recall; a.tp_cnt = tp_cnt; a.fp_cnt = fp_cnt;

    const size_t n_cell = (size_t)n_thr * n_cat * n_cfg;
    TA_CUDA(cudaMemsetAsync(precision, 0, n_cell * n_rec * sizeof(double), st));
    k_pr_plan<<<1, 32, 0, st>>>(a);
    if ((rc = ta_check_launch(ctx, "k_pr_plan"))) return rc;
    if (n_chunks_ub > 0) {
        k_pr_count<<<n_chunks_ub, PR_CHUNK, 0, st>>>(a);
        if ((rc = ta_check_launch(ctx, "k_pr_count"))) return rc;
    }
    k_pr_scan<<<n_cat, 128, 0, st>>>(a);
    if ((rc = ta_check_launch(ctx, "k_pr_scan"))) return rc;
    if (n_chunks_ub > 0) {
        const size_t smem = (size_t)PR_CPB * n_thr * n_rec * 8 + (size_t)PR_CPB * n_rec * 4 +
                            (size_t)PR_WARPS * PR_CPB * 32 * 4;
        if (smem > (size_t)ctx->smem_optin)
            return ta_set_err(TA_ERR_TOO_LARGE, "ta_pr_accumulate: too many recall thresholds");
        TA_CUDA(cudaFuncSetAttribute(k_pr_bucket, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(n_chunks_ub, (n_cfg + PR_CPB - 1) / PR_CPB);
        k_pr_bucket<<<grid, PR_CHUNK, smem, st>>>(a);
        if ((rc = ta_check_launch(ctx, "k_pr_bucket"))) return rc;
    }
    const int fin_blocks = (int)((n_cell + 255) / 256);
    k_pr_finalize<<<fin_blocks, 256, 0, st>>>(a);
    return ta_check_launch(ctx, "k_pr_finalize");
}
