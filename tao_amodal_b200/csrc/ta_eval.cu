// ta_eval.cu — sm_100a kernels + C ABI of the TAO-Amodal evaluation hot path.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared
//        -Xcompiler -fPIC   (see __graft_entry__.build()).  -fmad=false is load-bearing:
// every product/sum below must round exactly like the reference's Python floats
// (tao_amodal/evaluation/tao_amodal/eval.py:15-48), so no FMA contraction is allowed.
//
// Kernels (DESIGN.md has the roofline of each):
//   k_track_iou_tiled   spatio-temporal IoU, GT tracks staged densely in shared memory
//   k_track_iou_pair    one thread per track pair, sequential merge (alt. modes, fallback)
//   k_box_iou           per-(image,category) box IoU (pycocotools bbIou semantics)
//   k_match_greedy      sequential greedy assignment, one lane per (range cfg, threshold)
//   k_pr_accumulate     PR curve + 101-point interpolation, one warp per threshold
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <limits.h>
#include <math.h>

#include "ta_eval.h"
#include "ta_device_fns.cuh"

// ------------------------------------------------------------------------------------------
// context / errors
// ------------------------------------------------------------------------------------------
struct ta_ctx {
    int device;
    int sm_count;
    int smem_optin;       // max dynamic shared memory per block (opt-in)
    int64_t launches;
    int* d_flags;         // [0]: "i > u" assertion counter (eval.py:95)
    cudaStream_t own_stream;
};

static thread_local char g_err[512] = "";

static int set_err(int code, const char* fmt, const char* a = "", long long b = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

#define TA_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return set_err(TA_ERR_CUDA, "CUDA error %s at line %lld", cudaGetErrorString(e_), \
                           (long long)__LINE__);                                           \
    } while (0)

extern "C" int ta_abi_version(void) { return TA_ABI_VERSION; }
extern "C" const char* ta_last_error(void) { return g_err; }

extern "C" int ta_ctx_create(int device, ta_ctx** out) {
    if (!out) return set_err(TA_ERR_INVALID, "ta_ctx_create: out is NULL");
    int n = 0;
    TA_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return set_err(TA_ERR_INVALID, "ta_ctx_create: bad device %s%lld", "", device);
    TA_CUDA(cudaSetDevice(device));
    ta_ctx* c = new ta_ctx();
    c->device = device;
    c->launches = 0;
    TA_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    TA_CUDA(cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    TA_CUDA(cudaMalloc(&c->d_flags, 4 * sizeof(int)));
    TA_CUDA(cudaMemset(c->d_flags, 0, 4 * sizeof(int)));
    TA_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    *out = c;
    return TA_OK;
}

extern "C" int ta_ctx_destroy(ta_ctx* c) {
    if (!c) return TA_OK;
    cudaSetDevice(c->device);
    cudaFree(c->d_flags);
    cudaStreamDestroy(c->own_stream);
    delete c;
    return TA_OK;
}

extern "C" int ta_ctx_sm_count(const ta_ctx* c) { return c ? c->sm_count : 0; }
extern "C" int64_t ta_ctx_launch_count(const ta_ctx* c) { return c ? c->launches : 0; }

// ------------------------------------------------------------------------------------------
// warp helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
    // fixed xor tree -> deterministic association
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// K1a: tiled spatio-temporal IoU (TA_IOU_3D)
//
// One CTA per (video, category) group.  GT tracks of the group are expanded into dense
// per-slot arrays in shared memory (absent slots hold a sentinel box whose intersection with
// anything is empty), GT_TILE tracks at a time.  Each warp then streams one predicted track:
// lane k loads box k (coalesced, 32 B + 4 B per box, each box read from HBM once), looks up
// the GT box of the same frame slot for every staged GT track and accumulates the
// intersection area.  The union needs no per-frame work:
//     sum_t U = sum_{dt frames} area + sum_{gt frames} area - sum_{common frames} I
// which equals the reference's running sum (eval.py:87-94) whenever its partial sums are
// exact, and is within a few ulp otherwise (the reference's own value then depends on
// CPython set iteration order, eval.py:83).
// ------------------------------------------------------------------------------------------
#define GT_TILE 8
#define TI_WARPS 8

struct TrackIouArgs {
    const int64_t* grp_dt_off;
    const int64_t* grp_gt_off;
    const int64_t* dt_off;
    const double* dt_box;
    const int32_t* dt_slot;
    const int64_t* gt_off;
    const double* gt_box;
    const int32_t* gt_slot;
    const int64_t* iou_off;
    double* iou;
    int S;  // slots per window
};

__global__ void __launch_bounds__(TI_WARPS * 32)
k_track_iou_tiled(TrackIouArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.S;
    // planes: A = (x1, y1), B = (x2, y2); [GT_TILE][S + 1] each (index S = sentinel slot)
    double2* shA = reinterpret_cast<double2*>(smem_raw);
    double2* shB = shA + GT_TILE * (S + 1);
    __shared__ double ga_sh[GT_TILE];
    __shared__ int span_sh[2 * GT_TILE];

    const int grp = blockIdx.x;
    const int64_t d0 = a.grp_dt_off[grp], d1 = a.grp_dt_off[grp + 1];
    const int64_t g0 = a.grp_gt_off[grp], g1 = a.grp_gt_off[grp + 1];
    const int D = (int)(d1 - d0), G = (int)(g1 - g0);
    if (D == 0 || G == 0) return;
    double* out = a.iou + a.iou_off[grp];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);

    for (int gt0 = 0; gt0 < G; gt0 += GT_TILE) {
        const int gcnt = min(GT_TILE, G - gt0);
        // slot span of this GT tile (tracks are sorted by slot: first / last box)
        if (threadIdx.x < gcnt) {
            const int64_t b0 = a.gt_off[g0 + gt0 + threadIdx.x];
            const int64_t b1 = a.gt_off[g0 + gt0 + threadIdx.x + 1];
            span_sh[2 * threadIdx.x] = a.gt_slot[b0];
            span_sh[2 * threadIdx.x + 1] = a.gt_slot[b1 - 1];
        }
        __syncthreads();
        int smin = INT_MAX, smax = INT_MIN;
        for (int j = 0; j < gcnt; ++j) {
            smin = min(smin, span_sh[2 * j]);
            smax = max(smax, span_sh[2 * j + 1]);
        }
        const int n_win = (smax - smin) / S + 1;

        for (int win = 0; win < n_win; ++win) {
            const int w0 = smin + win * S;
            __syncthreads();  // previous window fully consumed
            // 1) sentinel fill
            for (int idx = threadIdx.x; idx < GT_TILE * (S + 1); idx += blockDim.x) {
                shA[idx] = make_double2(INF, 0.0);
                shB[idx] = make_double2(-INF, 0.0);
            }
            __syncthreads();
            // 2) scatter GT boxes, one warp per GT track; total area of the track
            for (int j = warp; j < gcnt; j += TI_WARPS) {
                const int64_t b0 = a.gt_off[g0 + gt0 + j], b1 = a.gt_off[g0 + gt0 + j + 1];
                double area = 0.0;
                for (int64_t k = b0 + lane; k < b1; k += 32) {
                    const double2 p = *reinterpret_cast<const double2*>(a.gt_box + 4 * k);
                    const double2 q = *reinterpret_cast<const double2*>(a.gt_box + 4 * k + 2);
                    area += q.x * q.y;
                    const int s = a.gt_slot[k] - w0;
                    if ((unsigned)s < (unsigned)S) {
                        shA[j * (S + 1) + s] = make_double2(p.x, p.y);
                        shB[j * (S + 1) + s] = make_double2(p.x + q.x, p.y + q.y);
                    }
                }
                area = warp_sum(area);
                if (lane == 0) ga_sh[j] = area;
            }
            __syncthreads();
            // 3) stream predicted tracks, one warp per track
            for (int i = warp; i < D; i += TI_WARPS) {
                const int64_t b0 = a.dt_off[d0 + i], b1 = a.dt_off[d0 + i + 1];
                double acc[GT_TILE];
#pragma unroll
                for (int j = 0; j < GT_TILE; ++j) acc[j] = 0.0;
                double da = 0.0;
                int64_t k = b0 + lane;
                double2 p = make_double2(0, 0), q = make_double2(0, 0);
                int sl = 0;
                if (k < b1) {
                    p = *reinterpret_cast<const double2*>(a.dt_box + 4 * k);
                    q = *reinterpret_cast<const double2*>(a.dt_box + 4 * k + 2);
                    sl = a.dt_slot[k];
                }
                while (k < b1) {
                    // prefetch the next box of this lane before computing on the current one
                    const int64_t kn = k + 32;
                    double2 pn = p, qn = q;
                    int sn = sl;
                    if (kn < b1) {
                        pn = *reinterpret_cast<const double2*>(a.dt_box + 4 * kn);
                        qn = *reinterpret_cast<const double2*>(a.dt_box + 4 * kn + 2);
                        sn = a.dt_slot[kn];
                    }
                    const double dx = p.x, dy = p.y;
                    const double dx2 = p.x + q.x, dy2 = p.y + q.y;
                    da += q.x * q.y;
                    int s = sl - w0;
                    s = ((unsigned)s < (unsigned)S) ? s : S;
#pragma unroll
                    for (int j = 0; j < GT_TILE; ++j) {
                        if (j < gcnt) {
                            const double2 A = shA[j * (S + 1) + s];
                            const double2 B = shB[j * (S + 1) + s];
                            acc[j] += ta_inter_corners(dx, dy, dx2, dy2, A.x, A.y, B.x, B.y);
                        }
                    }
                    p = pn; q = qn; sl = sn; k = kn;
                }
                da = warp_sum(da);
#pragma unroll
                for (int j = 0; j < GT_TILE; ++j) acc[j] = warp_sum(acc[j]);
                // lane j finalises pair (i, gt0 + j)
                double mine = 0.0;
#pragma unroll
                for (int j = 0; j < GT_TILE; ++j) if (lane == j) mine = acc[j];
                if (lane < gcnt) {
                    double* o = out + (int64_t)i * G + gt0 + lane;
                    double inter = mine;
                    if (win > 0) inter += *o;
                    if (win == n_win - 1) {
                        const double uni = (da + ga_sh[lane]) - inter;
                        *o = uni > 0.0 ? inter / uni : 0.0;
                    } else {
                        *o = inter;
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// K1a': one thread per track pair (alt. IoU flavours, sequential reference association)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_track_iou_pair(TrackIouArgs a, int mode, int* flags) {
    const int grp = blockIdx.x;
    const int64_t d0 = a.grp_dt_off[grp], d1 = a.grp_dt_off[grp + 1];
    const int64_t g0 = a.grp_gt_off[grp], g1 = a.grp_gt_off[grp + 1];
    const int D = (int)(d1 - d0), G = (int)(g1 - g0);
    if (D == 0 || G == 0) return;
    double* out = a.iou + a.iou_off[grp];
    for (int e = threadIdx.x; e < D * G; e += blockDim.x) {
        const int i = e / G, j = e % G;
        const int64_t db0 = a.dt_off[d0 + i], db1 = a.dt_off[d0 + i + 1];
        const int64_t gb0 = a.gt_off[g0 + j], gb1 = a.gt_off[g0 + j + 1];
        int bad = 0;
        out[e] = ta_pair_iou_merge(a.dt_box + 4 * db0, a.dt_slot + db0, (int)(db1 - db0),
                                   a.gt_box + 4 * gb0, a.gt_slot + gb0, (int)(gb1 - gb0),
                                   mode, &bad);
        if (bad) atomicAdd(flags, 1);
    }
}

// ------------------------------------------------------------------------------------------
// K1b: per-(image, category) box IoU; one warp per group, lanes over the D*G entries
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_box_iou(int64_t n_groups, const int64_t* __restrict__ grp_dt_off,
          const int64_t* __restrict__ grp_gt_off, const double* __restrict__ dt_box,
          const double* __restrict__ gt_box, const int64_t* __restrict__ iou_off,
          double* __restrict__ iou) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t grp = warp0; grp < n_groups; grp += nwarps) {
        const int64_t d0 = grp_dt_off[grp], g0 = grp_gt_off[grp];
        const int D = (int)(grp_dt_off[grp + 1] - d0), G = (int)(grp_gt_off[grp + 1] - g0);
        const int n = D * G;
        if (n == 0) continue;
        double* out = iou + iou_off[grp];
        for (int e = lane; e < n; e += 32) {
            const int d = e / G, g = e - d * G;
            const double2 dp = *reinterpret_cast<const double2*>(dt_box + 4 * (d0 + d));
            const double2 dq = *reinterpret_cast<const double2*>(dt_box + 4 * (d0 + d) + 2);
            const double2 gp = *reinterpret_cast<const double2*>(gt_box + 4 * (g0 + g));
            const double2 gq = *reinterpret_cast<const double2*>(gt_box + 4 * (g0 + g) + 2);
            out[e] = ta_bb_iou(dp.x, dp.y, dq.x, dq.y, gp.x, gp.y, gq.x, gq.y);
        }
    }
}

// ------------------------------------------------------------------------------------------
// K2: greedy matching.  One CTA per group; each lane is one (range cfg, IoU threshold)
// matcher; the n_thr lanes of a cfg sit in one warp so their TP/FP bits are OR-reduced with
// a segmented redux and written as one uint32 per (cfg, detection).
// ------------------------------------------------------------------------------------------
struct MatchArgs {
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
    const int64_t* dt_id;
    const double* gt_a;
    const double* gt_b;
    const int32_t* gt_hp;
    const uint8_t* gt_flag;
    const int64_t* gt_id;
    int64_t sentinel;
    uint32_t* dt_tpfp;
    int32_t* num_gt;
    int32_t* dt_match_gt;
    uint8_t* gt_ignore_out;
    int cfgs_per_warp;
};

__global__ void k_match_greedy(MatchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = blockIdx.x;
    const int64_t d0 = a.grp_dt_off[grp], g0 = a.grp_gt_off[grp];
    const int D = (int)(a.grp_dt_off[grp + 1] - d0), G = (int)(a.grp_gt_off[grp + 1] - g0);
    if (D == 0 && G == 0) return;
    const int nthreads = blockDim.x;
    const int words = (G + 31) >> 5;
    // shared layout: taken[words][nthreads] u32, then gt_ig[n_cfg][G] u8
    uint32_t* taken = reinterpret_cast<uint32_t*>(smem_raw);
    uint8_t* gt_ig = smem_raw + (size_t)words * nthreads * sizeof(uint32_t);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cpw = a.cfgs_per_warp;
    const int cw = lane / a.n_thr;
    const int t = lane - cw * a.n_thr;
    const int cfg = warp * cpw + cw;
    const bool active = (cw < cpw) && (cfg < a.n_cfg);

    // ---- GT ignore flags for every cfg (eval.py:348-368 / lvis eval.py:201-217)
    for (int idx = threadIdx.x; idx < a.n_cfg * G; idx += nthreads) {
        const int c = idx / G, g = idx - c * G;
        const uint8_t ig = ta_gt_ignored(a.cfgs[c], a.gt_a[g0 + g], a.gt_b[g0 + g],
                                         a.gt_hp[g0 + g], a.gt_flag[g0 + g]);
        gt_ig[idx] = ig;
        if (a.gt_ignore_out) a.gt_ignore_out[(int64_t)c * a.n_gt + g0 + g] = ig;
    }
    for (int w = 0; w < words; ++w) taken[w * nthreads + threadIdx.x] = 0u;
    __syncthreads();
    // ---- number of non-ignored GT per (category, cfg) (eval.py:520-522)
    if (threadIdx.x < a.n_cfg && G > 0) {
        int cnt = 0;
        for (int g = 0; g < G; ++g) cnt += gt_ig[threadIdx.x * G + g] == 0;
        if (cnt) atomicAdd(&a.num_gt[(int64_t)a.grp_cat[grp] * a.n_cfg + threadIdx.x], cnt);
    }
    if (D == 0) return;

    const double thr = active ? a.thrs[t] : 2.0;
    const double* iou = a.iou + a.iou_off[grp];
    const uint8_t* my_ig = gt_ig + (active ? cfg : 0) * G;
    ta_range_cfg rc;
    if (active) rc = a.cfgs[cfg];
    uint32_t* my_taken = taken + threadIdx.x;

    for (int d = 0; d < D; ++d) {
        uint32_t bits = 0;
        if (active) {
            int m = -1;
            if (G > 0) m = ta_match_one(iou + (int64_t)d * G, G, my_ig, my_taken, nthreads, thr);
            const int64_t did = a.dt_id[d0 + d];
            bool unmatched = true, ig = false;
            if (m >= 0) {
                if (did > 0) my_taken[(m >> 5) * nthreads] |= 1u << (m & 31);   // eval.py:407,428
                unmatched = a.gt_id[g0 + m] == a.sentinel;                     // eval.py:427,443
                ig = my_ig[m] != 0;                                             // eval.py:425
            }
            if (unmatched && !ig)
                ig = ta_dt_unmatched_ignored(rc, a.dt_a[d0 + d], a.dt_b[d0 + d], a.dt_flag[d0 + d]);
            if (!ig) bits = unmatched ? (1u << (16 + t)) : (1u << t);
            if (a.dt_match_gt)
                a.dt_match_gt[((int64_t)cfg * a.n_thr + t) * a.n_dt + d0 + d] = m;
        }
        // OR the bits of the n_thr lanes of each cfg
        // segmented OR of the n_thr consecutive lanes of each cfg, result in the lane t == 0
        uint32_t word = bits;
        for (int o = 1; o < a.n_thr; o <<= 1) {
            const uint32_t other = __shfl_down_sync(0xffffffffu, word, o);
            if (lane + o < 32 && ((lane + o) / a.n_thr) == cw) word |= other;
        }
        if (active && t == 0) a.dt_tpfp[(int64_t)cfg * a.n_dt + d0 + d] = word;
    }
}

// ------------------------------------------------------------------------------------------
// K3: precision / recall accumulation.  One CTA per (category, range cfg), one warp per IoU
// threshold.  A single forward pass over the category's detections in score order:
// TP/FP prefix counts come from ballots; the interpolated precision at recall threshold k
// is the maximum precision over all TPs whose running TP count reaches t_k (the suffix-max
// envelope of eval.py:557-559 + searchsorted of :561), so each TP only has to raise one
// bucket and a final suffix max over the recall thresholds finishes the row.
// ------------------------------------------------------------------------------------------
__global__ void k_pr_accumulate(int n_cat, const int64_t* __restrict__ cat_dt_off,
                                const int32_t* __restrict__ acc_perm, int64_t n_dt,
                                const uint32_t* __restrict__ dt_tpfp,
                                const int32_t* __restrict__ num_gt, int n_thr, int n_cfg,
                                int n_rec, const double* __restrict__ rec_thrs,
                                double* __restrict__ precision, double* __restrict__ recall,
                                int64_t* __restrict__ tp_cnt, int64_t* __restrict__ fp_cnt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int64_t* tk = reinterpret_cast<int64_t*>(smem_raw);                          // [n_rec]
    unsigned long long* bucket = reinterpret_cast<unsigned long long*>(tk + n_rec);  // [n_thr][n_rec]
    const int c = blockIdx.x / n_cfg, cfg = blockIdx.x - c * n_cfg;
    const int lane = threadIdx.x & 31, t = threadIdx.x >> 5;
    const int ngt = num_gt[(int64_t)c * n_cfg + cfg];
    const int64_t cell = ((int64_t)t * n_cat + c) * n_cfg + cfg;
    if (ngt == 0) {   // eval.py:522-525: cell keeps its -1 initialisation
        for (int k = lane; k < n_rec; k += 32)
            precision[(((int64_t)t * n_rec + k) * n_cat + c) * n_cfg + cfg] = -1.0;
        if (lane == 0) {
            recall[cell] = -1.0;
            if (tp_cnt) tp_cnt[cell] = 0;
            if (fp_cnt) fp_cnt[cell] = 0;
        }
        return;
    }
    for (int k = threadIdx.x; k < n_rec; k += blockDim.x) tk[k] = ta_min_tp_for_recall(rec_thrs[k], ngt);
    for (int k = threadIdx.x; k < n_thr * n_rec; k += blockDim.x) bucket[k] = 0ull;
    __syncthreads();

    const int64_t p0 = cat_dt_off[c], p1 = cat_dt_off[c + 1];
    const uint32_t* row = dt_tpfp + (int64_t)cfg * n_dt;
    unsigned long long* my_bucket = bucket + (int64_t)t * n_rec;
    int64_t tp_run = 0, fp_run = 0;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int64_t base = p0; base < p1; base += 32) {
        const int64_t p = base + lane;
        uint32_t w = 0;
        if (p < p1) w = row[acc_perm[p]];
        const unsigned tpb = (w >> t) & 1u, fpb = (w >> (16 + t)) & 1u;
        const unsigned bt = __ballot_sync(0xffffffffu, tpb), bf = __ballot_sync(0xffffffffu, fpb);
        if (tpb) {
            const int64_t tc = tp_run + __popc(bt & lt_mask) + 1;
            const int64_t fc = fp_run + __popc(bf & lt_mask);
            const double pr = ta_precision_at(tc, fc);
            // number of recall thresholds reachable with tc true positives, minus one
            int lo = 0, hi = n_rec;   // first k with tk[k] > tc
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (tk[mid] <= tc) lo = mid + 1; else hi = mid;
            }
            if (lo > 0) atomicMax(&my_bucket[lo - 1], (unsigned long long)__double_as_longlong(pr));
        }
        tp_run += __popc(bt);
        fp_run += __popc(bf);
    }
    __syncwarp();
    if (lane == 0) {
        unsigned long long best = 0ull;   // bit pattern of +0.0; precisions are positive
        for (int k = n_rec - 1; k >= 0; --k) {
            const unsigned long long v = my_bucket[k];
            best = v > best ? v : best;
            precision[(((int64_t)t * n_rec + k) * n_cat + c) * n_cfg + cfg] =
                __longlong_as_double((long long)best);
        }
        // eval.py:543-547: recall = rc[-1] if there are detections else 0
        recall[cell] = (p1 > p0) ? (double)tp_run / (double)ngt : 0.0;
        if (tp_cnt) tp_cnt[cell] = tp_run;
        if (fp_cnt) fp_cnt[cell] = fp_run;
    }
}

// ------------------------------------------------------------------------------------------
// C ABI: launches
// ------------------------------------------------------------------------------------------
static int check_launch(ta_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s launch failed: %s", what, cudaGetErrorString(e));
        return TA_ERR_CUDA;
    }
    ctx->launches++;
    return TA_OK;
}

extern "C" int ta_track_iou(ta_ctx* ctx, void* stream, int mode, int64_t n_groups,
                            const int64_t* grp_dt_off, const int64_t* grp_gt_off,
                            const int64_t* dt_trk_off, const double* dt_box, const int32_t* dt_slot,
                            const int64_t* gt_trk_off, const double* gt_box, const int32_t* gt_slot,
                            int32_t n_slots_max, const int64_t* iou_off, double* iou_out) {
    if (!ctx) return set_err(TA_ERR_INVALID, "ta_track_iou: ctx is NULL");
    if (n_groups < 0 || n_slots_max < 0) return set_err(TA_ERR_INVALID, "ta_track_iou: negative size");
    if (mode < 0 || mode > 3) return set_err(TA_ERR_INVALID, "ta_track_iou: unknown mode %s%lld", "", mode);
    if (n_groups == 0) return TA_OK;
    if (n_groups > INT_MAX) return set_err(TA_ERR_TOO_LARGE, "ta_track_iou: too many groups");
    TA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    TrackIouArgs a{grp_dt_off, grp_gt_off, dt_trk_off, dt_box, dt_slot,
                   gt_trk_off, gt_box, gt_slot, iou_off, iou_out, 0};
    if (mode == TA_IOU_3D) {
        int S = 64;
        while (S < n_slots_max && S < 384) S += (S < 128 ? 64 : 128);
        if (S > 384) S = 384;
        a.S = S;
        const size_t smem = (size_t)2 * GT_TILE * (S + 1) * sizeof(double2);
        TA_CUDA(cudaFuncSetAttribute(k_track_iou_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
        k_track_iou_tiled<<<(unsigned)n_groups, TI_WARPS * 32, smem, st>>>(a);
        return check_launch(ctx, "k_track_iou_tiled");
    }
    TA_CUDA(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), st));
    k_track_iou_pair<<<(unsigned)n_groups, 128, 0, st>>>(a, mode, ctx->d_flags);
    int rc = check_launch(ctx, "k_track_iou_pair");
    if (rc) return rc;
    if (mode == TA_IOU_3D_SEQ) {
        int bad = 0;
        TA_CUDA(cudaMemcpyAsync(&bad, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
        TA_CUDA(cudaStreamSynchronize(st));
        if (bad) return set_err(TA_ERR_ASSERT, "track IoU: intersection exceeds union in %s%lld pairs", "", bad);
    }
    return TA_OK;
}

extern "C" int ta_box_iou(ta_ctx* ctx, void* stream, int64_t n_groups,
                          const int64_t* grp_dt_off, const int64_t* grp_gt_off,
                          const double* dt_box, const double* gt_box,
                          const int64_t* iou_off, double* iou_out) {
    if (!ctx) return set_err(TA_ERR_INVALID, "ta_box_iou: ctx is NULL");
    if (n_groups < 0) return set_err(TA_ERR_INVALID, "ta_box_iou: negative size");
    if (n_groups == 0) return TA_OK;
    TA_CUDA(cudaSetDevice(ctx->device));
    const int64_t warps_needed = n_groups;
    int64_t blocks = (warps_needed + 7) / 8;
    const int64_t cap = (int64_t)ctx->sm_count * 32;   // persistent-ish: 8 CTAs of 8 warps per SM
    if (blocks > cap) blocks = cap;
    k_box_iou<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        n_groups, grp_dt_off, grp_gt_off, dt_box, gt_box, iou_off, iou_out);
    return check_launch(ctx, "k_box_iou");
}

extern "C" int ta_match_greedy(ta_ctx* ctx, void* stream, int64_t n_groups,
                               const int64_t* grp_dt_off, const int64_t* grp_gt_off,
                               const int32_t* grp_cat, const int64_t* iou_off, const double* iou,
                               int32_t n_thr, const double* iou_thrs,
                               int32_t n_cfg, const ta_range_cfg* cfgs,
                               int64_t n_dt, const double* dt_attr_a, const double* dt_attr_b,
                               const uint8_t* dt_flag, const int64_t* dt_id,
                               int64_t n_gt, const double* gt_attr_a, const double* gt_attr_b,
                               const int32_t* gt_hp, const uint8_t* gt_flag, const int64_t* gt_id,
                               int64_t sentinel, int32_t g_max, uint32_t* dt_tpfp, int32_t* num_gt,
                               int32_t* dt_match_gt, uint8_t* gt_ignore_out) {
    if (!ctx) return set_err(TA_ERR_INVALID, "ta_match_greedy: ctx is NULL");
    if (n_thr < 1 || n_thr > TA_MAX_THRS)
        return set_err(TA_ERR_INVALID, "ta_match_greedy: n_thr must be in [1,16], got %s%lld", "", n_thr);
    if (n_cfg < 1 || n_groups < 0 || g_max < 0) return set_err(TA_ERR_INVALID, "ta_match_greedy: bad sizes");
    if (n_groups == 0) return TA_OK;
    if (n_groups > INT_MAX) return set_err(TA_ERR_TOO_LARGE, "ta_match_greedy: too many groups");
    TA_CUDA(cudaSetDevice(ctx->device));
    const int cpw = 32 / n_thr;
    const int warps = (n_cfg + cpw - 1) / cpw;
    if (warps * 32 > 1024) return set_err(TA_ERR_TOO_LARGE, "ta_match_greedy: too many range cfgs");
    const int threads = warps * 32;
    const int words = (g_max + 31) / 32;
    const size_t smem = (size_t)words * threads * sizeof(uint32_t) + (size_t)n_cfg * g_max;
    if (smem > (size_t)ctx->smem_optin)
        return set_err(TA_ERR_TOO_LARGE,
                       "ta_match_greedy: a group with %s%lld ground-truth entities does not fit in shared memory",
                       "", (long long)g_max);
    MatchArgs a{grp_dt_off, grp_gt_off, grp_cat, iou_off, iou, n_thr, iou_thrs, n_cfg, cfgs,
                n_dt, n_gt, dt_attr_a, dt_attr_b, dt_flag, dt_id, gt_attr_a, gt_attr_b, gt_hp,
                gt_flag, gt_id, sentinel, dt_tpfp, num_gt, dt_match_gt, gt_ignore_out, cpw};
    if (smem > 48 * 1024)
        TA_CUDA(cudaFuncSetAttribute(k_match_greedy, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    k_match_greedy<<<(unsigned)n_groups, threads, smem, (cudaStream_t)stream>>>(a);
    return check_launch(ctx, "k_match_greedy");
}

extern "C" int ta_pr_accumulate(ta_ctx* ctx, void* stream, int32_t n_cat, const int64_t* cat_dt_off,
                                const int32_t* acc_perm, int64_t n_dt, const uint32_t* dt_tpfp,
                                const int32_t* num_gt, int32_t n_thr, int32_t n_cfg,
                                int32_t n_rec, const double* rec_thrs,
                                double* precision, double* recall, int64_t* tp_cnt, int64_t* fp_cnt) {
    if (!ctx) return set_err(TA_ERR_INVALID, "ta_pr_accumulate: ctx is NULL");
    if (n_thr < 1 || n_thr > TA_MAX_THRS || n_cfg < 1 || n_rec < 1 || n_cat < 0)
        return set_err(TA_ERR_INVALID, "ta_pr_accumulate: bad sizes");
    if (n_cat == 0) return TA_OK;
    TA_CUDA(cudaSetDevice(ctx->device));
    const size_t smem = (size_t)n_rec * 8 + (size_t)n_thr * n_rec * 8;
    if (smem > (size_t)ctx->smem_optin) return set_err(TA_ERR_TOO_LARGE, "ta_pr_accumulate: too many recall thresholds");
    TA_CUDA(cudaFuncSetAttribute(k_pr_accumulate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_pr_accumulate<<<(unsigned)((int64_t)n_cat * n_cfg), n_thr * 32, smem, (cudaStream_t)stream>>>(
        n_cat, cat_dt_off, acc_perm, n_dt, dt_tpfp, num_gt, n_thr, n_cfg, n_rec, rec_thrs,
        precision, recall, tp_cnt, fp_cnt);
    return check_launch(ctx, "k_pr_accumulate");
}

// ------------------------------------------------------------------------------------------
// Host-buffer pipeline: H2D -> IoU -> match -> accumulate -> D2H on the context's stream.
// Device memory comes from the stream-ordered pool (cudaMallocAsync), so repeated calls reuse
// the same pages without a device-wide synchronisation.
// ------------------------------------------------------------------------------------------
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
        if (count) {
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
        return set_err(TA_ERR_INVALID, "ta_eval_plan_host: NULL argument");
    TA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->own_stream;
    const bool track = pl->dt_trk_off != nullptr;
    const int64_t G = pl->n_groups;
    int rc = TA_OK;
    {
        DevArena ar(st);
        const int64_t *grp_dt_off, *grp_gt_off, *iou_off, *cat_dt_off, *dt_trk_off = nullptr,
                      *gt_trk_off = nullptr, *dt_id, *gt_id;
        const int32_t *grp_cat, *acc_perm, *dt_slot = nullptr, *gt_slot = nullptr, *gt_hp;
        const double *dt_box, *gt_box, *dt_a, *dt_b, *gt_a, *gt_b, *thrs, *recs;
        const uint8_t *dt_flag, *gt_flag;
        const ta_range_cfg* cfgs;
        TA_CUDA(ar.upload(&grp_dt_off, pl->grp_dt_off, G + 1));
        TA_CUDA(ar.upload(&grp_gt_off, pl->grp_gt_off, G + 1));
        TA_CUDA(ar.upload(&iou_off, pl->iou_off, G + 1));
        TA_CUDA(ar.upload(&cat_dt_off, pl->cat_dt_off, (size_t)pl->n_cat + 1));
        TA_CUDA(ar.upload(&grp_cat, pl->grp_cat, G));
        TA_CUDA(ar.upload(&acc_perm, pl->acc_perm, pl->n_dt));
        TA_CUDA(ar.upload(&dt_box, pl->dt_box, (size_t)pl->n_dt_boxes * 4));
        TA_CUDA(ar.upload(&gt_box, pl->gt_box, (size_t)pl->n_gt_boxes * 4));
        if (track) {
            TA_CUDA(ar.upload(&dt_trk_off, pl->dt_trk_off, pl->n_dt + 1));
            TA_CUDA(ar.upload(&gt_trk_off, pl->gt_trk_off, pl->n_gt + 1));
            TA_CUDA(ar.upload(&dt_slot, pl->dt_slot, pl->n_dt_boxes));
            TA_CUDA(ar.upload(&gt_slot, pl->gt_slot, pl->n_gt_boxes));
        }
        TA_CUDA(ar.upload(&dt_a, pl->dt_attr_a, pl->n_dt));
        TA_CUDA(ar.upload(&dt_b, pl->dt_attr_b, pl->n_dt));
        TA_CUDA(ar.upload(&gt_a, pl->gt_attr_a, pl->n_gt));
        TA_CUDA(ar.upload(&gt_b, pl->gt_attr_b, pl->n_gt));
        TA_CUDA(ar.upload(&dt_flag, pl->dt_flag, pl->n_dt));
        TA_CUDA(ar.upload(&gt_flag, pl->gt_flag, pl->n_gt));
        TA_CUDA(ar.upload(&gt_hp, pl->gt_hp, pl->n_gt));
        TA_CUDA(ar.upload(&dt_id, pl->dt_id, pl->n_dt));
        TA_CUDA(ar.upload(&gt_id, pl->gt_id, pl->n_gt));
        TA_CUDA(ar.upload(&thrs, pl->iou_thrs, pl->n_thr));
        TA_CUDA(ar.upload(&recs, pl->rec_thrs, pl->n_rec));
        TA_CUDA(ar.upload(&cfgs, pl->cfgs, pl->n_cfg));

        const int64_t n_iou = pl->iou_off[G];
        const size_t n_cell = (size_t)pl->n_thr * pl->n_cat * pl->n_cfg;
        const size_t n_prec = n_cell * pl->n_rec;
        void *d_iou, *d_tpfp, *d_numgt, *d_prec, *d_rec, *d_tp, *d_fp;
        TA_CUDA(ar.alloc(&d_iou, (size_t)n_iou * 8));
        TA_CUDA(ar.alloc(&d_tpfp, (size_t)pl->n_cfg * pl->n_dt * 4));
        TA_CUDA(ar.alloc(&d_numgt, (size_t)pl->n_cat * pl->n_cfg * 4));
        TA_CUDA(ar.alloc(&d_prec, n_prec * 8));
        TA_CUDA(ar.alloc(&d_rec, n_cell * 8));
        TA_CUDA(ar.alloc(&d_tp, n_cell * 8));
        TA_CUDA(ar.alloc(&d_fp, n_cell * 8));
        TA_CUDA(cudaMemsetAsync(d_numgt, 0, (size_t)pl->n_cat * pl->n_cfg * 4, st));

        if (track)
            rc = ta_track_iou(ctx, st, pl->iou_mode, G, grp_dt_off, grp_gt_off, dt_trk_off, dt_box,
                              dt_slot, gt_trk_off, gt_box, gt_slot, pl->n_slots_max, iou_off,
                              (double*)d_iou);
        else
            rc = ta_box_iou(ctx, st, G, grp_dt_off, grp_gt_off, dt_box, gt_box, iou_off,
                            (double*)d_iou);
        if (rc == TA_OK)
            rc = ta_match_greedy(ctx, st, G, grp_dt_off, grp_gt_off, grp_cat, iou_off,
                                 (const double*)d_iou, pl->n_thr, thrs, pl->n_cfg, cfgs, pl->n_dt,
                                 dt_a, dt_b, dt_flag, dt_id, pl->n_gt, gt_a, gt_b, gt_hp, gt_flag,
                                 gt_id, pl->sentinel, pl->g_max, (uint32_t*)d_tpfp, (int32_t*)d_numgt,
                                 nullptr, nullptr);
        if (rc == TA_OK)
            rc = ta_pr_accumulate(ctx, st, pl->n_cat, cat_dt_off, acc_perm, pl->n_dt,
                                  (const uint32_t*)d_tpfp, (const int32_t*)d_numgt, pl->n_thr,
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
    TA_CUDA(cudaStreamSynchronize(st));
    return rc;
}
