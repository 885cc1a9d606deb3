// ta_mask.cpp — run-length mask codec of the segm evaluation path (host side, part of
// libta_ingest.so; C ABI in include/ta_mask.h).
//
// The reference converts annotations to column-major run lengths with pycocotools
// (lvis_amodal/lvis.py:155-192, results.py:58-66); the in-tree copy of that C code is
// visualization/tao/third_party/pysot/training_dataset/coco/pycocotools/common/maskApi.c.
// The conversions here are written from that algorithm's definition and must reproduce its
// run structure exactly (the IoU pre-filter of rleIou, :80-82, depends on the run boundaries
// through rleToBbox), so every step cites the lines it follows.  Only conversion lives on the
// host; the mask IoU of the evaluation is the CUDA kernel in ta_rle.cu.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "ta_mask.h"

namespace {

thread_local std::string g_err;

struct Mask {
    uint64_t h = 0, w = 0;
    std::vector<uint32_t> c;      // run lengths, zeros first
};

// C's (int) conversion of a double as x86-64 performs it (cvttsd2si): truncation toward zero,
// INT_MIN for NaN / out of range — maskApi.c relies on it for degenerate (repeated) points
inline int to_int(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT32_MIN;
    return (int)v;
}

// maskApi.c:164-216.  Boundary of the polygon sampled on a 5x finer grid, the crossings of
// pixel-column boundaries turned into run starts / ends, sorted, differenced.
void poly_to_mask(const double* xy, int64_t k, uint64_t h, uint64_t w, Mask& out) {
    const double scale = 5.0;
    std::vector<int> px((size_t)k + 1), py((size_t)k + 1);
    for (int64_t j = 0; j < k; ++j) {
        px[j] = to_int(scale * xy[2 * j] + .5);           // :168
        py[j] = to_int(scale * xy[2 * j + 1] + .5);       // :169
    }
    px[k] = px[0];
    py[k] = py[0];
    // dense integer points along every edge (:170-184)
    std::vector<int> u, v;
    for (int64_t j = 0; j < k; ++j) {
        int xs = px[j], xe = px[j + 1], ys = py[j], ye = py[j + 1];
        const int dx = abs(xe - xs), dy = abs(ys - ye);
        const bool steep = dx < dy;
        const bool flip = (!steep && xs > xe) || (steep && ys > ye);
        if (flip) { std::swap(xs, xe); std::swap(ys, ye); }
        const double s = !steep ? (double)(ye - ys) / dx : (double)(xe - xs) / dy;
        const int n = steep ? dy : dx;
        for (int d = 0; d <= n; ++d) {
            const int t = flip ? n - d : d;
            if (!steep) { u.push_back(t + xs); v.push_back(to_int(ys + s * t + .5)); }
            else        { v.push_back(t + ys); u.push_back(to_int(xs + s * t + .5)); }
        }
    }
    // column crossings, back on the pixel grid (:186-194)
    std::vector<uint32_t> a;
    const double wlim = (double)(w - 1);          // unsigned arithmetic first, as `xd>w-1` does
    for (size_t j = 1; j < u.size(); ++j) {
        if (u[j] == u[j - 1]) continue;
        double xd = (double)(u[j] < u[j - 1] ? u[j] : u[j] - 1);
        xd = (xd + .5) / scale - .5;
        if (floor(xd) != xd || xd < 0 || xd > wlim) continue;
        double yd = (double)(v[j] < v[j - 1] ? v[j] : v[j - 1]);
        yd = (yd + .5) / scale - .5;
        if (yd < 0) yd = 0; else if (yd > (double)h) yd = (double)h;
        yd = ceil(yd);
        a.push_back((uint32_t)(to_int(xd) * (int)h + to_int(yd)));      // :197
    }
    a.push_back((uint32_t)(h * w));                                      // :198
    std::sort(a.begin(), a.end());                                       // :199 (qsort, uint order)
    uint32_t prev = 0;
    for (uint32_t& x : a) { const uint32_t t = x; x -= prev; prev = t; } // :200
    // zero-length runs cancel against their neighbours (:201-203)
    out.h = h; out.w = w; out.c.clear();
    size_t j = 0;
    out.c.push_back(a[j++]);
    while (j < a.size()) {
        if (a[j] > 0) out.c.push_back(a[j++]);
        else { ++j; if (j < a.size()) out.c.back() += a[j++]; }
    }
}

// maskApi.c:50-71 with intersect = 0: running union of the masks, pairwise run walk
void union_masks(const std::vector<Mask>& parts, Mask& out) {
    if (parts.empty()) { out = Mask(); return; }
    if (parts.size() == 1) { out = parts[0]; return; }
    uint64_t h = parts[0].h, w = parts[0].w;
    std::vector<uint32_t> cur = parts[0].c, nxt;
    for (size_t i = 1; i < parts.size(); ++i) {
        const Mask& B = parts[i];
        if (B.h != h || B.w != w) { h = w = 0; cur.clear(); break; }      // :59
        const std::vector<uint32_t>& A = cur;
        nxt.clear();
        uint32_t ca = A.empty() ? 0u : A[0], cb = B.c.empty() ? 0u : B.c[0], cc = 0, ct = 1;
        bool v = false, va = false, vb = false;
        size_t ia = 1, ib = 1;
        while (ct > 0) {
            const uint32_t c = ca < cb ? ca : cb;
            cc += c; ct = 0;
            ca -= c; if (!ca && ia < A.size()) { ca = A[ia++]; va = !va; } ct += ca;
            cb -= c; if (!cb && ib < B.c.size()) { cb = B.c[ib++]; vb = !vb; } ct += cb;
            const bool vp = v;
            v = va || vb;
            if (v != vp || ct == 0) { nxt.push_back(cc); cc = 0; }
        }
        cur.swap(nxt);
    }
    out.h = h; out.w = w; out.c = cur;
}

// maskApi.c:233-246: 6 bits per character (5 payload + continuation), offset 48; runs after
// the third are stored as differences to the run two places back
bool string_to_mask(const char* s, int64_t len, uint64_t h, uint64_t w, Mask& out) {
    out.h = h; out.w = w; out.c.clear();
    int64_t p = 0;
    while (p < len && s[p]) {
        long x = 0;
        int k = 0;
        bool more = true;
        while (more) {
            if (p >= len) { g_err = "ValueError: truncated RLE counts string"; return false; }
            const char c = (char)(s[p] - 48);
            x |= (long)(c & 0x1f) << 5 * k;
            more = (c & 0x20) != 0;
            ++p; ++k;
            if (!more && (c & 0x10)) x |= -1L << 5 * k;
        }
        const size_t m = out.c.size();
        if (m > 2) x += (long)out.c[m - 2];
        out.c.push_back((uint32_t)x);
    }
    return true;
}

// maskApi.c:218-231
std::string mask_to_string(const Mask& R) {
    std::string s;
    for (size_t i = 0; i < R.c.size(); ++i) {
        long x = (long)R.c[i];
        if (i > 2) x -= (long)R.c[i - 2];
        bool more = true;
        while (more) {
            char c = (char)(x & 0x1f);
            x >>= 5;
            more = (c & 0x10) ? x != -1 : x != 0;
            if (more) c |= 0x20;
            s.push_back((char)(c + 48));
        }
    }
    return s;
}

// maskApi.c:133-151 (this vintage derives the box from run starts / ends only)
void mask_bbox(const Mask& R, double* bb) {
    const uint32_t h = (uint32_t)R.h, w = (uint32_t)R.w;
    const size_t m = (R.c.size() / 2) * 2;
    if (m == 0 || h == 0) { bb[0] = bb[1] = bb[2] = bb[3] = 0; return; }
    uint32_t xs = w, ys = h, xe = 0, ye = 0, cc = 0;
    for (size_t j = 0; j < m; ++j) {
        cc += R.c[j];
        const uint32_t t = cc - (uint32_t)(j % 2);
        const uint32_t y = t % h, x = (t - y) / h;
        xs = std::min(xs, x); xe = std::max(xe, x);
        ys = std::min(ys, y); ye = std::max(ye, y);
    }
    bb[0] = xs; bb[2] = xe - xs + 1;
    bb[1] = ys; bb[3] = ye - ys + 1;
}

uint32_t mask_area(const Mask& R) {                  // maskApi.c:73-76
    uint32_t a = 0;
    for (size_t j = 1; j < R.c.size(); j += 2) a += R.c[j];
    return a;
}

}  // namespace

struct ta_rle_pool {
    std::vector<Mask> masks;
    int64_t total = 0;
    int64_t push(Mask&& m) {
        total += (int64_t)m.c.size();
        masks.push_back(std::move(m));
        return (int64_t)masks.size() - 1;
    }
};

extern "C" {

ta_rle_pool* ta_rle_pool_create(void) { return new ta_rle_pool(); }
void ta_rle_pool_destroy(ta_rle_pool* p) { delete p; }
const char* ta_mask_error(void) { return g_err.c_str(); }

int64_t ta_rle_pool_add_polygons(ta_rle_pool* p, int64_t n_parts, const int64_t* part_off,
                                 const double* xy, int64_t h, int64_t w) {
    if (!p || n_parts < 0 || (n_parts && (!part_off || !xy)) || h < 0 || w < 0) {
        g_err = "ValueError: ta_rle_pool_add_polygons: bad argument";
        return -1;
    }
    std::vector<Mask> parts((size_t)n_parts);
    for (int64_t i = 0; i < n_parts; ++i) {
        const int64_t n = part_off[i + 1] - part_off[i];
        if (n < 2) { g_err = "ValueError: polygon part with fewer than one point"; return -1; }
        poly_to_mask(xy + part_off[i], n / 2, (uint64_t)h, (uint64_t)w, parts[i]);   // int(len(p)/2), _mask.pyx:266
    }
    Mask m;
    union_masks(parts, m);
    return p->push(std::move(m));
}

int64_t ta_rle_pool_add_boxes(ta_rle_pool* p, int64_t n, const double* boxes,
                              const int64_t* h, const int64_t* w) {
    if (!p || n < 0 || (n && (!boxes || !h || !w))) {
        g_err = "ValueError: ta_rle_pool_add_boxes: bad argument";
        return -1;
    }
    const int64_t first = (int64_t)p->masks.size();
    for (int64_t i = 0; i < n; ++i) {
        const double xs = boxes[4 * i], xe = xs + boxes[4 * i + 2];
        const double ys = boxes[4 * i + 1], ye = ys + boxes[4 * i + 3];
        const double xy[8] = {xs, ys, xs, ye, xe, ye, xe, ys};           // maskApi.c:156-158
        Mask m;
        poly_to_mask(xy, 4, (uint64_t)h[i], (uint64_t)w[i], m);
        p->push(std::move(m));
    }
    return first;
}

int64_t ta_rle_pool_add_counts(ta_rle_pool* p, int64_t m, const uint32_t* counts, int64_t h, int64_t w) {
    if (!p || m < 0 || (m && !counts)) { g_err = "ValueError: ta_rle_pool_add_counts: bad argument"; return -1; }
    Mask k;
    k.h = (uint64_t)h; k.w = (uint64_t)w;
    k.c.assign(counts, counts + m);
    return p->push(std::move(k));
}

int64_t ta_rle_pool_add_string(ta_rle_pool* p, const char* s, int64_t len, int64_t h, int64_t w) {
    if (!p || !s || len < 0) { g_err = "ValueError: ta_rle_pool_add_string: bad argument"; return -1; }
    Mask k;
    if (!string_to_mask(s, len, (uint64_t)h, (uint64_t)w, k)) return -1;
    return p->push(std::move(k));
}

int64_t ta_rle_pool_size(const ta_rle_pool* p) { return p ? (int64_t)p->masks.size() : 0; }
int64_t ta_rle_pool_total_counts(const ta_rle_pool* p) { return p ? p->total : 0; }

int ta_rle_pool_export(const ta_rle_pool* p, int64_t* off, uint32_t* counts, uint32_t* hw,
                       double* bbox, uint32_t* area) {
    if (!p) { g_err = "ValueError: ta_rle_pool_export: pool is NULL"; return -1; }
    int64_t o = 0;
    for (size_t i = 0; i < p->masks.size(); ++i) {
        const Mask& m = p->masks[i];
        if (off) off[i] = o;
        if (counts && !m.c.empty()) memcpy(counts + o, m.c.data(), m.c.size() * sizeof(uint32_t));
        if (hw) { hw[2 * i] = (uint32_t)m.h; hw[2 * i + 1] = (uint32_t)m.w; }
        if (bbox) mask_bbox(m, bbox + 4 * i);
        if (area) area[i] = mask_area(m);
        o += (int64_t)m.c.size();
    }
    if (off) off[p->masks.size()] = o;
    return 0;
}

int64_t ta_rle_pool_to_string(const ta_rle_pool* p, int64_t i, char* buf, int64_t cap) {
    if (!p || i < 0 || i >= (int64_t)p->masks.size()) { g_err = "IndexError: mask index out of range"; return INT64_MIN; }
    const std::string s = mask_to_string(p->masks[(size_t)i]);
    if ((int64_t)s.size() + 1 > cap || !buf) return -((int64_t)s.size() + 1);
    memcpy(buf, s.c_str(), s.size() + 1);
    return (int64_t)s.size();
}

}  // extern "C"
