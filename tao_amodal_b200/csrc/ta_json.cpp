// ta_json.cpp — single-pass columnar reader of the two JSON files of the evaluation CLI
// (host-only; built into libta_ingest.so with g++, C ABI in include/ta_ingest.h).
//
// Replaces what the reference does with json.load + per-dict Python loops before any
// arithmetic: Tao._create_index / LVIS._create_index (tao_amodal/evaluation/tao_amodal/
// tao.py:108-160, lvis_amodal/lvis.py:37-61) and the result-list walks of TaoResults /
// LVISResults (results.py:38-109, lvis_amodal/results.py:29-71).  The reference parses each
// file twice and deep-copies the annotation dict twice; here each file is memory-mapped and
// scanned once, and only the fields the evaluation reads are kept, as flat columns
// (tao_amodal_b200/columnar.py: GtColumns / DtColumns).
//
// Numbers are converted with std::from_chars (correctly rounded, same doubles as Python's
// float()); ints stay exact int64.  Unknown keys and nested values are skipped structurally.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "ta_ingest.h"

namespace {

thread_local char g_err[512] = "";
thread_local int g_last_workers = 0;       // workers that produced the last document of this thread

struct Fail {
    std::string msg;
};

struct Column {
    int elem = 8;                 // bytes per element
    std::vector<char> data;
    template <typename T>
    void push(T v) {
        const size_t o = data.size();
        data.resize(o + sizeof(T));
        memcpy(data.data() + o, &v, sizeof(T));
    }
};

struct Doc {
    std::map<std::string, Column> col;
    Column& c(const char* name, int elem) {
        Column& x = col[name];
        x.elem = elem;
        return x;
    }
};

struct Parser {
    const char* p;
    const char* end;

    [[noreturn]] void fail(const char* what) const {
        char buf[160];
        snprintf(buf, sizeof(buf), "JSONDecodeError: %s at byte %lld", what, (long long)(p - begin));
        throw Fail{buf};
    }
    const char* begin;

    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
    }
    bool eat(char ch) {
        ws();
        if (p < end && *p == ch) { ++p; return true; }
        return false;
    }
    void need(char ch) {
        if (!eat(ch)) fail("unexpected character");
    }
    // string without unescaping: [s, e) is the raw content between the quotes
    void raw_string(const char*& s, const char*& e) {
        ws();
        if (p >= end || *p != '"') fail("expected string");
        s = ++p;
        while (p < end && *p != '"') {
            if (*p == '\\') ++p;
            ++p;
        }
        if (p >= end) fail("unterminated string");
        e = p++;
    }
    bool key_is(const char* s, const char* e, const char* lit) const {
        const size_t n = strlen(lit);
        return (size_t)(e - s) == n && memcmp(s, lit, n) == 0;
    }
    // number / true / false / null / NaN / Infinity  ->  double (+ exact int when integral text)
    struct Num { double d; int64_t i; bool is_int; bool is_null; };
    Num scalar() {
        ws();
        if (p >= end) fail("unexpected end");
        Num r{0.0, 0, false, false};
        const char c = *p;
        auto lit = [&](const char* w) {
            const size_t n = strlen(w);
            if ((size_t)(end - p) >= n && memcmp(p, w, n) == 0) { p += n; return true; }
            return false;
        };
        if (c == 't') { if (!lit("true")) fail("bad literal"); r.d = 1; r.i = 1; r.is_int = true; return r; }
        if (c == 'f') { if (!lit("false")) fail("bad literal"); r.is_int = true; return r; }
        if (c == 'n') { if (!lit("null")) fail("bad literal"); r.is_null = true; return r; }
        if (c == 'N') { if (!lit("NaN")) fail("bad literal"); r.d = NAN; return r; }
        if (c == 'I') { if (!lit("Infinity")) fail("bad literal"); r.d = INFINITY; return r; }
        if (c == '-' && p + 1 < end && p[1] == 'I') { ++p; if (!lit("Infinity")) fail("bad literal"); r.d = -INFINITY; return r; }
        const char* s = p;
        const char* q = p;
        if (q < end && *q == '-') ++q;
        bool integral = true;
        while (q < end && ((*q >= '0' && *q <= '9') || *q == '.' || *q == 'e' || *q == 'E' || *q == '+' || *q == '-')) {
            if (*q == '.' || *q == 'e' || *q == 'E') integral = false;
            ++q;
        }
        if (q == s) fail("expected a number");
        if (integral) {
            auto res = std::from_chars(s, q, r.i);
            if (res.ec == std::errc() && res.ptr == q) {
                r.is_int = true;
                r.d = (double)r.i;
                p = q;
                return r;
            }
        }
        auto res = std::from_chars(s, q, r.d);
        if (res.ec != std::errc() || res.ptr != q) fail("bad number");
        r.i = (int64_t)r.d;
        p = q;
        return r;
    }
    void skip_value() {
        ws();
        if (p >= end) fail("unexpected end");
        if (*p == '"') { const char *s, *e; raw_string(s, e); return; }
        if (*p == '{') {
            ++p;
            if (eat('}')) return;
            do {
                const char *s, *e;
                raw_string(s, e);
                need(':');
                skip_value();
            } while (eat(','));
            need('}');
            return;
        }
        if (*p == '[') {
            ++p;
            if (eat(']')) return;
            do skip_value(); while (eat(','));
            need(']');
            return;
        }
        scalar();
    }
    int64_t as_int() {
        const Num n = scalar();
        if (n.is_null) fail("null where an integer is required");
        return n.is_int ? n.i : (int64_t)n.d;
    }
    double as_double() {
        const Num n = scalar();
        if (n.is_null) fail("null where a number is required");
        return n.d;
    }
    bool truthy() {          // Python truthiness of a scalar (`if x.get("ignore", 0)`)
        ws();
        if (p < end && (*p == '"' || *p == '[' || *p == '{')) {
            const char* s = p;
            skip_value();
            return (p - s) > 2;      // non-empty string / list / dict
        }
        const Num n = scalar();
        return !n.is_null && n.d != 0.0;
    }
    // [int, int, ...] appended to vals; returns count
    int64_t int_list(Column& vals) {
        need('[');
        int64_t n = 0;
        if (eat(']')) return 0;
        do { vals.push<int64_t>(as_int()); ++n; } while (eat(','));
        need(']');
        return n;
    }
    void bbox(Column& out) {
        need('[');
        int n = 0;
        if (!eat(']')) {
            do { out.push<double>(as_double()); ++n; } while (eat(','));
            need(']');
        }
        if (n != 4) throw Fail{"ValueError: bbox does not have 4 elements"};
    }
};

[[noreturn]] void key_error(const char* k) {
    throw Fail{std::string("KeyError: ") + k};
}

// generic walk over the objects of a JSON array; f(key_begin, key_end) consumes one value
template <typename F, typename G>
void each_object(Parser& P, F&& per_key, G&& per_object_end) {
    P.need('[');
    if (P.eat(']')) return;
    do {
        P.need('{');
        if (!P.eat('}')) {
            do {
                const char *s, *e;
                P.raw_string(s, e);
                P.need(':');
                per_key(s, e);
            } while (P.eat(','));
            P.need('}');
        }
        per_object_end();
    } while (P.eat(','));
    P.need(']');
}

void finish_ragged(Column& off, int64_t total) { off.push<int64_t>(total); }

void parse_annotation_file(Parser& P, Doc& D) {
    Column &img_id = D.c("img_id", 8), &img_vid = D.c("img_video_id", 8), &img_fi = D.c("img_frame_index", 8);
    Column &img_neg_off = D.c("img_neg__off", 8), &img_neg_val = D.c("img_neg__val", 8);
    Column &img_nel_off = D.c("img_nel__off", 8), &img_nel_val = D.c("img_nel__val", 8);
    Column &vid_id = D.c("vid_id", 8);
    Column &vid_neg_off = D.c("vid_neg__off", 8), &vid_neg_val = D.c("vid_neg__val", 8);
    Column &vid_nel_off = D.c("vid_nel__off", 8), &vid_nel_val = D.c("vid_nel__val", 8);
    Column &trk_id = D.c("trk_id", 8), &trk_cat = D.c("trk_category_id", 8), &trk_vid = D.c("trk_video_id", 8);
    Column &trk_ign = D.c("trk_ignore", 1);
    Column &cat_id = D.c("cat_id", 8), &cat_freq = D.c("cat_freq", 1);
    Column &mm_src = D.c("merge_map__k", 8), &mm_dst = D.c("merge_map__v", 8);
    Column &ann_id = D.c("ann_id", 8), &ann_img = D.c("ann_image_id", 8), &ann_trk = D.c("ann_track_id", 8);
    Column &ann_cat = D.c("ann_category_id", 8), &ann_bbox = D.c("ann_bbox", 8), &ann_area = D.c("ann_area", 8);
    Column &ann_vis = D.c("ann_visibility", 8), &ann_oof = D.c("ann_oof", 1), &ann_ign = D.c("ann_ignore", 1);
    Column &flags = D.c("flags", 8);      // [has_image_lists, has_video_lists, seen sections bitmask]
    int64_t img_missing_neg = 0, vid_missing_neg = 0, img_missing_nel = 0, vid_missing_nel = 0;
    int64_t seen = 0;

    P.need('{');
    if (!P.eat('}')) {
        do {
            const char *ks, *ke;
            P.raw_string(ks, ke);
            P.need(':');
            if (P.key_is(ks, ke, "images")) {
                seen |= 1;
                int64_t id = 0, vid = -1, fi = 0, n_neg = 0, n_nel = 0;
                bool has_id = false, has_neg = false, has_nel = false;
                each_object(P, [&](const char* s, const char* e) {
                    if (P.key_is(s, e, "id")) { id = P.as_int(); has_id = true; }
                    else if (P.key_is(s, e, "video_id")) vid = P.as_int();
                    else if (P.key_is(s, e, "frame_index")) fi = P.as_int();
                    else if (P.key_is(s, e, "neg_category_ids")) { n_neg = P.int_list(img_neg_val); has_neg = true; }
                    else if (P.key_is(s, e, "not_exhaustive_category_ids")) { n_nel = P.int_list(img_nel_val); has_nel = true; }
                    else P.skip_value();
                }, [&]() {
                    if (!has_id) key_error("id");
                    img_id.push(id); img_vid.push(vid); img_fi.push(fi);
                    img_neg_off.push<int64_t>((int64_t)img_neg_val.data.size() / 8 - n_neg);
                    img_nel_off.push<int64_t>((int64_t)img_nel_val.data.size() / 8 - n_nel);
                    img_missing_neg += !has_neg; img_missing_nel += !has_nel;
                    vid = -1; fi = 0; n_neg = n_nel = 0; has_id = has_neg = has_nel = false;
                });
            } else if (P.key_is(ks, ke, "videos")) {
                seen |= 2;
                int64_t id = 0, n_neg = 0, n_nel = 0;
                bool has_id = false, has_neg = false, has_nel = false;
                each_object(P, [&](const char* s, const char* e) {
                    if (P.key_is(s, e, "id")) { id = P.as_int(); has_id = true; }
                    else if (P.key_is(s, e, "neg_category_ids")) { n_neg = P.int_list(vid_neg_val); has_neg = true; }
                    else if (P.key_is(s, e, "not_exhaustive_category_ids")) { n_nel = P.int_list(vid_nel_val); has_nel = true; }
                    else P.skip_value();
                }, [&]() {
                    if (!has_id) key_error("id");
                    vid_id.push(id);
                    vid_neg_off.push<int64_t>((int64_t)vid_neg_val.data.size() / 8 - n_neg);
                    vid_nel_off.push<int64_t>((int64_t)vid_nel_val.data.size() / 8 - n_nel);
                    vid_missing_neg += !has_neg; vid_missing_nel += !has_nel;
                    n_neg = n_nel = 0; has_id = has_neg = has_nel = false;
                });
            } else if (P.key_is(ks, ke, "tracks")) {
                seen |= 4;
                int64_t id = 0, cat = 0, vid = 0;
                int have = 0;
                uint8_t ign = 0;
                each_object(P, [&](const char* s, const char* e) {
                    if (P.key_is(s, e, "id")) { id = P.as_int(); have |= 1; }
                    else if (P.key_is(s, e, "category_id")) { cat = P.as_int(); have |= 2; }
                    else if (P.key_is(s, e, "video_id")) { vid = P.as_int(); have |= 4; }
                    else if (P.key_is(s, e, "ignore")) ign = P.truthy() ? 1 : 0;
                    else P.skip_value();
                }, [&]() {
                    if (!(have & 1)) key_error("id");
                    if (!(have & 2)) key_error("category_id");
                    if (!(have & 4)) key_error("video_id");
                    trk_id.push(id); trk_cat.push(cat); trk_vid.push(vid); trk_ign.push(ign);
                    have = 0; ign = 0;
                });
            } else if (P.key_is(ks, ke, "categories")) {
                seen |= 8;
                int64_t id = 0;
                bool has_id = false;
                uint8_t freq = 255;
                std::vector<int64_t> merged;
                each_object(P, [&](const char* s, const char* e) {
                    if (P.key_is(s, e, "id")) { id = P.as_int(); has_id = true; }
                    else if (P.key_is(s, e, "frequency")) {
                        P.ws();
                        if (P.p < P.end && *P.p == '"') {
                            const char *fs, *fe;
                            P.raw_string(fs, fe);
                            freq = (fe - fs == 1) ? (*fs == 'r' ? 0 : *fs == 'c' ? 1 : *fs == 'f' ? 2 : 255) : 255;
                        } else P.skip_value();
                    } else if (P.key_is(s, e, "merged")) {
                        each_object(P, [&](const char* ms, const char* me) {
                            if (P.key_is(ms, me, "id")) merged.push_back(P.as_int());
                            else P.skip_value();
                        }, []() {});
                    } else P.skip_value();
                }, [&]() {
                    if (!has_id) key_error("id");
                    cat_id.push(id); cat_freq.push(freq);
                    for (int64_t m : merged) { mm_src.push(m); mm_dst.push(id); }
                    merged.clear(); has_id = false; freq = 255;
                });
            } else if (P.key_is(ks, ke, "annotations")) {
                seen |= 16;
                int64_t id = 0, img = 0, trk = -1, cat = 0;
                double area = 0.0, vis = NAN;
                uint8_t oof = 2, ign = 0;
                int have = 0;
                each_object(P, [&](const char* s, const char* e) {
                    if (P.key_is(s, e, "id")) { id = P.as_int(); have |= 1; }
                    else if (P.key_is(s, e, "image_id")) { img = P.as_int(); have |= 2; }
                    else if (P.key_is(s, e, "category_id")) { cat = P.as_int(); have |= 4; }
                    else if (P.key_is(s, e, "bbox")) { P.bbox(ann_bbox); have |= 8; }
                    else if (P.key_is(s, e, "area")) { area = P.as_double(); have |= 16; }
                    else if (P.key_is(s, e, "track_id")) trk = P.as_int();
                    else if (P.key_is(s, e, "visibility")) vis = P.as_double();
                    else if (P.key_is(s, e, "out_of_frame")) oof = P.truthy() ? 1 : 0;
                    else if (P.key_is(s, e, "ignore")) ign = P.truthy() ? 1 : 0;
                    else P.skip_value();
                }, [&]() {
                    if (!(have & 1)) key_error("id");
                    if (!(have & 2)) key_error("image_id");
                    if (!(have & 4)) key_error("category_id");
                    if (!(have & 8)) key_error("bbox");
                    if (!(have & 16)) key_error("area");
                    ann_id.push(id); ann_img.push(img); ann_trk.push(trk); ann_cat.push(cat);
                    ann_area.push(area); ann_vis.push(vis); ann_oof.push(oof); ann_ign.push(ign);
                    trk = -1; vis = NAN; oof = 2; ign = 0; have = 0;
                });
            } else {
                if (P.key_is(ks, ke, "info")) seen |= 32;
                P.skip_value();
            }
        } while (P.eat(','));
        P.need('}');
    }
    finish_ragged(img_neg_off, (int64_t)img_neg_val.data.size() / 8);
    finish_ragged(img_nel_off, (int64_t)img_nel_val.data.size() / 8);
    finish_ragged(vid_neg_off, (int64_t)vid_neg_val.data.size() / 8);
    finish_ragged(vid_nel_off, (int64_t)vid_nel_val.data.size() / 8);
    if (!(seen & 1)) key_error("images");
    if (!(seen & 8)) key_error("categories");
    if (!(seen & 16)) key_error("annotations");
    // columnar.GtColumns.from_dict: lists are used only when EVERY image / video carries them;
    // then a missing not_exhaustive list is a KeyError
    const bool has_img = img_missing_neg == 0, has_vid = vid_missing_neg == 0;
    if (has_img && img_missing_nel) key_error("not_exhaustive_category_ids");
    if (has_vid && vid_missing_nel) key_error("not_exhaustive_category_ids");
    flags.push<int64_t>(has_img);
    flags.push<int64_t>(has_vid);
    flags.push<int64_t>(seen);
}

// One result object at P (just past its '{'): fields into the six columns.
struct ResultCols {
    Column img, trk, cat, vid, bbox, score;
    int64_t missing_trk = 0, missing_vid = 0;     // objects without "track_id" / "video_id"
};

inline void parse_result_object(Parser& P, ResultCols& R) {
    int64_t i = 0, t = -1, c = 0, v = -1;
    double s = 0.0;
    int have = 0;
    if (!P.eat('}')) {
        do {
            const char *ks, *ke;
            P.raw_string(ks, ke);
            P.need(':');
            if (P.key_is(ks, ke, "image_id")) { i = P.as_int(); have |= 1; }
            else if (P.key_is(ks, ke, "category_id")) { c = P.as_int(); have |= 2; }
            else if (P.key_is(ks, ke, "bbox")) { P.bbox(R.bbox); have |= 4; }
            else if (P.key_is(ks, ke, "score")) { s = P.as_double(); have |= 8; }
            else if (P.key_is(ks, ke, "track_id")) { t = P.as_int(); have |= 16; }
            else if (P.key_is(ks, ke, "video_id")) { v = P.as_int(); have |= 32; }
            else P.skip_value();
        } while (P.eat(','));
        P.need('}');
    }
    // the frame evaluator needs neither key (lvis_amodal/results.py); the track path raises
    // KeyError for them (tools/eval_on_tao_amodal.py:57, tao_amodal/results.py:71) — counted here
    if (!(have & 16)) ++R.missing_trk;
    if (!(have & 32)) ++R.missing_vid;
    if (!(have & 1)) key_error("image_id");
    if (!(have & 2)) key_error("category_id");
    if (!(have & 4)) key_error("bbox");
    if (!(have & 8)) key_error("score");
    R.img.push(i); R.trk.push(t); R.cat.push(c); R.vid.push(v); R.score.push(s);
}

void store_result_cols(Doc& D, std::vector<ResultCols>& parts) {
    const char* names[6] = {"image_id", "track_id", "category_id", "video_id", "bbox", "score"};
    {
        int64_t mt = 0, mv = 0;
        for (auto& r : parts) { mt += r.missing_trk; mv += r.missing_vid; }
        Column& m = D.c("dt_missing", 8);
        m.push(mt);
        m.push(mv);
    }
    for (int k = 0; k < 6; ++k) {
        Column& dst = D.c(names[k], 8);
        size_t total = 0;
        auto pick = [&](ResultCols& r) -> Column& {
            return k == 0 ? r.img : k == 1 ? r.trk : k == 2 ? r.cat : k == 3 ? r.vid : k == 4 ? r.bbox : r.score;
        };
        for (auto& r : parts) total += pick(r).data.size();
        dst.data.resize(total);
        size_t o = 0;
        for (auto& r : parts) {
            Column& c = pick(r);
            if (!c.data.empty()) memcpy(dst.data.data() + o, c.data.data(), c.data.size());
            o += c.data.size();
        }
    }
}

// Sequential parse: the authority for results and for error messages.
void parse_result_file(Parser& P, Doc& D) {
    std::vector<ResultCols> parts(1);
    P.ws();
    if (P.p >= P.end || *P.p != '[') throw Fail{"AssertionError: results is not a list."};
    P.need('[');
    if (!P.eat(']')) {
        do {
            P.need('{');
            parse_result_object(P, parts[0]);
        } while (P.eat(','));
        P.need(']');
    }
    store_result_cols(D, parts);
}

// Speculative parallel parse of a large result list.  The list is cut at guessed object starts
// (the first "{" after a "}" + "," near each cut); worker k parses whole objects from its start
// until it reaches worker k + 1's start.  The guesses are then VERIFIED by chaining: worker 0
// starts at the true beginning, so its object boundaries are true; if it ends exactly on
// worker 1's start, that start was a true boundary too, and so on.  Any mismatch, or any error
// in any worker, discards everything and the caller parses sequentially (which also produces
// the reference-compatible error message).  Returns false when it did not succeed.
bool parse_result_file_parallel(const char* begin, const char* end, Doc& D, int n_workers) {
    Parser H{begin, end, begin};
    H.ws();
    if (H.p >= end || *H.p != '[') return false;
    ++H.p;
    H.ws();
    if (H.p >= end || *H.p != '{') return false;
    const char* first = H.p;
    std::vector<const char*> start((size_t)n_workers + 1, nullptr);
    start[0] = first;
    const size_t span = (size_t)(end - first);
    for (int k = 1; k < n_workers; ++k) {
        const char* q = first + span / n_workers * k;
        // guess: '}' [ws] ',' [ws] '{'
        const char* found = nullptr;
        for (; q < end; ++q) {
            if (*q != '}') continue;
            const char* r = q + 1;
            while (r < end && (*r == ' ' || *r == '\n' || *r == '\t' || *r == '\r')) ++r;
            if (r >= end || *r != ',') continue;
            ++r;
            while (r < end && (*r == ' ' || *r == '\n' || *r == '\t' || *r == '\r')) ++r;
            if (r < end && *r == '{') { found = r; break; }
        }
        if (!found || found <= start[k - 1]) return false;
        start[k] = found;
    }
    start[n_workers] = end;          // the last worker runs to the closing bracket
    std::vector<ResultCols> parts((size_t)n_workers);
    std::vector<const char*> stop((size_t)n_workers, nullptr);
    std::vector<int> ok((size_t)n_workers, 0);
    std::vector<std::thread> th;
    for (int k = 0; k < n_workers; ++k)
        th.emplace_back([&, k]() {
            try {
                Parser P{start[k], end, begin};
                const bool last = k == n_workers - 1;
                while (true) {
                    P.need('{');
                    parse_result_object(P, parts[k]);
                    if (!P.eat(',')) {               // end of the list: only the last worker may see it
                        P.need(']');
                        P.ws();
                        if (!last || P.p != end) return;
                        stop[k] = end;
                        ok[k] = 1;
                        return;
                    }
                    P.ws();
                    if (!last && P.p >= start[k + 1]) { stop[k] = P.p; ok[k] = 1; return; }
                }
            } catch (...) {
            }
        });
    for (auto& t : th) t.join();
    for (int k = 0; k < n_workers; ++k) {
        if (!ok[k]) return false;
        if (k + 1 < n_workers && stop[k] != start[k + 1]) return false;
    }
    store_result_cols(D, parts);
    return true;
}

}  // namespace

struct ta_json_doc {
    Doc doc;
};

extern "C" const char* ta_json_error(void) { return g_err; }
extern "C" int ta_json_last_parse_workers(void) { return g_last_workers; }

extern "C" int ta_json_open(const char* path, int kind, ta_json_doc** out) {
    if (!path || !out) { snprintf(g_err, sizeof(g_err), "ta_json_open: NULL argument"); return -1; }
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { snprintf(g_err, sizeof(g_err), "FileNotFoundError: %s", path); return -2; }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); snprintf(g_err, sizeof(g_err), "OSError: fstat failed"); return -2; }
    const size_t len = (size_t)st.st_size;
    void* map = len ? mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    close(fd);
    if (len && map == MAP_FAILED) { snprintf(g_err, sizeof(g_err), "OSError: mmap failed"); return -2; }
    if (len) madvise(map, len, MADV_SEQUENTIAL);
    ta_json_doc* d = new ta_json_doc();
    int rc = 0;
    try {
        Parser P{static_cast<const char*>(map), static_cast<const char*>(map) + len,
                 static_cast<const char*>(map)};
        bool done = false;
        g_last_workers = 1;
        const char* min_env = getenv("TA_INGEST_PAR_MIN_BYTES");        // testing hook
        const size_t par_min = min_env ? (size_t)atoll(min_env) : ((size_t)8 << 20);
        if (kind == TA_JSON_RESULTS && len >= par_min) {
            // large result lists: speculative parallel parse, verified; else fall through
            unsigned hw = std::thread::hardware_concurrency();
            const char* env = getenv("TA_INGEST_THREADS");
            int nw = env ? atoi(env) : (int)(hw ? (hw > 16 ? 16 : hw) : 1);
            if (nw > 1) {
                done = parse_result_file_parallel(P.begin, P.end, d->doc, nw);
                if (!done) d->doc.col.clear();
                else g_last_workers = nw;
            }
        }
        if (!done) {
            if (kind == TA_JSON_ANNOTATIONS) parse_annotation_file(P, d->doc);
            else parse_result_file(P, d->doc);
            P.ws();
            if (P.p != P.end) P.fail("extra data");
        }
    } catch (const Fail& f) {
        snprintf(g_err, sizeof(g_err), "%s", f.msg.c_str());
        rc = -3;
    } catch (const std::exception& e) {
        snprintf(g_err, sizeof(g_err), "RuntimeError: %s", e.what());
        rc = -3;
    }
    if (len) munmap(map, len);
    if (rc) { delete d; return rc; }
    *out = d;
    return 0;
}

extern "C" int64_t ta_json_count(const ta_json_doc* d, const char* column) {
    if (!d || !column) return -1;
    auto it = d->doc.col.find(column);
    if (it == d->doc.col.end()) return -1;
    return (int64_t)(it->second.data.size() / it->second.elem);
}

extern "C" int ta_json_copy(const ta_json_doc* d, const char* column, void* dst, int64_t bytes) {
    if (!d || !column) return -1;
    auto it = d->doc.col.find(column);
    if (it == d->doc.col.end()) { snprintf(g_err, sizeof(g_err), "unknown column %s", column); return -1; }
    if ((int64_t)it->second.data.size() != bytes) { snprintf(g_err, sizeof(g_err), "size mismatch for %s", column); return -1; }
    if (bytes) memcpy(dst, it->second.data.data(), (size_t)bytes);
    return 0;
}

extern "C" void ta_json_close(ta_json_doc* d) { delete d; }
