/*
 * ta_ingest.h — C ABI of the host-side JSON ingest (libta_ingest.so, built from
 * tao_amodal_b200/csrc/ta_json.cpp with g++; no CUDA).
 *
 * One call parses one file of the evaluation CLI into flat columns:
 *   TA_JSON_ANNOTATIONS  the annotation file read by Tao.__init__ / LVIS.__init__
 *                        (tao_amodal/evaluation/tao_amodal/tao.py:69-160,
 *                        lvis_amodal/lvis.py:19-61)
 *   TA_JSON_RESULTS      the prediction list read by TaoResults / LVISResults
 *                        (results.py:29-40, lvis_amodal/results.py:29-38)
 * Column names equal the field names of columnar.GtColumns / DtColumns (ragged columns as
 * <name>__off / <name>__val, the merge map as merge_map__k / merge_map__v, plus "flags" =
 * [has_image_lists, has_video_lists, bitmask of top-level sections seen:
 *  1 images, 2 videos, 4 tracks, 8 categories, 16 annotations, 32 info]).
 * Errors: non-zero return, message in ta_json_error() prefixed with the Python exception the
 * dict-based path would raise ("KeyError: image_id", "JSONDecodeError: ...").
 */
#ifndef TA_INGEST_H
#define TA_INGEST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TA_JSON_ANNOTATIONS 0
#define TA_JSON_RESULTS 1

typedef struct ta_json_doc ta_json_doc;

int         ta_json_open(const char* path, int kind, ta_json_doc** out);
const char* ta_json_error(void);
/* Workers that parsed the last document opened on the calling thread: > 1 when a large result
 * list went through the speculative parallel parse (guessed object starts, verified by
 * chaining; any doubt falls back to the sequential parse), else 1. */
int         ta_json_last_parse_workers(void);
int64_t     ta_json_count(const ta_json_doc* doc, const char* column);   /* elements, -1 if unknown */
int         ta_json_copy(const ta_json_doc* doc, const char* column, void* dst, int64_t bytes);
void        ta_json_close(ta_json_doc* doc);

#ifdef __cplusplus
}
#endif
#endif /* TA_INGEST_H */
