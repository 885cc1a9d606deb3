"""Native JSON ingest: annotation / prediction files -> GtColumns / DtColumns through
libta_ingest.so (include/ta_ingest.h, csrc/ta_json.cpp).  Same columns, dtypes and exceptions
as ``GtColumns.from_dict(json.load(f))`` / ``DtColumns.from_list(json.load(f))``
(tests/test_ingest.py compares them field by field), without materialising Python dicts."""
from __future__ import annotations

import ctypes as C
import json
import os
import threading
from typing import Dict, Tuple

import numpy as np

from .columnar import DtColumns, GtColumns

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libta_ingest.so")
_lib = None
_CACHE: Dict[Tuple[str, float, int, int], object] = {}

_EXC = {"KeyError": KeyError, "ValueError": ValueError, "AssertionError": AssertionError,
        "FileNotFoundError": FileNotFoundError, "OSError": OSError, "RuntimeError": RuntimeError}


def load_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; "
                               "g.build()'`" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.ta_json_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        lib.ta_json_error.restype = C.c_char_p
        lib.ta_json_count.argtypes = [C.c_void_p, C.c_char_p]
        lib.ta_json_count.restype = C.c_int64
        lib.ta_json_copy.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
        lib.ta_json_close.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def _raise(msg: str):
    kind, _, rest = msg.partition(": ")
    if kind == "JSONDecodeError":
        raise json.JSONDecodeError(rest, "", 0)
    if kind == "KeyError":
        raise KeyError(rest)
    raise _EXC.get(kind, RuntimeError)(rest if kind in _EXC else msg)


class _Doc:
    def __init__(self, path: str, kind: int):
        self.lib = load_lib()
        self.h = C.c_void_p()
        if self.lib.ta_json_open(os.fsencode(path), kind, C.byref(self.h)) != 0:
            _raise(self.lib.ta_json_error().decode("utf-8", "replace"))

    def col(self, name: str, dtype) -> np.ndarray:
        n = self.lib.ta_json_count(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=dtype)
        if self.lib.ta_json_copy(self.h, name.encode(), out.ctypes.data_as(C.c_void_p),
                                 out.nbytes) != 0:
            raise RuntimeError(self.lib.ta_json_error().decode())
        return out

    def close(self):
        if self.h:
            self.lib.ta_json_close(self.h)
            self.h = None


_LOCK = threading.Lock()
_INFLIGHT: Dict[Tuple[str, float, int, int], threading.Event] = {}


def _cached(path: str, kind: int, build):
    """One parse per (path, mtime, size, kind), also when a prefetch thread is already on it: a
    second caller waits for the first instead of parsing again.  Failures are not cached — the
    waiting caller then parses itself and raises the error in its own thread."""
    st = os.stat(path)
    key = (os.path.abspath(path), st.st_mtime, st.st_size, kind)
    while True:
        with _LOCK:
            if key in _CACHE:
                return _CACHE[key]
            ev = _INFLIGHT.get(key)
            mine = ev is None
            if mine:
                ev = _INFLIGHT[key] = threading.Event()
        if not mine:
            ev.wait()
            continue
        try:
            val = build()
            with _LOCK:
                _CACHE[key] = val
            return val
        finally:
            with _LOCK:
                _INFLIGHT.pop(key, None)
            ev.set()


def prefetch(path: str, kind: str) -> threading.Thread:
    """Start parsing a file ("gt" annotations / "dt" results) on a background thread (the native
    parser runs without the GIL), so that the two files of the CLI are read concurrently.
    Errors are left to the regular load_gt / load_dt call that follows."""
    def work():
        try:
            (load_gt if kind == "gt" else load_dt)(path)
        except Exception:           # noqa: BLE001 - reported by the foreground load
            pass
    t = threading.Thread(target=work, name="ta-ingest-prefetch", daemon=True)
    t.start()
    return t


def load_gt(path: str, need_videos_tracks: bool = False) -> GtColumns:
    """Annotation file -> GtColumns (cached per (path, mtime, size))."""
    def build():
        d = _Doc(path, 0)
        try:
            i64, f64, u8 = np.int64, np.float64, np.uint8
            flags = d.col("flags", i64)
            cols = GtColumns(
                img_id=d.col("img_id", i64), img_video_id=d.col("img_video_id", i64),
                img_frame_index=d.col("img_frame_index", i64),
                img_neg=(d.col("img_neg__off", i64), d.col("img_neg__val", i64)),
                img_nel=(d.col("img_nel__off", i64), d.col("img_nel__val", i64)),
                vid_id=d.col("vid_id", i64),
                vid_neg=(d.col("vid_neg__off", i64), d.col("vid_neg__val", i64)),
                vid_nel=(d.col("vid_nel__off", i64), d.col("vid_nel__val", i64)),
                trk_id=d.col("trk_id", i64), trk_category_id=d.col("trk_category_id", i64),
                trk_video_id=d.col("trk_video_id", i64), trk_ignore=d.col("trk_ignore", u8),
                cat_id=d.col("cat_id", i64), cat_freq=d.col("cat_freq", u8),
                merge_map=dict(zip(d.col("merge_map__k", i64).tolist(),
                                   d.col("merge_map__v", i64).tolist())),
                ann_id=d.col("ann_id", i64), ann_image_id=d.col("ann_image_id", i64),
                ann_track_id=d.col("ann_track_id", i64),
                ann_category_id=d.col("ann_category_id", i64),
                ann_bbox=d.col("ann_bbox", f64).reshape(-1, 4), ann_area=d.col("ann_area", f64),
                ann_visibility=d.col("ann_visibility", f64), ann_oof=d.col("ann_oof", u8),
                ann_ignore=d.col("ann_ignore", u8),
                has_image_lists=bool(flags[0]), has_video_lists=bool(flags[1]))
            if not cols.has_image_lists:
                n = cols.img_id.size
                cols.img_neg = cols.img_nel = (np.zeros(n + 1, dtype=i64), np.zeros(0, dtype=i64))
            if not cols.has_video_lists:
                n = cols.vid_id.size
                cols.vid_neg = cols.vid_nel = (np.zeros(n + 1, dtype=i64), np.zeros(0, dtype=i64))
            cols._sections = int(flags[2])
            return cols
        finally:
            d.close()
    cols = _cached(path, 0, build)
    if need_videos_tracks:
        for bit, name in ((2, "videos"), (4, "tracks")):
            if not cols._sections & bit:
                raise KeyError(name)                       # tao.py:113-114
    return cols


def load_dt(path: str) -> DtColumns:
    """Prediction file -> DtColumns (cached per (path, mtime, size)); callers that modify the
    columns (make_track_ids_unique) must copy first."""
    def build():
        d = _Doc(path, 1)
        try:
            i64, f64 = np.int64, np.float64
            miss = d.col("dt_missing", i64)
            return DtColumns(image_id=d.col("image_id", i64), track_id=d.col("track_id", i64),
                             category_id=d.col("category_id", i64), video_id=d.col("video_id", i64),
                             bbox=d.col("bbox", f64).reshape(-1, 4), score=d.col("score", f64),
                             missing_track_id=int(miss[0]) if miss.size else 0,
                             missing_video_id=int(miss[1]) if miss.size > 1 else 0)
        finally:
            d.close()
    return _cached(path, 1, build)
