"""Run-length mask codec of the segm evaluation path: ctypes binding of the ``ta_rle_pool_*``
entry points of libta_ingest.so (include/ta_mask.h, csrc/ta_mask.cpp).

Replaces what the reference takes from pycocotools around ``LVISEval(iou_type="segm")``:
``LVIS.ann_to_rle`` (lvis_amodal/lvis.py:155-178: polygons -> frPyObjects + merge, uncompressed
counts -> frPyObjects, compressed RLE as is) and the box polygons ``LVISResults`` synthesises
for bbox results (results.py:50-52).  Conversion only — the mask IoU runs on the GPU
(``ta_rle_iou``)."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import ingest

_bound = False


def _lib():
    global _bound
    lib = ingest.load_lib()
    if not _bound:
        P, I64 = C.c_void_p, C.c_int64
        lib.ta_rle_pool_create.restype = P
        lib.ta_rle_pool_destroy.argtypes = [P]
        lib.ta_mask_error.restype = C.c_char_p
        lib.ta_rle_pool_add_polygons.argtypes = [P, I64, P, P, I64, I64]
        lib.ta_rle_pool_add_polygons.restype = I64
        lib.ta_rle_pool_add_boxes.argtypes = [P, I64, P, P, P]
        lib.ta_rle_pool_add_boxes.restype = I64
        lib.ta_rle_pool_add_counts.argtypes = [P, I64, P, I64, I64]
        lib.ta_rle_pool_add_counts.restype = I64
        lib.ta_rle_pool_add_string.argtypes = [P, C.c_char_p, I64, I64, I64]
        lib.ta_rle_pool_add_string.restype = I64
        lib.ta_rle_pool_size.argtypes = [P]
        lib.ta_rle_pool_size.restype = I64
        lib.ta_rle_pool_total_counts.argtypes = [P]
        lib.ta_rle_pool_total_counts.restype = I64
        lib.ta_rle_pool_export.argtypes = [P, P, P, P, P, P]
        lib.ta_rle_pool_to_string.argtypes = [P, I64, C.c_char_p, I64]
        lib.ta_rle_pool_to_string.restype = I64
        _bound = True
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class RlePool:
    """Append-only list of masks held by the native library."""

    def __init__(self):
        self._lib = _lib()
        self._h = C.c_void_p(self._lib.ta_rle_pool_create())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ta_rle_pool_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise ValueError(self._lib.ta_mask_error().decode())
        return int(rc)

    def __len__(self):
        return int(self._lib.ta_rle_pool_size(self._h))

    def add_polygons(self, parts: Sequence[Sequence[float]], h: int, w: int) -> int:
        off = np.zeros(len(parts) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(p) for p in parts])
        xy = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1)
                                                  for p in parts]) if len(parts) else np.zeros(0))
        return self._check(self._lib.ta_rle_pool_add_polygons(self._h, len(parts), _ptr(off), _ptr(xy),
                                                              int(h), int(w)))

    def add_boxes(self, boxes, h, w) -> int:
        boxes = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 4)
        n = boxes.shape[0]
        h = np.ascontiguousarray(np.broadcast_to(np.asarray(h, dtype=np.int64), (n,)))
        w = np.ascontiguousarray(np.broadcast_to(np.asarray(w, dtype=np.int64), (n,)))
        return self._check(self._lib.ta_rle_pool_add_boxes(self._h, n, _ptr(boxes), _ptr(h), _ptr(w)))

    def add_counts(self, counts, h: int, w: int) -> int:
        c = np.ascontiguousarray(counts, dtype=np.uint32)
        return self._check(self._lib.ta_rle_pool_add_counts(self._h, c.size, _ptr(c), int(h), int(w)))

    def add_string(self, s, h: int, w: int) -> int:
        b = s.encode() if isinstance(s, str) else bytes(s)
        return self._check(self._lib.ta_rle_pool_add_string(self._h, b, len(b), int(h), int(w)))

    def add_segmentation(self, segm, h: int, w: int) -> int:
        """LVIS.ann_to_rle (lvis.py:165-178) with the dispatch of frPyObjects
        (_mask.pyx:288-308): a list whose first element has 4 numbers is a list of boxes, longer
        first elements make it a list of polygons; the parts are merged (union)."""
        if isinstance(segm, list):
            if len(segm[0]) == 4:
                return self._merge_boxes(segm, h, w)
            return self.add_polygons(segm, h, w)
        if isinstance(segm["counts"], list):
            return self.add_counts(segm["counts"], h, w)       # size taken from the image, lvis.py:175
        return self.add_string(segm["counts"], segm["size"][0], segm["size"][1])

    def _merge_boxes(self, boxes, h, w):
        parts = []
        for x, y, bw, bh in boxes:                             # rleFrBbox's corner order
            parts.append([x, y, x, y + bh, x + bw, y + bh, x + bw, y])
        return self.add_polygons(parts, h, w)

    def export(self):
        """(off int64 [n+1], counts uint32, hw uint32 [n,2], bbox f64 [n,4], area uint32 [n])."""
        n, tot = len(self), int(self._lib.ta_rle_pool_total_counts(self._h))
        off = np.zeros(n + 1, dtype=np.int64)
        cnt = np.zeros(max(tot, 1), dtype=np.uint32)
        hw = np.zeros((max(n, 1), 2), dtype=np.uint32)
        bb = np.zeros((max(n, 1), 4), dtype=np.float64)
        ar = np.zeros(max(n, 1), dtype=np.uint32)
        self._check(self._lib.ta_rle_pool_export(self._h, _ptr(off), _ptr(cnt), _ptr(hw), _ptr(bb), _ptr(ar)))
        return off, cnt[:tot], hw[:n], bb[:n], ar[:n]

    def to_rle(self, i: int) -> dict:
        """Compressed RLE dict of mask i, as pycocotools returns it."""
        cap = 256
        while True:
            buf = C.create_string_buffer(cap)
            r = int(self._lib.ta_rle_pool_to_string(self._h, i, buf, cap))
            if r >= 0:
                break
            if r == -(2 ** 63):
                raise IndexError(self._lib.ta_mask_error().decode())
            cap = -r
        hw = np.zeros((len(self), 2), dtype=np.uint32)
        self._lib.ta_rle_pool_export(self._h, None, None, _ptr(hw), None, None)
        return {"size": [int(hw[i, 0]), int(hw[i, 1])], "counts": buf.value}
