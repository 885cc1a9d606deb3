"""TEST INFRASTRUCTURE — ctypes binding of ``oracle/_ref/libmaskapi_ref.so``.

That library is the reference tree's own C source of the box / mask arithmetic the frame
evaluator gets from ``pycocotools`` (visualization/tao/third_party/pysot/training_dataset/
coco/pycocotools/common/maskApi.c), compiled where it lies by ``oracle/Makefile``.  The
functions below restate the thin Cython wrapper next to it (``_mask.pyx``; cited per
function) so that ``pycocotools.mask.iou / frPyObjects / merge / area / toBbox`` can be
served by the REAL C code when the unmodified reference is run in the build container
(``oracle/ref_shims.py``) and when the numpy restatements of ``oracle/`` are pinned
(``tests/test_oracle_ref_c.py``).

Not part of the product path: only ``tests/``, ``oracle/`` tooling, ``__graft_entry__``'s
build step and ``bench.py``'s CPU legs may touch it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libmaskapi_ref.so")


class RLE(C.Structure):
    # maskApi.h:10  typedef struct { siz h, w, m; uint *cnts; } RLE;   (siz = unsigned long)
    _fields_ = [("h", C.c_ulong), ("w", C.c_ulong), ("m", C.c_ulong),
                ("cnts", C.POINTER(C.c_uint))]


_lib = None


def build(quiet: bool = True) -> bool:
    """Run oracle/Makefile (no-op without the reference tree). True if the library exists."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=False,
                   stdout=subprocess.DEVNULL if quiet else None)
    return os.path.exists(LIB_PATH)


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libmaskapi_ref.so missing: run `make -C oracle` "
                               "in a container that has /root/reference")
        L = C.CDLL(LIB_PATH)
        P, UL = C.c_void_p, C.c_ulong
        L.bbIou.argtypes = [P, P, UL, UL, P, P]
        L.rleIou.argtypes = [P, P, UL, UL, P, P]
        L.rleInit.argtypes = [P, UL, UL, UL, P]
        L.rleFree.argtypes = [P]
        L.rleMerge.argtypes = [P, P, UL, C.c_int]
        L.rleArea.argtypes = [P, UL, P]
        L.rleToBbox.argtypes = [P, P, UL]
        L.rleFrBbox.argtypes = [P, P, UL, UL, UL]
        L.rleFrPoly.argtypes = [P, P, UL, UL, UL]
        L.rleFrString.argtypes = [P, C.c_char_p, UL, UL]
        L.rleToString.argtypes = [P]
        L.rleToString.restype = C.c_void_p
        L.rleEncode.argtypes = [P, P, UL, UL, UL]
        L.rleDecode.argtypes = [P, P, UL]
        _lib = L
    return _lib


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


class RLEs:
    """_mask.pyx:56-75 — an owned array of C RLE structs."""

    def __init__(self, n):
        self.n = n
        self.arr = (RLE * max(n, 1))()

    def ptr(self, i=0):
        return C.byref(self.arr[i]) if i else C.cast(self.arr, C.c_void_p)

    def counts(self, i):
        r = self.arr[i]
        return np.ctypeslib.as_array(r.cnts, shape=(r.m,)).copy() if r.m else np.zeros(0, np.uint32)

    def __del__(self):
        try:
            L = lib()
            for i in range(self.n):
                if self.arr[i].cnts:
                    L.rleFree(C.byref(self.arr[i]))
        except Exception:   # interpreter shutdown
            pass


def _to_string(rs: RLEs):
    """_mask.pyx:103-116."""
    L = lib()
    out = []
    for i in range(rs.n):
        p = L.rleToString(C.byref(rs.arr[i]))
        s = C.string_at(p)
        _libc.free(p)
        out.append({"size": [int(rs.arr[i].h), int(rs.arr[i].w)], "counts": s})
    return out


def _fr_string(objs) -> RLEs:
    """_mask.pyx:119-134."""
    L = lib()
    rs = RLEs(len(objs))
    for i, o in enumerate(objs):
        s = o["counts"].encode() if isinstance(o["counts"], str) else bytes(o["counts"])
        L.rleFrString(C.byref(rs.arr[i]), s, o["size"][0], o["size"][1])
    return rs


def rle_counts(obj) -> np.ndarray:
    """Uncompressed run lengths (uint32) of one compressed RLE dict."""
    return _fr_string([obj]).counts(0)


def merge(objs, intersect=0):
    """_mask.pyx:152-157."""
    rs = _fr_string(objs)
    r = RLEs(1)
    lib().rleMerge(rs.ptr(), r.ptr(), rs.n, int(intersect))
    return _to_string(r)[0]


def area(objs):
    """_mask.pyx:159-168 behind mask.py:93-97 (a single dict gives a numpy uint32 scalar)."""
    single = not isinstance(objs, list)
    rs = _fr_string([objs] if single else objs)
    a = np.zeros(rs.n, dtype=np.uint32)
    lib().rleArea(rs.ptr(), rs.n, a.ctypes.data_as(C.c_void_p))
    return a[0] if single else a


def toBbox(objs):
    """_mask.pyx:241-251 behind mask.py:99-103."""
    single = not isinstance(objs, list)
    rs = _fr_string([objs] if single else objs)
    bb = np.zeros((rs.n, 4), dtype=np.double)
    lib().rleToBbox(rs.ptr(), bb.ctypes.data_as(C.c_void_p), rs.n)
    return bb[0] if single else bb


def frBbox(bb, h, w):
    """_mask.pyx:253-258."""
    bb = np.ascontiguousarray(bb, dtype=np.double).reshape(-1, 4)
    rs = RLEs(bb.shape[0])
    lib().rleFrBbox(rs.ptr(), bb.ctypes.data_as(C.c_void_p), h, w, bb.shape[0])
    return _to_string(rs)


def frPoly(poly, h, w):
    """_mask.pyx:260-268."""
    rs = RLEs(len(poly))
    for i, p in enumerate(poly):
        xy = np.array(p, dtype=np.double)
        lib().rleFrPoly(C.byref(rs.arr[i]), xy.ctypes.data_as(C.c_void_p), int(len(p) / 2), h, w)
    return _to_string(rs)


def frUncompressedRLE(uc, h, w):
    """_mask.pyx:270-286."""
    out = []
    for o in uc:
        cnts = np.array(o["counts"], dtype=np.uint32)
        rs = RLEs(1)
        lib().rleInit(rs.ptr(), o["size"][0], o["size"][1], len(cnts),
                      cnts.ctypes.data_as(C.c_void_p))
        out.append(_to_string(rs)[0])
    return out


def frPyObjects(pyobj, h, w):
    """_mask.pyx:288-308 (same dispatch order, including the 4-number "polygon is a box" case)."""
    if type(pyobj) == np.ndarray:
        return frBbox(pyobj, h, w)
    if type(pyobj) == list and len(pyobj[0]) == 4:
        return frBbox(pyobj, h, w)
    if type(pyobj) == list and len(pyobj[0]) > 4:
        return frPoly(pyobj, h, w)
    if type(pyobj) == list and type(pyobj[0]) == dict and "counts" in pyobj[0] and "size" in pyobj[0]:
        return frUncompressedRLE(pyobj, h, w)
    if type(pyobj) == list and len(pyobj) == 4:
        return frBbox([pyobj], h, w)[0]
    if type(pyobj) == list and len(pyobj) > 4:
        return frPoly([pyobj], h, w)[0]
    if type(pyobj) == dict and "counts" in pyobj and "size" in pyobj:
        return frUncompressedRLE([pyobj], h, w)[0]
    raise Exception("input type is not supported.")


def encode(mask):
    """_mask.pyx:137-143 — mask uint8 [h, w, n] Fortran order."""
    mask = np.asfortranarray(mask, dtype=np.uint8)
    h, w, n = mask.shape
    rs = RLEs(n)
    lib().rleEncode(rs.ptr(), mask.ctypes.data_as(C.c_void_p), h, w, n)
    return _to_string(rs)


def decode(objs):
    """_mask.pyx:145-150 behind mask.py:87-91."""
    single = not isinstance(objs, list)
    rs = _fr_string([objs] if single else objs)
    h, w = int(rs.arr[0].h), int(rs.arr[0].w)
    m = np.zeros((h, w, rs.n), dtype=np.uint8, order="F")
    lib().rleDecode(rs.ptr(), m.ctypes.data_as(C.c_void_p), rs.n)
    return m[:, :, 0] if single else m


def iou(dt, gt, pyiscrowd):
    """_mask.pyx:171-239: boxes (list of 4-lists / Nx4 array) -> bbIou, RLE dicts -> rleIou;
    [] when either side is empty; result [len(dt), len(gt)] (column-major C output)."""
    def prep(objs):
        if len(objs) == 0:
            return objs
        if isinstance(objs, np.ndarray):
            return np.ascontiguousarray(objs, dtype=np.double).reshape(-1, 4)
        if all(isinstance(o, dict) for o in objs):
            return _fr_string(objs)
        if all(len(o) == 4 and isinstance(o, (list, np.ndarray)) for o in objs):
            return np.ascontiguousarray(np.array(objs, dtype=np.double)).reshape(-1, 4)
        raise Exception("list input can be bounding box (Nx4) or RLEs ([RLE])")

    crowd = np.array(pyiscrowd, dtype=np.uint8)
    dt, gt = prep(dt), prep(gt)
    m = dt.n if isinstance(dt, RLEs) else len(dt)
    n = gt.n if isinstance(gt, RLEs) else len(gt)
    if m == 0 or n == 0:
        return []
    if type(dt) != type(gt):
        raise Exception("The dt and gt should have the same data type, either RLEs, list or np.ndarray")
    out = np.zeros(m * n, dtype=np.double)
    cp = crowd.ctypes.data_as(C.c_void_p) if crowd.size else None
    if isinstance(dt, RLEs):
        lib().rleIou(dt.ptr(), gt.ptr(), m, n, cp, out.ctypes.data_as(C.c_void_p))
    else:
        lib().bbIou(dt.ctypes.data_as(C.c_void_p), gt.ctypes.data_as(C.c_void_p), m, n, cp,
                    out.ctypes.data_as(C.c_void_p))
    return out.reshape((m, n), order="F")
