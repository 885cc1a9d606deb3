"""The C-ABI library loads on a machine without a GPU and exports every symbol that
include/ta_eval.h declares; argument validation that needs no device work."""
import ctypes as C
import os
import re

from tao_amodal_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ta_eval.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ta_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert set(_lib.EXPORTS) <= set(names)
    for n in names:
        assert getattr(lib, n) is not None, n


def test_abi_version_and_null_ctx_errors():
    lib = _lib.load()
    assert lib.ta_abi_version() == 3
    rc = lib.ta_track_iou(None, None, 0, 0, None, None, None, None, None, None, None, None, 0,
                          None, None)
    assert rc == _lib.TA_ERR_INVALID
    assert b"ctx" in lib.ta_last_error()
    rc = lib.ta_ctx_create(0, None)
    assert rc == _lib.TA_ERR_INVALID


def test_struct_layouts_match_header():
    from tao_amodal_b200 import prep
    assert C.sizeof(_lib.RangeCfg) == prep.RANGE_CFG_DTYPE.itemsize == 72
    assert C.sizeof(_lib.PlanHost) == 6 * 8 + 8 * 4 + 23 * 8


def test_mask_codec_exports_every_declared_symbol():
    """include/ta_mask.h (host-side run-length codec, part of libta_ingest.so)."""
    from tao_amodal_b200 import ingest
    src = open(os.path.join(ROOT, "include", "ta_mask.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(ta_[a-z0-9_]+)\s*\(", src)))
    assert len(names) >= 10
    lib = ingest.load_lib()
    for n in names:
        assert getattr(lib, n) is not None, n
