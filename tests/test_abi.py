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
    assert lib.ta_abi_version() == 4
    rc = lib.ta_track_iou(None, None, 0, 0, None, None, None, None, None, None, None, None, 0,
                          None, None)
    assert rc == _lib.TA_ERR_INVALID
    assert b"ctx" in lib.ta_last_error()
    rc = lib.ta_ctx_create(0, None)
    assert rc == _lib.TA_ERR_INVALID


def test_struct_layouts_match_header():
    from tao_amodal_b200 import prep
    assert C.sizeof(_lib.RangeCfg) == prep.RANGE_CFG_DTYPE.itemsize == 72
    assert C.sizeof(_lib.PlanHost) == 6 * 8 + 8 * 4 + 24 * 8 + 2 * 4
    assert C.sizeof(_lib.HostOut) == 5 * 8


def test_struct_layouts_match_the_c_compiler(tmp_path):
    """sizeof / offsetof of the public structs as gcc sees include/ta_eval.h (which must stay
    plain C) against the ctypes mirrors in _lib.py."""
    import subprocess
    fields = {"ta_plan_host": ("PlanHost", ["n_groups", "n_big", "n_cat", "flags", "grp_dt_off", "grp_cat",
                                            "dt_box", "gt_hp", "cfgs", "dt_box_idx", "dt_box_pool"]),
              "ta_host_out": ("HostOut", ["precision", "recall", "tp_cnt", "fp_cnt", "num_gt"]),
              "ta_range_cfg": ("RangeCfg", [n for n, _ in _lib.RangeCfg._fields_])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ta_eval.h"', 'int main(void) {']
    for st, (_, names) in fields.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (st, st))
        for n in names:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (st, n, st, n))
    lines.append('return 0; }')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True,
                                                 text=True).stdout.splitlines())
    for st, (cls, names) in fields.items():
        T = getattr(_lib, cls)
        assert int(got[st]) == C.sizeof(T), st
        for n in names:
            assert int(got["%s.%s" % (st, n)]) == getattr(T, n).offset, (st, n)


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
