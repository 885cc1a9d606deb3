"""TEST INFRASTRUCTURE — not part of the product path.

Runs the UNMODIFIED reference (``/root/reference``, read-only) in this
container so its outputs can pin the oracle and generate ``tests/golden``.
The reference needs four import shims here; none of them
touches the reference's files:

1. ``numpy.float``  (removed in numpy 1.24; used at
   tao_amodal/evaluation/tao_amodal/eval.py:530-531 and
   lvis_amodal/eval.py:373-374)  -> alias of ``float``.
2. ``pycocotools.mask``  (third-party, not vendored under tao_amodal/, not
   installed here).  Served by the reference tree's OWN C source of it —
   visualization/tao/third_party/pysot/training_dataset/coco/pycocotools/
   common/maskApi.c compiled where it lies into oracle/_ref/ (oracle/Makefile,
   bound by oracle/maskapi_ref.py with the wrapper semantics of _mask.pyx) —
   so ``iou`` (bbox: bbIou :109-120, segm: rleIou :75-96), ``frPyObjects``,
   ``merge``, ``area``, ``toBbox`` run the real code.  Without that library
   only ``iou`` on boxes is available, as a numpy restatement of bbIou
   (_mask.pyx:218-239: [] when either side is empty, shape [D,G]).
3. ``matplotlib``  (imported by lvis_amodal/vis.py:6,8 through
   lvis_amodal/__init__.py:5; never called)  -> empty stub modules.
4. ``detectron2``  (tools/eval_on_tao_amodal.py:20-21; only
   ``create_small_table`` is called, :110)  -> the tabulate call detectron2
   uses for that helper.

``/root/reference`` does not exist on the GPU box: nothing imported by the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call into this module.
"""
from __future__ import annotations

import contextlib
import io
import os
import runpy
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("TAO_AMODAL_REF", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "tao_amodal", "evaluation"))


def _bb_iou_stub(dt, gt, iscrowd):
    """maskApi.c:109-120 bbIou with iscrowd all zero; _mask.pyx:218-239 wrapper."""
    if len(dt) == 0 or len(gt) == 0:
        return []
    D = np.asarray(dt, dtype=np.double).reshape(-1, 4)
    G = np.asarray(gt, dtype=np.double).reshape(-1, 4)
    out = np.zeros((D.shape[0], G.shape[0]), dtype=np.double)
    for g in range(G.shape[0]):
        ga = G[g, 2] * G[g, 3]
        for d in range(D.shape[0]):
            da = D[d, 2] * D[d, 3]
            w = min(D[d, 2] + D[d, 0], G[g, 2] + G[g, 0]) - max(D[d, 0], G[g, 0])
            if w <= 0:
                continue
            h = min(D[d, 3] + D[d, 1], G[g, 3] + G[g, 1]) - max(D[d, 1], G[g, 1])
            if h <= 0:
                continue
            i = w * h
            u = da + ga - i
            out[d, g] = i / u
    return out


def _small_table(small_dict):
    # what detectron2.utils.logger.create_small_table does
    from tabulate import tabulate
    keys, values = tuple(zip(*small_dict.items()))
    return tabulate([values], headers=keys, tablefmt="pipe", floatfmt=".3f",
                    stralign="center", numalign="center")


def install_shims() -> None:
    if not hasattr(np, "float"):
        np.float = float  # type: ignore[attr-defined]
    if "pycocotools" not in sys.modules:
        pk = types.ModuleType("pycocotools")
        mk = types.ModuleType("pycocotools.mask")
        from . import maskapi_ref
        if maskapi_ref.available() or maskapi_ref.build():
            for fn in ("iou", "merge", "frPyObjects", "encode", "decode", "area", "toBbox"):
                setattr(mk, fn, getattr(maskapi_ref, fn))
            mk.BACKEND = "reference C (oracle/_ref/libmaskapi_ref.so)"
        else:
            mk.iou = _bb_iou_stub
            mk.BACKEND = "numpy restatement of bbIou"
        pk.mask = mk
        sys.modules["pycocotools"] = pk
        sys.modules["pycocotools.mask"] = mk
    if "matplotlib" not in sys.modules:
        mp = types.ModuleType("matplotlib")
        pp = types.ModuleType("matplotlib.pyplot")
        pa = types.ModuleType("matplotlib.patches")
        pa.Polygon = object
        mp.pyplot, mp.patches = pp, pa
        sys.modules.update({"matplotlib": mp, "matplotlib.pyplot": pp,
                            "matplotlib.patches": pa})
    if "detectron2" not in sys.modules:
        d2 = types.ModuleType("detectron2")
        du = types.ModuleType("detectron2.utils")
        dl = types.ModuleType("detectron2.utils.logger")
        de = types.ModuleType("detectron2.evaluation")
        dl.create_small_table = _small_table
        de.inference_on_dataset = None
        de.print_csv_format = None
        d2.utils, du.logger, d2.evaluation = du, dl, de
        sys.modules.update({"detectron2": d2, "detectron2.utils": du,
                            "detectron2.utils.logger": dl, "detectron2.evaluation": de})


def load_reference():
    """Import the reference evaluators. Returns a namespace with TaoEval, Tao,
    TaoResults, LVISEval, LVIS, LVISResults and the tao eval module."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    install_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import tao_amodal.evaluation.tao_amodal as rt
    import tao_amodal.evaluation.tao_amodal.eval as rte
    import tao_amodal.evaluation.lvis_amodal as rl
    ns = types.SimpleNamespace(
        TaoEval=rt.TaoEval, Tao=rt.Tao, TaoResults=rt.TaoResults,
        LVISEval=rl.LVISEval, LVIS=rl.LVIS, LVISResults=rl.LVISResults,
        tao_eval_module=rte)
    return ns


def run_reference_driver(annotation: str, track_result: str, output_log: str):
    """Run tools/eval_on_tao_amodal.py unmodified (cwd must be tools/, :14).
    Returns (stdout_text, stderr_text)."""
    install_shims()
    tools = os.path.join(REF_ROOT, "tools")
    argv, cwd = sys.argv, os.getcwd()
    out, err = io.StringIO(), io.StringIO()
    annotation, track_result, output_log = map(os.path.abspath,
                                               (annotation, track_result, output_log))
    try:
        os.chdir(tools)
        sys.argv = ["eval_on_tao_amodal.py", "--track_result", track_result,
                    "--output_log", output_log, "--annotation", annotation]
        with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
            runpy.run_path(os.path.join(tools, "eval_on_tao_amodal.py"), run_name="__main__")
    finally:
        os.chdir(cwd)
        sys.argv = argv
        import logging
        lg = logging.getLogger("__main__")
        for h in list(lg.handlers):
            h.close()
            lg.removeHandler(h)
    return out.getvalue(), err.getvalue()
