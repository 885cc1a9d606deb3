"""TEST INFRASTRUCTURE — generate tests/golden from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden tiny small # some cases

For every case of oracle/cases.py it runs the reference exactly as its CLI
does (tools/eval_on_tao_amodal.py:68-151): ``LVISEval(annotation_path,
result_path, "bbox").run()`` and ``TaoEval(Tao(path), json.load(results) after
make_track_ids_unique).run()``, and stores every intermediate the parity tests
compare: IoU matrices, per-cell match / ignore decisions, precision, recall,
TP/FP counts, the summary metrics, and the CLI's log file + stdout.
"""
from __future__ import annotations

import ast
import itertools
import json
import os
import sys
import tempfile
from collections import defaultdict

import numpy as np

from . import cases, golden_io, ref_shims

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")


def reference_make_track_ids_unique():
    """Extract the function object from the CLI script without running the script."""
    path = os.path.join(ref_shims.REF_ROOT, "tools", "eval_on_tao_amodal.py")
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef)
          and n.name == "make_track_ids_unique"][0]
    ns = {"tqdm": lambda x: x, "itertools": itertools, "defaultdict": defaultdict}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["make_track_ids_unique"]


def counts_from_pointers(ev, shape):
    """TP/FP totals per (thr, cat, ranges...) from eval['dt_pointers']."""
    tp = np.zeros(shape, dtype=np.int64)
    fp = np.zeros(shape, dtype=np.int64)

    def walk(node, idx):
        if "tps" in node:
            sl = (slice(None),) + tuple(idx)
            tp[sl] = node["tps"].sum(1)
            fp[sl] = node["fps"].sum(1)
            return
        for k, v in node.items():
            if isinstance(v, dict):
                walk(v, idx + [k])
    walk(ev["dt_pointers"], [])
    return tp, fp


def run_case(name: str, ref) -> dict:
    gt, res = cases.build(name)
    # the inputs travel with the golden so the tests never depend on RNG stream stability
    out = {"in_gt_json": np.asarray(json.dumps(gt)), "in_dt_json": np.asarray(json.dumps(res))}
    with tempfile.TemporaryDirectory() as td:
        ap, rp, lp = (os.path.join(td, f) for f in ("gt.json", "dt.json", "eval.log"))
        json.dump(gt, open(ap, "w"))
        json.dump(res, open(rp, "w"))

        # ---- frame AP (tools/eval_on_tao_amodal.py:68-116)
        le = ref.LVISEval(ap, rp, "bbox")
        le.run()
        n_img = len(le.params.img_ids)
        n_r = len(le.params.visibility_rng)
        cells = {}
        for flat, e in enumerate(le.eval_imgs):
            if e is not None:
                c, rem = divmod(flat, n_r * n_img)
                r, i = divmod(rem, n_img)
                cells[c, r, i] = e
        for k, v in golden_io.flatten_cells(cells).items():
            out["lvis_" + k] = v
        for k, v in golden_io.flatten_ious(le.ious).items():
            out["lvis_" + k] = v
        out["lvis_precision"] = le.eval["precision"]
        out["lvis_recall"] = le.eval["recall"]
        tp, fp = counts_from_pointers(le.eval, le.eval["recall"].shape)
        out["lvis_tp_cnt"], out["lvis_fp_cnt"] = tp, fp
        out["lvis_results"] = golden_io.results_vector(le.results)
        out["lvis_results_keys"] = np.asarray(golden_io.results_keys(le.results))

        # ---- track AP (tools/eval_on_tao_amodal.py:118-151)
        tao_gt = ref.Tao(ap)
        res2 = json.load(open(rp))
        reference_make_track_ids_unique()(res2)
        te = ref.TaoEval(tao_gt, res2)
        te.run()
        for k, v in golden_io.flatten_cells(dict(te.eval_vids)).items():
            out["tao_" + k] = v
        for k, v in golden_io.flatten_ious(te.ious).items():
            out["tao_" + k] = v
        out["tao_precision"] = te.eval["precision"]
        out["tao_recall"] = te.eval["recall"]
        tp, fp = counts_from_pointers(te.eval, te.eval["recall"].shape)
        out["tao_tp_cnt"], out["tao_fp_cnt"] = tp, fp
        out["tao_results"] = golden_io.results_vector(te.results)
        out["tao_results_keys"] = np.asarray(golden_io.results_keys(te.results))

        # ---- the whole CLI: log file + stdout (paths normalised)
        so, se = ref_shims.run_reference_driver(ap, rp, lp)
        log = open(lp).read().replace(rp, "<RESULTS>").replace(ap, "<ANNOTATION>")
        out["cli_log"] = np.asarray(log)
        out["cli_stdout"] = np.asarray(so.replace(rp, "<RESULTS>").replace(ap, "<ANNOTATION>"))
    return out


def run_nocats(name: str, ref) -> dict:
    """Params.use_cats = 0 (eval.py:257-260, :293-303): all categories of a video / image are
    evaluated as one.  LVISEval.summarize indexes the single pseudo category with the
    per-category frequency groups and raises IndexError in the reference, so only its
    accumulate() outputs are stored."""
    gt, res = cases.build(name)
    out = {"in_gt_json": np.asarray(json.dumps(gt)), "in_dt_json": np.asarray(json.dumps(res))}
    with tempfile.TemporaryDirectory() as td:
        ap, rp = (os.path.join(td, f) for f in ("gt.json", "dt.json"))
        json.dump(gt, open(ap, "w"))
        json.dump(res, open(rp, "w"))
        le = ref.LVISEval(ap, rp, "bbox")
        le.params.use_cats = 0
        le.evaluate()
        le.accumulate()
        out["lvis_precision"], out["lvis_recall"] = le.eval["precision"], le.eval["recall"]
        try:
            le.summarize()
            out["lvis_summarize_error"] = np.asarray("")
        except Exception as e:          # noqa: BLE001 - recording the reference's behaviour
            out["lvis_summarize_error"] = np.asarray(type(e).__name__)
        res2 = json.load(open(rp))
        reference_make_track_ids_unique()(res2)
        te = ref.TaoEval(ref.Tao(ap), res2)
        te.params.use_cats = 0
        te.run()
        out["tao_precision"], out["tao_recall"] = te.eval["precision"], te.eval["recall"]
        out["tao_results"] = golden_io.results_vector(te.results)
        for k, v in golden_io.flatten_cells(dict(te.eval_vids)).items():
            out["tao_" + k] = v
    return out


def run_alt_iou(name: str, mode: str, ref) -> dict:
    """TaoEval with Params.iou_3d_type = avg_iou / imagenetvid (eval.py:51-70, 99-117, 329-334)."""
    gt, res = cases.build(name)
    out = {"in_gt_json": np.asarray(json.dumps(gt)), "in_dt_json": np.asarray(json.dumps(res))}
    with tempfile.TemporaryDirectory() as td:
        ap, rp = (os.path.join(td, f) for f in ("gt.json", "dt.json"))
        json.dump(gt, open(ap, "w"))
        json.dump(res, open(rp, "w"))
        res2 = json.load(open(rp))
        reference_make_track_ids_unique()(res2)
        te = ref.TaoEval(ref.Tao(ap), res2, iou_3d_type=mode)
        te.run()
        for k, v in golden_io.flatten_ious(te.ious).items():
            out["tao_" + k] = v
        for k, v in golden_io.flatten_cells(dict(te.eval_vids)).items():
            out["tao_" + k] = v
        out["tao_precision"], out["tao_recall"] = te.eval["precision"], te.eval["recall"]
        out["tao_results"] = golden_io.results_vector(te.results)
    return out


def run_segm(name: str, ref) -> dict:
    """LVISEval(annotation, results, "segm").run() (lvis_amodal/eval.py:54-57, :70-72, :180-191):
    annotations become masks through the reference tree's own maskApi.c (oracle/_ref)."""
    import pycocotools.mask as pm
    assert pm.BACKEND.startswith("reference C"), "segm goldens need oracle/_ref (make -C oracle)"
    gt, res = cases.build(name)
    out = {"in_gt_json": np.asarray(json.dumps(gt)), "in_dt_json": np.asarray(json.dumps(res))}
    with tempfile.TemporaryDirectory() as td:
        ap, rp = (os.path.join(td, f) for f in ("gt.json", "dt.json"))
        json.dump(gt, open(ap, "w"))
        json.dump(res, open(rp, "w"))
        le = ref.LVISEval(ap, rp, "segm")
        le.run()
        n_img, n_r = len(le.params.img_ids), len(le.params.visibility_rng)
        cells = {}
        for flat, e in enumerate(le.eval_imgs):
            if e is not None:
                c, rem = divmod(flat, n_r * n_img)
                r, i = divmod(rem, n_img)
                cells[c, r, i] = e
        for k, v in golden_io.flatten_cells(cells).items():
            out["lvis_" + k] = v
        for k, v in golden_io.flatten_ious(le.ious).items():
            out["lvis_" + k] = v
        out["lvis_precision"], out["lvis_recall"] = le.eval["precision"], le.eval["recall"]
        tp, fp = counts_from_pointers(le.eval, le.eval["recall"].shape)
        out["lvis_tp_cnt"], out["lvis_fp_cnt"] = tp, fp
        out["lvis_results"] = golden_io.results_vector(le.results)
        out["lvis_results_keys"] = np.asarray(golden_io.results_keys(le.results))
    return out


def main(argv):
    names = argv or list(cases.CASES)
    ref = ref_shims.load_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    if names == ["alt_iou"]:
        for mode in ("avg_iou", "imagenetvid"):
            out = run_alt_iou("small", mode, ref)
            path = os.path.join(GOLDEN_DIR, "small_%s.npz" % mode)
            np.savez_compressed(path, **out)
            print("%-12s -> %s  TAO AP=%.6f" % (mode, path, out["tao_results"][0]))
        return
    if names == ["segm"]:
        for n in cases.SEGM_CASES:
            out = run_segm(n, ref)
            path = os.path.join(GOLDEN_DIR, n + ".npz")
            np.savez_compressed(path, **out)
            print("%-12s -> %s (%d KB)  LVIS segm AP=%.6f, %d IoU entries, %d positive" % (
                n, path, os.path.getsize(path) // 1024, out["lvis_results"][0],
                out["lvis_iou_vals"].size, int((out["lvis_iou_vals"] > 0).sum())))
        return
    if names == ["nocats"]:
        for n in ("small", "edge_mix"):
            out = run_nocats(n, ref)
            path = os.path.join(GOLDEN_DIR, n + "_nocats.npz")
            np.savez_compressed(path, **out)
            print("%-18s -> %s  TAO AP=%.6f  lvis summarize: %r" % (
                n, path, out["tao_results"][0], str(out["lvis_summarize_error"])))
        return
    for n in names:
        out = run_case(n, ref)
        path = os.path.join(GOLDEN_DIR, n + ".npz")
        np.savez_compressed(path, **out)
        print("%-18s -> %s (%d KB)  TAO AP=%.6f  LVIS AP=%.6f" % (
            n, path, os.path.getsize(path) // 1024, out["tao_results"][0], out["lvis_results"][0]))


if __name__ == "__main__":
    main(sys.argv[1:])
