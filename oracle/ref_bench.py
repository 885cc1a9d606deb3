"""TEST / BENCH INFRASTRUCTURE — times the UNMODIFIED reference on a fixed sample.

Used only by ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` leg).  Never imported
by the product.  The reference tree is looked up in this order:

1. ``$TAO_AMODAL_REF``;
2. ``baseline/_ref`` — a verbatim copy of ``tao_amodal/evaluation`` + ``tools/eval_on_tao_amodal.py``
   made by ``__graft_entry__.build()`` where ``/root/reference`` exists.  The directory is
   git-ignored (no reference source enters the history) but travels to the GPU box with the
   snapshot, like ``oracle/_ref``;
3. ``/root/reference``.

What is timed is what the reference's CLI runs (tools/eval_on_tao_amodal.py:68-151), on one
process / one core — the reference is single-threaded:

    LVISEval(annotation.json, results.json, "bbox").run()
    TaoEval(Tao(annotation.json), make_track_ids_unique(json.load(results.json))).run()

including its JSON parsing, index building and deep copies.  The sample is FIXED (no
time-budget-dependent prefix): whole videos of the named synthetic workload, generated from
the workload's own seed, written as the JSON files the CLI takes.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference():
    for cand in (os.environ.get("TAO_AMODAL_REF"), os.path.join(ROOT, "baseline", "_ref"),
                 "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "tao_amodal", "evaluation")):
            return cand
    return None


def write_sample(workload: str, videos: int, out_dir: str, frames: int = 0):
    """The first `videos` videos of the workload (its own seed), optionally only their first
    `frames` frames, as annotation.json / results.json.  Returns (annotation path, results
    path, box_pairs, description); box_pairs is counted on exactly the data written."""
    sys.path.insert(0, ROOT)
    from tao_amodal_b200 import prep, synth
    from tao_amodal_b200.columnar import DtColumns, GtColumns
    cfg = synth.CONFIGS[workload]
    gt, dt = synth.generate_named(workload, videos=videos, seed=cfg.seed)
    gd, dl = gt.to_dict(), dt.to_list()
    if frames and frames < cfg.frames:
        keep = {im["id"] for im in gd["images"] if im["frame_index"] < frames}
        gd["images"] = [im for im in gd["images"] if im["id"] in keep]
        gd["annotations"] = [a for a in gd["annotations"] if a["image_id"] in keep]
        live = {a["track_id"] for a in gd["annotations"]}
        gd["tracks"] = [t for t in gd["tracks"] if t["id"] in live]
        dl = [r for r in dl if r["image_id"] in keep]
        gt, dt = GtColumns.from_dict(gd), DtColumns.from_list(dl)
    lvis_plan = prep.prepare_lvis(gt, dt)
    d2 = dt.copy()
    prep.make_track_ids_unique(d2)
    tao_plan = prep.prepare_tao(gt, d2)
    pairs = prep.count_box_pair_visits(tao_plan) + prep.count_box_pair_visits(lvis_plan)
    ap, rp = os.path.join(out_dir, "annotation.json"), os.path.join(out_dir, "results.json")
    json.dump(gd, open(ap, "w"))
    json.dump(dl, open(rp, "w"))
    desc = ("first %d of %d frames of %d %s-shaped video(s) (%d predicted + %d GT tracks per video, "
            "%d categories), seed %d" % (frames or cfg.frames, cfg.frames, videos, workload,
                                         cfg.pred_tracks, cfg.gt_tracks, cfg.categories, cfg.seed))
    return ap, rp, int(pairs), desc


_REF = None


def _load():
    global _REF
    if _REF is None:
        root = find_reference()
        if root is None:
            raise RuntimeError("reference tree not found (baseline/_ref, /root/reference)")
        os.environ["TAO_AMODAL_REF"] = root
        from . import ref_shims
        ref_shims.REF_ROOT = root
        ref = ref_shims.load_reference()
        from .make_golden import reference_make_track_ids_unique
        _REF = (ref, reference_make_track_ids_unique())
    return _REF


def run_once(ap: str, rp: str):
    """One pass of both evaluators exactly as the reference CLI drives them.  Returns
    (seconds, {"lvis_AP": .., "tao_AP": ..})."""
    ref, uniq = _load()
    import contextlib
    import io
    import logging
    logging.disable(logging.CRITICAL)
    try:
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            le = ref.LVISEval(ap, rp, "bbox")
            le.run()
            tao_gt = ref.Tao(ap)
            res = json.load(open(rp))
            uniq(res)
            te = ref.TaoEval(tao_gt, res)
            te.run()
        dt = time.perf_counter() - t0
    finally:
        logging.disable(logging.NOTSET)
    return dt, {"lvis_AP": float(le.results["AP"]), "tao_AP": float(te.results["AP"])}


def run_subprocess(ap: str, rp: str, numba_disable_jit: bool, repeats: int = 1):
    """The same pass in a fresh interpreter (NUMBA_DISABLE_JIT has to be set before numba is
    imported).  Returns the list of per-pass seconds."""
    env = dict(os.environ)
    if numba_disable_jit:
        env["NUMBA_DISABLE_JIT"] = "1"
    code = ("import sys, json; sys.path.insert(0, %r); from oracle import ref_bench as rb; "
            "print(json.dumps([rb.run_once(%r, %r)[0] for _ in range(%d)]))" % (ROOT, ap, rp, repeats))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                         check=True)
    return json.loads(out.stdout.strip().splitlines()[-1])


if __name__ == "__main__":
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    nv = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    fr = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    with tempfile.TemporaryDirectory() as td:
        ap, rp, pairs, desc = write_sample(wl, nv, td, fr)
        print(desc, "box pairs", pairs, "reference at", find_reference())
        s, res = run_once(ap, rp)
        print("first pass (numba JIT compile included) %.2f s" % s, res)
        s, res = run_once(ap, rp)
        print("second pass %.2f s -> %.0f box-pairs/s" % (s, pairs / s))
        nj = run_subprocess(ap, rp, True, 1)
        print("NUMBA_DISABLE_JIT=1: %.2f s -> %.0f box-pairs/s" % (nj[0], pairs / nj[0]))
