"""TEST INFRASTRUCTURE — BASELINE-size fixtures from the UNMODIFIED reference.

    python -m oracle.make_fullsize_golden cfg2            # configs[1], both evaluators, ~40 s
    python -m oracle.make_fullsize_golden cfg3x4          # 4 whole cfg3 videos, both, ~2 min
    python -m oracle.make_fullsize_golden cfg5_tao        # configs[4], TaoEval half, ~8 min

Runs the reference (``/root/reference`` or ``baseline/_ref``) exactly as its CLI does
(tools/eval_on_tao_amodal.py:68-151) on the seeded synthetic set of that name and stores what a
parity test needs WITHOUT the input (the generator is deterministic: tao_amodal_b200/synth.py):

* ``*_recall``, ``*_tp_cnt``, ``*_fp_cnt`` — the full tensors (small);
* ``*_precision_sha256`` — digest of the precision tensor's bytes (C order, float64): equality
  is bit-exactness of all T x R x C x cfg entries; ``*_precision_sum`` / ``_n_valid`` as a
  readable second check;
* ``*_results`` + keys — the summary table (19 track metrics / 25 frame metrics).

The LVISEval half of cfg5 cannot run in the reference (its empty-cell bookkeeping needs ~135 GB,
SURVEY.md §3.4), so ``cfg5_tao`` holds the TaoEval half only — the "full TrackAP table" of
BASELINE.json configs[4].
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile
import time

import numpy as np

from . import golden_io, ref_bench, ref_shims
from .make_golden import counts_from_pointers, reference_make_track_ids_unique

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

SETS = {
    "cfg2": dict(workload="cfg2", videos=0, lvis=True),
    "cfg3x4": dict(workload="cfg3", videos=4, lvis=True),
    "cfg5_tao": dict(workload="cfg5", videos=0, lvis=False),
}


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def pack(prefix: str, ev, out: dict):
    prec, rec = ev.eval["precision"], ev.eval["recall"]
    out[prefix + "_precision_sha256"] = np.asarray(digest(prec))
    out[prefix + "_precision_shape"] = np.asarray(prec.shape)
    out[prefix + "_precision_sum"] = np.asarray(float(prec[prec > -1].sum()))
    out[prefix + "_precision_n_valid"] = np.asarray(int((prec > -1).sum()))
    out[prefix + "_recall"] = rec
    tp, fp = counts_from_pointers(ev.eval, rec.shape)
    out[prefix + "_tp_cnt"], out[prefix + "_fp_cnt"] = tp, fp
    out[prefix + "_results"] = golden_io.results_vector(ev.results)
    out[prefix + "_results_keys"] = np.asarray(golden_io.results_keys(ev.results))


def main(argv):
    sys.path.insert(0, ROOT)
    from tao_amodal_b200 import synth
    root = ref_bench.find_reference()
    ref_shims.REF_ROOT = root
    os.environ["TAO_AMODAL_REF"] = root
    ref = ref_shims.load_reference()
    uniq = reference_make_track_ids_unique()
    for name in argv or list(SETS):
        spec = SETS[name]
        over = {"videos": spec["videos"]} if spec["videos"] else {}
        gt, dt = synth.generate_named(spec["workload"], **over)
        out = {"workload": np.asarray(spec["workload"]), "videos": np.asarray(spec["videos"]),
               "n_pred_boxes": np.asarray(dt.n()), "n_gt_boxes": np.asarray(gt.n_anns())}
        timing = {}
        with tempfile.TemporaryDirectory() as td:
            ap, rp = os.path.join(td, "gt.json"), os.path.join(td, "dt.json")
            json.dump(gt.to_dict(), open(ap, "w"))
            json.dump(dt.to_list(), open(rp, "w"))
            if spec["lvis"]:
                t0 = time.perf_counter()
                le = ref.LVISEval(ap, rp, "bbox")
                le.run()
                timing["lvis_s"] = time.perf_counter() - t0
                pack("lvis", le, out)
                del le
            t0 = time.perf_counter()
            tao_gt = ref.Tao(ap)
            res = json.load(open(rp))
            uniq(res)
            te = ref.TaoEval(tao_gt, res)
            te.run()
            timing["tao_s"] = time.perf_counter() - t0
            pack("tao", te, out)
        out["reference_wall_s"] = np.asarray(json.dumps(timing))
        path = os.path.join(GOLDEN_DIR, "full_%s.npz" % name)
        np.savez_compressed(path, **out)
        print("%-10s -> %s (%d KB) %s  TAO AP=%.6f" % (
            name, path, os.path.getsize(path) // 1024, timing, out["tao_results"][0]), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
