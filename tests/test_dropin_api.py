"""The drop-in Python classes and the CLI against the unmodified reference's outputs
(tests/golden: metrics, per-cell structures, log file and stdout of the reference CLI)."""
import contextlib
import io
import json
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_inputs
from oracle import golden_io

pytestmark = pytest.mark.gpu


def _write(tmp_path, golden):
    gt, res = golden_inputs(golden)
    ap, rp = str(tmp_path / "gt.json"), str(tmp_path / "dt.json")
    json.dump(gt, open(ap, "w"))
    json.dump(res, open(rp, "w"))
    return ap, rp


def test_cli_log_and_stdout_identical(golden, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import eval_on_tao_amodal as cli
    ap, rp = _write(tmp_path, golden)
    lp = str(tmp_path / "out" / "eval.log")
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        assert cli.main(["--track_result", rp, "--output_log", lp, "--annotation", ap]) == 0
    log = open(lp).read().replace(rp, "<RESULTS>").replace(ap, "<ANNOTATION>")
    assert log == str(golden["cli_log"])
    assert out.getvalue() == str(golden["cli_stdout"])


def test_taoeval_attributes_match_reference(golden, tmp_path):
    from tao_amodal_b200.evaluation.tao_amodal import Tao, TaoEval
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import eval_on_tao_amodal as cli
    ap, rp = _write(tmp_path, golden)
    res = json.load(open(rp))
    cli.make_track_ids_unique(res)
    te = TaoEval(Tao(ap), res)
    te.run()
    assert golden_io.results_keys(te.results) == [str(k) for k in golden["tao_results_keys"]]
    assert np.array_equal(golden["tao_results"], golden_io.results_vector(te.results))
    assert np.array_equal(golden["tao_precision"], te.eval["precision"])
    assert np.array_equal(golden["tao_recall"], te.eval["recall"])
    assert te.eval["counts"] == [10, 101, te.eval["precision"].shape[2], 5, 4]
    fc = golden_io.flatten_cells(dict(te.eval_vids))
    for k, v in fc.items():
        assert np.array_equal(golden["tao_" + k], v), k
    exact = golden["_name"] != "small_float"
    fi = golden_io.flatten_ious(dict(te.ious))
    assert np.array_equal(golden["tao_iou_keys"], fi["iou_keys"])
    if exact:
        assert np.array_equal(golden["tao_iou_vals"], fi["iou_vals"])
    # dt_pointers: TP / FP totals per (category, area, time) cell
    tp = np.zeros(golden["tao_tp_cnt"].shape, dtype=np.int64)
    for c, by_a in te.eval["dt_pointers"].items():
        for a, by_t in by_a.items():
            for t, node in by_t.items():
                if node:
                    tp[:, c, a, t] = node["tps"].sum(1)
    assert np.array_equal(golden["tao_tp_cnt"], tp)


def test_lviseval_attributes_match_reference(golden, tmp_path):
    from tao_amodal_b200.evaluation.lvis_amodal import LVISEval
    ap, rp = _write(tmp_path, golden)
    le = LVISEval(ap, rp, "bbox")
    le.run()
    assert golden_io.results_keys(le.results) == [str(k) for k in golden["lvis_results_keys"]]
    assert np.array_equal(golden["lvis_results"], golden_io.results_vector(le.results))
    assert np.array_equal(golden["lvis_precision"], le.eval["precision"])
    n_img, n_r = len(le.params.img_ids), len(le.params.visibility_rng)
    cells = {}
    for flat, e in enumerate(le.eval_imgs):
        if e is not None:
            c, rem = divmod(flat, n_r * n_img)
            r, i = divmod(rem, n_img)
            cells[c, r, i] = e
    for k, v in golden_io.flatten_cells(cells).items():
        assert np.array_equal(golden["lvis_" + k], v), k
    fi = golden_io.flatten_ious(dict(le.ious))
    assert np.array_equal(golden["lvis_iou_vals"], fi["iou_vals"])


def test_constructor_errors_match_reference(tmp_path):
    from tao_amodal_b200.evaluation.tao_amodal import Tao, TaoEval, TaoResults
    from tao_amodal_b200 import synth
    gtc, dtc = synth.generate_named("tiny")
    gt, res = gtc.to_dict(), dtc.to_list()
    tao = Tao(gt)
    with pytest.raises(ValueError):
        TaoEval(tao, res, iou_type="keypoints")
    with pytest.raises(TypeError):
        TaoEval(42, res)
    with pytest.raises(TypeError):
        TaoEval(tao, 42)
    bad = [dict(r) for r in res]
    bad[-1]["video_id"] = bad[0]["video_id"] + 1
    bad[-1]["track_id"] = bad[0]["track_id"]
    with pytest.raises(AssertionError, match="appears in more than one video"):
        TaoResults(tao, bad)
    bad = [dict(r) for r in res]
    bad[0]["image_id"] = 10 ** 9
    with pytest.raises(AssertionError, match="Results do not correspond"):
        TaoResults(tao, bad)


def test_use_cats_zero_through_the_classes(tmp_path):
    from conftest import load_golden
    from tao_amodal_b200.evaluation.lvis_amodal import LVISEval
    from tao_amodal_b200.evaluation.tao_amodal import Tao, TaoEval
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import eval_on_tao_amodal as cli
    g = load_golden("small_nocats")
    ap, rp = _write(tmp_path, g)
    res = json.load(open(rp))
    cli.make_track_ids_unique(res)
    te = TaoEval(Tao(ap), res)
    te.params.use_cats = 0
    te.run()
    assert np.array_equal(g["tao_precision"], te.eval["precision"])
    assert np.array_equal(g["tao_results"], golden_io.results_vector(te.results))
    le = LVISEval(ap, rp, "bbox")
    le.params.use_cats = 0
    le.evaluate()
    le.accumulate()
    assert np.array_equal(g["lvis_precision"], le.eval["precision"])
    with pytest.raises(IndexError):        # the reference fails the same way (frequency groups)
        le.summarize()
