"""LVISEval(iou_type="segm") (lvis_amodal/eval.py:54-57, :70-72, :180-191) against goldens of the
unmodified reference run with its own maskApi.c: host logic + per-thread kernel arithmetic on
the CPU, the CUDA path with ``-m gpu``."""
import copy

import numpy as np
import pytest

from conftest import golden_inputs, load_golden
from plan_backends import compare_with_golden, run_hostsim

SEGM_CASES = ["segm_box", "segm_poly", "segm_rle"]


def _evaluator(case):
    from tao_amodal_b200.evaluation.lvis_amodal import LVIS, LVISEval
    g = load_golden(case)
    gt, res = golden_inputs(g)
    ev = LVISEval(LVIS(gt), copy.deepcopy(res), "segm")
    return g, ev


@pytest.mark.parametrize("case", SEGM_CASES)
def test_segm_plan_hostsim_matches_reference(case):
    g, ev = _evaluator(case)
    ev._prepare()
    plan = ev._plan
    assert plan.masks is not None and plan.masks["dt"][0].size == plan.n_dt + 1
    out = run_hostsim(plan)
    compare_with_golden(g, "lvis_", plan, out, exact_iou=True)


def test_segm_results_without_bbox_get_box_and_area_from_the_mask():
    """results.py:58-66: area = mask area, bbox = toBbox(mask)."""
    from tao_amodal_b200.evaluation.lvis_amodal import LVIS, LVISResults
    g = load_golden("segm_rle")
    gt, res = golden_inputs(g)
    res = copy.deepcopy(res)
    assert "bbox" not in res[0]
    r = LVISResults(LVIS(gt), res)
    a = r.dataset["annotations"][0]
    assert a["area"] > 0 and len(a["bbox"]) == 4 and a["id"] == 1


def test_ann_to_rle_and_mask_accessors():
    from tao_amodal_b200.evaluation.lvis_amodal import LVIS
    g = load_golden("segm_box")
    gt, _ = golden_inputs(g)
    lv = LVIS(gt)
    for ann in gt["annotations"][:12]:
        rle = lv.ann_to_rle(ann)
        m = lv.ann_to_mask(ann)
        img = lv.imgs[ann["image_id"]]
        assert rle["size"] == [img["height"], img["width"]] and m.shape == (img["height"], img["width"])
        assert m.dtype == np.uint8 and set(np.unique(m)) <= {0, 1}


@pytest.mark.gpu
@pytest.mark.parametrize("case", SEGM_CASES)
def test_segm_device_path_matches_reference(case):
    from tao_amodal_b200.evaluation._common import get_engine
    g, ev = _evaluator(case)
    ev.run()
    assert np.array_equal(g["lvis_precision"], ev.eval["precision"])
    assert np.array_equal(g["lvis_recall"], ev.eval["recall"])
    from oracle import golden_io
    assert np.array_equal(g["lvis_results"], golden_io.results_vector(ev.results))
    out = get_engine(ev.device).evaluate_device(ev._dev, detail=True)
    compare_with_golden(g, "lvis_", ev._plan, out, exact_iou=True)
    # the reference-shaped accessors
    flat = golden_io.flatten_ious(dict(ev.ious.items()))
    assert np.array_equal(g["lvis_iou_vals"], flat["iou_vals"])
