"""Host logic (prep.py: ordering / filtering / CSR layout) + the kernels' per-thread
arithmetic (csrc/ta_device_fns.cuh built for the host, tests/hostsim) against goldens from the
unmodified reference.  CPU only; the CUDA kernels themselves are checked in test_gpu_parity.py."""
import numpy as np
import pytest

from conftest import golden_inputs
from plan_backends import compare_with_golden, plans_from_json, run_hostsim


def test_tao_plan_hostsim_matches_reference(golden):
    gt, res = golden_inputs(golden)
    tao_plan, _ = plans_from_json(gt, res)
    off_grid = golden["_name"] == "small_float"
    # the tiled kernel re-associates the union (U = DA + GA - I); exact on grid data
    out = run_hostsim(tao_plan, "3d_iou")
    compare_with_golden(golden, "tao_", tao_plan, out, exact_iou=not off_grid, iou_atol=1e-12)


def test_lvis_plan_hostsim_matches_reference(golden):
    gt, res = golden_inputs(golden)
    _, lvis_plan = plans_from_json(gt, res)
    out = run_hostsim(lvis_plan)
    compare_with_golden(golden, "lvis_", lvis_plan, out, exact_iou=True)


@pytest.mark.parametrize("case", ["small_nocats", "edge_mix_nocats"])
def test_use_cats_zero_matches_reference(case):
    """Params.use_cats = 0: one pseudo category per video / image (eval.py:257-260, :293-303)."""
    from conftest import load_golden
    from oracle import golden_io
    from tao_amodal_b200 import materialize, prep
    from tao_amodal_b200.columnar import DtColumns, GtColumns
    g = load_golden(case)
    gt_d, res = golden_inputs(g)
    gt, dt = GtColumns.from_dict(gt_d), DtColumns.from_list(res)
    lvis_plan = prep.prepare_lvis(gt, dt, use_cats=False)
    out = run_hostsim(lvis_plan)
    assert np.array_equal(g["lvis_precision"], out.precision)
    assert np.array_equal(g["lvis_recall"], out.recall)
    prep.make_track_ids_unique(dt)
    tao_plan = prep.prepare_tao(gt, dt, use_cats=False)
    out = run_hostsim(tao_plan)
    assert np.array_equal(g["tao_precision"], out.precision.reshape(g["tao_precision"].shape))
    assert np.array_equal(g["tao_recall"], out.recall.reshape(g["tao_recall"].shape))
    cells = materialize.cells_dict(tao_plan, 10, out)
    for k, v in golden_io.flatten_cells(cells).items():
        assert np.array_equal(g["tao_" + k], v), k


@pytest.mark.parametrize("mode", ["avg_iou", "imagenetvid"])
def test_alternative_iou_modes_match_reference(mode):
    """Params.iou_3d_type = avg_iou / imagenetvid (eval.py:51-70, :99-117).  imagenetvid is a
    ratio of counts (exact); avg_iou is numpy's mean over CPython-set-ordered frames in the
    reference, a sequential ascending-frame sum here: equal to 1e-12, same decisions."""
    from conftest import load_golden
    from oracle import golden_io
    from tao_amodal_b200 import materialize
    g = load_golden("small_" + mode)
    gt, res = golden_inputs(g)
    tao_plan, _ = plans_from_json(gt, res)
    out = run_hostsim(tao_plan, mode)
    flat = golden_io.flatten_ious(materialize.iou_dict(tao_plan, out.iou))
    assert np.array_equal(g["tao_iou_keys"], flat["iou_keys"])
    if mode == "imagenetvid":
        assert np.array_equal(g["tao_iou_vals"], flat["iou_vals"])
    else:
        np.testing.assert_allclose(flat["iou_vals"], g["tao_iou_vals"], rtol=0, atol=1e-12)
    cells = materialize.cells_dict(tao_plan, 10, out)
    for k, v in golden_io.flatten_cells(cells).items():
        assert np.array_equal(g["tao_" + k], v), k
    assert np.array_equal(g["tao_precision"], out.precision.reshape(g["tao_precision"].shape))
