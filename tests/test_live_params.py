"""Evaluator knobs (``Params`` fields a user may change before ``run()``) against the UNMODIFIED
reference, where the reference tree exists: subsets of videos / images / categories, custom IoU
and recall thresholds, custom area / duration / visibility ranges, a smaller max_dets.  The
drop-in classes build their plan from the same ``params``; IoU, matching and PR then run through
the host build of the kernels' per-thread functions.  CPU only."""
import copy
import json

import numpy as np
import pytest

from oracle import ref_shims
from plan_backends import random_small_set, run_hostsim

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(),
                                reason="reference tree not present on this machine")


@pytest.fixture(scope="module")
def ref():
    return ref_shims.load_reference()


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    gt, res = random_small_set(77)
    d = tmp_path_factory.mktemp("params")
    ap, rp = str(d / "gt.json"), str(d / "dt.json")
    json.dump(gt, open(ap, "w"))
    json.dump(res, open(rp, "w"))
    return gt, res, ap, rp


def _tao_pair(ref, files, mutate, max_dets=300):
    from oracle.make_golden import reference_make_track_ids_unique
    from tao_amodal_b200.evaluation.tao_amodal import Tao, TaoEval, TaoResults
    gt, res, ap, rp = files
    res_r = json.load(open(rp))
    reference_make_track_ids_unique()(res_r)
    rgt = ref.Tao(ap)
    te = ref.TaoEval(rgt, ref.TaoResults(rgt, res_r, max_dets=max_dets))
    mutate(te.params)
    te.evaluate()
    te.accumulate()
    res_m = json.load(open(rp))
    reference_make_track_ids_unique()(res_m)
    mgt = Tao(copy.deepcopy(gt))
    ev = TaoEval(mgt, TaoResults(mgt, res_m, max_dets=max_dets))
    mutate(ev.params)
    ev.params.vid_ids = list(np.unique(ev.params.vid_ids))       # evaluate() does this, eval.py:255
    ev._prepare()
    out = run_hostsim(ev._plan, iou_thrs=ev.params.iou_thrs, rec_thrs=ev.params.rec_thrs)
    return te, ev, out


def _lvis_pair(ref, files, mutate, max_dets=300):
    from tao_amodal_b200.evaluation.lvis_amodal import LVIS, LVISEval, LVISResults
    gt, res, ap, rp = files
    rgt = ref.LVIS(ap)
    le = ref.LVISEval(rgt, ref.LVISResults(rgt, json.load(open(rp)), max_dets=max_dets), "bbox")
    mutate(le.params)
    le.evaluate()
    le.accumulate()
    mgt = LVIS(copy.deepcopy(gt))
    ev = LVISEval(mgt, LVISResults(mgt, json.load(open(rp)), max_dets=max_dets), "bbox")
    mutate(ev.params)
    ev.params.img_ids = list(np.unique(ev.params.img_ids))
    ev._prepare()
    out = run_hostsim(ev._plan, iou_thrs=ev.params.iou_thrs, rec_thrs=ev.params.rec_thrs)
    return le, ev, out


def _subset_vids(p):
    p.vid_ids = p.vid_ids[1::2]


def _subset_cats(p):
    p.cat_ids = p.cat_ids[::3]


def _thresholds(p):
    p.iou_thrs = np.array([0.3, 0.5, 0.75])
    p.rec_thrs = np.linspace(0.0, 1.0, 11)


def _tao_ranges(p):
    p.area_rng = [[0 ** 2, 1e5 ** 2], [0, 50 ** 2], [50 ** 2, 1e5 ** 2], [0, 1e5 ** 2], [0, 1e5 ** 2]]
    p.time_rng = [[0, 1e5], [0, 5], [5, 20], [20, 1e5]]


@pytest.mark.parametrize("mutate", [_subset_vids, _subset_cats, _thresholds, _tao_ranges],
                         ids=lambda f: f.__name__.strip("_"))
def test_tao_params(ref, files, mutate):
    te, ev, out = _tao_pair(ref, files, mutate)
    shp = te.eval["precision"].shape
    assert np.array_equal(te.eval["precision"], out.precision.reshape(shp))
    assert np.array_equal(te.eval["recall"], out.recall.reshape(te.eval["recall"].shape))


def test_tao_max_dets(ref, files):
    te, ev, out = _tao_pair(ref, files, lambda p: None, max_dets=3)
    assert np.array_equal(te.eval["precision"], out.precision.reshape(te.eval["precision"].shape))


def _subset_imgs(p):
    p.img_ids = p.img_ids[::2]


def _vis_ranges(p):
    p.visibility_rng = [[0, 1.0], [0, 0.3], [0.3, 0.6], [0.6, 1.0], [0, 0.5], [0, 1.0]]


@pytest.mark.parametrize("mutate", [_subset_imgs, _subset_cats, _thresholds, _vis_ranges],
                         ids=lambda f: f.__name__.strip("_"))
def test_lvis_params(ref, files, mutate):
    le, ev, out = _lvis_pair(ref, files, mutate)
    assert np.array_equal(le.eval["precision"], out.precision)
    assert np.array_equal(le.eval["recall"], out.recall)


def test_lvis_max_dets(ref, files):
    le, ev, out = _lvis_pair(ref, files, lambda p: None, max_dets=2)
    assert np.array_equal(le.eval["precision"], out.precision)


def _odd_thresholds(p):
    p.rec_thrs = np.array([0.5, 0.1, 0.9, 0.1, 1.0, 0.0])      # unsorted, duplicated
    p.iou_thrs = np.array([0.75, 0.5, 0.5])


@pytest.mark.parametrize("pair", ["tao", "lvis"])
def test_recall_thresholds_out_of_order(ref, files, pair):
    """The reference's loop over the recall thresholds stops at the first unreachable one
    (eval.py:565-571); the evaluators feed the kernels ascending thresholds and restore the
    given order (evaluation/_common.py)."""
    from tao_amodal_b200.evaluation._common import ascending_rec_thrs, restore_rec_order
    fn = _tao_pair if pair == "tao" else _lvis_pair
    a, ev, _ = fn(ref, files, _odd_thresholds)
    asc, inv = ascending_rec_thrs(ev.params.rec_thrs)
    assert inv is not None
    out = run_hostsim(ev._plan, iou_thrs=ev.params.iou_thrs, rec_thrs=asc)
    prec = restore_rec_order(out.precision, out.recall, ev.params.rec_thrs, inv)
    assert np.array_equal(a.eval["precision"].reshape(prec.shape), prec)
    assert np.array_equal(a.eval["recall"].reshape(out.recall.shape), out.recall)
    assert (prec == 0.0).any() and (prec > 0.0).any()
