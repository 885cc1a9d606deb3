"""Result checks at the sizes BASELINE.json names.

* configs[1] in full (cfg2: 10 videos x 300 frames, 50 predicted + 10 GT tracks per video, 100
  categories): the CUDA path against (a) the pure-Python oracle run live on the same seeded
  input and (b) the committed fixture `tests/golden/full_cfg2.npz`, which holds outputs of the
  UNMODIFIED reference on that input (oracle/make_fullsize_golden.py): counts, recall, the
  summary tables, and the SHA-256 of the precision tensors — bit-exact, all of them.
* configs[2]'s shape (cfg3: 1203 categories, 300-frame videos, 200 + 30 tracks per video) on 4
  whole videos evaluated jointly: same two comparators (`full_cfg3x4.npz`).
* configs[4] (cfg5: 5 000 videos, ~1 M predicted boxes, 1203 categories): the TaoEval half — the
  "full TrackAP table" — against the reference fixture `full_cfg5_tao.npz`; north_star asks for
  1e-6 on AP, the test asks for equality of every float.
"""
import copy
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def _plans(gt, dt):
    from tao_amodal_b200 import prep
    lvis = prep.prepare_lvis(gt, dt)
    d2 = dt.copy()
    prep.make_track_ids_unique(d2)
    return prep.prepare_tao(gt, d2), lvis


def _results(plan, out):
    from oracle import golden_io
    from tao_amodal_b200 import engine, materialize
    if plan.kind == "tao":
        prec = out.precision.reshape(out.precision.shape[:3] + (5, 4))
        rec = out.recall.reshape(out.recall.shape[:2] + (5, 4))
        res = materialize.summarize_tao(prec, rec, engine.IOU_THRS)
    else:
        res = materialize.summarize_lvis(out.precision, out.recall, engine.IOU_THRS, plan.freq_groups)
    return golden_io.results_keys(res), golden_io.results_vector(res)


def _check_fixture(fx, prefix, plan, out):
    shape = tuple(fx[prefix + "_recall"].shape)
    assert np.array_equal(fx[prefix + "_tp_cnt"], out.tp_cnt.reshape(shape)), prefix + " tp"
    assert np.array_equal(fx[prefix + "_fp_cnt"], out.fp_cnt.reshape(shape)), prefix + " fp"
    assert np.array_equal(fx[prefix + "_recall"], out.recall.reshape(shape)), prefix + " recall"
    assert tuple(fx[prefix + "_precision_shape"]) == tuple(
        out.precision.reshape(out.precision.shape[:3] + shape[2:]).shape)
    assert str(fx[prefix + "_precision_sha256"]) == _sha(out.precision), prefix + " precision"
    keys, vec = _results(plan, out)
    assert keys == [str(k) for k in fx[prefix + "_results_keys"]]
    assert np.array_equal(fx[prefix + "_results"], vec), prefix + " summary table"


def _check_oracle(gt, dt, tao_plan, lvis_plan, o_t, o_l):
    from oracle import lvis_frame, tao_track
    gd, res = gt.to_dict(), dt.to_list()
    res2 = copy.deepcopy(res)
    tao_track.uniquify_track_ids(res2)
    ref_t = tao_track.evaluate_tao(copy.deepcopy(gd), res2, keep_cells=False)
    ref_l = lvis_frame.evaluate_lvis(gd, res, keep_cells=False)
    for k in ("precision", "recall", "tp_cnt", "fp_cnt", "num_gt"):
        assert np.array_equal(ref_t[k], getattr(o_t, k).reshape(ref_t[k].shape)), "tao " + k
        assert np.array_equal(ref_l[k], getattr(o_l, k).reshape(ref_l[k].shape)), "lvis " + k


@pytest.fixture(scope="module")
def eng():
    import torch
    assert torch.cuda.is_available()
    from tao_amodal_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("name,workload,videos", [("cfg2", "cfg2", 0), ("cfg3x4", "cfg3", 4)])
def test_baseline_config_matches_reference_fixture_and_oracle(eng, name, workload, videos):
    from tao_amodal_b200 import synth
    gt, dt = synth.generate_named(workload, **({"videos": videos} if videos else {}))
    fx = np.load(os.path.join(GOLDEN, "full_%s.npz" % name))
    assert int(fx["n_pred_boxes"]) == dt.n() and int(fx["n_gt_boxes"]) == gt.n_anns()
    tao_plan, lvis_plan = _plans(gt, dt)
    # the single C call the drop-in evaluators make (host buffers), and the resident route
    o_t, o_l = eng.evaluate_host(tao_plan), eng.evaluate_host(lvis_plan)
    _check_fixture(fx, "tao", tao_plan, o_t)
    _check_fixture(fx, "lvis", lvis_plan, o_l)
    for plan, ref in ((tao_plan, o_t), (lvis_plan, o_l)):
        q = eng.evaluate_device(eng.upload(plan))
        for k in ("precision", "recall", "tp_cnt", "fp_cnt", "num_gt"):
            assert np.array_equal(getattr(q, k), getattr(ref, k)), k
    _check_oracle(gt, dt, tao_plan, lvis_plan, o_t, o_l)


def test_cfg5_track_ap_table_matches_reference_fixture(eng):
    """BASELINE configs[4] on one GPU (the 8-GPU run of the same set is `bench.py --workload
    cfg5 --gpus 8`, which compares its merged tensors with the 1-GPU ones and with this
    fixture)."""
    from tao_amodal_b200 import synth
    gt, dt = synth.generate_named("cfg5")
    fx = np.load(os.path.join(GOLDEN, "full_cfg5_tao.npz"))
    assert int(fx["n_pred_boxes"]) == dt.n() and int(fx["n_gt_boxes"]) == gt.n_anns()
    tao_plan, lvis_plan = _plans(gt, dt)
    o_t = eng.evaluate_host(tao_plan)
    _check_fixture(fx, "tao", tao_plan, o_t)
    keys, vec = _results(tao_plan, o_t)
    assert vec[keys.index("AP")] > 0.05, "the stress set must exercise real matches (AP > 0)"
    # frame half: no reference fixture can exist (the reference needs ~135 GB there); check it
    # against the resident route and the size-independent identities
    o_l = eng.evaluate_host(lvis_plan)
    q = eng.evaluate_device(eng.upload(lvis_plan))
    for k in ("precision", "recall", "tp_cnt", "fp_cnt", "num_gt"):
        assert np.array_equal(getattr(q, k), getattr(o_l, k)), k
    assert ((o_l.tp_cnt + o_l.fp_cnt) <= np.diff(lvis_plan.cat_dt_off)[None, :, None]).all()
    assert (o_l.tp_cnt <= o_l.num_gt[None]).all()
