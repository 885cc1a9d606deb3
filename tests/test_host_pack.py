"""Host-side logic of the host-buffer call's transport forms (tao_amodal_b200/engine.py) that
needs no GPU: which detection boxes the track plan can take from the frame plan's upload
(shared_box_index), and the lossless-float decision."""
import dataclasses

import numpy as np
import pytest

from conftest import golden_inputs, load_golden
from plan_backends import plans_from_json
from tao_amodal_b200.engine import lossless_f32_boxes, shared_box_index


@pytest.mark.parametrize("case", ["tiny", "small", "edge_mix", "small_float", "small_sparse_ids"])
def test_shared_boxes_reproduce_the_track_plan(case):
    g = load_golden(case)
    tao, lvis = plans_from_json(*golden_inputs(g))
    got = shared_box_index(lvis, tao)
    assert got is not None
    idx, extra = got
    assert idx.dtype == np.int32 and idx.shape == (tao.dt_box.shape[0],)
    pool = np.concatenate([lvis.dt_box, extra])
    assert np.array_equal(pool[idx], tao.dt_box)
    # the extras are exactly the result rows the frame evaluator filtered out
    assert extra.shape[0] == np.setdiff1d(tao.dt_box_src, lvis.dt_box_src).size
    assert (idx[np.isin(tao.dt_box_src, lvis.dt_box_src)] < lvis.dt_box.shape[0]).all()


def test_sharing_is_refused_when_it_cannot_be_verified():
    g = load_golden("small")
    tao, lvis = plans_from_json(*golden_inputs(g))
    # plans that do not say where their boxes came from
    assert shared_box_index(dataclasses.replace(lvis, dt_box_src=None), tao) is None
    assert shared_box_index(lvis, dataclasses.replace(tao, dt_box_src=None)) is None
    # a source column that points at other boxes: the value check catches it
    wrong = dataclasses.replace(tao, dt_box_src=np.roll(tao.dt_box_src, 1))
    assert shared_box_index(lvis, wrong) is None
    # too many boxes the pool does not have
    assert shared_box_index(lvis, tao, max_extra=0.0) is None
    few = dataclasses.replace(lvis, dt_box=lvis.dt_box[:10].copy(), dt_box_src=lvis.dt_box_src[:10].copy())
    assert shared_box_index(few, tao) is None


def test_lossless_float_decision():
    g = load_golden("small")
    tao, lvis = plans_from_json(*golden_inputs(g))
    f = lossless_f32_boxes(lvis)
    assert f is not None and f[0].dtype == np.float32
    assert np.array_equal(f[0].astype(np.float64), lvis.dt_box)
    g = load_golden("small_float")
    tao, lvis = plans_from_json(*golden_inputs(g))
    assert lossless_f32_boxes(tao) is None and lossless_f32_boxes(lvis) is None
