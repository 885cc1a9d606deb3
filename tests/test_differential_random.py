"""Differential test on CPU: host prep + host build of the kernel arithmetic against the
pure-Python oracle on randomly drawn small datasets (seeds fixed).  Complements the goldens:
every run of the generator mixes different overlaps, score ties, sparse ids and empty cells."""
import copy

import numpy as np
import pytest

from oracle import lvis_frame, tao_track
from plan_backends import plans_from_json, run_hostsim
from tao_amodal_b200 import synth


from plan_backends import random_small_set


@pytest.mark.parametrize("seed", range(8))
def test_random_small_sets_match_oracle(seed):
    gt, res = random_small_set(seed)
    tao_plan, lvis_plan = plans_from_json(copy.deepcopy(gt), copy.deepcopy(res))
    got_t, got_l = run_hostsim(tao_plan), run_hostsim(lvis_plan)
    res2 = copy.deepcopy(res)
    tao_track.uniquify_track_ids(res2)
    ref_t = tao_track.evaluate_tao(copy.deepcopy(gt), res2, keep_cells=False)
    ref_l = lvis_frame.evaluate_lvis(copy.deepcopy(gt), copy.deepcopy(res), keep_cells=False)
    assert np.array_equal(ref_t["precision"], got_t.precision.reshape(ref_t["precision"].shape))
    assert np.array_equal(ref_t["recall"], got_t.recall.reshape(ref_t["recall"].shape))
    assert np.array_equal(ref_t["tp_cnt"], got_t.tp_cnt.reshape(ref_t["tp_cnt"].shape))
    assert np.array_equal(ref_t["num_gt"], got_t.num_gt.reshape(ref_t["num_gt"].shape))
    assert np.array_equal(ref_l["precision"], got_l.precision)
    assert np.array_equal(ref_l["fp_cnt"], got_l.fp_cnt)
    assert np.array_equal(ref_l["num_gt"], got_l.num_gt)
