"""Live differential test against the UNMODIFIED reference (only where /root/reference exists,
i.e. in the build container; skipped elsewhere — the committed goldens cover those machines).
Random small datasets go through the reference's own TaoEval / LVISEval and through this
repo's host prep + host build of the kernel arithmetic; the reference's accessor methods are
compared with the drop-in classes'.  CPU only."""
import copy
import json

import numpy as np
import pytest

from oracle import golden_io, ref_shims
from plan_backends import plans_from_json, random_small_set, run_hostsim
from tao_amodal_b200 import materialize

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(),
                                reason="reference tree not present on this machine")


@pytest.fixture(scope="module")
def ref():
    return ref_shims.load_reference()


@pytest.mark.parametrize("seed", [20, 21, 22, 23])
def test_random_sets_against_the_reference(seed, ref, tmp_path):
    from oracle.make_golden import reference_make_track_ids_unique
    gt, res = random_small_set(seed)
    ap, rp = str(tmp_path / "gt.json"), str(tmp_path / "dt.json")
    json.dump(gt, open(ap, "w"))
    json.dump(res, open(rp, "w"))
    tao_plan, lvis_plan = plans_from_json(copy.deepcopy(gt), copy.deepcopy(res))
    got_t, got_l = run_hostsim(tao_plan), run_hostsim(lvis_plan)

    le = ref.LVISEval(ap, rp, "bbox")
    le.run()
    assert np.array_equal(le.eval["precision"], got_l.precision)
    assert np.array_equal(le.eval["recall"], got_l.recall)
    mine = materialize.summarize_lvis(got_l.precision, got_l.recall, le.params.iou_thrs,
                                      lvis_plan.freq_groups)
    assert list(mine.keys()) == list(le.results.keys())
    assert np.array_equal(golden_io.results_vector(mine), golden_io.results_vector(le.results))

    res2 = json.load(open(rp))
    reference_make_track_ids_unique()(res2)
    te = ref.TaoEval(ref.Tao(ap), res2)
    te.run()
    shape = te.eval["precision"].shape
    assert np.array_equal(te.eval["precision"], got_t.precision.reshape(shape))
    cells = materialize.cells_dict(tao_plan, 10, got_t)
    want = golden_io.flatten_cells(dict(te.eval_vids))
    for k, v in golden_io.flatten_cells(cells).items():
        assert np.array_equal(want[k], v), k


def test_accessors_match_the_reference(ref, tmp_path):
    from tao_amodal_b200.evaluation.lvis_amodal import LVIS
    from tao_amodal_b200.evaluation.tao_amodal import Tao, TaoResults
    gt, res = random_small_set(30)
    ap, rp = str(tmp_path / "gt.json"), str(tmp_path / "dt.json")
    json.dump(gt, open(ap, "w"))
    json.dump(res, open(rp, "w"))
    r, m = ref.Tao(ap), Tao(ap)
    assert sorted(r.get_vid_ids()) == sorted(m.get_vid_ids())
    assert sorted(r.get_cat_ids()) == sorted(m.get_cat_ids())
    assert sorted(r.get_img_ids()) == sorted(m.get_img_ids())
    vids, cats = sorted(r.get_vid_ids()), sorted(r.get_cat_ids())
    assert r.get_ann_ids(vid_ids=vids, cat_ids=cats) == m.get_ann_ids(vid_ids=vids, cat_ids=cats)
    assert r.get_ann_ids(vid_ids=vids[:1]) == m.get_ann_ids(vid_ids=vids[:1])
    assert r.get_ann_ids(cat_ids=cats[:2], area_rng=[100, 5000]) == \
        m.get_ann_ids(cat_ids=cats[:2], area_rng=[100, 5000])
    ids = r.get_ann_ids(vid_ids=vids, cat_ids=cats)
    gr, gm = r.group_ann_tracks(r.load_anns(ids)), m.group_ann_tracks(m.load_anns(ids))
    assert [t["id"] for t in gr] == [t["id"] for t in gm]
    assert [t["area"] for t in gr] == [t["area"] for t in gm]
    assert [[a["id"] for a in t["annotations"]] for t in gr] == \
        [[a["id"] for a in t["annotations"]] for t in gm]
    assert r.get_track_ids(cat_ids=cats[:3]) == m.get_track_ids(cat_ids=cats[:3])
    rr, mr = ref.TaoResults(ref.Tao(ap), json.load(open(rp))), TaoResults(Tao(ap), json.load(open(rp)))
    assert [a["id"] for a in rr.dataset["annotations"]] == [a["id"] for a in mr.dataset["annotations"]]
    assert [a["score"] for a in rr.dataset["annotations"]] == [a["score"] for a in mr.dataset["annotations"]]
    assert sorted(rr.tracks.keys()) == sorted(mr.tracks.keys())
    lr, lm = ref.LVIS(ap), LVIS(ap)
    imgs = sorted(lr.get_img_ids())
    assert lr.get_ann_ids(img_ids=imgs[:5], cat_ids=cats) == lm.get_ann_ids(img_ids=imgs[:5], cat_ids=cats)
    assert lr.load_cats(cats[:2]) == lm.load_cats(cats[:2])


@pytest.mark.parametrize("seed,dt_mode", [(41, "box"), (42, "poly"), (43, "rle")])
def test_segm_random_sets_match_the_reference(seed, dt_mode, ref, tmp_path):
    """LVISEval(iou_type="segm") of the unmodified reference (masks through its own maskApi.c,
    oracle/_ref) against the plan + mask codec + per-thread kernel arithmetic on the host."""
    from oracle import cases, maskapi_ref
    if not maskapi_ref.available():
        pytest.skip("oracle/_ref not built")
    gt, res = cases.case_segm(dt_mode, seed)
    ap, rp = str(tmp_path / "gt.json"), str(tmp_path / "dt.json")
    json.dump(gt, open(ap, "w"))
    json.dump(res, open(rp, "w"))
    le = ref.LVISEval(ap, rp, "segm")
    le.run()
    from tao_amodal_b200.evaluation.lvis_amodal import LVIS, LVISEval
    ev = LVISEval(LVIS(copy.deepcopy(gt)), copy.deepcopy(res), "segm")
    ev._prepare()
    out = run_hostsim(ev._plan)
    assert np.array_equal(le.eval["precision"], out.precision)
    assert np.array_equal(le.eval["recall"], out.recall)
