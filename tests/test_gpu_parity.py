"""Parity of the CUDA path (through the C ABI) with the unmodified reference's goldens and
with the CPU oracle.  Needs a B200; run with ``pytest -m gpu``."""
import numpy as np
import pytest

from conftest import golden_inputs
from plan_backends import compare_with_golden, plans_from_json, run_hostsim

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from tao_amodal_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def test_tao_golden_device_path(golden, eng):
    gt, res = golden_inputs(golden)
    plan, _ = plans_from_json(gt, res)
    out = eng.evaluate_device(eng.upload(plan), detail=True)
    off_grid = golden["_name"] == "small_float"
    compare_with_golden(golden, "tao_", plan, out, exact_iou=not off_grid, iou_atol=1e-12)


def test_lvis_golden_device_path(golden, eng):
    gt, res = golden_inputs(golden)
    _, plan = plans_from_json(gt, res)
    out = eng.evaluate_device(eng.upload(plan), detail=True)
    compare_with_golden(golden, "lvis_", plan, out, exact_iou=True)


def test_lvis_golden_unfused_kernels(golden, eng):
    """ta_box_iou + ta_match_greedy (the stand-alone entry points) give the same result as
    the fused frame kernel."""
    gt, res = golden_inputs(golden)
    _, plan = plans_from_json(gt, res)
    out = eng.evaluate_device(eng.upload(plan), detail=True, fused=False)
    compare_with_golden(golden, "lvis_", plan, out, exact_iou=True)


def test_frame_path_oversize_groups(eng):
    """Groups beyond the on-chip limits of ta_frame_eval (more than 32 GT boxes, or more than
    512 box pairs) are routed through the generic kernels; fused and unfused agree and both
    match the host arithmetic."""
    from tao_amodal_b200 import synth
    import numpy as np
    gtc, dtc = synth.generate_named("tiny", videos=2, frames=6, gt_tracks=48, pred_tracks=60,
                                    categories=3, max_present=1, seed=77)
    from tao_amodal_b200 import prep
    plan = prep.prepare_lvis(gtc, dtc)
    big = eng.big_list(plan)
    assert big.size > 0 and big.size < plan.n_groups
    a = eng.evaluate_device(eng.upload(plan), detail=True, fused=True)
    b = eng.evaluate_device(eng.upload(plan), detail=True, fused=False)
    ref = run_hostsim(plan)
    for o in (a, b):
        assert np.array_equal(o.iou, ref.iou)
        assert np.array_equal(o.dt_tpfp, ref.dt_tpfp)
        assert np.array_equal(o.dt_match_gt, ref.dt_match_gt)
        assert np.array_equal(o.gt_ignore, ref.gt_ignore)
        assert np.array_equal(o.num_gt, ref.num_gt)
        assert np.array_equal(o.precision, ref.precision)
        assert np.array_equal(o.recall, ref.recall)
    h = eng.evaluate_host(plan)
    assert np.array_equal(h.precision, ref.precision)
    assert np.array_equal(h.tp_cnt, ref.tp_cnt)


def test_host_buffer_call_matches_golden(golden, eng):
    """ta_eval_plan_host: the single C call the drop-in evaluators make."""
    gt, res = golden_inputs(golden)
    tao_plan, lvis_plan = plans_from_json(gt, res)
    o = eng.evaluate_host(tao_plan)
    assert np.array_equal(golden["tao_precision"], o.precision.reshape(golden["tao_precision"].shape))
    assert np.array_equal(golden["tao_recall"], o.recall.reshape(golden["tao_recall"].shape))
    assert np.array_equal(golden["tao_tp_cnt"], o.tp_cnt.reshape(golden["tao_tp_cnt"].shape))
    assert np.array_equal(golden["tao_fp_cnt"], o.fp_cnt.reshape(golden["tao_fp_cnt"].shape))
    assert o.h2d_bytes > 0 and o.d2h_bytes > 0
    o = eng.evaluate_host(lvis_plan)
    assert np.array_equal(golden["lvis_precision"], o.precision)
    assert np.array_equal(golden["lvis_recall"], o.recall)
    assert np.array_equal(golden["lvis_tp_cnt"], o.tp_cnt)
    assert np.array_equal(golden["lvis_fp_cnt"], o.fp_cnt)


@pytest.mark.parametrize("mode", ["3d_iou_seq", "avg_iou", "imagenetvid"])
def test_alt_iou_modes_match_host_arithmetic(mode, eng):
    """The pairwise kernel (alternative Params.iou_3d_type values, eval.py:51-117) against the
    same per-thread function compiled for the host."""
    from conftest import load_golden
    g = load_golden("small")
    gt, res = golden_inputs(g)
    plan, _ = plans_from_json(gt, res)
    out = eng.evaluate_device(eng.upload(plan), detail=True, iou_mode=mode)
    ref = run_hostsim(plan, mode)
    assert np.array_equal(out.iou, ref.iou)
    assert np.array_equal(out.dt_match_gt, ref.dt_match_gt)
    assert np.array_equal(out.precision, ref.precision)


def test_cfg2_sized_synthetic_matches_oracle(eng):
    """A BASELINE configs[1]-shaped set, reduced in videos so the pure-Python oracle finishes
    in seconds; integer decisions and AP compared bit for bit."""
    from oracle import lvis_frame, tao_track
    from tao_amodal_b200 import synth
    import copy
    gtc, dtc = synth.generate_named("cfg2", videos=2, frames=120, seed=4242)
    gt, res = gtc.to_dict(), dtc.to_list()
    tao_plan, lvis_plan = plans_from_json(copy.deepcopy(gt), copy.deepcopy(res))
    o_t = eng.evaluate_host(tao_plan)
    o_l = eng.evaluate_host(lvis_plan)
    res2 = copy.deepcopy(res)
    tao_track.uniquify_track_ids(res2)
    ref_t = tao_track.evaluate_tao(copy.deepcopy(gt), res2, keep_cells=False)
    ref_l = lvis_frame.evaluate_lvis(copy.deepcopy(gt), copy.deepcopy(res), keep_cells=False)
    assert np.array_equal(ref_t["precision"], o_t.precision.reshape(ref_t["precision"].shape))
    assert np.array_equal(ref_t["tp_cnt"], o_t.tp_cnt.reshape(ref_t["tp_cnt"].shape))
    assert np.array_equal(ref_t["fp_cnt"], o_t.fp_cnt.reshape(ref_t["fp_cnt"].shape))
    assert np.array_equal(ref_t["num_gt"], o_t.num_gt.reshape(ref_t["num_gt"].shape))
    assert np.array_equal(ref_l["precision"], o_l.precision)
    assert np.array_equal(ref_l["tp_cnt"], o_l.tp_cnt)
    assert np.array_equal(ref_l["fp_cnt"], o_l.fp_cnt)
    assert np.array_equal(ref_l["num_gt"], o_l.num_gt)


def test_lossless_f32_box_transport(eng):
    """Boxes that are exactly representable in float32 travel as float32 and are widened on the
    device: results identical to the fp64 transport; off-grid boxes fall back to fp64."""
    from conftest import load_golden
    from tao_amodal_b200.engine import lossless_f32_boxes
    g = load_golden("small")
    tao_plan, lvis_plan = plans_from_json(*golden_inputs(g))
    for plan in (tao_plan, lvis_plan):
        assert lossless_f32_boxes(plan) is not None
        a = eng.evaluate_host(plan, compress_boxes=True)
        b = eng.evaluate_host(plan, compress_boxes=False)
        assert a.h2d_bytes < b.h2d_bytes
        assert np.array_equal(a.precision, b.precision) and np.array_equal(a.tp_cnt, b.tp_cnt)
    g = load_golden("small_float")
    tao_plan, _ = plans_from_json(*golden_inputs(g))
    assert lossless_f32_boxes(tao_plan) is None
    o = eng.evaluate_host(tao_plan)
    assert np.array_equal(g["tao_precision"], o.precision.reshape(g["tao_precision"].shape))


def test_both_plans_in_one_call(eng):
    """ta_eval_plans_host: the track and the frame plan of one result file in one C call (one
    ta_ctx / stream each, one upload stream), against the reference goldens, in every transport
    form: shared box upload, separate boxes, uncompressed, page-locked memory."""
    from conftest import load_golden
    for case in ("edge_mix", "small", "small_ties", "small_float"):
        g = load_golden(case)
        tao_plan, lvis_plan = plans_from_json(*golden_inputs(g))
        variants = [dict(), dict(share_boxes=False), dict(compress=False), dict(pinned=True)]
        for kw in variants:
            pack = eng.pack_host([tao_plan, lvis_plan], **kw)
            if not kw or "pinned" in kw:
                assert pack.shared == {0: 1}, "the two plans hold the same result boxes"
                assert pack.structs[0].dt_box_idx and not pack.structs[0].dt_box
            else:
                assert not pack.shared
            for _ in range(2):
                o_t, o_l = eng.evaluate_pack(pack)
                assert np.array_equal(g["tao_precision"], o_t.precision.reshape(g["tao_precision"].shape))
                assert np.array_equal(g["tao_recall"], o_t.recall.reshape(g["tao_recall"].shape))
                assert np.array_equal(g["lvis_precision"], o_l.precision)
                assert np.array_equal(g["lvis_tp_cnt"], o_l.tp_cnt)
                assert np.array_equal(g["lvis_fp_cnt"], o_l.fp_cnt)
        # bytes over PCIe: sharing + compact tables must not cost more than separate uploads
        a = eng.evaluate_pack(eng.pack_host([tao_plan, lvis_plan]))
        b = eng.evaluate_pack(eng.pack_host([tao_plan, lvis_plan], compress=False))
        assert sum(o.h2d_bytes for o in a) < sum(o.h2d_bytes for o in b)
        # the order of the plans in the list is the order of the outputs
        o_l2, o_t2 = eng.evaluate_host_many([lvis_plan, tao_plan])
        assert np.array_equal(o_l2.precision, a[1].precision) and np.array_equal(o_t2.precision, a[0].precision)


def test_resident_plans_refreshed_from_a_pack(eng):
    """DevicePlan.reload_pack: the compact transport forms (float boxes, uint16 slots / group
    sizes, track boxes gathered from the frame plan's upload) restore wiped device buffers to
    the exact inputs — the route the multi-GPU end-to-end step takes."""
    from conftest import load_golden
    for case in ("edge_mix", "small_float"):
        g = load_golden(case)
        tao_plan, lvis_plan = plans_from_json(*golden_inputs(g))
        d_tao, d_lvis = eng.upload(tao_plan), eng.upload(lvis_plan)
        pack = eng.pack_host([tao_plan, lvis_plan], pinned=True)
        assert pack.shared == {0: 1}
        for rnd in range(2):
            for dv in (d_tao, d_lvis):
                for k in dv._input_keys:
                    if k not in ("iou_thrs", "rec_thrs"):
                        dv.t[k].zero_()
            n = d_lvis.reload_pack(pack, 1)
            n += d_tao.reload_pack(pack, 0, pool=d_lvis)
            assert n > 0
            o_t, o_l = eng.evaluate_device(d_tao), eng.evaluate_device(d_lvis)
            assert np.array_equal(g["tao_precision"], o_t.precision.reshape(g["tao_precision"].shape))
            assert np.array_equal(g["lvis_precision"], o_l.precision)
            assert np.array_equal(g["lvis_tp_cnt"], o_l.tp_cnt)
            assert np.array_equal(g["tao_recall"], o_t.recall.reshape(g["tao_recall"].shape))


def test_group_tables_as_counts(eng):
    """TA_PLAN_GRP_U16: the device prefix sum over uint16 group sizes reproduces the int64 offsets
    (sizes that cross several scan blocks; empty groups on either side)."""
    import ctypes as C
    import torch
    from tao_amodal_b200 import _lib
    from bench import make_workload
    gt, dt, tao_plan, lvis_plan = make_workload("cfg2", 0, 3)
    assert lvis_plan.n_groups > 3 * 2048
    pack = eng.pack_host([lvis_plan])
    assert pack.structs[0].flags & 4
    ref = eng.evaluate_device(eng.upload(lvis_plan))
    out = eng.evaluate_pack(pack)[0]
    for k in ("precision", "recall", "tp_cnt", "fp_cnt", "num_gt"):
        assert np.array_equal(getattr(out, k), getattr(ref, k)), k
    out2 = eng.evaluate_host(lvis_plan, compress_boxes=False)
    assert np.array_equal(out2.precision, ref.precision)


@pytest.mark.parametrize("seed", range(12))
def test_random_small_sets_gpu_vs_host_arithmetic(seed, eng):
    """Random small datasets (incl. overlapping GT -> several candidates per detection, the
    general matcher behind the flat frame kernel) through every CUDA route: host-buffer call,
    device stages with and without the per-cell outputs."""
    import copy
    from plan_backends import random_small_set
    gt, res = random_small_set(seed)
    tao_plan, lvis_plan = plans_from_json(copy.deepcopy(gt), copy.deepcopy(res))
    for plan in (tao_plan, lvis_plan):
        ref = run_hostsim(plan)
        h = eng.evaluate_host(plan)
        d = eng.evaluate_device(eng.upload(plan), detail=True)
        q = eng.evaluate_device(eng.upload(plan), detail=False)
        for o in (h, d, q):
            assert np.array_equal(o.precision, ref.precision)
            assert np.array_equal(o.recall, ref.recall)
            assert np.array_equal(o.tp_cnt, ref.tp_cnt)
            assert np.array_equal(o.fp_cnt, ref.fp_cnt)
            assert np.array_equal(o.num_gt, ref.num_gt)
        assert np.array_equal(d.dt_tpfp, ref.dt_tpfp)
        assert np.array_equal(d.dt_match_gt, ref.dt_match_gt)


def _gpu_pr(eng, c, iou_thrs, rec_thrs, impl):
    """ta_pr_accumulate on explicit arrays with TA_PR_IMPL = impl."""
    import ctypes as C
    import os
    import torch
    from tao_amodal_b200 import _lib
    dev = torch.device("cuda", eng.device)
    T, R, n_cat, n_cfg = len(iou_thrs), len(rec_thrs), c["n_cat"], c["n_cfg"]
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    cat_off, perm = up(c["cat_dt_off"], np.int64), up(c["acc_perm"], np.int32)
    tpfp, ngt = up(c["tpfp"].view(np.int32), np.int32), up(c["num_gt"], np.int32)
    rec = up(rec_thrs, np.float64)
    prec = torch.full((T, R, n_cat, n_cfg), 7.0, dtype=torch.float64, device=dev)
    rc = torch.full((T, n_cat, n_cfg), 7.0, dtype=torch.float64, device=dev)
    tp = torch.zeros((T, n_cat, n_cfg), dtype=torch.int64, device=dev)
    fp = torch.zeros_like(tp)
    p = lambda t: C.c_void_p(t.data_ptr())
    old = os.environ.get("TA_PR_IMPL")
    os.environ["TA_PR_IMPL"] = str(impl)
    try:
        st = C.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)
        _lib.check(eng.lib.ta_pr_accumulate(
            eng._ctx, st, n_cat, p(cat_off), p(perm), int(c["tpfp"].shape[0]), p(tpfp), None, p(ngt),
            T, n_cfg, R, p(rec), p(prec), p(rc), p(tp), p(fp)))
        torch.cuda.synchronize()
    finally:
        if old is None:
            os.environ.pop("TA_PR_IMPL")
        else:
            os.environ["TA_PR_IMPL"] = old
    return prec.cpu().numpy(), rc.cpu().numpy(), tp.cpu().numpy(), fp.cpu().numpy()


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("seed,kw", [(0, {}), (1, {}), (2, dict(n_cat=5, n_cfg=20, max_len=700)),
                                     (3, dict(n_cat=40, n_cfg=6, max_len=5000)),
                                     (100, dict(n_cat=4, n_cfg=3, tp_rate=0.0)),
                                     (101, dict(n_cat=4, n_cfg=3, tp_rate=1.0))])
def test_pr_accumulate_both_implementations(eng, seed, kw, impl):
    """Position-walk (TA_PR_IMPL=0) and bit-plane (TA_PR_IMPL=1, default) kernels of ta_pr_accumulate on
    random multi-chunk categories against the plain serial accumulation (tests/hostsim) and the
    oracle's accumulate cell."""
    from plan_backends import hostsim_pr
    from pr_cases import random_pr_case
    from tao_amodal_b200 import engine
    c = random_pr_case(seed, **kw)
    ref = hostsim_pr(c["n_cat"], c["cat_dt_off"], c["acc_perm"], c["tpfp"], c["num_gt"], c["n_cfg"])
    prec, rc, tp, fp = _gpu_pr(eng, c, engine.IOU_THRS, engine.REC_THRS, impl)
    assert np.array_equal(ref.precision, prec)
    assert np.array_equal(ref.recall, rc)
    assert np.array_equal(ref.tp_cnt, tp)
    assert np.array_equal(ref.fp_cnt, fp)
    if c["tpfp"].shape[0] * c["n_cfg"] <= 200000:
        # and against the oracle's accumulate cell (oracle.common.pr_curve), not only against the
        # emulation that shares the device header with the kernels
        from pr_cases import oracle_pr
        o_prec, o_rc, o_tp, o_fp = oracle_pr(c, engine.IOU_THRS, engine.REC_THRS)
        assert np.array_equal(o_prec, prec) and np.array_equal(o_rc, rc)
        assert np.array_equal(o_tp, tp) and np.array_equal(o_fp, fp)


@pytest.mark.parametrize("impl", [0, 1])
def test_goldens_with_each_pr_implementation(golden, eng, impl, monkeypatch):
    monkeypatch.setenv("TA_PR_IMPL", str(impl))
    gt, res = golden_inputs(golden)
    tao_plan, lvis_plan = plans_from_json(gt, res)
    off_grid = golden["_name"] == "small_float"
    compare_with_golden(golden, "tao_", tao_plan, eng.evaluate_device(eng.upload(tao_plan), detail=True),
                        exact_iou=not off_grid, iou_atol=1e-12)
    compare_with_golden(golden, "lvis_", lvis_plan, eng.evaluate_device(eng.upload(lvis_plan), detail=True))
