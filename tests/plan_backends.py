"""Test helpers: run an EvalPlan through (a) the host simulation of the kernel arithmetic
(tests/hostsim, CPU, test artefact only) or (b) the CUDA library via the C ABI, and compare
the outcome with a golden file generated from the unmodified reference."""
import ctypes as C
import os
import subprocess

import numpy as np

from tao_amodal_b200 import engine, materialize, prep
from tao_amodal_b200.columnar import DtColumns, GtColumns
from oracle import golden_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HS_DIR = os.path.join(ROOT, "tests", "hostsim")
HS_SO = os.path.join(HS_DIR, "_hostsim.so")


def build_hostsim():
    src = os.path.join(HS_DIR, "hostsim.cpp")
    deps = [src, os.path.join(ROOT, "tao_amodal_b200", "csrc", "ta_device_fns.cuh"),
            os.path.join(ROOT, "include", "ta_eval.h")]
    if (not os.path.exists(HS_SO)
            or os.path.getmtime(HS_SO) < max(os.path.getmtime(d) for d in deps)):
        subprocess.check_call([
            "g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++",
            "-I" + os.path.join(ROOT, "include"),
            "-I" + os.path.join(ROOT, "tao_amodal_b200", "csrc"), src, "-o", HS_SO])
    return C.CDLL(HS_SO)


def _p(a):
    return None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)


def run_hostsim(plan, iou_mode="3d_iou", iou_thrs=engine.IOU_THRS, rec_thrs=engine.REC_THRS,
                pr_impl="serial"):
    hs = build_hostsim()
    n_thr, n_rec, n_cfg, n_cat = len(iou_thrs), len(rec_thrs), plan.n_cfg, len(plan.cat_ids)
    n_iou = int(plan.iou_off[-1])
    iou = np.zeros(max(n_iou, 1))
    I64, I32 = C.c_int64, C.c_int32
    if plan.kind == "tao":
        hs.hs_track_iou.argtypes = [C.c_int, I64] + [C.c_void_p] * 10
        hs.hs_track_iou(engine._lib.IOU_MODES[iou_mode], plan.n_groups, _p(plan.grp_dt_off),
                        _p(plan.grp_gt_off), _p(plan.dt_trk_box_off), _p(plan.dt_box),
                        _p(plan.dt_box_slot), _p(plan.gt_trk_box_off), _p(plan.gt_box),
                        _p(plan.gt_box_slot), _p(plan.iou_off), _p(iou))
    elif plan.masks is not None:
        (do, dc, dhw, dbb), (go, gc, ghw, gbb) = plan.masks["dt"], plan.masks["gt"]
        hs.hs_rle_iou.argtypes = [I64] + [C.c_void_p] * 12
        hs.hs_rle_iou(plan.n_groups, _p(plan.grp_dt_off), _p(plan.grp_gt_off), _p(do), _p(dc),
                      _p(dhw), _p(dbb), _p(go), _p(gc), _p(ghw), _p(gbb), _p(plan.iou_off), _p(iou))
    else:
        hs.hs_box_iou.argtypes = [I64] + [C.c_void_p] * 6
        hs.hs_box_iou(plan.n_groups, _p(plan.grp_dt_off), _p(plan.grp_gt_off), _p(plan.dt_box),
                      _p(plan.gt_box), _p(plan.iou_off), _p(iou))
    tpfp = np.zeros((plan.n_dt, n_cfg), dtype=np.uint32)
    num_gt = np.zeros((n_cat, n_cfg), dtype=np.int32)
    match_gt = np.full((n_cfg, n_thr, plan.n_dt), -1, dtype=np.int32)
    gt_ig = np.zeros((n_cfg, plan.n_gt), dtype=np.uint8)
    thr = np.ascontiguousarray(iou_thrs, dtype=np.float64)
    rec = np.ascontiguousarray(rec_thrs, dtype=np.float64)
    P = C.c_void_p
    hs.hs_match_greedy.argtypes = ([I64, P, P, P, P, P, I32, P, I32, P, I64, P, P, P, I64,
                                    P, P, P, P, I32, P, P, P, P])
    g_max, _ = engine.plan_limits(plan)
    hs.hs_match_greedy(plan.n_groups, _p(plan.grp_dt_off), _p(plan.grp_gt_off), _p(plan.grp_cat),
                       _p(plan.iou_off), _p(iou), n_thr, _p(thr), n_cfg, _p(plan.range_cfgs),
                       plan.n_dt, _p(plan.dt_attr_a), _p(plan.dt_attr_b), _p(plan.dt_flag),
                       plan.n_gt, _p(plan.gt_attr_a), _p(plan.gt_attr_b),
                       _p(plan.gt_hp), _p(plan.gt_flag), g_max,
                       _p(tpfp), _p(num_gt), _p(match_gt), _p(gt_ig))
    out = hostsim_pr(n_cat, plan.cat_dt_off, plan.acc_perm, tpfp, num_gt, n_cfg, iou_thrs, rec_thrs,
                     impl=pr_impl)
    out.iou, out.dt_match_gt, out.gt_ignore = iou[:n_iou], match_gt, gt_ig
    return out


def hostsim_pr(n_cat, cat_dt_off, acc_perm, tpfp, num_gt, n_cfg, iou_thrs=engine.IOU_THRS,
               rec_thrs=engine.REC_THRS, impl="serial"):
    """PR accumulation of the host simulation on explicit arrays (also used by the
    multi-process exchange test).  impl="serial": the plain per-cell loop; impl="bits_tile":
    the serial emulation of the bit-plane kernels as shipped (chunks, warp transpose, TP-only
    walk, cell-major answers); impl="bits": the same with answers in the precision layout."""
    hs = build_hostsim()
    I64, I32, P = C.c_int64, C.c_int32, C.c_void_p
    n_thr, n_rec = len(iou_thrs), len(rec_thrs)
    n_dt = int(tpfp.shape[0])
    rec = np.ascontiguousarray(rec_thrs, dtype=np.float64)
    tpfp = np.ascontiguousarray(tpfp, dtype=np.uint32)
    num_gt = np.ascontiguousarray(num_gt, dtype=np.int32)
    out = engine.EvalOutput(
        precision=np.empty((n_thr, n_rec, n_cat, n_cfg)), recall=np.empty((n_thr, n_cat, n_cfg)),
        tp_cnt=np.empty((n_thr, n_cat, n_cfg), dtype=np.int64),
        fp_cnt=np.empty((n_thr, n_cat, n_cfg), dtype=np.int64), num_gt=num_gt, dt_tpfp=tpfp)
    fn = {"serial": hs.hs_pr_accumulate, "bits": hs.hs_pr_accumulate_bits,
          "bits_tile": hs.hs_pr_accumulate_bits_tile,
          "bits_seg": hs.hs_pr_accumulate_bits_seg}[impl]
    fn.argtypes = [I32, P, P, I64, P, P, I32, I32, I32, P, P, P, P, P]
    fn(n_cat, _p(np.asarray(cat_dt_off, dtype=np.int64)),
       _p(np.asarray(acc_perm, dtype=np.int32)), n_dt, _p(tpfp),
       _p(num_gt), n_thr, n_cfg, n_rec, _p(rec), _p(out.precision),
       _p(out.recall), _p(out.tp_cnt), _p(out.fp_cnt))
    return out


def plans_from_json(gt_dict, res_list):
    """(tao_plan, lvis_plan) the way the CLI builds them (track ids uniquified for TAO only,
    tools/eval_on_tao_amodal.py:127-129)."""
    gt = GtColumns.from_dict(gt_dict)
    dt = DtColumns.from_list(res_list)
    lvis_plan = prep.prepare_lvis(gt, dt)
    dt2 = dt.copy()
    prep.make_track_ids_unique(dt2)
    tao_plan = prep.prepare_tao(gt, dt2)
    return tao_plan, lvis_plan


def compare_with_golden(g, prefix, plan, out, exact_iou=True, iou_atol=0.0):
    """Every quantity the golden holds for one evaluator: IoU matrices, per-cell integer
    decisions, precision/recall tensors, TP/FP totals, summary metrics."""
    n_thr = out.recall.shape[0]
    ious = materialize.iou_dict(plan, out.iou)
    flat = golden_io.flatten_ious(ious)
    assert np.array_equal(g[prefix + "iou_keys"], flat["iou_keys"])
    assert np.array_equal(g[prefix + "iou_shape"], flat["iou_shape"])
    if exact_iou:
        assert np.array_equal(g[prefix + "iou_vals"], flat["iou_vals"])
    else:
        np.testing.assert_allclose(flat["iou_vals"], g[prefix + "iou_vals"], rtol=0, atol=iou_atol)
    cells = materialize.cells_dict(plan, n_thr, out)
    fc = golden_io.flatten_cells(cells)
    for k, v in fc.items():
        assert np.array_equal(g[prefix + k], v), k
    if plan.kind == "tao":
        shape5 = out.precision.shape[:3] + (5, 4)
        prec = out.precision.reshape(shape5)
        rec = out.recall.reshape(out.recall.shape[:2] + (5, 4))
        tp = out.tp_cnt.reshape(rec.shape)
        fp = out.fp_cnt.reshape(rec.shape)
        res = materialize.summarize_tao(prec, rec, engine.IOU_THRS)
    else:
        prec, rec, tp, fp = out.precision, out.recall, out.tp_cnt, out.fp_cnt
        res = materialize.summarize_lvis(prec, rec, engine.IOU_THRS, plan.freq_groups)
    assert np.array_equal(g[prefix + "precision"], prec)
    assert np.array_equal(g[prefix + "recall"], rec)
    assert np.array_equal(g[prefix + "tp_cnt"], tp)
    assert np.array_equal(g[prefix + "fp_cnt"], fp)
    assert golden_io.results_keys(res) == [str(k) for k in g[prefix + "results_keys"]]
    assert np.array_equal(g[prefix + "results"], golden_io.results_vector(res))


def random_small_set(seed: int):
    """A small random dataset (reference JSON structures): shape, tie density, id sparsity and
    GT overlap vary with the seed; every 4th seed duplicates GT boxes across tracks so that
    detections have several candidate GTs (the general matcher route)."""
    from tao_amodal_b200 import synth
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    gtc, dtc = synth.generate_named(
        "tiny", seed=500 + seed, videos=int(rng.integers(2, 5)), frames=int(rng.integers(6, 30)),
        pred_tracks=int(rng.integers(4, 16)), gt_tracks=int(rng.integers(2, 7)),
        categories=int(rng.integers(4, 14)), max_present=int(rng.integers(1, 4)),
        score_quantum=[0.0, 0.1, 0.25][seed % 3], sparse_image_ids=bool(seed % 2),
        keep_prob=float(rng.uniform(0.5, 1.0)))
    gt, res = gtc.to_dict(), dtc.to_list()
    if seed % 4 == 3:
        by_trk = {}
        for a in gt["annotations"]:
            by_trk.setdefault(a["track_id"], []).append(a)
        tids = sorted(by_trk)
        for t_src, t_dst in zip(tids[::2], tids[1::2]):
            src = {a["image_id"]: a for a in by_trk[t_src]}
            for a in by_trk[t_dst]:
                if a["image_id"] in src:
                    a["bbox"] = list(src[a["image_id"]]["bbox"])
                    a["area"] = a["bbox"][2] * a["bbox"][3]
        cat_of = {t["id"]: t["category_id"] for t in gt["tracks"]}
        for t_src, t_dst in zip(tids[::2], tids[1::2]):     # same category, or they never meet
            for t in gt["tracks"]:
                if t["id"] == t_dst and t["video_id"] == next(
                        x["video_id"] for x in gt["tracks"] if x["id"] == t_src):
                    t["category_id"] = cat_of[t_src]
                    for a in by_trk[t_dst]:
                        a["category_id"] = cat_of[t_src]
    return gt, res
