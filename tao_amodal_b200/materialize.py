"""Reference-shaped views of device results.

The kernels return flat, group-contiguous arrays (engine.EvalOutput).  The reference keeps
per-cell dicts (``eval_vids`` tao_amodal/evaluation/tao_amodal/eval.py:445-457,
``eval_imgs`` lvis_amodal/eval.py:292-303), a dict of IoU matrices (``ious`` :264-268) and
``dt_pointers`` (:533-537).  Downstream analysis scripts read those, so they are rebuilt
here on demand from the flat arrays — pure re-indexing, no evaluation logic.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List

import numpy as np

from .prep import (EvalPlan, LVIS_VIS_LBL, MAX_DETS, TAO_AREA_LBL, TAO_TIME_LBL)


def iou_dict(plan: EvalPlan, iou_flat: np.ndarray) -> Dict[tuple, np.ndarray]:
    """``ious[(unit_id, cat_id)]`` for every non-empty group: ndarray [D,G] when both
    sides are non-empty, else ``[]`` (eval.py:309-310 / lvis eval.py:171-172)."""
    out = {}
    for g in range(plan.n_groups):
        D = int(plan.grp_dt_off[g + 1] - plan.grp_dt_off[g])
        G = int(plan.grp_gt_off[g + 1] - plan.grp_gt_off[g])
        key = (int(plan.unit_ids[plan.grp_unit[g]]), int(plan.cat_ids[plan.grp_cat[g]]))
        if D == 0 or G == 0:
            # TaoEval builds np.zeros([D, G]) when one side is empty (eval.py:327);
            # pycocotools returns [] (_mask.pyx:218-219)
            out[key] = [] if plan.kind == "lvis" else np.zeros((D, G))
            continue
        o = int(plan.iou_off[g])
        out[key] = iou_flat[o:o + D * G].reshape(D, G)
    return out


def cell_record(plan: EvalPlan, g: int, cfg: int, n_thr: int, dt_tpfp: np.ndarray,
                dt_match_gt: np.ndarray, gt_ignore: np.ndarray) -> dict:
    """One ``eval_vids`` / ``eval_imgs`` entry (group g, range cfg) in the reference's key
    set and array conventions (float id arrays, sentinel for unmatched)."""
    d0, d1 = int(plan.grp_dt_off[g]), int(plan.grp_dt_off[g + 1])
    g0, g1 = int(plan.grp_gt_off[g]), int(plan.grp_gt_off[g + 1])
    D, G = d1 - d0, g1 - g0
    sent = float(plan.sentinel)
    gig = gt_ignore[cfg, g0:g1]
    gsel = np.argsort(gig, kind="mergesort")                      # eval.py:371
    pos = np.empty(G, dtype=np.int64)
    pos[gsel] = np.arange(G)
    gt_ids = plan.gt_id[g0:g1][gsel]
    dt_ids = plan.dt_id[d0:d1]
    m = dt_match_gt[cfg, :, d0:d1].astype(np.int64)               # [T,D]
    dt_m = np.full((n_thr, D), sent)
    gt_m = np.full((n_thr, G), sent)
    hit = m >= 0
    if hit.any():
        dt_m[hit] = plan.gt_id[g0:g1][m[hit]].astype(np.float64)
        for t in range(n_thr):
            h = hit[t]
            if h.any():
                gt_m[t, pos[m[t, h]]] = dt_ids[h].astype(np.float64)   # later dt overwrites
    w = dt_tpfp[d0:d1, cfg].astype(np.uint32)
    t_idx = np.arange(n_thr, dtype=np.uint32)[:, None]
    counted = ((w[None, :] >> t_idx) | (w[None, :] >> (t_idx + np.uint32(16)))) & np.uint32(1)
    unit_key = "video_id" if plan.kind == "tao" else "image_id"
    return {
        unit_key: int(plan.unit_ids[plan.grp_unit[g]]),
        "category_id": int(plan.cat_ids[plan.grp_cat[g]]),
        "dt_ids": dt_ids.tolist(), "gt_ids": gt_ids.tolist(),
        "dt_matches": dt_m, "gt_matches": gt_m,
        "dt_scores": plan.dt_score[d0:d1].tolist(),
        "gt_ignore": gig[gsel].astype(np.int64), "dt_ignore": counted == 0,
    }


def cells_dict(plan: EvalPlan, n_thr: int, out) -> Dict[tuple, dict]:
    """All non-None cells keyed like the reference: (vid_idx, cat_idx, area_idx, time_idx) for
    the track path (eval.py:271-276), (cat_idx, range_idx, img_idx) for the frame path
    (lvis eval.py:140-145, flat index c*R*I + r*I + i)."""
    cells = {}
    n_time = len(TAO_TIME_LBL)
    for g in range(plan.n_groups):
        u, c = int(plan.grp_unit[g]), int(plan.grp_cat[g])
        for cfg in range(plan.n_cfg):
            rec = cell_record(plan, g, cfg, n_thr, out.dt_tpfp, out.dt_match_gt, out.gt_ignore)
            if plan.kind == "tao":
                cells[u, c, cfg // n_time, cfg % n_time] = rec
            else:
                cells[c, cfg, u] = rec
    return cells


def dt_pointers(plan: EvalPlan, n_thr: int, dt_tpfp: np.ndarray, num_gt: np.ndarray) -> dict:
    """``eval['dt_pointers'][cat_idx][...]`` = {dt_ids, tps, fps} in accumulate order
    (eval.py:486-537); only cells with non-ignored GT exist with content."""
    out: Dict[int, dict] = {}
    n_time = len(TAO_TIME_LBL)
    t_idx = np.arange(n_thr, dtype=np.uint32)[:, None]
    for c in range(len(plan.cat_ids)):
        p0, p1 = int(plan.cat_dt_off[c]), int(plan.cat_dt_off[c + 1])
        order = plan.acc_perm[p0:p1]
        per_cfg: List[dict] = []
        for cfg in range(plan.n_cfg):
            if num_gt[c, cfg] == 0 or int(plan.cat_grp_off[c + 1] - plan.cat_grp_off[c]) == 0:
                per_cfg.append({})
                continue
            w = dt_tpfp[order, cfg].astype(np.uint32)[None, :]
            per_cfg.append({"dt_ids": plan.dt_id[order],
                            "tps": ((w >> t_idx) & 1).astype(bool),
                            "fps": ((w >> (t_idx + np.uint32(16))) & 1).astype(bool)})
        if plan.kind == "tao":
            out[c] = {a: {t: per_cfg[a * n_time + t] for t in range(n_time)}
                      for a in range(plan.n_cfg // n_time)}
        else:
            out[c] = {r: per_cfg[r] for r in range(plan.n_cfg)}
    return out


# ---- summaries -----------------------------------------------------------------------------
def _masked_mean(s: np.ndarray):
    """eval.py:619-623 — numpy's pairwise mean over the > -1 entries, or -1."""
    sel = s[s > -1]
    if len(sel) == 0:
        return -1
    return np.mean(sel)


def summarize_tao(precision: np.ndarray, recall: np.ndarray, iou_thrs: np.ndarray,
                  area_lbl=TAO_AREA_LBL, time_lbl=TAO_TIME_LBL, max_dets: int = MAX_DETS):
    """TaoEval.summarize (eval.py:625-660) on precision [T,R,C,A,Tm], recall [T,C,A,Tm]."""
    def pick(kind, thr=None, area="all", time="all"):
        ai = [i for i, l in enumerate(area_lbl) if l == area]
        ti = [i for i, l in enumerate(time_lbl) if l == time]
        s = precision if kind == "ap" else recall
        if thr is not None:
            s = s[np.where(thr == iou_thrs)[0]]
        s = s[:, :, :, ai, ti] if kind == "ap" else s[:, :, ai, ti]
        return _masked_mean(s)

    hp = "highly-and-partially-occluded"
    res = OrderedDict()
    res["AP"] = pick("ap")
    res["AP50"] = pick("ap", thr=0.50)
    res["AP75"] = pick("ap", thr=0.75)
    res["AP-HP"] = pick("ap", area=hp)
    res["AP50-HP"] = pick("ap", area=hp, thr=0.50)
    res["AP75-HP"] = pick("ap", area=hp, thr=0.75)
    for a in ["small", "medium", "large"]:
        res[("AP", "area", a, max_dets)] = pick("ap", area=a)
    for t in ["short", "medium", "long"]:
        res[("AP", "time", t, max_dets)] = pick("ap", time=t)
    res["AR@{}".format(max_dets)] = pick("ar")
    for a in ["small", "medium", "large"]:
        res[("AR", "area", a, max_dets)] = pick("ar", area=a)
    for t in ["short", "medium", "long"]:
        res[("AR", "time", t, max_dets)] = pick("ar", time=t)
    return res


def summarize_lvis(precision: np.ndarray, recall: np.ndarray, iou_thrs: np.ndarray,
                   freq_groups, vis_lbl=LVIS_VIS_LBL, max_dets: int = MAX_DETS):
    """LVISEval.summarize (lvis_amodal/eval.py:459-499) on precision [T,R,C,6]; the AR keys
    are built from the first letter of the range label, so three ranges collide on
    ``ARh@300`` and the last one written wins (:497-499)."""
    def pick(kind, thr=None, vis="all", freq=None):
        ri = [i for i, l in enumerate(vis_lbl) if l == vis]
        s = precision if kind == "ap" else recall
        if thr is not None:
            s = s[np.where(thr == iou_thrs)[0]]
        if kind == "ap":
            s = s[:, :, freq_groups[freq], ri] if freq is not None else s[:, :, :, ri]
        else:
            s = s[:, :, ri]
        return _masked_mean(s)

    res = OrderedDict()
    res["AP"] = pick("ap")
    res["AP50"] = pick("ap", thr=0.50)
    res["AP75"] = pick("ap", thr=0.75)
    for tag, lbl in (("HO", "highly-occluded"), ("PO", "partially-occluded"),
                     ("HP", "highly-and-partially-occluded"), ("HV", "highly-visible"),
                     ("OOF", "out-of-frame")):
        res["AP-" + tag] = pick("ap", vis=lbl)
        res["AP50-" + tag] = pick("ap", thr=0.50, vis=lbl)
        res["AP75-" + tag] = pick("ap", thr=0.75, vis=lbl)
    res["APr"] = pick("ap", freq=0)
    res["APc"] = pick("ap", freq=1)
    res["APf"] = pick("ap", freq=2)
    res["AR@{}".format(max_dets)] = pick("ar")
    for lbl in ["highly-occluded", "partially-occluded", "highly-visible",
                "highly-and-partially-occluded", "out-of-frame"]:
        res["AR{}@{}".format(lbl[0], max_dets)] = pick("ar", vis=lbl)
    return res
