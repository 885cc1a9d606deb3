"""TEST INFRASTRUCTURE — CPU oracle for the visibility-split frame-AP path.

Restates ``LVISEval.run()`` (tao_amodal/evaluation/lvis_amodal/eval.py:501-505)
for ``iou_type="bbox"`` over the reference's dict structures, with the
indexing it depends on (lvis_amodal/lvis.py, lvis_amodal/results.py).  The box
IoU the reference takes from third-party ``pycocotools.mask.iou`` (unpinned
pip dependency, absent from /root/reference/tao_amodal) is restated from the
in-tree copy of the same function (maskApi.c:109-120, see oracle/common.py).
Pinned against the unmodified reference by tests/test_oracle_golden.py.
Not product code.
"""
from __future__ import annotations

import copy
from collections import OrderedDict, defaultdict

import numpy as np

from .common import IOU_THRS, REC_THRS, frame_box_iou, greedy_match, masked_mean, pr_curve

VIS_RNG = [[0, 1.0], [0, 0.1], [0.1, 0.8], [0.8, 1.0], [0, 0.8], [0, 1.0]]
VIS_LBL = ["all", "highly-occluded", "partially-occluded", "highly-visible",
           "highly-and-partially-occluded", "out-of-frame"]
FREQ_LBL = ["r", "c", "f"]
MAX_DETS = 300


class FrameIndex:
    """lvis.py:37-61."""

    def __init__(self, ds):
        self.ds = ds
        self.img_anns = defaultdict(list)
        self.anns = {}
        for a in ds["annotations"]:
            self.img_anns[a["image_id"]].append(a)
            self.anns[a["id"]] = a
        self.imgs = {im["id"]: im for im in ds["images"]}
        self.cats = {c["id"]: c for c in ds["categories"]}

    def select_anns(self, img_ids, cat_ids):
        """lvis.py:63-97."""
        anns = []
        for i in img_ids:
            anns.extend(self.img_anns[i])
        cats = set(cat_ids)
        return [a for a in anns
                if a["category_id"] in cats and a["area"] > 0 and a["area"] < float("inf")]


def build_frame_results(gt_ds, results, max_dets=MAX_DETS):
    """lvis_amodal/results.py:9-71 for bbox results."""
    ds = copy.deepcopy(gt_ds)
    assert isinstance(results, list), "results is not a list."
    if max_dets >= 0:
        per_img = defaultdict(list)
        for r in results:
            per_img[r["image_id"]].append(r)
        for k, lst in per_img.items():
            if len(lst) > max_dets:
                per_img[k] = sorted(lst, key=lambda r: r["score"], reverse=True)[:max_dets]
        results = [r for lst in per_img.values() for r in lst]
    for n, r in enumerate(results):
        x1, y1, w, h = r["bbox"]
        r["area"] = w * h
        r["id"] = n + 1
    ds["annotations"] = results
    idx = FrameIndex(ds)
    got = set(r["image_id"] for r in results)
    assert got == (got & set(idx.imgs.keys())), "Results do not correspond to current LVIS set."
    return idx


def evaluate_lvis(gt_ds, results, keep_cells=True):
    """Whole LVISEval.run() (bbox).  Returns ious {(img,cat)}, cells
    {(ci, ri, ii)}, precision [T,R,C,6], recall [T,C,6], results."""
    gidx = FrameIndex(gt_ds)
    didx = build_frame_results(gt_ds, results)
    img_ids = list(np.unique(sorted(gidx.imgs.keys())))
    cat_ids = sorted(gidx.cats.keys())

    # ---- _prepare, eval.py:59-105
    gts = gidx.select_anns(img_ids, cat_ids)
    dts = didx.select_anns(img_ids, cat_ids)
    g_cell = defaultdict(list)
    d_cell = defaultdict(list)
    present = defaultdict(set)
    for g in gts:
        g.setdefault("ignore", 0)
        g_cell[g["image_id"], g["category_id"]].append(g)
        present[g["image_id"]].add(g["category_id"])
    neg = {i: gidx.imgs[i]["neg_category_ids"] for i in img_ids}
    nel = {i: gidx.imgs[i]["not_exhaustive_category_ids"] for i in img_ids}
    for d in dts:
        i, c = d["image_id"], d["category_id"]
        if c not in neg[i] and c not in present[i]:
            continue
        d_cell[i, c].append(d)
    freq_groups = [[] for _ in FREQ_LBL]
    for k, c in enumerate(cat_ids):
        freq_groups[FREQ_LBL.index(gidx.cats[c]["frequency"])].append(k)

    # ---- compute_iou, eval.py:168-192 (non-empty cells only)
    nonempty = sorted(set(g_cell.keys()) | set(d_cell.keys()))
    ious = {}
    n_pairs = 0
    for (i, c) in nonempty:
        gt, dt = g_cell.get((i, c), []), d_cell.get((i, c), [])
        order = np.argsort([-d["score"] for d in dt], kind="mergesort")
        dt = [dt[k] for k in order]
        if len(dt) == 0 or len(gt) == 0:
            ious[i, c] = []
            continue
        m = np.zeros((len(dt), len(gt)))
        for a, d in enumerate(dt):
            for b, g in enumerate(gt):
                m[a, b] = frame_box_iou([float(x) for x in d["bbox"]],
                                        [float(x) for x in g["bbox"]])
        n_pairs += m.size
        ious[i, c] = m

    # ---- evaluate_img, eval.py:194-303
    ipos = {i: k for k, i in enumerate(img_ids)}
    cpos = {c: k for k, c in enumerate(cat_ids)}
    cells = {}
    for (i, c) in nonempty:
        gt0, dt0 = g_cell.get((i, c), []), d_cell.get((i, c), [])
        for ri, rng in enumerate(VIS_RNG):
            oof = ri == len(VIS_RNG) - 1
            for g in gt0:
                if oof:
                    bad = g["ignore"] or (not g["out_of_frame"])
                else:
                    bad = g["ignore"] or (g["visibility"] < rng[0] or g["visibility"] > rng[1])
                g["_ignore"] = 1 if bad else 0
            gsel = np.argsort([g["_ignore"] for g in gt0], kind="mergesort")
            gt = [gt0[k] for k in gsel]
            dsel = np.argsort([-d["score"] for d in dt0], kind="mergesort")
            dt = [dt0[k] for k in dsel]
            m = ious[i, c][:, gsel] if len(ious[i, c]) > 0 else ious[i, c]
            gflag = np.array([g["_ignore"] for g in gt])
            dmask = [d["area"] < 0 or d["area"] > 1e5 ** 2
                     or d["category_id"] in nel[d["image_id"]] for d in dt]
            if len(m) == 0:
                m = np.zeros((len(dt), len(gt)))
                if len(gt) and len(dt):
                    raise AssertionError("unreachable")
            dt_m, gt_m, dt_ig = greedy_match(
                m, gflag, [g["id"] for g in gt], [d["id"] for d in dt], dmask, IOU_THRS, 0)
            cells[cpos[c], ri, ipos[i]] = {
                "image_id": i, "category_id": c,
                "dt_ids": [d["id"] for d in dt], "gt_ids": [g["id"] for g in gt],
                "dt_matches": dt_m, "gt_matches": gt_m,
                "dt_scores": [d["score"] for d in dt],
                "gt_ignore": gflag, "dt_ignore": dt_ig,
            }

    # ---- accumulate, eval.py:305-426
    T, R, C, NR = len(IOU_THRS), len(REC_THRS), len(cat_ids), len(VIS_RNG)
    precision = -np.ones((T, R, C, NR))
    recall = -np.ones((T, C, NR))
    tp_cnt = np.zeros((T, C, NR), dtype=np.int64)
    fp_cnt = np.zeros((T, C, NR), dtype=np.int64)
    num_gt = np.zeros((C, NR), dtype=np.int64)
    by_cat = defaultdict(list)
    for (ci, ri, ii) in cells:
        if ri == 0:
            by_cat[ci].append(ii)
    for ci, iis in by_cat.items():
        iis.sort()
        for ri in range(NR):
            E = [cells[ci, ri, ii] for ii in iis]
            sc = np.concatenate([e["dt_scores"] for e in E], axis=0)
            dm = np.concatenate([e["dt_matches"] for e in E], axis=1)
            di = np.concatenate([e["dt_ignore"] for e in E], axis=1)
            gi = np.concatenate([e["gt_ignore"] for e in E])
            num_gt[ci, ri] = np.count_nonzero(gi == 0)
            got = pr_curve(sc, dm, di, gi, 0, REC_THRS)
            if got is None:
                continue
            p, r, _, tps, fps = got
            precision[:, :, ci, ri] = p
            recall[:, ci, ri] = r
            tp_cnt[:, ci, ri] = tps.sum(1)
            fp_cnt[:, ci, ri] = fps.sum(1)

    out = {
        "img_ids": img_ids, "cat_ids": cat_ids, "ious": ious,
        "precision": precision, "recall": recall, "tp_cnt": tp_cnt, "fp_cnt": fp_cnt,
        "num_gt": num_gt, "freq_groups": freq_groups,
        "results": summarize_lvis(precision, recall, freq_groups),
        "box_pairs": n_pairs,
    }
    if keep_cells:
        out["cells"] = cells
    return out


def summarize_lvis(precision, recall, freq_groups):
    """lvis_amodal/eval.py:428-499; same keys, order and AR key collision."""
    def pick(kind, thr=None, vis="all", freq=None):
        ri = [k for k, l in enumerate(VIS_LBL) if l == vis]
        s = precision if kind == "ap" else recall
        if thr is not None:
            s = s[np.where(thr == IOU_THRS)[0]]
        if kind == "ap":
            s = s[:, :, freq_groups[freq], ri] if freq is not None else s[:, :, :, ri]
        else:
            s = s[:, :, ri]
        return masked_mean(s)

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
    res["AR@{}".format(MAX_DETS)] = pick("ar")
    for lbl in ["highly-occluded", "partially-occluded", "highly-visible",
                "highly-and-partially-occluded", "out-of-frame"]:
        res["AR{}@{}".format(lbl[0], MAX_DETS)] = pick("ar", vis=lbl)
    return res
