"""TEST INFRASTRUCTURE — CPU oracle for the TAO track-AP path.  Not product code.

Restates, in plain Python over the reference's own dict structures, what
``TaoEval.run()`` computes (tao_amodal/evaluation/tao_amodal/eval.py:662-666)
including the dataset/result indexing it depends on (tao.py, results.py) and
``make_track_ids_unique`` of the CLI (tools/eval_on_tao_amodal.py:44-66).
Pinned against the unmodified reference by tests/test_oracle_golden.py (see
oracle/common.py header).  Only tests/, smoke() and bench.py's CPU-baseline
legs may import this.
"""
from __future__ import annotations

import copy
from collections import OrderedDict, defaultdict

import numpy as np

from .common import (IOU_THRS, REC_THRS, box_inter_union, greedy_match, masked_mean,
                     pr_curve)

AREA_RNG = [[0 ** 2, 1e5 ** 2], [0 ** 2, 32 ** 2], [32 ** 2, 96 ** 2],
            [96 ** 2, 1e5 ** 2], [0 ** 2, 1e5 ** 2]]
AREA_LBL = ["all", "small", "medium", "large", "highly-and-partially-occluded"]
TIME_RNG = [[0, 1e5], [0, 3], [3, 10], [10, 1e5]]
TIME_LBL = ["all", "short", "medium", "long"]
MAX_DETS = 300


# --------------------------------------------------------------------------- CLI helper
def uniquify_track_ids(results):
    """tools/eval_on_tao_amodal.py:44-66."""
    first_video = {}
    clash = set()
    top = 0
    for r in results:
        t = r["track_id"]
        first_video.setdefault(t, r["video_id"])
        if r["video_id"] != first_video[t]:
            clash.add(t)
        top = max(top, t)
    if clash:
        fresh = {}
        nxt = top + 1
        for r in results:
            t, v = r["track_id"], r["video_id"]
            if t in clash:
                if (t, v) not in fresh:
                    fresh[t, v] = nxt
                    nxt += 1
                r["track_id"] = fresh[t, v]
    return len(clash)


# --------------------------------------------------------------------------- indexing
def merge_map_of(ds):
    """tao.py:97-106."""
    mm = {}
    for c in ds["categories"]:
        for m in c.get("merged", []):
            mm[m["id"]] = c["id"]
    return mm


class Index:
    """The dict indices of tao.py:108-160 that the evaluation reads."""

    def __init__(self, ds):
        self.ds = ds
        mm = merge_map_of(ds)
        for x in ds["annotations"] + ds["tracks"]:
            if x["category_id"] in mm:
                x["category_id"] = mm[x["category_id"]]
        self.vids = {v["id"]: v for v in ds["videos"]}
        self.tracks = {t["id"]: t for t in ds["tracks"]}
        self.cats = {c["id"]: c for c in ds["categories"]}
        self.imgs = {}
        self.vid_imgs = defaultdict(list)
        for im in ds["images"]:
            self.imgs[im["id"]] = im
            self.vid_imgs[im["video_id"]].append(im)
        self.anns = {}
        self.img_anns = defaultdict(list)
        self.track_anns = defaultdict(list)
        for a in ds["annotations"]:
            a["bbox"] = [float(x) for x in a["bbox"]]                       # tao.py:142
            assert a["category_id"] == self.tracks[a["track_id"]]["category_id"]  # :148-149
            self.track_anns[a["track_id"]].append(a)
            self.img_anns[a["image_id"]].append(a)
            self.anns[a["id"]] = a

    def select_anns(self, vid_ids, cat_ids):
        """tao.py:203-254 with vid_ids and cat_ids given, img_ids/area_rng None."""
        vimgs = []
        for v in vid_ids:
            vimgs.extend(im["id"] for im in self.vid_imgs[v])
        img_order = list(set(vimgs) & set(vimgs))      # CPython set order, tao.py:230
        anns = []
        for i in img_order:
            anns.extend(self.img_anns[i])
        cats = set(cat_ids)
        return [a for a in anns
                if a["category_id"] in cats and a["area"] > 0 and a["area"] < float("inf")]

    def group_tracks(self, anns):
        """tao.py:172-188."""
        out = {}
        for a in anns:
            t = a["track_id"]
            if t not in out:
                out[t] = dict(self.tracks[t])
                out[t]["annotations"] = []
            out[t]["annotations"].append(a)
        for tr in out.values():
            tr["annotations"] = sorted(
                tr["annotations"], key=lambda a: self.imgs[a["image_id"]]["frame_index"])
            tr["area"] = sum(a["area"] for a in tr["annotations"]) / len(tr["annotations"])
        return list(out.values())


def build_results_index(gt_ds, results, max_dets=MAX_DETS):
    """results.py:12-109 (TaoResults.__init__) for a list of result dicts."""
    ds = copy.deepcopy(gt_ds)
    mm = merge_map_of(ds)
    for r in results:
        if r["category_id"] in mm:
            r["category_id"] = mm[r["category_id"]]
    assert isinstance(results, list), "results is not a list."
    seen = {}
    for r in results:                                                     # :111-119
        t = r["track_id"]
        seen.setdefault(t, r["video_id"])
        assert r["video_id"] == seen[t], "Track id %s appears in more than one video" % t
    if max_dets >= 0:                                                     # :121-132
        per_img = defaultdict(list)
        for r in results:
            per_img[r["image_id"]].append(r)
        for k, lst in per_img.items():
            if len(lst) > max_dets:
                per_img[k] = sorted(lst, key=lambda r: r["score"], reverse=True)[:max_dets]
        results = [r for lst in per_img.values() for r in lst]
    tracks = {}
    if "bbox" in results[0]:
        for n, r in enumerate(results):
            x1, y1, w, h = r["bbox"]
            t = r["track_id"]
            if t not in tracks:
                tracks[t] = {"id": t, "video_id": r["video_id"],
                             "category_id": r["category_id"]}
            assert tracks[t]["category_id"] == r["category_id"]
            r["area"] = w * h
            r["id"] = n + 1
    ds["annotations"] = results
    ds["tracks"] = list(tracks.values())
    idx = Index(ds)
    for t, lst in idx.track_anns.items():                                  # :88-98
        sc = [float(a["score"]) for a in lst]
        uniq = set(sc)
        if len(uniq) > 1:
            avg = np.mean(sc)
            idx.tracks[t]["score"] = avg
            for a in lst:
                a["score"] = avg
        elif len(uniq) == 1:
            idx.tracks[t]["score"] = uniq.pop()
    got = set(r["image_id"] for r in results)
    assert got == (got & set(idx.imgs.keys())), "Results do not correspond to current Tao set."
    return idx


# --------------------------------------------------------------------------- IoU
def track_iou_3d(dt_track, gt_track):
    """eval.py:73-96."""
    i = 0
    u = 0
    for image in set(gt_track.keys()) | set(dt_track.keys()):
        g = gt_track.get(image, None)
        d = dt_track.get(image, None)
        if d and g:
            a, b = box_inter_union(d, g)
            i += a
            u += b
        elif not d and g:
            u += g[2] * g[3]
        elif d and not g:
            u += d[2] * d[3]
    assert i <= u
    return i / u if u > 0 else 0


def track_iou_avg(dt_track, gt_track):
    """eval.py:99-117."""
    vals = []
    for image in set(gt_track.keys()) | set(dt_track.keys()):
        g = gt_track.get(image, None)
        d = dt_track.get(image, None)
        if d and g:
            a, b = box_inter_union(d, g)
            vals.append(a / b if b > 0 else 0)
        elif (not d and g) or (d and not g):
            vals.append(0)
    return np.mean(vals)


def track_iou_imagenetvid(dt_track, gt_track, threshold=0.5):
    """eval.py:51-70."""
    hit = 0
    tot = 0
    for image in set(gt_track.keys()) | set(dt_track.keys()):
        g = gt_track.get(image, None)
        d = dt_track.get(image, None)
        if d and g:
            a, b = box_inter_union(d, g)
            if a > threshold * b:
                hit += 1
        if d or g:
            tot += 1
    return hit / tot


_PAIR = {"3d_iou": track_iou_3d, "avg_iou": track_iou_avg, "imagenetvid": track_iou_imagenetvid}


# --------------------------------------------------------------------------- evaluation
def evaluate_tao(gt_ds, results, iou_3d_type="3d_iou", keep_cells=True):
    """Whole TaoEval.run() on parsed JSON structures.

    ``gt_ds`` and ``results`` are consumed (mutated) like the reference does.
    Returns a dict: vid_ids, cat_ids, ious {(v,c): ndarray}, cells
    {(vi,ci,ai,ti): record}, precision [T,R,C,A,Tm], recall, results (the 19
    summary entries in reference order), counts of tp/fp per cell.
    """
    gidx = Index(gt_ds)
    didx = build_results_index(gt_ds, results)
    vid_ids = list(np.unique(sorted(gidx.vids.keys())))
    cat_ids = sorted(gidx.cats.keys())

    # ---- _prepare, eval.py:178-233
    gt_anns = gidx.select_anns(vid_ids, cat_ids)
    dt_anns = didx.select_anns(vid_ids, cat_ids)
    if len(gt_anns) == 0:
        raise ValueError("Found no groundtruth annotations for given params")
    if len(dt_anns) == 0:
        raise ValueError("Found no predicted annotations for given params")
    gts = gidx.group_tracks(gt_anns)
    dts = didx.group_tracks(dt_anns)
    g_cell = defaultdict(list)
    d_cell = defaultdict(list)
    present = defaultdict(set)
    for g in gts:
        g.setdefault("ignore", 0)
        g_cell[g["video_id"], g["category_id"]].append(g)
        present[g["video_id"]].add(g["category_id"])
    neg = {v: gidx.vids[v]["neg_category_ids"] for v in vid_ids}
    nel = {v: gidx.vids[v]["not_exhaustive_category_ids"] for v in vid_ids}
    for d in dts:
        v, c = d["video_id"], d["category_id"]
        if c not in neg[v] and c not in present[v]:
            continue
        d_cell[v, c].append(d)

    # ---- compute_iou, eval.py:306-335 (only non-empty cells are visited here; the
    # reference walks the full videos x categories grid and stores [] for the rest)
    pair = _PAIR[iou_3d_type]
    cells_nonempty = sorted(set(g_cell.keys()) | set(d_cell.keys()))
    ious = {}
    n_visits = 0
    for (v, c) in cells_nonempty:
        gt, dt = g_cell.get((v, c), []), d_cell.get((v, c), [])
        order = np.argsort([-d["score"] for d in dt], kind="mergesort")
        dt = [dt[i] for i in order]
        gmaps = [{a["image_id"]: a["bbox"] for a in g["annotations"]} for g in gt]
        dmaps = [{a["image_id"]: a["bbox"] for a in d["annotations"]} for d in dt]
        m = np.zeros([len(dt), len(gt)])
        for i in range(len(dt)):
            for j in range(len(gt)):
                m[i, j] = pair(dmaps[i], gmaps[j])
                n_visits += len(set(gmaps[j]) | set(dmaps[i]))
        ious[v, c] = m

    # ---- evaluate_vid, eval.py:337-457
    vpos = {v: i for i, v in enumerate(vid_ids)}
    cpos = {c: i for i, c in enumerate(cat_ids)}
    cells = {}
    for (v, c) in cells_nonempty:
        gt0, dt0 = g_cell.get((v, c), []), d_cell.get((v, c), [])
        for ai, arng in enumerate(AREA_RNG):
            occl = ai == len(AREA_RNG) - 1
            for ti, trng in enumerate(TIME_RNG):
                for g in gt0:
                    n = len(g["annotations"])
                    bad = (g["ignore"] or g["area"] < arng[0] or g["area"] > arng[1]
                           or n < trng[0] or n > trng[1])
                    if occl:
                        hp = sum(a["visibility"] < 0.8 for a in g["annotations"])
                        bad = bad or hp <= 5
                    g["_ignore"] = 1 if bad else 0
                gsel = np.argsort([g["_ignore"] for g in gt0], kind="mergesort")
                gt = [gt0[i] for i in gsel]
                dsel = np.argsort([-d["score"] for d in dt0], kind="mergesort")
                dt = [dt0[i] for i in dsel]
                m = ious[v, c][:, gsel] if len(ious[v, c]) > 0 else ious[v, c]
                gflag = np.array([g["_ignore"] for g in gt])
                dmask = [d["area"] < arng[0] or d["area"] > arng[1]
                         or len(d["annotations"]) < trng[0] or len(d["annotations"]) > trng[1]
                         or d["category_id"] in nel[d["video_id"]] for d in dt]
                dt_m, gt_m, dt_ig = greedy_match(
                    m, gflag, [g["id"] for g in gt], [d["id"] for d in dt], dmask,
                    IOU_THRS, -1)
                cells[vpos[v], cpos[c], ai, ti] = {
                    "video_id": v, "category_id": c,
                    "dt_ids": [d["id"] for d in dt], "gt_ids": [g["id"] for g in gt],
                    "dt_matches": dt_m, "gt_matches": gt_m,
                    "dt_scores": [d["score"] for d in dt],
                    "gt_ignore": gflag, "dt_ignore": dt_ig,
                }

    # ---- accumulate, eval.py:459-584
    T, R, C = len(IOU_THRS), len(REC_THRS), len(cat_ids)
    A, Tm, V = len(AREA_RNG), len(TIME_RNG), len(vid_ids)
    precision = -np.ones((T, R, C, A, Tm))
    recall = -np.ones((T, C, A, Tm))
    tp_cnt = np.zeros((T, C, A, Tm), dtype=np.int64)
    fp_cnt = np.zeros((T, C, A, Tm), dtype=np.int64)
    num_gt = np.zeros((C, A, Tm), dtype=np.int64)
    by_cat = defaultdict(list)
    for (vi, ci, ai, ti) in cells:
        if ai == 0 and ti == 0:
            by_cat[ci].append(vi)
    for ci, vis_ in by_cat.items():
        vis_.sort()
        for ai in range(A):
            for ti in range(Tm):
                E = [cells[vi, ci, ai, ti] for vi in vis_]
                sc = np.concatenate([e["dt_scores"] for e in E], axis=0)
                dm = np.concatenate([e["dt_matches"] for e in E], axis=1)
                di = np.concatenate([e["dt_ignore"] for e in E], axis=1)
                gi = np.concatenate([e["gt_ignore"] for e in E])
                num_gt[ci, ai, ti] = np.count_nonzero(gi == 0)
                got = pr_curve(sc, dm, di, gi, -1, REC_THRS)
                if got is None:
                    continue
                p, r, _, tps, fps = got
                precision[:, :, ci, ai, ti] = p
                recall[:, ci, ai, ti] = r
                tp_cnt[:, ci, ai, ti] = tps.sum(1)
                fp_cnt[:, ci, ai, ti] = fps.sum(1)

    out = {
        "vid_ids": vid_ids, "cat_ids": cat_ids, "ious": ious,
        "precision": precision, "recall": recall, "tp_cnt": tp_cnt, "fp_cnt": fp_cnt,
        "num_gt": num_gt, "results": summarize_tao(precision, recall),
        "box_pair_visits": n_visits,
    }
    if keep_cells:
        out["cells"] = cells
    return out


def summarize_tao(precision, recall):
    """eval.py:586-660; keys and order identical to TaoEval.results."""
    def pick(kind, thr=None, area="all", time="all"):
        ai = [i for i, l in enumerate(AREA_LBL) if l == area]
        ti = [i for i, l in enumerate(TIME_LBL) if l == time]
        s = precision if kind == "ap" else recall
        if thr is not None:
            s = s[np.where(thr == IOU_THRS)[0]]
        s = s[:, :, :, ai, ti] if kind == "ap" else s[:, :, ai, ti]
        return masked_mean(s)

    res = OrderedDict()
    hp = "highly-and-partially-occluded"
    res["AP"] = pick("ap")
    res["AP50"] = pick("ap", thr=0.50)
    res["AP75"] = pick("ap", thr=0.75)
    res["AP-HP"] = pick("ap", area=hp)
    res["AP50-HP"] = pick("ap", area=hp, thr=0.50)
    res["AP75-HP"] = pick("ap", area=hp, thr=0.75)
    for a in ["small", "medium", "large"]:
        res["AP", "area", a, MAX_DETS] = pick("ap", area=a)
    for t in ["short", "medium", "long"]:
        res["AP", "time", t, MAX_DETS] = pick("ap", time=t)
    res["AR@{}".format(MAX_DETS)] = pick("ar")
    for a in ["small", "medium", "large"]:
        res["AR", "area", a, MAX_DETS] = pick("ar", area=a)
    for t in ["short", "medium", "long"]:
        res["AR", "time", t, MAX_DETS] = pick("ar", time=t)
    return res
