"""Host prep: columnar annotations -> the CSR "plan" the CUDA library consumes.

This is the vectorised (numpy) equivalent of everything the reference does
between ``json.load`` and the first IoU: ``TaoResults.__init__``
(tao_amodal/evaluation/tao_amodal/results.py:12-109), ``Tao.get_ann_ids`` /
``group_ann_tracks`` (tao.py:172-254), ``TaoEval._prepare`` (eval.py:178-233)
and, for the frame-AP path, ``LVISResults`` (lvis_amodal/results.py:9-84),
``LVIS.get_ann_ids`` (lvis.py:63-97) and ``LVISEval._prepare``
(lvis_amodal/eval.py:59-113).  All ordering rules that influence results are
reproduced exactly:

* image iteration order ``list(set(img_ids) & set(video_images))`` (tao.py:230)
  is evaluated literally with CPython sets;
* tracks appear in first-appearance order of that annotation stream
  (tao.py:174-180), annotations inside a track are stably sorted by
  ``frame_index`` (:182-184), ``track['area']`` is CPython's ``sum()`` (Neumaier
  compensated since 3.12) divided by the count (:186-187);
* max-300-detections-per-image with a stable descending score sort
  (results.py:121-132), ids assigned after limiting (:80-81), per-track mean
  score only when the track's box scores differ (:88-98);
* the strict ``0 < area < inf`` filter (tao.py:247-253) and the federated
  filter (eval.py:228-233);
* detections of a group in stable descending-score order (eval.py:313), groups
  of a category in video (image) order, which fixes the tie order of the
  stable merge sort in ``accumulate`` (eval.py:498-511).

Layout produced (all offsets are exclusive prefix sums):

    groups  sorted by (category index, video|image index)
    dt/gt entities (tracks for TAO, boxes for LVIS) stored group-contiguous
    TAO boxes stored track-contiguous, sorted by frame slot inside the track
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

from .columnar import DtColumns, GtColumns, Ragged

MAX_DETS = 300

# ---- range configurations (Params of both evaluators) ------------------------------------
# tao_amodal/evaluation/tao_amodal/eval.py:735-744
TAO_AREA_RNG = [[0 ** 2, 1e5 ** 2], [0 ** 2, 32 ** 2], [32 ** 2, 96 ** 2],
                [96 ** 2, 1e5 ** 2], [0 ** 2, 1e5 ** 2]]
TAO_AREA_LBL = ["all", "small", "medium", "large", "highly-and-partially-occluded"]
TAO_TIME_RNG = [[0, 1e5], [0, 3], [3, 10], [10, 1e5]]
TAO_TIME_LBL = ["all", "short", "medium", "long"]
# lvis_amodal/eval.py:567-575
LVIS_VIS_RNG = [[0, 1.0], [0, 0.1], [0.1, 0.8], [0.8, 1.0], [0, 0.8], [0, 1.0]]
LVIS_VIS_LBL = ["all", "highly-occluded", "partially-occluded", "highly-visible",
                "highly-and-partially-occluded", "out-of-frame"]

INF = float("inf")

# struct ta_range_cfg of include/ta_eval.h as a numpy record
RANGE_CFG_DTYPE = np.dtype([
    ("gt_a_lo", "<f8"), ("gt_a_hi", "<f8"), ("gt_b_lo", "<f8"), ("gt_b_hi", "<f8"),
    ("dt_a_lo", "<f8"), ("dt_a_hi", "<f8"), ("dt_b_lo", "<f8"), ("dt_b_hi", "<f8"),
    ("gt_hp_min", "<i4"), ("gt_need_oof", "<i4"),
], align=True)


def tao_range_cfgs(area_rng=TAO_AREA_RNG, time_rng=TAO_TIME_RNG) -> np.ndarray:
    """One record per (area a, duration t), flattened a-major like the reference's
    precision[..., a, t] axes.  The last area range carries the occlusion rule
    (eval.py:272, :357-368): ignore unless more than 5 frames have visibility < 0.8."""
    out = np.zeros(len(area_rng) * len(time_rng), dtype=RANGE_CFG_DTYPE)
    k = 0
    for a, ar in enumerate(area_rng):
        for t, tr in enumerate(time_rng):
            out[k] = (ar[0], ar[1], tr[0], tr[1], ar[0], ar[1], tr[0], tr[1],
                      6 if a == len(area_rng) - 1 else -(2 ** 31), 0)
            k += 1
    return out


def lvis_range_cfgs(vis_rng=LVIS_VIS_RNG) -> np.ndarray:
    """One record per visibility range; the last is the out-of-frame pseudo range
    (lvis_amodal/eval.py:140-145, :201-217).  Unmatched-detection ignore uses the fixed
    area window [0, 1e10] (:281-290)."""
    out = np.zeros(len(vis_rng), dtype=RANGE_CFG_DTYPE)
    for r, vr in enumerate(vis_rng):
        oof = r == len(vis_rng) - 1
        out[r] = (-INF if oof else vr[0], INF if oof else vr[1], -INF, INF,
                  0.0, 1e5 ** 2, -INF, INF, -(2 ** 31), 1 if oof else 0)
    return out


# ---- plan --------------------------------------------------------------------------------
@dataclass
class EvalPlan:
    kind: str                    # "tao" (track path) | "lvis" (frame path)
    cat_ids: np.ndarray          # int64 [C] sorted
    unit_ids: np.ndarray         # int64 [U] video ids (tao) / image ids (lvis), sorted
    n_groups: int
    grp_cat: np.ndarray          # int32 [n_groups] category index
    grp_unit: np.ndarray         # int32 [n_groups] video/image index
    grp_dt_off: np.ndarray       # int64 [n_groups+1]
    grp_gt_off: np.ndarray       # int64 [n_groups+1]
    iou_off: np.ndarray          # int64 [n_groups+1]  (sum D*G)
    cat_grp_off: np.ndarray      # int64 [C+1] group range of each category
    cat_dt_off: np.ndarray       # int64 [C+1] dt-entity range of each category
    acc_perm: np.ndarray         # int32 [n_dt]  position in score order -> dt entity index
    # dt entities
    dt_score: np.ndarray         # f64
    dt_attr_a: np.ndarray        # f64 area (tao: mean track area)
    dt_attr_b: np.ndarray        # f64 number of annotations (tao) / zeros (lvis)
    dt_flag: np.ndarray          # u8 bit0: category not exhaustively annotated in this unit,
                                 #    bit1: id > 0, i.e. the detection locks the GT it matches
    dt_id: np.ndarray            # int64 track id (tao) / annotation id (lvis)
    # gt entities
    gt_attr_a: np.ndarray        # f64 mean area (tao) / visibility (lvis)
    gt_attr_b: np.ndarray        # f64 number of annotations (tao) / zeros (lvis)
    gt_hp: np.ndarray            # int32 frames with visibility < 0.8 (tao) / zeros
    gt_flag: np.ndarray          # u8 bit0: ignore, bit1: out_of_frame, bit2: id == sentinel
    gt_id: np.ndarray            # int64
    # boxes
    dt_box: np.ndarray           # f64 [N,4]; lvis: one per entity
    gt_box: np.ndarray
    dt_trk_box_off: Optional[np.ndarray] = None   # int64 [n_dt+1] (tao)
    gt_trk_box_off: Optional[np.ndarray] = None
    dt_box_slot: Optional[np.ndarray] = None      # int32 frame slot inside the video (tao)
    gt_box_slot: Optional[np.ndarray] = None
    range_cfgs: Optional[np.ndarray] = None
    sentinel: int = -1           # "unmatched" id value: -1 (tao, eval.py:390-391) / 0 (lvis :239-240)
    freq_groups: Optional[list] = None            # lvis: category indices per r/c/f
    # lvis, iou_type="segm": run-length masks of the entities (mask.RlePool.export()):
    # {"dt"|"gt": (rle_off int64 [n+1], counts uint32, hw uint32 [n,2], bbox f64 [n,4])}
    masks: Optional[dict] = None
    # row of the result file (DtColumns) every detection box was taken from: lets the engine see
    # that two plans of one result file hold the same boxes (engine.shared_box_index)
    dt_box_src: Optional[np.ndarray] = None
    stats: Dict[str, float] = field(default_factory=dict)

    @property
    def n_dt(self) -> int:
        return int(self.dt_score.shape[0])

    @property
    def n_gt(self) -> int:
        return int(self.gt_id.shape[0])

    @property
    def n_cfg(self) -> int:
        return int(self.range_cfgs.shape[0])

    def box_pair_units(self) -> int:
        """The metric's unit: sum over track pairs of |F_d ∪ F_g| for the
        track path, sum of D*G for the frame path."""
        return int(self.stats["box_pairs"])


# ---- small helpers -----------------------------------------------------------------------
def _index_of(sorted_keys: np.ndarray, q: np.ndarray) -> np.ndarray:
    """Position of each q in sorted_keys, -1 when absent."""
    if sorted_keys.size == 0:
        return np.full(q.shape, -1, dtype=np.int64)
    if (q.size >= 4096 and sorted_keys.dtype.kind in "iu" and q.dtype.kind in "iu"):
        # dense integer keys (ids of images, videos, categories, tracks): a lookup table instead
        # of a binary search per query
        lo, hi = int(sorted_keys[0]), int(sorted_keys[-1])
        span = hi - lo + 1
        if span <= max(8 * sorted_keys.size, 1 << 16) and span <= (1 << 26):
            lut = np.full(span, -1, dtype=np.int64)
            # duplicates: the leftmost position, as searchsorted returns it
            lut[(sorted_keys[::-1] - lo)] = np.arange(sorted_keys.size - 1, -1, -1, dtype=np.int64)
            rel = q.astype(np.int64) - lo
            ok = (rel >= 0) & (rel < span)
            out = lut[np.where(ok, rel, 0)]
            out[~ok] = -1
            return out
    pos = np.searchsorted(sorted_keys, q)
    pos_c = np.minimum(pos, sorted_keys.size - 1)
    return np.where(sorted_keys[pos_c] == q, pos_c, -1).astype(np.int64)


def _apply_merge(cat: np.ndarray, merge_map: Dict[int, int]) -> np.ndarray:
    if not merge_map:
        return cat
    out = cat.copy()
    for src, dst in merge_map.items():
        out[cat == src] = dst
    return out


def _last_row_of(ids: np.ndarray):
    """dict(id -> row) semantics: later duplicates overwrite. Returns (sorted unique ids, row)."""
    order = np.argsort(ids, kind="stable")
    s = ids[order]
    last = np.ones(s.size, dtype=bool)
    last[:-1] = s[1:] != s[:-1]
    return s[last], order[last]


def _ragged_contains(r: Ragged, row: np.ndarray, val: np.ndarray) -> np.ndarray:
    """For each i: val[i] in list(row[i]) of the ragged column (rows may be -1 -> False)."""
    off, vals = r
    n_rows = off.size - 1
    if vals.size == 0 or row.size == 0:
        return np.zeros(row.shape, dtype=bool)
    lens = np.diff(off)
    row_of = np.repeat(np.arange(n_rows, dtype=np.int64), lens)
    vmax = int(max(vals.max(), val.max())) + 1
    vmin = int(min(vals.min(), val.min()))
    span = vmax - vmin + 1
    keys = np.sort(row_of * span + (vals - vmin))      # membership only: duplicates are harmless
    q = np.where(row >= 0, row, 0) * span + (val - vmin)
    hit = _index_of(keys, q) >= 0
    return hit & (row >= 0)


def neumaier_segment_sum(x: np.ndarray, off: np.ndarray) -> np.ndarray:
    """CPython >= 3.12 ``sum()`` of each float segment x[off[i]:off[i+1]] (builtin sum
    switched to Neumaier compensated summation; bltinmodule.c).  Segments must be
    non-empty.  Vectorised across segments, sequential inside a segment."""
    n = off.size - 1
    lens = np.diff(off)
    f = x[off[:-1]].astype(np.float64) + 0.0     # int 0 + x0
    c = np.zeros(n)
    maxlen = int(lens.max()) if n else 0
    alive = np.arange(n)
    for k in range(1, maxlen):
        alive = alive[lens[alive] > k]
        if alive.size == 0:
            break
        xv = x[off[alive] + k]
        fv = f[alive]
        t = fv + xv
        big = np.abs(fv) >= np.abs(xv)
        c[alive] += np.where(big, (fv - t) + xv, (xv - t) + fv)
        f[alive] = t
    add = (c != 0) & np.isfinite(c)
    f[add] = f[add] + c[add]
    return f


def limit_dets_per_image(image_id: np.ndarray, score: np.ndarray, max_dets: int) -> np.ndarray:
    """Row selection+order equal to results.py:121-132 / lvis results.py:73-84: rows grouped
    by image in first-appearance order; an image with more than max_dets rows keeps its top
    max_dets by a stable descending score sort (and in that order)."""
    n = image_id.size
    if max_dets < 0 or n == 0:
        return np.arange(n, dtype=np.int64)
    # the usual file: every image's rows are one contiguous run and no image is over-full —
    # the selection is then the identity (checked on the run heads only)
    head = np.concatenate([[0], np.nonzero(image_id[1:] != image_id[:-1])[0] + 1])
    if (np.diff(np.concatenate([head, [n]])).max() <= max_dets
            and np.unique(image_id[head]).size == head.size):
        return np.arange(n, dtype=np.int64)
    uniq, first, inv, cnt = np.unique(image_id, return_index=True, return_inverse=True,
                                      return_counts=True)
    first_rank = np.empty(uniq.size, dtype=np.int64)
    first_rank[np.argsort(first, kind="stable")] = np.arange(uniq.size)
    img_rank = first_rank[inv]
    within = np.arange(n, dtype=np.int64)
    over = cnt[inv] > max_dets
    keep = np.ones(n, dtype=bool)
    if over.any():
        rows = np.nonzero(over)[0]
        # stable descending sort by score inside each over-full image
        o = np.lexsort((rows, -score[rows], img_rank[rows]))
        rows_sorted = rows[o]
        grp = img_rank[rows_sorted]
        start = np.r_[0, np.nonzero(grp[1:] != grp[:-1])[0] + 1]
        pos = np.arange(rows_sorted.size) - np.repeat(start, np.diff(np.r_[start, grp.size]))
        within[rows_sorted] = pos
        keep[rows_sorted[pos >= max_dets]] = False
    sel = np.nonzero(keep)[0]
    order = np.lexsort((within[sel], img_rank[sel]))
    return sel[order]


def require_track_keys(dt: DtColumns) -> None:
    """The track path reads res['track_id'] / res['video_id'] of every result
    (tools/eval_on_tao_amodal.py:57, tao_amodal/results.py:71-76): a file that lacks them raises
    KeyError there, and so does this repo instead of scoring each video as one giant track."""
    if getattr(dt, "missing_track_id", 0):
        raise KeyError("track_id")
    if getattr(dt, "missing_video_id", 0):
        raise KeyError("video_id")


def make_track_ids_unique(dt: DtColumns) -> int:
    """tools/eval_on_tao_amodal.py:44-66 on columns (in place). Returns #clashing ids."""
    require_track_keys(dt)
    t, v = dt.track_id, dt.video_id
    if t.size == 0:
        return 0
    uniq, first = np.unique(t, return_index=True)
    first_vid = v[first][np.searchsorted(uniq, t)]
    clash_rows = v != first_vid
    clash_ids = np.unique(t[clash_rows])
    if clash_ids.size == 0:
        return 0
    max_id = max(0, int(t.max()))
    rows = np.nonzero(np.isin(t, clash_ids))[0]
    pair = np.stack([t[rows], v[rows]], axis=1)
    up, pfirst, pinv = np.unique(pair, axis=0, return_index=True, return_inverse=True)
    rank = np.empty(up.shape[0], dtype=np.int64)
    rank[np.argsort(pfirst, kind="stable")] = np.arange(up.shape[0])
    new = t.copy()
    new[rows] = max_id + 1 + rank[pinv.reshape(-1)]
    dt.track_id = new
    return int(clash_ids.size)


def cpython_set_order(ids: list) -> list:
    """tao.py:230 ``list(set(img_ids) & set(video_images))`` with img_ids == video_images."""
    return list(set(ids) & set(list(ids)))


def _layout(n_units, g_unit, g_cat, g_key, d_unit, d_cat, d_key, d_score, use_cats=True):
    """Sort entities into (category, unit) groups.  Returns group table + permutations.

    use_cats=False (Params.use_cats = 0, eval.py:293-303): one group per unit holding the
    entities of all categories, concatenated in category order then in-cell order."""
    if use_cats:
        gk = g_cat.astype(np.int64) * n_units + g_unit
        dk = d_cat.astype(np.int64) * n_units + d_unit
        g_perm = np.lexsort((g_key, gk))
        d_perm = np.lexsort((d_key, -d_score, dk))
    else:
        gk = g_unit.astype(np.int64)
        dk = d_unit.astype(np.int64)
        g_perm = np.lexsort((g_key, g_cat, gk))
        d_perm = np.lexsort((d_key, d_cat, -d_score, dk))
    keys = np.unique(np.concatenate([gk, dk]))
    g_cnt = np.bincount(_index_of(keys, gk), minlength=keys.size)
    d_cnt = np.bincount(_index_of(keys, dk), minlength=keys.size)
    grp_gt_off = np.zeros(keys.size + 1, dtype=np.int64)
    grp_dt_off = np.zeros(keys.size + 1, dtype=np.int64)
    np.cumsum(g_cnt, out=grp_gt_off[1:])
    np.cumsum(d_cnt, out=grp_dt_off[1:])
    iou_off = np.zeros(keys.size + 1, dtype=np.int64)
    np.cumsum(g_cnt.astype(np.int64) * d_cnt.astype(np.int64), out=iou_off[1:])
    grp_cat = (keys // n_units).astype(np.int32)
    grp_unit = (keys % n_units).astype(np.int32)
    return keys, grp_cat, grp_unit, grp_gt_off, grp_dt_off, iou_off, g_perm, d_perm


def _cat_offsets(n_cat, grp_cat, grp_dt_off):
    cat_grp_off = np.searchsorted(grp_cat, np.arange(n_cat + 1)).astype(np.int64)
    cat_dt_off = grp_dt_off[cat_grp_off].astype(np.int64)
    return cat_grp_off, cat_dt_off


def _acc_perm(n_cat, cat_dt_off, dt_score):
    """Stable descending-score order of each category's detections (eval.py:511)."""
    cat_of = np.repeat(np.arange(n_cat, dtype=np.int64), np.diff(cat_dt_off))
    return np.lexsort((-dt_score, cat_of)).astype(np.int32)


# ---- TAO track path ----------------------------------------------------------------------
def prepare_tao(gt: GtColumns, dt: DtColumns, max_dets: int = MAX_DETS,
                area_rng=TAO_AREA_RNG, time_rng=TAO_TIME_RNG,
                vid_ids=None, cat_ids=None, use_cats: bool = True,
                allow_empty: bool = False) -> EvalPlan:
    """Plan of the track path.  vid_ids / cat_ids default to everything in the annotation
    file (TaoEval.__init__, eval.py:169-170); subsets mirror assigning Params.vid_ids /
    Params.cat_ids before evaluate().  allow_empty: a video SHARD of a multi-GPU run may hold
    no ground truth / no predictions (or no videos); the reference's "found no ... annotations"
    errors (eval.py:188-192) then apply to the whole set, not to the shard (the caller checks
    the summed counts)."""
    cat_ids = np.unique(gt.cat_id if cat_ids is None else np.asarray(cat_ids, dtype=np.int64))
    vid_ids = np.unique(gt.vid_id if vid_ids is None else np.asarray(vid_ids, dtype=np.int64))
    n_cat, n_vid = cat_ids.size, vid_ids.size
    mm = gt.merge_map
    g_cat_raw = _apply_merge(gt.ann_category_id, mm)
    trk_cat = _apply_merge(gt.trk_category_id, mm)
    d_cat_raw = _apply_merge(dt.category_id, mm)

    # --- images: S order (tao.py:222-230), frame slots
    img_keys, img_row = _last_row_of(gt.img_id)
    img_vidx_all = _index_of(vid_ids, gt.img_video_id)
    listed = np.nonzero(img_vidx_all >= 0)[0]
    L_rows = listed[np.argsort(img_vidx_all[listed], kind="stable")]
    S = cpython_set_order(gt.img_id[L_rows].tolist())
    S = np.asarray(S, dtype=np.int64)
    n_img = S.size
    s_sort = np.argsort(S, kind="stable")
    S_sorted = S[s_sort]

    def s_rank(ids):
        p = _index_of(S_sorted, ids)
        return np.where(p >= 0, s_sort[np.maximum(p, 0)], -1)

    # attributes of each S image (dict semantics: last row with that id)
    S_row = img_row[_index_of(img_keys, S)]
    S_vidx = img_vidx_all[S_row]
    S_frame = gt.img_frame_index[S_row]
    slot_order = np.lexsort((np.arange(n_img), S_frame, S_vidx))
    S_slot = np.empty(n_img, dtype=np.int64)
    vstart = np.searchsorted(S_vidx[slot_order], np.arange(n_vid))
    S_slot[slot_order] = np.arange(n_img) - vstart[S_vidx[slot_order]]

    # --- ground truth
    trk_keys, trk_row = _last_row_of(gt.trk_id)
    a_trk = _index_of(trk_keys, gt.ann_track_id)
    if (a_trk < 0).any():
        raise KeyError(int(gt.ann_track_id[np.nonzero(a_trk < 0)[0][0]]))   # tao.py:148
    a_trk_row = trk_row[a_trk]
    if (g_cat_raw != trk_cat[a_trk_row]).any():
        raise AssertionError("annotation category differs from its track's category")  # :148-149
    g_rank = s_rank(gt.ann_image_id)
    g_valid = ((g_rank >= 0) & (_index_of(cat_ids, g_cat_raw) >= 0)
               & (gt.ann_area > 0) & (gt.ann_area < INF))
    g_rows = np.nonzero(g_valid)[0]
    if g_rows.size == 0 and not allow_empty:
        raise ValueError("Found no groundtruth annotations for given params")    # eval.py:188-190
    gt_ent = _tao_tracks(
        rows=g_rows, rank=g_rank[g_rows], trk_key=a_trk_row[g_rows],
        S_frame=S_frame, S_slot=S_slot, bbox=gt.ann_bbox, area=gt.ann_area,
        vis=gt.ann_visibility)
    g_trow = gt_ent["trk_key"]                       # row in gt track table
    g_unit = _index_of(vid_ids, gt.trk_video_id[g_trow])
    g_cidx = _index_of(cat_ids, trk_cat[g_trow])
    if np.isnan(gt_ent["vis_min"]).any():
        # evaluate_vid reads x['visibility'] for the occlusion range (eval.py:360)
        raise KeyError("visibility")

    # --- detections (results.py:12-109)
    require_track_keys(dt)
    if dt.n() == 0 and not allow_empty:
        raise IndexError("list index out of range")                              # results.py:63
    tu, tfirst = np.unique(dt.track_id, return_index=True)
    if dt.n() and (dt.video_id != dt.video_id[tfirst][np.searchsorted(tu, dt.track_id)]).any():
        bad = np.nonzero(dt.video_id != dt.video_id[tfirst][np.searchsorted(tu, dt.track_id)])[0][0]
        raise AssertionError("Track id %d appears in more than one video" % int(dt.track_id[bad]))
    sel = limit_dets_per_image(dt.image_id, dt.score, max_dets)
    d_img = dt.image_id[sel]
    d_trk = dt.track_id[sel]
    d_cat = d_cat_raw[sel]
    d_vid = dt.video_id[sel]
    d_box = dt.bbox[sel]
    d_score = dt.score[sel].astype(np.float64)
    d_area = d_box[:, 2] * d_box[:, 3]
    if (_index_of(img_keys, d_img) < 0).any():
        raise AssertionError("Results do not correspond to current Tao set.")    # results.py:105-109
    # track table: first appearance, category consistency (results.py:71-79)
    tu, tfirst, tinv = np.unique(d_trk, return_index=True, return_inverse=True)
    if (d_cat != d_cat[tfirst][tinv]).any():
        raise AssertionError("Annotations for a track have multiple categories")
    t_vid = d_vid[tfirst]
    t_cat = d_cat[tfirst]
    # track score (results.py:88-98): mean only when the track's scores differ
    o = np.argsort(tinv, kind="stable")
    seg = np.zeros(tu.size + 1, dtype=np.int64)
    np.cumsum(np.bincount(tinv, minlength=tu.size), out=seg[1:])
    sc_sorted = d_score[o]
    smin = np.minimum.reduceat(sc_sorted, seg[:-1]) if tu.size else np.zeros(0)
    smax = np.maximum.reduceat(sc_sorted, seg[:-1]) if tu.size else np.zeros(0)
    t_score = smin.copy()
    for k in np.nonzero(smin != smax)[0]:
        t_score[k] = np.mean(sc_sorted[seg[k]:seg[k + 1]].tolist())
    d_rank = s_rank(d_img)
    d_valid = ((d_rank >= 0) & (_index_of(cat_ids, d_cat) >= 0) & (d_area > 0) & (d_area < INF))
    d_rows = np.nonzero(d_valid)[0]
    if d_rows.size == 0 and not allow_empty:
        raise ValueError("Found no predicted annotations for given params")      # eval.py:191-192
    dt_ent = _tao_tracks(
        rows=d_rows, rank=d_rank[d_rows], trk_key=tinv[d_rows],
        S_frame=S_frame, S_slot=S_slot, bbox=d_box, area=d_area, vis=None)
    d_t = dt_ent["trk_key"]                          # index into tu
    d_unit = _index_of(vid_ids, t_vid[d_t])
    if (d_unit < 0).any():
        raise KeyError(int(t_vid[d_t][np.nonzero(d_unit < 0)[0][0]]))            # eval.py:230
    d_cidx = _index_of(cat_ids, t_cat[d_t])

    # --- federated filter (eval.py:214-233)
    if not gt.has_video_lists:
        raise KeyError("neg_category_ids")
    vid_keys, vid_row = _last_row_of(gt.vid_id)
    v_row_of_unit = vid_row[_index_of(vid_keys, vid_ids)]
    present = np.unique(g_unit.astype(np.int64) * n_cat + g_cidx)
    in_present = _index_of(present, d_unit.astype(np.int64) * n_cat + d_cidx) >= 0
    in_neg = _ragged_contains(gt.vid_neg, v_row_of_unit[d_unit], t_cat[d_t])
    keep = in_present | in_neg
    if not use_cats:
        keep = np.ones_like(keep)          # eval.py:230: the federated filter needs use_cats
    in_nel = _ragged_contains(gt.vid_nel, v_row_of_unit[d_unit], t_cat[d_t])

    kept = np.nonzero(keep)[0]
    (keys, grp_cat, grp_unit, grp_gt_off, grp_dt_off, iou_off, g_perm, d_perm) = _layout(
        n_vid, g_unit, g_cidx, gt_ent["first_key"],
        d_unit[kept], d_cidx[kept], dt_ent["first_key"][kept], t_score[d_t][kept], use_cats)
    d_sel = kept[d_perm]
    if not use_cats:
        cat_ids, n_cat = np.asarray([-1], dtype=np.int64), 1      # eval.py:259-260

    gb_off, gb, gslot, _ = _gather_track_boxes(gt_ent, g_perm)
    db_off, db, dslot, db_src = _gather_track_boxes(dt_ent, d_sel)
    cat_grp_off, cat_dt_off = _cat_offsets(n_cat, grp_cat, grp_dt_off)
    dt_score = np.ascontiguousarray(t_score[d_t][d_sel])

    plan = EvalPlan(
        kind="tao", cat_ids=cat_ids, unit_ids=vid_ids, n_groups=int(keys.size),
        grp_cat=grp_cat, grp_unit=grp_unit, grp_dt_off=grp_dt_off, grp_gt_off=grp_gt_off,
        iou_off=iou_off, cat_grp_off=cat_grp_off, cat_dt_off=cat_dt_off,
        acc_perm=_acc_perm(n_cat, cat_dt_off, dt_score),
        dt_score=dt_score,
        dt_attr_a=np.ascontiguousarray(dt_ent["area_mean"][d_sel]),
        dt_attr_b=np.ascontiguousarray(dt_ent["n_anns"][d_sel].astype(np.float64)),
        dt_flag=np.ascontiguousarray(in_nel[d_sel].astype(np.uint8)
                                     | ((tu[d_t][d_sel] > 0).astype(np.uint8) << 1)),
        dt_id=np.ascontiguousarray(tu[d_t][d_sel]),
        gt_attr_a=np.ascontiguousarray(gt_ent["area_mean"][g_perm]),
        gt_attr_b=np.ascontiguousarray(gt_ent["n_anns"][g_perm].astype(np.float64)),
        gt_hp=np.ascontiguousarray(gt_ent["n_hp"][g_perm].astype(np.int32)),
        gt_flag=np.ascontiguousarray((gt.trk_ignore[g_trow][g_perm].astype(np.uint8) & 1)
                                     | ((gt.trk_id[g_trow][g_perm] == -1).astype(np.uint8) << 2)),
        gt_id=np.ascontiguousarray(gt.trk_id[g_trow][g_perm]),
        dt_box=db, gt_box=gb, dt_trk_box_off=db_off, gt_trk_box_off=gb_off,
        dt_box_slot=dslot, gt_box_slot=gslot,
        range_cfgs=tao_range_cfgs(area_rng, time_rng), sentinel=-1,
        dt_box_src=np.ascontiguousarray(sel[db_src]),
    )
    plan.stats["n_dt_boxes"] = float(db.shape[0])
    plan.stats["n_gt_boxes"] = float(gb.shape[0])
    return plan


def _tao_tracks(rows, rank, trk_key, S_frame, S_slot, bbox, area, vis):
    """group_ann_tracks (tao.py:172-188) on the selected annotation rows.

    rows: annotation rows (dataset / result order index) that passed the filters
    rank: S rank of each row's image;  trk_key: integer key of each row's track.
    Returns per-track arrays in arbitrary (key-sorted) order plus ``first_key`` which orders
    tracks by first appearance in the annotation stream sorted by (S rank, row)."""
    stream = np.lexsort((rows, rank))                 # get_ann_ids order (tao.py:233-234)
    pos_in_stream = np.empty(rows.size, dtype=np.int64)
    pos_in_stream[stream] = np.arange(rows.size)
    keys, inv = np.unique(trk_key, return_inverse=True)
    n_trk = keys.size
    # inside a track: stable sort by frame_index over the stream order (tao.py:182-184)
    order = np.lexsort((pos_in_stream, S_frame[rank], inv))
    seg = np.zeros(n_trk + 1, dtype=np.int64)
    np.cumsum(np.bincount(inv, minlength=n_trk), out=seg[1:])
    # first appearance of every track in the stream (tracks are contiguous in `order`)
    first_key = (np.minimum.reduceat(pos_in_stream[order], seg[:-1]) if n_trk
                 else np.zeros(0, dtype=np.int64))
    r_sorted = rows[order]
    a_sorted = area[r_sorted].astype(np.float64)
    area_mean = neumaier_segment_sum(a_sorted, seg) / np.diff(seg)
    n_anns = np.diff(seg)
    if vis is not None:
        v_sorted = vis[r_sorted]
        n_hp = np.add.reduceat((v_sorted < 0.8).astype(np.int64), seg[:-1])
        vis_min = np.minimum.reduceat(np.where(np.isnan(v_sorted), np.nan, 0.0), seg[:-1])
    else:
        n_hp = np.zeros(n_trk, dtype=np.int64)
        vis_min = np.zeros(n_trk)
    # boxes keyed by image: a later annotation on the same image replaces the earlier one
    # (dict comprehension, eval.py:322-325)
    slot_sorted = S_slot[rank[order]]
    trk_sorted = inv[order]
    # one stable sort by (track, slot) — on a single integer key, nearly sorted already — serves
    # both steps: duplicates of an image need not be adjacent in `order` when two images share a
    # frame_index, and the kept boxes are wanted in (track, slot) order
    n_slot = int(slot_sorted.max()) + 1 if slot_sorted.size else 1
    o2 = np.argsort(trk_sorted.astype(np.int64) * n_slot + slot_sorted, kind="stable")
    ts, ss = trk_sorted[o2], slot_sorted[o2]
    last = np.ones(o2.size, dtype=bool)
    last[:-1] = (ts[1:] != ts[:-1]) | (ss[1:] != ss[:-1])
    kb = o2[last]                           # the last of every (track, slot) run, in that order
    box_seg = np.zeros(n_trk + 1, dtype=np.int64)
    np.cumsum(np.bincount(trk_sorted[kb], minlength=n_trk), out=box_seg[1:])
    return {
        "trk_key": keys, "first_key": first_key, "area_mean": area_mean, "n_anns": n_anns,
        "n_hp": n_hp, "vis_min": vis_min, "box_seg": box_seg,
        "box": bbox[r_sorted[kb]].astype(np.float64), "slot": slot_sorted[kb].astype(np.int32),
        "src": r_sorted[kb],
    }


def _gather_track_boxes(ent, perm):
    """Concatenate the box runs of tracks ``perm`` (in that order)."""
    seg = ent["box_seg"]
    lens = (seg[1:] - seg[:-1])[perm]
    off = np.zeros(perm.size + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    total = int(off[-1])
    idx = np.arange(total, dtype=np.int64) - np.repeat(off[:-1], lens) + np.repeat(seg[:-1][perm], lens)
    return (off, np.ascontiguousarray(ent["box"][idx]), np.ascontiguousarray(ent["slot"][idx]),
            ent["src"][idx])


# ---- LVIS frame path ---------------------------------------------------------------------
def prepare_lvis(gt: GtColumns, dt: DtColumns, max_dets: int = MAX_DETS,
                 vis_rng=LVIS_VIS_RNG, img_ids=None, cat_ids=None,
                 use_cats: bool = True, allow_empty: bool = False) -> EvalPlan:
    """Plan of the frame path; img_ids / cat_ids as in prepare_tao (lvis eval.py:51-52);
    allow_empty as in prepare_tao (a shard without predictions)."""
    cat_ids = np.unique(gt.cat_id if cat_ids is None else np.asarray(cat_ids, dtype=np.int64))
    all_img_ids = np.unique(gt.img_id)
    img_ids = all_img_ids if img_ids is None else np.unique(np.asarray(img_ids, dtype=np.int64))
    n_cat, n_img = cat_ids.size, img_ids.size
    img_keys, img_row = _last_row_of(gt.img_id)

    # ground truth (lvis.py:63-97, eval.py:64-80); no category merge on this path
    g_unit_all = _index_of(img_ids, gt.ann_image_id)
    g_valid = ((g_unit_all >= 0) & (_index_of(cat_ids, gt.ann_category_id) >= 0)
               & (gt.ann_area > 0) & (gt.ann_area < INF))
    g_rows = np.nonzero(g_valid)[0]
    g_unit = g_unit_all[g_rows]
    g_cidx = _index_of(cat_ids, gt.ann_category_id[g_rows])
    if np.isnan(gt.ann_visibility[g_rows]).any():
        raise KeyError("visibility")                                   # eval.py:204
    if (gt.ann_oof[g_rows] == 2).any():
        raise KeyError("out_of_frame")                                 # eval.py:213

    # detections (lvis results.py:29-71)
    if dt.n() == 0 and not allow_empty:
        raise IndexError("list index out of range")
    sel = limit_dets_per_image(dt.image_id, dt.score, max_dets)
    d_img = dt.image_id[sel]
    d_cat = dt.category_id[sel]
    d_box = dt.bbox[sel]
    d_score = dt.score[sel].astype(np.float64)
    d_area = d_box[:, 2] * d_box[:, 3]
    if (_index_of(all_img_ids, d_img) < 0).any():
        raise AssertionError("Results do not correspond to current LVIS set.")
    d_unit_all = _index_of(img_ids, d_img)
    d_valid = ((d_unit_all >= 0) & (_index_of(cat_ids, d_cat) >= 0)
               & (d_area > 0) & (d_area < INF))
    d_rows = np.nonzero(d_valid)[0]
    d_unit = d_unit_all[d_rows]
    d_cidx = _index_of(cat_ids, d_cat[d_rows])

    # federated filter with image-level lists (eval.py:88-103)
    if not gt.has_image_lists:
        raise KeyError("neg_category_ids")
    row_of_unit = img_row[_index_of(img_keys, img_ids)]
    present = np.unique(g_unit.astype(np.int64) * n_cat + g_cidx)
    in_present = _index_of(present, d_unit.astype(np.int64) * n_cat + d_cidx) >= 0
    in_neg = _ragged_contains(gt.img_neg, row_of_unit[d_unit], d_cat[d_rows])
    in_nel = _ragged_contains(gt.img_nel, row_of_unit[d_unit], d_cat[d_rows])
    kept = np.nonzero(in_present | in_neg)[0]

    (keys, grp_cat, grp_unit, grp_gt_off, grp_dt_off, iou_off, g_perm, d_perm) = _layout(
        n_img, g_unit, g_cidx, g_rows, d_unit[kept], d_cidx[kept], d_rows[kept],
        d_score[d_rows][kept], use_cats)
    d_sel = d_rows[kept][d_perm]
    g_sel = g_rows[g_perm]

    # frequency groups (eval.py:107-113): indices into the category list
    cat_keys, cat_row = _last_row_of(gt.cat_id)
    freq = gt.cat_freq[cat_row[_index_of(cat_keys, cat_ids)]]
    if (freq > 2).any():
        raise KeyError("frequency")
    freq_groups = [np.nonzero(freq == k)[0].tolist() for k in range(3)]
    if not use_cats:
        cat_ids, n_cat = np.asarray([-1], dtype=np.int64), 1      # lvis eval.py:127-128
    cat_grp_off, cat_dt_off = _cat_offsets(n_cat, grp_cat, grp_dt_off)
    dt_score = np.ascontiguousarray(d_score[d_sel])

    plan = EvalPlan(
        kind="lvis", cat_ids=cat_ids, unit_ids=img_ids, n_groups=int(keys.size),
        grp_cat=grp_cat, grp_unit=grp_unit, grp_dt_off=grp_dt_off, grp_gt_off=grp_gt_off,
        iou_off=iou_off, cat_grp_off=cat_grp_off, cat_dt_off=cat_dt_off,
        acc_perm=_acc_perm(n_cat, cat_dt_off, dt_score),
        dt_score=dt_score,
        dt_attr_a=np.ascontiguousarray(d_area[d_sel]),
        dt_attr_b=np.zeros(d_sel.size),
        dt_flag=np.ascontiguousarray(in_nel[kept][d_perm].astype(np.uint8) | np.uint8(2)),
        dt_id=np.ascontiguousarray((d_sel + 1).astype(np.int64)),
        gt_attr_a=np.ascontiguousarray(gt.ann_visibility[g_sel]),
        gt_attr_b=np.zeros(g_sel.size),
        gt_hp=np.zeros(g_sel.size, dtype=np.int32),
        gt_flag=np.ascontiguousarray(((gt.ann_ignore[g_sel] & 1)
                                      | ((gt.ann_oof[g_sel] == 1).astype(np.uint8) << 1)
                                      | ((gt.ann_id[g_sel] == 0).astype(np.uint8) << 2))
                                     .astype(np.uint8)),
        gt_id=np.ascontiguousarray(gt.ann_id[g_sel]),
        dt_box=np.ascontiguousarray(d_box[d_sel].astype(np.float64)),
        gt_box=np.ascontiguousarray(gt.ann_bbox[g_sel].astype(np.float64)),
        range_cfgs=lvis_range_cfgs(vis_rng), sentinel=0, freq_groups=freq_groups,
        dt_box_src=np.ascontiguousarray(sel[d_sel]),
    )
    plan.stats["n_dt_boxes"] = float(d_sel.size)
    plan.stats["n_gt_boxes"] = float(g_sel.size)
    plan.stats["box_pairs"] = float(iou_off[-1])
    return plan


# ---- workload accounting -----------------------------------------------------------------
def count_box_pair_visits(plan: EvalPlan) -> int:
    """The metric's unit.  Track path: one iteration of the reference's
    per-frame loop (tao_amodal/evaluation/tao_amodal/eval.py:83-84), i.e. the sum over all
    (dt track, gt track) pairs of a group of |F_d ∪ F_g|.  Frame path: sum of D*G."""
    if plan.kind != "tao":
        return int(plan.iou_off[-1])
    total = 0
    n_slots = 1 + max(int(plan.dt_box_slot.max()) if plan.dt_box_slot.size else 0,
                      int(plan.gt_box_slot.max()) if plan.gt_box_slot.size else 0)
    d_len = np.diff(plan.dt_trk_box_off)
    g_len = np.diff(plan.gt_trk_box_off)
    for g in np.nonzero(np.diff(plan.iou_off) > 0)[0]:
        d0, d1 = int(plan.grp_dt_off[g]), int(plan.grp_dt_off[g + 1])
        g0, g1 = int(plan.grp_gt_off[g]), int(plan.grp_gt_off[g + 1])
        D, G = d1 - d0, g1 - g0
        pd = np.zeros((D, n_slots), dtype=np.float32)
        pg = np.zeros((G, n_slots), dtype=np.float32)
        b0, b1 = int(plan.dt_trk_box_off[d0]), int(plan.dt_trk_box_off[d1])
        pd[np.repeat(np.arange(D), d_len[d0:d1]), plan.dt_box_slot[b0:b1]] = 1.0
        b0, b1 = int(plan.gt_trk_box_off[g0]), int(plan.gt_trk_box_off[g1])
        pg[np.repeat(np.arange(G), g_len[g0:g1]), plan.gt_box_slot[b0:b1]] = 1.0
        common = int(round(float((pd @ pg.T).sum())))
        total += int(d_len[d0:d1].sum()) * G + int(g_len[g0:g1].sum()) * D - common
    return total
