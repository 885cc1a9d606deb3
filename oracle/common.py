"""TEST INFRASTRUCTURE — CPU oracle, shared pieces.  Not part of the product path.

Plain-Python / numpy restatement of the arithmetic, the greedy matcher and the
precision-recall accumulation that both reference evaluators share.  Every
function cites the reference lines it restates (paths relative to the
reference root).  Parity of this oracle is PINNED against outputs of the
unmodified reference run in the build container: ``oracle/make_golden.py``
generated ``tests/golden/*.npz`` from the reference, and
``tests/test_oracle_golden.py`` checks this module against them, plus the
three valid doctest vectors of ``bb_intersect_union``
(tao_amodal/evaluation/tao_amodal/eval.py:21-24,29-30) and the known-answer
cases of SURVEY.md §8c.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.
"""
from __future__ import annotations

import numpy as np

IOU_THRS = np.linspace(0.5, 0.95, int(np.round((0.95 - 0.5) / 0.05)) + 1, endpoint=True)
REC_THRS = np.linspace(0.0, 1.00, int(np.round((1.00 - 0.0) / 0.01)) + 1, endpoint=True)


def box_inter_union(d, g):
    """tao_amodal/evaluation/tao_amodal/eval.py:15-48 (bb_intersect_union).

    Python ``max``/``min`` semantics (first argument wins ties), IEEE fp64, no
    fused multiply-add; ``(da + ga) - i`` in that association."""
    dx, dy, dw, dh = d
    gx, gy, gw, gh = g
    da = dw * dh
    ga = gw * gh
    left = gx if gx > dx else dx
    r0, r1 = dx + dw, gx + gw
    right = r1 if r1 < r0 else r0
    top = gy if gy > dy else dy
    b0, b1 = dy + dh, gy + gh
    bottom = b1 if b1 < b0 else b0
    w = right - left
    if 0 > w:
        w = 0
    h = bottom - top
    if 0 > h:
        h = 0
    inter = w * h
    return inter, da + ga - inter


def frame_box_iou(d, g):
    """pycocotools bbIou with iscrowd=0 — in-tree copy at visualization/tao/
    third_party/pysot/training_dataset/coco/pycocotools/common/maskApi.c:109-120."""
    da = d[2] * d[3]
    ga = g[2] * g[3]
    w = min(d[2] + d[0], g[2] + g[0]) - max(d[0], g[0])
    if w <= 0:
        return 0.0
    h = min(d[3] + d[1], g[3] + g[1]) - max(d[1], g[1])
    if h <= 0:
        return 0.0
    i = w * h
    u = da + ga - i
    return i / u


def greedy_match(ious, gt_flag, gt_ids, dt_ids, dt_unmatched_ignore, iou_thrs, sentinel):
    """COCO-style sequential greedy assignment at every IoU threshold.

    tao_amodal/evaluation/tao_amodal/eval.py:385-443 (sentinel -1) and
    lvis_amodal/eval.py:233-292 (sentinel 0).

    ``ious`` is [D,G] with dts in descending-score order and gts ALREADY in the
    ignore-last order; ``gt_flag`` the matching _ignore flags.  Returns
    (dt_m [T,D], gt_m [T,G], dt_ig [T,D] bool)."""
    T = len(iou_thrs)
    G = len(gt_ids)
    D = len(dt_ids)
    gt_m = np.zeros((T, G)) + sentinel
    dt_m = np.zeros((T, D)) + sentinel
    dt_ig = np.zeros((T, D))
    if D and G:
        for ti, thr in enumerate(iou_thrs):
            for di in range(D):
                best = min([thr, 1 - 1e-10])
                m = -1
                for gi in range(G):
                    if gt_m[ti, gi] > 0:          # eval.py:407 — "taken" tests the stored dt id
                        continue
                    if m > -1 and gt_flag[m] == 0 and gt_flag[gi] == 1:
                        break
                    if ious[di, gi] < best:
                        continue
                    best = ious[di, gi]
                    m = gi
                if m == -1:
                    continue
                dt_ig[ti, di] = gt_flag[m]
                dt_m[ti, di] = gt_ids[m]
                gt_m[ti, m] = dt_ids[di]
    mask = np.array(dt_unmatched_ignore, dtype=bool).reshape((1, D))
    mask = np.repeat(mask, T, 0)
    dt_ig = np.logical_or(dt_ig, np.logical_and(dt_m == sentinel, mask))
    return dt_m, gt_m, dt_ig


def pr_curve(dt_scores, dt_m, dt_ig, gt_ig, sentinel, rec_thrs):
    """One (category, range) cell of accumulate():
    tao_amodal/evaluation/tao_amodal/eval.py:508-573, lvis_amodal/eval.py:353-413.

    Inputs are the per-cell concatenations in reference order.  Returns None
    when the cell has no non-ignored GT (precision/recall stay -1), else
    (precision [T,R], recall [T], order, tps, fps)."""
    order = np.argsort(-dt_scores, kind="mergesort")
    dt_m = dt_m[:, order]
    dt_ig = dt_ig[:, order]
    num_gt = np.count_nonzero(gt_ig == 0)
    if num_gt == 0:
        return None
    matched = dt_m != sentinel
    tps = np.logical_and(matched, np.logical_not(dt_ig))
    fps = np.logical_and(np.logical_not(matched), np.logical_not(dt_ig))
    tp_sum = np.cumsum(tps, axis=1).astype(dtype=float)
    fp_sum = np.cumsum(fps, axis=1).astype(dtype=float)
    T = dt_m.shape[0]
    R = len(rec_thrs)
    precision = np.zeros((T, R))
    recall = np.zeros(T)
    for ti in range(T):
        tp = tp_sum[ti]
        fp = fp_sum[ti]
        n = len(tp)
        rc = tp / num_gt
        recall[ti] = rc[-1] if n else 0
        pr = (tp / (fp + tp + np.spacing(1))).tolist()
        for i in range(n - 1, 0, -1):
            if pr[i] > pr[i - 1]:
                pr[i - 1] = pr[i]
        where = np.searchsorted(rc, rec_thrs, side="left")
        row = [0.0] * R
        for k, pi in enumerate(where):
            if pi >= n:      # the reference's bare try/except IndexError, :567-571
                break
            row[k] = pr[pi]
        precision[ti] = np.array(row)
    return precision, recall, order, tps, fps


def masked_mean(s):
    """tao_amodal/evaluation/tao_amodal/eval.py:619-623."""
    sel = s[s > -1]
    if len(sel) == 0:
        return -1
    return np.mean(sel)
