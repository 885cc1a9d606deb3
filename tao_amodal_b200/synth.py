"""Seeded synthetic TAO-Amodal annotation + prediction generator.

Concrete restatement of the workload shapes in BASELINE.json ``configs``
(SURVEY.md §8d).  All box coordinates are quantised to 1/8 px so every partial
sum the evaluator forms (areas, intersections, unions) is exactly
representable in fp64; the track IoU is then independent of the frame
summation order, which is what makes bit-exact parity with the reference's
set-ordered loop (tao_amodal/evaluation/tao_amodal/eval.py:83-94) a guarantee
rather than a likelihood.

Output is columnar (``GtColumns`` / ``DtColumns``); ``.to_dict()`` /
``.to_list()`` on those give the reference JSON structures for sizes where
JSON is practical.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from .columnar import DtColumns, GtColumns, ragged_from_lists

Q = 8.0  # coordinates live on a 1/8 px grid


@dataclass
class SynthConfig:
    name: str
    videos: int
    frames: int
    pred_tracks: int
    gt_tracks: int
    categories: int
    seed: int
    gt_span: Optional[int] = None     # fixed alive span for every GT track (cfg5 uses 10)
    pred_full_length: bool = False    # predicted tracks that are not copies of a GT track cover
                                      # all frames (cfg5); copies always live on the GT's span —
                                      # a 40-frame copy of a 10-frame GT track would have 3-D IoU
                                      # <= 0.25 and the set's AP would be identically 0
    score_quantum: float = 0.0        # >0: scores are multiples of this (tie stress)
    sparse_image_ids: bool = False    # non-consecutive ids -> CPython set order != ascending
    keep_prob: float = 0.9            # per-frame keep probability of a copied track
    variety: bool = True              # mix in short / small / highly-visible GT tracks so every
                                      # area, duration and occlusion range of Params is exercised
    max_present: int = 5


# BASELINE.json configs[1..4] (SURVEY.md §8d names them cfg2..cfg5)
CONFIGS = {
    "tiny": SynthConfig("tiny", videos=3, frames=12, pred_tracks=8, gt_tracks=4,
                        categories=12, seed=7),
    "small": SynthConfig("small", videos=6, frames=40, pred_tracks=20, gt_tracks=6,
                         categories=30, seed=11),
    "cfg2": SynthConfig("cfg2", videos=10, frames=300, pred_tracks=50, gt_tracks=10,
                        categories=100, seed=1001),
    "cfg3": SynthConfig("cfg3", videos=500, frames=300, pred_tracks=200, gt_tracks=30,
                        categories=1203, seed=1002),
    # 1 M predicted boxes: 8 tracks per video, half of them 10-frame copies of a GT track, the
    # others 40-frame static boxes -> 5000 x 8 x 25 boxes on average
    "cfg5": SynthConfig("cfg5", videos=5000, frames=40, pred_tracks=8, gt_tracks=2,
                        categories=1203, seed=1004, gt_span=10, pred_full_length=True),
}


def _quant(a: np.ndarray) -> np.ndarray:
    return np.round(a * Q) / Q


def _walk(rng: np.random.Generator, start: np.ndarray, n: int, step: float) -> np.ndarray:
    """Random walk of n frames from ``start`` ([4]) with +-step px increments."""
    inc = rng.uniform(-step, step, size=(n, 4))
    inc[0] = 0.0
    path = start[None, :] + np.cumsum(inc, axis=0)
    path[:, 2] = np.maximum(path[:, 2], 4.0)
    path[:, 3] = np.maximum(path[:, 3], 4.0)
    path[:, 0] = np.maximum(path[:, 0], 0.0)
    path[:, 1] = np.maximum(path[:, 1], 0.0)
    return _quant(path)


def _rand_box(rng: np.random.Generator) -> np.ndarray:
    return np.array([rng.uniform(0, 900), rng.uniform(0, 500),
                     rng.uniform(10, 300), rng.uniform(10, 200)])


def generate(cfg: SynthConfig) -> Tuple[GtColumns, DtColumns]:
    rng = np.random.Generator(np.random.PCG64(cfg.seed))
    V, F, P, G, C = cfg.videos, cfg.frames, cfg.pred_tracks, cfg.gt_tracks, cfg.categories
    cat_ids = np.arange(1, C + 1, dtype=np.int64)

    n_img = V * F
    if cfg.sparse_image_ids:
        img_ids = np.sort(rng.choice(np.arange(1, 50 * n_img + 1000), size=n_img,
                                     replace=False)).astype(np.int64)
        # shuffle which video gets which id block so ids are not monotone in dataset order
        img_ids = img_ids.reshape(V, F)[rng.permutation(V)].reshape(-1)
    else:
        img_ids = np.arange(1, n_img + 1, dtype=np.int64)
    img_vid = np.repeat(np.arange(1, V + 1, dtype=np.int64), F)
    img_fi = np.tile(np.arange(F, dtype=np.int64), V)

    vid_neg, vid_nel = [], []
    g_img, g_trk, g_cat, g_box, g_vis, g_oof = [], [], [], [], [], []
    t_id, t_cat, t_vid = [], [], []
    p_img, p_trk, p_cat, p_vid, p_box, p_score = [], [], [], [], [], []
    next_gt_track = 1
    next_pred_track = 1

    for v in range(V):
        vid = v + 1
        base = v * F
        n_present = min(G, cfg.max_present)
        present = rng.choice(cat_ids, size=n_present, replace=False)
        absent_pool = np.setdiff1d(cat_ids, present, assume_unique=True)
        n_neg = min(10, max(0, len(absent_pool) - 1))
        neg = rng.choice(absent_pool, size=n_neg, replace=False)
        other_pool = np.setdiff1d(absent_pool, neg, assume_unique=True)
        nel = rng.choice(present, size=1)
        vid_neg.append(neg.tolist())
        vid_nel.append(nel.tolist())

        # ---- ground-truth tracks
        gt_paths = []
        for g in range(G):
            cat = int(present[g % n_present])
            if cfg.gt_span is not None:
                span = min(cfg.gt_span, F)
                s = int(rng.integers(0, F - span + 1))
                e = s + span
            else:
                s = int(rng.integers(0, max(1, F // 3)))
                e = int(rng.integers(max(s + 1, (2 * F) // 3), F + 1))
            if cfg.variety and cfg.gt_span is None and rng.uniform() < 0.12:
                n_short = int(rng.integers(1, 13))
                s = int(rng.integers(0, max(1, F - n_short + 1)))
                e = min(F, s + n_short)
            n = e - s
            start = _rand_box(rng)
            if cfg.variety and rng.uniform() < 0.15:
                start[2:] = rng.uniform(8, 30, size=2)
            path = _walk(rng, start, n, 2.0)
            vis = rng.uniform(0.0, 1.0, size=n)
            if cfg.variety:
                mode = rng.uniform()
                if mode < 0.3:
                    vis = rng.uniform(0.8, 1.0, size=n)
                elif mode < 0.55:
                    vis = rng.uniform(0.8, 1.0, size=n)
                    n_occ = int(rng.integers(0, 9))
                    if n_occ:
                        vis[rng.choice(n, size=min(n, n_occ), replace=False)] = rng.uniform(0, 0.8)
            tid = next_gt_track
            next_gt_track += 1
            t_id.append(tid)
            t_cat.append(cat)
            t_vid.append(vid)
            g_img.append(img_ids[base + s: base + e])
            g_trk.append(np.full(n, tid, dtype=np.int64))
            g_cat.append(np.full(n, cat, dtype=np.int64))
            g_box.append(path)
            g_vis.append(vis)
            g_oof.append((rng.uniform(size=n) < 0.1).astype(np.uint8))
            gt_paths.append((s, e, cat, path))

        # ---- predicted tracks
        kinds = rng.choice(4, size=P, p=[0.5, 0.2, 0.2, 0.1])
        for k in range(P):
            kind = int(kinds[k])
            tid = next_pred_track
            next_pred_track += 1
            if kind == 0:
                s, e, cat, path = gt_paths[int(rng.integers(0, G))]
                keep = rng.uniform(size=e - s) < cfg.keep_prob
                if not keep.any():
                    keep[0] = True
                frames = np.arange(s, e)[keep]
                boxes = path[keep]
                boxes = boxes + _quant(rng.uniform(-3.0, 3.0, size=boxes.shape))
                boxes[:, 2:] = np.maximum(boxes[:, 2:], 1.0)
            else:
                if kind == 1:
                    cat = int(rng.choice(present))
                elif kind == 2 and len(neg):
                    cat = int(rng.choice(neg))
                elif len(other_pool):
                    cat = int(rng.choice(other_pool))
                else:
                    cat = int(rng.choice(present))
                if cfg.pred_full_length:
                    s, e = 0, F
                else:
                    s = int(rng.integers(0, max(1, F // 3)))
                    e = int(rng.integers(max(s + 1, (2 * F) // 3), F + 1))
                frames = np.arange(s, e)
                boxes = np.repeat(_quant(_rand_box(rng))[None, :], e - s, axis=0)
            n = len(frames)
            p_img.append(img_ids[base + frames])
            p_trk.append(np.full(n, tid, dtype=np.int64))
            p_cat.append(np.full(n, cat, dtype=np.int64))
            p_vid.append(np.full(n, vid, dtype=np.int64))
            p_box.append(boxes)
            sc = rng.uniform(0.0, 1.0)
            if cfg.score_quantum > 0:
                sc = float(np.round(sc / cfg.score_quantum) * cfg.score_quantum)
            p_score.append(np.full(n, sc))

    g_box_a = np.concatenate(g_box, axis=0)
    n_gt = g_box_a.shape[0]
    # image-level federated lists = the video's lists (TAO-Amodal LVIS-style files carry both)
    img_neg = ragged_from_lists(vid_neg[v] for v in range(V) for _ in range(F))
    img_nel = ragged_from_lists(vid_nel[v] for v in range(V) for _ in range(F))
    gt = GtColumns(
        img_id=img_ids, img_video_id=img_vid, img_frame_index=img_fi,
        img_neg=img_neg, img_nel=img_nel,
        vid_id=np.arange(1, V + 1, dtype=np.int64),
        vid_neg=ragged_from_lists(vid_neg), vid_nel=ragged_from_lists(vid_nel),
        trk_id=np.asarray(t_id, dtype=np.int64),
        trk_category_id=np.asarray(t_cat, dtype=np.int64),
        trk_video_id=np.asarray(t_vid, dtype=np.int64),
        trk_ignore=np.zeros(len(t_id), dtype=np.uint8),
        cat_id=cat_ids,
        cat_freq=(np.arange(C) % 3).astype(np.uint8),
        merge_map={},
        ann_id=np.arange(1, n_gt + 1, dtype=np.int64),
        ann_image_id=np.concatenate(g_img),
        ann_track_id=np.concatenate(g_trk),
        ann_category_id=np.concatenate(g_cat),
        ann_bbox=np.ascontiguousarray(g_box_a),
        ann_area=g_box_a[:, 2] * g_box_a[:, 3],
        ann_visibility=np.concatenate(g_vis),
        ann_oof=np.concatenate(g_oof),
        ann_ignore=np.zeros(n_gt, dtype=np.uint8),
    )
    dt = DtColumns(
        image_id=np.concatenate(p_img), track_id=np.concatenate(p_trk),
        category_id=np.concatenate(p_cat), video_id=np.concatenate(p_vid),
        bbox=np.ascontiguousarray(np.concatenate(p_box, axis=0)),
        score=np.concatenate(p_score),
    )
    # predictions arrive frame-major in real tracker output; interleave so that neither the
    # per-image nor the per-track order is trivially sorted.
    order = np.lexsort((dt.track_id, dt.image_id))
    for name in ("image_id", "track_id", "category_id", "video_id", "bbox", "score"):
        setattr(dt, name, np.ascontiguousarray(getattr(dt, name)[order]))
    return gt, dt


def generate_named(name: str, **overrides) -> Tuple[GtColumns, DtColumns]:
    cfg = CONFIGS[name]
    if overrides:
        cfg = SynthConfig(**{**cfg.__dict__, **overrides})
    return generate(cfg)
