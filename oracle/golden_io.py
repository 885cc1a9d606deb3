"""TEST INFRASTRUCTURE — canonical flattening of evaluator outputs for golden files.

A "cell" is what the reference stores per (video|image, category, range...)
grid entry (tao_amodal/evaluation/tao_amodal/eval.py:445-457,
lvis_amodal/eval.py:292-303).  Golden files hold, for every non-None cell in
sorted key order, the integer decisions (matched ids, ignore flags) flattened
into 1-D arrays so they compare with ``np.array_equal``.
"""
from __future__ import annotations

import numpy as np


def flatten_cells(cells: dict) -> dict:
    keys = sorted(k for k, v in cells.items() if v is not None)
    key_arr = np.asarray(keys, dtype=np.int64).reshape(len(keys), -1)
    nd, ng = [], []
    dt_ids, gt_ids, dt_m, gt_m, dt_ig, gt_ig, dt_sc = [], [], [], [], [], [], []
    for k in keys:
        e = cells[k]
        D, G = len(e["dt_ids"]), len(e["gt_ids"])
        nd.append(D)
        ng.append(G)
        dt_ids.append(np.asarray(e["dt_ids"], dtype=np.int64).reshape(-1))
        gt_ids.append(np.asarray(e["gt_ids"], dtype=np.int64).reshape(-1))
        dt_m.append(np.asarray(e["dt_matches"]).astype(np.int64).reshape(-1))
        gt_m.append(np.asarray(e["gt_matches"]).astype(np.int64).reshape(-1))
        dt_ig.append(np.asarray(e["dt_ignore"]).astype(np.uint8).reshape(-1))
        gt_ig.append(np.asarray(e["gt_ignore"]).astype(np.uint8).reshape(-1))
        dt_sc.append(np.asarray(e["dt_scores"], dtype=np.float64).reshape(-1))

    def cat(xs, dt):
        return np.concatenate(xs) if xs else np.zeros(0, dtype=dt)

    return {
        "cell_keys": key_arr, "cell_nd": np.asarray(nd, dtype=np.int64),
        "cell_ng": np.asarray(ng, dtype=np.int64),
        "cell_dt_ids": cat(dt_ids, np.int64), "cell_gt_ids": cat(gt_ids, np.int64),
        "cell_dt_m": cat(dt_m, np.int64), "cell_gt_m": cat(gt_m, np.int64),
        "cell_dt_ig": cat(dt_ig, np.uint8), "cell_gt_ig": cat(gt_ig, np.uint8),
        "cell_dt_scores": cat(dt_sc, np.float64),
    }


def flatten_ious(ious: dict) -> dict:
    keys = sorted(k for k, v in ious.items() if len(v) > 0 and np.asarray(v).size > 0)
    shp = [np.asarray(ious[k]).shape for k in keys]
    vals = [np.asarray(ious[k], dtype=np.float64).reshape(-1) for k in keys]
    return {
        "iou_keys": np.asarray(keys, dtype=np.int64).reshape(len(keys), -1),
        "iou_shape": np.asarray(shp, dtype=np.int64).reshape(len(keys), -1),
        "iou_vals": np.concatenate(vals) if vals else np.zeros(0),
    }


def results_vector(results) -> np.ndarray:
    return np.asarray([float(v) for v in results.values()], dtype=np.float64)


def results_keys(results) -> list:
    return [k if isinstance(k, str) else "|".join(str(x) for x in k) for k in results.keys()]
