"""TEST INFRASTRUCTURE — mask known-answer vectors from the reference's own C code.

    python -m oracle.make_mask_vectors       # build container only (needs oracle/_ref)

Polygons, multi-part polygons, box polygons (incl. sub-pixel, out-of-frame, repeated points),
uncompressed and compressed RLE are converted by maskApi.c (rleFrPoly :164-216, rleMerge
:50-71, rleFrString :233-246) and compared pairwise by rleIou (:78-96); counts, boxes, areas,
strings and the IoU matrix go to tests/golden/mask_ref_c.json.
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import maskapi_ref as M

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden", "mask_ref_c.json")


def main():
    rng = np.random.Generator(np.random.PCG64(2024))
    H, W = 48, 64
    segs = []
    for t in range(24):
        k = int(rng.integers(3, 8))
        pts = (rng.uniform(-6, 70, (k, 2)) * [1, 0.75])
        if t % 3 == 0:
            pts = np.round(pts * 2) / 2
        segs.append([pts.reshape(-1).tolist()])
    for t in range(8):
        segs.append([(rng.uniform(0, 60, 8) * 0.8).tolist(), (rng.uniform(0, 60, 10) * 0.8).tolist()])
    for x, y, w, h in [(3, 4, 20, 10), (3.3, 4.7, 20.2, 9.9), (-5, -5, 12, 12), (60, 40, 10, 10),
                       (10, 10, 0.05, 0.05), (0, 0, 64, 48), (30, 20, 0, 5), (5, 5, 1, 1)]:
        segs.append([[x, y, x, y + h, x + w, y + h, x + w, y]])
    segs.append([[4.0, 4.0, 4.0, 4.0, 20.0, 30.0]])        # repeated point
    masks = []
    rles = []
    for s in segs:
        r = M.merge(M.frPyObjects(s, H, W))
        rles.append(r)
        masks.append({"segm": s, "h": H, "w": W})
    m = (rng.random((H, W)) < 0.3).astype(np.uint8)
    r = M.encode(np.asfortranarray(m[:, :, None]))[0]
    rles.append(r)
    masks.append({"segm": {"size": [H, W], "counts": r["counts"].decode()}, "h": H, "w": W})
    rles.append(r)
    masks.append({"segm": {"size": [H, W], "counts": M.rle_counts(r).tolist()}, "h": H, "w": W})
    for rec, r in zip(masks, rles):
        rec["counts"] = M.rle_counts(r).tolist()
        rec["bbox"] = M.toBbox(r).tolist()
        rec["area"] = int(M.area(r))
        rec["string"] = r["counts"].decode()
    iou = M.iou(rles, rles, [0] * len(rles))
    json.dump({"masks": masks, "iou": iou.tolist()}, open(OUT, "w"))
    print("wrote", OUT, len(masks), "masks;", int((iou > 0).sum()), "overlapping pairs")


if __name__ == "__main__":
    main()
