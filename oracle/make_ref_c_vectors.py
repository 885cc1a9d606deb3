"""TEST INFRASTRUCTURE — known-answer vectors from the reference's own C code.

    python -m oracle.make_ref_c_vectors        # build container only (needs oracle/_ref)

Feeds seeded box pairs (dyadic-grid, off-grid, touching, contained, zero-area, negative
width) to bbIou of the reference tree's maskApi.c (:109-120, compiled by oracle/Makefile)
and stores inputs + outputs in tests/golden/bbiou_ref_c.npz, so the pin of
oracle.common.frame_box_iou and of the device function ta_bb_iou travels to machines that
have neither the reference tree nor oracle/_ref.
"""
from __future__ import annotations

import os

import numpy as np

from . import maskapi_ref

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden", "bbiou_ref_c.npz")


def boxes(seed=77, n=64):
    rng = np.random.Generator(np.random.PCG64(seed))
    grid = np.round(rng.uniform(0, 300, (n, 4)) * 8) / 8          # 1/8-px grid
    grid[:, 2:] = np.maximum(grid[:, 2:], 0.125)
    free = rng.uniform(0, 300, (n, 4))                            # arbitrary doubles
    near = grid[rng.integers(0, n, n)] + rng.uniform(-3, 3, (n, 4))
    near[:, 2:] = np.abs(near[:, 2:]) + 1e-3
    special = np.array([
        [0, 0, 20, 20], [0, 0, 10, 10], [10, 20, 10, 10], [10, 20, 5, 5],   # eval.py:21-30 doctest boxes
        [20, 0, 10, 10],          # touches [0,0,20,20] on an edge: w == 0
        [5, 5, 0, 10],            # zero area
        [5, 5, -4, 10],           # negative width
        [0, 0, 1e-3, 1e6], [1e5, 1e5, 1e5, 1e5], [0.1, 0.2, 0.3, 0.4],
    ], dtype=np.double)
    return np.concatenate([grid, free, near, special])


def main():
    b = boxes()
    dt, gt = b, b[::-1].copy()
    out = maskapi_ref.iou(dt, gt, [0] * len(gt))
    np.savez_compressed(OUT, dt=dt, gt=gt, iou=out)
    print("wrote", OUT, out.shape, "non-zero:", int((out > 0).sum()))


if __name__ == "__main__":
    main()
