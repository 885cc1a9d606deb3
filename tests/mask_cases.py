"""Seeded mask inputs shared by the CPU and GPU tests of the segm path."""
import numpy as np


def random_mask_groups(seed, n_groups=12, H=72, W=128):
    """Groups of detection / GT masks (boxes, polygons, multi-part polygons, empty masks) on
    one canvas size, plus one group whose GT canvas differs (IoU -1, maskApi.c:85)."""
    from tao_amodal_b200.mask import RlePool
    rng = np.random.Generator(np.random.PCG64(seed))
    dt, gt = RlePool(), RlePool()
    segs = {"dt": [], "gt": []}
    d_off, g_off = [0], [0]
    for grp in range(n_groups):
        D, G = int(rng.integers(0, 7)), int(rng.integers(0, 5))
        centre = rng.uniform(20, 100, 2) * [1, 0.5]
        for side, pool, n in (("dt", dt, D), ("gt", gt, G)):
            for _ in range(n):
                kind = rng.integers(0, 4)
                hh, ww = (H, W) if not (grp == 5 and side == "gt") else (H + 1, W)
                if kind == 0:      # box-shaped polygon near the group centre
                    x, y = centre + rng.uniform(-15, 15, 2)
                    bw, bh = rng.uniform(0.5, 40, 2)
                    seg = [[x, y, x, y + bh, x + bw, y + bh, x + bw, y]]
                elif kind == 1:    # general polygon
                    k = int(rng.integers(3, 8))
                    pts = centre + rng.uniform(-30, 30, (k, 2))
                    seg = [pts.reshape(-1).tolist()]
                elif kind == 2:    # two parts
                    seg = []
                    for _p in range(2):
                        pts = centre + rng.uniform(-25, 25, (4, 2))
                        seg.append(pts.reshape(-1).tolist())
                else:              # far away / empty
                    seg = [[-50.0, -50.0, -50.0, -40.0, -40.0, -40.0, -40.0, -50.0]]
                pool.add_segmentation(seg, hh, ww)
                segs[side].append((seg, hh, ww))
        d_off.append(d_off[-1] + D)
        g_off.append(g_off[-1] + G)
    d_off, g_off = np.asarray(d_off, dtype=np.int64), np.asarray(g_off, dtype=np.int64)
    iou_off = np.concatenate([[0], np.cumsum(np.diff(d_off) * np.diff(g_off))]).astype(np.int64)
    return dt, gt, d_off, g_off, iou_off, segs
