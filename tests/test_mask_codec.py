"""Run-length mask codec (csrc/ta_mask.cpp) and the per-pair mask IoU the CUDA kernel runs
(ta_rle_pair_iou, host build) against the reference tree's own C code (oracle/_ref, built from
maskApi.c) and against committed vectors generated from it."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import plan_backends
from mask_cases import random_mask_groups
from oracle import maskapi_ref as M
from tao_amodal_b200.mask import RlePool

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "mask_ref_c.json")
needs_ref = pytest.mark.skipif(not M.available(), reason="oracle/_ref not built (no reference tree)")


def _hs_rle_iou(dt, gt, d_off, g_off, iou_off):
    hs = plan_backends.build_hostsim()
    do, dc, dhw, dbb, _ = dt.export()
    go, gc, ghw, gbb, _ = gt.export()
    out = np.full(max(int(iou_off[-1]), 1), 9.0)
    p = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
    keep = [do, dc, dhw, dbb, go, gc, ghw, gbb, d_off, g_off, iou_off]
    hs.hs_rle_iou.argtypes = [C.c_int64] + [C.c_void_p] * 12
    hs.hs_rle_iou(len(d_off) - 1, *[p(a) for a in (d_off, g_off, do, dc, dhw, dbb, go, gc, ghw, gbb, iou_off)], p(out))
    del keep
    return out[:int(iou_off[-1])]


def test_committed_vectors_from_reference_c():
    g = json.load(open(GOLDEN))
    pool = RlePool()
    for rec in g["masks"]:
        i = pool.add_segmentation(rec["segm"], rec["h"], rec["w"])
        off, cnt, hw, bb, ar = pool.export()
        assert cnt[off[i]:off[i + 1]].tolist() == rec["counts"]
        assert bb[i].tolist() == rec["bbox"] and int(ar[i]) == rec["area"]
        assert pool.to_rle(i)["counts"].decode() == rec["string"]
    # pair IoUs: every mask against every mask of the same canvas, as one group
    n = len(pool)
    d_off = g_off = np.asarray([0, n], dtype=np.int64)
    got = _hs_rle_iou(pool, pool, d_off, g_off, np.asarray([0, n * n], dtype=np.int64)).reshape(n, n)
    assert np.array_equal(got, np.asarray(g["iou"]))


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_live_reference_c_groups(seed):
    dt, gt, d_off, g_off, iou_off, segs = random_mask_groups(seed)
    got = _hs_rle_iou(dt, gt, d_off, g_off, iou_off)
    for grp in range(len(d_off) - 1):
        ds = [M.merge(M.frPyObjects(s, h, w)) for s, h, w in segs["dt"][d_off[grp]:d_off[grp + 1]]]
        gs = [M.merge(M.frPyObjects(s, h, w)) for s, h, w in segs["gt"][g_off[grp]:g_off[grp + 1]]]
        ref = M.iou(ds, gs, [0] * len(gs))
        mine = got[iou_off[grp]:iou_off[grp + 1]]
        if len(ds) == 0 or len(gs) == 0:
            assert mine.size == 0
        else:
            assert np.array_equal(mine.reshape(len(ds), len(gs)), ref), grp


@needs_ref
def test_codec_matches_reference_c_on_random_annotations():
    rng = np.random.Generator(np.random.PCG64(11))
    H, W = 60, 90
    pool = RlePool()
    for t in range(200):
        parts = [rng.uniform(-10, 100, int(rng.integers(3, 8)) * 2).tolist() for _ in range(int(rng.integers(1, 4)))]
        i = pool.add_segmentation(parts, H, W)
        ref = M.merge(M.frPyObjects(parts, H, W))
        assert pool.to_rle(i) == ref
    boxes = np.concatenate([rng.uniform(-10, 80, (100, 2)), rng.uniform(0, 50, (100, 2))], 1)
    first = pool.add_boxes(boxes, H, W)
    for k, ref in enumerate(M.frPyObjects(boxes, H, W)):
        assert pool.to_rle(first + k) == ref
    off, cnt, hw, bb, ar = pool.export()
    refs = [pool.to_rle(i) for i in range(len(pool))]
    assert np.array_equal(bb, M.toBbox(refs)) and np.array_equal(ar, M.area(refs))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_rle_iou_kernel_matches_host_arithmetic(seed):
    """ta_rle_iou on the device against the same per-pair function built for the host."""
    import torch
    from tao_amodal_b200 import _lib
    from tao_amodal_b200.engine import Engine
    dt, gt, d_off, g_off, iou_off, _ = random_mask_groups(seed, n_groups=40)
    ref = _hs_rle_iou(dt, gt, d_off, g_off, iou_off)
    eng = Engine(0)
    dev = torch.device("cuda", 0)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    do, dc, dhw, dbb, _ = dt.export()
    go, gc, ghw, gbb, _ = gt.export()
    pad = lambda a, shape, t: a if a.size else np.zeros(shape, dtype=t)
    ts = [up(x) for x in (d_off, g_off, do, pad(dc, 1, np.uint32).view(np.int32), pad(dhw, (1, 2), np.uint32).view(np.int32),
                          pad(dbb, (1, 4), np.float64), go, pad(gc, 1, np.uint32).view(np.int32),
                          pad(ghw, (1, 2), np.uint32).view(np.int32), pad(gbb, (1, 4), np.float64), iou_off)]
    out = torch.full((max(int(iou_off[-1]), 1),), 9.0, dtype=torch.float64, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream(0).cuda_stream)
    _lib.check(eng.lib.ta_rle_iou(eng._ctx, st, len(d_off) - 1, None, 0, p(ts[0]), p(ts[1]),
                                  p(ts[2]), p(ts[3]), p(ts[4]), p(ts[5]), p(ts[6]), p(ts[7]),
                                  p(ts[8]), p(ts[9]), p(ts[10]), p(out)))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy()[:int(iou_off[-1])], ref)
    eng.close()
