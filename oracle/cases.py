"""TEST INFRASTRUCTURE — named parity cases (inputs only).

Each case returns ``(gt_dict, results_list)`` in the reference's JSON
structures (tao.py:4-60, README.md:108-115).  ``oracle/make_golden.py`` feeds
them to the unmodified reference; the tests feed them to the oracle and to the
CUDA path.  Cases cover the edge conditions of SURVEY.md §8d: CPython set
order on sparse ids, score ties, >300 detections per frame, zero-area boxes,
track id 0, duplicate track ids across videos, per-box scores inside a track,
ignored GT, merged categories, GT annotation id 0, duplicate frames in a track,
real-valued (off-grid) boxes.
"""
from __future__ import annotations

import copy

import numpy as np

from tao_amodal_b200 import synth


def _base(name, **kw):
    gt, dt = synth.generate_named(name, **kw)
    return gt.to_dict(), dt.to_list()


def case_tiny():
    return _base("tiny")


def case_small():
    return _base("small")


def case_small_sparse_ids():
    return _base("small", sparse_image_ids=True, seed=12)


def case_small_ties():
    return _base("small", score_quantum=0.05, seed=13)


def case_small_float():
    """Off-grid real-valued boxes: partial sums are inexact, so the frame
    summation order can change the last bit of a track IoU."""
    gt, res = _base("small", seed=14)
    f = float(np.pi / 3.0)
    for a in gt["annotations"]:
        a["bbox"] = [x * f for x in a["bbox"]]
        a["area"] = a["bbox"][2] * a["bbox"][3]
    for r in res:
        r["bbox"] = [x * f for x in r["bbox"]]
    return gt, res


def case_edge_mix():
    gt, res = _base("small", seed=15)
    rng = np.random.Generator(np.random.PCG64(99))
    imgs = gt["images"]
    vids = gt["videos"]
    cats = [c["id"] for c in gt["categories"]]

    # merged categories: the last category absorbs a fresh id 1000 (tao.py:97-106)
    gt["categories"][-1]["merged"] = [{"id": 1000}]
    # re-label one predicted track into the merged-away id
    tid0 = res[0]["track_id"]
    for r in res:
        if r["track_id"] == tid0:
            r["category_id"] = 1000

    # > 300 detections on one frame: 330 extra single-box tracks in a present category
    img0 = imgs[3]
    vid0 = img0["video_id"]
    present_cat = next(t["category_id"] for t in gt["tracks"] if t["video_id"] == vid0)
    base_tid = max(r["track_id"] for r in res) + 1
    for k in range(330):
        res.append({"image_id": img0["id"], "track_id": base_tid + k, "category_id": present_cat,
                    "video_id": vid0,
                    "bbox": [float(k % 40) * 8, float(k // 40) * 8, 48.0, 36.0],
                    "score": float(np.round(rng.uniform(), 3))})

    # zero-area and negative-width predictions (dropped by 0 < area, tao.py:251-252)
    res.append({"image_id": imgs[5]["id"], "track_id": base_tid + 400, "category_id": present_cat,
                "video_id": imgs[5]["video_id"], "bbox": [10.0, 10.0, 0.0, 20.0], "score": 0.9})
    res.append({"image_id": imgs[6]["id"], "track_id": base_tid + 401, "category_id": present_cat,
                "video_id": imgs[6]["video_id"], "bbox": [10.0, 10.0, -5.0, 20.0], "score": 0.9})

    # track id 0 on a high-scoring copy of a GT track (the gt_m > 0 quirk, eval.py:407)
    g_tid = gt["tracks"][0]["id"]
    g_anns = [a for a in gt["annotations"] if a["track_id"] == g_tid]
    for a in g_anns:
        res.append({"image_id": a["image_id"], "track_id": 0,
                    "category_id": a["category_id"],
                    "video_id": gt["tracks"][0]["video_id"],
                    "bbox": list(a["bbox"]), "score": 0.999})

    # the same track id reused in two videos (tools/eval_on_tao_amodal.py:44-66)
    other = [r for r in res if r["video_id"] != res[0]["video_id"]]
    reuse_id = other[0]["track_id"]
    v_a = vids[-1]["id"]
    img_a = [im for im in imgs if im["video_id"] == v_a][:4]
    cat_a = next(t["category_id"] for t in gt["tracks"] if t["video_id"] == v_a)
    if other[0]["video_id"] != v_a:
        for im in img_a:
            res.append({"image_id": im["id"], "track_id": reuse_id, "category_id": cat_a,
                        "video_id": v_a, "bbox": [100.0, 100.0, 50.0, 50.0], "score": 0.77})

    # per-box scores inside one track (averaged, results.py:88-98)
    tid1 = res[len(res) // 3]["track_id"]
    k = 0
    for r in res:
        if r["track_id"] == tid1:
            r["score"] = float(np.round(0.2 + 0.05 * (k % 7), 3))
            k += 1

    # ignored GT: one track-level flag (TAO) and a few ann-level flags (LVIS)
    gt["tracks"][1]["ignore"] = 1
    for a in gt["annotations"][10:16]:
        a["ignore"] = 1

    # GT annotation id 0 (lvis_amodal matched-id quirk, eval.py:239-240)
    gt["annotations"][20]["id"] = 0

    # a GT track with two annotations on the same image: last one wins (eval.py:322-325)
    dup = copy.deepcopy(g_anns[2])
    dup["id"] = max(a["id"] for a in gt["annotations"]) + 1
    dup["bbox"] = [dup["bbox"][0] + 8.0, dup["bbox"][1], dup["bbox"][2], dup["bbox"][3]]
    gt["annotations"].append(dup)

    # a prediction in a category that is not in the category list at all
    res.append({"image_id": imgs[7]["id"], "track_id": base_tid + 402, "category_id": 777777,
                "video_id": imgs[7]["video_id"], "bbox": [1.0, 1.0, 9.0, 9.0], "score": 0.5})

    # integer-typed boxes / area in the JSON
    gt["annotations"][30]["bbox"] = [int(x) + 1 for x in gt["annotations"][30]["bbox"]]
    gt["annotations"][30]["area"] = (gt["annotations"][30]["bbox"][2]
                                     * gt["annotations"][30]["bbox"][3])
    order = rng.permutation(len(res))
    res = [res[i] for i in order]
    return gt, res


def _inset_polygon(b, rng):
    """A convex-ish polygon inside box b (x, y, w, h): corners cut by random fractions."""
    x, y, w, h = b
    c = rng.uniform(0.05, 0.35, 4)
    return [x + c[0] * w, y, x + w - c[1] * w, y, x + w, y + c[1] * h, x + w, y + h - c[2] * h,
            x + w - c[2] * w, y + h, x + c[3] * w, y + h, x, y + h - c[3] * h, x, y + c[0] * h]


def add_segmentations(gt, res, seed, dt_mode):
    """GT annotations get a ``segmentation`` in all the forms LVIS.ann_to_rle accepts
    (lvis.py:165-178): polygon, two-part polygon, list-of-boxes, uncompressed counts, compressed
    RLE.  dt_mode: "box" results keep only their bbox (results.py:50-52 derives a polygon),
    "poly" results carry polygons of their own, "rle" results carry ONLY a compressed RLE
    (results.py:58-66).  The two RLE forms are produced with the run-length codec under test
    and stored in the golden file, so the reference reads exactly the same bytes."""
    from tao_amodal_b200.mask import RlePool
    rng = np.random.Generator(np.random.PCG64(seed))
    img = {i["id"]: i for i in gt["images"]}

    def rle_of(segm, im, compressed):
        pool = RlePool()
        pool.add_segmentation(segm, im["height"], im["width"])
        if compressed:
            r = pool.to_rle(0)
            return {"size": r["size"], "counts": r["counts"].decode()}
        off, cnt, hw, _, _ = pool.export()
        return {"size": [int(hw[0, 0]), int(hw[0, 1])], "counts": cnt.tolist()}

    for n, a in enumerate(gt["annotations"]):
        b = a["bbox"]
        kind = n % 6
        if kind in (0, 1):
            a["segmentation"] = [_inset_polygon(b, rng)]
        elif kind == 2:      # two parts: left and right half, cut
            x, y, w, h = b
            a["segmentation"] = [_inset_polygon([x, y, w * 0.45, h], rng),
                                 _inset_polygon([x + w * 0.55, y, w * 0.45, h], rng)]
        elif kind == 3:      # a list whose first element has 4 numbers is a list of BOXES
            x, y, w, h = b
            a["segmentation"] = [[x, y, w, h * 0.5], [x + w * 0.25, y + h * 0.5, w * 0.5, h * 0.5]]
        elif kind == 4:
            a["segmentation"] = rle_of([_inset_polygon(b, rng)], img[a["image_id"]], False)
        else:
            a["segmentation"] = rle_of([_inset_polygon(b, rng)], img[a["image_id"]], True)
    if dt_mode == "poly":
        for n, r in enumerate(res):
            if n % 3:
                r["segmentation"] = [_inset_polygon(r["bbox"], rng)]
    elif dt_mode == "rle":
        for r in res:
            r["segmentation"] = rle_of([_inset_polygon(r.pop("bbox"), rng)], img[r["image_id"]], True)
    return gt, res


def case_segm(dt_mode, seed):
    gt, res = synth.generate_named("tiny", seed=seed)
    gt, res = gt.to_dict(), res.to_list()
    return add_segmentations(gt, res, seed + 1, dt_mode)


SEGM_CASES = {
    "segm_box": lambda: case_segm("box", 31),
    "segm_poly": lambda: case_segm("poly", 32),
    "segm_rle": lambda: case_segm("rle", 33),
}

CASES = {
    "tiny": case_tiny,
    "small": case_small,
    "small_sparse_ids": case_small_sparse_ids,
    "small_ties": case_small_ties,
    "small_float": case_small_float,
    "edge_mix": case_edge_mix,
}


def build(name):
    gt, res = (CASES.get(name) or SEGM_CASES[name])()
    return gt, res
