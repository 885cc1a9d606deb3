"""Columnar (struct-of-arrays) containers for TAO-Amodal annotations and predictions.

The reference keeps every annotation as a Python dict inside dict-of-list
indices (tao_amodal/evaluation/tao_amodal/tao.py:108-160,
lvis_amodal/lvis.py:37-61).  Everything downstream of JSON parsing in this
repo works on flat numpy columns instead, so the host prep can be vectorised
and the arrays can be handed to the CUDA library as plain pointers.

A *ragged* column (per-video / per-image category lists) is the pair
``(offsets int64[n+1], values int64[m])``.
"""
from __future__ import annotations

from dataclasses import dataclass, field, fields
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

Ragged = Tuple[np.ndarray, np.ndarray]

FREQ_CODE = {"r": 0, "c": 1, "f": 2}
FREQ_MISSING = 255


def ragged_from_lists(lists: Iterable[Sequence[int]]) -> Ragged:
    lens = []
    vals: List[int] = []
    for l in lists:
        lens.append(len(l))
        vals.extend(l)
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    if lens:
        np.cumsum(np.asarray(lens, dtype=np.int64), out=off[1:])
    return off, np.asarray(vals, dtype=np.int64)


def ragged_to_lists(r: Ragged) -> List[List[int]]:
    off, vals = r
    v = vals.tolist()
    o = off.tolist()
    return [v[o[i]:o[i + 1]] for i in range(len(o) - 1)]


def ragged_empty(n: int) -> Ragged:
    return np.zeros(n + 1, dtype=np.int64), np.zeros(0, dtype=np.int64)


def _i64(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.int64))


def _f64(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def _u8(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint8))


@dataclass
class GtColumns:
    """Ground-truth annotation file as columns (tao.py:4-60 schema)."""

    # images (dataset order)
    img_id: np.ndarray
    img_video_id: np.ndarray
    img_frame_index: np.ndarray
    img_neg: Ragged            # image-level neg_category_ids (LVIS path)
    img_nel: Ragged            # image-level not_exhaustive_category_ids
    # videos (dataset order)
    vid_id: np.ndarray
    vid_neg: Ragged
    vid_nel: Ragged
    # tracks (dataset order)
    trk_id: np.ndarray
    trk_category_id: np.ndarray
    trk_video_id: np.ndarray
    trk_ignore: np.ndarray     # uint8; "ignore" key of the track dict (eval.py:202-204)
    # categories (dataset order)
    cat_id: np.ndarray
    cat_freq: np.ndarray       # uint8 code of "frequency" (lvis_amodal/eval.py:107-113)
    merge_map: Dict[int, int]  # tao.py:97-106
    # annotations (dataset order)
    ann_id: np.ndarray
    ann_image_id: np.ndarray
    ann_track_id: np.ndarray
    ann_category_id: np.ndarray
    ann_bbox: np.ndarray       # f64 [N,4] x,y,w,h
    ann_area: np.ndarray       # f64 (the JSON "area", NOT w*h)
    ann_visibility: np.ndarray  # f64, NaN when the key is absent
    ann_oof: np.ndarray        # uint8 out_of_frame (2 = key absent)
    ann_ignore: np.ndarray     # uint8 "ignore" key of the ann (lvis_amodal/eval.py:76-78)
    # flags describing optional keys
    has_image_lists: bool = True   # images carry neg/not_exhaustive lists
    has_video_lists: bool = True

    def n_anns(self) -> int:
        return int(self.ann_id.shape[0])

    # ------------------------------------------------------------------ io
    def save_npz(self, path: str) -> None:
        d = {}
        for f in fields(self):
            v = getattr(self, f.name)
            if isinstance(v, tuple):
                d[f.name + "__off"], d[f.name + "__val"] = v
            elif isinstance(v, dict):
                d[f.name + "__k"] = _i64(list(v.keys()))
                d[f.name + "__v"] = _i64(list(v.values()))
            elif isinstance(v, bool):
                d[f.name + "__b"] = np.asarray([v])
            else:
                d[f.name] = v
        np.savez(path, **d)

    @classmethod
    def load_npz(cls, path: str) -> "GtColumns":
        z = np.load(path)
        kw = {}
        for f in fields(cls):
            n = f.name
            if n + "__off" in z:
                kw[n] = (z[n + "__off"], z[n + "__val"])
            elif n + "__k" in z:
                kw[n] = dict(zip(z[n + "__k"].tolist(), z[n + "__v"].tolist()))
            elif n + "__b" in z:
                kw[n] = bool(z[n + "__b"][0])
            else:
                kw[n] = z[n]
        return cls(**kw)

    # ----------------------------------------------------- dict conversion
    @classmethod
    def from_dict(cls, ds: dict) -> "GtColumns":
        """Build columns from a parsed annotation JSON (reference schema).

        Mirrors what Tao._create_index reads (tao.py:108-160): the merge map,
        the float conversion of bbox (:142).  Key errors for missing required
        keys surface as KeyError exactly like the dict accesses they replace.
        """
        images = ds["images"]
        videos = ds.get("videos", [])
        tracks = ds.get("tracks", [])
        cats = ds["categories"]
        anns = ds["annotations"]

        merge_map: Dict[int, int] = {}
        for c in cats:
            if "merged" in c:
                for m in c["merged"]:
                    merge_map[m["id"]] = c["id"]

        has_img_lists = all("neg_category_ids" in im for im in images)
        has_vid_lists = all("neg_category_ids" in v for v in videos)
        return cls(
            img_id=_i64([im["id"] for im in images]),
            img_video_id=_i64([im.get("video_id", -1) for im in images]),
            img_frame_index=_i64([im.get("frame_index", 0) for im in images]),
            img_neg=(ragged_from_lists(im["neg_category_ids"] for im in images)
                     if has_img_lists else ragged_empty(len(images))),
            img_nel=(ragged_from_lists(im["not_exhaustive_category_ids"] for im in images)
                     if has_img_lists else ragged_empty(len(images))),
            vid_id=_i64([v["id"] for v in videos]),
            vid_neg=(ragged_from_lists(v["neg_category_ids"] for v in videos)
                     if has_vid_lists else ragged_empty(len(videos))),
            vid_nel=(ragged_from_lists(v["not_exhaustive_category_ids"] for v in videos)
                     if has_vid_lists else ragged_empty(len(videos))),
            trk_id=_i64([t["id"] for t in tracks]),
            trk_category_id=_i64([t["category_id"] for t in tracks]),
            trk_video_id=_i64([t["video_id"] for t in tracks]),
            trk_ignore=_u8([1 if t.get("ignore", 0) else 0 for t in tracks]),
            cat_id=_i64([c["id"] for c in cats]),
            cat_freq=_u8([FREQ_CODE.get(c.get("frequency"), FREQ_MISSING) for c in cats]),
            merge_map=merge_map,
            ann_id=_i64([a["id"] for a in anns]),
            ann_image_id=_i64([a["image_id"] for a in anns]),
            ann_track_id=_i64([a.get("track_id", -1) for a in anns]),
            ann_category_id=_i64([a["category_id"] for a in anns]),
            ann_bbox=_f64([a["bbox"] for a in anns]).reshape(-1, 4),
            ann_area=_f64([a["area"] for a in anns]),
            ann_visibility=_f64([a.get("visibility", np.nan) for a in anns]),
            ann_oof=_u8([(1 if a["out_of_frame"] else 0) if "out_of_frame" in a else 2
                         for a in anns]),
            ann_ignore=_u8([1 if a.get("ignore", 0) else 0 for a in anns]),
            has_image_lists=has_img_lists,
            has_video_lists=has_vid_lists,
        )

    def to_dict(self) -> dict:
        """Emit the reference's annotation-JSON structure (tao.py:4-60)."""
        img_neg = ragged_to_lists(self.img_neg)
        img_nel = ragged_to_lists(self.img_nel)
        vid_neg = ragged_to_lists(self.vid_neg)
        vid_nel = ragged_to_lists(self.vid_nel)
        inv_freq = {v: k for k, v in FREQ_CODE.items()}
        images = []
        for k, (i, v, f) in enumerate(zip(self.img_id.tolist(), self.img_video_id.tolist(),
                                          self.img_frame_index.tolist())):
            im = {"id": i, "video_id": v, "frame_index": f, "width": 1280, "height": 720,
                  "file_name": "v%d/f%06d.jpg" % (v, f)}
            if self.has_image_lists:
                im["neg_category_ids"] = img_neg[k]
                im["not_exhaustive_category_ids"] = img_nel[k]
            images.append(im)
        videos = []
        for k, v in enumerate(self.vid_id.tolist()):
            vd = {"id": v, "name": "v%d" % v, "width": 1280, "height": 720, "metadata": {}}
            if self.has_video_lists:
                vd["neg_category_ids"] = vid_neg[k]
                vd["not_exhaustive_category_ids"] = vid_nel[k]
            videos.append(vd)
        tracks = []
        for i, c, v, ig in zip(self.trk_id.tolist(), self.trk_category_id.tolist(),
                               self.trk_video_id.tolist(), self.trk_ignore.tolist()):
            t = {"id": i, "category_id": c, "video_id": v}
            if ig:
                t["ignore"] = 1
            tracks.append(t)
        cats = []
        for i, f in zip(self.cat_id.tolist(), self.cat_freq.tolist()):
            c = {"id": i, "name": "c%d" % i, "synset": "unknown"}
            if f != FREQ_MISSING:
                c["frequency"] = inv_freq[f]
            cats.append(c)
        merged: Dict[int, List[int]] = {}
        for src, dst in self.merge_map.items():
            merged.setdefault(dst, []).append(src)
        for c in cats:
            if c["id"] in merged:
                c["merged"] = [{"id": s} for s in merged[c["id"]]]
        anns = []
        bb = self.ann_bbox.tolist()
        vis = self.ann_visibility.tolist()
        for k, (i, im, t, c, a, o, ig) in enumerate(zip(
                self.ann_id.tolist(), self.ann_image_id.tolist(), self.ann_track_id.tolist(),
                self.ann_category_id.tolist(), self.ann_area.tolist(), self.ann_oof.tolist(),
                self.ann_ignore.tolist())):
            d = {"id": i, "image_id": im, "track_id": t, "category_id": c,
                 "bbox": bb[k], "area": a}
            if vis[k] == vis[k]:
                d["visibility"] = vis[k]
            if o != 2:
                d["out_of_frame"] = bool(o)
            if ig:
                d["ignore"] = 1
            anns.append(d)
        return {"info": {}, "images": images, "videos": videos, "tracks": tracks,
                "annotations": anns, "categories": cats, "licenses": []}


@dataclass
class DtColumns:
    """Prediction file as columns (README.md:108-115: image_id, category_id,
    bbox, score, track_id, video_id)."""

    image_id: np.ndarray
    track_id: np.ndarray
    category_id: np.ndarray
    video_id: np.ndarray
    bbox: np.ndarray     # f64 [N,4]
    score: np.ndarray    # f64
    # number of result objects that lack the key (their column entry is -1).  The frame
    # evaluator never reads these keys; the track path raises KeyError like the reference
    # (tools/eval_on_tao_amodal.py:57, tao_amodal/results.py:71) — prep.require_track_keys.
    missing_track_id: int = 0
    missing_video_id: int = 0

    _ARRAYS = ("image_id", "track_id", "category_id", "video_id", "bbox", "score")

    def n(self) -> int:
        return int(self.image_id.shape[0])

    def save_npz(self, path: str) -> None:
        np.savez(path, **{k: getattr(self, k) for k in self._ARRAYS})

    @classmethod
    def load_npz(cls, path: str) -> "DtColumns":
        z = np.load(path)
        return cls(**{k: z[k] for k in cls._ARRAYS})

    @classmethod
    def from_list(cls, results: list) -> "DtColumns":
        if not isinstance(results, list):
            raise AssertionError("results is not a list.")  # results.py:52
        return cls(
            image_id=_i64([r["image_id"] for r in results]),
            track_id=_i64([r.get("track_id", -1) for r in results]),
            category_id=_i64([r["category_id"] for r in results]),
            video_id=_i64([r.get("video_id", -1) for r in results]),
            bbox=_f64([r["bbox"] for r in results]).reshape(-1, 4),
            score=_f64([r["score"] for r in results]),
            missing_track_id=sum(1 for r in results if "track_id" not in r),
            missing_video_id=sum(1 for r in results if "video_id" not in r),
        )

    def to_list(self) -> list:
        bb = self.bbox.tolist()
        return [
            {"image_id": i, "track_id": t, "category_id": c, "video_id": v,
             "bbox": bb[k], "score": s}
            for k, (i, t, c, v, s) in enumerate(zip(
                self.image_id.tolist(), self.track_id.tolist(), self.category_id.tolist(),
                self.video_id.tolist(), self.score.tolist()))
        ]

    def copy(self) -> "DtColumns":
        return DtColumns(**{k: getattr(self, k).copy() for k in self._ARRAYS},
                         missing_track_id=self.missing_track_id,
                         missing_video_id=self.missing_video_id)


def _ragged_take(r: Ragged, rows: np.ndarray) -> Ragged:
    off, vals = r
    lens = (off[1:] - off[:-1])[rows]
    new_off = np.zeros(rows.size + 1, dtype=np.int64)
    np.cumsum(lens, out=new_off[1:])
    idx = (np.arange(int(new_off[-1]), dtype=np.int64) - np.repeat(new_off[:-1], lens)
           + np.repeat(off[:-1][rows], lens))
    return new_off, vals[idx] if vals.size else vals


def subset_videos(gt: GtColumns, dt: DtColumns, video_ids) -> Tuple[GtColumns, DtColumns]:
    """The annotation / prediction columns restricted to some videos (categories are kept
    whole).  Used to shard a dataset across GPUs and to cut bounded CPU-baseline samples."""
    vids = np.unique(np.asarray(video_ids, dtype=np.int64))
    vrow = np.nonzero(np.isin(gt.vid_id, vids))[0]
    irow = np.nonzero(np.isin(gt.img_video_id, vids))[0]
    trow = np.nonzero(np.isin(gt.trk_video_id, vids))[0]
    arow = np.nonzero(np.isin(gt.ann_image_id, gt.img_id[irow]))[0]
    g = GtColumns(
        img_id=gt.img_id[irow], img_video_id=gt.img_video_id[irow],
        img_frame_index=gt.img_frame_index[irow],
        img_neg=_ragged_take(gt.img_neg, irow), img_nel=_ragged_take(gt.img_nel, irow),
        vid_id=gt.vid_id[vrow], vid_neg=_ragged_take(gt.vid_neg, vrow),
        vid_nel=_ragged_take(gt.vid_nel, vrow),
        trk_id=gt.trk_id[trow], trk_category_id=gt.trk_category_id[trow],
        trk_video_id=gt.trk_video_id[trow], trk_ignore=gt.trk_ignore[trow],
        cat_id=gt.cat_id, cat_freq=gt.cat_freq, merge_map=dict(gt.merge_map),
        ann_id=gt.ann_id[arow], ann_image_id=gt.ann_image_id[arow],
        ann_track_id=gt.ann_track_id[arow], ann_category_id=gt.ann_category_id[arow],
        ann_bbox=np.ascontiguousarray(gt.ann_bbox[arow]), ann_area=gt.ann_area[arow],
        ann_visibility=gt.ann_visibility[arow], ann_oof=gt.ann_oof[arow],
        ann_ignore=gt.ann_ignore[arow],
        has_image_lists=gt.has_image_lists, has_video_lists=gt.has_video_lists)
    drow = np.nonzero(np.isin(dt.video_id, vids))[0]
    d = DtColumns(image_id=dt.image_id[drow], track_id=dt.track_id[drow],
                  category_id=dt.category_id[drow], video_id=dt.video_id[drow],
                  bbox=np.ascontiguousarray(dt.bbox[drow]), score=dt.score[drow],
                  missing_track_id=dt.missing_track_id, missing_video_id=dt.missing_video_id)
    return g, d
