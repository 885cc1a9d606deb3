"""``LVIS`` — annotation index of the frame evaluator (mirror of
tao_amodal/evaluation/lvis_amodal/lvis.py:18-205; the mask helpers :155-192 go through the
native run-length codec, tao_amodal_b200/mask.py)."""
from __future__ import annotations

import logging
from collections import defaultdict

from ... import ingest
from ...columnar import GtColumns
from .._common import load_json

_INDEX_ATTRS = ("img_ann_map", "cat_img_map", "anns", "cats", "imgs")


class LVIS:
    def __init__(self, annotation_path):
        self.logger = logging.getLogger(__name__)
        self.logger.info("Loading annotations.")
        if isinstance(annotation_path, dict):
            self.dataset = annotation_path
            self.columns = GtColumns.from_dict(self.dataset)
        else:
            # native single-pass reader; ``self.dataset`` is parsed only when read
            self._path = annotation_path
            self.columns = ingest.load_gt(annotation_path)
        self._indexed = False

    def __getattr__(self, name):
        if name == "dataset" and "_path" in self.__dict__:
            ds = load_json(self.__dict__["_path"])
            assert type(ds) == dict, "Annotation file format {} not supported.".format(type(ds))
            self.__dict__["dataset"] = ds
            return ds
        if name in _INDEX_ATTRS and not self.__dict__.get("_indexed", True):
            self._create_index()
            return self.__dict__[name]
        raise AttributeError(name)

    def _create_index(self):
        """lvis.py:38-61, on first use."""
        self.logger.info("Creating index.")
        self._indexed = True
        self.img_ann_map, self.cat_img_map = defaultdict(list), defaultdict(list)
        self.anns, self.cats, self.imgs = {}, {}, {}
        for ann in self.dataset["annotations"]:
            self.img_ann_map[ann["image_id"]].append(ann)
            self.anns[ann["id"]] = ann
            self.cat_img_map[ann["category_id"]].append(ann["image_id"])
        for img in self.dataset["images"]:
            self.imgs[img["id"]] = img
        for cat in self.dataset["categories"]:
            self.cats[cat["id"]] = cat
        self.logger.info("Index created.")

    def get_ann_ids(self, img_ids=None, cat_ids=None, area_rng=None):
        """lvis.py:63-97."""
        if img_ids is None:
            pool = self.dataset["annotations"]
        else:
            pool = [a for i in img_ids for a in self.img_ann_map[i]]
        if cat_ids is None and area_rng is None:
            return [a["id"] for a in pool]
        cats = set(cat_ids)
        lo, hi = (0, float("inf")) if area_rng is None else area_rng
        return [a["id"] for a in pool
                if a["category_id"] in cats and a["area"] > lo and a["area"] < hi]

    def get_cat_ids(self):
        return self.columns.cat_id.tolist() if not self._indexed else list(self.cats.keys())

    def get_img_ids(self):
        return self.columns.img_id.tolist() if not self._indexed else list(self.imgs.keys())

    @staticmethod
    def _load_helper(_dict, ids):
        if ids is None:
            return list(_dict.values())
        return [_dict[i] for i in ids]

    def load_anns(self, ids=None):
        return self._load_helper(self.anns, ids)

    def load_cats(self, ids):
        return self._load_helper(self.cats, ids)

    def load_imgs(self, ids):
        return self._load_helper(self.imgs, ids)

    def ann_to_rle(self, ann):
        """lvis.py:155-178: polygons / uncompressed counts / compressed RLE -> compressed RLE
        dict ({"size": [h, w], "counts": bytes}), through the native codec."""
        from ...mask import RlePool
        segm = ann["segmentation"]
        if not isinstance(segm, list) and not isinstance(segm["counts"], list):
            return segm
        img = self.imgs[ann["image_id"]]
        pool = RlePool()
        pool.add_segmentation(segm, img["height"], img["width"])
        return pool.to_rle(0)

    def ann_to_mask(self, ann):
        """lvis.py:180-192: binary mask uint8 [h, w] of the annotation."""
        import numpy as np
        from ...mask import RlePool
        pool = RlePool()
        segm = ann["segmentation"]
        img = self.imgs[ann["image_id"]]
        pool.add_segmentation(segm, img["height"], img["width"])
        off, cnt, hw, _, _ = pool.export()
        vals = np.arange(cnt.size, dtype=np.uint8) & 1
        return np.repeat(vals, cnt).reshape((int(hw[0, 1]), int(hw[0, 0]))).T.copy()
