"""``LVISResults`` — prediction index of the frame evaluator (mirror of
tao_amodal/evaluation/lvis_amodal/results.py:9-89; results must carry ``bbox``, optionally with
their own ``segmentation`` for iou_type="segm").  Predictions are kept
as columns; the reference-shaped ``dataset`` is materialised on demand."""
from __future__ import annotations

import copy
import logging
from collections import defaultdict

import numpy as np

from ... import ingest
from ...columnar import DtColumns
from .._common import load_json
from .lvis import LVIS


class LVISResults(LVIS):
    def __init__(self, lvis_gt, results, max_dets=300):
        if isinstance(lvis_gt, LVIS):
            self._gt = lvis_gt
        elif isinstance(lvis_gt, str):
            self._gt = LVIS(lvis_gt)
        else:
            raise TypeError("Unsupported type {} of lvis_gt.".format(lvis_gt))
        self.logger = logging.getLogger(__name__)
        self.logger.info("Loading and preparing results.")
        self.max_dets = max_dets
        self._results_path = None
        self._has_segm = None
        self._mask_only = False
        from_file = isinstance(results, str)
        dt = None
        if isinstance(results, DtColumns):
            self._result_anns, dt, self._has_segm = None, results, False
        elif from_file:
            self._results_path, self._result_anns = results, None
            try:
                dt = ingest.load_dt(results)       # native single-pass reader (bbox results)
            except KeyError:                       # mask-only results file: the dict path below
                results, self._results_path = load_json(results), None
        if dt is None:
            if not from_file:
                self.logger.warn("Assuming user provided the results in correct format.")
            assert isinstance(results, list), "results is not a list."
            self._mask_only = bool(len(results)) and "bbox" not in results[0]
            if self._mask_only:
                self._box_and_area_from_masks(results)
            self._result_anns = results
            self._has_segm = any("segmentation" in r for r in results)
            dt = DtColumns.from_list(results)
        if dt.n() == 0:
            raise IndexError("list index out of range")                # results.py:42
        self.columns = self._gt.columns
        self.dt_columns = dt
        if not np.isin(dt.image_id, np.unique(self.columns.img_id)).all():   # results.py:67-71
            raise AssertionError("Results do not correspond to current LVIS set.")
        self._indexed = False

    @property
    def dataset(self):
        if "_dataset" not in self.__dict__:
            ds = copy.deepcopy(self._gt.dataset)
            if self._result_anns is not None:
                anns = self._result_anns
            elif self._results_path is not None:
                anns = load_json(self._results_path)
            else:
                anns = self.dt_columns.to_list()
            if self.max_dets >= 0:
                anns = self.limit_dets_per_image(anns, self.max_dets)
            if self._has_segm is None:
                self._has_segm = any("segmentation" in r for r in anns)
            for n, r in enumerate(anns):
                x1, y1, w, h = r["bbox"]
                if "segmentation" not in r:
                    r["segmentation"] = [[x1, y1, x1, y1 + h, x1 + w, y1 + h, x1 + w, y1]]
                if not self._mask_only:
                    r["area"] = w * h
                r["id"] = n + 1
            ds["annotations"] = anns
            self.__dict__["_dataset"] = ds
        return self.__dict__["_dataset"]

    @staticmethod
    def _box_and_area_from_masks(results):
        """results.py:58-66 — results that carry only a compressed RLE: ``area`` is the mask
        area, ``bbox`` the box of the mask (pycocotools area / toBbox, here the native codec)."""
        from ...mask import RlePool
        pool = RlePool()
        for r in results:
            seg = r["segmentation"]
            pool.add_string(seg["counts"], seg["size"][0], seg["size"][1])
        _, _, _, bb, area = pool.export()
        for k, r in enumerate(results):
            r["area"] = area[k]
            if "bbox" not in r:
                r["bbox"] = bb[k].copy()
        pool.close()

    def given_segmentations(self):
        """True when some result came with a ``segmentation`` of its own; None when they all get
        the box polygon of results.py:50-52."""
        if self._has_segm is None:
            self.dataset            # a results FILE: the dict form records it while it is built
        return True if self._has_segm else None

    def limit_dets_per_image(self, anns, max_dets):
        """results.py:73-84."""
        per_img = defaultdict(list)
        for r in anns:
            per_img[r["image_id"]].append(r)
        for k, lst in per_img.items():
            if len(lst) > max_dets:
                per_img[k] = sorted(lst, key=lambda r: r["score"], reverse=True)[:max_dets]
        return [r for lst in per_img.values() for r in lst]

    def get_top_results(self, img_id, score_thrs):
        """results.py:86-89."""
        anns = self.load_anns(self.get_ann_ids(img_ids=[img_id]))
        return [a for a in anns if a["score"] > score_thrs]
