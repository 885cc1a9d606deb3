"""``TaoResults`` — prediction index of the track evaluator.

Mirror of tao_amodal/evaluation/tao_amodal/results.py:11-140.  The reference deep-copies
the whole ground-truth dataset and rewrites every result dict; here the predictions live
in columns (``self.dt_columns``) and the reference-shaped ``dataset`` / dict indices are
produced only when somebody asks for them.  All validation the reference performs in its
constructor (unique track ids per video, results' images known to the GT file) happens in
the constructor here as well, with the same exception types and messages.
"""
from __future__ import annotations

import copy
import logging
from collections import defaultdict

import numpy as np

from ... import ingest
from ...columnar import DtColumns
from ... import prep
from .._common import load_json
from .tao import Tao


class TaoResults(Tao):
    def __init__(self, tao_gt, results, max_dets=300):
        if isinstance(tao_gt, Tao):
            self._gt = tao_gt
        elif isinstance(tao_gt, str):
            self._gt = Tao(tao_gt)
        else:
            raise TypeError("Unsupported type {} of tao_gt.".format(tao_gt))
        self.logger = logging.getLogger('tao.results')
        self.max_dets = max_dets
        self._results_path = None
        if isinstance(results, DtColumns):
            self._result_anns, dt = None, results
        elif isinstance(results, str):
            # native single-pass reader; the list-of-dicts form is parsed only if
            # ``dataset`` is read
            self._results_path, self._result_anns = results, None
            dt = ingest.load_dt(results)
        else:
            self.logger.warn("Assuming results file is a list of dicts.")   # results.py:40
            assert isinstance(results, list), "results is not a list."
            self._result_anns = results
            dt = DtColumns.from_list(results)
        self.merge_map = dict(self._gt.merge_map)
        self.columns = self._gt.columns
        self.dt_columns = dt
        self._check()
        self._indexed = False

    def _check(self):
        dt = self.dt_columns
        if dt.n() == 0:
            raise IndexError("list index out of range")               # results.py:63
        tu, tfirst = np.unique(dt.track_id, return_index=True)        # results.py:111-119
        first_vid = dt.video_id[tfirst][np.searchsorted(tu, dt.track_id)]
        bad = np.nonzero(dt.video_id != first_vid)[0]
        if bad.size:
            t = int(dt.track_id[bad[0]])
            raise AssertionError(
                f'Track id {t} appears in more than one video: '
                f'{int(first_vid[bad[0]])} and {int(dt.video_id[bad[0]])}')
        known = np.unique(self.columns.img_id)                        # results.py:105-109
        if not np.isin(dt.image_id, known).all():
            raise AssertionError("Results do not correspond to current Tao set.")

    @property
    def dataset(self):
        """The reference's merged dataset dict (results.py:28-87), materialised on demand."""
        if "_dataset" not in self.__dict__:
            self.__dict__["_dataset"] = self._materialise()
        return self.__dict__["_dataset"]

    def _materialise(self):
        ds = copy.deepcopy(self._gt.dataset)
        if self._result_anns is not None:
            anns = self._result_anns
        elif self._results_path is not None:
            anns = load_json(self._results_path)
        else:
            anns = self.dt_columns.to_list()
        mm = self.merge_map
        for r in anns:
            if r["category_id"] in mm:
                r["category_id"] = mm[r["category_id"]]
        anns = self.limit_dets_per_image(anns, self.max_dets) if self.max_dets >= 0 else anns
        tracks = {}
        for n, r in enumerate(anns):
            w, h = r["bbox"][2], r["bbox"][3]
            t = r["track_id"]
            if t not in tracks:
                tracks[t] = {"id": t, "video_id": r["video_id"], "category_id": r["category_id"]}
            r["area"] = w * h
            r["id"] = n + 1
        ds["annotations"] = anns
        ds["tracks"] = list(tracks.values())
        # track scores: mean only when the per-box scores differ (results.py:88-98)
        per_track = defaultdict(list)
        for r in anns:
            per_track[r["track_id"]].append(r)
        for t, lst in per_track.items():
            sc = [float(r["score"]) for r in lst]
            if len(set(sc)) > 1:
                avg = np.mean(sc)
                tracks[t]["score"] = avg
                for r in lst:
                    r["score"] = avg
            else:
                tracks[t]["score"] = sc[0]
        return ds

    def ensure_unique_track_ids(self, result_anns):
        """results.py:111-119."""
        seen = {}
        for r in result_anns:
            t = r['track_id']
            seen.setdefault(t, r['video_id'])
            assert r['video_id'] == seen[t], (
                f'Track id {t} appears in more than one video: {seen[t]} and {r["video_id"]}')

    def limit_dets_per_image(self, anns, max_dets):
        """results.py:121-132: per image (first-appearance order) the top max_dets by score."""
        per_img = defaultdict(list)
        for r in anns:
            per_img[r["image_id"]].append(r)
        for k, lst in per_img.items():
            if len(lst) > max_dets:
                per_img[k] = sorted(lst, key=lambda r: r["score"], reverse=True)[:max_dets]
        return [r for lst in per_img.values() for r in lst]

    def get_top_results(self, img_id, score_thrs):
        raise NotImplementedError('Unclear if this should be per image or per video')   # results.py:134-136
