"""``Tao`` — annotation index of the track evaluator.

Mirror of the reference class (tao_amodal/evaluation/tao_amodal/tao.py:68-341): same
constructor, same public dict attributes and accessors.  The evaluation itself reads the
columnar form (``self.columns``); the reference-style dict indices (``anns``, ``imgs``,
``img_ann_map`` ...) are built on first use only.
"""
from __future__ import annotations

import logging
from collections import defaultdict

from ... import ingest
from ...columnar import GtColumns
from .._common import load_json

_INDEX_ATTRS = ("vids", "tracks", "cats", "imgs", "anns", "vid_img_map", "vid_track_map",
                "img_ann_map", "cat_img_map", "track_ann_map", "negative_categories")


class Tao:
    def __init__(self, annotation_path, logger=None):
        if not logger:
            self.logger = logging.getLogger('tao.tao')
        elif isinstance(logger, str):
            self.logger = logging.getLogger(logger)
        else:
            self.logger = logger
        self.logger.info("Loading annotations.")
        if isinstance(annotation_path, dict):
            for key in ('info', 'images', 'annotations', 'categories', 'videos', 'tracks'):
                assert key in annotation_path, (
                    f'Provided dictionary does not contain key {key}')
            self.dataset = annotation_path
            assert type(self.dataset) == dict, (
                "Annotation file format {} not supported.".format(type(self.dataset)))
            self.columns = GtColumns.from_dict(self.dataset)
        else:
            # native single-pass reader (csrc/ta_json.cpp); the dict form of the file
            # (``self.dataset``) is only parsed if somebody reads it
            self._path = annotation_path
            self.columns = ingest.load_gt(annotation_path, need_videos_tracks=True)
        self._init_columns()

    # -------------------------------------------------------------------------------- columns
    def _init_columns(self):
        self.merge_map = dict(self.columns.merge_map)
        if not self.merge_map:
            logging.error('Did not merge any categories.')      # tao.py:104-105
        self._indexed = False

    def __getattr__(self, name):
        if name == "dataset" and "_path" in self.__dict__:
            ds = load_json(self.__dict__["_path"])
            assert type(ds) == dict, "Annotation file format {} not supported.".format(type(ds))
            self.__dict__["dataset"] = ds
            return ds
        # dict indices of tao.py:108-160, built lazily
        if name in _INDEX_ATTRS and not self.__dict__.get("_indexed", True):
            self._create_index()
            return self.__dict__[name]
        raise AttributeError(name)

    def _create_index(self):
        """tao.py:108-160, including its in-place normalisation of the dataset dicts
        (merged category ids, float boxes)."""
        self._indexed = True
        ds = self.dataset
        self.vids = {x['id']: x for x in ds['videos']}
        self.tracks = {x['id']: x for x in ds['tracks']}
        self.cats = {x['id']: x for x in ds['categories']}
        self.imgs, self.anns = {}, {}
        self.vid_img_map, self.vid_track_map = defaultdict(list), defaultdict(list)
        self.img_ann_map, self.cat_img_map = defaultdict(list), defaultdict(list)
        self.track_ann_map = defaultdict(list)
        self.negative_categories = {}
        mm = self.merge_map
        for x in ds['annotations'] + ds['tracks']:
            if x['category_id'] in mm:
                x['category_id'] = mm[x['category_id']]
        for img in ds['images']:
            self.imgs[img['id']] = img
            self.vid_img_map[img['video_id']].append(img)
        for t in ds['tracks']:
            self.vid_track_map[t['video_id']].append(t)
        for ann in ds['annotations']:
            ann['bbox'] = [float(x) for x in ann['bbox']]
            assert ann['category_id'] == self.tracks[ann['track_id']]['category_id'], \
                'Annotation category differs from its track category'
            self.track_ann_map[ann['track_id']].append(ann)
            self.img_ann_map[ann['image_id']].append(ann)
            self.cat_img_map[ann['category_id']].append(ann['image_id'])
            self.anns[ann['id']] = ann

    # -------------------------------------------------------------------------------- accessors
    def get_track_ids(self, img_ids=None, cat_ids=None):
        if img_ids is not None:
            raise NotImplementedError("Searching track ids by image ids not yet supported")
        if cat_ids is None:
            return list(self.tracks.keys())
        cats = set(cat_ids)
        return [t['id'] for t in self.tracks.values() if t['category_id'] in cats]

    def group_ann_tracks(self, anns):
        """tao.py:172-188."""
        out = {}
        for a in anns:
            tid = a['track_id']
            if tid not in out:
                out[tid] = dict(self.tracks[tid])
                out[tid]['annotations'] = []
            out[tid]['annotations'].append(a)
        for tr in out.values():
            tr['annotations'] = sorted(
                tr['annotations'], key=lambda x: self.imgs[x['image_id']]['frame_index'])
            tr['area'] = sum(x['area'] for x in tr['annotations']) / len(tr['annotations'])
        return list(out.values())

    def get_kth_annotation(self, track_id, k):
        """tao.py:198-201."""
        return sorted(self.track_ann_map[track_id],
                      key=lambda x: self.imgs[x['image_id']]['frame_index'])[k]

    def get_single_object_init(self, track_id, init_type='first'):
        """tao.py:190-196."""
        if init_type == 'first':
            return self.get_kth_annotation(track_id, k=0)
        raise NotImplementedError(f'Unsupported init type, {init_type}')

    def get_ann_ids(self, vid_ids=None, img_ids=None, cat_ids=None, area_rng=None):
        """tao.py:203-254: same selection and ordering rules, including the CPython set
        order of ``set(img_ids) & set(video_images)`` (:230) and the strict area bounds."""
        if vid_ids is not None:
            in_videos = [im['id'] for v in vid_ids for im in self.vid_img_map[v]]
            wanted = in_videos if img_ids is None else img_ids
            img_ids = list(set(wanted) & set(in_videos))
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

    def get_vid_ids(self):
        return self.columns.vid_id.tolist() if not self._indexed else list(self.vids.keys())

    def get_img_ids(self):
        return self.columns.img_id.tolist() if not self._indexed else list(self.imgs.keys())

    @staticmethod
    def _load_helper(_dict, ids):
        if ids is None:
            return list(_dict.values())
        return [_dict[i] for i in ids]

    def load_anns(self, ids=None):
        return self._load_helper(self.anns, ids)

    def load_tracks(self, ids=None):
        return self._load_helper(self.tracks, ids)

    def load_cats(self, ids):
        return self._load_helper(self.cats, ids)

    def load_imgs(self, ids):
        return self._load_helper(self.imgs, ids)

    def load_vids(self, ids):
        return self._load_helper(self.vids, ids)
