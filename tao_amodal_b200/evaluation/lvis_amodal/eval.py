"""``LVISEval`` — visibility-split frame-AP evaluator on the CUDA library.

Mirror of tao_amodal/evaluation/lvis_amodal/eval.py:14-583 (``iou_type='bbox'`` and ``'segm'``): same
constructor, ``params``, ``evaluate / accumulate / summarize / run / print_results /
get_results`` and the public attributes ``ious``, ``eval_imgs``, ``eval``, ``results``,
``freq_groups``.  The (image, category, visibility range) Python grid of the reference is one
fused IoU + matching kernel here (engine.stage_frame_eval); with ``iou_type='segm'`` the
annotations become run-length masks (mask.RlePool, native codec) and the IoU matrices come from
the mask-IoU kernel (ta_rle_iou) followed by the generic matcher.
"""
from __future__ import annotations

import datetime
import logging
from collections import OrderedDict

import numpy as np

from ... import materialize, prep
from ...columnar import DtColumns
from .._common import (LazyDict, ascending_rec_thrs, dist_accumulate, dist_check_nonempty, dist_info, get_engine,
                       restore_rec_order)
from .lvis import LVIS
from .results import LVISResults


class Params:
    def __init__(self, iou_type):
        """Same fields and defaults as lvis_amodal/eval.py:553-583."""
        self.img_ids = []
        self.cat_ids = []
        self.iou_thrs = np.linspace(0.5, 0.95, int(np.round((0.95 - 0.5) / 0.05)) + 1,
                                    endpoint=True)
        self.rec_thrs = np.linspace(0.0, 1.00, int(np.round((1.00 - 0.0) / 0.01)) + 1,
                                    endpoint=True)
        self.max_dets = 300
        self.visibility_rng = [[0, 1.0], [0, 0.1], [0.1, 0.8], [0.8, 1.0], [0, 0.8],
                               [0, 1.0]]   # last: pseudo range for out-of-frame boxes
        self.visibility_rng_lbl = ["all", "highly-occluded", "partially-occluded",
                                   "highly-visible", "highly-and-partially-occluded",
                                   'out-of-frame']
        self.use_cats = 1
        self.img_count_lbl = ["r", "c", "f"]
        self.iou_type = iou_type


class _CellList:
    """``eval_imgs`` of the reference is a flat list indexed c*R*I + r*I + i with None for
    empty cells (lvis eval.py:140-145, :340-346); this view builds the non-empty ones lazily."""

    def __init__(self, loader, n_cat, n_rng, n_img):
        self._loader, self._cells = loader, None
        self._shape = (n_cat, n_rng, n_img)

    def _get(self):
        if self._cells is None:
            self._cells = self._loader()
        return self._cells

    def __len__(self):
        c, r, i = self._shape
        return c * r * i

    def __getitem__(self, flat):
        c, r, i = self._shape
        if flat < 0:
            flat += len(self)
        if not 0 <= flat < len(self):
            raise IndexError("list index out of range")
        ci, rem = divmod(flat, r * i)
        ri, ii = divmod(rem, i)
        return self._get().get((ci, ri, ii))

    def __iter__(self):
        for k in range(len(self)):
            yield self[k]

    def __bool__(self):
        return True


class LVISEval:
    def __init__(self, lvis_gt, lvis_dt, iou_type="segm", device=0):
        self.logger = logging.getLogger(__name__)
        if iou_type not in ["bbox", "segm"]:
            raise ValueError("iou_type: {} is not supported.".format(iou_type))
        if isinstance(lvis_gt, LVIS):
            self.lvis_gt = lvis_gt
        elif isinstance(lvis_gt, str):
            self.lvis_gt = LVIS(lvis_gt)
        else:
            raise TypeError("Unsupported type {} of lvis_gt.".format(lvis_gt))
        if isinstance(lvis_dt, LVISResults):
            self.lvis_dt = lvis_dt
        elif isinstance(lvis_dt, (str, list, DtColumns)):
            self.lvis_dt = LVISResults(self.lvis_gt, lvis_dt)
        else:
            raise TypeError("Unsupported type {} of lvis_dt.".format(lvis_dt))
        self.eval_imgs = []
        self.eval = {}
        self.params = Params(iou_type=iou_type)
        self.results = OrderedDict()
        self.ious = {}
        self.params.img_ids = sorted(self.lvis_gt.get_img_ids())
        self.params.cat_ids = sorted(self.lvis_gt.get_cat_ids())
        self.device = device
        self._plan = self._dev = self._detail = self._host_out = None
        self._rank, self._world = 0, 1

    def _prepare(self):
        """Columnar equivalent of lvis eval.py:59-113 (prep.prepare_lvis)."""
        p = self.params
        if p.iou_type not in ("bbox", "segm"):
            raise ValueError("Unknown iou_type for iou computation.")        # lvis eval.py:184
        if len(p.iou_thrs) > 16:
            raise ValueError("at most 16 IoU thresholds are supported")
        img_ids = p.img_ids
        self._rank, self._world = dist_info()
        if self._world > 1:
            # images travel with their video (same partition as the track evaluator); files
            # without video ids are split by image id
            from ... import parallel
            cols = self.lvis_gt.columns
            sel = np.isin(cols.img_id, np.asarray(img_ids, dtype=np.int64))
            vid = cols.img_video_id[sel]
            if (vid >= 0).all() and vid.size:
                mine = parallel.shard_videos(np.unique(vid), self._world)[self._rank]
                img_ids = cols.img_id[sel][np.isin(vid, mine)]
            else:
                img_ids = np.array_split(np.unique(cols.img_id[sel]), self._world)[self._rank]
        self._plan = prep.prepare_lvis(
            self.lvis_gt.columns, self.lvis_dt.dt_columns, max_dets=self.lvis_dt.max_dets,
            vis_rng=p.visibility_rng, img_ids=img_ids,
            cat_ids=p.cat_ids if p.cat_ids else None, use_cats=bool(p.use_cats),
            allow_empty=self._world > 1)
        self.freq_groups = self._plan.freq_groups
        if p.iou_type == "segm":
            self._plan.masks = self._build_masks(self._plan)

    def _build_masks(self, plan):
        """``_to_mask`` of lvis eval.py:54-57, :70-72 for the entities of the plan: every GT /
        detection annotation through LVIS.ann_to_rle's conversion (lvis.py:155-178), kept as
        flat run-length arrays for the device.  Detections that carry no ``segmentation`` get
        the 4-corner polygon of their box (lvis_amodal/results.py:50-52), converted in one
        batched call."""
        from ...mask import RlePool
        gt_anns, imgs = self.lvis_gt.anns, self.lvis_gt.imgs
        gpool = RlePool()
        for aid in plan.gt_id.tolist():
            ann = gt_anns[aid]
            img = imgs[ann["image_id"]]
            gpool.add_segmentation(ann["segmentation"], img["height"], img["width"])
        dpool = RlePool()
        given = self.lvis_dt.given_segmentations()
        if given is None:
            img_of = np.asarray(plan.unit_ids)[plan.grp_unit[np.repeat(
                np.arange(plan.n_groups), np.diff(plan.grp_dt_off))]]
            hh = np.asarray([imgs[i]["height"] for i in img_of.tolist()], dtype=np.int64)
            ww = np.asarray([imgs[i]["width"] for i in img_of.tolist()], dtype=np.int64)
            if plan.n_dt:
                dpool.add_boxes(plan.dt_box, hh, ww)
        else:
            dt_anns = self.lvis_dt.dataset["annotations"]
            for aid in plan.dt_id.tolist():
                ann = dt_anns[aid - 1]                       # ids are 1 + position, results.py:56
                img = imgs[ann["image_id"]]
                dpool.add_segmentation(ann["segmentation"], img["height"], img["width"])
        out = {}
        for side, pool in (("dt", dpool), ("gt", gpool)):
            off, cnt, hw, bb, _ = pool.export()
            out[side] = (off, np.ascontiguousarray(cnt if cnt.size else np.zeros(1, np.uint32)),
                         np.ascontiguousarray(hw.reshape(-1, 2) if hw.size else np.zeros((1, 2), np.uint32)),
                         np.ascontiguousarray(bb.reshape(-1, 4) if bb.size else np.zeros((1, 4))))
            pool.close()
        return out

    def evaluate(self):
        """Per-image evaluation on the GPU (lvis eval.py:115-145)."""
        self.logger.info("Running per image evaluation.")
        self.logger.info("Evaluate annotation type *{}*".format(self.params.iou_type))
        self.params.img_ids = list(np.unique(self.params.img_ids))
        self._prepare()
        eng = get_engine(self.device)
        rec_sorted, self._rec_inv = ascending_rec_thrs(self.params.rec_thrs)
        self._rec_sorted = rec_sorted
        self._detail = None
        self._dev = self._host_out = None
        if self._world == 1 and self._plan.masks is None:
            # single GPU, boxes: ONE C call from host buffers (ta_eval_plan_host) — no torch
            self._host_out = eng.evaluate_host(
                self._plan, iou_thrs=np.ascontiguousarray(self.params.iou_thrs, dtype=np.float64),
                rec_thrs=rec_sorted)
        else:
            self._dev = eng.upload(self._plan, self.params.iou_thrs, rec_sorted)
            if self._world > 1:
                dist_check_nonempty(self._dev, "Found no groundtruth annotations for given params",
                                    "list index out of range", dt_error=IndexError)
            if self._plan.masks is None:
                eng.stage_frame_eval(self._dev)
            else:
                eng.stage_iou(self._dev)
                eng.stage_match(self._dev)
        self.ious = LazyDict(lambda: materialize.iou_dict(self._plan, self._need_detail().iou))
        plan = self._plan
        self.eval_imgs = _CellList(
            lambda: materialize.cells_dict(plan, len(self.params.iou_thrs), self._need_detail()),
            len(plan.cat_ids), plan.n_cfg, len(plan.unit_ids))

    def _need_detail(self):
        if self._world > 1:
            # every rank holds the cells of its own videos only; a merged view would have to
            # gather per-cell arrays that the reference builds in one process
            raise RuntimeError("ious / eval_imgs / dt_pointers are per-process structures: "
                               "not available in multi-GPU mode (run single-GPU to read them)")
        if self._detail is None:
            # a second pass over the same plan; the accumulated tensors it rewrites are
            # identical (same plan, same kernels)
            eng = get_engine(self.device)
            if self._dev is None:
                self._dev = eng.upload(self._plan, self.params.iou_thrs, self._rec_sorted)
            self._detail = eng.evaluate_device(self._dev, detail=True)
        return self._detail

    def compute_iou(self, img_id, cat_id):
        return self.ious.get((img_id, cat_id), [])

    def accumulate(self):
        """PR accumulation on the GPU (lvis eval.py:305-426)."""
        self.logger.info("Accumulating evaluation results.")
        if self._plan is None:
            self.logger.warn("Please run evaluate first.")
            return
        p = self.params
        T, R, C, NR = len(p.iou_thrs), len(p.rec_thrs), len(self._plan.cat_ids), \
            len(p.visibility_rng)
        if self._host_out is not None:
            prec_raw, recall = self._host_out.precision, self._host_out.recall
            self._num_gt = self._host_out.num_gt
        else:
            eng = get_engine(self.device)
            if self._world > 1:
                dist_accumulate(eng, self._dev, self._rank, self._world)
            else:                               # segm plans: staged device path
                eng.stage_accumulate(self._dev)
            t = self._dev.t
            prec_raw, recall = t["precision"].cpu().numpy(), t["recall"].cpu().numpy()
            self._num_gt = t["num_gt"].cpu().numpy()
        self.eval = {
            "params": p,
            "counts": [T, R, C, NR],
            "date": datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S"),
            "precision": restore_rec_order(prec_raw, recall, p.rec_thrs, self._rec_inv),
            "recall": recall,
            "dt_pointers": LazyDict(lambda: materialize.dt_pointers(
                self._plan, T, self._need_detail().dt_tpfp, self._num_gt)),
        }

    def _summarize(self, summary_type, iou_thr=None, visibility_rng="all", freq_group_idx=None):
        """lvis eval.py:428-457."""
        p = self.params
        aidx = [i for i, l in enumerate(p.visibility_rng_lbl) if l == visibility_rng]
        s = self.eval["precision"] if summary_type == 'ap' else self.eval["recall"]
        if iou_thr is not None:
            s = s[np.where(iou_thr == p.iou_thrs)[0]]
        if summary_type == 'ap':
            s = s[:, :, self.freq_groups[freq_group_idx], aidx] if freq_group_idx is not None \
                else s[:, :, :, aidx]
        else:
            s = s[:, :, aidx]
        sel = s[s > -1]
        return -1 if len(sel) == 0 else np.mean(sel)

    def summarize(self):
        """lvis eval.py:459-499 (same keys / order, including the AR key collision)."""
        if not self.eval:
            raise RuntimeError("Please run accumulate() first.")
        p = self.params
        self.results = materialize.summarize_lvis(
            self.eval["precision"], self.eval["recall"], p.iou_thrs, self.freq_groups,
            p.visibility_rng_lbl, p.max_dets)

    def run(self):
        self.evaluate()
        self.accumulate()
        self.summarize()

    def print_results(self):
        """lvis eval.py:507-545: printed to stdout (not logged), reference template."""
        template = (" {:<18} {} @[ IoU={:<9} | visibility={:>6s} | maxDets={:>3d} "
                    "catIds={:>3s}] = {:0.3f}")
        full = {'HO': 'Highly Occluded (vis < 0.1)',
                'PO': 'Partially Occluded (0.1 < vis < 0.8)',
                'HP': 'Highly + Partially Occluded (vis < 0.8)',
                'HV': 'Highly Visible (vis > 0.8)'}
        for key, value in self.results.items():
            max_dets = self.params.max_dets
            if "AP" in key:
                title, _type = "Average Precision", "(AP)"
            else:
                title, _type = "Average Recall", "(AR)"
            if len(key) > 2 and key[2].isdigit():
                iou = "{:0.2f}".format(float(key[2:4]) / 100)
            else:
                iou = "{:0.2f}:{:0.2f}".format(self.params.iou_thrs[0], self.params.iou_thrs[-1])
            cat_group_name = key[2] if (len(key) > 2 and key[2] in ["r", "c", "f"]) else "all"
            if len(key) > 2 and key[-2:] in full:
                vis = full[key[-2:]]
            elif len(key) > 2 and key[-3:] == "OOF":
                vis = "Out-of-Frame"
            else:
                vis = "all"
            print(template.format(title, _type, iou, vis, max_dets, cat_group_name, value))

    def get_results(self):
        if not self.results:
            self.logger.warn("results is empty. Call run().")
        return self.results
