"""``TaoEval`` — Track-AP evaluator on the CUDA library.

Mirror of tao_amodal/evaluation/tao_amodal/eval.py:120-757: constructor, ``params``,
``evaluate / accumulate / summarize / run / print_results / get_results`` and the public
attributes ``ious``, ``eval_vids``, ``eval``, ``results`` keep the reference's names, shapes
and log lines.  The per-(video, category) Python loops of the reference are replaced by
three device stages (engine.py); ``ious`` / ``eval_vids`` / ``eval['dt_pointers']`` are
rebuilt lazily from the device results when read.
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
from .results import TaoResults
from .tao import Tao


class Params:
    def __init__(self, iou_type, iou_3d_type='3d_iou'):
        """Same fields and defaults as eval.py:720-757."""
        self.vid_ids = []
        self.cat_ids = []
        self.iou_thrs = np.linspace(0.5, 0.95, int(np.round((0.95 - 0.5) / 0.05)) + 1,
                                    endpoint=True)
        self.rec_thrs = np.linspace(0.0, 1.00, int(np.round((1.00 - 0.0) / 0.01) + 1),
                                    endpoint=True)
        self.max_dets = 300
        self.area_rng = [[0 ** 2, 1e5 ** 2], [0 ** 2, 32 ** 2], [32 ** 2, 96 ** 2],
                         [96 ** 2, 1e5 ** 2], [0 ** 2, 1e5 ** 2]]
        self.area_rng_lbl = ["all", "small", "medium", "large", "highly-and-partially-occluded"]
        self.time_rng = [[0, 1e5], [0, 3], [3, 10], [10, 1e5]]
        self.time_rng_lbl = ["all", "short", "medium", "long"]
        self.use_cats = 1
        self.vid_count_lbl = ["r", "c", "f"]
        self.iou_type = iou_type
        self.iou_3d_type = iou_3d_type


class TaoEval:
    def __init__(self, tao_gt, tao_dt, logger=None, iou_type="bbox", iou_3d_type="3d_iou",
                 device=0, _plan=None):
        if not logger:
            self.logger = logging.getLogger('tao.eval')
        elif isinstance(logger, str):
            self.logger = logging.getLogger(logger)
        else:
            self.logger = logger
        if iou_type not in ["bbox", "segm"]:
            raise ValueError("iou_type: {} is not supported.".format(iou_type))
        if isinstance(tao_gt, Tao):
            self.tao_gt = tao_gt
        elif isinstance(tao_gt, str):
            self.tao_gt = Tao(tao_gt)
        else:
            raise TypeError("Unsupported type {} of tao_gt.".format(tao_gt))
        if isinstance(tao_dt, TaoResults):
            self.tao_dt = tao_dt
        elif isinstance(tao_dt, (str, list, DtColumns)):
            self.tao_dt = TaoResults(self.tao_gt, tao_dt)
        else:
            raise TypeError("Unsupported type {} of tao_dt.".format(tao_dt))
        self.eval_vids = {}
        self.eval = {}
        self.params = Params(iou_type=iou_type, iou_3d_type=iou_3d_type)
        self.results = OrderedDict()
        self.ious = {}
        self.params.vid_ids = sorted(self.tao_gt.get_vid_ids())
        self.params.cat_ids = sorted(self.tao_gt.get_cat_ids())
        self.device = device
        self._plan = None
        self._dev = None
        self._detail = None
        self._host_out = None
        self._rank, self._world = 0, 1
        # a plan built ahead of time for exactly these inputs and the default Params (the CLI's
        # BackgroundPlan); used only if the Params that shape the plan are still the defaults
        self._given_plan = _plan
        self._plan_defaults = self._plan_signature()

    def _plan_signature(self):
        p = self.params
        return (tuple(p.vid_ids), tuple(p.cat_ids), repr(p.area_rng), repr(p.time_rng),
                int(p.use_cats), int(self.tao_dt.max_dets))

    # ------------------------------------------------------------------------ device stages
    def _prepare(self):
        """Columnar equivalent of eval.py:178-233 (prep.prepare_tao)."""
        p = self.params
        if p.iou_type != "bbox":
            # the reference fails here too: _to_mask (eval.py:173-176) calls Tao.ann_to_rle, which
            # tao.py does not define
            raise AttributeError("'Tao' object has no attribute 'ann_to_rle'")
        if len(p.iou_thrs) > 16:
            raise ValueError("at most 16 IoU thresholds are supported")
        vid_ids = p.vid_ids
        self._rank, self._world = dist_info()
        if self._world > 1:
            # one shard of videos per GPU (torchrun): IoU + matching are local, accumulate()
            # exchanges the per-track records (parallel.py)
            from ... import parallel
            vid_ids = parallel.shard_videos(np.unique(vid_ids), self._world)[self._rank]
        if (self._given_plan is not None and self._world == 1
                and self._plan_signature() == self._plan_defaults):
            self._plan = self._given_plan
            return
        self._plan = prep.prepare_tao(
            self.tao_gt.columns, self.tao_dt.dt_columns, max_dets=self.tao_dt.max_dets,
            area_rng=p.area_rng, time_rng=p.time_rng, vid_ids=vid_ids,
            cat_ids=p.cat_ids if p.cat_ids else None, use_cats=bool(p.use_cats),
            allow_empty=self._world > 1)

    def evaluate(self, show_progress=False):
        """Per-video evaluation: IoU matrices + greedy matching of every (video, category,
        area range, duration range, threshold) on the GPU (eval.py:246-276)."""
        self.logger.info("Running per video evaluation.")
        self.logger.info("Evaluate annotation type *{}*".format(self.params.iou_type))
        self.params.vid_ids = list(np.unique(self.params.vid_ids))
        self._prepare()
        eng = get_engine(self.device)
        rec_sorted, self._rec_inv = ascending_rec_thrs(self.params.rec_thrs)
        self._rec_sorted = rec_sorted
        self._detail = None
        self._dev = self._host_out = None
        if self._world == 1:
            # single GPU: ONE C call from host buffers (ta_eval_plan_host: H2D, IoU, matching,
            # accumulation, D2H) — no torch anywhere on this path
            self._host_out = eng.evaluate_host(
                self._plan, iou_mode=self.params.iou_3d_type,
                iou_thrs=np.ascontiguousarray(self.params.iou_thrs, dtype=np.float64),
                rec_thrs=rec_sorted)
        else:
            self._dev = eng.upload(self._plan, self.params.iou_thrs, rec_sorted)
            dist_check_nonempty(self._dev, "Found no groundtruth annotations for given params",
                                "Found no predicted annotations for given params")
            eng.stage_iou(self._dev, self.params.iou_3d_type)
            eng.stage_match(self._dev)
        self.ious = LazyDict(lambda: materialize.iou_dict(self._plan, self._need_detail().iou))
        self.eval_vids = LazyDict(lambda: materialize.cells_dict(
            self._plan, len(self.params.iou_thrs), self._need_detail()))

    def _need_detail(self):
        """Second pass with the optional per-cell outputs switched on (only when someone
        reads ``ious`` / ``eval_vids`` / ``dt_pointers``)."""
        if self._world > 1:
            raise RuntimeError("ious / eval_vids / dt_pointers are per-process structures: "
                               "not available in multi-GPU mode (run single-GPU to read them)")
        if self._detail is None:
            eng = get_engine(self.device)
            if self._dev is None:
                self._dev = eng.upload(self._plan, self.params.iou_thrs, self._rec_sorted)
            self._detail = eng.evaluate_device(self._dev, detail=True,
                                               iou_mode=self.params.iou_3d_type)
        return self._detail

    def compute_iou(self, vid_id, cat_id):
        """ious of one (video, category) group (eval.py:306-335)."""
        return self.ious.get((vid_id, cat_id), [])

    def accumulate(self):
        """PR accumulation on the GPU (eval.py:459-584)."""
        self.logger.info("Accumulating evaluation results.")
        if self._plan is None:
            self.logger.warn("Please run evaluate first.")
            return
        p = self.params
        T, R, C = len(p.iou_thrs), len(p.rec_thrs), len(self._plan.cat_ids)
        A, Tm = len(p.area_rng), len(p.time_rng)
        if self._host_out is not None:
            recall, prec_raw = self._host_out.recall, self._host_out.precision
            self._num_gt = self._host_out.num_gt
        else:
            eng = get_engine(self.device)
            dist_accumulate(eng, self._dev, self._rank, self._world)
            t = self._dev.t
            recall, prec_raw = t["recall"].cpu().numpy(), t["precision"].cpu().numpy()
            self._num_gt = t["num_gt"].cpu().numpy()
        precision = restore_rec_order(prec_raw, recall, p.rec_thrs,
                                      self._rec_inv).reshape(T, R, C, A, Tm)
        recall = recall.reshape(T, C, A, Tm)
        self.eval = {
            "params": p,
            "counts": [T, R, C, A, Tm],
            "date": datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S"),
            "precision": precision,
            "recall": recall,
            "dt_pointers": LazyDict(lambda: materialize.dt_pointers(
                self._plan, T, self._need_detail().dt_tpfp, self._num_gt)),
        }

    # ------------------------------------------------------------------------ summaries
    def _summarize(self, summary_type, iou_thr=None, area_rng="all", time_rng="all",
                   freq_group_idx=None):
        """eval.py:586-623."""
        p = self.params
        aidx = [i for i, l in enumerate(p.area_rng_lbl) if l == area_rng]
        tidx = [i for i, l in enumerate(p.time_rng_lbl) if l == time_rng]
        s = self.eval["precision"] if summary_type == 'ap' else self.eval["recall"]
        if iou_thr is not None:
            s = s[np.where(iou_thr == p.iou_thrs)[0]]
        s = s[:, :, :, aidx, tidx] if summary_type == 'ap' else s[:, :, aidx, tidx]
        sel = s[s > -1]
        return -1 if len(sel) == 0 else np.mean(sel)

    def summarize(self):
        """eval.py:625-660 (same keys, same order)."""
        if not self.eval:
            raise RuntimeError("Please run accumulate() first.")
        p = self.params
        self.results = materialize.summarize_tao(
            self.eval["precision"], self.eval["recall"], p.iou_thrs, p.area_rng_lbl,
            p.time_rng_lbl, p.max_dets)

    def run(self, show_progress=False):
        self.evaluate(show_progress=show_progress)
        self.accumulate()
        self.summarize()

    def print_results(self):
        """eval.py:668-712: one logger.info line per metric, reference template."""
        template = (" {:<18} {} @[ IoU={:<9} | area={:>6s} | dur={:>6s} | maxDets={:>3d} "
                    "catIds={:>3s}] = {:0.3f}")
        for key, value in self.results.items():
            max_dets = self.params.max_dets
            if "AP" in key:
                title, _type = "Average Precision", "(AP)"
            else:
                title, _type = "Average Recall", "(AR)"
            area_rng = time_rng = "all"
            if isinstance(key, tuple):
                subset_type, subset_rng, max_dets = key[1:]
                if subset_type == "time":
                    time_rng = subset_rng[0]
                elif subset_type == "area":
                    area_rng = subset_rng[0]
                else:
                    raise ValueError('This should not happen')
            if len(key) > 2 and key[2].isdigit():
                iou = "{:0.2f}".format(float(key[2:4]) / 100)
            else:
                iou = "{:0.2f}:{:0.2f}".format(self.params.iou_thrs[0], self.params.iou_thrs[-1])
            cat_group_name = key[2] if (len(key) > 2 and key[2] in ["r", "c", "f"]) else "all"
            self.logger.info(template.format(title, _type, iou, area_rng, time_rng, max_dets,
                                             cat_group_name, value))

    def get_results(self):
        if not self.results:
            self.logger.warn("results is empty. Call run().")
        return self.results
