"""Device engine: runs an ``EvalPlan`` (prep.py) through the C ABI of include/ta_eval.h.

Two entry levels, mirroring how the reference is used:

* ``Engine.evaluate_host(plan)`` — ONE C call (``ta_eval_plan_host``) that takes the plan
  from host memory, runs IoU -> greedy match -> PR accumulation on the GPU and returns the
  reference's ``eval['precision']`` / ``eval['recall']`` tensors.  This is what
  ``TaoEval.run()`` / ``LVISEval.run()`` of this repo call (reference:
  tao_amodal/evaluation/tao_amodal/eval.py:662-666, lvis_amodal/eval.py:501-505).
* ``Engine.upload(plan)`` + ``Engine.evaluate_device(dev, detail=...)`` — the three
  stage entry points (``ta_track_iou``/``ta_box_iou``, ``ta_match_greedy``,
  ``ta_pr_accumulate``) on device-resident buffers; torch tensors are used purely as
  device-memory containers.  ``detail=True`` also returns the IoU matrices, the matched GT
  per (range, threshold, detection) and the GT ignore flags, from which the reference's
  ``ious`` / ``eval_vids`` / ``eval_imgs`` structures are rebuilt lazily (materialize.py).

No CPU fallback exists: every method raises if libta_eval.so is missing or a call fails.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from . import _lib
from .prep import EvalPlan

IOU_THRS = np.linspace(0.5, 0.95, int(np.round((0.95 - 0.5) / 0.05)) + 1, endpoint=True)
REC_THRS = np.linspace(0.0, 1.00, int(np.round((1.00 - 0.0) / 0.01)) + 1, endpoint=True)


@dataclass
class EvalOutput:
    """Accumulated tensors in the reference's layouts (eval.py:474-484, lvis eval.py:320-326):
    precision [T,R,C,n_cfg], recall [T,C,n_cfg] (−1 where the reference leaves −1), the
    TP / FP totals per cell and the non-ignored GT count per (category, cfg)."""
    precision: np.ndarray
    recall: np.ndarray
    tp_cnt: np.ndarray
    fp_cnt: np.ndarray
    num_gt: np.ndarray
    h2d_bytes: int = 0
    d2h_bytes: int = 0
    # detail (optional)
    iou: Optional[np.ndarray] = None          # f64 [sum D*G], group g at plan.iou_off[g]
    dt_tpfp: Optional[np.ndarray] = None      # u32 [n_dt, n_cfg]
    dt_match_gt: Optional[np.ndarray] = None  # i32 [n_cfg, n_thr, n_dt]
    gt_ignore: Optional[np.ndarray] = None    # u8  [n_cfg, n_gt]


def plan_limits(plan: EvalPlan):
    """(g_max, n_slots_max) of a plan: the largest GT count of a group and 1 + the largest
    frame slot (0 on the frame path)."""
    cached = getattr(plan, "_limits", None)
    if cached is not None:
        return cached
    g_cnt = np.diff(plan.grp_gt_off)
    g_max = int(g_cnt.max()) if g_cnt.size else 0
    n_slots = 0
    if plan.kind == "tao":
        for s in (plan.dt_box_slot, plan.gt_box_slot):
            if s is not None and s.size:
                n_slots = max(n_slots, int(s.max()) + 1)
    plan._limits = (g_max, n_slots)
    return g_max, n_slots


def frame_big_groups(plan: EvalPlan, max_gt: int, max_dt: int, max_pairs: int) -> np.ndarray:
    """Groups of a frame-path plan that exceed the on-chip limits of ta_frame_eval."""
    D = np.diff(plan.grp_dt_off)
    G = np.diff(plan.grp_gt_off)
    return np.nonzero((G > 0) & ((G > max_gt) | (D > max_dt) | (D * G > max_pairs)))[0].astype(np.int32)


def lossless_f32_boxes(plan: EvalPlan):
    """(dt_box, gt_box) as float32 when every coordinate survives the round trip through
    float32 exactly (true for detector outputs that were float32 to begin with), else None.
    Cached on the plan; ta_eval_plan_host then ships half the box bytes over PCIe and widens
    them back to the identical fp64 values on the device."""
    got = getattr(plan, "_f32_boxes", 0)
    if got == 0:
        d32, g32 = plan.dt_box.astype(np.float32), plan.gt_box.astype(np.float32)
        ok = (np.array_equal(d32.astype(np.float64), plan.dt_box)
              and np.array_equal(g32.astype(np.float64), plan.gt_box))
        got = (np.ascontiguousarray(d32), np.ascontiguousarray(g32)) if ok else None
        plan._f32_boxes = got
    return got


def expand_words(dt_word: np.ndarray, dt_tpfp: np.ndarray, n_thr: int, n_cfg: int) -> np.ndarray:
    """Full [n_dt, n_cfg] TP/FP rows from the compact result words of ta_frame_eval
    (include/ta_eval.h): rows of detections whose word has bit 31 set are taken from dt_tpfp."""
    w = dt_word.view(np.uint32).astype(np.uint32)
    thr_all = np.uint32((1 << n_thr) - 1)
    M = (w & thr_all)[:, None]
    c = np.arange(n_cfg, dtype=np.uint32)[None, :]
    A = (w[:, None] >> (n_thr + c)) & 1
    B = (w[:, None] >> (n_thr + n_cfg + c)) & 1
    U = (w[:, None] >> (n_thr + 2 * n_cfg + c)) & 1
    tp = np.where(A == 1, M, 0)
    fp = np.where(B == 1, M, 0) | np.where(U == 1, thr_all & ~M, 0)
    rows = (tp | (fp << 16)).astype(np.uint32)
    full = (w >> 31) == 1
    rows[full] = dt_tpfp.view(np.uint32).reshape(-1, n_cfg)[full]
    return rows


def _ptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "plan arrays must be contiguous"
    return a.ctypes.data_as(C.c_void_p)


def shared_box_index(pool: EvalPlan, user: EvalPlan, max_extra: float = 0.25):
    """(idx, extra) such that concat(pool.dt_box, extra)[idx] == user.dt_box: for every detection
    box of `user` the row of the pool plan's box array that holds the same box; `extra` are the
    boxes of `user` the pool plan does not have (the two evaluators filter the result file
    differently), appended to the pool.  None when the plans do not say where their boxes came
    from or more than max_extra of the pool would be extras.  The identity of the boxes is
    checked value by value, not assumed."""
    a, b = pool.dt_box_src, user.dt_box_src
    if a is None or b is None or a.size == 0 or b.size == 0 or pool.dt_box.shape[0] != a.size:
        return None
    n_rows = int(max(a.max(), b.max())) + 1
    pos = np.full(n_rows, -1, dtype=np.int64)
    pos[a] = np.arange(a.size)
    idx = pos[b]
    missing = np.nonzero(idx < 0)[0]
    extra = np.zeros((0, 4), dtype=np.float64)
    if missing.size:
        msrc, first = np.unique(b[missing], return_index=True)
        if msrc.size > max_extra * a.size:
            return None
        extra = np.ascontiguousarray(user.dt_box[missing[first]])
        pos[msrc] = a.size + np.arange(msrc.size)
        idx = pos[b]
    if a.size + extra.shape[0] >= 2 ** 31:
        return None
    step = 1 << 20
    for i in range(0, idx.size, step):
        j = idx[i:i + step]
        own = j < a.size
        got = np.empty((j.size, 4), dtype=np.float64)
        got[own] = pool.dt_box[j[own]]
        got[~own] = extra[j[~own] - a.size]
        if not np.array_equal(got, user.dt_box[i:i + step]):
            return None
    return np.ascontiguousarray(idx.astype(np.int32)), extra


class HostPack:
    """The host-side transport form of the plans of one result set, built once and passed to
    ``Engine.evaluate_pack`` any number of times (include/ta_eval.h: ta_plan_host):

    * box coordinates as float32 when that is lossless (TA_PLAN_BOX_F32), frame slots as uint16
      (TA_PLAN_SLOT_U16), group tables as uint16 counts (TA_PLAN_GRP_U16);
    * with several plans, the track plan's detection boxes as int32 rows of the frame plan's box
      array (checked value by value; the few boxes only the track plan keeps are appended to
      that array): the boxes of the result file then cross PCIe once;
    * the plans ordered so that the one with the larger result tensors goes first (its download
      overlaps the upload of the next).

    ``pinned=True`` puts every array the C call reads into page-locked memory."""

    def __init__(self, eng: "Engine", plans, iou_mode, iou_thrs, rec_thrs, compress, pinned,
                 share_boxes):
        self.eng, self.plans = eng, plans
        self.n_thr, self.n_rec = len(iou_thrs), len(rec_thrs)
        self.pinned = pinned
        hold = eng.host_copy if pinned else np.ascontiguousarray
        iou_thrs = hold(np.asarray(iou_thrs, dtype=np.float64))
        rec_thrs = hold(np.asarray(rec_thrs, dtype=np.float64))
        self.structs, self.keep, self.shared = [], [], {}
        # plan with the larger result tensors first
        self.order = sorted(range(len(plans)),
                            key=lambda k: -len(plans[k].cat_ids) * plans[k].n_cfg)
        slot_of = {k: s for s, k in enumerate(self.order)}
        pool_idx = {}
        if share_boxes and compress and len(plans) > 1:
            frames = [k for k, p in enumerate(plans) if p.kind == "lvis" and p.masks is None]
            for k, p in enumerate(plans):
                if p.kind != "tao" or not frames:
                    continue
                if any(v[0] == frames[0] for v in pool_idx.values()):
                    continue                       # one user per pool
                got = shared_box_index(plans[frames[0]], p)
                if got is not None and (lossless_f32_boxes(p) is None) == (
                        lossless_f32_boxes(plans[frames[0]]) is None):
                    pool_idx[k] = (frames[0], got[0], got[1])
        self.shared = {k: v[0] for k, v in pool_idx.items()}
        extra_of = {v[0]: v[2] for v in pool_idx.values()}
        for k, plan in enumerate(plans):
            if plan.masks is not None:
                raise NotImplementedError("ta_eval_plan_host carries box plans only; use upload() + "
                                          "evaluate_device() for iou_type='segm'")
            g_max, n_slots = plan_limits(plan)
            track = plan.kind == "tao"
            ph = _lib.PlanHost()
            ph.n_groups, ph.n_dt, ph.n_gt = plan.n_groups, plan.n_dt, plan.n_gt
            ph.n_dt_boxes, ph.n_gt_boxes = plan.dt_box.shape[0], plan.gt_box.shape[0]
            ph.n_cat, ph.n_cfg = len(plan.cat_ids), plan.n_cfg
            ph.n_thr, ph.n_rec = self.n_thr, self.n_rec
            ph.n_slots_max, ph.g_max = n_slots, g_max
            ph.iou_mode = _lib.IOU_MODES[iou_mode]
            big = None if track else eng.big_list(plan)
            ph.n_big = 0 if big is None else int(big.size)
            flags = 0
            f32 = lossless_f32_boxes(plan) if compress else None
            dt_box, gt_box = f32 if f32 is not None else (plan.dt_box, plan.gt_box)
            if f32 is not None:
                flags |= 1
            if k in extra_of and extra_of[k].shape[0]:
                # boxes only the sharing plan uses ride behind this plan's own (rows >= n_dt)
                dt_box = np.concatenate([dt_box, extra_of[k].astype(dt_box.dtype)])
                ph.n_dt_boxes = dt_box.shape[0]
            dt_box_idx = None
            if k in pool_idx:
                # the pool's transport format decides how the shared boxes travel
                dt_box, dt_box_idx = None, pool_idx[k][1]
                ph.dt_box_pool = slot_of[pool_idx[k][0]]
            dt_slot, gt_slot = (plan.dt_box_slot, plan.gt_box_slot) if track else (None, None)
            if track and compress and n_slots <= 65536:
                dt_slot, gt_slot = plan.dt_box_slot.astype(np.uint16), plan.gt_box_slot.astype(np.uint16)
                flags |= 2
            grp_dt, grp_gt, grp_cat = plan.grp_dt_off, plan.grp_gt_off, plan.grp_cat
            if compress and plan.n_groups:
                d_cnt, g_cnt = np.diff(plan.grp_dt_off), np.diff(plan.grp_gt_off)
                if max(int(d_cnt.max()), int(g_cnt.max()), int(plan.grp_cat.max())) < 65536:
                    grp_dt, grp_gt = d_cnt.astype(np.uint16), g_cnt.astype(np.uint16)
                    grp_cat = plan.grp_cat.astype(np.uint16)
                    flags |= 4
            ph.flags = flags
            need_iou_off = track or ph.n_big > 0
            keep = dict(
                grp_dt_off=grp_dt, grp_gt_off=grp_gt, grp_cat=grp_cat,
                iou_off=plan.iou_off, cat_dt_off=plan.cat_dt_off, acc_perm=plan.acc_perm,
                big_list=big if (big is not None and big.size) else None,
                dt_box=dt_box, gt_box=gt_box, dt_box_idx=dt_box_idx,
                dt_trk_off=plan.dt_trk_box_off if track else None,
                gt_trk_off=plan.gt_trk_box_off if track else None,
                dt_slot=dt_slot, gt_slot=gt_slot,
                dt_attr_a=plan.dt_attr_a if track else None,
                dt_attr_b=plan.dt_attr_b if track else None,
                gt_attr_a=plan.gt_attr_a, gt_attr_b=plan.gt_attr_b if track else None,
                dt_flag=plan.dt_flag, gt_flag=plan.gt_flag, gt_hp=plan.gt_hp if track else None,
                cfgs=plan.range_cfgs)
            for name, v in list(keep.items()):
                if v is None:
                    continue
                if name == "iou_off" and not need_iou_off:
                    # only its last entry is read (on the host): no copy needed
                    v = np.ascontiguousarray(v)
                elif v.dtype.fields is not None:
                    raw = hold(np.ascontiguousarray(v).view(np.uint8))
                    v = raw.view(v.dtype)
                else:
                    v = hold(v)
                keep[name] = v
                setattr(ph, name, _ptr(v))
            ph.iou_thrs, ph.rec_thrs = _ptr(iou_thrs), _ptr(rec_thrs)
            keep["iou_thrs"], keep["rec_thrs"] = iou_thrs, rec_thrs
            self.structs.append(ph)
            self.keep.append(keep)

    def new_output(self, k: int) -> EvalOutput:
        plan = self.plans[k]
        mk = self.eng.host_empty if self.pinned else (lambda shape, dt: np.empty(shape, dtype=dt))
        T, R, Cn, K = self.n_thr, self.n_rec, len(plan.cat_ids), plan.n_cfg
        return EvalOutput(precision=mk((T, R, Cn, K), np.float64), recall=mk((T, Cn, K), np.float64),
                          tp_cnt=mk((T, Cn, K), np.int64), fp_cnt=mk((T, Cn, K), np.int64),
                          num_gt=mk((Cn, K), np.int32))


class Engine:
    """One ta_ctx on one CUDA device."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        if self.lib.ta_abi_version() != 4:
            raise RuntimeError("libta_eval.so ABI version mismatch")
        self.device = int(device)
        h = C.c_void_p()
        _lib.check(self.lib.ta_ctx_create(self.device, C.byref(h)))
        self._ctx = h
        self._extra = []
        self.fe_max_gt = int(self.lib.ta_frame_eval_max_gt())
        self.fe_max_dt = int(self.lib.ta_frame_eval_max_dt())
        self.fe_max_pairs = int(self.lib.ta_frame_eval_max_pairs())

    def big_list(self, plan: EvalPlan) -> np.ndarray:
        bl = getattr(plan, "_big_list", None)
        if bl is None:
            bl = frame_big_groups(plan, self.fe_max_gt, self.fe_max_dt, self.fe_max_pairs)
            plan._big_list = bl
        return bl

    def close(self):
        if getattr(self, "_ctx", None):
            for h in getattr(self, "_extra", []):
                self.lib.ta_ctx_destroy(h)
            self._extra = []
            self.lib.ta_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return sum(int(self.lib.ta_ctx_launch_count(h)) for h in [self._ctx] + self._extra)

    def timing(self, enable: bool):
        """Switch the per-kernel event log of the main context on / off (bench.py)."""
        _lib.check(self.lib.ta_ctx_timing(self._ctx, 1 if enable else 0))

    def timing_read(self):
        """{kernel name: (total ms, launches)} since the last read."""
        names = C.create_string_buffer(8192)
        ms = (C.c_double * 128)()
        cnt = (C.c_int * 128)()
        n = self.lib.ta_ctx_timing_read(self._ctx, names, 8192, ms, cnt, 128)
        if n < 0:
            _lib.check(n)
        keys = names.value.decode().split("\n")[:n]
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(keys)}

    @property
    def sm_count(self) -> int:
        return int(self.lib.ta_ctx_sm_count(self._ctx))

    # ------------------------------------------------------------------ host-buffer call
    def host_empty(self, shape, dtype) -> np.ndarray:
        """Uninitialised array in page-locked host memory (ta_host_alloc), freed with the array."""
        import weakref
        dtype = np.dtype(dtype)
        count = int(np.prod(shape, dtype=np.int64)) if np.ndim(shape) else int(shape)
        nbytes = max(1, count * dtype.itemsize)
        p = self.lib.ta_host_alloc(nbytes)
        if not p:
            raise _lib.TaEvalError(_lib.TA_ERR_CUDA, self.lib.ta_last_error().decode())
        buf = (C.c_ubyte * nbytes).from_address(p)
        weakref.finalize(buf, self.lib.ta_host_free, C.c_void_p(p))
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def host_copy(self, a: np.ndarray) -> np.ndarray:
        out = self.host_empty(a.shape, a.dtype)
        out[...] = a
        return out

    def pack_host(self, plans, iou_mode: str = "3d_iou", iou_thrs: np.ndarray = IOU_THRS,
                  rec_thrs: np.ndarray = REC_THRS, compress: bool = True, pinned: bool = False,
                  share_boxes: bool = True) -> "HostPack":
        """Transport form of one or several plans for ta_eval_plans_host (see HostPack)."""
        return HostPack(self, list(plans), iou_mode, iou_thrs, rec_thrs, compress, pinned,
                        share_boxes)

    def aux_ctx(self):
        """A second ta_ctx of this device, for library calls issued on another stream while the
        main context is busy (a context's scratch serves one stream at a time)."""
        if not self._extra:
            h = C.c_void_p()
            _lib.check(self.lib.ta_ctx_create(self.device, C.byref(h)))
            self._extra.append(h)
        return self._extra[0]

    def evaluate_pack(self, pack: "HostPack", outs=None):
        """ONE C call: every plan of the pack from host memory to its result tensors on the
        host.  Returns the EvalOutputs in the order the plans were given to pack_host."""
        n = len(pack.structs)
        while len(self._extra) < n - 1:
            h = C.c_void_p()
            _lib.check(self.lib.ta_ctx_create(self.device, C.byref(h)))
            self._extra.append(h)
        outs = list(outs) if outs is not None else [None] * n
        for i in range(n):
            if outs[i] is None:
                outs[i] = pack.new_output(i)
        ctxs = (C.c_void_p * n)(*([self._ctx] + self._extra)[:n])
        plans = (C.c_void_p * n)(*[C.addressof(pack.structs[k]) for k in pack.order])
        ho = (_lib.HostOut * n)()
        for slot, k in enumerate(pack.order):
            o = outs[k]
            ho[slot].precision, ho[slot].recall = _ptr(o.precision), _ptr(o.recall)
            ho[slot].tp_cnt, ho[slot].fp_cnt, ho[slot].num_gt = _ptr(o.tp_cnt), _ptr(o.fp_cnt), _ptr(o.num_gt)
        h2d, d2h = (C.c_int64 * n)(), (C.c_int64 * n)()
        _lib.check(self.lib.ta_eval_plans_host(n, ctxs, plans, ho, h2d, d2h))
        for slot, k in enumerate(pack.order):
            outs[k].h2d_bytes, outs[k].d2h_bytes = int(h2d[slot]), int(d2h[slot])
        return outs

    def evaluate_host_many(self, plans, outs=None, **kw):
        """Several plans of one result set at once (the CLI's track + frame evaluations)."""
        return self.evaluate_pack(self.pack_host(plans, **kw), outs)

    def evaluate_host(self, plan: EvalPlan, iou_mode: str = "3d_iou",
                      iou_thrs: np.ndarray = IOU_THRS, rec_thrs: np.ndarray = REC_THRS,
                      out: Optional[EvalOutput] = None, compress_boxes: bool = True) -> EvalOutput:
        pack = self.pack_host([plan], iou_mode, iou_thrs, rec_thrs, compress=compress_boxes)
        return self.evaluate_pack(pack, [out])[0]

    # ------------------------------------------------------------------ device-resident path
    def upload(self, plan: EvalPlan, iou_thrs: np.ndarray = IOU_THRS,
               rec_thrs: np.ndarray = REC_THRS) -> "DevicePlan":
        return DevicePlan(self, plan, iou_thrs, rec_thrs)

    def stage_iou(self, dev: "DevicePlan", iou_mode: str = "3d_iou"):
        """compute_iou of every group (eval.py:306-335 / lvis eval.py:168-192)."""
        import torch
        p, plan = dev.ptr, dev.plan
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if plan.kind == "tao":
            _lib.check(self.lib.ta_track_iou(
                self._ctx, st, _lib.IOU_MODES[iou_mode], plan.n_groups, p["grp_dt_off"],
                p["grp_gt_off"], p["dt_trk_off"], p["dt_box"], p["dt_slot"], p["gt_trk_off"],
                p["gt_box"], p["gt_slot"], dev.n_slots, p["iou_off"], p["iou"]))
        elif plan.masks is not None:
            # iou_type="segm": mask IoU (lvis eval.py:180-191 -> rleIou)
            _lib.check(self.lib.ta_rle_iou(
                self._ctx, st, plan.n_groups, None, 0, p["grp_dt_off"], p["grp_gt_off"],
                p["dt_rle_off"], p["dt_rle_counts"], p["dt_rle_hw"], p["dt_rle_bb"],
                p["gt_rle_off"], p["gt_rle_counts"], p["gt_rle_hw"], p["gt_rle_bb"],
                p["iou_off"], p["iou"]))
        else:
            _lib.check(self.lib.ta_box_iou(
                self._ctx, st, plan.n_groups, None, 0, p["grp_dt_off"], p["grp_gt_off"],
                p["dt_box"], p["gt_box"], p["iou_off"], p["iou"]))

    def stage_match(self, dev: "DevicePlan", detail: bool = False):
        """evaluate_vid / evaluate_img of every group x range x threshold
        (eval.py:337-457 / lvis eval.py:194-303) from the IoU matrices of stage_iou."""
        import torch
        plan = dev.plan
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.ta_zero(self._ctx, st, C.c_void_p(dev.t["num_gt"].data_ptr()),
                                    dev.t["num_gt"].numel() * 4))
        if detail:
            dev.ensure_detail()
        p = dev.ptr
        dev.words_valid = False
        _lib.check(self.lib.ta_match_greedy(
            self._ctx, st, plan.n_groups, None, 0, p["grp_dt_off"], p["grp_gt_off"], p["grp_cat"],
            p["iou_off"], p["iou"], dev.n_thr, p["iou_thrs"], plan.n_cfg, p["cfgs"],
            plan.n_dt, p["dt_attr_a"], p["dt_attr_b"], p["dt_flag"],
            plan.n_gt, p["gt_attr_a"], p["gt_attr_b"], p["gt_hp"], p["gt_flag"],
            dev.g_max, p["dt_tpfp"], p["num_gt"],
            p["dt_match_gt"] if detail else None, p["gt_ignore"] if detail else None))

    def stage_frame_eval(self, dev: "DevicePlan", detail: bool = False):
        """Fused frame path: LVISEval.compute_iou + evaluate_img in one kernel
        (lvis_amodal/eval.py:168-303).  Without `detail` the result of a detection is one compact
        word (dev.t["dt_word"]) unless dev.compact is False (full rows in dt_tpfp)."""
        import torch
        plan = dev.plan
        assert plan.kind == "lvis"
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.ta_zero(self._ctx, st, C.c_void_p(dev.t["num_gt"].data_ptr()),
                                    dev.t["num_gt"].numel() * 4))
        if detail:
            dev.ensure_detail()
        p = dev.ptr
        dev.words_valid = bool(dev.compact and not detail and "dt_word" in dev.t)
        _lib.check(self.lib.ta_frame_eval(
            self._ctx, st, plan.n_groups, p["grp_dt_off"], p["grp_gt_off"], p["grp_cat"],
            p["dt_box"], p["gt_box"], dev.n_thr, p["iou_thrs"], plan.n_cfg, p["cfgs"],
            plan.n_dt, p["dt_flag"], plan.n_gt, p["gt_attr_a"], p["gt_flag"],
            dev.n_big, p["big_list"] if dev.n_big else None, dev.g_max,
            p["iou_off"], p["iou"], 1 if detail else 0,
            p.get("sched"), p["dt_word"] if dev.words_valid else None,
            p["dt_tpfp"], p["num_gt"],
            p["dt_match_gt"] if detail else None, p["gt_ignore"] if detail else None))

    def stage_accumulate(self, dev: "DevicePlan"):
        """accumulate (eval.py:459-584 / lvis eval.py:305-426)."""
        import torch
        p, plan = dev.ptr, dev.plan
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.ta_pr_accumulate(
            self._ctx, st, dev.n_cat, p["cat_dt_off"], p["acc_perm"], plan.n_dt, p["dt_tpfp"],
            p["dt_word"] if dev.words_valid else None,
            p["num_gt"], dev.n_thr, plan.n_cfg, dev.n_rec, p["rec_thrs"],
            p["precision"], p["recall"], p["tp_cnt"], p["fp_cnt"]))

    def evaluate_device(self, dev: "DevicePlan", detail: bool = False,
                        iou_mode: str = "3d_iou", fetch: bool = True, fused: bool = True):
        """IoU -> match -> accumulate on dev's buffers (asynchronous on torch's current
        stream until results are fetched).  The frame path runs the fused kernel unless
        fused=False (then ta_box_iou + ta_match_greedy).  Returns EvalOutput or None."""
        plan = dev.plan
        if plan.kind == "lvis" and fused and plan.masks is None:
            self.stage_frame_eval(dev, detail)
        else:
            self.stage_iou(dev, iou_mode)
            self.stage_match(dev, detail)
        self.stage_accumulate(dev)
        if not fetch:
            return None
        if plan.kind == "tao" and iou_mode == "3d_iou":
            import torch
            bad = C.c_int32(0)
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self.lib.ta_ctx_take_assert_count(self._ctx, st, C.byref(bad)))
            if bad.value:                       # the reference asserts i <= u (eval.py:95)
                raise AssertionError("track IoU: intersection exceeds union in %d pairs" % bad.value)
        t = dev.t
        out = EvalOutput(
            precision=t["precision"].cpu().numpy(), recall=t["recall"].cpu().numpy(),
            tp_cnt=t["tp_cnt"].cpu().numpy(), fp_cnt=t["fp_cnt"].cpu().numpy(),
            num_gt=t["num_gt"].cpu().numpy())
        if detail:
            n_iou = int(plan.iou_off[-1]) if plan.iou_off.size else 0
            out.iou = t["iou"].cpu().numpy()[:n_iou]
            out.dt_tpfp = (t["dt_tpfp"].cpu().numpy().view(np.uint32)[:plan.n_cfg * plan.n_dt]
                           .reshape(plan.n_dt, plan.n_cfg))
            out.dt_match_gt = t["dt_match_gt"].cpu().numpy()[:, :, :plan.n_dt]
            out.gt_ignore = t["gt_ignore"].cpu().numpy()[:, :plan.n_gt]
        return out


class DevicePlan:
    """A plan resident in HBM: one torch tensor per column (device-memory containers only)."""

    def __init__(self, eng: Engine, plan: EvalPlan, iou_thrs, rec_thrs):
        import torch
        self.plan = plan
        self.eng = eng
        dev = torch.device("cuda", eng.device)
        self.dev = dev
        self.g_max, self.n_slots = plan_limits(plan)
        self.n_thr, self.n_rec, self.n_cat = len(iou_thrs), len(rec_thrs), len(plan.cat_ids)
        n_cfg = plan.n_cfg
        track = plan.kind == "tao"
        host: Dict[str, np.ndarray] = dict(
            grp_dt_off=plan.grp_dt_off, grp_gt_off=plan.grp_gt_off, iou_off=plan.iou_off,
            cat_dt_off=plan.cat_dt_off, grp_cat=plan.grp_cat, acc_perm=plan.acc_perm,
            dt_box=plan.dt_box, gt_box=plan.gt_box,
            dt_attr_a=plan.dt_attr_a, dt_attr_b=plan.dt_attr_b,
            gt_attr_a=plan.gt_attr_a, gt_attr_b=plan.gt_attr_b,
            dt_flag=plan.dt_flag, gt_flag=plan.gt_flag, gt_hp=plan.gt_hp,
            iou_thrs=np.ascontiguousarray(iou_thrs, dtype=np.float64),
            rec_thrs=np.ascontiguousarray(rec_thrs, dtype=np.float64),
            cfgs=plan.range_cfgs.view(np.uint8))
        self.n_big = 0
        if not track:
            big = eng.big_list(plan)
            self.n_big = int(big.size)
            host["big_list"] = big
        if track:
            host.update(dt_trk_off=plan.dt_trk_box_off, gt_trk_off=plan.gt_trk_box_off,
                        dt_slot=plan.dt_box_slot, gt_slot=plan.gt_box_slot)
        if plan.masks is not None:
            for side in ("dt", "gt"):
                off, cnt, hw, bb = plan.masks[side]
                host.update({side + "_rle_off": off, side + "_rle_counts": cnt,
                             side + "_rle_hw": hw, side + "_rle_bb": bb})
        self.t = {}
        self.input_bytes = 0
        self._input_keys = []
        for k, v in host.items():
            v = np.ascontiguousarray(v)
            if v.dtype == np.uint32:        # torch tensors are only memory containers here
                v = v.view(np.int32)
            self.input_bytes += v.nbytes
            if v.size == 0:
                self.t[k] = torch.zeros(16, dtype=torch.uint8, device=dev)
            else:
                self.t[k] = torch.from_numpy(v).to(dev)
                self._input_keys.append(k)
        self._host_of = lambda pl: dict(
            grp_dt_off=pl.grp_dt_off, grp_gt_off=pl.grp_gt_off, iou_off=pl.iou_off,
            cat_dt_off=pl.cat_dt_off, grp_cat=pl.grp_cat, acc_perm=pl.acc_perm,
            dt_box=pl.dt_box, gt_box=pl.gt_box, dt_attr_a=pl.dt_attr_a, dt_attr_b=pl.dt_attr_b,
            gt_attr_a=pl.gt_attr_a, gt_attr_b=pl.gt_attr_b, dt_flag=pl.dt_flag,
            gt_flag=pl.gt_flag, gt_hp=pl.gt_hp, dt_trk_off=pl.dt_trk_box_off,
            gt_trk_off=pl.gt_trk_box_off, dt_slot=pl.dt_box_slot, gt_slot=pl.gt_box_slot)
        n_iou = int(plan.iou_off[-1]) if plan.iou_off.size else 0
        C_, T, R = self.n_cat, self.n_thr, self.n_rec
        self.t["iou"] = torch.empty(max(n_iou, 1), dtype=torch.float64, device=dev)
        self.t["dt_tpfp"] = torch.zeros(max(n_cfg * plan.n_dt, 1), dtype=torch.int32, device=dev)
        self.t["num_gt"] = torch.zeros((C_, n_cfg), dtype=torch.int32, device=dev)
        # frame path: compact per-detection result words + the per-plan task schedule of
        # ta_frame_eval (depends on the CSR offsets only: built once, here)
        self.compact = True
        self.words_valid = False
        if not track and plan.masks is None:
            self.t["dt_word"] = torch.zeros(max(plan.n_dt, 1), dtype=torch.int32, device=dev)
            if plan.n_groups > 0 and plan.n_dt > 0:
                nb = int(eng.lib.ta_frame_sched_bytes(plan.n_groups, plan.n_dt, plan.n_gt,
                                                      self.n_cat, n_cfg))
                self.t["sched"] = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
                st = C.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)
                P = lambda k: C.c_void_p(self.t[k].data_ptr())
                _lib.check(eng.lib.ta_frame_sched_build(
                    eng._ctx, st, plan.n_groups, P("grp_dt_off"), P("grp_gt_off"), P("grp_cat"),
                    plan.n_dt, P("dt_flag"), plan.n_gt, P("gt_attr_a"), P("gt_flag"),
                    self.n_cat, n_cfg, P("cfgs"), P("sched")))
        self.t["precision"] = torch.empty((T, R, C_, n_cfg), dtype=torch.float64, device=dev)
        self.t["recall"] = torch.empty((T, C_, n_cfg), dtype=torch.float64, device=dev)
        self.t["tp_cnt"] = torch.empty((T, C_, n_cfg), dtype=torch.int64, device=dev)
        self.t["fp_cnt"] = torch.empty((T, C_, n_cfg), dtype=torch.int64, device=dev)
        self._refresh()

    def reload(self, plan: EvalPlan) -> int:
        """Copy the arrays of `plan` (same shapes as the uploaded one, e.g. its pinned twin) into
        the existing device buffers, asynchronously on the current stream.  Returns the bytes."""
        import torch
        src = self._host_of(plan)
        f32 = lossless_f32_boxes(plan)
        st = C.c_void_p(torch.cuda.current_stream(self.eng.device).cuda_stream)
        n = 0
        # the fused frame path reads neither the per-detection attributes (area = w * h of the
        # box) nor, without oversize groups, the IoU offsets: ta_eval_plan_host does not ship
        # them either
        unused = set()
        if plan.kind == "lvis" and plan.masks is None:
            unused = {"dt_attr_a", "dt_attr_b", "gt_attr_b", "gt_hp"} | ({"iou_off"} if not self.n_big else set())
        for k in self._input_keys:
            v = src.get(k)
            if v is None or k in unused:
                continue
            if f32 is not None and k in ("dt_box", "gt_box"):
                # lossless float transport: half the PCIe bytes, widened on the device
                h = f32[0] if k == "dt_box" else f32[1]
                stage = self.t.get(k + "_f32")
                if stage is None:
                    stage = self.t[k + "_f32"] = torch.empty(h.shape, dtype=torch.float32, device=self.dev)
                stage.copy_(torch.from_numpy(h), non_blocking=True)
                _lib.check(self.eng.lib.ta_widen_boxes(
                    self.eng._ctx, st, h.shape[0], C.c_void_p(stage.data_ptr()),
                    C.c_void_p(self.t[k].data_ptr())))
                n += h.nbytes
                continue
            self.t[k].copy_(torch.from_numpy(v), non_blocking=True)
            n += v.nbytes
        return n

    def reload_pack(self, pack: "HostPack", k: int, pool: Optional["DevicePlan"] = None,
                    only=None, skip=(), ctx=None) -> int:
        """Refresh the device buffers from plan `k` of a HostPack (the compact transport forms of
        ta_plan_host: float boxes, uint16 slots and group sizes, detection boxes shared with the
        plan resident in `pool`), asynchronously on the current stream, and rebuild what depends
        on them (the frame path's schedule).  `only` / `skip` select columns by name, so that a
        shared box array can go ahead of the rest; `ctx` is the context to issue the calls on
        (Engine.aux_ctx() when the current stream is not the one the main context works on).
        Returns the bytes copied from the host."""
        import torch
        eng, lib = self.eng, self.eng.lib
        keep, flags, plan = pack.keep[k], int(pack.structs[k].flags), self.plan
        ctx = ctx or eng._ctx
        assert pack.plans[k] is plan or pack.plans[k].n_dt == plan.n_dt
        st = C.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)
        P = lambda t: C.c_void_p(t.data_ptr())
        n = 0

        def staged(name, v):
            nonlocal n
            t = self.t.get(name + "_stage")
            if t is None or t.numel() != v.size or t.element_size() != v.itemsize:
                w = v.view(np.int16) if v.dtype == np.uint16 else v
                t = self.t[name + "_stage"] = torch.empty(w.shape, dtype=torch.from_numpy(w[:0]).dtype,
                                                          device=self.dev)
            t.copy_(torch.from_numpy(v.view(np.int16) if v.dtype == np.uint16 else v), non_blocking=True)
            n += v.nbytes
            return t

        for name, v in keep.items():
            if v is None or name in ("iou_thrs", "rec_thrs") or v.size == 0:
                continue
            if (only is not None and name not in only) or name in skip:
                continue
            if name == "iou_off" and not (plan.kind == "tao" or self.n_big):
                continue
            if name == "dt_box_idx":
                assert pool is not None, "this plan takes its boxes from another resident plan"
                _lib.check(lib.ta_gather_boxes(ctx, st, v.size, P(pool.t["dt_box_pool"]),
                                               P(staged(name, v)), P(self.t["dt_box"])))
            elif name in ("dt_box", "gt_box"):
                rows = v.shape[0]
                dst = self.t[name]
                if name == "dt_box" and rows != plan.dt_box.shape[0]:
                    # pool plan: the boxes only the sharing plan uses ride behind this plan's own
                    big = self.t.get("dt_box_pool")
                    if big is None or big.shape[0] != rows:
                        big = self.t["dt_box_pool"] = torch.empty((rows, 4), dtype=torch.float64, device=self.dev)
                        self.t["dt_box"] = big[:plan.dt_box.shape[0]]
                        self._refresh()
                    dst = big
                elif name == "dt_box":
                    self.t["dt_box_pool"] = dst
                if flags & 1:
                    _lib.check(lib.ta_widen_boxes(ctx, st, rows, P(staged(name, v)), P(dst)))
                else:
                    dst.copy_(torch.from_numpy(v), non_blocking=True)
                    n += v.nbytes
            elif name in ("dt_slot", "gt_slot") and flags & 2:
                _lib.check(lib.ta_widen_u16(ctx, st, v.size, P(staged(name, v)), P(self.t[name])))
            elif name in ("grp_dt_off", "grp_gt_off") and flags & 4:
                _lib.check(lib.ta_offsets_from_counts(ctx, st, v.size, P(staged(name, v)), P(self.t[name])))
            elif name == "grp_cat" and flags & 4:
                _lib.check(lib.ta_widen_u16(ctx, st, v.size, P(staged(name, v)), P(self.t[name])))
            else:
                w = v.view(np.uint8) if v.dtype.fields is not None else v
                self.t[name].copy_(torch.from_numpy(w.view(np.int32) if w.dtype == np.uint32 else w),
                                   non_blocking=True)
                n += v.nbytes
        if "sched" in self.t and only is None:
            Pk = lambda key: C.c_void_p(self.t[key].data_ptr())
            _lib.check(lib.ta_frame_sched_build(
                ctx, st, plan.n_groups, Pk("grp_dt_off"), Pk("grp_gt_off"), Pk("grp_cat"),
                plan.n_dt, Pk("dt_flag"), plan.n_gt, Pk("gt_attr_a"), Pk("gt_flag"),
                self.n_cat, plan.n_cfg, Pk("cfgs"), Pk("sched")))
        return n

    def ensure_detail(self):
        import torch
        if "dt_match_gt" in self.t:
            return
        plan = self.plan
        self.t["dt_match_gt"] = torch.empty(
            (plan.n_cfg, self.n_thr, max(plan.n_dt, 1)), dtype=torch.int32, device=self.dev)
        self.t["gt_ignore"] = torch.empty(
            (plan.n_cfg, max(plan.n_gt, 1)), dtype=torch.uint8, device=self.dev)
        self._refresh()

    def _refresh(self):
        self.ptr = {k: C.c_void_p(v.data_ptr()) for k, v in self.t.items()}
