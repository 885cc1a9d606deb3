"""Multi-GPU evaluation: videos shard across ranks, one exchange before the PR accumulation.

IoU and greedy matching are independent per (video|image, category) group, so every rank
evaluates its own videos with no communication.  ``accumulate`` of the reference, however,
orders ALL detections of a category by score across all videos, ties broken by video (image)
order and then in-group order (tao_amodal/evaluation/tao_amodal/eval.py:498-511,
lvis_amodal/eval.py:340-361) — a histogram all-reduce cannot reproduce that, so the exchange
moves one compact record per detection to the rank that owns its category:

  setup (once per plan — the multi-GPU part of "building the plan", like ``acc_perm`` in prep):
      all_to_all of (score f64, order key i64, category i32); the owner orders its records by
      (category, -score, key) and keeps that permutation as its ``acc_perm``.
      key = unit_id * 2**24 + position in the group  (unit_id = video id / image id: the
      reference iterates units in sorted-id order, so this is its tie order).
  every evaluation:
      all_to_all of the TP/FP words  u32 [n_dt][n_cfg]   (the only per-step payload)
      all_reduce(SUM) of num_gt      i32 [C][n_cfg]
      ta_pr_accumulate on the owner's categories, renumbered densely (local index =
      category index // world), so each rank produces a [T, R, ceil(C / world), n_cfg] slice
      gather of the slices to rank 0, which interleaves them back into [T, R, C, n_cfg].

Categories are dealt round-robin (owner = category index mod world) to spread the skew of
category sizes.  Works with the NCCL backend on CUDA tensors and with gloo on CPU tensors (the
world_size-2 CPU tests inject a host PR function).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from .prep import EvalPlan

KEY_SHIFT = 24


def order_keys(plan: EvalPlan) -> np.ndarray:
    """int64 [n_dt]: unit_id * 2**24 + position of the detection inside its group."""
    D = np.diff(plan.grp_dt_off)
    unit = np.repeat(plan.unit_ids[plan.grp_unit].astype(np.int64), D)
    pos = np.arange(plan.n_dt, dtype=np.int64) - np.repeat(plan.grp_dt_off[:-1], D)
    if plan.n_dt:
        if int(pos.max()) >= (1 << KEY_SHIFT) or int(unit.max()) >= (1 << (62 - KEY_SHIFT)):
            raise ValueError("unit id / group size exceed the order-key encoding")
        if int(unit.min()) < 0:
            raise ValueError("negative unit ids are not supported by the multi-GPU exchange")
    return unit * (1 << KEY_SHIFT) + pos


def category_of_dt(plan: EvalPlan) -> np.ndarray:
    return np.repeat(plan.grp_cat.astype(np.int32), np.diff(plan.grp_dt_off))


class DistAccumulator:
    """Exchange state of one plan on one rank."""

    def __init__(self, plan: EvalPlan, rank: int, world: int, device, group=None):
        import torch
        import torch.distributed as dist
        self.plan, self.rank, self.world, self.group = plan, rank, world, group
        self.device = device
        self.n_cat, self.n_cfg = len(plan.cat_ids), plan.n_cfg
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        cat = t(category_of_dt(plan)).to(torch.int64)
        score = t(plan.dt_score)
        key = t(order_keys(plan))
        owner = cat % world
        # records grouped by destination rank, local order preserved inside a destination
        self.send_order = torch.sort(owner, stable=True).indices
        send_counts = torch.bincount(owner, minlength=world).to(torch.int64)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=group)
        self.send_splits = send_counts.cpu().tolist()
        self.recv_splits = recv_counts.cpu().tolist()
        self.n_recv = int(sum(self.recv_splits))

        def xchg(x):
            out = torch.empty((self.n_recv,) + tuple(x.shape[1:]), dtype=x.dtype, device=device)
            dist.all_to_all_single(out, x.index_select(0, self.send_order).contiguous(),
                                   self.recv_splits, self.send_splits, group=group)
            return out

        r_cat, r_score, r_key = xchg(cat), xchg(score), xchg(key)
        r_loc = r_cat // world              # dense local category index on the owner
        # (category, -score, key) order through three stable sorts, least significant first
        p = torch.sort(r_key, stable=True).indices
        p = p[torch.sort(-r_score[p], stable=True).indices]
        p = p[torch.sort(r_loc[p], stable=True).indices]
        self.acc_perm = p.to(torch.int32).contiguous()
        self.n_loc = -(-self.n_cat // world)                        # ceil(C / world), padded
        self.n_own = len(range(rank, self.n_cat, world))
        cnt = torch.bincount(r_loc, minlength=self.n_loc)
        self.cat_dt_off = torch.zeros(self.n_loc + 1, dtype=torch.int64, device=device)
        self.cat_dt_off[1:] = torch.cumsum(cnt, 0)
        self.recv_rows = torch.empty((max(self.n_recv, 1), self.n_cfg), dtype=torch.int32,
                                     device=device)

    # ---------------------------------------------------------------------------- per step
    def exchange_tpfp(self, tpfp_local):
        """tpfp_local: int32 [n_dt, n_cfg] (device order) -> this rank's received rows."""
        import torch.distributed as dist
        send = tpfp_local.index_select(0, self.send_order).contiguous()
        out = self.recv_rows[:self.n_recv]
        dist.all_to_all_single(out, send, self.recv_splits, self.send_splits, group=self.group)
        return out

    def global_num_gt(self, num_gt_local):
        """Sum over ranks; returns (global counts [C, n_cfg], this rank's categories
        [ceil(C / world), n_cfg], zero-padded)."""
        import torch
        import torch.distributed as dist
        g = num_gt_local.clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        own = torch.zeros((self.n_loc, g.shape[1]), dtype=g.dtype, device=g.device)
        own[:self.n_own] = g[self.rank::self.world]
        return g, own

    def merge_to_root(self, parts, full=None):
        """parts: this rank's [.., n_loc, n_cfg] tensors (category axis = -2).  Rank 0 receives
        every rank's slice and interleaves them into the matching `full` tensors."""
        import torch
        import torch.distributed as dist
        for i, x in enumerate(parts):
            x = x.contiguous()
            recv = [torch.empty_like(x) for _ in range(self.world)] if self.rank == 0 else None
            dist.gather(x, recv, dst=0, group=self.group)
            if self.rank == 0:
                for r in range(self.world):
                    n_r = len(range(r, self.n_cat, self.world))
                    full[i][..., r::self.world, :] = recv[r][..., :n_r, :]


class DeviceDistAccumulator(DistAccumulator):
    """DistAccumulator driving ta_pr_accumulate on the owner's slice (CUDA / NCCL)."""

    def __init__(self, eng, dev, rank: int, world: int, group=None):
        import torch
        super().__init__(dev.plan, rank, world, dev.dev, group)
        self.eng, self.dev = eng, dev
        dev.compact = False          # the exchange ships full TP/FP rows
        T, R, L, K = dev.n_thr, dev.n_rec, self.n_loc, self.n_cfg
        d = dev.dev
        self.part = {"precision": torch.empty((T, R, L, K), dtype=torch.float64, device=d),
                     "recall": torch.empty((T, L, K), dtype=torch.float64, device=d),
                     "tp_cnt": torch.empty((T, L, K), dtype=torch.int64, device=d),
                     "fp_cnt": torch.empty((T, L, K), dtype=torch.int64, device=d)}

    def accumulate(self):
        import ctypes as C
        import torch
        from . import _lib
        dev, eng = self.dev, self.eng
        t = dev.t
        n_dt, n_cfg = dev.plan.n_dt, self.n_cfg
        rows = self.exchange_tpfp(t["dt_tpfp"][:n_dt * n_cfg].view(n_dt, n_cfg))
        num_gt, num_gt_own = self.global_num_gt(t["num_gt"])
        st = C.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)
        P = lambda x: C.c_void_p(x.data_ptr())
        q = self.part
        _lib.check(eng.lib.ta_pr_accumulate(
            eng._ctx, st, self.n_loc, P(self.cat_dt_off), P(self.acc_perm), self.n_recv, P(rows),
            None, P(num_gt_own), dev.n_thr, n_cfg, dev.n_rec, dev.ptr["rec_thrs"],
            P(q["precision"]), P(q["recall"]), P(q["tp_cnt"]), P(q["fp_cnt"])))
        names = ("precision", "recall", "tp_cnt", "fp_cnt")
        self.merge_to_root([q[k] for k in names], [t[k] for k in names])
        t["num_gt"].copy_(num_gt)


def shard_videos(video_ids, world: int, weights=None):
    """Greedy bin packing of videos onto ranks by weight (default: equal), SURVEY §8e.
    Returns a list of `world` sorted id arrays."""
    vids = np.asarray(video_ids, dtype=np.int64)
    w = np.ones(vids.size) if weights is None else np.asarray(weights, dtype=np.float64)
    order = np.argsort(-w, kind="stable")
    load = np.zeros(world)
    bins = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        bins[r].append(int(vids[i]))
        load[r] += w[i]
    return [np.sort(np.asarray(b, dtype=np.int64)) for b in bins]
