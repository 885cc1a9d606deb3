"""Multi-GPU evaluation: videos shard across ranks, one exchange before the PR accumulation.

IoU and greedy matching are independent per (video|image, category) group, so every rank
evaluates its own videos with no communication.  ``accumulate`` of the reference, however,
orders ALL detections of a category by score across all videos, ties broken by video (image)
order and then in-group order (tao_amodal/evaluation/tao_amodal/eval.py:498-511,
lvis_amodal/eval.py:340-361) — a histogram all-reduce cannot reproduce that, so the exchange
moves one compact record per detection to the rank that owns its category:

  ownership   contiguous BLOCKS of categories, balanced by the global number of detections
              (one all-reduce of the per-category counts at setup).  A rank's detections are
              stored category-major (prep.py), so the records bound for one owner are one
              contiguous slice of the local result arrays: the send side needs no packing.
  setup       (once per plan — the multi-GPU part of "building the plan", like ``acc_perm``):
              all-to-all of (score f64, order key i64, category i32); the owner orders its records
              by (category, -score, key) and keeps that permutation as its ``acc_perm``.
              key = unit_id * 2**24 + position in the group  (unit_id = video id / image id: the
              reference iterates units in sorted-id order, so this is its tie order).
  every evaluation  (ONE fused NCCL launch per plan, ``ta_exchange_*`` of include/ta_eval.h):
              all-to-all of the result records — frame path: the compact 4-byte words of
              ta_frame_eval plus the full rows of the few detections the general matcher
              handled; track path: the full rows u32 [n_cfg] — and the all-reduce(SUM) of
              num_gt i32 [C][n_cfg];
              ta_pr_accumulate on the owner's category block, which yields the owner's
              [T, R, C_own, n_cfg] slice of precision (and recall / counts).
  results     stay with their owners (each copies its own slice out);  ``gather_to_root``
              assembles the reference's full tensors on rank 0 when a caller wants them.

Two transports carry the same logic: ``AbiTransport`` (the C ABI over NCCL, CUDA tensors — the
product path) and ``TorchTransport`` (``torch.distributed``; the world_size-2 gloo tests on CPU,
which inject a host PR function).
"""
from __future__ import annotations

import ctypes as C
from typing import List

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


def balanced_blocks(weights: np.ndarray, world: int) -> np.ndarray:
    """int64 [world + 1]: category block boundaries such that every block holds about the same
    total weight (number of detections); all-zero weights split by category count."""
    w = np.asarray(weights, dtype=np.float64)
    n = w.size
    if w.sum() <= 0:
        w = np.ones(n)
    pre = np.concatenate([[0.0], np.cumsum(w)])
    targets = pre[-1] * np.arange(1, world) / world
    inner = np.searchsorted(pre, targets, side="left")
    b = np.concatenate([[0], np.minimum(inner, n), [n]]).astype(np.int64)
    return np.maximum.accumulate(b)


# ------------------------------------------------------------------------------ transports
class TorchTransport:
    """torch.distributed (gloo on CPU tensors for the tests, or nccl)."""

    def __init__(self, rank: int, world: int, device, group=None):
        self.rank, self.world, self.device, self.group = rank, world, device, group

    def counts(self, send_counts: List[int]) -> List[int]:
        import torch
        import torch.distributed as dist
        s = torch.tensor(send_counts, dtype=torch.int64, device=self.device)
        r = torch.empty_like(s)
        dist.all_to_all_single(r, s, group=self.group)
        return [int(v) for v in r.cpu().tolist()]

    def all_to_all(self, send, send_counts, recv_counts, out=None):
        import torch
        import torch.distributed as dist
        n = int(sum(recv_counts))
        if out is None:
            out = torch.empty((n,) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(out[:n], send.contiguous(), list(recv_counts), list(send_counts),
                               group=self.group)
        return out[:n]

    def all_reduce(self, t):
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def begin(self):
        pass

    def end(self):
        pass


class AbiTransport:
    """The C ABI over NCCL (ta_exchange_*): device tensors, torch's current stream."""

    _DT = {"torch.int32": 0, "torch.int64": 1, "torch.float64": 2}

    def __init__(self, eng, rank: int, world: int, unique_id: bytes):
        import torch
        from . import _lib
        self.eng, self.rank, self.world = eng, rank, world
        self.lib = eng.lib
        h = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _lib.check(self.lib.ta_exchange_create(eng._ctx, rank, world, buf, C.byref(h)))
        self._h = h
        self.device = torch.device("cuda", eng.device)

    @staticmethod
    def unique_id(lib) -> bytes:
        from . import _lib
        buf = C.create_string_buffer(128)
        _lib.check(lib.ta_exchange_unique_id(buf, 128))
        return buf.raw

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ta_exchange_destroy(self._h)
            self._h = None

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.eng.device).cuda_stream)

    def counts(self, send_counts):
        import torch
        s = torch.tensor(send_counts, dtype=torch.int64, device=self.device)
        r = self.all_to_all(s, [1] * self.world, [1] * self.world)
        return [int(v) for v in r.cpu().tolist()]

    def all_to_all(self, send, send_counts, recv_counts, out=None):
        import torch
        from . import _lib
        row = int(np.prod(send.shape[1:], dtype=np.int64)) * send.element_size()
        n = int(sum(recv_counts))
        if out is None:
            out = torch.empty((max(n, 1),) + tuple(send.shape[1:]), dtype=send.dtype, device=self.device)
        so = np.concatenate([[0], np.cumsum(send_counts)]).astype(np.int64) * row
        ro = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64) * row
        assert send.is_contiguous() and out.is_contiguous()
        _lib.check(self.lib.ta_exchange_alltoallv(
            self._h, self._stream(), C.c_void_p(send.data_ptr()), so.ctypes.data_as(C.c_void_p),
            C.c_void_p(out.data_ptr()), ro.ctypes.data_as(C.c_void_p)))
        return out[:n]

    def all_reduce(self, t):
        from . import _lib
        assert t.is_contiguous()
        _lib.check(self.lib.ta_exchange_allreduce_sum(self._h, self._stream(), C.c_void_p(t.data_ptr()),
                                                      t.numel(), self._DT[str(t.dtype)]))
        return t

    def peer_window(self, n_bytes: int):
        """A PeerWindow of n_bytes on every rank (collective), or None when peer access is not
        available / switched off (TA_XCHG=nccl): the caller then exchanges through NCCL."""
        import os
        if os.environ.get("TA_XCHG", "peer") == "nccl":
            return None
        try:
            return PeerWindow(self, n_bytes)
        except Exception as e:       # noqa: BLE001 - e.g. no peer access between two devices
            import sys
            sys.stderr.write("peer windows unavailable (%s): exchanging through NCCL\n" % e)
            return None

    def begin(self):
        from . import _lib
        _lib.check(self.lib.ta_exchange_group_begin(self._h))

    def end(self):
        from . import _lib
        _lib.check(self.lib.ta_exchange_group_end(self._h))


class PeerWindow:
    """ta_peer_window: a device buffer of this rank that all ranks have mapped (CUDA IPC), and the
    kernel that pulls an owner's slices out of the peers' windows over NVLink."""

    def __init__(self, tr: AbiTransport, n_bytes: int):
        from . import _lib
        self.tr, self.lib = tr, tr.lib
        h = C.c_void_p()
        _lib.check(self.lib.ta_peer_window_create(tr._h, int(n_bytes), C.byref(h)))
        self._h = h
        self.base = int(self.lib.ta_peer_window_ptr(h))
        self.n_bytes = int(n_bytes)

    def set_plan(self, copies, sums):
        """copies: (peer, byte offset in the peer's window, bytes, local device pointer);
        sums: (byte offset, count, local int32 device pointer)."""
        from . import _lib
        cs = (_lib.PeerCopy * max(len(copies), 1))()
        for i, (peer, off, nb, dst) in enumerate(copies):
            cs[i].peer, cs[i].src_off, cs[i].bytes, cs[i].dst = int(peer), int(off), int(nb), int(dst)
        ss = (_lib.PeerSum * max(len(sums), 1))()
        for i, (off, cnt, dst) in enumerate(sums):
            ss[i].off, ss[i].count, ss[i].dst = int(off), int(cnt), int(dst)
        _lib.check(self.lib.ta_peer_window_set_plan(self._h, len(copies), cs, len(sums), ss))

    def acquire(self):
        from . import _lib
        _lib.check(self.lib.ta_peer_window_acquire(self._h, self.tr._stream()))

    def put(self, off: int, t, n_bytes: int):
        from . import _lib
        _lib.check(self.lib.ta_peer_window_put(self._h, self.tr._stream(), int(off),
                                               C.c_void_p(t.data_ptr()), int(n_bytes)))

    def exchange(self):
        from . import _lib
        _lib.check(self.lib.ta_peer_window_exchange(self._h, self.tr._stream()))

    def timed_out(self) -> bool:
        from . import _lib
        v = C.c_int32(0)
        _lib.check(self.lib.ta_peer_window_check(self._h, self.tr._stream(), C.byref(v)))
        return bool(v.value)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ta_peer_window_destroy(self._h)
            self._h = None


# ------------------------------------------------------------------------------ routing
class ExchangePlan:
    """Routing state of one plan on one rank (built once per plan): category ownership, the
    contiguous send slices, and the owner-side accumulate order of the received records."""

    def __init__(self, plan: EvalPlan, tr, device):
        import torch
        self.plan, self.tr, self.device = plan, tr, device
        self.rank, self.world = tr.rank, tr.world
        self.n_cat, self.n_cfg = len(plan.cat_ids), plan.n_cfg
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        cnt = t(np.diff(plan.cat_dt_off).astype(np.int64))
        tr.all_reduce(cnt)
        self.bounds = balanced_blocks(cnt.cpu().numpy(), self.world)        # [world + 1] categories
        dt_b = np.asarray(plan.cat_dt_off, dtype=np.int64)[self.bounds]     # local detection slices
        self.send_counts = [int(v) for v in np.diff(dt_b)]
        self.recv_counts = tr.counts(self.send_counts)
        self.n_recv = int(sum(self.recv_counts))
        self.c_lo, self.c_hi = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        self.n_loc = self.c_hi - self.c_lo
        r_cat = tr.all_to_all(t(category_of_dt(plan)).to(torch.int64), self.send_counts, self.recv_counts)
        r_score = tr.all_to_all(t(plan.dt_score), self.send_counts, self.recv_counts)
        r_key = tr.all_to_all(t(order_keys(plan)), self.send_counts, self.recv_counts)
        r_loc = r_cat - self.c_lo                # dense local category index on the owner
        # (category, -score, key) order through three stable sorts, least significant first
        p = torch.sort(r_key, stable=True).indices
        p = p[torch.sort(-r_score[p], stable=True).indices]
        p = p[torch.sort(r_loc[p], stable=True).indices]
        self.acc_perm = p.to(torch.int32).contiguous()
        cnt_loc = torch.bincount(r_loc, minlength=max(self.n_loc, 1))
        self.cat_dt_off = torch.zeros(self.n_loc + 1, dtype=torch.int64, device=device)
        if self.n_loc:
            self.cat_dt_off[1:] = torch.cumsum(cnt_loc[:self.n_loc], 0)

    # ---------------------------------------------------------------------------- peer windows
    def flag_routing(self, flag_idx: np.ndarray):
        """(send, recv) counts per rank of the detections `flag_idx` (sorted local indices) whose
        full rows travel next to the compact words."""
        b = np.asarray(self.plan.cat_dt_off, dtype=np.int64)[self.bounds]
        send = [int(v) for v in np.diff(np.searchsorted(flag_idx, b, side="left"))]
        return send, self.tr.counts(send)

    def window_plan(self, rec_bytes: int, flag=None):
        """Layout of this rank's peer window and what this owner pulls out of every rank's window
        (ta_peer_window_set_plan).  Byte offsets; the first two sections sit at the same place on
        every rank: GT counts [C, K] int32 | one record of rec_bytes per detection in local order
        (an owner's share is one contiguous slice) | with flag = (send counts, recv counts, row
        bytes): the rows of the flagged detections in send order.  A rank only knows where ITS
        slices start inside its own window, so those offsets are exchanged (tr.counts).
        copies: (peer, byte offset in the peer's window, bytes, "rec" | "flag", byte offset in the
        owner's receive buffer of that kind)."""
        al = lambda v: (int(v) + 255) & ~255
        n_dt = self.plan.n_dt
        lay = {"numgt": 0, "numgt_count": self.n_cat * self.n_cfg}
        lay["rec"] = al(lay["numgt_count"] * 4)
        lay["flag"] = lay["rec"] + al(n_dt * rec_bytes)
        n_flag = int(sum(flag[0])) if flag else 0
        row = int(flag[2]) if flag else 0
        lay["bytes"] = lay["flag"] + al(n_flag * row)
        dt_b = np.asarray(self.plan.cat_dt_off, dtype=np.int64)[self.bounds]
        rec_at = self.tr.counts([lay["rec"] + int(dt_b[r]) * rec_bytes for r in range(self.world)])
        copies, off = [], 0
        for p in range(self.world):
            copies.append((p, rec_at[p], self.recv_counts[p] * rec_bytes, "rec", off))
            off += self.recv_counts[p] * rec_bytes
        if flag:
            f_start = np.concatenate([[0], np.cumsum(flag[0])]).astype(np.int64)
            flag_at = self.tr.counts([lay["flag"] + int(f_start[r]) * row for r in range(self.world)])
            off = 0
            for p in range(self.world):
                copies.append((p, flag_at[p], flag[1][p] * row, "flag", off))
                off += flag[1][p] * row
        lay["copies"] = copies
        return lay

    # ---------------------------------------------------------------------------- per step
    def exchange(self, records, out=None):
        """records: [n_dt, ...] in local detection order -> this owner's received records."""
        return self.tr.all_to_all(records, self.send_counts, self.recv_counts, out=out)

    def global_num_gt(self, num_gt_local):
        """Sum over ranks (in place); returns (global counts [C, n_cfg], the owner's block)."""
        g = self.tr.all_reduce(num_gt_local)
        return g, g[self.c_lo:self.c_hi]

    def gather_to_root(self, parts, full=None):
        """parts: this rank's [.., n_loc, n_cfg] tensors (category axis = -2).  Rank 0 receives
        every owner's slice and places it at the owner's category block of the matching `full`
        tensor.  Not part of an evaluation step: results live with their owners."""
        for i, x in enumerate(parts):
            x = x.contiguous()
            lead = int(np.prod(x.shape[:-2], dtype=np.int64))
            per_cat = lead * x.shape[-1]
            sizes = [int(self.bounds[r + 1] - self.bounds[r]) * per_cat for r in range(self.world)]
            send_counts = [x.numel()] + [0] * (self.world - 1)
            recv_counts = sizes if self.rank == 0 else [0] * self.world
            got = self.tr.all_to_all(x.reshape(-1), send_counts, recv_counts)
            if self.rank == 0:
                off = 0
                for r in range(self.world):
                    lo, hi = int(self.bounds[r]), int(self.bounds[r + 1])
                    if hi > lo:
                        blk = got[off:off + sizes[r]].reshape(tuple(x.shape[:-2]) + (hi - lo, x.shape[-1]))
                        full[i][..., lo:hi, :] = blk
                    off += sizes[r]


class DeviceExchange(ExchangePlan):
    """ExchangePlan driving ta_pr_accumulate on the owner's block (CUDA, AbiTransport)."""

    def __init__(self, eng, dev, tr):
        import torch
        super().__init__(dev.plan, tr, dev.dev)
        self.eng, self.dev = eng, dev
        T, R, L, K = dev.n_thr, dev.n_rec, self.n_loc, self.n_cfg
        d = dev.dev
        self.part = {"precision": torch.empty((T, R, L, K), dtype=torch.float64, device=d),
                     "recall": torch.empty((T, L, K), dtype=torch.float64, device=d),
                     "tp_cnt": torch.empty((T, L, K), dtype=torch.int64, device=d),
                     "fp_cnt": torch.empty((T, L, K), dtype=torch.int64, device=d)}
        n_dt, n = dev.plan.n_dt, max(self.n_recv, 1)
        self.compact = bool(dev.compact and "dt_word" in dev.t)
        self.recv_rows = torch.zeros((n, K), dtype=torch.int32, device=d)
        self.num_gt_global = torch.zeros_like(dev.t["num_gt"])
        if self.compact:
            # one dry run of the local matcher tells which detections carry a full row (groups
            # the general matcher handled): their rows travel next to the compact words
            eng.stage_frame_eval(dev)
            words = dev.t["dt_word"][:n_dt]
            self.flag_idx = torch.nonzero(words < 0).flatten().to(torch.int32).contiguous()
            self.flag_send, self.flag_recv = self.flag_routing(self.flag_idx.cpu().numpy())
            self.recv_words = torch.zeros(n, dtype=torch.int32, device=d)
            self.exchange(words, out=self.recv_words)
            self.recv_flag_idx = (torch.nonzero(self.recv_words[:self.n_recv] < 0).flatten()
                                  .to(torch.int32).contiguous())
            assert int(self.recv_flag_idx.numel()) == int(sum(self.flag_recv))
            nf, nr = max(int(self.flag_idx.numel()), 1), max(int(self.recv_flag_idx.numel()), 1)
            self.send_flag_rows = torch.zeros((nf, K), dtype=torch.int32, device=d)
            self.recv_flag_rows = torch.zeros((nr, K), dtype=torch.int32, device=d)
        self.window = None
        if hasattr(tr, "peer_window"):
            self._open_window()

    def _open_window(self):
        """The peer window of this plan (ExchangePlan.window_plan) and its pull plan."""
        tr, K = self.tr, self.n_cfg
        rec = 4 if self.compact else 4 * K
        flag = (self.flag_send, self.flag_recv, 4 * K) if self.compact else None
        lay = self.window_plan(rec, flag)
        self.w_numgt, self.w_rec, self.w_flag = lay["numgt"], lay["rec"], lay["flag"]
        self.window = tr.peer_window(lay["bytes"])
        if self.window is None:
            return
        base = {"rec": (self.recv_words if self.compact else self.recv_rows).data_ptr(),
                "flag": self.recv_flag_rows.data_ptr() if self.compact else 0}
        copies = [(p, off, nb, base[kind] + dst) for p, off, nb, kind, dst in lay["copies"]]
        self.window.set_plan(copies, [(self.w_numgt, lay["numgt_count"], self.num_gt_global.data_ptr())])

    def accumulate(self):
        """Exchange + owner-side PR of the matcher outputs currently in dev's buffers."""
        import torch
        from . import _lib
        dev, eng, tr = self.dev, self.eng, self.tr
        t = dev.t
        n_dt, K = dev.plan.n_dt, self.n_cfg
        st = C.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)
        P = lambda x: C.c_void_p(x.data_ptr())
        rows_all = t["dt_tpfp"][:n_dt * K].view(n_dt, K)
        self.num_gt_global.copy_(t["num_gt"])
        use_words = self.compact and dev.words_valid
        if self.window is not None and use_words == self.compact:
            # the library's own exchange kernel over NVLink peer memory
            w = self.window
            w.acquire()
            w.put(self.w_numgt, t["num_gt"], self.n_cat * K * 4)
            if use_words:
                w.put(self.w_rec, t["dt_word"], n_dt * 4)
                nf = int(self.flag_idx.numel())
                if nf:
                    _lib.check(eng.lib.ta_exchange_gather(eng._ctx, st, nf, K, P(self.flag_idx), P(rows_all),
                                                          C.c_void_p(w.base + self.w_flag)))
            else:
                w.put(self.w_rec, rows_all, n_dt * K * 4)
            w.exchange()
            nr = int(self.recv_flag_idx.numel()) if use_words else 0
            if nr:
                _lib.check(eng.lib.ta_exchange_scatter(eng._ctx, st, nr, K, P(self.recv_flag_idx),
                                                       P(self.recv_flag_rows), P(self.recv_rows)))
        elif use_words:
            nf = int(self.flag_idx.numel())
            if nf:
                _lib.check(eng.lib.ta_exchange_gather(eng._ctx, st, nf, K, P(self.flag_idx),
                                                      P(rows_all), P(self.send_flag_rows)))
            tr.begin()
            self.exchange(t["dt_word"][:n_dt], out=self.recv_words)
            tr.all_to_all(self.send_flag_rows[:nf], self.flag_send, self.flag_recv,
                          out=self.recv_flag_rows)
            tr.all_reduce(self.num_gt_global)
            tr.end()
            nr = int(self.recv_flag_idx.numel())
            if nr:
                _lib.check(eng.lib.ta_exchange_scatter(eng._ctx, st, nr, K, P(self.recv_flag_idx),
                                                       P(self.recv_flag_rows), P(self.recv_rows)))
        else:
            tr.begin()
            self.exchange(rows_all, out=self.recv_rows)
            tr.all_reduce(self.num_gt_global)
            tr.end()
        own = self.num_gt_global[self.c_lo:self.c_hi]
        q = self.part
        if self.n_loc:
            _lib.check(eng.lib.ta_pr_accumulate(
                eng._ctx, st, self.n_loc, P(self.cat_dt_off), P(self.acc_perm), self.n_recv,
                P(self.recv_rows), P(self.recv_words) if use_words else None, P(own),
                dev.n_thr, K, dev.n_rec, dev.ptr["rec_thrs"],
                P(q["precision"]), P(q["recall"]), P(q["tp_cnt"]), P(q["fp_cnt"])))

    def to_root(self):
        """Assemble the reference's full tensors in dev.t on rank 0 (API / parity checks)."""
        names = ("precision", "recall", "tp_cnt", "fp_cnt")
        if self.window is not None and self.window.timed_out():
            raise RuntimeError("peer exchange: a rank never published its window (a peer process "
                               "died or did not take part in this evaluation)")
        self.gather_to_root([self.part[k] for k in names], [self.dev.t[k] for k in names])
        self.dev.t["num_gt"].copy_(self.num_gt_global)


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
