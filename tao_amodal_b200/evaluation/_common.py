"""Shared plumbing of the two drop-in evaluators."""
from __future__ import annotations

import json
import os
from typing import Dict, Tuple

from ..columnar import DtColumns, GtColumns

_ENGINES: Dict[int, object] = {}
_JSON_CACHE: Dict[Tuple[str, float, int], object] = {}

import numpy as np


import threading

_ENGINE_LOCK = threading.Lock()


def get_engine(device: int = 0):
    """One Engine (ta_ctx) per device per process (thread-safe: the CLI creates it on a
    background thread while the JSON files are parsed — CUDA context creation takes a few
    hundred milliseconds)."""
    from ..engine import Engine
    with _ENGINE_LOCK:
        if device not in _ENGINES:
            _ENGINES[device] = Engine(device)
        return _ENGINES[device]


def warm_engine(device: int = 0):
    t = threading.Thread(target=get_engine, args=(device,), daemon=True)
    t.start()
    return t


class BackgroundPlan:
    """Builds the track evaluator's plan on a worker thread while the frame evaluator runs
    (numpy's sorts / searches release the GIL).  Any exception is swallowed here: the consumer
    then builds the plan itself and raises at the place the reference would."""

    def __init__(self, annotation: str, results: str):
        self.annotation, self.results = annotation, results
        self.gt = self.dt = self.plan = None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        try:
            from .. import ingest, prep
            g = ingest.load_gt(self.annotation)
            d = ingest.load_dt(self.results).copy()
            prep.make_track_ids_unique(d)
            plan = prep.prepare_tao(g, d)
            self.gt, self.dt, self.plan = g, d, plan
        except BaseException:       # noqa: BLE001 - the foreground path reports errors
            self.gt = self.dt = self.plan = None

    def take(self):
        """(gt columns, uniquified dt columns, plan) or (None, None, None)."""
        self._thread.join()
        return self.gt, self.dt, self.plan


def load_json(path: str):
    """json.load with a per-process cache keyed by (path, mtime, size): the reference's CLI
    parses the annotation file twice and the result file twice
    (tools/eval_on_tao_amodal.py:100,122,127; lvis.py:27, results.py:29-30).  Every caller gets
    its OWN deep copy: the dataset views mutate the parsed objects in place (merged categories,
    averaged track scores, added id / area / segmentation — results.py:47-98, lvis
    results.py:44-66), and the reference parses a fresh copy for each class."""
    import copy
    st = os.stat(path)
    key = (os.path.abspath(path), st.st_mtime, st.st_size)
    if key not in _JSON_CACHE:
        with open(path, "r") as f:
            _JSON_CACHE[key] = json.load(f)
    return copy.deepcopy(_JSON_CACHE[key])


class LazyDict(dict):
    """A dict whose content is produced on first access (the reference-shaped ``ious`` /
    ``eval_vids`` views of device results are only materialised when somebody reads them)."""

    def __init__(self, loader):
        super().__init__()
        self._loader = loader

    def _fill(self):
        if self._loader is not None:
            loader, self._loader = self._loader, None
            super().update(loader())

    def __getitem__(self, k):
        self._fill()
        return super().__getitem__(k)

    def __iter__(self):
        self._fill()
        return super().__iter__()

    def __len__(self):
        self._fill()
        return super().__len__()

    def __contains__(self, k):
        self._fill()
        return super().__contains__(k)

    def __bool__(self):
        return True

    def keys(self):
        self._fill()
        return super().keys()

    def values(self):
        self._fill()
        return super().values()

    def items(self):
        self._fill()
        return super().items()

    def get(self, k, default=None):
        self._fill()
        return super().get(k, default)


def dist_info():
    """(rank, world) of the default torch.distributed group, (0, 1) when not initialised.
    With world > 1 the evaluators shard videos across ranks (parallel.py)."""
    import sys
    if "torch.distributed" not in sys.modules and int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return 0, 1          # single process: never import torch (seconds of start-up time)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


_TRANSPORTS = {}


def dist_transport(eng, rank, world):
    """The NCCL communicator of the C ABI (ta_exchange_*) for this engine, created once: rank 0
    draws the id, torch.distributed (already initialised by the caller) hands it round."""
    import torch.distributed as dist
    from .. import parallel
    tr = _TRANSPORTS.get(id(eng))
    if tr is None:
        box = [parallel.AbiTransport.unique_id(eng.lib) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        tr = parallel.AbiTransport(eng, rank, world, box[0])
        _TRANSPORTS[id(eng)] = tr
    return tr


def dist_check_nonempty(dev, what_gt, what_dt, dt_error=ValueError):
    """The reference's "found no ... annotations" errors (eval.py:188-192; an empty result list
    is an IndexError in lvis_amodal/results.py) on the WHOLE set: a shard alone may be empty
    (prep allow_empty), the sum over ranks may not."""
    import torch
    import torch.distributed as dist
    n = torch.tensor([dev.plan.n_gt, dev.plan.n_dt], dtype=torch.int64, device=dev.dev)
    dist.all_reduce(n)
    n_gt, n_dt = (int(v) for v in n.cpu().tolist())
    if n_gt == 0 and dt_error is ValueError:
        raise ValueError(what_gt)
    if n_dt == 0:
        raise dt_error(what_dt)


def dist_accumulate(eng, dev, rank, world):
    """Cross-rank PR accumulation (parallel.DeviceExchange over the C ABI's NCCL exchange);
    every rank ends up with the merged tensors, as the single-process API promises."""
    import torch.distributed as dist
    from .. import parallel
    ex = parallel.DeviceExchange(eng, dev, dist_transport(eng, rank, world))
    if dev.plan.kind == "lvis" and ex.compact:
        pass                     # the setup's dry run left this evaluation's words in place
    ex.accumulate()
    ex.to_root()
    for k in ("precision", "recall", "tp_cnt", "fp_cnt"):
        dist.broadcast(dev.t[k], src=0)


def ascending_rec_thrs(rec_thrs):
    """The PR kernels take the recall thresholds in ascending order.  Returns (ascending copy,
    inv) with ``rec_thrs[k] == ascending[inv[k]]``; inv is None when the given order already is
    ascending (the evaluators' default, np.linspace)."""
    r = np.ascontiguousarray(rec_thrs, dtype=np.float64)
    if r.size < 2 or bool(np.all(r[1:] >= r[:-1])):
        return r, None
    order = np.argsort(r, kind="stable")
    inv = np.empty(r.size, dtype=np.int64)
    inv[order] = np.arange(r.size)
    return np.ascontiguousarray(r[order]), inv


def restore_rec_order(precision, recall, rec_thrs, inv):
    """Undo ascending_rec_thrs on precision [T, R, C, K] (recall [T, C, K]), including the
    reference's behaviour for thresholds given out of order: its loop over the recall
    thresholds stops at the FIRST one no detection reaches (bare ``except: pass`` around the
    loop, eval.py:565-571 / lvis_amodal/eval.py:405-411), so every later entry of the cell stays
    0.0 even if its threshold is reachable."""
    if inv is None:
        return precision
    out = np.ascontiguousarray(precision[:, inv])
    r = np.asarray(rec_thrs, dtype=np.float64)
    unreachable = r[None, :, None, None] > recall[:, None]
    stopped = np.maximum.accumulate(unreachable, axis=1)
    out[stopped & (out != -1)] = 0.0
    return out
