#!/usr/bin/env python
"""Benchmark of the TAO-Amodal evaluation hot path on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3]

A *step* is one pass of the whole hot path — track-AP (3-D IoU -> greedy match -> PR
accumulate) and visibility-split frame-AP (box IoU -> greedy match -> PR accumulate) — over
one synthetic prediction + annotation set of BASELINE.json's configs[2] shape
(500 videos x 300 frames, 200 predicted / 30 GT tracks per video, 1203 categories), per GPU.
``value`` = box-pairs / s with the prepared columns resident in HBM; ``e2e`` = the same metric
through ``ta_eval_plan_host`` (host buffers in pinned memory, H2D + D2H inside the timed
region).  With N > 1 (torchrun) every rank evaluates its own 500-video shard of an N x 500
video set ("weak" scaling) and the per-detection records are exchanged over NCCL for the
category-sharded PR accumulation (parallel.py).

``--impl reference`` times the CPU implementation of the same path (the oracle port of the
reference, oracle/; the Python reference itself cannot travel to the GPU box) on all host
cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "box_pairs_per_s"
UNIT = "box-pairs/s"


# ------------------------------------------------------------------------------ workload
def make_workload(name: str, rank: int, videos: int = 0):
    from tao_amodal_b200 import prep, synth
    over = {"seed": synth.CONFIGS[name].seed + 7919 * rank}
    if videos:
        over["videos"] = videos
    gt, dt = synth.generate_named(name, **over)
    if rank:
        # every rank holds different videos of one N x 500-video dataset: make the video and
        # image ids globally unique (they define the cross-rank tie order, parallel.py)
        v_off, i_off = rank * 1_000_000, rank * 100_000_000
        gt.vid_id = gt.vid_id + v_off
        gt.img_video_id = gt.img_video_id + v_off
        gt.trk_video_id = gt.trk_video_id + v_off
        dt.video_id = dt.video_id + v_off
        gt.img_id = gt.img_id + i_off
        gt.ann_image_id = gt.ann_image_id + i_off
        dt.image_id = dt.image_id + i_off
    lvis_plan = prep.prepare_lvis(gt, dt)
    dt2 = dt.copy()
    prep.make_track_ids_unique(dt2)
    tao_plan = prep.prepare_tao(gt, dt2)
    return gt, dt, tao_plan, lvis_plan


def algorithmic_bytes(plan, cells_with_gt: int = 0) -> dict:
    """Algorithmic HBM bytes of each kernel for one launch on this plan (DESIGN.md §4,
    SURVEY §8d): every array a kernel has to read or write, counted once.  cells_with_gt =
    number of (category, cfg) cells with non-ignored GT (only those produce precision values
    other than the -1 fill)."""
    nd_box, ng_box = plan.dt_box.shape[0], plan.gt_box.shape[0]
    n_iou = int(plan.iou_off[-1])
    n_cfg, n_dt, n_gt = plan.n_cfg, plan.n_dt, plan.n_gt
    per_box = 36 if plan.kind == "tao" else 32
    n_cat = len(plan.cat_ids)
    T, R = 10, 101
    if plan.kind == "tao":
        # the IoU kernel only touches groups that have both detections and GT
        D, G = np.diff(plan.grp_dt_off), np.diff(plan.grp_gt_off)
        act = (D > 0) & (G > 0)
        db, gb = plan.dt_trk_box_off[plan.grp_dt_off], plan.gt_trk_box_off[plan.grp_gt_off]
        nd_box = int((db[1:] - db[:-1])[act].sum())
        ng_box = int((gb[1:] - gb[:-1])[act].sum())
    n_chunks = int(np.ceil(np.diff(plan.cat_dt_off) / 256.0).sum())
    n_tasks = n_dt // 256 + n_gt // 64 + 1
    compact = plan.kind == "lvis" and T + 3 * n_cfg <= 31
    word_bytes = 4 if compact else 4 * n_cfg          # result bytes per detection
    return {
        "iou": per_box * (nd_box + ng_box) + 8 * n_iou,
        # IoU matrix + dt (area, n_anns, flag) + gt (attr a, b, hp, flag) read, TP/FP words written
        "match": 8 * n_iou + 17 * n_dt + 21 * n_gt + 4 * n_cfg * n_dt,
        # streamed lane-per-detection frame kernel: boxes (each once), the per-detection
        # descriptor word, the per-GT word, the task table; one result word per detection
        "frame_flat": 32 * (nd_box + ng_box) + 4 * n_dt + 4 * n_gt + 16 * n_tasks + word_bytes * n_dt,
        # group table + GT visibility/flags read, per-GT words written
        "frame_prep": 20 * plan.n_groups + 9 * n_gt + 4 * n_gt,
        # permutation + TP/FP words read, chunk counters written
        "pr_count": n_dt * (4 + 4 * n_cfg) + 128 * n_cfg * n_chunks,
        "pr_envelope": n_dt * (4 + 4 * n_cfg) + (128 + 8 * T) * n_cfg * n_chunks
                       + 8 * T * R * cells_with_gt,
        # bit-plane path: k_pr_bits reads permutation + result words, writes the TP / FP planes
        # (64 B per cell and chunk) and the chunk counters; k_pr_envelope_bits reads the planes +
        # counters and writes chunk bests + the answered recall levels of the cells with GT
        "pr_bits": n_dt * (4 + word_bytes) + (128 + 64 * T) * n_cfg * n_chunks,
        "pr_envelope_bits": (128 + 64 * T + 8 * T) * n_cfg * n_chunks + 8 * T * R * cells_with_gt,
        # precision tensor written once, answered entries read once
        "pr_finalize": 8 * T * R * n_cat * n_cfg + 8 * T * R * cells_with_gt,
    }


# ------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), 0
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": float(self.max_mhz) if self.max_mhz else None,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------ CPU legs
def _cpu_sample_worker(args):
    """One bounded sample: the oracle port of the reference on one video's sub-dataset."""
    gt_dict, res_list = args
    import copy
    from oracle import lvis_frame, tao_track
    t0 = time.perf_counter()
    o_l = lvis_frame.evaluate_lvis(copy.deepcopy(gt_dict), copy.deepcopy(res_list), keep_cells=False)
    res2 = copy.deepcopy(res_list)
    tao_track.uniquify_track_ids(res2)
    o_t = tao_track.evaluate_tao(gt_dict, res2, keep_cells=False)
    return o_t["box_pair_visits"] + o_l["box_pairs"], time.perf_counter() - t0


def cpu_samples(gt, dt, n_samples: int, frames: int):
    """n one-video samples (first `frames` frames of each) in the reference's JSON structures."""
    from tao_amodal_b200.columnar import subset_videos
    out = []
    vids = np.unique(gt.vid_id)[:n_samples]
    for v in vids:
        g, d = subset_videos(gt, dt, [int(v)])
        gd, dl = g.to_dict(), d.to_list()
        if frames:
            keep = {im["id"] for im in gd["images"] if im["frame_index"] < frames}
            gd["images"] = [im for im in gd["images"] if im["id"] in keep]
            gd["annotations"] = [a for a in gd["annotations"] if a["image_id"] in keep]
            live = {a["track_id"] for a in gd["annotations"]}
            gd["tracks"] = [t for t in gd["tracks"] if t["id"] in live]
            dl = [r for r in dl if r["image_id"] in keep]
        out.append((gd, dl))
    return out


def run_cpu_port(samples, cores: int):
    """All samples through a process pool of `cores` workers; returns (pairs, wall seconds)."""
    t0 = time.perf_counter()
    if cores <= 1:
        res = [_cpu_sample_worker(s) for s in samples]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_sample_worker, samples)
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res), wall


CPU_SAMPLE_FRAMES = 0     # 0 = whole videos (cfg3 videos have 300 frames)

# The reference arm and the cpu_baseline leg time the UNMODIFIED reference (baseline/_ref, a
# verbatim copy made by __graft_entry__.build(); oracle/ref_bench.py) on ONE core — it is
# single-threaded — over this FIXED sample of the bench workload: one video, its first 150
# frames, all 1203 categories (the reference's cost is the (image x category) grid walk, linear
# in images: ~8 s per pass, so --steps 20 --warmup 5 ends in a few minutes).
REF_SAMPLE_VIDEOS = 1
REF_SAMPLE_FRAMES = 150


def reference_sample(workload, td):
    from oracle import ref_bench
    return ref_bench.write_sample(workload, REF_SAMPLE_VIDEOS, td, REF_SAMPLE_FRAMES)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import tempfile
    from oracle import ref_bench
    ref_root = ref_bench.find_reference()
    extra = {}
    if ref_root is not None:
        with tempfile.TemporaryDirectory() as td:
            ap, rp, pairs_1, sample = reference_sample(args.workload, td)
            for _ in range(args.warmup):
                ref_bench.run_once(ap, rp)
            wall, res = 0.0, None
            for _ in range(args.steps):
                s, res = ref_bench.run_once(ap, rp)
                wall += s
            pairs = pairs_1 * args.steps
            try:        # best-case CPU variant: numba's per-call dispatch switched off
                nj = ref_bench.run_subprocess(ap, rp, True, 1)[0]
                extra["value_numba_disabled"] = pairs_1 / nj
            except Exception as e:      # noqa: BLE001
                extra["value_numba_disabled"] = None
                extra["numba_disabled_error"] = str(e)[:200]
        cores, kind, warm = 1, "reference", args.warmup
        extra["reference_results"] = res
        extra["reference_root"] = os.path.relpath(ref_root, ROOT) if ref_root.startswith(ROOT) else ref_root
        sample = "per step: " + sample + "; unmodified reference through its CLI call sequence, 1 process"
    else:
        # no reference tree on this box: the oracle port on all cores (round-1 behaviour)
        from tao_amodal_b200 import synth
        cores = os.cpu_count() or 1
        cfg = synth.CONFIGS[args.workload]
        gt, dt = synth.generate_named(args.workload, videos=max(cores, 2), seed=cfg.seed)
        samples = cpu_samples(gt, dt, cores, REF_SAMPLE_FRAMES)
        warm = min(args.warmup, 1)
        for _ in range(warm):
            run_cpu_port(samples, cores)
        pairs, wall = 0, 0.0
        for _ in range(args.steps):
            p, w = run_cpu_port(samples, cores)
            pairs += p
            wall += w
        kind = "port"
        sample = "%d one-video samples per step (first %d of %d frames of %s-shaped videos), one per core" % (
            len(samples), REF_SAMPLE_FRAMES, cfg.frames, args.workload)
    value = pairs / wall
    cb = {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
    cb.update(extra)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, 1),
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cpu_baseline_leg(args, gt, dt, eng=None):
    """cpu_baseline of the main arm (rank 0, N = 1): the unmodified reference, one warm-up pass
    (numba JIT compile) + one timed pass over the fixed sample; the oracle port on all cores
    when the reference tree is absent.  With an engine the same sample also runs through this
    repo's pipeline and the summary metrics of both are compared (parity.reference_sample_*)."""
    import tempfile
    from oracle import ref_bench
    if ref_bench.find_reference() is not None:
        with tempfile.TemporaryDirectory() as td:
            ap, rp, pairs, sample = reference_sample(args.workload, td)
            ref_bench.run_once(ap, rp)
            s, ref_res = ref_bench.run_once(ap, rp)
            out = {"value": pairs / s, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": sample + "; unmodified reference (baseline/_ref), 1 process, %.1f s" % s}
            try:
                nj = ref_bench.run_subprocess(ap, rp, True, 1)[0]
                out["value_numba_disabled"] = pairs / nj
            except Exception:           # noqa: BLE001
                out["value_numba_disabled"] = None
            if eng is not None:
                from tao_amodal_b200 import engine, ingest, materialize, prep
                g, d = ingest.load_gt(ap), ingest.load_dt(rp)
                lp = prep.prepare_lvis(g, d)
                d2 = d.copy()
                prep.make_track_ids_unique(d2)
                tp = prep.prepare_tao(g, d2)
                o_t, o_l = eng.evaluate_host(tp), eng.evaluate_host(lp)
                mine = {
                    "tao_AP": float(materialize.summarize_tao(
                        o_t.precision.reshape(o_t.precision.shape[:3] + (5, 4)),
                        o_t.recall.reshape(o_t.recall.shape[:2] + (5, 4)), engine.IOU_THRS)["AP"]),
                    "lvis_AP": float(materialize.summarize_lvis(
                        o_l.precision, o_l.recall, engine.IOU_THRS, lp.freq_groups)["AP"])}
                out["parity"] = {"reference_sample_AP": ref_res, "ours_sample_AP": mine,
                                 "reference_sample_identical": bool(mine == ref_res)}
        return out
    cores = os.cpu_count() or 1
    samples = cpu_samples(gt, dt, cores, CPU_SAMPLE_FRAMES)
    pairs, wall = run_cpu_port(samples, cores)
    return {"value": pairs / wall, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d whole one-video samples of the same workload, one per core, %.1f s wall"
                      % (len(samples), wall)}


def workload_config(args, world):
    from tao_amodal_b200 import synth
    c = synth.CONFIGS[args.workload]
    return {"workload": "%s: %d videos x %d frames, %d predicted + %d GT tracks per video, %d "
                        "categories, 10 IoU thresholds; track-AP + frame-AP paths per step"
                        % (args.workload, args.videos or c.videos, c.frames, c.pred_tracks,
                           c.gt_tracks, c.categories),
            "per_gpu_videos": args.videos or c.videos, "world": world,
            "l2": "inputs (>0.6 GB per step) exceed the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------ main arm
OUT_KEYS = ("precision", "recall", "tp_cnt", "fp_cnt", "num_gt")


def plans_of(gt, dt):
    from tao_amodal_b200 import prep
    lvis_plan = prep.prepare_lvis(gt, dt, allow_empty=True)
    dt2 = dt.copy()
    prep.make_track_ids_unique(dt2)
    tao_plan = prep.prepare_tao(gt, dt2, vid_ids=np.unique(gt.vid_id), allow_empty=True)
    return tao_plan, lvis_plan


class Pipeline:
    """The bench step on one rank: both evaluators on resident plans, with the cross-rank
    exchange when world > 1."""

    STAGES = ["tao_iou", "tao_match", "tao_acc", "lvis_eval", "lvis_acc"]

    def __init__(self, eng, tao_plan, lvis_plan, transport=None):
        self.eng = eng
        self.d_tao, self.d_lvis = eng.upload(tao_plan), eng.upload(lvis_plan)
        self.exch = None
        if transport is not None:
            from tao_amodal_b200 import parallel
            self.exch = {id(d): parallel.DeviceExchange(eng, d, transport)
                         for d in (self.d_tao, self.d_lvis)}
        eng_, d_tao, d_lvis = eng, self.d_tao, self.d_lvis
        self.stages = [lambda: eng_.stage_iou(d_tao), lambda: eng_.stage_match(d_tao),
                       lambda: self.acc(d_tao), lambda: eng_.stage_frame_eval(d_lvis),
                       lambda: self.acc(d_lvis)]

    def acc(self, dev):
        if self.exch is None:
            self.eng.stage_accumulate(dev)
        else:
            self.exch[id(dev)].accumulate()

    def step(self, record=None):
        for k, fn in enumerate(self.stages):
            if record is not None:
                record[k][0].record()
            fn()
            if record is not None:
                record[k][1].record()

    def outputs(self, root_only=True):
        """The reference-layout tensors as numpy (after to_root when sharded); None off-root."""
        if self.exch is not None:
            for d in (self.d_tao, self.d_lvis):
                self.exch[id(d)].to_root()
            if root_only and self.exch[id(self.d_tao)].rank != 0:
                return None
        return {n + "_" + k: d.t[k].cpu().numpy().copy()
                for n, d in (("tao", self.d_tao), ("lvis", self.d_lvis)) for k in OUT_KEYS}


def timed(torch, dist, world, pipe, steps, warmup, eng, local, clock=True):
    """W warm-up + K timed steps: CUDA events on the launch stream around the region and around
    every stage, per-kernel events inside the library; barrier + synchronize on both sides."""
    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    for _ in range(warmup):
        pipe.step()
    sync()
    sampler = ClockSampler(local) if clock else None
    if sampler:
        sampler.start()
    ev = [[[torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
           for _ in pipe.STAGES] for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launches
    eng.timing(True)           # per-kernel CUDA events inside the library (ta_ctx_timing)
    e0.record()
    for s in range(steps):
        pipe.step(ev[s])
    e1.record()
    sync()
    kernel_ms = eng.timing_read()
    eng.timing(False)
    clocks = sampler.stop() if sampler else None
    stage_ms = {n: float(np.mean([ev[s][i][0].elapsed_time(ev[s][i][1]) for s in range(steps)]))
                for i, n in enumerate(pipe.STAGES)}
    return {"dev_ms": e0.elapsed_time(e1), "launches": eng.launches - l0, "kernel_ms": kernel_ms,
            "clocks": clocks, "stage_ms": stage_ms}


def bind_near_gpu(torch, local):
    """N > 1: run this rank on the CPUs of its GPU's NUMA node (sysfs local_cpulist of the PCI
    device), so that its pinned host buffers are first touched — and stay — in the memory next to
    the GPU's PCIe root instead of crossing the socket interconnect under eight-fold load.
    Returns a description for the JSON line; does nothing when the topology is not exposed."""
    try:
        pr = torch.cuda.get_device_properties(local)
        dev = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + dev
        node = int(open(base + "/numa_node").read().strip())
        if node < 0:
            return "unchanged (no NUMA node reported for %s)" % dev
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "unchanged (no usable CPU near %s)" % dev
        os.sched_setaffinity(0, cpus)
        return "NUMA node %d of GPU %s (%d CPUs)" % (node, dev, len(cpus))
    except Exception as e:          # noqa: BLE001 - topology files missing: leave the affinity alone
        return "unchanged (%s)" % type(e).__name__


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--videos", type=int, default=0, help="override videos per GPU (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true",
                    help="N > 1: skip the strong-scaling section (same set sharded + parity)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    affinity = bind_near_gpu(torch, local) if world > 1 else "unchanged"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from tao_amodal_b200.engine import Engine
    from tao_amodal_b200 import prep, synth
    t_prep = time.perf_counter()
    gt, dt, tao_plan, lvis_plan = make_workload(args.workload, rank, args.videos)
    pairs_trk = prep.count_box_pair_visits(tao_plan)
    pairs_img = prep.count_box_pair_visits(lvis_plan)
    pairs_local = pairs_trk + pairs_img
    t_prep = time.perf_counter() - t_prep

    eng = Engine(local)
    transport = None
    if world > 1:
        # the C ABI's own NCCL communicator (ta_exchange_*): rank 0 draws the id
        from tao_amodal_b200 import parallel
        box = [parallel.AbiTransport.unique_id(eng.lib) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        transport = parallel.AbiTransport(eng, rank, world, box[0])
    pipe = Pipeline(eng, tao_plan, lvis_plan, transport)
    d_tao, d_lvis = pipe.d_tao, pipe.d_lvis

    res = timed(torch, dist, world, pipe, args.steps, args.warmup, eng, local)
    dev_ms, kernel_ms, clocks, launches, stage_ms = (res["dev_ms"], res["kernel_ms"], res["clocks"],
                                                     res["launches"], res["stage_ms"])
    parity = {}
    single = None
    if world == 1 or rank == 0:
        # what was just timed, on this rank's own (whole) set, single GPU: the comparator of the
        # host-buffer call (N = 1) and of the sharded run of the same set (N > 1)
        if world > 1:
            eng.stage_iou(d_tao)
            eng.stage_match(d_tao)
            eng.stage_accumulate(d_tao)
            eng.stage_frame_eval(d_lvis)
            eng.stage_accumulate(d_lvis)
            torch.cuda.synchronize()
            single = {n + "_" + k: d.t[k].cpu().numpy().copy()
                      for n, d in (("tao", d_tao), ("lvis", d_lvis)) for k in OUT_KEYS}
        else:
            single = pipe.outputs()
    cells_with_gt = [0, 0]
    if single is not None:
        cells_with_gt = [int((single["tao_num_gt"] > 0).sum()), int((single["lvis_num_gt"] > 0).sum())]

    # ---- e2e: host buffers (pinned) through the C call(s), H2D + D2H inside the region
    e2e_steps = max(3, min(args.steps, 10))
    if world == 1:
        # transport form of both plans in page-locked memory (ta_host_alloc), built once; one
        # C call per step: both plans host -> device -> host
        pack = eng.pack_host([tao_plan, lvis_plan], pinned=True)
        outs = [pack.new_output(0), pack.new_output(1)]

        def e2e_step():
            eng.evaluate_pack(pack, outs)
            return (outs[0].h2d_bytes + outs[1].h2d_bytes, outs[0].d2h_bytes + outs[1].d2h_bytes)
    else:
        # transport form of both plans in page-locked memory, as at N = 1
        pack = eng.pack_host([tao_plan, lvis_plan], pinned=True)
        # pinned host slices for every owner's results: each rank copies its OWN category block out
        host_part = {id(dv): {k: torch.empty(v.shape, dtype=v.dtype).pin_memory()
                              for k, v in pipe.exch[id(dv)].part.items()} for dv in (d_tao, d_lvis)}

        # the resident pass' slices, to check the end-to-end route against; the device slices
        # are wiped so that only a step that really recomputes them can pass
        torch.cuda.synchronize()
        resident_part = {id(dv): {k: v.cpu().clone() for k, v in pipe.exch[id(dv)].part.items()}
                         for dv in (d_tao, d_lvis)}
        for dv in (d_tao, d_lvis):
            for v in pipe.exch[id(dv)].part.values():
                v.zero_()
        copy_in, copy_out = torch.cuda.Stream(), torch.cuda.Stream()
        aux = eng.aux_ctx()       # the upload stream's library calls run beside the main context's

        def e2e_step():
            # pinned host plans -> HBM on an upload stream (compact transport forms; the frame
            # plan's boxes first, the track plan takes its boxes from them), local IoU + matching,
            # cross-rank exchange of the result records, owner-side PR, every owner's slice back
            # to ITS host (pinned) on a third stream: the track plan's kernels, exchange and
            # download overlap the upload of the rest of the frame plan.  The frame schedule is
            # rebuilt from the fresh inputs every step; the exchange's routing tables are part
            # of the plan, like acc_perm, and stay resident.
            h2d = d2h = 0
            main = torch.cuda.current_stream()
            ready = {}
            with torch.cuda.stream(copy_in):
                shared = bool(pack.shared)
                if shared:
                    h2d += d_lvis.reload_pack(pack, 1, only=("dt_box",), ctx=aux)
                h2d += d_tao.reload_pack(pack, 0, pool=d_lvis if shared else None, ctx=aux)
                ready[id(d_tao)] = torch.cuda.Event()
                ready[id(d_tao)].record(copy_in)
                h2d += d_lvis.reload_pack(pack, 1, skip=("dt_box",) if shared else (), ctx=aux)
                ready[id(d_lvis)] = torch.cuda.Event()
                ready[id(d_lvis)].record(copy_in)
            for dev in (d_tao, d_lvis):
                main.wait_event(ready[id(dev)])
                if dev.plan.kind == "tao":
                    eng.stage_iou(dev)
                    eng.stage_match(dev)
                else:
                    eng.stage_frame_eval(dev)
                ex = pipe.exch[id(dev)]
                ex.accumulate()
                done = torch.cuda.Event()
                done.record(main)
                copy_out.wait_event(done)
                with torch.cuda.stream(copy_out):
                    for k, v in ex.part.items():
                        host_part[id(dev)][k].copy_(v, non_blocking=True)
                        d2h += v.numel() * v.element_size()
            torch.cuda.synchronize()
            return h2d, d2h
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h2d, d2h = e2e_step()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        same = all(torch.equal(host_part[i][k], resident_part[i][k])
                   for i in host_part for k in host_part[i])
        flag = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        e2e_same = bool(flag.item())
        parity["e2e_slices_equal_resident"] = e2e_same
        if not e2e_same:
            if rank == 0:
                sys.stderr.write("PARITY FAILURE: the end-to-end step's result slices differ from the resident pass'\n")
            dist.destroy_process_group()
            return 3
    if world == 1:
        # the host-buffer call and the resident route must agree bit for bit
        host_out = {n + "_" + k: getattr(o, k) for n, o in (("tao", outs[0]), ("lvis", outs[1]))
                    for k in OUT_KEYS}
        parity["host_call_equals_resident"] = bool(all(
            np.array_equal(host_out[k].reshape(single[k].shape), single[k]) for k in single))

    # ---- N > 1: the SAME set sharded per video over the ranks (BASELINE configs[3] / [4]):
    # strong-scaling time and bit-identity of the merged tensors with the 1-GPU ones
    strong = None
    if world > 1 and not args.no_strong:
        from tao_amodal_b200 import parallel
        from tao_amodal_b200.columnar import subset_videos
        cfg = synth.CONFIGS[args.workload]
        over = {"seed": cfg.seed}
        if args.videos:
            over["videos"] = args.videos
        t_s = time.perf_counter()
        gt0, dt0 = (gt, dt) if rank == 0 else synth.generate_named(args.workload, **over)
        shards = parallel.shard_videos(np.unique(gt0.vid_id), world)
        g_s, d_s = subset_videos(gt0, dt0, shards[rank])
        s_tao, s_lvis = plans_of(g_s, d_s)
        t_s = time.perf_counter() - t_s
        del pipe.exch
        spipe = Pipeline(eng, s_tao, s_lvis, transport)
        sres = timed(torch, dist, world, spipe, args.steps, args.warmup, eng, local, clock=False)
        merged = spipe.outputs()
        tt = torch.tensor([sres["dev_ms"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ok = torch.tensor([1], device="cuda", dtype=torch.int64)
        if rank == 0:
            ident = {k: bool(np.array_equal(merged[k], single[k])) for k in single}
            parity.update({"ranks": world, "identical": bool(all(ident.values())),
                           "tensors": ident, "compared_with": "the same %d-video set evaluated on one GPU"
                                                              % (args.videos or cfg.videos)})
            fx_path = os.path.join(ROOT, "tests", "golden", "full_%s_tao.npz" % args.workload)
            if os.path.exists(fx_path) and not args.videos:
                # configs[4]: the full TrackAP table against the unmodified reference's
                import hashlib
                from tao_amodal_b200 import engine, materialize
                fx = np.load(fx_path)
                prec = merged["tao_precision"]
                tab = materialize.summarize_tao(prec.reshape(prec.shape[:3] + (5, 4)),
                                                merged["tao_recall"].reshape(merged["tao_recall"].shape[:2] + (5, 4)),
                                                engine.IOU_THRS)
                vec = np.asarray([float(v) for v in tab.values()])
                parity["trackap_table_max_abs_diff_vs_reference"] = float(np.abs(vec - fx["tao_results"]).max())
                parity["trackap_precision_sha256_equal"] = bool(
                    hashlib.sha256(np.ascontiguousarray(prec).tobytes()).hexdigest() == str(fx["tao_precision_sha256"]))
                parity["track_AP"] = float(vec[0])
            ok[0] = 1 if parity["identical"] else 0
            strong_ms = float(tt[0]) / args.steps
            strong = {"scaling": "strong", "videos_total": int(np.unique(gt0.vid_id).size),
                      "ms_per_step": strong_ms, "value": pairs_local / (strong_ms * 1e-3),
                      "unit": UNIT, "stages_ms": sres["stage_ms"], "host_prep_s": t_s,
                      "note": "the rank-0 set of the weak run sharded per video over all ranks; "
                              "value = its box-pairs / max-over-ranks device time"}
        dist.broadcast(ok, src=0)
        if int(ok[0]) != 1:
            if rank == 0:
                sys.stderr.write("PARITY FAILURE: sharded tensors differ from the 1-GPU ones: %s\n" % parity)
            dist.destroy_process_group()
            return 3

    # ---- reduce over ranks: max time, sum units
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(t[0]), float(t[1])
        u = torch.tensor([pairs_local], device="cuda", dtype=torch.int64)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        pairs_total = int(u[0])
    else:
        pairs_total = pairs_local
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    ms_per_step = dev_ms / args.steps
    value = pairs_total / (ms_per_step * 1e-3)
    e2e_value = pairs_total / (e2e_s / e2e_steps)

    # ---- roofline of the dominant kernel (largest share of the step by summed launch time)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    ab_t, ab_l = algorithmic_bytes(tao_plan, cells_with_gt[0]), algorithmic_bytes(lvis_plan, cells_with_gt[1])
    # algorithmic bytes of each kernel per STEP (summed over its launches in one step)
    kernel_bytes = {
        "k_track_iou_tiled": ab_t["iou"], "k_match_greedy": ab_t["match"],
        "k_frame_flat": ab_l["frame_flat"], "k_frame_prep": ab_l["frame_prep"],
        "k_pr_count": ab_t["pr_count"] + ab_l["pr_count"],
        "k_pr_envelope": ab_t["pr_envelope"] + ab_l["pr_envelope"],
        "k_pr_bits": ab_t["pr_bits"] + ab_l["pr_bits"],
        "k_pr_envelope_bits": ab_t["pr_envelope_bits"] + ab_l["pr_envelope_bits"],
        "k_pr_finalize_tile": ab_t["pr_finalize"] + ab_l["pr_finalize"],
        "k_pr_finalize": ab_t["pr_finalize"] + ab_l["pr_finalize"],
    }
    per_kernel = {}
    for name, (ms_tot, n_launch) in kernel_ms.items():
        ms_step = ms_tot / args.steps
        b = kernel_bytes.get(name) if world == 1 else None
        per_kernel[name] = {"ms_per_step": ms_step, "launches_per_step": n_launch / args.steps,
                            "alg_bytes_per_step": b,
                            "gbs": (b / (ms_step * 1e-3) / 1e9) if (b and ms_step > 0) else None}
    ranked = [k for k in sorted(per_kernel, key=lambda k: -per_kernel[k]["ms_per_step"])
              if per_kernel[k]["gbs"] is not None]
    roofline = None
    if ranked:
        dom = ranked[0]
        achieved = per_kernel[dom]["gbs"]
        traffic = None
        tr_path = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr_path):
            tr = json.load(open(tr_path)).get(dom)
            if isinstance(tr, list):     # ncu dram bytes of the kernel's launches of one step
                traffic = float(sum(tr[:max(1, int(round(per_kernel[dom]["launches_per_step"])))]))
            else:
                traffic = tr
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak,
                    "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic,
                    "note": "achieved = algorithmic bytes of the kernel's launches in one step / their "
                            "summed CUDA-event time (events recorded by the library after every launch)",
                    "kernels": per_kernel,
                    "stages_ms": stage_ms}
    else:
        roofline = {"bound": "hbm", "kernels": per_kernel, "stages_ms": stage_ms,
                    "note": "per-kernel byte accounting is reported at N = 1"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, world), cpu_affinity=affinity),
        "box_pairs_per_step": pairs_total, "box_pairs_track_path": pairs_trk,
        "box_pairs_frame_path": pairs_img,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / e2e_steps,
                "api": ("Engine.evaluate_pack -> ta_eval_plans_host (both plans in one call: pinned host plans -> precision/recall on host; shared_boxes=%s)" % bool(pack.shared) if world == 1
                        else "per rank: DevicePlan.reload_pack (pinned host plans, shared boxes=%s) + stages + ta_exchange_* + owner-side PR + the owner's slice to its pinned host buffer" % bool(pack.shared))},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "parity": parity,
        "host_prep_s": t_prep,
    }
    if strong is not None:
        line["strong"] = strong
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_leg(args, gt, dt, eng)
        line["parity"].update(line["cpu_baseline"].pop("parity", {}))
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    # stdout carries exactly one JSON line: anything libraries print there (e.g. NCCL's version
    # banner) is diverted to stderr while the benchmark runs
    _real = os.dup(1)
    os.dup2(2, 1)
    _buf = []
    _print = print

    def print(*a, **k):          # noqa: A001 - bench-local capture of the JSON line
        _buf.append(" ".join(str(x) for x in a))

    try:
        rc = main()
    finally:
        sys.stdout.flush()
        os.dup2(_real, 1)
        os.close(_real)
    for line in _buf:
        _print(line, flush=True)
    sys.exit(rc)
