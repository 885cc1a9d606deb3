"""A/B of the runtime-selectable kernel variants on one GPU, at the bench workload (cfg3).

    python tools/ab_variants.py [--videos N] [--steps K] [--out gpurun_out/ab_variants.json]

Builds the plans once, then for every value of
    TA_PR_IMPL   0 position-walk kernels | 1 bit-plane kernels (csrc/ta_pr.cu)
runs warm-up + K timed steps (CUDA events around the steps, per-kernel events inside the
library) and compares EVERY output tensor of both evaluators bit for bit with the baseline
variant 0, which the reference goldens pin — a full-size parity check of the variants.
Writes one JSON document; prints a short table.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--variants", default="0,1", help="TA_PR_IMPL values; the first is the baseline")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ab_variants.json"))
    args = ap.parse_args()
    import torch
    import bench
    from tao_amodal_b200 import prep
    from tao_amodal_b200.engine import Engine
    t0 = time.perf_counter()
    gt, dt, tao_plan, lvis_plan = bench.make_workload("cfg3", 0, args.videos)
    pairs = prep.count_box_pair_visits(tao_plan) + prep.count_box_pair_visits(lvis_plan)
    prep_s = time.perf_counter() - t0
    eng = Engine(0)
    d_tao, d_lvis = eng.upload(tao_plan), eng.upload(lvis_plan)

    def list_count():
        import ctypes as C
        n = C.c_int32(0)
        st = C.c_void_p(torch.cuda.current_stream(0).cuda_stream)
        eng.lib.ta_ctx_debug_list_count(eng._ctx, st, C.byref(n))
        return int(n.value)

    eng.stage_iou(d_tao)
    eng.stage_match(d_tao)
    n_list_tao = list_count()
    eng.stage_frame_eval(d_lvis)
    n_list_lvis = list_count()
    print("groups handed to the general matcher: track path %d of %d, frame path %d of %d" % (
        n_list_tao, tao_plan.n_groups, n_list_lvis, lvis_plan.n_groups), flush=True)

    def step():
        eng.stage_iou(d_tao)
        eng.stage_match(d_tao)
        eng.stage_accumulate(d_tao)
        eng.stage_frame_eval(d_lvis)
        eng.stage_accumulate(d_lvis)

    def outputs():
        out = {}
        for name, dv in (("tao", d_tao), ("lvis", d_lvis)):
            for k in ("precision", "recall", "tp_cnt", "fp_cnt", "num_gt"):
                out[name + "_" + k] = dv.t[k].cpu().numpy().copy()
            out[name + "_dt_tpfp"] = dv.t["dt_tpfp"].cpu().numpy().copy()
        return out

    results, base = {}, None
    for pr in [int(x) for x in args.variants.split(',')]:
        os.environ["TA_PR_IMPL"] = str(pr)
        key = "pr%d" % pr
        rec = {"TA_PR_IMPL": pr}
        try:
            # poison the outputs so that a kernel that writes nothing cannot pass
            for dv in (d_tao, d_lvis):
                dv.t["precision"].fill_(123.0)
                dv.t["recall"].fill_(123.0)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            out = outputs()
            if base is None:
                base = out
            rec["identical_to_baseline"] = {k: bool(np.array_equal(base[k], v)) for k, v in out.items()}
            rec["all_identical"] = all(rec["identical_to_baseline"].values())
            rec["sha1_precision"] = {n: hashlib.sha1(out[n + "_precision"].tobytes()).hexdigest()[:16]
                                     for n in ("tao", "lvis")}
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.timing(True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            kms = eng.timing_read()
            eng.timing(False)
            rec["ms_per_step"] = e0.elapsed_time(e1) / args.steps
            rec["box_pairs_per_s"] = pairs / (rec["ms_per_step"] * 1e-3)
            rec["kernels_ms_per_step"] = {k: v[0] / args.steps for k, v in sorted(kms.items())}
        except Exception as e:          # noqa: BLE001 - record the failure and go on
            rec["error"] = "%s: %s" % (type(e).__name__, e)
        results[key] = rec
        print(key, json.dumps({k: rec.get(k) for k in ("ms_per_step", "all_identical", "error")}), flush=True)
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump({"workload": "cfg3", "videos": args.videos or 500, "box_pairs_per_step": pairs,
                   "host_prep_s": prep_s, "steps": args.steps, "variants": results},
                  open(args.out, "w"), indent=1)
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
