#!/usr/bin/env python
"""Wall-clock of the drop-in CLI on the headline set (BASELINE.json: 1 M predictions / 100 k GT
boxes, 5 000 videos, 1203 categories), JSON in -> metrics out.  The CPU side of the comparison
is bench.py's business (``cpu_baseline`` / ``--impl reference``: the only places outside tests/
that may run oracle/); SURVEY.md §3.4 estimates ~73 min for the unmodified reference on this
shape.

    python tools/bench_cli.py [--config cfg5] [--videos N] [--out gpurun_out/cli.json] [--cold]

--cold: what a user sees — every run is a FRESH `python tools/eval_on_tao_amodal.py ...`
subprocess, wall-clock around it: interpreter start, imports, CUDA context creation, both JSON
parses, host prep, GPU work and printing all included.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg5")
    ap.add_argument("--videos", type=int, default=0)
    ap.add_argument("--out", default="")
    ap.add_argument("--cold", action="store_true")
    ap.add_argument("--runs", type=int, default=3)
    ap.add_argument("--reference", action="store_true",
                    help="--cold: also time the UNMODIFIED reference CLI (baseline/_ref through "
                         "oracle/ref_shims.py) on the same two files, fresh process, 1 core")
    args = ap.parse_args()
    from tao_amodal_b200 import synth
    over = {"videos": args.videos} if args.videos else {}
    gt, dt = synth.generate_named(args.config, **over)
    td = tempfile.mkdtemp(prefix="ta_cli_")
    ap_, rp, lp = (os.path.join(td, f) for f in ("gt.json", "dt.json", "eval.log"))
    json.dump(gt.to_dict(), open(ap_, "w"))
    json.dump(dt.to_list(), open(rp, "w"))
    res = {"config": args.config, "videos": int(len(gt.vid_id)), "pred_boxes": int(dt.n()),
           "gt_boxes": int(gt.n_anns()), "annotation_mb": os.path.getsize(ap_) / 1e6,
           "prediction_mb": os.path.getsize(rp) / 1e6}

    if args.cold:
        import subprocess
        cli_py = os.path.join(ROOT, "tools", "eval_on_tao_amodal.py")
        runs = []
        for _ in range(args.runs):
            t0 = time.perf_counter()
            p = subprocess.run([sys.executable, cli_py, "--track_result", rp, "--output_log", lp,
                                "--annotation", ap_], capture_output=True, text=True)
            runs.append(time.perf_counter() - t0)
            if p.returncode != 0:
                res["error"] = p.stderr[-2000:]
                break
        res["cold_runs_s"] = runs
        res["cold_wall_s"] = min(runs) if runs else None
        res["cold_first_run_s"] = runs[0] if runs else None
        if os.path.exists(lp):
            res["log_tail"] = open(lp).read().strip().splitlines()[-1]
        res["torch_imported_by_cli"] = None
        chk = subprocess.run([sys.executable, "-c",
                              "import sys; sys.argv=['x','--track_result',%r,'--output_log',%r,'--annotation',%r];"
                              "sys.path.insert(0,%r); import contextlib, io\n"
                              "import eval_on_tao_amodal as c\n"
                              "with contextlib.redirect_stdout(io.StringIO()): c.main()\n"
                              "print('TORCH' if 'torch' in sys.modules else 'NOTORCH', file=sys.stderr)"
                              % (rp, lp, ap_, os.path.join(ROOT, "tools"))],
                             capture_output=True, text=True)
        res["torch_imported_by_cli"] = "NOTORCH" not in chk.stderr
        if args.reference:
            # the reference's own script, unmodified, in a fresh interpreter (its JSON parsing,
            # index building, numba JIT compile and both evaluators included)
            ref_log = os.path.join(td, "ref.log")
            code = ("import sys, os; sys.path.insert(0, %r)\n"
                    "from oracle import ref_bench, ref_shims\n"
                    "root = ref_bench.find_reference(); ref_shims.REF_ROOT = root\n"
                    "ref_shims.run_reference_driver(%r, %r, %r)\n" % (ROOT, ap_, rp, ref_log))
            t0 = time.perf_counter()
            p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
            res["reference_cold_wall_s"] = time.perf_counter() - t0
            if p.returncode != 0:
                res["reference_error"] = p.stderr[-1500:]
            else:
                same = open(ref_log).read().replace(ref_log, "") == open(lp).read().replace(lp, "")
                res["reference_log_identical"] = bool(same)
                res["speedup_cold_vs_reference"] = res["reference_cold_wall_s"] / res["cold_wall_s"]
        print(json.dumps(res))
        if args.out:
            json.dump(res, open(args.out, "w"), indent=1)
        return

    import eval_on_tao_amodal as cli
    import torch
    torch.cuda.init()
    from tao_amodal_b200.evaluation._common import _JSON_CACHE, get_engine
    get_engine(0)                                     # CUDA context creation is not the CLI's work
    runs = []
    from tao_amodal_b200 import ingest
    for _ in range(2):
        _JSON_CACHE.clear()
        ingest._CACHE.clear()            # every run parses both files from disk
        out = io.StringIO()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(out):
            cli.main(["--track_result", rp, "--output_log", lp, "--annotation", ap_])
        runs.append(time.perf_counter() - t0)
    res["cli_wall_s"] = min(runs)
    res["cli_runs_s"] = runs
    res["log_tail"] = open(lp).read().strip().splitlines()[-1]

    res["reference_estimate_s"] = 73 * 60 if args.config == "cfg5" and not args.videos else None
    print(json.dumps(res))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
