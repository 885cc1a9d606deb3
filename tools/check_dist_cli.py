#!/usr/bin/env python
"""Multi-GPU CLI check: the torchrun run on N GPUs must write the same log file and stdout as the
single-GPU run (bit-identical metrics).  Needs N GPUs on one node.

    python tools/check_dist_cli.py [--gpus 2] [--config cfg2] [--out gpurun_out/dist_cli.json]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    from tao_amodal_b200 import synth
    gt, dt = synth.generate_named(args.config)
    td = tempfile.mkdtemp(prefix="ta_dist_")
    a, p = os.path.join(td, "gt.json"), os.path.join(td, "dt.json")
    json.dump(gt.to_dict(), open(a, "w"))
    json.dump(dt.to_list(), open(p, "w"))
    cli = os.path.join(ROOT, "tools", "eval_on_tao_amodal.py")
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    res = {}
    outs = {}
    for name, cmd in (
            ("single", [sys.executable, cli]),
            ("dist", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                      "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
                      "--master-port", "29531", cli])):
        log = os.path.join(td, name + ".log")
        t0 = time.perf_counter()
        r = subprocess.run(cmd + ["--track_result", p, "--output_log", log, "--annotation", a],
                           env=env, capture_output=True, text=True)
        res[name + "_wall_s"] = time.perf_counter() - t0
        if r.returncode != 0:
            print(r.stderr[-3000:])
            raise SystemExit("%s run failed" % name)
        # NCCL prints its version banner on stdout at communicator creation
        so = "\n".join(l for l in r.stdout.splitlines() if not l.startswith("NCCL version"))
        outs[name] = (open(log).read(), so)
    same_log = outs["single"][0] == outs["dist"][0]
    same_out = outs["single"][1] == outs["dist"][1]
    res.update(gpus=args.gpus, config=args.config, log_identical=same_log, stdout_identical=same_out,
               last_line=outs["dist"][0].strip().splitlines()[-1])
    print(json.dumps(res))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)
    if not (same_log and same_out):
        import difflib
        for k in (0, 1):
            print("\n".join(list(difflib.unified_diff(outs["single"][k].splitlines(),
                                                       outs["dist"][k].splitlines(), lineterm=""))[:40]))
        raise SystemExit(1)


if __name__ == "__main__":
    main()
