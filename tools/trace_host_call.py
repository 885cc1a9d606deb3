"""Timeline of the host-buffer call (ta_eval_plans_host) at the bench workload: sets
TA_PIPE_TRACE=1, packs both cfg3 plans in page-locked memory and runs the call a few times; the
library prints, per plan, when its uploads, kernels and downloads finished (ms since the first
upload).  Also times a plain pinned H2D / D2H copy of the same byte counts for comparison.

    python tools/trace_host_call.py [--videos N] [--runs K]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=0)
    ap.add_argument("--runs", type=int, default=4)
    ap.add_argument("--workload", default="cfg3")
    args = ap.parse_args()
    import torch
    import bench
    from tao_amodal_b200.engine import Engine
    gt, dt, tao_plan, lvis_plan = bench.make_workload(args.workload, 0, args.videos)
    eng = Engine(0)
    pack = eng.pack_host([tao_plan, lvis_plan], pinned=True)
    outs = [pack.new_output(0), pack.new_output(1)]
    for _ in range(2):
        eng.evaluate_pack(pack, outs)
    os.environ["TA_PIPE_TRACE"] = "1"
    for k in range(args.runs):
        t0 = time.perf_counter()
        eng.evaluate_pack(pack, outs)
        print("call %d: %.2f ms wall" % (k, 1e3 * (time.perf_counter() - t0)), file=sys.stderr, flush=True)
    del os.environ["TA_PIPE_TRACE"]
    h2d = sum(o.h2d_bytes for o in outs)
    d2h = sum(o.d2h_bytes for o in outs)
    src = torch.empty(h2d, dtype=torch.uint8).pin_memory()
    dst = torch.empty(h2d, dtype=torch.uint8, device="cuda")
    back = torch.empty(d2h, dtype=torch.uint8).pin_memory()
    for name, fn in (("H2D %d MB" % (h2d // 10**6), lambda: dst.copy_(src, non_blocking=True)),
                     ("D2H %d MB" % (d2h // 10**6), lambda: back.copy_(dst[:d2h], non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / 3
        print("plain pinned copy, %s: %.2f ms" % (name, ms), file=sys.stderr)
    eng.close()


if __name__ == "__main__":
    main()
