#!/usr/bin/env python
"""Drop-in replacement of the reference's ``tools/eval_on_tao_amodal.py``: same flags, same
log-file / console / stdout lines, evaluated on the GPU.

    python tools/eval_on_tao_amodal.py --track_result P --output_log L --annotation A

Differences that do not change any output: runs from any working directory, parses each JSON
file once (the reference parses both twice), ``--annotation`` has no machine-specific default,
``--device`` (additive) selects the GPU.

Multi-GPU: launch with torchrun, e.g.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
        tools/eval_on_tao_amodal.py --track_result P --output_log L --annotation A
Every rank evaluates its shard of the videos on its own GPU; rank 0 writes the log and stdout.
"""
import argparse
import logging
import os
import sys
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tao_amodal_b200.columnar import DtColumns                      # noqa: E402
from tao_amodal_b200.evaluation._common import BackgroundPlan, load_json, warm_engine   # noqa: E402
from tao_amodal_b200.evaluation.lvis_amodal import LVISEval        # noqa: E402
from tao_amodal_b200.evaluation.tao_amodal import Tao, TaoEval     # noqa: E402
from tao_amodal_b200 import ingest, prep                           # noqa: E402

BBOX_METRICS = ["AP", "AP50", "AP75", "AP-HO", "AP50-HO", "AP75-HO", "AP-PO", "AP50-PO",
                "AP75-PO", "AP-HV", "AP50-HV", "AP75-HV", "AP-OOF", "AP50-OOF", "AP75-OOF",
                "AP-HP", "AP50-HP", "AP75-HP", "APr", "APc", "APf"]


def create_small_table(small_dict):
    """detectron2.utils.logger.create_small_table (the only detectron2 function the reference
    CLI calls, tools/eval_on_tao_amodal.py:20,110)."""
    try:
        from detectron2.utils.logger import create_small_table as d2
        return d2(small_dict)
    except ImportError:
        from tabulate import tabulate
        keys, values = tuple(zip(*small_dict.items()))
        return tabulate([values], headers=keys, tablefmt="pipe", floatfmt=".3f",
                        stralign="center", numalign="center")


def make_track_ids_unique(result_anns):
    """tools/eval_on_tao_amodal.py:44-66 for a list of result dicts (in place)."""
    first_video, clash, top = {}, set(), 0
    for r in result_anns:
        t = r['track_id']
        first_video.setdefault(t, r['video_id'])
        if r['video_id'] != first_video[t]:
            clash.add(t)
        top = max(top, t)
    if clash:
        fresh, nxt = {}, top + 1
        for r in result_anns:
            k = (r['track_id'], r['video_id'])
            if k[0] in clash:
                if k not in fresh:
                    fresh[k] = nxt
                    nxt += 1
                r['track_id'] = fresh[k]
    return len(clash)


def evaluate_predictions_on_lvis(lvis_gt, lvis_results, iou_type, logger, device=0):
    """_custom_evaluate_predictions_on_lvis, tools/eval_on_tao_amodal.py:68-116."""
    metrics = BBOX_METRICS
    if len(lvis_results) == 0:
        logger.warn("No predictions from the model!")
        return {metric: float("nan") for metric in metrics}
    logger.info('Evaluating {} on LVIS...'.format(lvis_results))
    lvis_eval = LVISEval(lvis_gt, lvis_results, iou_type, device=device)
    lvis_eval.run()
    lvis_eval.print_results()
    results = lvis_eval.get_results()
    results = {metric: float(results[metric] * 100) for metric in metrics}
    logger.info("Evaluation results for {}: \n".format(iou_type) + create_small_table(results))
    important = [(metric, results[metric]) for metric in metrics]
    logger.info("copypaste: " + ",".join([k[0] for k in important]))
    logger.info("copypaste: " + ",".join(["{0:.4f}".format(k[1]) for k in important]))
    return results


def eval_tao_track(ann_path, results_path, logger, device=0, background=None):
    """tools/eval_on_tao_amodal.py:118-151.  `background`: a BackgroundPlan that has been
    building the track plan of these two files while the frame evaluation ran."""
    logger.setLevel(logging.INFO)
    results = {}
    logger.info("Loading gt {}...".format(ann_path))
    tao_gt = Tao(ann_path)
    logger.info('Done')
    logger.info('Loading results...')
    bg_gt, dt, plan = background.take() if background is not None else (None, None, None)
    if plan is None or bg_gt is not tao_gt.columns:
        plan = None
        dt = ingest.load_dt(results_path).copy()  # native reader; the cached columns stay intact
        prep.make_track_ids_unique(dt)
    logger.info('Done')
    logger.info('Building')
    tao_eval = TaoEval(tao_gt, dt, logger=logger, device=device, _plan=plan)
    logger.info('Done')
    tao_eval.run()
    tao_eval.print_results()
    res = tao_eval.get_results()
    results["TAO 3DmAP50"] = res["AP50"] * 100
    results["TAO 3DmAP50-HP"] = res["AP50-HP"] * 100
    results["TAO 3DmAP"] = res["AP"] * 100
    results["TAO 3DmAP-HP"] = res["AP-HP"] * 100
    logger.info("TAO 3DmAP50:{:.4f}".format(results["TAO 3DmAP50"]))
    logger.info("TAO 3DmAP50-HP:{:.4f}".format(results["TAO 3DmAP50-HP"]))
    logger.info("TAO 3DmAP:{:.4f}".format(results["TAO 3DmAP"]))
    logger.info("TAO 3DmAP-HP:{:.4f}".format(results["TAO 3DmAP-HP"]))
    keys = ["TAO 3DmAP50", "TAO 3DmAP50-HP", "TAO 3DmAP", "TAO 3DmAP-HP"]
    logger.info("copypaste: " + ",".join(keys))
    logger.info("copypaste: " + ",".join(["{:.4f}".format(results[k]) for k in keys]))
    return results


def main(argv=None):
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('--track_result', type=str, required=True)
    parser.add_argument('--output_log', type=str, required=True)
    parser.add_argument('--annotation', type=str, default=None)
    parser.add_argument('--device', type=int, default=0, help="CUDA device (additive flag)")
    args = parser.parse_args(argv)
    if not args.annotation:
        parser.error("--annotation is required (the reference's default is a path on its "
                     "authors' machine, tools/eval_on_tao_amodal.py:39)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    device = args.device
    if world > 1:
        import torch
        import torch.distributed as dist
        device = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(device)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    output_log = Path(args.output_log)
    logger = logging.getLogger("__main__")
    handler = None
    if rank == 0:
        logger.setLevel(logging.INFO)
        output_log.parent.mkdir(parents=True, exist_ok=True)
        handler = logging.FileHandler(output_log, mode='w')
        logger.addHandler(handler)
    else:           # other ranks compute their shard silently
        logging.disable(logging.CRITICAL)
        sys.stdout = open(os.devnull, "w")
    # the results file is parsed on a background thread while the annotation file is read; the
    # CUDA context is created meanwhile; single-GPU runs also build the track evaluator's plan
    # in the background while the frame evaluator works
    background = None
    if isinstance(args.track_result, str) and os.path.exists(args.track_result):
        ingest.prefetch(args.track_result, "dt")
        if world == 1:
            warm_engine(device)
            if os.path.exists(args.annotation):
                background = BackgroundPlan(args.annotation, args.track_result)
    try:
        evaluate_predictions_on_lvis(args.annotation, args.track_result, "bbox", logger,
                                     device=device)
        eval_tao_track(args.annotation, args.track_result, logger, device=device,
                       background=background)
    finally:
        if handler is not None:
            handler.close()
            logger.removeHandler(handler)
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
