"""world_size-2 test of the multi-GPU exchange (tao_amodal_b200/parallel.py) on CPU with the
gloo backend: each rank evaluates its shard of videos with the host simulation of the kernels,
records are exchanged by category owner, and rank 0's merged precision / recall / counts must
equal the unmodified reference's result on the whole dataset."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, kind, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import golden_inputs, load_golden
    from plan_backends import hostsim_pr, run_hostsim
    from tao_amodal_b200 import parallel, prep
    from tao_amodal_b200.columnar import DtColumns, GtColumns, subset_videos
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank,
                            world_size=world)
    try:
        g = load_golden(case)
        gt_d, res = golden_inputs(g)
        gt, dt = GtColumns.from_dict(gt_d), DtColumns.from_list(res)
        if kind == "tao":
            prep.make_track_ids_unique(dt)
        shards = parallel.shard_videos(np.unique(gt.vid_id), world)
        g_s, d_s = subset_videos(gt, dt, shards[rank])
        plan = prep.prepare_tao(g_s, d_s) if kind == "tao" else prep.prepare_lvis(g_s, d_s)
        local = run_hostsim(plan)                       # per-rank IoU + matching
        acc = parallel.DistAccumulator(plan, rank, world, torch.device("cpu"))
        rows = acc.exchange_tpfp(torch.from_numpy(local.dt_tpfp.view(np.int32)))
        num_gt, num_gt_own = acc.global_num_gt(torch.from_numpy(local.num_gt))
        # owner-side PR through the emulation of the shipped (bit-plane) kernels: padded empty
        # categories and categories without detections included
        out = hostsim_pr(acc.n_loc, acc.cat_dt_off.numpy(), acc.acc_perm.numpy(),
                         rows.numpy().view(np.uint32), num_gt_own.numpy(), plan.n_cfg,
                         impl="bits_tile")
        parts = [torch.from_numpy(x) for x in (out.precision, out.recall, out.tp_cnt, out.fp_cnt)]
        C_, K = len(plan.cat_ids), plan.n_cfg
        pr, rc = torch.empty((10, 101, C_, K), dtype=torch.float64), torch.empty((10, C_, K), dtype=torch.float64)
        tp, fp = torch.empty((10, C_, K), dtype=torch.int64), torch.empty((10, C_, K), dtype=torch.int64)
        acc.merge_to_root(parts, [pr, rc, tp, fp])
        if rank == 0:
            shape = g[kind + "_precision"].shape
            ok = (np.array_equal(g[kind + "_precision"], pr.numpy().reshape(shape))
                  and np.array_equal(g[kind + "_recall"], rc.numpy().reshape(g[kind + "_recall"].shape))
                  and np.array_equal(g[kind + "_tp_cnt"], tp.numpy().reshape(g[kind + "_tp_cnt"].shape))
                  and np.array_equal(g[kind + "_fp_cnt"], fp.numpy().reshape(g[kind + "_fp_cnt"].shape)))
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["tao", "lvis"])
@pytest.mark.parametrize("case", ["small", "small_ties", "edge_mix"])
def test_two_rank_exchange_matches_reference(case, kind):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_videos_balances_and_covers():
    from tao_amodal_b200.parallel import shard_videos
    vids = np.arange(1, 12)
    w = np.array([5, 1, 1, 1, 9, 1, 1, 1, 1, 1, 4], dtype=float)
    bins = shard_videos(vids, 3, w)
    assert sorted(np.concatenate(bins).tolist()) == vids.tolist()
    loads = [w[np.isin(vids, b)].sum() for b in bins]
    assert max(loads) - min(loads) <= 2
