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


def _worker(rank, world, port, case, kind, q, only_video=None):
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
        if only_video is not None:
            # predictions of ONE video only: every other shard is prediction-free
            keep = dt.video_id == only_video
            for f in ("image_id", "track_id", "category_id", "video_id", "bbox", "score"):
                setattr(dt, f, np.ascontiguousarray(getattr(dt, f)[keep]))
        shards = parallel.shard_videos(np.unique(gt.vid_id), world)
        g_s, d_s = subset_videos(gt, dt, shards[rank])
        if kind == "tao":
            plan = prep.prepare_tao(g_s, d_s, vid_ids=shards[rank], allow_empty=True)
        else:
            plan = prep.prepare_lvis(g_s, d_s, img_ids=g_s.img_id, allow_empty=True)
        local = run_hostsim(plan)                       # per-rank IoU + matching
        tr = parallel.TorchTransport(rank, world, torch.device("cpu"))
        ex = parallel.ExchangePlan(plan, tr, torch.device("cpu"))
        rows = ex.exchange(torch.from_numpy(np.ascontiguousarray(local.dt_tpfp).view(np.int32)))
        num_gt, num_gt_own = ex.global_num_gt(torch.from_numpy(local.num_gt.copy()))
        _check_peer_window_plans(dist, ex, plan, local, rows, num_gt)
        # owner-side PR through the emulation of the shipped (bit-plane) kernels: categories
        # without detections and owners without categories included
        T, R, K = 10, 101, plan.n_cfg
        if ex.n_loc:
            out = hostsim_pr(ex.n_loc, ex.cat_dt_off.numpy(), ex.acc_perm.numpy(),
                             rows.numpy().view(np.uint32).reshape(-1, K),
                             np.ascontiguousarray(num_gt_own.numpy()), K, impl="bits_tile")
            parts = [torch.from_numpy(x) for x in (out.precision, out.recall, out.tp_cnt, out.fp_cnt)]
        else:
            parts = [torch.empty((T, R, 0, K), dtype=torch.float64), torch.empty((T, 0, K), dtype=torch.float64),
                     torch.empty((T, 0, K), dtype=torch.int64), torch.empty((T, 0, K), dtype=torch.int64)]
        C_ = len(plan.cat_ids)
        pr, rc = torch.empty((T, R, C_, K), dtype=torch.float64), torch.empty((T, C_, K), dtype=torch.float64)
        tp, fp = torch.empty((T, C_, K), dtype=torch.int64), torch.empty((T, C_, K), dtype=torch.int64)
        ex.gather_to_root(parts, [pr, rc, tp, fp])
        if rank == 0:
            if only_video is None:
                want = {k: g[kind + "_" + k] for k in ("precision", "recall", "tp_cnt", "fp_cnt")}
            else:
                # single-process plan of the same reduced input through the same emulation
                if kind == "tao":
                    whole = prep.prepare_tao(gt, dt)
                else:
                    whole = prep.prepare_lvis(gt, dt)
                w = run_hostsim(whole)
                want = {"precision": w.precision, "recall": w.recall, "tp_cnt": w.tp_cnt, "fp_cnt": w.fp_cnt}
            got = {"precision": pr, "recall": rc, "tp_cnt": tp, "fp_cnt": fp}
            ok = all(np.array_equal(np.asarray(want[k]).reshape(got[k].shape), got[k].numpy()) for k in got)
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


def _check_peer_window_plans(dist, ex, plan, local, rows_nccl, num_gt_sum):
    """The peer-window route on CPU: every rank fills a window laid out by
    ExchangePlan.window_plan, the windows are made visible to everybody (an all-gather stands in
    for the CUDA IPC mapping), and applying the owner's pull plan must reproduce what the
    all-to-all route delivered — full rows (track path), and words + flagged rows (frame path,
    with every third local detection flagged)."""
    import torch
    K, n_dt = plan.n_cfg, plan.n_dt
    tpfp = np.ascontiguousarray(local.dt_tpfp).view(np.int32).reshape(n_dt, K)

    def windows_of(lay, fill):
        w = np.zeros(lay["bytes"], dtype=np.uint8)
        w[lay["numgt"]:lay["numgt"] + 4 * lay["numgt_count"]] = local.num_gt.astype(np.int32).reshape(-1).view(np.uint8)
        fill(w)
        box = [None] * ex.world
        dist.all_gather_object(box, w)
        return box

    def pull(lay, wins, sizes):
        out = {k: np.zeros(n, dtype=np.uint8) for k, n in sizes.items()}
        for peer, off, nb, kind, dst in lay["copies"]:
            assert off % 4 == 0 and nb % 4 == 0 and off + nb <= wins[peer].size
            out[kind][dst:dst + nb] = wins[peer][off:off + nb]
        tot = sum(w[lay["numgt"]:lay["numgt"] + 4 * lay["numgt_count"]].view(np.int32).astype(np.int64)
                  for w in wins)
        return out, tot

    # track-path form: one full row per detection
    lay = ex.window_plan(4 * K)

    def fill_rows(w):
        w[lay["rec"]:lay["rec"] + tpfp.nbytes] = tpfp.reshape(-1).view(np.uint8)
    got, tot = pull(lay, windows_of(lay, fill_rows), {"rec": ex.n_recv * 4 * K})
    assert np.array_equal(got["rec"].view(np.int32).reshape(-1, K), rows_nccl.numpy().reshape(-1, K))
    assert np.array_equal(tot, num_gt_sum.numpy().reshape(-1))

    # frame-path form: a 4-byte word per detection + the rows of the flagged ones
    words = (np.arange(n_dt, dtype=np.int32) * 8 + ex.rank).astype(np.int32)
    flag_idx = np.arange(0, n_dt, 3, dtype=np.int64)
    send, recv = ex.flag_routing(flag_idx)
    lay = ex.window_plan(4, (send, recv, 4 * K))
    flag_rows = np.ascontiguousarray(tpfp[flag_idx])

    def fill_words(w):
        w[lay["rec"]:lay["rec"] + words.nbytes] = words.view(np.uint8)
        w[lay["flag"]:lay["flag"] + flag_rows.nbytes] = flag_rows.reshape(-1).view(np.uint8)
    got, tot = pull(lay, windows_of(lay, fill_words), {"rec": ex.n_recv * 4, "flag": int(sum(recv)) * 4 * K})
    want_words = ex.exchange(torch.from_numpy(words)).numpy()
    assert np.array_equal(got["rec"].view(np.int32), want_words)
    want_rows = ex.tr.all_to_all(torch.from_numpy(flag_rows), send, recv).numpy()
    assert np.array_equal(got["flag"].view(np.int32).reshape(-1, K), want_rows.reshape(-1, K))
    assert np.array_equal(tot, num_gt_sum.numpy().reshape(-1))


def _run(world, case, kind, **kw):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, kind, q), kwargs=kw)
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


@pytest.mark.parametrize("kind", ["tao", "lvis"])
@pytest.mark.parametrize("case", ["small", "small_ties", "edge_mix"])
def test_two_rank_exchange_matches_reference(case, kind):
    _run(2, case, kind)


@pytest.mark.parametrize("kind", ["tao", "lvis"])
def test_prediction_free_shard_takes_part_in_the_exchange(kind):
    """Only one video has predictions: the other rank's shard has ground truth but nothing to
    send; the merged result equals the single-process one (no deadlock, no per-shard error)."""
    _run(2, "small", kind, only_video=1)


@pytest.mark.parametrize("kind", ["tao", "lvis"])
def test_more_ranks_than_videos(kind):
    """world = 4 over the 3 videos of `tiny`: one rank holds no video at all."""
    _run(4, "tiny", kind)


def test_balanced_blocks_cover_and_balance():
    from tao_amodal_b200.parallel import balanced_blocks
    w = np.array([0, 5, 0, 0, 7, 1, 1, 1, 9, 0, 3, 0])
    for world in (1, 2, 3, 5, 20):
        b = balanced_blocks(w, world)
        assert b[0] == 0 and b[-1] == w.size and (np.diff(b) >= 0).all() and b.size == world + 1
    b = balanced_blocks(w, 3)
    loads = [w[b[i]:b[i + 1]].sum() for i in range(3)]
    assert max(loads) <= 16
    assert (np.diff(balanced_blocks(np.zeros(6), 3)) == 2).all()


def test_shard_videos_balances_and_covers():
    from tao_amodal_b200.parallel import shard_videos
    vids = np.arange(1, 12)
    w = np.array([5, 1, 1, 1, 9, 1, 1, 1, 1, 1, 4], dtype=float)
    bins = shard_videos(vids, 3, w)
    assert sorted(np.concatenate(bins).tolist()) == vids.tolist()
    loads = [w[np.isin(vids, b)].sum() for b in bins]
    assert max(loads) - min(loads) <= 2
