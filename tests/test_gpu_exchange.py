"""The C ABI's exchange entry points (ta_exchange_*, csrc/ta_exchange.cu) on ONE GPU: a
single-rank communicator takes the same code path as N ranks — NCCL loaded with dlopen, the
library's own communicator, zero-copy sends of the owner slices (here: the self segment), the
sparse full rows through k_xchg_gather / k_xchg_scatter, owner-side ta_pr_accumulate on the
received records — and must reproduce the local accumulation bit for bit.  The N > 1 routing is
covered on CPU by tests/test_parallel_gloo.py and on 2 / 4 / 8 GPUs by bench.py's
`parity.identical` (profiles/r2_bench_n*.json), which runs the peer-window route."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_peer_window_copies_any_alignment():
    """k_peer_exchange's copy loops: every combination of source / destination misalignment
    and lengths around the 16-byte vector width, and the sum over ranks."""
    import ctypes as C
    import torch
    from tao_amodal_b200 import parallel
    from tao_amodal_b200.engine import Engine
    eng = Engine(0)
    tr = parallel.AbiTransport(eng, 0, 1, parallel.AbiTransport.unique_id(eng.lib))
    try:
        n_words = 1 << 16
        w = parallel.PeerWindow(tr, 4 * n_words)
        src = torch.arange(n_words, dtype=torch.int32, device="cuda") * 7 + 3
        dst = torch.zeros(n_words + 64, dtype=torch.int32, device="cuda")
        total = torch.zeros(100, dtype=torch.int32, device="cuda")
        cases = [(s_off, d_off, n) for s_off in (0, 1, 2, 3, 5) for d_off in (0, 1, 2, 3)
                 for n in (0, 1, 3, 4, 5, 17, 1000, 40001)]
        for rnd in range(2):
            for s_off, d_off, n in cases:
                dst.zero_()
                w.set_plan([(0, 4 * s_off, 4 * n, dst.data_ptr() + 4 * d_off)],
                           [(4 * 8, 100, total.data_ptr())])
                w.acquire()
                w.put(0, src, 4 * n_words)
                w.exchange()
                torch.cuda.synchronize()
                assert torch.equal(dst[d_off:d_off + n], src[s_off:s_off + n]), (s_off, d_off, n)
                assert int(dst[:d_off].abs().sum()) == 0 and int(dst[d_off + n:].abs().sum()) == 0
                assert torch.equal(total, src[8:108])
        assert not w.timed_out()
        w.close()
    finally:
        tr.close()
        eng.close()
@pytest.mark.parametrize("route", ["peer", "nccl"])
@pytest.mark.parametrize("case", ["small", "edge_mix"])
def test_single_rank_exchange_equals_local_accumulate(case, route, monkeypatch):
    """route: "peer" = the library's own exchange kernel on peer windows (ta_peer_window_*; with
    one rank the self slices, through the same copy loops incl. unaligned heads and tails),
    "nccl" = grouped ncclSend / ncclRecv."""
    import torch
    monkeypatch.setenv("TA_XCHG", route)
    from conftest import golden_inputs, load_golden
    from plan_backends import plans_from_json
    from tao_amodal_b200 import parallel
    from tao_amodal_b200.engine import Engine
    g = load_golden(case)
    gt, res = golden_inputs(g)
    tao_plan, lvis_plan = plans_from_json(gt, res)
    eng = Engine(0)
    tr = parallel.AbiTransport(eng, 0, 1, parallel.AbiTransport.unique_id(eng.lib))
    try:
        for plan in (tao_plan, lvis_plan):
            ref = eng.evaluate_device(eng.upload(plan))             # local route
            dev = eng.upload(plan)
            ex = parallel.DeviceExchange(eng, dev, tr)
            assert ex.n_loc == len(plan.cat_ids) and ex.n_recv == plan.n_dt
            assert (ex.window is not None) == (route == "peer")
            if plan.kind == "tao":
                eng.stage_iou(dev)
                eng.stage_match(dev)
            else:
                eng.stage_frame_eval(dev)
                assert ex.compact and dev.words_valid
            for _ in range(3):          # epochs: acquire / publish / release more than once
                ex.accumulate()
            ex.to_root()
            torch.cuda.synchronize()
            if ex.window is not None:
                assert not ex.window.timed_out()
            for k in ("precision", "recall", "tp_cnt", "fp_cnt", "num_gt"):
                assert np.array_equal(dev.t[k].cpu().numpy(), getattr(ref, k)), (plan.kind, k)
            shape = g[plan.kind + "_precision"].shape
            assert np.array_equal(g[plan.kind + "_precision"],
                                  dev.t["precision"].cpu().numpy().reshape(shape))
    finally:
        tr.close()
        eng.close()
