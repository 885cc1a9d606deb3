"""The C ABI's exchange entry points (ta_exchange_*, csrc/ta_exchange.cu) on ONE GPU: a
single-rank communicator takes the same code path as N ranks — NCCL loaded with dlopen, the
library's own communicator, zero-copy sends of the owner slices (here: the self segment), the
sparse full rows through k_xchg_gather / k_xchg_scatter, owner-side ta_pr_accumulate on the
received records — and must reproduce the local accumulation bit for bit.  The N > 1 routing is
covered on CPU by tests/test_parallel_gloo.py and on 2 / 4 / 8 GPUs by bench.py's
`parity.identical` (profiles/r2_bench_n*.json)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["small", "edge_mix"])
def test_single_rank_exchange_equals_local_accumulate(case):
    import torch
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
            if plan.kind == "tao":
                eng.stage_iou(dev)
                eng.stage_match(dev)
            else:
                eng.stage_frame_eval(dev)
                assert ex.compact and dev.words_valid
            ex.accumulate()
            ex.to_root()
            torch.cuda.synchronize()
            for k in ("precision", "recall", "tp_cnt", "fp_cnt", "num_gt"):
                assert np.array_equal(dev.t[k].cpu().numpy(), getattr(ref, k)), (plan.kind, k)
            shape = g[plan.kind + "_precision"].shape
            assert np.array_equal(g[plan.kind + "_precision"],
                                  dev.t["precision"].cpu().numpy().reshape(shape))
    finally:
        tr.close()
        eng.close()
