"""Host logic (prep.py: ordering / filtering / CSR layout) + the kernels' per-thread
arithmetic (csrc/ta_device_fns.cuh built for the host, tests/hostsim) against goldens from the
unmodified reference.  CPU only; the CUDA kernels themselves are checked in test_gpu_parity.py."""
import numpy as np
import pytest

from conftest import golden_inputs
from plan_backends import compare_with_golden, plans_from_json, run_hostsim


def test_tao_plan_hostsim_matches_reference(golden):
    gt, res = golden_inputs(golden)
    tao_plan, _ = plans_from_json(gt, res)
    off_grid = golden["_name"] == "small_float"
    # the tiled kernel re-associates the union (U = DA + GA - I); exact on grid data
    out = run_hostsim(tao_plan, "3d_iou")
    compare_with_golden(golden, "tao_", tao_plan, out, exact_iou=not off_grid, iou_atol=1e-12)


def test_lvis_plan_hostsim_matches_reference(golden):
    gt, res = golden_inputs(golden)
    _, lvis_plan = plans_from_json(gt, res)
    out = run_hostsim(lvis_plan)
    compare_with_golden(golden, "lvis_", lvis_plan, out, exact_iou=True)
