"""The reference's own C box IoU (maskApi.c bbIou, built into oracle/_ref by oracle/Makefile)
pins the oracle's numpy restatement and the device function the kernels use (host build)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import common, maskapi_ref
import plan_backends

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "bbiou_ref_c.npz")


def _oracle_matrix(dt, gt):
    return np.array([[common.frame_box_iou(d, g) for g in gt] for d in dt])


def _device_fn_matrix(dt, gt):
    hs = plan_backends.build_hostsim()
    D, G = len(dt), len(gt)
    out = np.zeros(D * G)
    off = lambda *v: np.asarray(v, dtype=np.int64)
    hs.hs_box_iou.argtypes = [C.c_int64] + [C.c_void_p] * 6
    p = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
    d_off, g_off, i_off = off(0, D), off(0, G), off(0, D * G)
    dt, gt = np.ascontiguousarray(dt), np.ascontiguousarray(gt)
    hs.hs_box_iou(1, p(d_off), p(g_off), p(dt), p(gt), p(i_off), p(out))
    return out.reshape(D, G)


def test_committed_vectors_from_reference_c():
    z = np.load(GOLDEN)
    assert np.array_equal(_oracle_matrix(z["dt"], z["gt"]), z["iou"])
    assert np.array_equal(_device_fn_matrix(z["dt"], z["gt"]), z["iou"])


def test_doctest_boxes_through_reference_c():
    # tao_amodal/evaluation/tao_amodal/eval.py:21-24, 29-30 give (I, U); bbIou gives I / U
    z = np.load(GOLDEN)
    sp = z["dt"][-10:]
    iou = z["iou"][-10:, :][:, ::-1][:, -10:]          # gt is dt reversed
    assert iou[0, 0] == 1.0 and iou[0, 1] == 100.0 / 400.0 and iou[2, 3] == 25.0 / 100.0
    assert iou[0, 4] == 0.0                            # touching boxes: w <= 0
    assert np.array_equal(sp[0], [0, 0, 20, 20])


@pytest.mark.skipif(not maskapi_ref.available(), reason="oracle/_ref not built (no reference tree)")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_live_reference_c_random(seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    dt = rng.uniform(-50, 400, (96, 4))
    gt = dt[rng.integers(0, 96, 80)] + rng.normal(0, 4, (80, 4))
    dt[::7, 2] = 0.0
    gt[::9, 3] *= -1.0
    ref = maskapi_ref.iou(dt, gt, [0] * len(gt))
    assert np.array_equal(_oracle_matrix(dt, gt), ref, equal_nan=True)
    assert np.array_equal(_device_fn_matrix(dt, gt), ref, equal_nan=True)


@pytest.mark.skipif(not maskapi_ref.available(), reason="oracle/_ref not built (no reference tree)")
def test_goldens_were_generated_with_reference_c():
    from oracle import ref_shims
    if not ref_shims.reference_available():
        pytest.skip("reference tree absent")
    ref_shims.install_shims()
    import pycocotools.mask as pm
    assert pm.BACKEND.startswith("reference C") and pm.iou is maskapi_ref.iou
