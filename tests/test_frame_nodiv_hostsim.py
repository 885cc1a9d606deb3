"""Flat frame kernel, NODIV variant (one division per detection): its candidate screening
(ta_frame_candidates_nodiv, host build) never loses a candidate of the exact per-pair quotient
and agrees with it whenever it reports a single one."""
import ctypes as C

import numpy as np
import pytest

import plan_backends


def _run(det, gt, thr_min):
    hs = plan_backends.build_hostsim()
    n, G = len(det), len(gt)
    out = [np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32), np.zeros(n),
           np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32), np.zeros(n)]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    det, gt = np.ascontiguousarray(det, dtype=np.float64), np.ascontiguousarray(gt, dtype=np.float64)
    hs.hs_frame_candidates.argtypes = [C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_double] + [C.c_void_p] * 6
    hs.hs_frame_candidates(n, p(det), G, p(gt), thr_min, *[p(o) for o in out])
    return out


def _check(det, gt, thr_min):
    cn, gn, vn, ce, ge, ve = _run(det, gt, thr_min)
    assert (cn >= ce).all()                                   # no candidate lost
    one = cn == 1
    ok = one & ~(vn < thr_min)                                # survivor passes the exact re-check
    assert (ce[ok] == 1).all() and (gn[ok] == ge[ok]).all()
    assert np.array_equal(vn[ok], ve[ok])                     # and carries the exact quotient
    assert (ce[one & (vn < thr_min)] == 0).all()
    assert (ce[cn == 0] == 0).all()
    return int(((cn >= 2) & (ce <= 1)).sum())                 # needless trips to the general matcher


@pytest.mark.parametrize("seed", range(6))
def test_random_boxes(seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    gt = np.concatenate([rng.uniform(0, 200, (6, 2)), rng.uniform(5, 120, (6, 2))], 1)
    det = gt[rng.integers(0, 6, 4000)] + rng.normal(0, rng.choice([0.5, 4, 15]), (4000, 4))
    det[::50, 2] = 0.0
    det[::77, 3] *= -1
    det[::91] = np.nan
    assert _check(det, gt, 0.5) == 0
    assert _check(det, gt, 0.05) == 0


def test_quotients_at_and_next_to_the_threshold():
    # GT 10 x 20 at the origin; detections 10 x h inside it have IoU h / 20: exactly 0.5 at h = 10
    gt = np.array([[0.0, 0.0, 10.0, 20.0]])
    hs = 10.0 + np.arange(-40, 41) * np.spacing(10.0)
    det = np.stack([np.zeros_like(hs), np.zeros_like(hs), np.full_like(hs, 10.0), hs], 1)
    _check(det, gt, 0.5)
    cn, gn, vn, ce, ge, ve = _run(det, gt, 0.5)
    assert ce.min() == 0 and ce.max() == 1                    # the sweep straddles the threshold
    assert np.array_equal((cn == 1) & ~(vn < 0.5), ce == 1)


def test_non_positive_threshold_makes_every_pair_a_candidate():
    rng = np.random.Generator(np.random.PCG64(9))
    gt = np.concatenate([rng.uniform(0, 200, (3, 2)), rng.uniform(5, 60, (3, 2))], 1)
    det = np.concatenate([rng.uniform(0, 200, (500, 2)), rng.uniform(5, 60, (500, 2))], 1)
    for thr in (0.0, -1.0):
        cn, gn, vn, ce, ge, ve = _run(det, gt, thr)
        assert (cn == 3).all() and (ce == 3).all() and (gn == ge).all()
        cn, gn, vn, ce, ge, ve = _run(det, gt[:1], thr)
        assert (cn == 1).all() and np.array_equal(vn, ve)
