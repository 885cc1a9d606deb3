"""Bit-plane PR kernels (k_pr_bits / k_pr_envelope_bits, csrc/ta_pr.cu): their per-thread
functions, emulated serially on the host, against the plain serial accumulation that the
goldens of the unmodified reference pin."""
import numpy as np
import pytest

import plan_backends
from plan_backends import hostsim_pr
from pr_cases import random_pr_case
from tao_amodal_b200 import engine


def _same(a, b):
    for k in ("precision", "recall", "tp_cnt", "fp_cnt"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


@pytest.mark.parametrize("impl", ["bits", "bits_tile", "bits_seg"])
@pytest.mark.parametrize("seed", range(8))
def test_random_multi_chunk(seed, impl):
    c = random_pr_case(seed)
    args = (c["n_cat"], c["cat_dt_off"], c["acc_perm"], c["tpfp"], c["num_gt"], c["n_cfg"])
    _same(hostsim_pr(*args, impl=impl), hostsim_pr(*args, impl="serial"))


@pytest.mark.parametrize("seed,kw", [(0, {}), (1, {}), (2, dict(n_cat=5, n_cfg=20, max_len=700)),
                                     (100, dict(n_cat=4, n_cfg=3, tp_rate=0.0)),
                                     (101, dict(n_cat=4, n_cfg=3, tp_rate=1.0)),
                                     (7, dict(n_cat=3, n_cfg=2, tp_rate=0.97, max_len=2600))])
def test_emulations_match_the_oracle_accumulate_cell(seed, kw):
    """The random multi-chunk cases against the ORACLE's accumulate cell (not only against the
    serial emulation built from the same device header): serial and bit-plane emulations."""
    from pr_cases import oracle_pr
    c = random_pr_case(seed, **kw)
    args = (c["n_cat"], c["cat_dt_off"], c["acc_perm"], c["tpfp"], c["num_gt"], c["n_cfg"])
    prec, rec, tp, fp = oracle_pr(c, engine.IOU_THRS, engine.REC_THRS)
    for impl in ("serial", "bits_tile", "bits_seg"):
        out = hostsim_pr(*args, impl=impl)
        assert np.array_equal(out.precision, prec), impl
        assert np.array_equal(out.recall, rec), impl
        assert np.array_equal(out.tp_cnt, tp) and np.array_equal(out.fp_cnt, fp), impl


@pytest.mark.parametrize("L", [1, 31, 32, 33, 255, 256, 257, 511, 512, 513, 1024, 1300])
def test_structured_runs_against_the_oracle(L):
    """Runs of true positives that end at word and chunk boundaries, lone true positives at chunk
    ends, all-TP / all-FP / alternating lists, lengths around the word and chunk sizes: every
    emulation of the run-end walk against the oracle's accumulate cell."""
    from pr_cases import oracle_pr, structured_patterns, structured_pr_case
    c = structured_pr_case(structured_patterns(L, np.random.Generator(np.random.PCG64(L))))
    args = (c["n_cat"], c["cat_dt_off"], c["acc_perm"], c["tpfp"], c["num_gt"], c["n_cfg"])
    prec, rec, tp, fp = oracle_pr(c, engine.IOU_THRS, engine.REC_THRS)
    for impl in ("serial", "bits", "bits_tile", "bits_seg"):
        out = hostsim_pr(*args, impl=impl)
        assert np.array_equal(out.precision, prec) and np.array_equal(out.recall, rec), impl
        assert np.array_equal(out.tp_cnt, tp) and np.array_equal(out.fp_cnt, fp), impl


@pytest.mark.parametrize("tp_rate", [0.0, 1.0])
def test_all_or_no_true_positives(tp_rate):
    c = random_pr_case(100, n_cat=4, n_cfg=3, tp_rate=tp_rate)
    args = (c["n_cat"], c["cat_dt_off"], c["acc_perm"], c["tpfp"], c["num_gt"], c["n_cfg"])
    _same(hostsim_pr(*args, impl="bits"), hostsim_pr(*args, impl="serial"))


@pytest.mark.parametrize("tp_rate", [0.02, 0.3, 0.9, 0.995])
@pytest.mark.parametrize("seed", range(4))
def test_true_positive_runs_of_every_length(seed, tp_rate):
    """The walk only visits the true positives that end a run: long categories (many chunks, so
    that the state is carried over several of them), from almost no to almost only true positives,
    with ignored detections (neither TP nor FP) in between."""
    c = random_pr_case(200 + seed, n_cat=3, n_cfg=3, tp_rate=tp_rate, max_len=2600)
    args = (c["n_cat"], c["cat_dt_off"], c["acc_perm"], c["tpfp"], c["num_gt"], c["n_cfg"])
    want = hostsim_pr(*args, impl="serial")
    for impl in ("bits", "bits_tile", "bits_seg"):
        _same(hostsim_pr(*args, impl=impl), want)


def test_track_shape_and_odd_thresholds():
    c = random_pr_case(5, n_cat=5, n_cfg=20, n_thr=10, max_len=700)
    args = (c["n_cat"], c["cat_dt_off"], c["acc_perm"], c["tpfp"], c["num_gt"], c["n_cfg"])
    _same(hostsim_pr(*args, impl="bits"), hostsim_pr(*args, impl="serial"))
    c = random_pr_case(6, n_cat=5, n_cfg=2, n_thr=3, max_len=900)
    args = (c["n_cat"], c["cat_dt_off"], c["acc_perm"], c["tpfp"], c["num_gt"], c["n_cfg"])
    thr, rec = engine.IOU_THRS[:3], np.array([0.0, 0.25, 0.5, 0.5, 0.99, 1.0])
    for impl in ("bits", "bits_tile", "bits_seg"):
        _same(hostsim_pr(*args, iou_thrs=thr, rec_thrs=rec, impl=impl),
              hostsim_pr(*args, iou_thrs=thr, rec_thrs=rec, impl="serial"))


def test_transpose_is_a_warp_of_ballots():
    import ctypes as C
    hs = plan_backends.build_hostsim()
    rng = np.random.Generator(np.random.PCG64(3))
    x = rng.integers(0, 2**32, 32, dtype=np.uint64).astype(np.uint32)
    out = np.zeros(32, dtype=np.uint32)
    hs.hs_transpose32(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    for b in range(32):
        ballot = sum(((int(x[l]) >> b) & 1) << l for l in range(32))
        assert int(out[b]) == ballot


def test_goldens_through_the_bit_plane_emulation(golden):
    """Every golden of the unmodified reference, PR step done by the bit-plane emulation."""
    from conftest import golden_inputs
    from plan_backends import compare_with_golden, plans_from_json, run_hostsim
    gt, res = golden_inputs(golden)
    tao_plan, lvis_plan = plans_from_json(gt, res)
    off_grid = golden["_name"] == "small_float"
    for impl in ("bits", "bits_tile", "bits_seg"):
        compare_with_golden(golden, "tao_", tao_plan, run_hostsim(tao_plan, pr_impl=impl),
                            exact_iou=not off_grid, iou_atol=1e-12)
        compare_with_golden(golden, "lvis_", lvis_plan, run_hostsim(lvis_plan, pr_impl=impl))


@pytest.mark.parametrize("n_thr,n_cfg", [(10, 6), (16, 5), (1, 10), (3, 2), (13, 6)])
def test_compact_words_give_the_planes_of_the_expanded_rows(n_thr, n_cfg):
    """k_pr_bits, compact-word path: one bit transpose of the words + pr_cell_planes per cell
    equals expanding every word into its full TP/FP row (pr_expand) and transposing per cfg;
    and pr_expand equals the host-side expansion (engine.expand_words)."""
    import ctypes as C
    hs = plan_backends.build_hostsim()
    rng = np.random.Generator(np.random.PCG64(n_thr * 100 + n_cfg))
    assert n_thr + 3 * n_cfg <= 31
    for _ in range(50):
        w = rng.integers(0, 1 << (n_thr + 3 * n_cfg), size=32, dtype=np.uint64).astype(np.uint32)
        w[rng.uniform(size=32) < 0.1] |= np.uint32(1 << 31)      # detections with a full row
        w[rng.uniform(size=32) < 0.1] = 0                       # dead lanes
        tp = np.zeros(n_thr * n_cfg, dtype=np.uint32)
        fp = np.zeros_like(tp)
        bad = hs.hs_pr_compact_planes(w.ctypes.data_as(C.c_void_p), n_thr, n_cfg,
                                      tp.ctypes.data_as(C.c_void_p), fp.ctypes.data_as(C.c_void_p))
        assert bad == 0
        rows = engine.expand_words(w, np.zeros((32, n_cfg), dtype=np.uint32), n_thr, n_cfg)
        live = (w >> 31) == 0
        for cfg in range(n_cfg):
            for k in range(n_thr):
                want_tp = sum(int((rows[l, cfg] >> k) & 1) << l for l in range(32) if live[l])
                want_fp = sum(int((rows[l, cfg] >> (16 + k)) & 1) << l for l in range(32) if live[l])
                assert int(tp[cfg * n_thr + k]) == want_tp and int(fp[cfg * n_thr + k]) == want_fp
