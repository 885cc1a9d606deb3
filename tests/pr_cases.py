"""Random inputs of the precision/recall accumulation (ta_pr_accumulate) shared by the CPU and
GPU tests: categories spanning many 256-detection chunks, dense / sparse / absent true
positives, empty categories, categories without GT, exact chunk multiples."""
import numpy as np


def random_pr_case(seed, n_cat=7, n_cfg=6, n_thr=10, max_len=1500, tp_rate=None):
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = rng.integers(0, max_len, n_cat)
    lens[rng.integers(0, n_cat)] = 0                       # an empty category
    lens[rng.integers(0, n_cat)] = 256 * int(rng.integers(1, 4))   # exact chunk multiple
    if n_cat > 3:
        lens[3] = 1
    cat_dt_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n_dt = int(cat_dt_off[-1])
    # accumulate order: a permutation inside every category
    acc_perm = np.concatenate([cat_dt_off[c] + rng.permutation(lens[c]) for c in range(n_cat)]
                              + [np.zeros(0, dtype=np.int64)]).astype(np.int32)
    rate = rng.uniform(0.02, 0.9, (n_cat, n_cfg, 1)) if tp_rate is None else np.full((n_cat, n_cfg, 1), tp_rate)
    thr_decay = np.linspace(1.0, 0.3, n_thr)[None, None, :]
    cat_of = np.repeat(np.arange(n_cat), lens)
    u = rng.random((n_dt, n_cfg, n_thr))
    tp = u < (rate * thr_decay)[cat_of]
    ignored = rng.random((n_dt, n_cfg, 1)) < 0.2            # neither TP nor FP in this cfg
    fp = ~tp & ~ignored & (rng.random((n_dt, n_cfg, n_thr)) < 0.9)
    tp &= ~ignored
    w = np.zeros((n_dt, n_cfg), dtype=np.uint32)
    for t in range(n_thr):
        w |= tp[:, :, t].astype(np.uint32) << t
        w |= fp[:, :, t].astype(np.uint32) << (16 + t)
    tp_tot = np.stack([np.bincount(cat_of, weights=tp[:, c, 0], minlength=n_cat) for c in range(n_cfg)], 1)
    num_gt = (tp_tot + rng.integers(0, 40, (n_cat, n_cfg))).astype(np.int32)
    num_gt[rng.random((n_cat, n_cfg)) < 0.15] = 0           # cells without GT stay -1
    num_gt[(num_gt == 0) & (tp_tot > 0) & (rng.random((n_cat, n_cfg)) < 0.5)] = 1   # tp > num_gt never happens in
    num_gt = np.maximum(num_gt, np.where(num_gt > 0, tp_tot, 0)).astype(np.int32)   # the evaluators; keep tp <= num_gt
    return dict(n_cat=n_cat, n_cfg=n_cfg, cat_dt_off=cat_dt_off, acc_perm=acc_perm, tpfp=w, num_gt=num_gt)


def oracle_pr(c, iou_thrs, rec_thrs):
    """The case through the ORACLE's accumulate cell (oracle.common.pr_curve, the restatement of
    tao_amodal/evaluation/tao_amodal/eval.py:508-573 that the reference goldens pin): precision
    [T, R, C, K], recall / tp_cnt / fp_cnt [T, C, K] in the layouts of ta_pr_accumulate.  A
    detection's word says per threshold whether it is a true positive (matched, counted), a false
    positive (unmatched, counted) or ignored."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle.common import pr_curve
    T, R, C, K = len(iou_thrs), len(rec_thrs), c["n_cat"], c["n_cfg"]
    prec = -np.ones((T, R, C, K))
    rec = -np.ones((T, C, K))
    tp_cnt = np.zeros((T, C, K), dtype=np.int64)
    fp_cnt = np.zeros((T, C, K), dtype=np.int64)
    bits = np.arange(T, dtype=np.uint32)[:, None]
    for cat in range(C):
        idx = c["acc_perm"][c["cat_dt_off"][cat]:c["cat_dt_off"][cat + 1]]     # descending score
        for k in range(K):
            w = c["tpfp"][idx, k].astype(np.uint32)[None, :]
            tp = ((w >> bits) & 1).astype(bool)
            fp = ((w >> (16 + bits)) & 1).astype(bool)
            got = pr_curve(-np.arange(idx.size, dtype=np.float64), np.where(tp, 1, -1), ~(tp | fp),
                           np.zeros(int(c["num_gt"][cat, k])), -1, rec_thrs)
            if got is None:
                continue
            prec[:, :, cat, k], rec[:, cat, k] = got[0], got[1]
            tp_cnt[:, cat, k], fp_cnt[:, cat, k] = got[3].sum(1), got[4].sum(1)
    return prec, rec, tp_cnt, fp_cnt


def structured_pr_case(patterns, n_cfg=2, n_thr=10, seed=5):
    """A case from explicit per-category position patterns (0 ignored, 1 true positive, 2 false
    positive, in descending-score order); higher thresholds lose some true positives, the second
    cfg is the pattern rotated by three positions."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = [len(p) for p in patterns]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    w = np.zeros((int(off[-1]), n_cfg), dtype=np.uint32)
    for c, p in enumerate(patterns):
        p = np.asarray(p)
        for k in range(n_cfg):
            for t in range(n_thr):
                q = p.copy()
                if t >= 5:
                    q = np.where((np.arange(len(p)) % (t + 2) == 0) & (q == 1), 2, q)
                if k == 1:
                    q = np.roll(q, 3)
                w[off[c]:off[c + 1], k] |= ((q == 1).astype(np.uint32) << t) | ((q == 2).astype(np.uint32) << (16 + t))
    tp_tot = np.array([[max(int(((w[off[c]:off[c + 1], k] >> t) & 1).sum()) for t in range(n_thr))
                        for k in range(n_cfg)] for c in range(len(patterns))], dtype=np.int64)
    num_gt = (tp_tot + rng.integers(0, 5, tp_tot.shape)).astype(np.int32)
    return dict(n_cat=len(patterns), n_cfg=n_cfg, cat_dt_off=off,
                acc_perm=np.arange(int(off[-1]), dtype=np.int32), tpfp=w, num_gt=num_gt)


def structured_patterns(L, rng):
    i = np.arange(L)
    return [np.ones(L, int), np.full(L, 2), i % 2 + 1, np.where(i < L // 2, 1, 2), np.where(i < L // 2, 2, 1),
            np.where(i % 32 == 31, 1, np.where(i % 32 == 0, 2, 0)),     # TP ends a word, FP starts the next
            np.where(i % 256 == 255, 1, 0),                              # a lone TP at every chunk end
            rng.integers(0, 3, L), np.where(rng.random(L) < 0.05, 2, np.where(rng.random(L) < 0.5, 1, 0))]
