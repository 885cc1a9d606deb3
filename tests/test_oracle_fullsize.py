"""The oracle (CPU restatement) against the BASELINE-size fixture of the UNMODIFIED reference:
configs[1] in full (cfg2, `tests/golden/full_cfg2.npz`, oracle/make_fullsize_golden.py)."""
import copy
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def test_oracle_reproduces_the_reference_at_cfg2():
    from oracle import lvis_frame, tao_track
    from tao_amodal_b200 import synth
    fx = np.load(os.path.join(GOLDEN, "full_cfg2.npz"))
    gt, dt = synth.generate_named("cfg2")
    assert int(fx["n_pred_boxes"]) == dt.n() and int(fx["n_gt_boxes"]) == gt.n_anns()
    gd, res = gt.to_dict(), dt.to_list()
    res2 = copy.deepcopy(res)
    tao_track.uniquify_track_ids(res2)
    for prefix, out in (("tao", tao_track.evaluate_tao(copy.deepcopy(gd), res2, keep_cells=False)),
                        ("lvis", lvis_frame.evaluate_lvis(gd, res, keep_cells=False))):
        shape = tuple(fx[prefix + "_recall"].shape)
        assert np.array_equal(fx[prefix + "_tp_cnt"], out["tp_cnt"].reshape(shape))
        assert np.array_equal(fx[prefix + "_fp_cnt"], out["fp_cnt"].reshape(shape))
        assert np.array_equal(fx[prefix + "_recall"], out["recall"].reshape(shape))
        assert str(fx[prefix + "_precision_sha256"]) == _sha(out["precision"])
        got = np.asarray([float(v) for v in out["results"].values()])
        assert np.array_equal(fx[prefix + "_results"], got)
