"""Native JSON ingest (libta_ingest.so) against the dict-based column builders: identical
columns on every golden input, same exceptions on malformed / incomplete files.  CPU only."""
import json
from dataclasses import fields

import numpy as np
import pytest

from conftest import golden_inputs
from tao_amodal_b200 import ingest
from tao_amodal_b200.columnar import DtColumns, GtColumns


def _same(a, b):
    if isinstance(a, tuple):
        return all(_same(x, y) for x, y in zip(a, b))
    if isinstance(a, np.ndarray):
        return (a.dtype == b.dtype and a.shape == b.shape
                and np.array_equal(a, b, equal_nan=a.dtype.kind == "f"))
    return a == b


def test_columns_identical_on_golden_inputs(golden, tmp_path):
    gt, res = golden_inputs(golden)
    ap, rp = str(tmp_path / "gt.json"), str(tmp_path / "dt.json")
    json.dump(gt, open(ap, "w"))
    json.dump(res, open(rp, "w"), indent=1)          # whitespace variations
    g_n, d_n = ingest.load_gt(ap, need_videos_tracks=True), ingest.load_dt(rp)
    g_p, d_p = GtColumns.from_dict(gt), DtColumns.from_list(res)
    for f in fields(GtColumns):
        assert _same(getattr(g_n, f.name), getattr(g_p, f.name)), f.name
    for f in fields(DtColumns):
        assert _same(getattr(d_n, f.name), getattr(d_p, f.name)), f.name


def test_odd_but_valid_json(tmp_path):
    gt = {"info": {"nested": {"a": [1, {"b": 'x"y'}]}}, "licenses": [],
          "images": [{"id": 7, "file_name": 'a\\b".jpg', "video_id": 3, "frame_index": 2,
                      "neg_category_ids": [], "not_exhaustive_category_ids": [5, 6]}],
          "videos": [{"id": 3, "neg_category_ids": [9], "not_exhaustive_category_ids": []}],
          "tracks": [{"id": 1, "category_id": 5, "video_id": 3, "ignore": True}],
          "categories": [{"id": 5, "frequency": "r", "merged": [{"id": 50, "name": "q"}, {"id": 51}]},
                         {"id": 6, "name": "no frequency"}],
          "annotations": [{"id": 1, "image_id": 7, "track_id": 1, "category_id": 5,
                           "bbox": [1, 2.5, 3e1, 4], "area": 120, "visibility": 0.25,
                           "out_of_frame": False, "ignore": 0, "extra": None},
                          {"id": 2, "image_id": 7, "category_id": 5, "bbox": [0, 0, 1, 1],
                           "area": 1.0, "out_of_frame": 1}]}
    p = str(tmp_path / "gt.json")
    json.dump(gt, open(p, "w"))
    a, b = ingest.load_gt(p, need_videos_tracks=True), GtColumns.from_dict(gt)
    for f in fields(GtColumns):
        assert _same(getattr(a, f.name), getattr(b, f.name)), f.name
    assert a.merge_map == {50: 5, 51: 5} and a.ann_track_id.tolist() == [1, -1]
    assert np.isnan(a.ann_visibility[1]) and a.ann_oof.tolist() == [0, 1]
    res = [{"image_id": 7, "category_id": 5, "bbox": [1, 2, 3, 4], "score": 1, "track_id": 4,
            "video_id": 3, "segmentation": {"counts": "abc", "size": [1, 2]}},
           {"score": 2.5e-1, "bbox": [0.1, 0.2, 0.3, 0.4], "category_id": 6, "image_id": 7}]
    p = str(tmp_path / "dt.json")
    json.dump(res, open(p, "w"))
    a, b = ingest.load_dt(p), DtColumns.from_list(res)
    for f in fields(DtColumns):
        assert _same(getattr(a, f.name), getattr(b, f.name)), f.name


def test_errors_mirror_the_dict_path(tmp_path):
    p = str(tmp_path / "x.json")
    open(p, "w").write('[{"image_id": 1, "category_id": 2, "bbox": [1,2,3,4]}]')
    with pytest.raises(KeyError, match="score"):
        ingest.load_dt(p)
    open(p, "w").write('{"a": 1}')
    with pytest.raises(AssertionError, match="not a list"):
        ingest.load_dt(p)
    open(p, "w").write('[{"image_id": 1, "category_id": 2, "bbox": [1,2,3,4], "score": 0.5}')
    with pytest.raises(json.JSONDecodeError):
        ingest.load_dt(p)
    open(p, "w").write('{"images": [], "annotations": [], "categories": []}')
    ingest.load_gt(p)                                   # enough for the frame evaluator
    with pytest.raises(KeyError, match="videos"):
        ingest.load_gt(p, need_videos_tracks=True)      # Tao needs videos and tracks
    with pytest.raises(FileNotFoundError):
        ingest.load_gt(str(tmp_path / "missing.json"))


def test_prefetch_shares_one_parse_with_the_foreground_load(tmp_path):
    """ingest.prefetch + load_dt: the foreground call waits for the background parse and gets
    the very same columns object; a broken file raises in the foreground as before."""
    import json
    from tao_amodal_b200 import ingest, synth
    gt, dt = synth.generate_named("tiny")
    p = str(tmp_path / "dt.json")
    json.dump(dt.to_list(), open(p, "w"))
    ingest._CACHE.clear()
    t = ingest.prefetch(p, "dt")
    a = ingest.load_dt(p)
    t.join()
    assert ingest.load_dt(p) is a and len(ingest._INFLIGHT) == 0
    assert np.array_equal(a.bbox, dt.bbox)
    bad = str(tmp_path / "bad.json")
    open(bad, "w").write('[{"image_id": 1, "bbox": [0, 0, 1')
    ingest.prefetch(bad, "dt").join()
    with pytest.raises(Exception):
        ingest.load_dt(bad)


def _dt_equal(a, b):
    for f in ("image_id", "track_id", "category_id", "video_id", "bbox", "score"):
        assert np.array_equal(getattr(a, f), getattr(b, f), equal_nan=True), f


@pytest.mark.parametrize("threads", [2, 3, 8])
def test_parallel_result_parse_equals_sequential(tmp_path, monkeypatch, threads):
    """Large result lists are parsed by several workers from GUESSED object starts that are then
    verified by chaining (csrc/ta_json.cpp); forced here on small files, including values that
    contain the very byte pattern the guess looks for."""
    import json
    from tao_amodal_b200 import ingest, synth
    from tao_amodal_b200.columnar import DtColumns
    gt, dt = synth.generate_named("small")
    clean = dt.to_list()
    nasty = dt.to_list()
    for k, r in enumerate(nasty):
        if k % 5 == 0:      # nested values and strings with '}, {' inside: wrong guesses
            r["segmentation"] = {"size": [4, 4], "counts": "ab}, {\"image_id\": 7}, {cd"}
        if k % 7 == 0:
            r["extra"] = [{"a": [1, 2, {"b": "}, {"}]}, "}, {"]
    for name, res, indent in (("clean", clean, None), ("clean_indent", clean, 1),
                              ("nasty", nasty, None), ("nasty_indent", nasty, 1)):
        p = str(tmp_path / ("dt_%s.json" % name))
        json.dump(res, open(p, "w"), indent=indent)
        monkeypatch.setenv("TA_INGEST_THREADS", "1")
        ingest._CACHE.clear()
        seq = ingest.load_dt(p)
        monkeypatch.setenv("TA_INGEST_THREADS", str(threads))
        monkeypatch.setenv("TA_INGEST_PAR_MIN_BYTES", "0")
        ingest._CACHE.clear()
        par = ingest.load_dt(p)
        if name.startswith("clean"):      # guesses hold: really parsed by `threads` workers
            assert ingest.load_lib().ta_json_last_parse_workers() == threads
        _dt_equal(seq, par)
        _dt_equal(par, DtColumns.from_list(json.load(open(p))))
    ingest._CACHE.clear()


def test_parallel_result_parse_reports_the_sequential_error(tmp_path, monkeypatch):
    import json
    from tao_amodal_b200 import ingest, synth
    gt, dt = synth.generate_named("tiny")
    res = dt.to_list()
    del res[len(res) // 2]["score"]
    p = str(tmp_path / "bad.json")
    json.dump(res, open(p, "w"))
    monkeypatch.setenv("TA_INGEST_PAR_MIN_BYTES", "0")
    monkeypatch.setenv("TA_INGEST_THREADS", "4")
    ingest._CACHE.clear()
    with pytest.raises(KeyError, match="score"):
        ingest.load_dt(p)
    open(p, "w").write(json.dumps(dt.to_list()) + " trailing")
    with pytest.raises(json.JSONDecodeError):
        ingest.load_dt(p)
    open(p, "w").write("[]")
    assert ingest.load_dt(p).n() == 0
    ingest._CACHE.clear()


def test_results_without_track_or_video_id_raise_like_the_reference(tmp_path):
    """a results file that lacks track_id (but has video_id) must raise
    KeyError('track_id') on the track path (tools/eval_on_tao_amodal.py:57) instead of scoring
    every video as one giant track; the frame path never reads those keys."""
    import json
    import pytest
    from tao_amodal_b200 import ingest, prep, synth
    from tao_amodal_b200.columnar import DtColumns
    gt, dt = synth.generate_named("tiny")
    res = dt.to_list()
    for drop, want in (("track_id", "track_id"), ("video_id", "video_id")):
        lst = [{k: v for k, v in r.items() if k != drop} for r in res]
        p = tmp_path / ("no_%s.json" % drop)
        json.dump(lst, open(p, "w"))
        for cols in (ingest.load_dt(str(p)), DtColumns.from_list(lst)):
            assert (cols.missing_track_id > 0) == (drop == "track_id")
            assert (cols.missing_video_id > 0) == (drop == "video_id")
            prep.prepare_lvis(gt, cols)                          # frame path: fine
            with pytest.raises(KeyError, match=want):
                prep.make_track_ids_unique(cols.copy())
            with pytest.raises(KeyError, match=want):
                prep.prepare_tao(gt, cols)
    p = tmp_path / "full.json"
    json.dump(res, open(p, "w"))
    full = ingest.load_dt(str(p))
    assert full.missing_track_id == 0 and full.missing_video_id == 0
