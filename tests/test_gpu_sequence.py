"""The sequence driver (main.py flow) and the sweep (benchmark.py flow) on the GPU, against the
per-frame IoUs the UNMODIFIED reference produced for the same clips, boxes and label maps
(tests/golden/seq_*.npz: training with scikit-learn, update() per frame, computeBenchmark)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import yaml

from helpers import PKG, SEQ_NAMES, GoldenSeq, polygons

pytestmark = pytest.mark.gpu


def _config(g):
    with open(os.path.join(PKG, "config_benchmark.yaml")) as f:
        base = yaml.full_load(f)
    v = g.meta["video"]
    return {**base, "input_video": "Input/SegTrack2/Video/%s.mp4" % v, "input_truth": "Input/SegTrack2/Truth/%s.mp4" % v,
            "multi_selection": g.meta["multi_selection"], "params": dict(g.params), **polygons()[v]}


@pytest.mark.parametrize("name", SEQ_NAMES)
def test_run_sequence_reproduces_reference_iou(name):
    from pcm.providers import make_segment_provider
    from pcm.sequence import run_sequence
    g = GoldenSeq(name)
    if not g.frames_match():
        pytest.skip("video decoder output differs from the one the goldens were made with")
    n = g.meta["n_frames"]
    prior_fn = None
    if g.has_recorded_priors():            # the reference's FLANN matcher is randomised: replay the golden run's priors
        calls = iter(range(1, n))
        prior_fn = lambda *a, **k: g.recorded_priors(next(calls))
    r = run_sequence(_config(g), segment_fn=make_segment_provider(g.meta["segments"]), tracker_provider="truth",
                     max_frames=n, prior_fn=prior_fn)
    assert r["n_frames"] == n and r["n_updates"] == n and r["tracker"] == "truth"
    want = g.z["iou"]
    assert len(r["iou"]) == n
    assert np.array_equal(np.asarray(r["iou"]), want, equal_nan=True), "per-frame IoU differs from the reference"
    assert r["mean_iou"] == float(np.mean(want))


def test_main_cli_writes_the_reference_result_file(tmp_path):
    g = GoldenSeq("worm_rgb3")
    cfg = _config(g)
    cfg["tracker_provider"] = "truth"
    cfg["params"]["over_segmentation"] = g.meta["segments"]
    cfg_path, out_path = tmp_path / "config-0-worm.yaml", tmp_path / "results-0-worm.csv"
    with open(cfg_path, "w") as f:
        yaml.dump(cfg, f, sort_keys=False)
    res = subprocess.run([sys.executable, os.path.join(PKG, "main.py"), str(cfg_path), str(out_path)],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    iou, secs = map(float, open(out_path).readline().split(";"))        # benchmark.py:20-21
    assert 0.0 < iou <= 1.0 and secs > 0.0
    if g.frames_match():
        # the first frames are the golden ones; the whole clip keeps scoring in the same band
        assert abs(iou - float(np.mean(g.z["iou"]))) < 0.25


def test_sweep_single_rank_smoke(tmp_path):
    """Four sequences of the reference grid (one combo x four videos), truncated clips: shared
    forests are fitted once, results come back in the reference's CSV layout."""
    from pcm import sweep
    with open(os.path.join(PKG, "config_benchmark.yaml")) as f:
        base = yaml.full_load(f)
    base["tracker_provider"] = "truth"
    out = tmp_path / "benchmark_results.csv"
    summary, table = sweep.run(base, polygons(), limit=4, max_frames=12, out_csv=str(out))
    assert summary["n_sequences"] == 4 and summary["sequences_per_s"] > 0
    for v in sweep.VIDEOS:
        assert 0.0 <= table.loc[0, v + "_benchmark"] <= 1.0 and table.loc[0, v + "_time"] > 0
    assert "avg_benchmark" in table.columns and os.path.isfile(out)


@pytest.mark.parametrize("video,over_seg,novelty,features", [("soldier", "quickshift", False, "8 hsv_lab"),
                                                              ("bmx", "felzenszwalb", True, "6 lab"),
                                                              ("worm", "quickshift", True, "6 lab")])
def test_clip_resident_sequence_equals_the_per_frame_path(video, over_seg, novelty, features):
    """pcm.fastseq (device-resident clip, precomputed boxes and label maps, asynchronous frames -- what the sweep
    runs) gives the per-frame IoUs of pcm.sequence.run_sequence (the main.py flow through Masker.update)."""
    from pcm import fastseq, sweep
    from pcm.sequence import run_sequence
    with open(os.path.join(PKG, "config_benchmark.yaml")) as f:
        base = yaml.full_load(f)
    base["tracker_provider"] = "truth"
    params = dict(n_estimators=20, max_depth=7, n_components=1, novelty_detection=novelty, over_segmentation=over_seg,
                  features=features, dilation_kernel=7, prior_weight=0.0)
    cfg = sweep.sequence_config(base, polygons(), video, params, "Input/SegTrack2/Video", "Input/SegTrack2/Truth")
    n = 40
    want = run_sequence(cfg, max_frames=n)
    clip = fastseq.ClipContext(cfg["input_video"], cfg["input_truth"], 1, 0, max_frames=n)
    got = fastseq.run_sequence_fast(cfg, clip)
    again = fastseq.run_sequence_fast(cfg, clip, model_cache=sweep.ModelCache(), cache_tag=video)
    clip.close()
    assert got["n_frames"] == want["n_frames"] and got["tracker"] == want["tracker"] == "truth"
    assert np.array_equal(np.asarray(got["iou"]), np.asarray(want["iou"]), equal_nan=True)
    assert np.array_equal(np.asarray(again["iou"]), np.asarray(want["iou"]), equal_nan=True)
    assert got["mean_iou"] == want["mean_iou"]


def test_clip_resident_sequence_with_sift_prior():
    """prior_weight = 0.1: the FLANN kd-tree matcher of the reference (:139-141) is randomised, so two runs of the SAME
    path differ in a few matches; the two drivers must agree to within that noise."""
    from pcm import fastseq, sweep
    from pcm.sequence import run_sequence
    with open(os.path.join(PKG, "config_benchmark.yaml")) as f:
        base = yaml.full_load(f)
    base["tracker_provider"] = "truth"
    params = dict(n_estimators=20, max_depth=7, n_components=1, novelty_detection=False, over_segmentation="quickshift",
                  features="6 lab", dilation_kernel=7, prior_weight=0.1)
    cfg = sweep.sequence_config(base, polygons(), "soldier", params, "Input/SegTrack2/Video", "Input/SegTrack2/Truth")
    want = run_sequence(cfg)
    clip = fastseq.ClipContext(cfg["input_video"], cfg["input_truth"], 1, 0)
    got = fastseq.run_sequence_fast(cfg, clip)
    clip.close()
    assert len(got["iou"]) == len(want["iou"])
    assert abs(got["mean_iou"] - want["mean_iou"]) < 0.02
    assert got["iou"][0] == want["iou"][0]              # frame 0 has no prior
