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
    r = run_sequence(_config(g), segment_fn=make_segment_provider(g.meta["segments"]), tracker_provider="truth",
                     max_frames=n)
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
