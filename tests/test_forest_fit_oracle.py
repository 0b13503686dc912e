"""oracle/forest_fit_oracle.c (the CPU restatement of scikit-learn's tree builder that the GPU trainer is
checked against) pinned against scikit-learn itself: every tree_ array must be EQUAL, node for node."""
import numpy as np
import pytest

import forest_fit_oracle as ffo


def _sk_fit(Xi, y, T, D):
    from sklearn.ensemble import RandomForestClassifier
    return RandomForestClassifier(random_state=42, n_estimators=T, max_depth=D).fit(Xi.astype(np.float64) / 255, y)


def assert_trees_equal_sklearn(clf, trees, full=True):
    assert len(clf.estimators_) == len(trees)
    for i, (est, t) in enumerate(zip(clf.estimators_, trees)):
        tr = est.tree_
        if isinstance(t, tuple):
            t = dict(feature=t[0], threshold=t[1], left=t[2], right=t[3], value1=t[4])
        assert tr.node_count == len(t["feature"]), "tree %d: %d vs %d nodes" % (i, tr.node_count, len(t["feature"]))
        assert np.array_equal(tr.feature, t["feature"]), i
        assert np.array_equal(tr.threshold, t["threshold"]), i
        assert np.array_equal(tr.children_left, t["left"]), i
        assert np.array_equal(tr.children_right, t["right"]), i
        assert np.array_equal(tr.value[:, 0, 1], t["value1"]), i
        if full and "impurity" in t:
            assert np.array_equal(tr.impurity, t["impurity"]), i
            assert np.array_equal(tr.n_node_samples, t["n_node_samples"]), i
            assert np.array_equal(tr.weighted_n_node_samples, t["weighted_n"]), i


def synthetic_rows(seed, n, F, constants=False):
    rng = np.random.default_rng(seed)
    Xi = rng.integers(-1, 256, (n, F)).astype(np.int16)
    y = ((Xi[:, 3 % F].astype(int) + Xi[:, 50 % F] + rng.integers(0, 80, n)) > 290).astype(np.int64)
    if constants:                      # constant columns, two-valued columns, the -1 border sentinel
        Xi[:, ::3] = 7
        Xi[:, 1::7] = (Xi[:, 1::7] > 128) * 255 - (Xi[:, 1::7] < 20)
    y[0], y[1] = 0, 1
    return Xi, y


@pytest.mark.parametrize("n,F,T,D,constants", [(3000, 147, 6, 4, False), (2500, 147, 8, 10, False), (2500, 390, 6, 7, True),
                                                (50, 27, 12, 30, True), (2, 5, 3, 5, False), (400, 1, 5, 6, False)])
def test_oracle_grows_scikit_learns_trees(n, F, T, D, constants):
    Xi, y = synthetic_rows(n + F, n, F, constants)
    assert_trees_equal_sklearn(_sk_fit(Xi, y, T, D), ffo.fit_forest(Xi, y, T, D))


def test_oracle_on_clip_training_rows():
    """Training rows of a real selection (worm, '6 lab': polygon bbox + RONI as addModel builds them, :170-197)."""
    import cv2 as cv
    import pcm_oracle as orc
    from helpers import polygons, read_video
    P = polygons()["worm"]
    f = read_video("Video", "worm")[P["pts_frame_numbers"][0]]
    pts, roni = P["pts"][0][0], P["bboxes_roni"][0][0]
    x, y, w, h = cv.boundingRect(np.array(pts))
    X = orc.get_features_int(orc.build_planes(f[y:y + h, x:x + w], ["lab"]), 6)
    Xn = orc.get_features_int(orc.build_planes(f[roni[1]:roni[1] + roni[3], roni[0]:roni[0] + roni[2]], ["lab"]), 6)
    roi = np.zeros((h, w), np.uint8)
    cv.fillPoly(roi, np.array([[(p[0] - x, p[1] - y) for p in pts]], dtype=np.int32), 255)
    Xi = np.concatenate([X, Xn]).astype(np.int16)
    lab = np.concatenate([(roi.reshape(-1) > 0).astype(np.int64), np.zeros(len(Xn), np.int64)])
    assert_trees_equal_sklearn(_sk_fit(Xi, lab, 5, 10), ffo.fit_forest(Xi, lab, 5, 10))


def test_product_draws_equal_the_oracles_and_scikit_learns():
    """pcm/train.py draws seeds and bootstrap counts like scikit-learn (no GPU needed for this part)."""
    from pcm import train
    n, T = 777, 9
    counts, seeds = train.tree_draws(n, T)
    for t, (c, s) in enumerate(ffo.tree_draws(n, T)):
        assert np.array_equal(counts[t], c) and int(seeds[t]) == s
    Xi, y = synthetic_rows(5, n, 20)
    clf = _sk_fit(Xi, y, T, 3)
    from sklearn.ensemble._forest import _generate_sample_indices
    for t, est in enumerate(clf.estimators_):
        idx = _generate_sample_indices(est.random_state, n, n, None)
        assert np.array_equal(np.bincount(idx, minlength=n), counts[t])


def test_draws_are_computed_once_when_many_threads_ask():
    """The fits of one row set start together in the sweep: they must share ONE computation of the bootstrap draws
    (30 ms under the interpreter lock each, otherwise) and get the very same arrays."""
    import threading
    from pcm import train
    n, T = 4321, 12
    got, errors = [], []

    def ask():
        try:
            got.append(train.tree_draws(n, T, 1234))
        except Exception as e:               # pragma: no cover
            errors.append(e)
    threads = [threading.Thread(target=ask) for _ in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors and len(got) == 8
    assert all(g[0] is got[0][0] and g[1] is got[0][1] for g in got)
    assert train.tree_draws(n, T, 1234)[0] is got[0][0]          # and later callers
    counts, seeds = got[0]
    for t, (c, s) in enumerate(ffo.tree_draws(n, T, 1234)):
        assert np.array_equal(counts[t], c) and int(seeds[t]) == s


def test_stage_timeline_records_intervals(tmp_path, monkeypatch):
    """pcm.stages with PCM_STAGE_TIMELINE: every stage interval is kept and dumped as JSON (tools/sweep_timeline.py)."""
    import importlib
    import json
    import time
    path = tmp_path / "timeline.json"
    monkeypatch.setenv("PCM_STAGE_TIMELINE", str(path))
    from pcm import stages
    stages = importlib.reload(stages)
    try:
        stages.reset()
        with stages.stage("outer"):
            with stages.stage("inner"):
                time.sleep(0.01)
        assert stages.dump_timeline(".x") == str(path) + ".x"
        events = json.load(open(str(path) + ".x"))
        assert [e["stage"] for e in events] == ["inner", "outer"]
        inner, outer = events
        assert outer["start"] <= inner["start"] <= inner["end"] <= outer["end"] and inner["end"] - inner["start"] >= 0.009
        assert stages.snapshot()["outer"] >= stages.snapshot()["inner"] > 0
    finally:
        monkeypatch.delenv("PCM_STAGE_TIMELINE")
        importlib.reload(stages)
