"""pcm_fit_forest (csrc/pcm_forest_fit.cuh): the forest of addModel grown on the GPU must be the forest
scikit-learn grows (reference maskers/pixel_classification.py:199-200) -- EQUAL tree_ arrays -- and equal to the
CPU restatement in oracle/forest_fit_oracle.c."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from test_forest_fit_oracle import _sk_fit, assert_trees_equal_sklearn, synthetic_rows  # noqa: E402


@pytest.fixture(scope="module")
def handle():
    from pcm import capi
    h = capi.Handle(0)
    yield h
    h.close()


@pytest.mark.parametrize("n,F,T,D,constants", [(3000, 147, 6, 4, False), (2500, 147, 8, 10, False), (2500, 390, 6, 7, True),
                                                (50, 27, 12, 24, True), (2, 5, 3, 5, False), (400, 1, 5, 6, False),
                                                (9000, 390, 30, 10, False), (700, 1161, 4, 5, True), (300, 30, 3, 1, False)])
def test_gpu_forest_is_scikit_learns_forest(handle, n, F, T, D, constants):
    from pcm import train
    import forest_fit_oracle as ffo
    assert train.gpu_fit_supported(), "the GPU trainer restates scikit-learn 1.9's tree builder"
    Xi, y = synthetic_rows(n + F, n, F, constants)
    got = train.fit_forest(handle, Xi, y, T, D)
    assert_trees_equal_sklearn(_sk_fit(Xi, y, T, D), got.trees)
    want = ffo.as_tree_arrays(ffo.fit_forest(Xi, y, T, D))
    for a, b in zip(got.trees, want):
        assert all(np.array_equal(u, v) for u, v in zip(a, b))


def test_rows_stay_resident_between_fits(handle):
    from pcm import train
    Xi, y = synthetic_rows(11, 4000, 147)
    a7 = train.fit_forest(handle, Xi, y, 5, 7, rows_id=41)
    a10 = train.fit_forest(handle, Xi, y, 5, 10, rows_id=41, rows_resident=True)
    assert_trees_equal_sklearn(_sk_fit(Xi, y, 5, 7), a7.trees)
    assert_trees_equal_sklearn(_sk_fit(Xi, y, 5, 10), a10.trees)
    from pcm import capi
    with pytest.raises(capi.PcmError):
        train.fit_forest(handle, Xi, y, 5, 7, rows_id=42, rows_resident=True)


@pytest.mark.parametrize("video,features", [("soldier", "8 hsv_lab"), ("bmx", "6 lab"), ("frog", "8 hsv_lab")])
def test_clip_models_equal_scikit_learn(video, features):
    """The sweep's real training sets (polygon bbox + RONI rows gathered on the device, :170-197), 30 trees, depth 10."""
    import cv2 as cv
    from helpers import polygons, read_video
    from pcm import capi, train
    P = polygons()[video]
    frames = read_video("Video", video)
    h = capi.Handle(0)
    tok = features.split()
    h.set_features(int(tok[0]), tok[1].split("_"))
    s = 1
    pts, roni = P["pts"][0][s], P["bboxes_roni"][0][s]
    f = frames[P["pts_frame_numbers"][s]]
    x, y, w, hh = cv.boundingRect(np.array(pts))
    roi = np.zeros((hh, w), np.uint8)
    cv.fillPoly(roi, np.array([[(p[0] - x, p[1] - y) for p in pts]], dtype=np.int32), 255)
    X = np.concatenate([h.gather_features(f, (x, y, w, hh)), h.gather_features(f, tuple(roni))])
    lab = np.concatenate([(roi.reshape(-1) > 0).astype(np.int64), np.zeros(roni[2] * roni[3], np.int64)])
    got = train.fit_forest(h, X, lab, 30, 10)
    from sklearn.ensemble import RandomForestClassifier
    clf = RandomForestClassifier(random_state=42, n_estimators=30, max_depth=10, n_jobs=8).fit(X.astype(np.float64) / 255, lab)
    assert_trees_equal_sklearn(clf, got.trees)
    h.close()
