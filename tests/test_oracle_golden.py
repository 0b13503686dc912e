"""Pins the oracle (oracle/pcm_oracle.py restatement AND oracle/ref_port.py port)
against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import pcm_oracle as orc
from helpers import GOLDEN, SEQ_NAMES, GoldenSeq, polygons, sha1
from pcm.providers import make_segment_provider
import os


@pytest.fixture(scope="module")
def stages():
    return np.load(os.path.join(GOLDEN, "stages.npz"))


def test_star_gather_matches_reference_getFeatures(stages):
    for ci in range(int(stages["feat_n"])):
        crop = stages["feat%d_crop" % ci]
        n, spaces = orc.parse_features(str(stages["feat%d_features" % ci]))
        X = orc.get_features_int(orc.build_planes(crop, spaces), n)
        assert X.shape == stages["feat%d_X" % ci].shape
        assert np.array_equal(X, stages["feat%d_X" % ci]), ci


def test_saliency_matches_reference_compileSaliencyMap(stages):
    for si in range(int(stages["sal_n"])):
        seg = stages["sal%d_segments" % si]
        scores, areas = orc.saliency_scores(stages["sal%d_p1" % si], stages["sal%d_sa" % si], seg,
                                            float(stages["sal%d_thr" % si]), stages["sal%d_priors" % si],
                                            float(stages["sal%d_pw" % si]))
        assert np.array_equal(orc.saliency_mask(scores, seg), stages["sal%d_map" % si]), si


def test_iou_matches_reference_computeBenchmark(stages):
    for bi in range(int(stages["iou_n"])):
        got = orc.iou(stages["iou%d_mask" % bi], stages["iou%d_truth" % bi])
        want = stages["iou%d_value" % bi]
        assert (np.isnan(got) and np.isnan(want)) or got == want


def test_bbox_quirk():
    # reference :49-51 (width/height use the UNCLAMPED x/y)
    assert orc.enlarge_bbox((5, 7, 30, 40), (100, 200, 3)) == (0, 0, 70, 80)
    assert orc.enlarge_bbox((150, 60, 45, 35), (100, 200, 3)) == (130, 40, 70, 60)
    assert orc.enlarge_bbox((-4, 50, 30, 20), (100, 200, 3)) == (0, 30, 70, 60)


@pytest.mark.parametrize("name", SEQ_NAMES)
def test_restatement_reproduces_reference_update(name):
    """forest_p1 (integer thresholds, ordered f64 sum), novelty_error, blend,
    saliency_scores, dilate: stage by stage against dumps of the reference."""
    g = GoldenSeq(name)
    if not g.frames_match():
        pytest.skip("video decoder output differs from the one the goldens were made with")
    n, spaces = orc.parse_features(g.params["features"])
    F = 3 * (1 + 8 * n) * len(spaces)
    forests = [orc.forest_from_arrays(g.tree_arrays(m), F) for m in range(g.n_models)]
    nf = g.model_frames()
    for i in g.meta["dump"]:
        z = g.z
        x, y, w, h = [int(v) for v in z["f%d_rect" % i]]
        eb = orc.enlarge_bbox(tuple(int(v) for v in z["bbox"][i]), g.frames[i].shape)
        assert (x, y) == (eb[0], eb[1])
        crop = g.frames[i][y:y + h, x:x + w]
        X = orc.get_features_int(orc.build_planes(crop, spaces), n)
        index, cur = g.state_at(i)
        p1 = orc.forest_p1(forests[cur], X)
        novelty = g.params["novelty_detection"]
        sa = orc.novelty_error(X, *g.pca(cur)) if novelty else np.zeros(h * w)
        if g.meta["multi_selection"] and cur + 1 < g.n_models:
            tau = (index - nf[cur]) / (nf[cur + 1] - nf[cur])
            p1 = orc.blend(p1, orc.forest_p1(forests[cur + 1], X), tau)
            if novelty:
                sa = orc.blend(sa, orc.novelty_error(X, *g.pca(cur + 1)), tau)
        assert np.array_equal(p1, z["f%d_p1" % i]), "P(fg) must be bit-equal (frame %d)" % i
        if novelty:
            np.testing.assert_allclose(sa, z["f%d_sa" % i].reshape(-1), rtol=1e-9, atol=0)
            sa = z["f%d_sa" % i].reshape(-1)      # decide on the reference's own values
        seg = z["f%d_segments" % i]
        assert np.array_equal(seg, make_segment_provider(g.meta["segments"])(crop))
        scores, _ = orc.saliency_scores(p1, sa, seg, g.novelty_threshold(cur), z["f%d_priors" % i],
                                        g.params["prior_weight"])
        pre = orc.saliency_mask(scores, seg)
        assert np.array_equal(pre, z["f%d_pre" % i])
        assert np.array_equal(orc.dilate(pre, g.params["dilation_kernel"]), z["f%d_post" % i])


@pytest.mark.parametrize("name", SEQ_NAMES)
def test_port_reproduces_reference_sequence(name):
    """RefPortMasker (the CPU baseline bench.py times): retrain + run the whole
    sequence; trees, return values, masks and IoU counts equal the reference's."""
    import cv2 as cv
    from ref_port import RefPortMasker, compute_benchmark
    g = GoldenSeq(name)
    if not g.frames_match():
        pytest.skip("video decoder output differs from the one the goldens were made with")
    poly = polygons()[g.meta["video"]]
    pts, ronis = poly["pts"][0], poly["bboxes_roni"][0]
    m = RefPortMasker(debug=False, frame=g.frames[0], config=g.config, poly_roi=pts[0],
                      segment_fn=make_segment_provider(g.meta["segments"]),
                      prior_fn=(lambda self_, crop, segs, labels: g.recorded_priors(self_.index))
                      if g.has_recorded_priors() else None)
    for s in range(g.n_models):
        fn = g.model_frames()[s]
        m.addModel(frame=g.frames[fn], poly_roi=pts[s], bbox=cv.boundingRect(np.array(pts[s])),
                   bbox_roni=ronis[s], n_frame=fn)
        got = orc.sklearn_tree_arrays(m.models[s]["model"])
        want = g.tree_arrays(s)
        assert len(got) == len(want)
        for a, b in zip(got, want):
            for u, v in zip(a, b):
                assert np.array_equal(u, v)
    for i in g.sample_frames():
        m.index, m.current_model = g.state_at(i)      # long goldens are sampled: the state is a function of the frame index
        mask = np.zeros_like(g.frames[i])
        ret = m.update(bbox=tuple(int(v) for v in g.z["bbox"][i]), frame=g.frames[i], mask=mask)
        assert (-1 if ret is None else ret) == int(g.z["ret"][i])
        assert sha1(mask[:, :, 2]) == str(g.z["mask_sha1"][i]), "mask differs at frame %d" % i
        tg = cv.cvtColor(g.truth[i], cv.COLOR_BGR2GRAY)
        assert orc.iou_counts(mask[:, :, 2], tg) == (int(g.z["inter"][i]), int(g.z["union"][i]))
        got = compute_benchmark(mask[:, :, 2], tg)
        assert got == g.z["iou"][i] or (np.isnan(got) and np.isnan(g.z["iou"][i]))
        if i in g.meta["dump"]:
            assert np.array_equal(m.last["p1"], g.z["f%d_p1" % i])
