"""pcm/priors.py: reusing the previous frame's unmasked SIFT result + a mask filter gives exactly
what the reference's second detectAndCompute(prevFrame, prevForegroundMask) gives (CPU only)."""
import cv2 as cv
import numpy as np

from helpers import read_video


def test_masked_sift_is_a_subset_of_the_unmasked_one():
    from pcm.priors import SiftPrior
    frames = read_video("Video", "frog")[:6]
    truth = read_video("Truth", "frog")[:6]
    cached, plain = SiftPrior(reuse=True), SiftPrior(reuse=False)
    prev_crop = prev_mask = None
    checked = 0
    for f, t in zip(frames, truth):
        crop = f[40:240, 150:470]
        mask = cv.dilate((cv.cvtColor(t, cv.COLOR_BGR2GRAY)[40:240, 150:470] > 127).astype(np.uint8) * 255,
                         np.ones((7, 7), np.uint8))
        if prev_crop is not None:
            a = cached.features(prev_crop, prev_mask, crop)
            b = plain.features(prev_crop, prev_mask, crop)
            assert len(b[0]) > 10, "test needs keypoints inside the mask"
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
            checked += 1
        else:
            cached.features(crop, mask, crop)          # primes the cache like the first update() would not:
            cached._last = None                         # the reference computes no prior at index 0
            cached.features(crop, np.zeros_like(mask), crop)
        prev_crop, prev_mask = crop, mask
    assert checked == 5


def _soldier_pairs(n=8):
    """(prev crop, prev mask, crop, label map) of consecutive soldier frames, masks from the truth clip."""
    from pcm.providers import grid_segments
    frames = read_video("Video", "soldier")[:n + 1]
    truth = read_video("Truth", "soldier")[:n + 1]
    box = (slice(0, 224), slice(140, 400))
    out = []
    for i in range(1, n + 1):
        prev, cur = frames[i - 1][box], frames[i][box]
        mask = cv.dilate((cv.cvtColor(truth[i - 1], cv.COLOR_BGR2GRAY)[box] > 127).astype(np.uint8) * 255, np.ones((7, 7), np.uint8))
        out.append((prev, mask, cur, grid_segments(cur, 8)))
    return out


def test_exact_matcher_restatement_against_the_reference_flann_path():
    """oracle/prior_oracle.py (what the GPU prior computes: EXACT 2-nearest neighbours) against the reference's own
    computePriors arithmetic with OpenCV's FLANN kd-trees (pcm/priors.SiftPrior, the port of :129-163).  FLANN is
    approximate and randomised, so the two are not identical -- two FLANN runs are not either; they must agree on
    almost every superpixel, and the exact matcher must not be further from a FLANN run than another FLANN run is."""
    import prior_oracle as po
    from pcm.priors import SiftPrior
    sift = cv.SIFT_create()
    diff_exact, diff_flann, total, positives = 0, 0, 0, 0
    for prev, mask, cur, seg in _soldier_pairs():
        S = int(seg.max()) + 1
        a = SiftPrior(reuse=False)(prev, mask, cur, seg, S)
        b = SiftPrior(reuse=False)(prev, mask, cur, seg, S)
        k1, d1 = sift.detectAndCompute(np.ascontiguousarray(prev), None)
        k2, d2 = sift.detectAndCompute(np.ascontiguousarray(cur), None)
        p1 = np.array([k.pt for k in k1], np.float32).reshape(-1, 2)
        p2 = np.array([k.pt for k in k2], np.float32).reshape(-1, 2)
        e = po.compute_priors(p1, po.as_u8_descriptors(d1), mask, p2, po.as_u8_descriptors(d2), seg, S)
        assert e.dtype == np.float32 and set(np.unique(e).tolist()) <= {-1.0, 1.0}
        diff_exact += int(np.count_nonzero(e != a))
        diff_flann += int(np.count_nonzero(a != b))
        total += S
        positives += int(np.count_nonzero(a == 1))
    assert positives > 40, "the test needs matched keypoints"
    assert diff_exact <= max(12, 3 * diff_flann + 8), (diff_exact, diff_flann, total)
    assert diff_exact / total < 0.01
