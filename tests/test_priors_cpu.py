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
