"""SIFT-match prior of the PC masker (reference maskers/pixel_classification.py:129-163,
`computePriors`).  Host-side OpenCV, outside the GPU hot path ("next" row f-2 of
SURVEY.md §8): only used when `prior_weight != 0` and `index > 0`.

A label gets prior +1 when a SIFT keypoint of the current crop that matches a
keypoint inside the previous foreground mask (Lowe ratio 0.7, 90th-percentile
displacement filter) falls on it; every other label keeps -1.

The reference runs SIFT twice per frame: on the previous crop with the previous mask, and on
the current crop.  The previous crop IS the crop of the previous call, and OpenCV applies the
mask only as a keypoint filter after detection (KeyPointsFilter::runByPixelsMask) and before
the per-keypoint descriptors, so the masked result is the subset of the unmasked one computed a
frame earlier: that one is kept and filtered here instead of being recomputed (`reuse=True`).
"""
import cv2 as cv
import numpy as np


class SiftPrior:
    def __init__(self, reuse=True):
        self.sift = cv.SIFT_create()
        self.flann = cv.FlannBasedMatcher(dict(algorithm=1, trees=5), dict(checks=50))
        self.reuse = reuse
        self._last = None          # (crop array object, keypoint xy float32 [n,2], descriptors)

    @staticmethod
    def _pts(kps):
        return np.array([k.pt for k in kps], np.float32).reshape(-1, 2)

    def features(self, prev_crop, prev_mask, crop, cache=None, key=None):
        """(pts1, des1) of the previous crop inside prev_mask, (pts2, des2) of the current crop.
        `cache` / `key`: a sweep's shared cache (get_or_compute) and an identifier of the current
        crop (clip, frame, rectangle) -- its unmasked SIFT result is the same for every
        hyper-parameter combination of the sweep."""
        if self.reuse and self._last is not None and self._last[0] is prev_crop:
            pts, des = self._last[1], self._last[2]
            if len(pts):
                # KeyPointsFilter::runByPixelsMask: keep mask[(int)(y + 0.5f), (int)(x + 0.5f)] != 0
                yy = (pts[:, 1] + np.float32(0.5)).astype(np.int32)
                xx = (pts[:, 0] + np.float32(0.5)).astype(np.int32)
                keep = prev_mask[yy, xx] != 0
                pts1, des1 = pts[keep], (des[keep] if des is not None else None)
            else:
                pts1, des1 = pts, des
        else:
            kp1, des1 = self.sift.detectAndCompute(np.ascontiguousarray(prev_crop), np.ascontiguousarray(prev_mask))
            pts1 = self._pts(kp1)
        def detect():
            kp2, des2 = self.sift.detectAndCompute(np.ascontiguousarray(crop), None)
            return self._pts(kp2), des2
        if cache is not None and key is not None and hasattr(cache, "get_or_compute"):
            pts2, des2 = cache.get_or_compute(("sift",) + tuple(key), detect)
        else:
            pts2, des2 = detect()
        self._last = (crop, pts2, des2)
        return pts1, des1, pts2, des2

    def __call__(self, prev_crop, prev_mask, crop, segments, n_labels, cache=None, key=None):
        priors = np.full(n_labels, -1, np.float32)
        if prev_crop is None or prev_mask is None:
            return priors
        pts1, des1, pts2, des2 = self.features(prev_crop, prev_mask, crop, cache, key)
        if len(pts1) == 0 or len(pts2) < 2:
            return priors
        good = [m for m, n in (pair for pair in self.flann.knnMatch(des1, des2, k=2) if len(pair) == 2)
                if m.distance < 0.7 * n.distance]
        if not good:
            return priors
        p1 = np.array([pts1[m.queryIdx] for m in good], np.float64)
        p2 = np.array([pts2[m.trainIdx] for m in good], np.float64)
        dist = np.sqrt(((p2 - p1) ** 2).sum(axis=1))
        keep = dist <= np.percentile(dist, 90)
        for px, py in p2[keep]:
            priors[segments[int(py), int(px)]] = 1
        return priors
