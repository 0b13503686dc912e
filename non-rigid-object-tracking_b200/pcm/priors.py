"""SIFT-match prior of the PC masker (reference maskers/pixel_classification.py:129-163,
`computePriors`).  Host-side OpenCV, outside the GPU hot path ("next" row f-2 of
SURVEY.md §8): only used when `prior_weight != 0` and `index > 0`.

A label gets prior +1 when a SIFT keypoint of the current crop that matches a
keypoint inside the previous foreground mask (Lowe ratio 0.7, 90th-percentile
displacement filter) falls on it; every other label keeps -1.
"""
import cv2 as cv
import numpy as np


class SiftPrior:
    def __init__(self):
        self.sift = cv.SIFT_create()
        self.flann = cv.FlannBasedMatcher(dict(algorithm=1, trees=5), dict(checks=50))

    def __call__(self, prev_crop, prev_mask, crop, segments, n_labels):
        priors = np.full(n_labels, -1, np.float32)
        if prev_crop is None or prev_mask is None:
            return priors
        kp1, des1 = self.sift.detectAndCompute(np.ascontiguousarray(prev_crop), np.ascontiguousarray(prev_mask))
        kp2, des2 = self.sift.detectAndCompute(np.ascontiguousarray(crop), None)
        if len(kp1) == 0 or len(kp2) < 2:
            return priors
        good = [m for m, n in (pair for pair in self.flann.knnMatch(des1, des2, k=2) if len(pair) == 2)
                if m.distance < 0.7 * n.distance]
        if not good:
            return priors
        p1 = np.array([kp1[m.queryIdx].pt for m in good])
        p2 = np.array([kp2[m.trainIdx].pt for m in good])
        dist = np.sqrt(((p2 - p1) ** 2).sum(axis=1))
        keep = dist <= np.percentile(dist, 90)
        for px, py in p2[keep]:
            priors[segments[int(py), int(px)]] = 1
        return priors
