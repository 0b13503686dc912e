"""CPU restatement of the SIFT-match prior as the GPU computes it -- TEST INFRASTRUCTURE ONLY (see pcm_oracle.py header).

Follows the reference's computePriors (maskers/pixel_classification.py:129-163) step for step, with ONE stated
difference: the reference asks OpenCV's FLANN matcher (randomised kd-trees, `trees=5, checks=50`, :40-43) for
APPROXIMATE 2-nearest neighbours -- its answers change from run to run and no golden vector can pin them -- while this
restatement (and csrc/pcm_prior.cuh) takes the EXACT 2-nearest neighbours FLANN approximates (ties to the smaller
index).  Everything else is the reference's arithmetic:
  * previous-crop keypoints = the unmasked detection filtered like cv::KeyPointsFilter::runByPixelsMask does inside
    detectAndCompute(prev, prevMask): keep mask[(int)(y + 0.5f), (int)(x + 0.5f)] != 0   (:135)
  * m.distance, n.distance = float32 sqrt of the float32 squared L2 distance, compared as Python floats with 0.7 (:147)
  * displacement in float64 from the float32 keypoint coordinates, np.percentile(dist, 90), dist <= thrs (:155-157)
  * priors[segments[int(y), int(x)]] = 1 (:159-161); fewer than 1 previous / 2 current keypoints -> all -1 (:139)
`tests/test_priors_cpu.py` measures how often the exact search and FLANN disagree on the shipped clips.
"""
import numpy as np


def as_u8_descriptors(des):
    """OpenCV SIFT descriptors (float32 holding integers 0..255) -> uint8; raises if they are not."""
    if des is None or len(des) == 0:
        return np.zeros((0, 128), np.uint8)
    d = np.asarray(des)
    if d.shape[1] != 128 or not np.array_equal(d, np.rint(d)) or d.min() < 0 or d.max() > 255:
        raise ValueError("SIFT descriptors are expected to be 128 integers in 0..255")
    return d.astype(np.uint8)


def exact_knn2(des1, des2):
    """(j1, d1, d2): nearest neighbour index, squared distances of the two nearest (int64), ties to the smaller index."""
    a = des1.astype(np.int64)
    b = des2.astype(np.int64)
    d = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
    order = np.argsort(d, axis=1, kind="stable")[:, :2]
    rows = np.arange(len(a))
    return order[:, 0], d[rows, order[:, 0]], d[rows, order[:, 1]]


def compute_priors(pts_prev, des_prev, prev_mask, pts_cur, des_cur, segments, n_labels):
    priors = np.full(n_labels, -1, np.float32)
    pts_prev = np.asarray(pts_prev, np.float32).reshape(-1, 2)
    pts_cur = np.asarray(pts_cur, np.float32).reshape(-1, 2)
    if len(pts_prev) == 0 or len(pts_cur) < 2:
        return priors
    yy = (pts_prev[:, 1] + np.float32(0.5)).astype(np.int32)
    xx = (pts_prev[:, 0] + np.float32(0.5)).astype(np.int32)
    inside = (yy >= 0) & (xx >= 0) & (yy < prev_mask.shape[0]) & (xx < prev_mask.shape[1])
    keep = np.zeros(len(pts_prev), bool)
    keep[inside] = prev_mask[yy[inside], xx[inside]] != 0
    p1, d1 = pts_prev[keep], des_prev[keep]
    if len(p1) == 0:
        return priors
    j1, s1, s2 = exact_knn2(d1, des_cur)
    m_dist = np.sqrt(s1.astype(np.float32)).astype(np.float64)
    n_dist = np.sqrt(s2.astype(np.float32)).astype(np.float64)
    good = m_dist < 0.7 * n_dist
    if not good.any():
        return priors
    a = p1[good].astype(np.float64)
    b = pts_cur[j1[good]].astype(np.float64)
    dist = np.sqrt((b[:, 0] - a[:, 0]) ** 2 + (b[:, 1] - a[:, 1]) ** 2)
    thr = np.percentile(dist, 90)
    for px, py in b[dist <= thr]:
        priors[segments[int(py), int(px)]] = 1
    return priors
