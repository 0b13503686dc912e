"""CPU port of the reference PC masker with the reference's own cost structure.

TEST INFRASTRUCTURE ONLY (see oracle/pcm_oracle.py header).  Used as
  * the CPU baseline that bench.py times (`cpu_baseline.kind == "port"`,
    `bench.py --impl reference`): same third-party calls as the reference
    (cv2.cvtColor, a numba-jitted tap gather that materialises X[N,F] float64,
    `X/255`, sklearn `predict_proba`, sklearn PCA, numba saliency loop,
    cv2.dilate) in the same order, single thread;
  * a second checker for the GPU path (it shares no code with
    pcm_oracle.py's integer restatement).

Follows /root/reference/maskers/pixel_classification.py:24-309 (class
PixelClassificationNonRigidMasker) and maskers/masker.py:3-14.  PINNED against
the unmodified reference by tests/test_oracle_golden.py through the golden
vectors in tests/golden/*.npz (made by tests/golden/make_golden.py).

Not restated: over-segmentation (skimage, absent -> `segment_fn` injected) and
the SIFT prior (`computePriors`, "next" row SURVEY §8 f-2 -> `prior_fn`
injected, default: all -1 as the reference returns for prior_weight == 0).
"""
import copy

import cv2 as cv
import numpy as np
from numba import njit
from sklearn.decomposition import PCA
from sklearn.ensemble import RandomForestClassifier


@njit(cache=True)
def _gather_star(planes, n_neighbors, X):
    """planes: [Q,h,w,3] u8 -> X[h*w, Q*K*3] (pre-filled with -1.0)  (:249-277)."""
    Q, h, w, _ = planes.shape
    K = 1 + 8 * n_neighbors
    dr = np.zeros(K, np.int64)
    dc = np.zeros(K, np.int64)
    k = 1
    for i in range(1, n_neighbors + 1):
        for a, b in ((-i, 0), (i, 0), (0, -i), (0, i), (i, i), (-i, -i), (i, -i), (-i, i)):
            dr[k] = a
            dc[k] = b
            k += 1
    for q in range(Q):
        base = q * K * 3
        for r in range(h):
            for c in range(w):
                row = r * w + c
                for k in range(K):
                    rr = r + dr[k]
                    cc = c + dc[k]
                    if rr >= 0 and rr < h and cc >= 0 and cc < w:
                        X[row, base + 3 * k + 0] = planes[q, rr, cc, 0]
                        X[row, base + 3 * k + 1] = planes[q, rr, cc, 1]
                        X[row, base + 3 * k + 2] = planes[q, rr, cc, 2]


@njit(cache=True)
def _saliency(p1, sa, segments, thr, areas, priors, prior_weight, out):
    """compileSaliencyMap (:230-246): float32 per-label accumulator, raster order."""
    h, w = segments.shape
    acc = np.zeros(areas.shape[0], np.float32)
    c = 0
    for i in range(h):
        for j in range(w):
            acc[segments[i, j]] += p1[c] - (max(sa[i, j], thr) - thr)
            c += 1
    on = np.zeros(areas.shape[0], np.bool_)
    for key in range(areas.shape[0]):
        if areas[key] > 0:
            acc[key] = (acc[key] / areas[key]) * (1 - prior_weight) + priors[key] * prior_weight
            on[key] = acc[key] > 0.5
    for i in range(h):
        for j in range(w):
            out[i, j] = 255 if on[segments[i, j]] else 0
    return acc


class RefPortMasker:
    """Same constructor/`addModel`/`update` protocol as the reference class."""

    def __init__(self, debug=False, frame=None, config=None, poly_roi=None, update_mask=None,
                 segment_fn=None, prior_fn=None, **others):
        self.debug = debug
        self.prevFrame = frame.copy() if frame is not None else None      # masker.py:7
        self.config = config
        self.poly_roi = copy.deepcopy(poly_roi)
        self.index = 0
        self.models = []
        self.novelty_det = []
        self.current_model = 0
        self.multi_selection = self.config.get("multi_selection")
        self.segment_fn = segment_fn
        self.prior_fn = prior_fn
        self.last = {}          # stage dumps of the latest update (parity tests)

    # -- features ---------------------------------------------------------
    def _planes(self, crop):
        params = self.config["params"]["features"].split()
        frames = []
        for e in params[1].split("_"):
            if e == "rgb":
                frames.append(crop)
            elif e == "hsv":
                frames.append(cv.cvtColor(crop, cv.COLOR_BGR2HSV))
            elif e == "lab":
                frames.append(cv.cvtColor(crop, cv.COLOR_BGR2LAB))
        return np.ascontiguousarray(np.stack(frames)), int(params[0])

    def features(self, crop):
        planes, n = self._planes(crop)
        Q, h, w, _ = planes.shape
        X = np.full((h * w, Q * (1 + 8 * n) * 3), -1.0)
        _gather_star(planes, n, X)
        return X

    # -- training (:166-228) ------------------------------------------------
    def addModel(self, frame, poly_roi, bbox, n_frame, bbox_roni=None, show_prob_map=False):
        if bbox_roni is None:
            raise ValueError("headless port: bbox_roni is required")
        crop = frame[bbox[1]:bbox[1] + bbox[3], bbox[0]:bbox[0] + bbox[2]]
        pts = np.array([[(p[0] - bbox[0], p[1] - bbox[1]) for p in poly_roi]], dtype=np.int32)
        roi = np.zeros([bbox[3], bbox[2]], dtype=np.uint8)
        cv.fillPoly(roi, pts, 255)
        X = self.features(crop)
        y = (roi.reshape(-1) > 0).astype(np.int64)
        roni = frame[bbox_roni[1]:bbox_roni[1] + bbox_roni[3], bbox_roni[0]:bbox_roni[0] + bbox_roni[2]]
        Xn = self.features(roni)
        X = np.concatenate([X, Xn], axis=0) / 255
        y = np.concatenate([y, np.zeros(Xn.shape[0], np.int64)])
        p = self.config["params"]
        clf = RandomForestClassifier(random_state=42, n_estimators=p["n_estimators"],
                                     max_depth=p["max_depth"]).fit(X, y)
        if p["novelty_detection"]:
            pca = PCA(n_components=p["n_components"]).fit(X[y == 1])
            err = np.sum(np.sqrt(np.power(X - pca.inverse_transform(pca.transform(X)), 2)), axis=1)
            threshold = np.percentile(err, 90)
        else:
            pca, threshold = None, 0.0
        self.models.append({"n_frame": n_frame, "model": clf})
        self.novelty_det.append({"n_frame": n_frame, "model": pca, "threshold": threshold})
        return bbox_roni

    # -- per-frame hot path (:45-126) ---------------------------------------
    def update(self, bbox, frame, mask, color=None):
        e = 20
        bbox = (max(bbox[0] - e, 0), max(bbox[1] - e, 0),
                min(bbox[0] + bbox[2] + e, frame.shape[1]) - bbox[0] + e,
                min(bbox[1] + bbox[3] + e, frame.shape[0]) - bbox[1] + e)
        ys, xs = slice(bbox[1], bbox[1] + bbox[3]), slice(bbox[0], bbox[0] + bbox[2])
        crop = frame[ys, xs]
        h, w = crop.shape[:2]
        p = self.config["params"]
        X = self.features(crop)
        X = X / 255
        cur = self.current_model
        novelty = bool(p["novelty_detection"])

        def recon_err(pca):
            return np.sum(np.sqrt(np.power(X - pca.inverse_transform(pca.transform(X)), 2)), axis=1)

        if novelty:
            sa = recon_err(self.novelty_det[cur]["model"]).reshape(h, w).copy()
        else:
            sa = np.zeros((h, w), dtype=np.uint8)
        segments = np.ascontiguousarray(self.segment_fn(crop))
        probs = self.models[cur]["model"].predict_proba(X)
        if self.multi_selection and len(self.models) > cur + 1:
            probs_next = self.models[cur + 1]["model"].predict_proba(X)
            span = self.models[cur + 1]["n_frame"] - self.models[cur]["n_frame"]
            tmp = self.index - self.models[cur]["n_frame"]
            wts = [1 - (tmp / span), tmp / span]
            probs = np.average([probs, probs_next], axis=0, weights=wts)
            if novelty:
                sa_next = recon_err(self.novelty_det[cur + 1]["model"]).reshape(h, w).copy()
                sa = np.average([sa, sa_next], axis=0, weights=wts)
        labels, areas = np.unique(segments, return_counts=True)
        if labels[0] != 0 or labels[-1] != len(labels) - 1:
            raise ValueError("labels must be contiguous 0..S-1 (reference indexes by label, :237,:241)")
        if self.prior_fn is not None and self.index != 0 and p["prior_weight"] != 0.0:
            priors = np.asarray(self.prior_fn(self, crop, segments, labels), np.float32)
        else:
            priors = np.full(labels.shape, -1, np.float32)
        sal = np.zeros((h, w), np.uint8)
        p1 = np.ascontiguousarray(probs[:, 1])
        scores = _saliency(p1, sa, segments, float(self.novelty_det[cur]["threshold"]), areas, priors,
                           float(p["prior_weight"]), sal)
        mask[ys, xs, 2] = sal
        k = p["dilation_kernel"]
        mask[ys, xs, 2] = cv.dilate(mask[ys, xs, 2], np.ones((k, k), np.uint8), iterations=1)
        self.last = dict(bbox=bbox, p1=p1, sa=np.asarray(sa, np.float64), segments=segments, areas=areas,
                         priors=priors, scores=scores, saliency=sal)
        self.index += 1
        self.prevFrame = crop
        self.prevForegroundMask = mask[ys, xs, 2]
        if self.multi_selection and len(self.models) > cur + 1 and self.index >= self.models[cur + 1]["n_frame"]:
            self.current_model += 1
            return self.current_model
        return None


class SiftFlannPrior:
    """Port of computePriors (:129-163) for `prior_fn`: cv.SIFT on the previous crop (inside the previous mask) and on
    the current crop, FLANN kd-tree 2-NN (algorithm 1, trees 5, checks 50, :40-43), Lowe ratio 0.7, matches farther
    than the 90th percentile of the displacement dropped, +1 for the superpixel under every surviving keypoint."""

    def __init__(self):
        self.sift = cv.SIFT_create()
        self.flann = cv.FlannBasedMatcher(dict(algorithm=1, trees=5), dict(checks=50))

    def __call__(self, masker, crop, segments, labels):
        priors = np.full(labels.shape, -1, np.float32)
        kp1, des1 = self.sift.detectAndCompute(np.ascontiguousarray(masker.prevFrame), np.ascontiguousarray(masker.prevForegroundMask))
        kp2, des2 = self.sift.detectAndCompute(np.ascontiguousarray(crop), None)
        if len(kp1) == 0 or len(kp2) < 2:          # the reference unpacks (m, n) pairs and needs two neighbours
            return priors
        good = [pair[0] for pair in self.flann.knnMatch(des1, des2, k=2) if len(pair) == 2 and pair[0].distance < 0.7 * pair[1].distance]
        if not good:
            return priors
        dist = np.array([((kp2[m.trainIdx].pt[0] - kp1[m.queryIdx].pt[0]) ** 2 +
                          (kp2[m.trainIdx].pt[1] - kp1[m.queryIdx].pt[1]) ** 2) ** 0.5 for m in good])
        thr = np.percentile(dist, 90)
        for m in np.array(good)[dist <= thr]:
            priors[segments[int(kp2[m.trainIdx].pt[1]), int(kp2[m.trainIdx].pt[0])]] = 1
        return priors


def compute_benchmark(mask, truth):
    """benchmark.py:8-14 (np.bool there; removed from numpy>=1.24 except as alias in 2.x)."""
    bm, bt = mask.astype(bool), truth.astype(bool)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.sum(bm & bt) / np.sum(bm | bt)
