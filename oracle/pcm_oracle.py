"""CPU oracle for the PixelClassification (PC) masker per-frame hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`non-rigid-object-tracking_b200/`) imports this file; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may use it, and only as the checker / baseline.

It restates, in numpy (integer/byte work) and small Python loops, what the
reference computes per frame.  Every function cites the reference lines it
follows.  Reference = /root/reference (materight/non-rigid-object-tracking);
third-party arithmetic on the path (OpenCV 4.13 `cvtColor`/`dilate`,
scikit-learn 1.9 forests/PCA, numba, numpy) is restated from first principles
and PINNED in tests/ against
  * the installed libraries themselves (exhaustive 2^24-colour images for the
    colour conversions, `predict_proba` equality for the forests, `cv2.dilate`),
  * golden vectors produced by running the UNMODIFIED reference class in the
    build container (tests/golden/make_golden.py -> tests/golden/*.npz).

Two layers are provided on purpose:
  * "restatement" functions (`bgr2hsv`, `bgr2lab`, `forest_p1`, ...): the
    arithmetic itself, integer-exact, no cv2/sklearn calls;
  * `RefPortMasker`: the reference's `update()` control flow with the same
    third-party calls and the same cost structure (materialised X, sklearn
    `predict_proba`) -- this is what `bench.py` times as the CPU baseline.
"""
import numpy as np

# --------------------------------------------------------------------------
# colour conversions (reference: maskers/pixel_classification.py:294-309 ->
# cv.cvtColor(COLOR_BGR2HSV / COLOR_BGR2LAB); main.py:285 -> COLOR_BGR2GRAY)
# --------------------------------------------------------------------------

HSV_SHIFT = 12


def hsv_tables():
    """sdiv[v] = rint(255*4096/v), hdiv[d] = rint(180*4096/(6 d)); [0] = 0.

    OpenCV's 8-bit RGB2HSV_b (hrange=180) fixed-point tables.
    np.rint is round-half-even like cvRound.
    """
    sdiv = np.zeros(256, np.int32)
    hdiv = np.zeros(256, np.int32)
    i = np.arange(1, 256, dtype=np.float64)
    sdiv[1:] = np.rint((255 << HSV_SHIFT) / i).astype(np.int32)
    hdiv[1:] = np.rint((180 << HSV_SHIFT) / (6.0 * i)).astype(np.int32)
    return sdiv, hdiv


def bgr2hsv(img):
    """HxWx3 u8 BGR -> HxWx3 u8 (H in [0,180), S, V).  Integer-exact."""
    sdiv, hdiv = hsv_tables()
    b = img[..., 0].astype(np.int32)
    g = img[..., 1].astype(np.int32)
    r = img[..., 2].astype(np.int32)
    v = np.maximum(np.maximum(b, g), r)
    vmin = np.minimum(np.minimum(b, g), r)
    diff = v - vmin
    # tie order: v==r first, then v==g, else b
    h = np.where(v == r, g - b, np.where(v == g, (b - r) + 2 * diff, (r - g) + 4 * diff))
    s = (diff * sdiv[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = (h * hdiv[diff] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = h + np.where(h < 0, 180, 0)
    out = np.empty(img.shape, np.uint8)
    out[..., 0] = h.astype(np.uint8)
    out[..., 1] = s.astype(np.uint8)
    out[..., 2] = v.astype(np.uint8)
    return out


LAB_SHIFT = 12
LAB_SHIFT2 = 15
LAB_GAMMA_SCALE = 2040          # 255 * (1 << 3)
LAB_CBRT_SIZE = 2041            # reachable indices 0..2040
# rint(4096 * M_sRGB->XYZ / D65 white), rows sum to 4096
LAB_COEFFS = ((1777, 1541, 778), (871, 2929, 296), (73, 448, 3575))
LAB_LSCALE = (116 * 255 + 50) // 100                       # 296
LAB_LSHIFT = (16 * 255 * (1 << LAB_SHIFT2) + 50) // 100     # 1336934


def lab_tables():
    """gamma[256] and cbrt[2041] uint16 tables of OpenCV's RGB2Lab_b (sRGB).

    gamma[i] = rint(2040 * srgb_inverse_gamma(i/255))
    cbrt[i]  = rint(32768 * f(i/2040)),  f(x) = x < 216/24389 ? x*841/108 + 16/116 : cbrt(x)
    OpenCV builds these with softfloat; a float64 evaluation differs in exactly
    two entries (49 and 628), patched here (verified over all
    2^24 colours in tests/test_oracle_color.py).
    """
    x = np.arange(256, dtype=np.float64) / 255.0
    lin = np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)
    gamma = np.rint(lin * LAB_GAMMA_SCALE).astype(np.uint16)
    t = np.arange(LAB_CBRT_SIZE, dtype=np.float64) / LAB_GAMMA_SCALE
    f = np.where(t < 216.0 / 24389.0, t * (841.0 / 108.0) + 16.0 / 116.0, np.cbrt(t))
    cb = np.rint(f * (1 << LAB_SHIFT2)).astype(np.uint16)
    cb[49] = 9454
    cb[628] = 22126
    return gamma, cb


def bgr2lab(img):
    """HxWx3 u8 BGR -> HxWx3 u8 (L, a, b) as OpenCV 8-bit Lab.  Integer-exact."""
    gamma, cb = lab_tables()
    B = gamma[img[..., 0]].astype(np.int64)
    G = gamma[img[..., 1]].astype(np.int64)
    R = gamma[img[..., 2]].astype(np.int64)
    rnd = 1 << (LAB_SHIFT - 1)
    C = LAB_COEFFS
    fX = cb[(R * C[0][0] + G * C[0][1] + B * C[0][2] + rnd) >> LAB_SHIFT].astype(np.int64)
    fY = cb[(R * C[1][0] + G * C[1][1] + B * C[1][2] + rnd) >> LAB_SHIFT].astype(np.int64)
    fZ = cb[(R * C[2][0] + G * C[2][1] + B * C[2][2] + rnd) >> LAB_SHIFT].astype(np.int64)
    rnd2 = 1 << (LAB_SHIFT2 - 1)
    L = (LAB_LSCALE * fY - LAB_LSHIFT + rnd2) >> LAB_SHIFT2
    a = (500 * (fX - fY) + 128 * (1 << LAB_SHIFT2) + rnd2) >> LAB_SHIFT2
    b = (200 * (fY - fZ) + 128 * (1 << LAB_SHIFT2) + rnd2) >> LAB_SHIFT2
    out = np.empty(img.shape, np.uint8)
    out[..., 0] = np.clip(L, 0, 255).astype(np.uint8)
    out[..., 1] = np.clip(a, 0, 255).astype(np.uint8)
    out[..., 2] = np.clip(b, 0, 255).astype(np.uint8)
    return out


def bgr2gray(img):
    """cv.cvtColor(BGR2GRAY) 8-bit: (3735 B + 19235 G + 9798 R + 2^14) >> 15  (main.py:285)."""
    b = img[..., 0].astype(np.int32)
    g = img[..., 1].astype(np.int32)
    r = img[..., 2].astype(np.int32)
    return ((3735 * b + 19235 * g + 9798 * r + (1 << 14)) >> 15).astype(np.uint8)


# --------------------------------------------------------------------------
# feature layout (reference: pixel_classification.py:249-277, :294-309)
# --------------------------------------------------------------------------

SPACE_IDS = {"rgb": 0, "hsv": 1, "lab": 2}


def parse_features(features):
    """'8 hsv_lab' -> (n_neighbors=8, ['hsv','lab'])  (pixel_classification.py:300-308)."""
    parts = features.split()
    n = int(parts[0])
    spaces = parts[1].split("_")
    for s in spaces:
        if s not in SPACE_IDS:
            raise ValueError("unknown colour space token %r" % s)
    return n, spaces


def star_taps(n):
    """[(drow, dcol)] in the reference's order (pixel_classification.py:254-258)."""
    taps = [(0, 0)]
    for i in range(1, n + 1):
        taps += [(-i, 0), (+i, 0), (0, -i), (0, +i), (+i, +i), (-i, -i), (+i, -i), (-i, +i)]
    return taps


def build_planes(crop_bgr, spaces):
    """List of HxWx3 u8 images, one per colour-space token (buildFramesParameter :294-309)."""
    out = []
    for s in spaces:
        if s == "rgb":
            out.append(crop_bgr)
        elif s == "hsv":
            out.append(bgr2hsv(crop_bgr))
        elif s == "lab":
            out.append(bgr2lab(crop_bgr))
    return out


def get_features_int(frames, n):
    """X[h*w, F] int16 of raw tap values, -1 outside the CROP (getFeatures :263-272).

    Column index = q*K*3 + k*3 + ch.  The reference stores the same integers as
    float64 and divides by 255 afterwards (:55); tests compare X_ref == this.
    """
    h, w = frames[0].shape[:2]
    taps = star_taps(n)
    K = len(taps)
    X = np.full((h, w, len(frames) * K * 3), -1, np.int16)
    for q, fr in enumerate(frames):
        for k, (dr, dc) in enumerate(taps):
            r0, r1 = max(0, -dr), min(h, h - dr)
            c0, c1 = max(0, -dc), min(w, w - dc)
            if r0 >= r1 or c0 >= c1:
                continue
            col = q * K * 3 + k * 3
            X[r0:r1, c0:c1, col:col + 3] = fr[r0 + dr:r1 + dr, c0 + dc:c1 + dc, :]
    return X.reshape(h * w, -1)


def features_at(frames, n, rows, cols):
    """Rows of get_features_int for selected pixels only (large frames)."""
    h, w = frames[0].shape[:2]
    rows = np.asarray(rows)
    cols = np.asarray(cols)
    taps = star_taps(n)
    K = len(taps)
    X = np.full((len(rows), len(frames) * K * 3), -1, np.int16)
    for q, fr in enumerate(frames):
        for k, (dr, dc) in enumerate(taps):
            r, c = rows + dr, cols + dc
            ok = (r >= 0) & (r < h) & (c >= 0) & (c < w)
            col = q * K * 3 + k * 3
            X[ok, col:col + 3] = fr[r[ok], c[ok], :]
    return X


# --------------------------------------------------------------------------
# forest scoring (reference: pixel_classification.py:80-95 -> sklearn
# RandomForestClassifier.predict_proba)
# --------------------------------------------------------------------------

# float32(v/255) for v in -1..255: the only values a feature can take after
# `X = X / 255` (:55) and sklearn's cast of X to float32.
_FEATURE_VALUES_F32 = (np.arange(-1, 256, dtype=np.float64) / 255).astype(np.float32)


def int_threshold(thr):
    """Largest integer t in [-2,255] such that  f32(v/255) <= thr  <=>  v <= t."""
    return int(np.count_nonzero(_FEATURE_VALUES_F32.astype(np.float64) <= thr)) - 2


def forest_from_arrays(tree_arrays, n_features):
    """tree_arrays: list of (feature, threshold_f64, children_left, children_right, value1)."""
    trees = []
    for feature, threshold, left, right, value1 in tree_arrays:
        thr = np.array([int_threshold(x) if l != -1 else 0 for x, l in zip(threshold, left)], np.int32)
        trees.append(dict(feature=np.asarray(feature, np.int32).copy(), thr=thr,
                          left=np.asarray(left, np.int32).copy(),
                          right=np.asarray(right, np.int32).copy(),
                          p1=np.asarray(value1, np.float64).copy()))
    return dict(trees=trees, n_features=int(n_features))


def class1_fraction(tree):
    """P(class 1) per node as predict_proba reports it: tree_.value holds class fractions from
    scikit-learn 1.4 on (returned unchanged) and weighted class counts before (the reference pins
    0.24.1), where predict_proba divides by the row sum with a zero sum replaced by 1."""
    v = np.asarray(tree.value[:, 0, :], np.float64)
    s = v.sum(axis=1)
    if np.all(np.abs(s - 1.0) < 1e-9):
        return v[:, 1].copy()
    s[s == 0.0] = 1.0
    return v[:, 1] / s


def sklearn_tree_arrays(clf):
    """Raw per-tree arrays of a fitted sklearn RandomForestClassifier (estimators_ order);
    value1 = P(class 1) of the node as predict_proba reports it (class1_fraction)."""
    if list(clf.classes_) != [0, 1]:
        raise ValueError("forest must be a binary {0,1} classifier (reference indexes probs[:,1])")
    out = []
    for est in clf.estimators_:
        t = est.tree_
        out.append((t.feature.copy(), t.threshold.copy(), t.children_left.copy(),
                    t.children_right.copy(), class1_fraction(t)))
    return out


def export_forest(clf):
    """Flatten a fitted sklearn RandomForestClassifier: per tree feature[], thr_int[],
    left[], right[], p1[]; root = 0, leaf <=> left == -1."""
    return forest_from_arrays(sklearn_tree_arrays(clf), clf.n_features_in_)


def forest_p1(forest, Xint):
    """P(class 1) per row, bit-equal to clf.predict_proba(X/255)[:,1].

    Each tree is traversed with INTEGER compares v <= t; the forest result is a
    float64 sum of leaf fractions in estimator order divided by n_estimators
    (sklearn ensemble/_forest.py: `all_proba += prediction` then `/= n`).
    """
    N = Xint.shape[0]
    acc = np.zeros(N, np.float64)
    rows = np.arange(N)
    for tr in forest["trees"]:
        node = np.zeros(N, np.int64)
        left, right, feat, thr = tr["left"], tr["right"], tr["feature"], tr["thr"]
        while True:
            is_int = left[node] != -1
            if not is_int.any():
                break
            idx = rows[is_int]
            nd = node[idx]
            v = Xint[idx, feat[nd]]
            node[idx] = np.where(v <= thr[nd], left[nd], right[nd])
        acc += tr["p1"][node]
    return acc / len(forest["trees"])


# --------------------------------------------------------------------------
# novelty (PCA reconstruction L1 error) and temporal blend
# (reference: pixel_classification.py:57-63, :81-95)
# --------------------------------------------------------------------------

def novelty_error(Xint, mean, comps):
    """sum_f |x - inverse_transform(transform(x))| with x = Xint/255, float64.

    sklearn PCA (whiten=False): t = X @ C.T - mean @ C.T ; Xhat = t @ C + mean.
    """
    X = Xint.astype(np.float64) / 255
    t = X @ comps.T - mean.reshape(1, -1) @ comps.T
    Xhat = t @ comps + mean
    return np.sum(np.abs(X - Xhat), axis=1)


def blend(a_cur, a_next, tau):
    """np.average([a_cur, a_next], axis=0, weights=[1-tau, tau])  (:87, :93)."""
    w0, w1 = 1 - tau, tau
    return (a_cur * w0 + a_next * w1) / (w0 + w1)


# --------------------------------------------------------------------------
# superpixel decision, dilation, IoU
# (reference: pixel_classification.py:97-112, :230-246; benchmark.py:8-14)
# --------------------------------------------------------------------------

def saliency_scores(p1, sa, segments, outlier_threshold, priors, prior_weight):
    """Per-label score exactly as numba's compileSaliencyMap computes it (:233-241).

    acc is a float32 array; every `+=` is evaluated in float64 and rounded back
    to float32, strictly in raster order.  Then
      s = f32( (f64(acc)/area) * (1-w) + f64(prior_f32) * w ).
    Returns (scores f32[S], areas int64[S]); labels must be 0..S-1.
    """
    seg = segments.reshape(-1)
    S = int(seg.max()) + 1
    areas = np.bincount(seg, minlength=S).astype(np.int64)
    thr = float(outlier_threshold)
    d = p1.astype(np.float64) - (np.maximum(sa.reshape(-1).astype(np.float64), thr) - thr)
    acc = np.zeros(S, np.float32)
    # sequential per label in raster order == sequential over all pixels
    order = np.argsort(seg, kind="stable")
    starts = np.concatenate([[0], np.cumsum(areas)])
    for lab in range(S):
        a = np.float32(0)
        for v in d[order[starts[lab]:starts[lab + 1]]]:
            a = np.float32(np.float64(a) + v)
        acc[lab] = a
    w = float(prior_weight)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = (acc.astype(np.float64) / areas) * (1 - w) + priors.astype(np.float32).astype(np.float64) * w
    return s.astype(np.float32), areas


def saliency_scores_fast(p1, segments):
    """float64 per-label means (novelty off, no prior): NOT the reference's float32
    accumulation -- only for large-size property tests, which exclude labels whose
    score is within 1e-4 of 0.5."""
    seg = segments.reshape(-1)
    S = int(seg.max()) + 1
    areas = np.bincount(seg, minlength=S)
    sums = np.bincount(seg, weights=p1, minlength=S)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (sums / areas).astype(np.float32), areas


def saliency_mask(scores, segments):
    """u8 map: 255 for every pixel of a label whose score > 0.5 (:242-245)."""
    return np.where(scores[segments] > 0.5, 255, 0).astype(np.uint8)


def dilate(mask, k):
    """cv.dilate(mask, ones((k,k)), iterations=1): max over the k x k window with
    anchor (k//2, k//2); neighbours outside the array are ignored (:112)."""
    h, w = mask.shape
    a = k // 2
    out = np.zeros_like(mask)
    for dy in range(-a, k - a):
        for dx in range(-a, k - a):
            r0, r1 = max(0, -dy), min(h, h - dy)
            c0, c1 = max(0, -dx), min(w, w - dx)
            if r0 >= r1 or c0 >= c1:
                continue
            np.maximum(out[r0:r1, c0:c1], mask[r0 + dy:r1 + dy, c0 + dx:c1 + dx], out=out[r0:r1, c0:c1])
    return out


def iou_counts(mask, truth):
    """(intersection, union) pixel counts of (mask != 0), (truth != 0)  (benchmark.py:12-13)."""
    m = mask != 0
    t = truth != 0
    return int(np.count_nonzero(m & t)), int(np.count_nonzero(m | t))


def iou(mask, truth):
    i, u = iou_counts(mask, truth)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.float64(i) / np.float64(u)


def enlarge_bbox(bbox, frame_shape, enlarge=20):
    """The reference's exact (quirky) enlargement (:49-51) followed by numpy slice
    clamping.  Returns (x, y, w, h) of the crop actually sliced."""
    H, W = frame_shape[:2]
    x = max(bbox[0] - enlarge, 0)
    y = max(bbox[1] - enlarge, 0)
    w = min(bbox[0] + bbox[2] + enlarge, W) - bbox[0] + enlarge
    h = min(bbox[1] + bbox[3] + enlarge, H) - bbox[1] + enlarge
    return x, y, w, h


def slice_extent(start, length, size):
    """numpy semantics of a[start:start+length] for start >= 0: (start, stop) clamped."""
    stop = start + length
    if stop < 0:
        stop = max(size + stop, 0)
    return min(start, size), min(max(stop, 0), size)
