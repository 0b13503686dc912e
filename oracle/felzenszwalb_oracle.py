"""CPU restatement of `skimage.segmentation.felzenszwalb` as the PC masker calls it
(reference maskers/pixel_classification.py:72-73:
 `felzenszwalb(crop_frame, scale=100, sigma=0.5, min_size=50)`).

TEST INFRASTRUCTURE ONLY (see oracle/pcm_oracle.py header).

PARITY UNPINNED against scikit-image (environment.yaml:12 pins 0.17.2; not installed, not
vendored, no reference golden vectors).  Restated from the published algorithm of that version,
skimage/segmentation/_felzenszwalb_cy.pyx::_felzenszwalb_cython for multichannel images:

  image = img_as_float64(image); scale = scale / 255
  image = scipy.ndimage.gaussian_filter(image, sigma=[sigma, sigma, 0])        (reflect, truncate 4)
  edge costs, 8-connectivity (right, down, down-right, up-right): Euclidean colour distance
  edges sorted by cost; greedy merge of the two components of an edge while
      cost < min(Int(C0) + scale/|C0|, Int(C1) + scale/|C1|)     (Int = cost of the last merge)
  second pass over the sorted edges: merge components smaller than min_size
  labels = np.unique(root, return_inverse=True)[1]

One deliberate, documented deviation: scikit-image sorts with `np.argsort(costs)` (unstable
quicksort, tie order unspecified); this restatement -- and the product -- use a STABLE sort
(cost, then edge index in the order right, down, down-right, up-right).  Zero-cost ties never
matter (all of them merge); other exact float64 ties are rare.

The Gaussian step IS pinned: `smooth()` is compared bit for bit with scipy.ndimage (installed).
"""
import math

import numpy as np
from numba import njit


def gaussian_kernel1d(sigma, truncate=4.0):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius): exp(-0.5 x^2 / sigma^2), normalised."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum(), radius


def _correlate1d_reflect(a, w, radius, axis):
    """scipy's symmetric correlate1d: centre tap first, then (left + right) * w from the outside in;
    'reflect' boundary (d c b a | a b c d | d c b a)."""
    a = np.moveaxis(a, axis, 0)
    n = a.shape[0]
    idx = np.arange(-radius, n + radius)
    idx = np.where(idx < 0, -idx - 1, idx)
    idx = np.where(idx >= n, 2 * n - 1 - idx, idx)
    # lines shorter than the radius reflect repeatedly
    while (idx < 0).any() or (idx >= n).any():
        idx = np.where(idx < 0, -idx - 1, idx)
        idx = np.where(idx >= n, 2 * n - 1 - idx, idx)
    p = a[idx]
    out = p[radius:radius + n] * w[radius]
    for j in range(radius, 0, -1):
        out = out + (p[radius - j:radius - j + n] + p[radius + j:radius + j + n]) * w[radius - j]
    return np.moveaxis(out, 0, axis)


def smooth(img_u8, sigma):
    """img_as_float64 + gaussian_filter(sigma=[sigma, sigma, 0])."""
    image = img_u8.astype(np.float64) / 255.0
    if sigma <= 0:
        return image
    w, radius = gaussian_kernel1d(sigma)
    return _correlate1d_reflect(_correlate1d_reflect(image, w, radius, 0), w, radius, 1)


@njit(cache=True)
def _find(parent, i):
    while parent[i] != i:
        i = parent[i]
    return i


@njit(cache=True)
def _join(parent, n, m):
    rn, rm = _find(parent, n), _find(parent, m)
    root = rn if rn < rm else rm
    for start in (n, m):
        i = start
        while parent[i] != i:
            nxt = parent[i]
            parent[i] = root
            i = nxt
        parent[i] = root


@njit(cache=True)
def _merge(e0, e1, costs, n, scale, min_size):
    parent = np.arange(n)
    size = np.ones(n, np.int64)
    cint = np.zeros(n, np.float64)
    for e in range(costs.size):
        s0, s1 = _find(parent, e0[e]), _find(parent, e1[e])
        if s0 == s1:
            continue
        if costs[e] < min(cint[s0] + scale / size[s0], cint[s1] + scale / size[s1]):
            _join(parent, s0, s1)
            r = _find(parent, s0)
            size[r] = size[s0] + size[s1]
            cint[r] = costs[e]
    for e in range(costs.size):
        s0, s1 = _find(parent, e0[e]), _find(parent, e1[e])
        if s0 == s1:
            continue
        if size[s0] < min_size or size[s1] < min_size:
            _join(parent, s0, s1)
            r = _find(parent, s0)
            size[r] = size[s0] + size[s1]
    for i in range(n):
        parent[i] = _find(parent, i)
    return parent


def felzenszwalb(img, scale=1.0, sigma=0.8, min_size=20):
    """Label map (int64 HxW, labels 0..S-1) of an HxWx3 u8 image."""
    image = smooth(img, sigma)
    h, w = image.shape[:2]
    sc = float(scale) / 255.0

    def cost(a, b):
        d = a - b
        return np.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2])
    seg = np.arange(h * w).reshape(h, w)
    costs = np.concatenate([cost(image[:, 1:], image[:, :w - 1]).ravel(), cost(image[1:], image[:h - 1]).ravel(),
                            cost(image[1:, 1:], image[:h - 1, :w - 1]).ravel(),
                            cost(image[1:, :w - 1], image[:h - 1, 1:]).ravel()])
    e0 = np.concatenate([seg[:, 1:].ravel(), seg[1:].ravel(), seg[1:, 1:].ravel(), seg[:h - 1, 1:].ravel()])
    e1 = np.concatenate([seg[:, :w - 1].ravel(), seg[:h - 1].ravel(), seg[:h - 1, :w - 1].ravel(), seg[1:, :w - 1].ravel()])
    order = np.argsort(costs, kind="stable")
    root = _merge(e0[order], e1[order], costs[order], h * w, sc, int(min_size))
    return np.unique(root, return_inverse=True)[1].reshape(h, w)
