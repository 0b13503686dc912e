"""CPU restatement of `skimage.segmentation.quickshift` as the PC masker calls it
(reference maskers/pixel_classification.py:70-71:
 `quickshift(crop_frame, kernel_size=3, max_dist=6, ratio=0.5, random_seed=42)`).

TEST INFRASTRUCTURE ONLY (see oracle/pcm_oracle.py header).

PARITY UNPINNED.  scikit-image is a dependency of the reference (environment.yaml:12 pins
scikit-image 0.17.2) whose source is not under /root/reference and which is not installed in
this image, and the reference has no test or golden vector for its output.  This file restates
the published algorithm of that version:

  skimage/segmentation/_quickshift.py        quickshift(): img_as_float -> rgb2lab -> (sigma = 0:
                                             no smoothing) -> image * ratio -> _quickshift_cython
  skimage/segmentation/_quickshift_cy.pyx    densities over a (2w+1)^2 window, w = ceil(3 *
                                             kernel_size), Gaussian of the 5-D distance; tiny
                                             seeded normal noise to break ties; parent = nearest
                                             window pixel of higher density; links longer than
                                             max_dist are cut; labels = np.unique(root, return_inverse)
  skimage/color/colorconv.py                 rgb2xyz (sRGB companding, xyz_from_rgb), xyz2lab
                                             (D65, 2 degree observer)

What is NOT claimed: bit-equality of the float64 Lab values with skimage's BLAS matrix product
(the order of the three products is fixed here as ((r*m0 + g*m1) + b*m2)), nor of libm's
exp / cbrt / pow against the CUDA math library.  The GPU path is checked against THIS
restatement (tests/test_gpu_quickshift.py): label maps must describe the same partition.
"""
import math

import numpy as np
from numba import njit

XYZ_FROM_RGB = np.array([[0.412453, 0.357580, 0.180423],
                         [0.212671, 0.715160, 0.072169],
                         [0.019334, 0.119193, 0.950227]])
XYZ_REF_WHITE_D65_2 = np.array([0.95047, 1.0, 1.08883])


def srgb_linear_table():
    """lin[v] for v in 0..255: skimage rgb2xyz companding of v/255 (float64)."""
    a = np.arange(256, dtype=np.float64) / 255.0
    out = np.empty(256, np.float64)
    m = a > 0.04045
    out[m] = np.power((a[m] + 0.055) / 1.055, 2.4)
    out[~m] = a[~m] / 12.92
    return out


def rgb2lab_u8(img):
    """HxWx3 u8, channels in the order given (the reference passes an OpenCV BGR crop, so
    channel 0 plays the role of 'R'), -> HxWx3 float64 Lab."""
    lin = srgb_linear_table()
    c0, c1, c2 = lin[img[..., 0]], lin[img[..., 1]], lin[img[..., 2]]
    M = XYZ_FROM_RGB
    xyz = [(c0 * M[i, 0] + c1 * M[i, 1]) + c2 * M[i, 2] for i in range(3)]
    f = []
    for i in range(3):
        t = xyz[i] / XYZ_REF_WHITE_D65_2[i]
        f.append(np.where(t > 0.008856, np.cbrt(t), 7.787 * t + 16.0 / 116.0))
    L = 116.0 * f[1] - 16.0
    a = 500.0 * (f[0] - f[1])
    b = 200.0 * (f[1] - f[2])
    return np.stack([L, a, b], axis=-1)


@njit(cache=True)
def _densities(image, kernel_size, kernel_width):
    h, w, nc = image.shape
    inv = -0.5 / (kernel_size * kernel_size)
    dens = np.zeros((h, w), np.float64)
    for r in range(h):
        r_min, r_max = max(r - kernel_width, 0), min(r + kernel_width + 1, h)
        for c in range(w):
            c_min, c_max = max(c - kernel_width, 0), min(c + kernel_width + 1, w)
            acc = 0.0
            for r_ in range(r_min, r_max):
                for c_ in range(c_min, c_max):
                    dist = 0.0
                    for ch in range(nc):
                        t = image[r, c, ch] - image[r_, c_, ch]
                        dist += t * t
                    t = float(r - r_)
                    dist += t * t
                    t = float(c - c_)
                    dist += t * t
                    acc += math.exp(dist * inv)
            dens[r, c] = acc
    return dens


@njit(cache=True)
def _parents(image, dens, kernel_width):
    h, w, nc = image.shape
    parent = np.empty(h * w, np.int64)
    dist_parent = np.empty(h * w, np.float64)
    for r in range(h):
        r_min, r_max = max(r - kernel_width, 0), min(r + kernel_width + 1, h)
        for c in range(w):
            c_min, c_max = max(c - kernel_width, 0), min(c + kernel_width + 1, w)
            cur = dens[r, c]
            closest = np.inf
            best = r * w + c
            for r_ in range(r_min, r_max):
                for c_ in range(c_min, c_max):
                    if dens[r_, c_] > cur:
                        dist = 0.0
                        for ch in range(nc):
                            t = image[r, c, ch] - image[r_, c_, ch]
                            dist += t * t
                        t = float(r - r_)
                        dist += t * t
                        t = float(c - c_)
                        dist += t * t
                        if dist < closest:
                            closest = dist
                            best = r_ * w + c_
            parent[r * w + c] = best
            dist_parent[r * w + c] = math.sqrt(closest)
    return parent, dist_parent


def tie_noise(shape, random_seed=42):
    """`random_state.normal(scale=0.00001, size=(height, width))` of _quickshift_cy.pyx."""
    return np.random.RandomState(random_seed).normal(scale=0.00001, size=shape)


def quickshift(img, ratio=1.0, kernel_size=5, max_dist=10, random_seed=42, stages=False):
    """Label map (int64, HxW, labels 0..S-1) of an HxWx3 u8 image."""
    image = np.ascontiguousarray(rgb2lab_u8(img) * ratio)
    kernel_width = int(math.ceil(3 * kernel_size))
    dens = _densities(image, float(kernel_size), kernel_width)
    dens = dens + tie_noise(dens.shape, random_seed)
    parent, dist_parent = _parents(image, dens, kernel_width)
    far = dist_parent > max_dist
    parent[far] = np.arange(parent.size)[far]
    old = np.zeros_like(parent)
    while (old != parent).any():
        old = parent
        parent = parent[parent]
    labels = np.unique(parent, return_inverse=True)[1].reshape(img.shape[:2])
    if stages:
        return labels, dict(lab=image, densities=dens, root=parent.reshape(img.shape[:2]))
    return labels


def same_partition(a, b):
    """True when two label maps induce the same partition with the same label ORDER
    (both come from np.unique ranks of root indices, so equal partitions give equal labels)."""
    return a.shape == b.shape and np.array_equal(a, b)


def partition_agreement(a, b):
    """Fraction of pixels on which the partitions agree (pairs (a, b) that are the majority
    image of their a-label)."""
    a, b = a.reshape(-1).astype(np.int64), b.reshape(-1).astype(np.int64)
    key = a * (int(b.max()) + 1) + b
    uniq, cnt = np.unique(key, return_counts=True)
    best = {}
    for k, n in zip(uniq, cnt):
        la = int(k) // (int(b.max()) + 1)
        best[la] = max(best.get(la, 0), int(n))
    return sum(best.values()) / a.size
