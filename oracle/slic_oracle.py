"""CPU restatement of `skimage.segmentation.slic` as the PC masker calls it
(reference maskers/pixel_classification.py:74-75:
 `slic(crop_frame, n_segments=250, compactness=10, sigma=1, start_label=0)`).

TEST INFRASTRUCTURE ONLY (see oracle/pcm_oracle.py header).

PARITY UNPINNED.  scikit-image (environment.yaml:12 pins 0.17.2) is not under /root/reference and not installed
here, and the reference holds no test or golden vector for this stage.  Restated from the published algorithm of
that version:

  skimage/segmentation/slic_superpixels.py   slic(): img_as_float -> [depth 1] -> gaussian_filter(sigma = [s, s, s, 0])
                                             -> rgb2lab -> regular_grid seeds -> image / compactness ->
                                             _slic_cython(max_iter = 10) -> _enforce_label_connectivity_cython
                                             (min_size_factor 0.5, max_size_factor 3)
  skimage/util/_regular_grid.py              regular_grid(shape, n_points)
  skimage/segmentation/_slic.pyx             k-means in (z, y, x, L, a, b): every cluster looks at the window
                                             +-2 steps around its centre; distance = squared position distance /
                                             step^2 + squared colour distance; a pixel takes the FIRST cluster (in
                                             cluster order) with the smallest distance; centres = means of their
                                             pixels; the "no pixel changed" exit never fires because the distance
                                             map is reset every iteration, so all max_iter iterations run;
                                             connectivity: raster scan, breadth-first component up to max_size,
                                             components smaller than min_size take the label of a neighbouring,
                                             already relabelled component met during the search (0 if none)

The Gaussian step reuses the scipy-exact correlate of oracle/felzenszwalb_oracle.py (pinned bit for bit against
scipy.ndimage there); the depth axis has length 1 and is still filtered, as scipy does (every reflected tap is the
pixel itself).  rgb2lab is the one of oracle/quickshift_oracle.py applied to float images.
"""
import math

import numpy as np
from numba import njit

from felzenszwalb_oracle import _correlate1d_reflect, gaussian_kernel1d
from quickshift_oracle import XYZ_FROM_RGB, XYZ_REF_WHITE_D65_2


def rgb2lab_float(img):
    """HxWx3 float64 in 0..1 (channel 0 plays 'R': the reference passes BGR) -> Lab, skimage.color formulas."""
    m = img > 0.04045
    lin = np.where(m, np.power((img + 0.055) / 1.055, 2.4), img / 12.92)
    c0, c1, c2 = lin[..., 0], lin[..., 1], lin[..., 2]
    M = XYZ_FROM_RGB
    xyz = [(c0 * M[i, 0] + c1 * M[i, 1]) + c2 * M[i, 2] for i in range(3)]
    f = []
    for i in range(3):
        t = xyz[i] / XYZ_REF_WHITE_D65_2[i]
        f.append(np.where(t > 0.008856, np.cbrt(t), 7.787 * t + 16.0 / 116.0))
    return np.stack([116.0 * f[1] - 16.0, 500.0 * (f[0] - f[1]), 200.0 * (f[1] - f[2])], axis=-1)


def smooth(img_u8, sigma):
    """img_as_float + gaussian_filter over (depth = 1, H, W, C) with sigma = [s, s, s, 0]: axes 0, 1, 2 in that order."""
    image = img_u8.astype(np.float64) / 255.0
    if sigma <= 0:
        return image
    w, radius = gaussian_kernel1d(sigma)
    vol = image[np.newaxis]
    for axis in (0, 1, 2):
        vol = _correlate1d_reflect(vol, w, radius, axis)
    return vol[0]


def grid_steps(h, w, n_segments):
    """(start_y, step_y, start_x, step_x) of regular_grid((1, h, w), n_segments); None when every pixel is a seed."""
    shape = np.array([1, h, w])
    unsort = np.argsort(np.argsort(shape))
    dims = np.sort(shape)
    space = float(np.prod(shape))
    if space <= n_segments:
        return None
    steps = np.full(3, (space / n_segments) ** (1.0 / 3), dtype=np.float64)
    if (dims < steps).any():
        for d in range(3):
            steps[d] = dims[d]
            space = float(np.prod(dims[d + 1:]))
            steps[d + 1:] = (space / n_segments) ** (1.0 / (3 - d - 1))
            if (dims >= steps).all():
                break
    starts = (steps // 2).astype(int)
    isteps = np.round(steps).astype(int)
    starts, isteps = starts[unsort], isteps[unsort]
    return int(starts[1]), int(isteps[1]), int(starts[2]), int(isteps[2])


@njit(cache=True)
def _kmeans(image, centers, step, step_y, step_x, max_iter):
    h, w, nc = image.shape
    K = centers.shape[0]
    nearest = np.zeros((h, w), np.int64)
    dist = np.empty((h, w), np.float64)
    count = np.zeros(K, np.int64)
    spatial_weight = 1.0 / (step * step)
    for _ in range(max_iter):
        dist[:, :] = np.inf                       # DBL_MAX in the original; only comparisons matter
        for k in range(K):
            cy, cx = centers[k, 1], centers[k, 2]
            if not (cy == cy):                    # a cluster that lost all its pixels (NaN centre) takes no part
                continue
            y_min, y_max = int(max(cy - 2 * step_y, 0.0)), int(min(cy + 2 * step_y + 1, float(h)))
            x_min, x_max = int(max(cx - 2 * step_x, 0.0)), int(min(cx + 2 * step_x + 1, float(w)))
            for y in range(y_min, y_max):
                dy = (cy - y) ** 2
                for x in range(x_min, x_max):
                    d = (0.0 + dy + (cx - x) ** 2) * spatial_weight
                    dc = 0.0
                    for c in range(nc):
                        t = image[y, x, c] - centers[k, 3 + c]
                        dc += t * t
                    d += dc
                    if dist[y, x] > d:
                        nearest[y, x] = k
                        dist[y, x] = d
        count[:] = 0
        centers[:, :] = 0.0
        for y in range(h):
            for x in range(w):
                k = nearest[y, x]
                count[k] += 1
                centers[k, 1] += y
                centers[k, 2] += x
                for c in range(nc):
                    centers[k, 3 + c] += image[y, x, c]
        for k in range(K):
            for c in range(3 + nc):
                centers[k, c] = centers[k, c] / count[k] if count[k] > 0 else np.nan
    return nearest


@njit(cache=True)
def _enforce_connectivity(seg, min_size, max_size, start_label):
    h, w = seg.shape
    out = -np.ones((h, w), np.int64)
    coords = np.empty((max(max_size, 1), 2), np.int64)
    dy = (0, 0, 1, -1)
    dx = (1, -1, 0, 0)
    new_label = start_label
    for y in range(h):
        for x in range(w):
            if out[y, x] >= 0:
                continue
            adjacent = 0
            label = seg[y, x]
            out[y, x] = new_label
            size = 1
            visited = 0
            coords[0, 0] = y
            coords[0, 1] = x
            while visited < size and size < max_size:
                for i in range(4):
                    yy = coords[visited, 0] + dy[i]
                    xx = coords[visited, 1] + dx[i]
                    if 0 <= xx < w and 0 <= yy < h:
                        if seg[yy, xx] == label and out[yy, xx] == -1:
                            out[yy, xx] = new_label
                            coords[size, 0] = yy
                            coords[size, 1] = xx
                            size += 1
                            if size >= max_size:
                                break
                        elif out[yy, xx] >= 0 and out[yy, xx] != new_label:
                            adjacent = out[yy, xx]
                visited += 1
            if size < min_size:
                for i in range(size):
                    out[coords[i, 0], coords[i, 1]] = adjacent
            else:
                new_label += 1
    return out


def slic(img, n_segments=100, compactness=10.0, sigma=0.0, max_iter=10, start_label=0, stages=False):
    """Label map (int64, HxW) of an HxWx3 u8 image (channel 0 treated as 'R', like the reference's BGR crops)."""
    h, w = img.shape[:2]
    lab = rgb2lab_float(smooth(img, float(sigma)))
    g = grid_steps(h, w, n_segments)
    if g is None:
        ys, xs, step_y, step_x = np.arange(h), np.arange(w), 1, 1
    else:
        sy, step_y, sx, step_x = g
        ys, xs = np.arange(sy, h, step_y), np.arange(sx, w, step_x)
    K = len(ys) * len(xs)
    centers = np.zeros((K, 6), np.float64)
    centers[:, 1] = np.repeat(ys, len(xs))
    centers[:, 2] = np.tile(xs, len(ys))
    step = float(max(1, step_y, step_x))
    image = np.ascontiguousarray(lab * (1.0 / compactness))
    nearest = _kmeans(image, centers, step, step_y, step_x, int(max_iter))
    segment_size = h * w / n_segments
    labels = _enforce_connectivity(nearest, int(0.5 * segment_size), int(3 * segment_size), int(start_label))
    if stages:
        return labels, dict(lab=lab, nearest=nearest, centers=centers)
    return labels
