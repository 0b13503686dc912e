"""pcm_felzenszwalb (host code of the library) against the restatement of
scikit-image 0.17.2's algorithm in oracle/felzenszwalb_oracle.py -- PARITY UNPINNED against
scikit-image itself; the Gaussian step of the oracle is pinned against scipy.ndimage.  Runs
without a GPU."""
import numpy as np
import pytest
import scipy.ndimage as ndi

import felzenszwalb_oracle as fo
from helpers import read_video


@pytest.mark.parametrize("sigma", [0.5, 0.8, 2.0])
def test_oracle_gaussian_step_equals_scipy(sigma):
    rng = np.random.default_rng(int(sigma * 10))
    for shape in [(37, 53, 3), (2, 3, 3), (1, 9, 3), (64, 5, 3)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        want = ndi.gaussian_filter(img.astype(np.float64) / 255.0, sigma=[sigma, sigma, 0])
        assert np.array_equal(fo.smooth(img, sigma), want), shape


def _check(frame, rect, **kw):
    from pcm import capi
    x, y, w, h = rect
    want = fo.felzenszwalb(frame[y:y + h, x:x + w], kw.get("scale", 100), kw.get("sigma", 0.5), kw.get("min_size", 50))
    got, n = capi.felzenszwalb(frame, rect, **kw)
    assert n == int(want.max()) + 1
    assert np.array_equal(got, want)
    return n


def test_felzenszwalb_segtrack_crops():
    assert _check(read_video("Video", "soldier")[0], (300, 0, 139, 224)) > 20
    _check(read_video("Video", "frog")[3], (150, 60, 133, 159))
    _check(read_video("Video", "parachute")[5], (0, 0, 414, 352))


@pytest.mark.parametrize("case", ["noise", "flat", "gradient", "tiny", "other_params"])
def test_felzenszwalb_synthetic(case):
    rng = np.random.default_rng(17)
    kw = {}
    if case == "noise":
        frame, rect = rng.integers(0, 256, (70, 90, 3), dtype=np.uint8), (3, 5, 81, 60)
    elif case == "flat":                       # every cost is an exact zero: one segment
        frame, rect = np.full((40, 50, 3), 99, np.uint8), (0, 0, 50, 40)
    elif case == "gradient":
        g = np.add.outer(np.arange(96), np.arange(120)).astype(np.uint8)
        frame, rect = np.stack([g, g[::-1], 255 - g], -1).copy(), (10, 7, 100, 80)
    elif case == "tiny":
        frame, rect = rng.integers(0, 256, (6, 7, 3), dtype=np.uint8), (1, 2, 3, 2)
    else:
        frame, rect = rng.integers(0, 64, (60, 60, 3), dtype=np.uint8), (0, 0, 60, 60)
        kw = dict(scale=1, sigma=0.8, min_size=20)       # scikit-image's own defaults
    n = _check(frame, rect, **kw)
    if case == "flat":
        assert n == 1
