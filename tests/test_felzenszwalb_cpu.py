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


@pytest.mark.parametrize("case", ["lower_half_ties", "few_distinct", "zeros_and_tiny", "wide_range"])
def test_edge_order_is_the_stable_cost_order(case):
    """The product sorts the upper half of the cost bits by radix and orders what ties there by the lower half; the
    partition of an explicit edge list (pcm_felzenszwalb_graph) must be the one a stable argsort of the whole costs
    gives -- on costs that differ ONLY in the lower half, on long runs of equal costs, on exact zeros and on costs
    spread over many binades."""
    from pcm import capi
    rng = np.random.default_rng(5)
    h, w = 30, 41
    seg = np.arange(h * w).reshape(h, w)
    e0 = np.concatenate([seg[:, 1:].ravel(), seg[1:].ravel(), seg[1:, 1:].ravel(), seg[:h - 1, 1:].ravel()])
    e1 = np.concatenate([seg[:, :w - 1].ravel(), seg[:h - 1].ravel(), seg[:h - 1, :w - 1].ravel(), seg[1:, :w - 1].ravel()])
    n = e0.size
    if case == "lower_half_ties":              # one upper half, 32 random lower bits
        costs = (np.float64(0.25).view(np.uint64) + rng.integers(0, 2 ** 32, n, dtype=np.uint64)).view(np.float64)
    elif case == "few_distinct":               # runs of thousands of equal keys: edge order decides
        costs = rng.choice(np.array([0.0, 0.125, 0.1250000001, 0.3, 1.7]), n)
    elif case == "zeros_and_tiny":
        costs = np.where(rng.random(n) < 0.5, 0.0, rng.random(n) * 1e-300)
    else:
        costs = np.exp(rng.uniform(-40, 1, n))
    order = np.argsort(costs, kind="stable")
    for scale, min_size in ((0.4, 1), (0.05, 12)):
        root = fo._merge(e0[order], e1[order], costs[order], h * w, scale, min_size)
        want = np.unique(root, return_inverse=True)[1]
        got, count = capi.felzenszwalb_graph(h * w, e0, e1, costs, scale, min_size)
        assert count == int(want.max()) + 1
        assert np.array_equal(got, want)


@pytest.mark.parametrize("rect", [(4, 3, 1, 20), (2, 5, 30, 1), (7, 7, 1, 1), (0, 0, 2, 2)])
def test_felzenszwalb_degenerate_crops(rect):
    """One-pixel-wide, one-pixel-high and single-pixel crops: edge classes that do not exist there are simply absent."""
    frame = np.random.default_rng(23).integers(0, 256, (32, 40, 3), dtype=np.uint8)
    _check(frame, rect, scale=100, sigma=0.5, min_size=3)


def test_felzenszwalb_from_many_threads():
    """The sweep computes the label maps of a clip's frames on a pool of threads; the library keeps its working memory
    per thread: crops of different sizes, interleaved on 8 threads, give the single-threaded maps."""
    from concurrent.futures import ThreadPoolExecutor
    from pcm import capi
    rng = np.random.default_rng(31)
    frame = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    frame[30:90, 40:120] //= 4                                  # some structure besides the noise
    rects = [(int(x), int(y), int(w), int(h)) for x, y, w, h in
             zip(rng.integers(0, 60, 24), rng.integers(0, 40, 24), rng.integers(5, 100, 24), rng.integers(5, 80, 24))]
    want = [capi.felzenszwalb(frame, r, scale=100, sigma=0.5, min_size=20) for r in rects]
    with ThreadPoolExecutor(max_workers=8) as pool:
        got = list(pool.map(lambda r: capi.felzenszwalb(frame, r, scale=100, sigma=0.5, min_size=20), rects * 3))
    for k, (seg, n) in enumerate(got):
        assert n == want[k % len(rects)][1] and np.array_equal(seg, want[k % len(rects)][0])
