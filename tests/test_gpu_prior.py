"""pcm_prior_device (csrc/pcm_prior.cuh): the SIFT-match prior of computePriors (reference :129-163) on the GPU must
equal oracle/prior_oracle.py bit for bit (exact 2-NN matcher; see that file for the one stated difference from the
reference's randomised FLANN search)."""
import cv2 as cv
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import prior_oracle as po  # noqa: E402


def gpu_priors(h, pts1, des1, plane, prev_xy, prev_wh, pts2, des2, seg, S):
    import torch
    dev = torch.device("cuda", 0)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_p1, d_d1 = up(pts1.reshape(-1, 2).astype(np.float32)), up(des1.reshape(-1, 128))
    d_p2, d_d2 = up(pts2.reshape(-1, 2).astype(np.float32)), up(des2.reshape(-1, 128))
    d_plane, d_seg = up(plane), up(seg.astype(np.int32))
    d_out = torch.full((S,), 7.0, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    W = plane.shape[1]
    h.prior_device(d_p1.data_ptr() if len(pts1) else 0, d_d1.data_ptr() if len(pts1) else 0, len(pts1),
                   d_plane.data_ptr() + prev_xy[1] * W + prev_xy[0], W, prev_wh[0], prev_wh[1],
                   d_p2.data_ptr() if len(pts2) else 0, d_d2.data_ptr() if len(pts2) else 0, len(pts2),
                   d_seg.data_ptr(), seg.shape[1], seg.shape[0], S, d_out.data_ptr())
    h.synchronize()
    return d_out.cpu().numpy()


@pytest.fixture(scope="module")
def handle():
    from pcm import capi
    h = capi.Handle(0)
    yield h
    h.close()


@pytest.mark.parametrize("video,box", [("soldier", (140, 0, 260, 224)), ("bmx", (100, 30, 330, 320)), ("worm", (100, 50, 300, 170))])
def test_prior_on_clip_frames_equals_the_oracle(handle, video, box):
    from helpers import read_video
    from pcm.providers import grid_segments, voronoi_segments
    frames, truth = read_video("Video", video), read_video("Truth", video)
    x, y, w, hh = box
    sift = cv.SIFT_create()
    positives = 0
    for i in range(1, 9):
        prev, cur = frames[i - 1][y:y + hh, x:x + w], frames[i][y:y + hh, x:x + w]
        plane = cv.dilate((cv.cvtColor(truth[i - 1], cv.COLOR_BGR2GRAY) > 127).astype(np.uint8) * 255, np.ones((7, 7), np.uint8))
        seg = grid_segments(cur, 8) if i % 2 else voronoi_segments(cur, 150, seed=i)
        S = int(seg.max()) + 1
        k1, d1 = sift.detectAndCompute(np.ascontiguousarray(prev), None)
        k2, d2 = sift.detectAndCompute(np.ascontiguousarray(cur), None)
        p1 = np.array([k.pt for k in k1], np.float32).reshape(-1, 2)
        p2 = np.array([k.pt for k in k2], np.float32).reshape(-1, 2)
        u1, u2 = po.as_u8_descriptors(d1), po.as_u8_descriptors(d2)
        want = po.compute_priors(p1, u1, plane[y:y + hh, x:x + w], p2, u2, seg, S)
        got = gpu_priors(handle, p1, u1, plane, (x, y), (w, hh), p2, u2, seg, S)
        assert np.array_equal(got, want), "frame %d: %d labels differ" % (i, np.count_nonzero(got != want))
        positives += int(np.count_nonzero(want == 1))
    assert positives > 10


@pytest.mark.parametrize("m1", list(range(0, 24)) + [31, 32, 33, 64, 101, 257, 1500])
def test_percentile_filter_and_edge_cases(handle, m1):
    """m1 keypoints with an exact twin in the current frame (every ratio test passes): the 90th-percentile
    displacement filter for every match count, equal displacements, masked-out keypoints, m2 < 2."""
    rng = np.random.default_rng(100 + m1)
    H, W = 90, 120
    m2 = m1 + 40
    des2 = rng.integers(0, 256, (m2, 128)).astype(np.uint8)
    pts2 = np.stack([rng.uniform(0, W - 1, m2), rng.uniform(0, H - 1, m2)], 1).astype(np.float32)
    pick = rng.permutation(m2)[:m1]
    des1 = des2[pick].copy()
    shift = rng.integers(-6, 7, (m1, 2)).astype(np.float32) if m1 % 3 else np.tile(np.float32([3, 4]), (m1, 1))   # all equal
    pts1 = np.clip(pts2[pick] + shift, 0, [W - 1, H - 1]).astype(np.float32)
    plane = (rng.random((H + 20, W + 30)) < 0.8).astype(np.uint8) * 255
    seg = (np.arange(H)[:, None] // 6 * 20 + np.arange(W)[None, :] // 6).astype(np.int32)
    S = int(seg.max()) + 1
    want = po.compute_priors(pts1, des1, plane[7:7 + H, 11:11 + W], pts2, des2, seg, S)
    got = gpu_priors(handle, pts1, des1, plane, (11, 7), (W, H), pts2, des2, seg, S)
    assert np.array_equal(got, want)
    if m1 == 5:                            # fewer than two current keypoints: the reference returns all -1 (:139)
        got = gpu_priors(handle, pts1, des1, plane, (11, 7), (W, H), pts2[:1], des2[:1], seg, S)
        assert np.array_equal(got, np.full(S, -1, np.float32))
