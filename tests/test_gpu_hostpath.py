"""Host-buffer entry points (pcm_update / pcm_iou through the plugin): page-locked caller buffers
(pcm_host_register), the mask mirror that spares pcm_iou the re-upload of the mask pcm_update has just
produced, and label maps that stay resident between updates.  Every shortcut must give exactly the
bytes and counts of the plain staged path and of the oracle.  Needs a GPU: `pytest -m gpu`."""
import numpy as np
import pytest

import pcm_oracle as orc
from test_gpu_parity import _random_forest_arrays

pytestmark = pytest.mark.gpu


def _setup(rng, hgt, wid):
    from pcm import capi
    F = 27
    trees = _random_forest_arrays(rng, 8, 4, F)
    h = capi.Handle(0)
    h.set_features(1, ["rgb"])
    h.add_model_arrays(0, trees)
    return h, orc.forest_from_arrays(trees, F)


def _want_mask(forest, frame, rect, seg, k=3):
    x, y, w, hh = rect
    crop = frame[y:y + hh, x:x + w]
    p1 = orc.forest_p1(forest, orc.get_features_int([crop], 1))
    S = int(seg.max()) + 1
    scores, _ = orc.saliency_scores(p1, np.zeros(seg.size), seg, 0.0, np.full(S, -1, np.float32), 0.0)
    return orc.dilate(orc.saliency_mask(scores, seg), k)


@pytest.mark.parametrize("hgt,wid", [(540, 960), (1000, 1100)])      # the larger frame travels in two row bands
def test_registered_buffers_and_mask_mirror(hgt, wid):
    from pcm import capi
    from pcm.providers import grid_segments
    rng = np.random.default_rng(5)
    h, forest = _setup(rng, hgt, wid)
    prm = capi.Handle.make_params(0, dilation_kernel=3)
    frames = [rng.integers(0, 256, (hgt, wid, 3), dtype=np.uint8) for _ in range(3)]
    truths = [(rng.random((hgt, wid)) < 0.4).astype(np.uint8) * 255 for _ in range(3)]
    rect = (0, 0, wid, hgt)
    seg = grid_segments(frames[0], 12)
    mask = np.zeros((hgt, wid, 3), np.uint8)

    def step(i, registered):
        b0 = h.transfer_bytes
        h.update(frames[i], rect, seg, 0, None, prm, mask)
        b1 = h.transfer_bytes
        counts = h.iou_counts(mask[:, :, 2], truths[i])
        b2 = h.transfer_bytes
        want = _want_mask(forest, frames[i], rect, seg)
        assert np.array_equal(mask[:, :, 2], want), (i, registered)
        assert not mask[:, :, :2].any()
        assert counts == orc.iou_counts(want, truths[i]), (i, registered)
        return b1[0] - b0[0], b2[0] - b1[0]

    step(0, False)                                        # first call: everything travels, the mirror gets synchronised
    up, iou = step(1, False)
    assert up == hgt * wid * 3                            # labels cached, frame staged
    assert iou == hgt * wid                               # truth only: the mask is already on the device
    for a in frames + truths:
        capi.host_register(a)
    try:
        for i in range(3):
            up, iou = step(i, True)
            assert up == hgt * wid * 3 and iou == hgt * wid
        # the caller edits the mask between update and iou: the changed band (and only it) is re-sent
        h.update(frames[0], rect, seg, 0, None, prm, mask)
        mask[100:103, 50:300, 2] = 255
        b0 = h.transfer_bytes
        counts = h.iou_counts(mask[:, :, 2], truths[0])
        sent = h.transfer_bytes[0] - b0[0] - hgt * wid
        assert counts == orc.iou_counts(mask[:, :, 2], truths[0])
        assert 0 < sent <= 2 * (1 << 16) + 2 * wid, sent
        # a different mask image with other content: whatever differs is refreshed
        other = np.zeros((hgt, wid, 3), np.uint8)
        other[:, :, 2] = (rng.random((hgt, wid)) < 0.5) * 255
        assert h.iou_counts(other[:, :, 2], truths[1]) == orc.iou_counts(other[:, :, 2], truths[1])
        assert h.iou_counts(mask[:, :, 2], truths[2]) == orc.iou_counts(mask[:, :, 2], truths[2])
        # a dense (pixel stride 1) mask and a BGR truth
        dense = np.ascontiguousarray(mask[:, :, 2])
        bgr = np.repeat(truths[0][:, :, None], 3, axis=2)
        assert h.iou_counts(dense, bgr) == orc.iou_counts(dense, truths[0])
    finally:
        for a in frames + truths:
            capi.host_unregister(a)
    step(2, False)
    h.close()


def test_registered_sub_crop_in_row_bands():
    """A crop that is not the whole frame, large enough to be fed in several row bands (2-D copies from the page-locked
    frame on the copy stream, K0 / K1 per band): same mask as the staged path and the oracle."""
    from pcm import capi
    from pcm.providers import grid_segments
    rng = np.random.default_rng(8)
    hgt, wid = 1300, 1500
    h, forest = _setup(rng, hgt, wid)
    prm = capi.Handle.make_params(0, dilation_kernel=5)
    frame = rng.integers(0, 256, (hgt, wid, 3), dtype=np.uint8)
    rect = (37, 21, 1401, 1233)                          # 1.7 M px: three bands; odd width
    x, y, w, hh = rect
    seg = grid_segments(frame[y:y + hh, x:x + w], 14)
    want = _want_mask(forest, frame, rect, seg, 5)
    staged = np.zeros((hgt, wid, 3), np.uint8)
    h.update(frame, rect, seg, 0, None, prm, staged)
    assert np.array_equal(staged[y:y + hh, x:x + w, 2], want)
    capi.host_register(frame)
    try:
        for _ in range(3):
            direct = np.zeros((hgt, wid, 3), np.uint8)
            h.update(frame, rect, seg, 0, None, prm, direct)
            assert np.array_equal(direct, staged)
    finally:
        capi.host_unregister(frame)
    h.close()


def test_two_targets_share_one_mask_image():
    """main.py:286-343 with two targets: each masker writes its own crop of the shared mask image and scores the WHOLE
    image; a handle's mirror has never seen the other handle's crop and must pick it up from the caller's bytes."""
    from pcm import capi
    from pcm.providers import grid_segments
    rng = np.random.default_rng(6)
    hgt, wid = 400, 700
    frame = rng.integers(0, 256, (hgt, wid, 3), dtype=np.uint8)
    truth = (rng.random((hgt, wid)) < 0.3).astype(np.uint8) * 255
    hs = [_setup(rng, hgt, wid) for _ in range(2)]
    rects = [(20, 30, 300, 250), (350, 100, 320, 280)]
    prm = capi.Handle.make_params(0, dilation_kernel=3)
    for rep in range(3):
        mask = np.zeros((hgt, wid, 3), np.uint8)
        want = np.zeros((hgt, wid), np.uint8)
        for (h, forest), rect in zip(hs, rects):
            x, y, w, hh = rect
            seg = grid_segments(frame[y:y + hh, x:x + w], 9 + rep)
            h.update(frame, rect, seg, 0, None, prm, mask)
            want[y:y + hh, x:x + w] = _want_mask(forest, frame, rect, seg)
            assert np.array_equal(mask[:, :, 2], want)
            assert h.iou_counts(mask[:, :, 2], truth) == orc.iou_counts(want, truth)
        frame = np.roll(frame, 7, axis=1)
    for h, _ in hs:
        h.close()


def test_resident_label_map_through_the_plugin():
    """A provider that returns the same read-only array lets the plugin skip the label upload (pcm_update with
    labels == NULL); a writeable array, another object or another crop size sends the map again."""
    from maskers import getMaskerByName
    from pcm.providers import grid_segments
    rng = np.random.default_rng(9)
    hgt, wid = 300, 420
    frames = [rng.integers(0, 256, (hgt, wid, 3), dtype=np.uint8) for _ in range(4)]
    cache = {}

    def provider(crop):
        key = crop.shape[:2]
        if key not in cache:
            cache[key] = grid_segments(crop, 10)
            cache[key].setflags(write=provider.read_only)
        return cache[key]
    provider.read_only = True

    def run(masker, box):
        sent = []
        masks = []
        for f in frames:
            mask = np.zeros_like(f)
            b0 = masker.native.transfer_bytes[0]
            masker.update(bbox=box, frame=f, mask=mask)
            sent.append(masker.native.transfer_bytes[0] - b0)
            masks.append(mask[:, :, 2].copy())
        return sent, masks

    cfg = dict(multi_selection=False, params=dict(n_estimators=6, max_depth=4, n_components=1, novelty_detection=False,
                                                  over_segmentation="grid:10", features="1 rgb", dilation_kernel=3, prior_weight=0.0))
    m = getMaskerByName("PC", debug=False, frame=frames[0], config=cfg, poly_roi=None, update_mask=False, segment_fn=provider,
                        device=0)
    trees = _random_forest_arrays(rng, 6, 4, m.native.num_features)
    m.native.add_model_arrays(0, trees)
    m.models.append({"n_frame": 0, "model": None, "n_trees": 6})
    m.novelty_det.append({"n_frame": 0, "model": None, "threshold": 0.0})
    box = (40, 30, 200, 150)
    from pcm import capi
    x, y, w, h = capi.crop_rect(box, hgt, wid)
    sent, masks = run(m, box)
    assert sent[0] == w * h * 3 + w * h * 4 and all(s == w * h * 3 for s in sent[1:]), sent
    # same frames, labels forced to travel every time: identical masks
    m.reuse_resident_labels = False
    m.native.set_label_cache(False)
    sent2, masks2 = run(m, box)
    assert all(s == w * h * 3 + w * h * 4 for s in sent2), sent2
    assert all(np.array_equal(a, b) for a, b in zip(masks, masks2))
    m.reuse_resident_labels = True
    m.native.set_label_cache(True)
    # a writeable array is never trusted (here the label cache still spares the upload, after comparing the bytes)
    cache.clear()
    provider.read_only = False
    sent3, masks3 = run(m, box)
    assert all(np.array_equal(a, b) for a, b in zip(masks, masks3))
    # another crop size: a new map travels once
    cache.clear()
    provider.read_only = True
    box2 = (100, 60, 150, 120)
    x2, y2, w2, h2 = capi.crop_rect(box2, hgt, wid)
    sent4, _ = run(m, box2)
    assert sent4[0] == w2 * h2 * 7 and all(s == w2 * h2 * 3 for s in sent4[1:]), sent4
    m.close()


def test_labels_null_needs_a_resident_map():
    from pcm import capi
    rng = np.random.default_rng(10)
    h, _ = _setup(rng, 64, 64)
    frame = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    mask = np.zeros_like(frame)
    with pytest.raises(capi.PcmError):
        h.update(frame, (0, 0, 64, 64), None, 0, None, capi.Handle.make_params(0, dilation_kernel=3), mask)
    h.close()
