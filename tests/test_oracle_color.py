"""Pins the oracle's colour conversions against OpenCV (the library the reference
calls, maskers/pixel_classification.py:305,307; main.py:285) over ALL 2^24 colours."""
import hashlib

import cv2 as cv
import numpy as np

import pcm_oracle as orc


def all_colours():
    v = np.arange(1 << 24, dtype=np.uint32)
    return np.stack([v & 255, (v >> 8) & 255, (v >> 16) & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)


def test_lab_table_checksums():
    g, cb = orc.lab_tables()
    assert hashlib.sha1(g.astype("<u2").tobytes()).hexdigest() == "a6933c257a23b6991f5c3173f267c82174a95556"
    assert hashlib.sha1(cb.astype("<u2").tobytes()).hexdigest() == "f0f257069ad742b7d40b2afebd71a4038133492e"
    assert list(g[:12]) == [0, 1, 1, 2, 2, 3, 4, 4, 5, 6, 6, 7] and list(g[252:]) == [1986, 2004, 2022, 2040]
    assert list(cb[:6]) == [4520, 4645, 4770, 4895, 5020, 5145] and cb[2040] == 32768


def test_hsv_exhaustive():
    img = all_colours()
    assert np.array_equal(orc.bgr2hsv(img), cv.cvtColor(img, cv.COLOR_BGR2HSV))


def test_lab_exhaustive():
    img = all_colours()
    assert np.array_equal(orc.bgr2lab(img), cv.cvtColor(img, cv.COLOR_BGR2LAB))


def test_gray_exhaustive():
    img = all_colours()
    assert np.array_equal(orc.bgr2gray(img), cv.cvtColor(img, cv.COLOR_BGR2GRAY))


def test_dilate_matches_cv2():
    rng = np.random.default_rng(3)
    for h, w, k in [(40, 50, 7), (9, 9, 3), (30, 17, 4), (5, 64, 7), (1, 1, 7), (12, 12, 2)]:
        m = ((rng.random((h, w)) < 0.05) * 255).astype(np.uint8)
        ref = cv.dilate(m, np.ones((k, k), np.uint8), iterations=1)
        assert np.array_equal(orc.dilate(m, k), ref), (h, w, k)
