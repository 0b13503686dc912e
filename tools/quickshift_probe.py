#!/usr/bin/env python
"""Times pcm_quickshift on a synthetic 1080p frame and on a SegTrack2-sized crop (run under ncu to
profile the qs_* kernels: `ncu --set full -k regex:qs_ ... python tools/quickshift_probe.py`)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "non-rigid-object-tracking_b200"))

import numpy as np  # noqa: E402
from pcm import capi  # noqa: E402
from pcm.synthetic import SyntheticSequence  # noqa: E402

seq = SyntheticSequence(1920, 1080, 4, seed=0)
f = seq.frame(1)
h = capi.Handle(0)
h.set_features(8, ["hsv", "lab"])
for rect in [(0, 0, 1920, 1080), (300, 200, 224, 139)]:
    noise = np.random.RandomState(42).normal(scale=1e-5, size=(rect[3], rect[2]))
    h.quickshift(f, rect, noise=noise, want_labels=False)
    t = time.perf_counter()
    reps = 3
    for _ in range(reps):
        _, n = h.quickshift(f, rect, noise=None, want_labels=False)
    print("quickshift %dx%d: %d segments, %.3f ms/call (host crop in, labels stay on the device)"
          % (rect[2], rect[3], n, (time.perf_counter() - t) / reps * 1e3))
h.close()
