"""CPU port of ONE sequence of the reference's sweep -- TEST INFRASTRUCTURE ONLY (see oracle/pcm_oracle.py header).

    python oracle/ref_sequence.py CONFIG.yaml RESULT_FILE [max_frames]

The command line and the result file ("{mean IoU};{seconds}", main.py:366-368) are those of the reference's
`python main.py CONFIG RESULT` that benchmark.py:16-24 launches once per (hyper-parameters, clip) on a ThreadPool.
bench.py's cpu_baseline leg launches THIS script the same way to time the reference's CPU path on the box's host
cores.  Flow of main.py:72-368 without the GUI: open clip, one masker (oracle/ref_port.RefPortMasker) per target,
addModel per polygon selection, per frame tracker box -> update -> computeBenchmark.  Not available in this image
and therefore substituted: skimage over-segmentation -> the restatements oracle/quickshift_oracle.py /
felzenszwalb_oracle.py (parity unpinned, same asymptotic cost: numba loops over the 19 x 19 window / edge sort);
cv.legacy CSRT -> boxes derived from the truth clip (as the product's sweep config does).
A third line of timing detail is appended for the extrapolation bench.py does:
"{seconds of imports + JIT};{seconds of training};{frames processed}".
"""
import os
import sys
import time

T0 = time.time()
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import cv2 as cv  # noqa: E402
import numpy as np  # noqa: E402
import yaml  # noqa: E402

import felzenszwalb_oracle  # noqa: E402
import quickshift_oracle  # noqa: E402
from ref_port import RefPortMasker, SiftFlannPrior, compute_benchmark  # noqa: E402


def read_clip(path):
    cap = cv.VideoCapture(path)
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(f)
    return frames


def truth_boxes(truth_frames, fallback):
    boxes, prev = [], tuple(int(v) for v in fallback)
    for t in truth_frames:
        ys, xs = np.nonzero(t[..., 0] > 127)
        if len(xs):
            prev = (int(xs.min()), int(ys.min()), int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1))
        boxes.append(prev)
    return boxes


def segmenter(name):
    if name == "quickshift":
        return lambda crop: quickshift_oracle.quickshift(np.ascontiguousarray(crop), kernel_size=3, max_dist=6, ratio=0.5, random_seed=42)
    if name == "felzenszwalb":
        return lambda crop: felzenszwalb_oracle.felzenszwalb(np.ascontiguousarray(crop), scale=100, sigma=0.5, min_size=50)
    raise ValueError(name)


def main(argv):
    cfg = yaml.full_load(open(argv[1]))
    max_frames = int(argv[3]) if len(argv) > 3 else None
    frames = read_clip(cfg["input_video"])
    truths = read_clip(cfg["input_truth"])
    pts, frame_numbers, ronis = cfg["pts"], cfg["pts_frame_numbers"], cfg["bboxes_roni"]
    seg = segmenter(cfg["params"]["over_segmentation"])
    seg(frames[0][:40, :40])                                      # numba JIT, like the reference's first call
    t_import = time.time() - T0
    t1 = time.time()
    maskers, first = [], []
    for t, selections in enumerate(pts):
        m = RefPortMasker(debug=False, frame=frames[0], config=cfg, poly_roi=pts[t][0], segment_fn=seg, prior_fn=SiftFlannPrior())
        for s, sel in enumerate(selections):
            if not cfg.get("multi_selection") and s > 0:
                continue
            bbox = cv.boundingRect(np.array(sel))
            if s == 0:
                first.append(bbox)
            m.addModel(frame=frames[frame_numbers[s]], poly_roi=sel, bbox=bbox, bbox_roni=ronis[t][s], n_frame=frame_numbers[s])
        maskers.append(m)
    t_train = time.time() - t1
    boxes = [truth_boxes(truths, b) for b in first]
    n = len(frames) if max_frames is None else min(len(frames), max_frames)
    ious = []
    start = time.time()
    for i in range(n):
        mask = np.zeros_like(frames[i])
        truth = cv.cvtColor(truths[i], cv.COLOR_BGR2GRAY) if i < len(truths) else None
        for t, m in enumerate(maskers):
            m.update(bbox=boxes[t][min(i, len(boxes[t]) - 1)], frame=frames[i], mask=mask, color=(0, 0, 255))
            if truth is not None:
                ious.append(compute_benchmark(mask[:, :, 2], truth))
    seconds = time.time() - start
    with open(argv[2], "w") as f:
        f.write("%s;%s\n%s;%s;%d" % (float(np.mean(ious)) if ious else float("nan"), seconds, t_import, t_train, n))


if __name__ == "__main__":
    main(sys.argv)
