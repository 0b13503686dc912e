"""Headless sequence driver with the flow of the reference's main.py (:72-368):

    config (YAML dict) -> open video -> one masker per target (getMaskerByName) ->
    addModel for every polygon selection -> rigid tracker on the first box ->
    per frame: tracker boxes -> masker.update(bbox, frame, mask, color) per target ->
    computeBenchmark(mask[:, :, 2], truth) per box -> "{mean IoU};{seconds}".

What differs from the reference is only what cannot run on a headless GPU box: the GUI
branches (manual ROI selection, imshow / waitKey, video writers) are not built, and the
rigid tracker is a provider (`pcm.providers`): `cv.legacy` CSRT/KCF when the OpenCV build has
it, else OpenCV's MIL tracker; boxes derived from the ground-truth clip only on request
(`tracker_provider: truth`, which the shipped sweep config asks for and reports).
The per-frame arithmetic is the CUDA path behind `maskers.getMaskerByName("PC")`; the IoU is
`pcm_iou` (benchmark.py:8-14).
"""
import os
import time

import cv2 as cv
import numpy as np
import yaml

from maskers import getMaskerByName
from . import providers, stages

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# fixed overlay colours (the reference draws a random CSS3 colour per target, colorutils.py:4-29;
# the PC masker ignores `color`)
COLORS = [(0, 0, 255), (0, 255, 0), (255, 0, 0), (0, 255, 255), (255, 0, 255), (255, 255, 0)]


def load_config(path):
    """YAML config with the reference's keys (config.yaml:1-37).  `polygons: <clip>` pulls pts /
    pts_frame_numbers / bboxes_roni of that clip from polygons.yaml next to the config (or the
    package's) unless they are given inline."""
    with open(path) as f:
        cfg = yaml.full_load(f)
    name = cfg.get("polygons")
    if name and cfg.get("pts") is None:
        for d in (os.path.dirname(os.path.abspath(path)), PKG):
            pp = os.path.join(d, "polygons.yaml")
            if os.path.isfile(pp):
                with open(pp) as f:
                    cfg.update(yaml.full_load(f)[name])
                break
    return cfg


def resolve_path(p):
    """Paths in the shipped configs are relative to the package directory (like the reference's
    are to its repository root)."""
    if p is None or os.path.isabs(p) or os.path.exists(p):
        return p
    q = os.path.join(PKG, p)
    return q if os.path.exists(q) else p


_clip_cache = {}
_truth_box_cache = {}


def read_clip(path, resize_factor=1, cache=True):
    """All frames of a clip, resized like main.py:99,284 (cv.resize(fx=fy=resize_factor))."""
    key = (os.path.abspath(path), float(resize_factor))
    if cache and key in _clip_cache:
        return _clip_cache[key]
    cap = cv.VideoCapture(path)
    if not cap.isOpened():
        raise IOError("Input video not opened correctly: %s" % path)
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        if resize_factor != 1:
            f = cv.resize(f, (0, 0), fx=resize_factor, fy=resize_factor)
        frames.append(f)
    cap.release()
    if cache:
        _clip_cache[key] = frames
    return frames


def int_box(b):
    """main.py:294-298: corner points truncated to int, then width / height from the corners."""
    x0, y0 = int(b[0]), int(b[1])
    x1, y1 = int(b[0] + b[2]), int(b[1] + b[3])
    return (x0, y0, x1 - x0, y1 - y0)


class _CvMultiTracker:
    """cv.legacy.MultiTracker when available (main.py:11-32,258-261), else one cv.TrackerMIL per target."""

    def __init__(self, name, frame, boxes):
        self.legacy = hasattr(cv, "legacy") and hasattr(cv.legacy, "MultiTracker_create")
        self.name = name
        if self.legacy:
            self.mt = cv.legacy.MultiTracker_create()
            for b in boxes:
                self.mt.add(self._make(name), frame, tuple(b))
        else:
            self.trackers = []
            for b in boxes:
                t = cv.TrackerMIL_create()
                t.init(frame, tuple(int(v) for v in b))
                self.trackers.append(t)
            self.last = [tuple(b) for b in boxes]

    @staticmethod
    def _make(name):
        table = {"BOOSTING": "TrackerBoosting_create", "MIL": "TrackerMIL_create", "KCF": "TrackerKCF_create",
                 "TLD": "TrackerTLD_create", "MEDIANFLOW": "TrackerMedianFlow_create", "MOSSE": "TrackerMOSSE_create",
                 "CSRT": "TrackerCSRT_create"}
        return getattr(cv.legacy, table.get(name, "TrackerCSRT_create"))()

    def update(self, frame):
        if self.legacy:
            return self.mt.update(frame)
        out = []
        for i, t in enumerate(self.trackers):
            ok, b = t.update(frame)
            if ok:
                self.last[i] = tuple(b)
            out.append(self.last[i])
        return True, out


def make_tracker(config, frame0, first_boxes, truth_frames, provider=None):
    """-> (factory(frame, boxes) -> tracker with .update(frame) -> (ok, boxes), name)."""
    provider = provider or config.get("tracker_provider") or "auto"
    if provider == "auto":
        # never the ground truth by default: boxes derived from the clip that also scores the masks are an
        # explicit opt-in (`tracker_provider: truth`).  "cv" = cv.legacy CSRT/KCF/... when the OpenCV build has
        # it (the reference's tracker, main.py:11-32), else OpenCV's MIL tracker; the name says which.
        provider = "cv"
    if provider == "truth":
        if truth_frames is None:
            raise ValueError("tracker_provider 'truth' needs input_truth")
        key = (id(truth_frames), len(truth_frames), tuple(int(v) for v in first_boxes[0]))
        if key not in _truth_box_cache:            # the clip cache hands out the same list per video
            _truth_box_cache[key] = providers.truth_boxes(truth_frames, first_boxes[0])
        boxes = _truth_box_cache[key]

        def factory(frame, start_boxes, start_index=0):
            t = providers.ScriptedBoxTracker([boxes])
            t.i = start_index
            return t
        return factory, "truth"
    if provider == "static":
        return (lambda frame, start_boxes, start_index=0: providers.ScriptedBoxTracker([[b] for b in start_boxes])), "static"
    if callable(provider):
        return provider, "custom"
    name = "cv.legacy:%s" % (config.get("tracker") or "CSRT") if hasattr(cv, "legacy") else "cv:MIL"
    return (lambda frame, start_boxes, start_index=0: _CvMultiTracker(config.get("tracker"), frame, start_boxes)), name


def run_sequence(config, device=0, out_path=None, segment_fn=None, prior_fn=None, tracker_provider=None,
                 model_cache=None, cache_tag=None, max_frames=None, verbose=False):
    """Run one sequence (one video, one hyper-parameter set).  Returns a dict with `mean_iou`,
    `seconds` (the reference's tot_time: the frame loop only, main.py:275,364), `iou` (per box
    per frame), `n_frames`, `n_updates`, `train_seconds`, `tracker`."""
    t_wall = time.time()
    rf = config.get("resize_factor") or 1
    frames = read_clip(resolve_path(config["input_video"]), rf)
    if not frames:
        raise IOError("Fatal error! no frames in %s" % config["input_video"])
    truth_path = config.get("input_truth")
    truths = read_clip(resolve_path(truth_path), rf) if truth_path is not None else None
    debug = bool(config.get("debug")) and False          # GUI / writers are not part of the headless driver

    pts = config.get("pts")
    frame_numbers = config.get("pts_frame_numbers")
    ronis = config.get("bboxes_roni")
    if config.get("manual_roi_selection"):
        raise ValueError("manual_roi_selection needs a GUI (main.py:167-245); use the `pts` polygons")

    t_decode = time.time() - t_wall
    t_train = time.time()
    maskers, bboxes, colors = [], [], []
    for n_target, target_selection in enumerate(pts):
        bboxes.append([])
        extra = {}
        if config.get("masker") == "PC":
            extra = dict(segment_fn=segment_fn, prior_fn=prior_fn, device=device, model_cache=model_cache,
                         cache_tag=(cache_tag, n_target) if cache_tag is not None else None,
                         train_jobs=config.get("train_jobs"), fit_estimators=config.get("fit_estimators"))
        maskers.append(getMaskerByName(config.get("masker"), debug=debug, frame=frames[0], config=config,
                                       poly_roi=pts[n_target][0], update_mask=config.get("update_mask"), **extra))
        for n_selection, selection in enumerate(target_selection):
            if not config.get("multi_selection") and n_selection > 0:
                continue
            bbox = cv.boundingRect(np.array(selection))
            bboxes[-1].append(bbox)
            colors.append(COLORS[len(colors) % len(COLORS)])
            n_frame = frame_numbers[n_selection]
            if n_frame >= len(frames):
                raise IOError("Fatal error! selection frame %d beyond the clip" % n_frame)
            maskers[n_target].addModel(frame=frames[n_frame], poly_roi=pts[n_target][n_selection], bbox=bbox,
                                       bbox_roni=ronis[n_target][n_selection] if ronis is not None else None,
                                       n_frame=n_frame)
    t_train = time.time() - t_train

    first_boxes = [b[0] for b in bboxes]
    factory, tracker_name = make_tracker(config, frames[0], first_boxes, truths, tracker_provider)
    tracker = factory(frames[0], first_boxes)
    native = maskers[0].native
    custom = config.get("custom_trackers") or []

    ious = []
    n_updates = 0
    start = time.time()
    n = len(frames) if max_frames is None else min(len(frames), max_frames)
    for index in range(n):
        frame = frames[index]
        truth = truths[index] if truths is not None and index < len(truths) else None
        masked = np.zeros_like(frame, dtype=np.uint8)
        ok, boxes = tracker.update(frame)
        for i, newbox in enumerate(boxes):
            box = int_box(newbox)
            if config.get("masker") not in custom:
                status = maskers[i].update(bbox=box, frame=frame, mask=masked, color=colors[i])
                n_updates += 1
                if status is not None:                   # re-initialise the tracker (main.py:303-339)
                    if verbose:
                        print("RE-INITIALIZE TRACKER n. %d" % status)
                    tracker = factory(frame, [bboxes[t][status] for t in range(len(bboxes))], index + 1)
            if truth is not None:
                # truth is the decoded BGR frame; pcm_iou applies the BGR2GRAY arithmetic of main.py:285
                with stages.stage("gpu_iou_call"):
                    inter, union = native.iou_counts(masked[:, :, 2], truth)
                ious.append(inter / union if union else float("nan"))
    seconds = time.time() - start
    stages.add("frame_loop", seconds)
    stages.add("decode", t_decode)
    mean_iou = float(np.mean(ious)) if ious else float("nan")
    if out_path is not None:
        with open(out_path, "w") as f:
            f.write("%s;%s" % (mean_iou, seconds))
    for m in maskers:
        if hasattr(m, "close"):
            m.close()
    return dict(mean_iou=mean_iou, seconds=seconds, iou=ious, n_frames=n, n_updates=n_updates,
                train_seconds=t_train, tracker=tracker_name, decode_seconds=t_decode, wall_seconds=time.time() - t_wall)
