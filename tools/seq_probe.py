"""Timing probe: one clip-resident sequence (pcm.fastseq) per configuration, single stream; host enqueue time,
total time and the per-kernel device times (CUDA events around every launch)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "non-rigid-object-tracking_b200")
sys.path[:0] = [PKG, os.path.join(ROOT, "tests")]
import yaml
from helpers import polygons
from pcm import fastseq, stages, sweep, capi

base = yaml.full_load(open(os.path.join(PKG, "config_benchmark.yaml")))
video = sys.argv[1] if len(sys.argv) > 1 else "frog"
clip = fastseq.ClipContext("Input/SegTrack2/Video/%s.mp4" % video, "Input/SegTrack2/Truth/%s.mp4" % video, 1, 0)
cache = sweep.ModelCache()
prof = {}
orig = capi.Handle.run_frames


def patched(self, *a):
    self.profile_enable(os.environ.get("PROBE_PROFILE", "1") == "1")
    t0 = time.perf_counter()
    orig(self, *a)
    t1 = time.perf_counter()
    self.synchronize()
    t2 = time.perf_counter()
    prof["enqueue_ms"], prof["wait_ms"] = 1e3 * (t1 - t0), 1e3 * (t2 - t1)
    prof["kernels"] = self.profile_read(reset=True)


capi.Handle.run_frames = patched
for feats, D, T, nov, seg, pw in [("8 hsv_lab", 7, 20, False, "quickshift", 0.0), ("8 hsv_lab", 7, 20, True, "quickshift", 0.0),
                                  ("8 hsv_lab", 10, 30, False, "quickshift", 0.0), ("8 hsv_lab", 10, 30, True, "felzenszwalb", 0.1),
                                  ("6 lab", 10, 30, True, "quickshift", 0.1), ("6 lab", 7, 20, False, "felzenszwalb", 0.0)]:
    params = dict(n_estimators=T, max_depth=D, n_components=1, novelty_detection=nov, over_segmentation=seg, features=feats,
                  dilation_kernel=7, prior_weight=pw)
    cfg = sweep.sequence_config(base, polygons(), video, params, "Input/SegTrack2/Video", "Input/SegTrack2/Truth")
    cfg["fit_estimators"] = 30
    for rep in range(2):
        stages.reset()
        r = fastseq.run_sequence_fast(cfg, clip, model_cache=cache, cache_tag=video)
    k = prof["kernels"]
    n = r["n_frames"]
    print("%s %s D%d T%d nov=%d %s pw=%.1f: loop %.1f ms (%.3f ms/frame) enqueue %.1f wait %.1f | per frame us: %s | stages %s" % (
        video, feats, D, T, nov, seg[:4], pw, 1e3 * r["seconds"], 1e3 * r["seconds"] / n, prof["enqueue_ms"], prof["wait_ms"],
        {kk: round(1e3 * v[0] / max(v[1], 1), 1) for kk, v in k.items() if v[1]},
        {a: round(b, 3) for a, b in stages.snapshot().items()}), flush=True)
clip.close()
