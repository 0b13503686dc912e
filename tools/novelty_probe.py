"""Short clip-resident sequence with PCA novelty detection on (for an ncu capture of score_kernel on a small crop)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "non-rigid-object-tracking_b200")
sys.path[:0] = [PKG, os.path.join(ROOT, "tests")]
import yaml
from helpers import polygons
from pcm import fastseq, sweep

base = yaml.full_load(open(os.path.join(PKG, "config_benchmark.yaml")))
params = dict(n_estimators=20, max_depth=7, n_components=1, novelty_detection=True, over_segmentation="quickshift",
              features="8 hsv_lab", dilation_kernel=7, prior_weight=0.0)
cfg = sweep.sequence_config(base, polygons(), "frog", params, "Input/SegTrack2/Video", "Input/SegTrack2/Truth")
clip = fastseq.ClipContext(cfg["input_video"], cfg["input_truth"], 1, 0, max_frames=int(sys.argv[1]) if len(sys.argv) > 1 else 12)
r = fastseq.run_sequence_fast(cfg, clip)
print("frames", r["n_frames"], "mean iou", r["mean_iou"])
clip.close()
