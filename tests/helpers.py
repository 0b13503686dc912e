"""Shared helpers for the test-suite (golden loading, video decode, model rebuild)."""
import hashlib
import json
import os

import cv2 as cv
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "non-rigid-object-tracking_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")

SEQ_NAMES = ["soldier_default", "parachute_novelty", "worm_rgb3", "frog_sweep", "bmx_sweep", "soldier_prior"]


def sha1(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


_video_cache = {}


def read_video(kind, name):
    key = (kind, name)
    if key not in _video_cache:
        cap = cv.VideoCapture(os.path.join(PKG, "Input", "SegTrack2", kind, name + ".mp4"))
        frames = []
        while True:
            ok, f = cap.read()
            if not ok:
                break
            frames.append(f)
        _video_cache[key] = frames
    return _video_cache[key]


class GoldenSeq:
    """One tests/golden/seq_<name>.npz with convenience accessors."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, "seq_%s.npz" % name), allow_pickle=False)
        self.meta = json.loads(str(self.z["meta"]))
        self.name = name
        self.params = self.meta["params"]
        self.config = dict(multi_selection=self.meta["multi_selection"], params=dict(self.params))
        self.frames = read_video("Video", self.meta["video"])
        self.truth = read_video("Truth", self.meta["video"])
        self.n_models = self.meta["n_models"]

    def frames_match(self):
        """The decoded frames are the ones the golden vectors were made from."""
        return all(sha1(self.frames[i]) == str(self.z["frame_sha1"][i]) for i in range(self.meta["n_frames"]))

    def tree_arrays(self, m):
        z = self.z
        offs = z["m%d_offsets" % m]
        out = []
        for t in range(len(offs) - 1):
            s = slice(int(offs[t]), int(offs[t + 1]))
            out.append((z["m%d_feature" % m][s], z["m%d_threshold" % m][s], z["m%d_left" % m][s],
                        z["m%d_right" % m][s], z["m%d_value1" % m][s]))
        return out

    def pca(self, m):
        z = self.z
        if ("m%d_pca_mean" % m) not in z:
            return None
        return z["m%d_pca_mean" % m], z["m%d_pca_comp" % m]

    def novelty_threshold(self, m):
        return float(self.z["m%d_threshold_novelty" % m])

    def has_recorded_priors(self):
        return "priors_recorded" in self.z

    def recorded_priors(self, i):
        """Priors the reference's computePriors returned at frame i (its FLANN matcher is randomised, so the golden
        run's priors are injected wherever a sequence with prior_weight != 0 is replayed)."""
        return self.z["pri%d" % i]

    def sample_frames(self, budget=28):
        """Frames for the (slow) CPU port: everything for short goldens, else the start, both sides of every model
        switch, the dumped frames and an even sprinkling -- per-frame results depend on (index, current model) only."""
        n = self.meta["n_frames"]
        if n <= budget or self.has_recorded_priors():
            return list(range(n))
        keep = set(range(6)) | set(self.meta["dump"]) | {n - 1}
        for f in self.model_frames()[1:]:
            keep |= {f - 2, f - 1, f, f + 1}
        keep |= set(np.linspace(0, n - 1, 8).astype(int).tolist())
        return sorted(k for k in keep if 0 <= k < n)

    def model_frames(self):
        return self.meta["pts_frame_numbers"][:self.n_models]

    def state_at(self, i):
        """(index, current_model) the masker holds when frame i is processed."""
        nf = self.model_frames()
        cur = 0
        if self.meta["multi_selection"]:
            while cur + 1 < len(nf) and i >= nf[cur + 1]:
                cur += 1
        return i, cur


def polygons():
    import yaml
    with open(os.path.join(PKG, "polygons.yaml")) as f:
        return yaml.full_load(f)


def native_masker_from_golden(g, device=0, **kw):
    """The product masker with its models injected from the golden tree arrays
    (no sklearn fitting): isolates the per-frame hot path."""
    from maskers import getMaskerByName
    from pcm.providers import make_segment_provider
    m = getMaskerByName("PC", debug=False, frame=g.frames[0], config=g.config, poly_roi=None, update_mask=False,
                        segment_fn=make_segment_provider(g.meta["segments"]), device=device, **kw)
    if g.has_recorded_priors():
        m.prior_fn = lambda *a, **k: g.recorded_priors(m.index)
    m.native.set_debug(True)
    for s in range(g.n_models):
        idx = m.native.add_model_arrays(g.model_frames()[s], g.tree_arrays(s))
        pca = g.pca(s)
        if pca is not None:
            m.native.set_novelty(idx, pca[0], pca[1][0])
        m.models.append({"n_frame": g.model_frames()[s], "model": None})
        m.novelty_det.append({"n_frame": g.model_frames()[s], "model": None, "threshold": g.novelty_threshold(s)})
    return m
