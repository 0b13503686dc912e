"""Deterministic synthetic sequences for the throughput configs of BASELINE.json
(configs[1]: 1920x1080, 300 frames, one target; configs[4]: 3840x2160, four targets).

Frame = smooth low-frequency background (bilinear-upsampled random field per
channel) + N(0, 6) noise, plus one textured blob per target: a sinusoidally
deforming 24-gon with a distinct colour distribution whose centre follows a
bounded random walk.  The blob outline doubles as the ROI polygon
(`polygons.yaml`-style `pts`), the filled blob as the ground-truth mask.
"""
import cv2 as cv
import numpy as np


class SyntheticSequence:
    def __init__(self, width=1920, height=1080, n_frames=300, seed=0, n_targets=1, radius=None):
        self.W, self.H, self.n_frames, self.seed, self.n_targets = width, height, n_frames, seed, n_targets
        rng = np.random.default_rng(seed)
        gw, gh = 32, 18
        self.field = rng.integers(40, 200, (gh, gw, 3)).astype(np.float32)
        self.bg = cv.resize(self.field, (width, height), interpolation=cv.INTER_LINEAR)
        R = radius or min(width, height) * (0.16 if n_targets == 1 else 0.09)
        self.R = R
        # per-target motion: bounded random walk of the centre, phases of the 24 vertices
        self.centres = []
        self.phases = rng.uniform(0, 2 * np.pi, (n_targets, 24))
        self.colors = []
        for t in range(n_targets):
            cx = width * (t + 1) / (n_targets + 1)
            cy = height * (0.5 if n_targets == 1 else (0.35 + 0.3 * (t % 2)))
            steps = rng.normal(0, 2.0, (n_frames, 2)).cumsum(axis=0)
            lim = np.array([width * 0.1, height * 0.1])
            steps = np.clip(steps, -lim, lim)
            self.centres.append(np.stack([cx + steps[:, 0], cy + steps[:, 1]], 1))
            self.colors.append(np.array([[30, 40, 220], [40, 200, 230], [220, 60, 200], [60, 230, 60]][t % 4],
                                        np.float32))

    def polygon(self, target, i):
        """24 integer [x, y] vertices of target `target` at frame i."""
        a = np.arange(24) * (2 * np.pi / 24)
        r = self.R * (1 + 0.22 * np.sin(0.07 * i + self.phases[target]))
        c = self.centres[target][i]
        pts = np.stack([c[0] + r * np.cos(a), c[1] + r * np.sin(a)], 1)
        pts[:, 0] = np.clip(pts[:, 0], 0, self.W - 1)
        pts[:, 1] = np.clip(pts[:, 1], 0, self.H - 1)
        return np.rint(pts).astype(np.int32).tolist()

    def truth(self, i):
        m = np.zeros((self.H, self.W), np.uint8)
        for t in range(self.n_targets):
            cv.fillPoly(m, np.array([self.polygon(t, i)], np.int32), 255)
        return m

    def frame(self, i):
        rng = np.random.default_rng((self.seed, i))
        img = self.bg + rng.normal(0, 6, self.bg.shape).astype(np.float32)
        for t in range(self.n_targets):
            m = np.zeros((self.H, self.W), np.uint8)
            cv.fillPoly(m, np.array([self.polygon(t, i)], np.int32), 255)
            ys, xs = np.nonzero(m)
            tex = 25 * np.sin(xs * 0.35 + ys * 0.2 + t)[:, None] + rng.normal(0, 10, (len(xs), 3))
            img[ys, xs] = self.colors[t] + tex.astype(np.float32)
        return np.clip(np.rint(img), 0, 255).astype(np.uint8)

    def bbox(self, target, i):
        return cv.boundingRect(np.array(self.polygon(target, i), np.int32))

    def roni(self, target=0):
        """A background rectangle [x, y, w, h] away from every blob's walk area."""
        w, h = max(8, self.W * 200 // 1920), max(8, self.H * 100 // 1080)
        return [4, 4, int(w), int(h)]
