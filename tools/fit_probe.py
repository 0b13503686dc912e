"""Timing probe: GPU forest fit (pcm_fit_forest) against scikit-learn on the sweep's training sets."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "non-rigid-object-tracking_b200"), os.path.join(ROOT, "tests")]
import cv2 as cv
import numpy as np
from helpers import polygons, read_video
from pcm import capi, train
from sklearn.ensemble import RandomForestClassifier

for video, features in [("soldier", "8 hsv_lab"), ("bmx", "8 hsv_lab"), ("frog", "6 lab")]:
    P = polygons()[video]
    frames = read_video("Video", video)
    h = capi.Handle(0)
    tok = features.split()
    h.set_features(int(tok[0]), tok[1].split("_"))
    pts, roni = P["pts"][0][0], P["bboxes_roni"][0][0]
    f = frames[P["pts_frame_numbers"][0]]
    x, y, w, hh = cv.boundingRect(np.array(pts))
    roi = np.zeros((hh, w), np.uint8)
    cv.fillPoly(roi, np.array([[(p[0] - x, p[1] - y) for p in pts]], dtype=np.int32), 255)
    t0 = time.time()
    X = np.concatenate([h.gather_features(f, (x, y, w, hh)), h.gather_features(f, tuple(roni))])
    lab = np.concatenate([(roi.reshape(-1) > 0).astype(np.int64), np.zeros(roni[2] * roni[3], np.int64)])
    t_rows = time.time() - t0
    for D in (7, 10):
        t0 = time.time()
        draws = train.tree_draws(len(lab), 30)
        t1 = time.time()
        g = train.fit_forest(h, X, lab, 30, D, draws=draws)
        t2 = time.time()
        g = train.fit_forest(h, X, lab, 30, D, draws=draws, rows_id=5)
        t3 = time.time()
        g = train.fit_forest(h, X, lab, 30, D, draws=draws, rows_id=5, rows_resident=True)
        t4 = time.time()
        clf = RandomForestClassifier(random_state=42, n_estimators=30, max_depth=D, n_jobs=os.cpu_count()).fit(X.astype(np.float64) / 255, lab)
        t5 = time.time()
        same = all(np.array_equal(e.tree_.threshold, t[1]) and np.array_equal(e.tree_.feature, t[0]) for e, t in zip(clf.estimators_, g.trees))
        print("%s %s rows %d D %d: rows %.3fs draws %.3fs gpu fit %.3fs / %.3fs, resident %.3fs; sklearn(%d jobs) %.3fs; equal %s; nodes/tree %.0f"
              % (video, features, len(lab), D, t_rows, t1 - t0, t2 - t1, t3 - t2, t4 - t3, os.cpu_count(), t5 - t4, same,
                 np.mean([len(t[0]) for t in g.trees])), flush=True)
    h.close()
