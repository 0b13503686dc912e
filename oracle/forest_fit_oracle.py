"""Python face of oracle/forest_fit_oracle.c -- TEST INFRASTRUCTURE ONLY (see that file's header).

`fit_forest(Xint, y, n_estimators, max_depth, random_state=42)` grows the trees that
`RandomForestClassifier(random_state=42, n_estimators=.., max_depth=..).fit(Xint / 255, y)` grows
(reference maskers/pixel_classification.py:199-200) and returns them as the raw arrays the C ABI's
pcm_add_model takes.  The per-tree seeds, bootstrap counts and splitter seeds are drawn with numpy exactly
as scikit-learn 1.9.0 draws them:
    sklearn/ensemble/_base.py:77-84      _set_random_states: tree seed = forest_rs.randint(MAX_INT32)
    sklearn/ensemble/_forest.py:95-103   _generate_sample_indices: RandomState(seed).randint(0, n, n)
    sklearn/ensemble/_forest.py:148-165  _parallel_build_trees: sample_weight = bincount(indices)
    sklearn/tree/_splitter.pyx:155       rand_r_state = RandomState(seed).randint(0, RAND_R_MAX)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "forest_fit_oracle.c")
LIB = os.path.join(HERE, "_build", "libforest_fit_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        P, I = C.c_void_p, C.c_int
        lib.forest_fit_oracle_tree.restype = I
        lib.forest_fit_oracle_tree.argtypes = [P, P, P, I, I, I, I, C.c_uint32, I, P, P, P, P, P, P, P, P, P]
        _lib = lib
    return _lib


def tree_draws(n_samples, n_estimators, random_state=42):
    """[(bootstrap counts int32[n], rand_r_state)] per tree, drawn like scikit-learn does."""
    rs = np.random.RandomState(random_state)
    out = []
    for _ in range(n_estimators):
        seed = rs.randint(np.iinfo(np.int32).max)
        idx = np.random.RandomState(seed).randint(0, n_samples, n_samples)
        counts = np.bincount(idx, minlength=n_samples).astype(np.int32)
        out.append((counts, int(np.random.RandomState(seed).randint(0, 2147483647))))
    return out


def fit_tree(Xint, y, counts, rand_r_state, max_depth, max_features=None):
    lib = _load()
    Xint = np.ascontiguousarray(Xint, np.int16)
    n, F = Xint.shape
    y8 = np.ascontiguousarray(y, np.uint8)
    counts = np.ascontiguousarray(counts, np.int32)
    if max_features is None:
        max_features = max(1, int(np.sqrt(F)))
    cap = int(min(2 ** (min(max_depth, 24) + 1), 2 * n + 1))
    feature = np.empty(cap, np.int32)
    thr = np.empty(cap, np.float64)
    left = np.empty(cap, np.int32)
    right = np.empty(cap, np.int32)
    v0 = np.empty(cap, np.float64)
    v1 = np.empty(cap, np.float64)
    nns = np.empty(cap, np.int32)
    wn = np.empty(cap, np.float64)
    imp = np.empty(cap, np.float64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    k = lib.forest_fit_oracle_tree(p(Xint), p(y8), p(counts), n, F, int(max_depth), int(max_features),
                                   C.c_uint32(rand_r_state), cap, p(feature), p(thr), p(left), p(right), p(v0), p(v1),
                                   p(nns), p(wn), p(imp))
    if k < 0:
        raise RuntimeError("forest_fit_oracle: node capacity exceeded")
    return dict(feature=feature[:k].copy(), threshold=thr[:k].copy(), left=left[:k].copy(), right=right[:k].copy(),
                value0=v0[:k].copy(), value1=v1[:k].copy(), n_node_samples=nns[:k].copy(), weighted_n=wn[:k].copy(),
                impurity=imp[:k].copy())


def fit_forest(Xint, y, n_estimators, max_depth, random_state=42):
    """List of per-tree dicts (see fit_tree)."""
    y = np.asarray(y)
    if set(np.unique(y).tolist()) != {0, 1}:
        raise ValueError("binary {0,1} labels with both classes present are required")
    return [fit_tree(Xint, y, c, s, max_depth) for c, s in tree_draws(len(y), n_estimators, random_state)]


def as_tree_arrays(trees):
    """[(feature, threshold, left, right, value1)] -- the layout of pcm_oracle.sklearn_tree_arrays."""
    return [(t["feature"], t["threshold"], t["left"], t["right"], t["value1"]) for t in trees]
