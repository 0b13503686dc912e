"""The C-ABI library on a machine WITHOUT a GPU: it loads, exports every function that
include/pcm_b200.h declares (and nothing is declared that the ctypes binding does not know),
and fails loudly -- there is no CPU fallback behind the boundary."""
import ctypes as C
import os
import re

import pytest

from helpers import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "pcm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcm_[a-z0-9_]+)\s*\(", text)))


def test_header_binding_and_library_agree():
    from pcm import capi
    names = _declared()
    assert len(names) >= 25
    lib = capi.load_library()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in pcm_b200.h but not exported: %s" % missing
    assert sorted(capi.EXPORTED_SYMBOLS) == names, "pcm/capi.py and include/pcm_b200.h list different entry points"
    assert lib.pcm_abi_version() == 1


def test_crop_rect_needs_no_device():
    """pcm_crop_rect is host integer logic (reference :49-51): bbox quirks incl. negative x."""
    from pcm import capi
    import pcm_oracle as orc
    H, W = 224, 528
    for bbox in [(419, 16, 30, 200), (5, 3, 40, 40), (-7, 10, 50, 30), (500, 200, 60, 60), (0, 0, 528, 224),
                 (-40, -40, 10, 10), (600, 300, 5, 5)]:
        eb = orc.enlarge_bbox(bbox, (H, W, 3))
        y0, y1 = orc.slice_extent(eb[1], eb[3], H)
        x0, x1 = orc.slice_extent(eb[0], eb[2], W)
        assert capi.crop_rect(bbox, H, W) == (x0, y0, max(x1 - x0, 0), max(y1 - y0, 0)), bbox


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from pcm import capi
    with pytest.raises(capi.PcmError) as e:
        capi.Handle(0)
    assert "no CUDA device" in str(e.value) and "no CPU path" in str(e.value)
    from maskers import getMaskerByName
    with pytest.raises(capi.PcmError):
        getMaskerByName("PC", debug=False, frame=None, config=dict(multi_selection=False, params=dict(
            features="8 hsv_lab", over_segmentation="quickshift")), poly_roi=None, update_mask=False)


def test_entry_points_reject_null_arguments_before_touching_a_device():
    """Argument validation of the device-pointer entry points needs no GPU: a NULL handle / pointer is PCM_E_INVALID
    with a message, not a crash (the reference's ctypes binding of prim/ passes raw pointers the same way)."""
    from pcm import capi
    lib = capi.load_library()
    n = C.c_int(0)
    rect = (C.c_int * 4)(0, 0, 4, 4)
    assert lib.pcm_quickshift_device(None, None, 8, 8, 24, rect, 0.5, 3.0, 6.0, None, None, C.byref(n)) != 0
    assert b"pcm_quickshift_device" in lib.pcm_last_error()
    assert lib.pcm_quickshift_device_batch(None, 1, None, 192, None, 8, 8, 24, None, 0.5, 3.0, 6.0, None, None, None, None) != 0
    assert b"pcm_quickshift_device_batch" in lib.pcm_last_error()
    assert lib.pcm_run_frames(None, 8, 8, 24, None, 8, None, 0) != 0
    assert lib.pcm_update_device(None, None, 8, 8, 24, None, None, 0, None, None, None, 8) != 0
    assert lib.pcm_iou_device(None, None, 8, None, 8, 1, 8, 8, None) != 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "non-rigid-object-tracking_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                t = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"^\s*(import|from)\s+(oracle|pcm_oracle|ref_port|quickshift_oracle)\b", t, flags=re.M) or \
                        re.search(r"sys\.path.*oracle", t):
                    bad.append(os.path.join(d, f))
            elif f.endswith((".cu", ".cuh", ".h", ".cpp")):
                t = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"#\s*include\s*[\"<][^\">]*oracle", t):
                    bad.append(os.path.join(d, f))
    assert not bad, "product files referring to oracle/: %s" % bad


def test_forest_export_is_predict_proba_for_fractions_and_counts():
    """`tree_.value` holds class fractions from scikit-learn 1.4 on and weighted class COUNTS
    before (the reference pins 0.24.1); the exporter must hand the C ABI what predict_proba reports in
    both cases (leaf values in [0, 1]; forest_p1 == clf.predict_proba bit for bit)."""
    import numpy as np
    from sklearn.ensemble import RandomForestClassifier
    from pcm import capi
    import pcm_oracle as orc
    rng = np.random.default_rng(3)
    Xi = rng.integers(-1, 256, (600, 27)).astype(np.int16)
    y = (Xi[:, 3] + Xi[:, 11] > 250).astype(np.int64)
    clf = RandomForestClassifier(random_state=42, n_estimators=7, max_depth=6).fit(Xi.astype(np.float64) / 255, y)
    want = clf.predict_proba(Xi.astype(np.float64) / 255)[:, 1]
    got = orc.forest_p1(orc.export_forest(clf), Xi)
    assert np.array_equal(got, want)
    for est in clf.estimators_:
        v = capi.class1_fraction(est.tree_)
        assert v.min() >= 0.0 and v.max() <= 1.0
        assert np.array_equal(v, orc.class1_fraction(est.tree_))
        # the same tree as an old scikit-learn would store it: weighted class counts per node
        class Counts:
            value = est.tree_.value * est.tree_.weighted_n_node_samples[:, None, None]
        c = capi.class1_fraction(Counts)
        assert c.min() >= 0.0 and c.max() <= 1.0
        old = Counts.value[:, 0, :].copy()
        norm = old.sum(axis=1)[:, None]
        norm[norm == 0.0] = 1.0
        assert np.array_equal(c, (old / norm)[:, 1])          # 0.24's predict_proba arithmetic
        assert np.allclose(c, v, rtol=0, atol=1e-12)
