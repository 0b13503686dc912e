"""Import the UNMODIFIED reference (/root/reference) in this container.

Test infrastructure only: used by tests/golden/make_golden.py (committed
generator of the golden fixtures) and by CPU tests that are skipped when
/root/reference is absent (it never exists on the GPU box).

The reference's `maskers/__init__.py:1-8` imports every masker, which pulls in
matplotlib / skimage (not installed here).  Those packages are not on the PC
hot path except `skimage.segmentation.{quickshift,felzenszwalb,slic}`
(`maskers/pixel_classification.py:70-75`), whose output we INJECT: both sides
of every parity test consume the same label map ("parity
unpinned" for the over-segmentation itself).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PCM_REFERENCE_ROOT", "/root/reference")

# label provider the stubbed skimage functions delegate to; set by the caller.
_segment_provider = {"fn": None}


def set_segment_provider(fn):
    """fn(crop_bgr: HxWx3 u8) -> HxW integer label map (labels 0..S-1)."""
    _segment_provider["fn"] = fn


def _segments(crop, *a, **k):
    fn = _segment_provider["fn"]
    if fn is None:
        raise RuntimeError("ref_shim: no segment provider set")
    return fn(crop)


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "maskers", "pixel_classification.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load():
    """Returns (maskers_module, benchmark_module) of the reference."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
        tk = _stub("mpl_toolkits")
        tk.mplot3d = _stub("mpl_toolkits.mplot3d", Axes3D=object)
    if "skimage" not in sys.modules:
        sk = _stub("skimage")
        none = lambda *a, **k: None
        sk.segmentation = _stub("skimage.segmentation", slic=_segments, quickshift=_segments,
                                felzenszwalb=_segments, mark_boundaries=none, watershed=none)
        sk.filters = _stub("skimage.filters", sobel=none)
        sk.color = _stub("skimage.color", rgb2gray=none)
    # the reference is a directory of scripts, imported as top-level modules;
    # keep them under private names so they cannot collide with the repo's
    # own `maskers` / `benchmark` drop-in modules.
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k == "maskers" or k.startswith("maskers.") or k == "benchmark" or k == "prim"}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import maskers as ref_maskers
        import benchmark as ref_benchmark
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in list(sys.modules):
            if k == "maskers" or k.startswith("maskers.") or k == "benchmark" or k == "prim" or k.startswith("prim."):
                sys.modules["_ref_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    return ref_maskers, ref_benchmark
