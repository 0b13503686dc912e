"""Masker factory with the reference's interface (maskers/__init__.py:10-22).

Only the PixelClassification masker ("PC") is in scope of this B200-native
re-implementation; the other reference maskers are different
algorithms and are reported as unavailable instead of silently substituted.
"""
from .masker import Masker
from .pixel_classification import PixelClassificationNonRigidMasker

_OUT_OF_SCOPE = ("OpticalFlow", "BgSub", "LinPuntracker", "GrabCut")


def getMaskerByName(name, **args):
    if name == "PC":
        return PixelClassificationNonRigidMasker(**args)
    if name in _OUT_OF_SCOPE:
        exit("Masker %r is not part of the B200-native PC hot path" % name)
    exit("Masker name not found")
