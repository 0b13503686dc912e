"""Masker plugin base class -- same contract as the reference's maskers/masker.py:3-14:
constructor keeps `debug`, a copy of the first frame (`prevFrame`) and the whole
config dict; `update()` and `addModel()` are no-ops to be overridden; unknown
keyword arguments are swallowed."""
from abc import ABC


class Masker(ABC):
    def __init__(self, debug=False, frame=None, config=None, **others):
        self.debug = debug
        self.prevFrame = None if frame is None else frame.copy()
        self.config = config

    def update(self, *args, **kwargs):
        return None

    def addModel(self, frame, poly_roi, bbox, n_frame, bbox_roni=None, show_prob_map=False):
        return None
