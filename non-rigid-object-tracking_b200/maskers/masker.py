"""Plugin contract shared by every masker (what the reference's maskers/masker.py:3-14 defines):

    m = SomeMasker(debug=..., frame=first_frame, config=whole_yaml_dict, **anything_else)
    m.addModel(frame, poly_roi, bbox, n_frame, bbox_roni=None, show_prob_map=False)
    status = m.update(bbox=..., frame=..., mask=..., color=...)

State kept for subclasses: `debug`, `config` (the entire configuration mapping, not only
`params`) and `prevFrame`, a private copy of the frame the masker was created on.  Keyword
arguments a subclass does not know are accepted and dropped, exactly like the reference does,
so that the sequence driver can pass one argument set to every masker type.
"""
import abc


class Masker(abc.ABC):
    """Base class; both hooks are harmless no-ops until a subclass overrides them."""

    def __init__(self, debug=False, frame=None, config=None, **others):
        del others                                   # tolerated, unused
        self.config = config
        self.debug = debug
        self.prevFrame = frame.copy() if frame is not None else None

    def addModel(self, frame, poly_roi, bbox, n_frame, bbox_roni=None, show_prob_map=False):
        """Register a (frame, polygon) training selection; returns the RONI box that was used."""
        return None

    def update(self, *args, **kwargs):
        """Process one frame; `None` means 'carry on', an int asks the driver to re-seat its tracker."""
        return None
