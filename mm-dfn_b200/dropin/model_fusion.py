"""Drop-in for the reference's `code/model_fusion.py` (imported by code/model.py:992 for att_type 'mfn')."""
import _bootstrap  # noqa: F401
from mmdfn_b200.modules import MFN  # noqa: F401


def _outside_hot_path(name):
    class _Missing:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is a fusion ablation outside the MM-DFN hot path (SURVEY.md section 2)")
    _Missing.__name__ = name
    return _Missing


TFN = _outside_hot_path("TFN")
LMF = _outside_hot_path("LMF")
