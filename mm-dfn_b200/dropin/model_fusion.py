"""Drop-in for the reference's `code/model_fusion.py` (imported by code/model.py:992-1001 for att_type 'mfn' / 'lmf_only')."""
import _bootstrap  # noqa: F401
from mmdfn_b200.modules import LMF, MFN  # noqa: F401


def _outside_hot_path(name):
    class _Missing:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is a fusion ablation outside the MM-DFN hot path (SURVEY.md section 2)")
    _Missing.__name__ = name
    return _Missing


TFN = _outside_hot_path("TFN")
