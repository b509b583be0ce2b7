"""Drop-in for the reference's `code/model_fusion.py` (imported by code/model.py:992-1001 for att_type 'mfn' / 'tfn_only' / 'lmf_only')."""
import _bootstrap  # noqa: F401
from mmdfn_b200.modules import LMF, MFN, TFN  # noqa: F401
