"""Drop-in for the reference's `code/model_mm.py`."""
import _bootstrap  # noqa: F401
from mmdfn_b200.modules import GraphConvolution, MM_GCN  # noqa: F401
