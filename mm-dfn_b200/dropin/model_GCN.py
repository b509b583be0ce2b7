"""Drop-in for the reference's `code/model_GCN.py`."""
import _bootstrap  # noqa: F401
from mmdfn_b200.modules import GCNII, GCNII_lyc, GraphConvolution  # noqa: F401
