"""mmdfn_b200 -- B200 (sm_100a) implementation of MM-DFN's per-dialogue forward/backward hot path.

    from mmdfn_b200 import DialogueGNNModel, FocalLoss          # same signatures as the reference
or put ``mm-dfn_b200/dropin`` first on ``sys.path`` and run the reference's
``code/run_train_erc.py`` unchanged."""
from ._lib import MMDFNError, SO_PATH, lib  # noqa: F401
from .modules import (BlockAdj, DialogueGNNModel, FocalLoss, GCNII_lyc, GraphConvolution, MaskedEdgeAttention,  # noqa: F401
                      LMF, MFN, MM_GCN, MMGatedAttention, TFN, simple_batch_graphify)
from .ops import DialogGeom  # noqa: F401

__version__ = "0.1.0"
