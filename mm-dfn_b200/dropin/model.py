"""Drop-in for the reference's `code/model.py` (imported by code/run_train_erc.py:10)."""
import _bootstrap  # noqa: F401
from mmdfn_b200.modules import (Attention, DialogueGNNModel, MaskedEdgeAttention, MatchingAttention,  # noqa: F401
                                MMGatedAttention, SimpleAttention, simple_batch_graphify)
from mmdfn_b200.relation import (GraphNetwork, attentive_node_features, batch_graphify, classify_node_features,  # noqa: F401
                                 edge_perms)


def _outside_hot_path(name):
    class _Missing:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is a baseline outside the MM-DFN hot path (SURVEY.md section 2)")
    _Missing.__name__ = name
    return _Missing


LSTMModel = _outside_hot_path("LSTMModel")
GRUModel = _outside_hot_path("GRUModel")
DialogRNNModel = _outside_hot_path("DialogRNNModel")
