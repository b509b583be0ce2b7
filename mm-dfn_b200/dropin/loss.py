"""Drop-in for the reference's `code/loss.py` (imported by code/run_train_erc.py:16)."""
import _bootstrap  # noqa: F401
from mmdfn_b200.modules import FocalLoss  # noqa: F401


class MaskedNLLLoss:
    """code/loss.py:38-58 is the loss of the non-graph baselines (LSTMModel / GRUModel / DialogRNNModel, reached at
    code/run_train_erc.py:510 only when --graph_model is off).  The name exists so the trainer's import succeeds."""

    def __init__(self, *a, **k):
        raise NotImplementedError("MaskedNLLLoss belongs to the baselines outside the MM-DFN hot path (SURVEY.md section 2)")
