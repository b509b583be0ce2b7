"""Drop-in for the reference's `code/dataloader.py` (imported by code/run_train_erc.py:9)."""
import _bootstrap  # noqa: F401
from mmdfn_b200.dataloader import DailyDialogueDataset, IEMOCAPDataset, MELDDataset  # noqa: F401
