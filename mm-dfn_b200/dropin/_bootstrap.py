"""Makes `mmdfn_b200` importable when only this directory is on sys.path (the reference's
trainer is started as `python code/run_train_erc.py` with PYTHONPATH=<repo>/mm-dfn_b200/dropin)."""
import os
import sys

_repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _repo not in sys.path:
    sys.path.insert(0, _repo)
import mmdfn_b200  # noqa: E402,F401
