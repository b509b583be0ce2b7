"""Import shim that lets the UNMODIFIED reference (zerohd4869/MM-DFN `code/*.py`) run on CPU with torch 2.x.

TEST / BASELINE INFRASTRUCTURE ONLY (same rules as the oracle: nothing under ``mm-dfn_b200/`` imports it).
Used by ``tests/golden/make_golden*.py`` (golden generation, build container) and by ``bench.py --impl reference`` /
the ``cpu_baseline`` leg (the reference arm, from the git-ignored copy under ``baseline/_ref/code`` that
``__graft_entry__.build()`` stages from ``/root/reference``).  No reference file is edited; the four
incompatibilities of SURVEY.md F5 are patched from the outside:

  (a) ``torch_geometric`` is imported at module top (code/model.py:11) but not installed: stub modules whose
      RGCNConv / GraphConv raise on construction (the GDF path never constructs them);
  (b) ``.cuda()`` is hard-coded on the GDF path (code/model_mm.py:85,98,125): identity on CPU;
  (c) torch<=1.x treated a 2-D numpy index as a tuple of index arrays (code/model_mm.py:172, code/model.py:465-466):
      ``Tensor.__getitem__/__setitem__`` convert such an index to a tuple;
  (d) nothing else -- train-mode backward with real dropout works as is.
"""
import os
import sys
import types

import numpy as np
import torch

_installed = [None]


def ref_code_dir(root=None):
    """baseline/_ref/code of this repo if staged, else /root/reference/code (build container), else None."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) if root is None else root
    for p in (os.path.join(here, "baseline", "_ref", "code"), "/root/reference/code"):
        if os.path.exists(os.path.join(p, "model.py")):
            return p
    return None


def ref_data_dir(root=None):
    """directory holding the reference's feature pickles (build container only: they are not staged)"""
    return "/root/reference/data" if os.path.exists("/root/reference/data/iemocap/IEMOCAP_features.pkl") else None


def install(code_dir=None):
    """Patch the environment and put the reference's code directory first on sys.path.  Idempotent."""
    code_dir = code_dir or ref_code_dir()
    if code_dir is None:
        raise FileNotFoundError("the reference sources are not staged (baseline/_ref/code) and /root/reference is absent")
    if _installed[0] == code_dir:
        return code_dir
    if code_dir not in sys.path:
        sys.path.insert(0, code_dir)
    tg, tgnn = types.ModuleType("torch_geometric"), types.ModuleType("torch_geometric.nn")

    class _NA(torch.nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError("torch_geometric is not installed")

    tgnn.RGCNConv = tgnn.GraphConv = _NA
    tg.nn = tgnn
    sys.modules.setdefault("torch_geometric", tg)
    sys.modules.setdefault("torch_geometric.nn", tgnn)
    if not torch.cuda.is_available() or os.environ.get("MMDFN_REF_FORCE_CPU", "1") == "1":
        torch.Tensor.cuda = lambda self, *a, **k: self
    if _installed[0] is None:
        _si, _gi = torch.Tensor.__setitem__, torch.Tensor.__getitem__

        def fix(i):
            return tuple(torch.as_tensor(r) for r in i) if isinstance(i, np.ndarray) and i.ndim == 2 else i

        torch.Tensor.__setitem__ = lambda self, i, v: _si(self, fix(i), v)
        torch.Tensor.__getitem__ = lambda self, i: _gi(self, fix(i))
    _installed[0] = code_dir
    return code_dir


def reference_modules(code_dir=None):
    """(model, loss, model_mm, model_GCN) modules of the unmodified reference."""
    install(code_dir)
    import importlib
    return tuple(importlib.import_module(n) for n in ("model", "loss", "model_mm", "model_GCN"))


def make_reference_model(model_mod, d_text, d_audio, d_visual, S, C, K, dataset="IEMOCAP", spk_w="3-0-1", dropout=0.4,
                         reason_flag=True):
    """DialogueGNNModel of the reference in the configuration the authors' scripts run (GDF, LSTM, avl)."""
    return model_mod.DialogueGNNModel(
        "LSTM", d_text, 150, 150, 100, 100, 100, 100, n_speakers=S, max_seq_len=200, window_past=10,
        window_future=10, n_classes=C, dropout=dropout, nodal_attention=True, no_cuda=True, graph_type="GDF",
        alpha=0.2, lamda=0.5, multiheads=6, graph_construct="direct", use_GCN=False, use_residue=True,
        D_m_v=d_visual, D_m_a=d_audio, modals="avl", att_type="concat_subsequently", av_using_lstm=False,
        Deep_GCN_nlayers=K, dataset=dataset, use_speaker=False, use_modal=False, reason_flag=reason_flag,
        multi_modal=True, use_crn_speaker=True, speaker_weights=spk_w, modal_weight=1.0)
