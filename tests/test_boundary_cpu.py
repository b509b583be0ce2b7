"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the ctypes binding is derived from the header, state_dict keys/shapes match the
reference's manifest, and the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import mmdfn_b200
from mmdfn_b200 import _lib
from helpers import manifest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    sigs = _lib.parse_header()
    assert len(sigs) >= 25
    assert os.path.exists(_lib.SO_PATH), "build with __graft_entry__.build()"
    L = ctypes.CDLL(_lib.SO_PATH)
    for name in sigs:
        assert hasattr(L, name), name
    L.mmdfn_abi_version.restype = ctypes.c_int
    assert L.mmdfn_abi_version() == 2


def test_header_prototypes_match_definitions():
    """every extern "C" definition in csrc has a prototype in the header with the same argument count"""
    sigs = _lib.parse_header()
    src = ""
    cs = os.path.join(ROOT, "mm-dfn_b200", "csrc")
    for f in os.listdir(cs):
        if f.endswith(".cu"):
            src += open(os.path.join(cs, f)).read()
    defs = re.findall(r'extern "C"\s+(?:int|long long)\s+(mmdfn_\w+)\s*\(([^)]*)\)\s*\{', src, flags=re.S)
    assert {d[0] for d in defs} == set(sigs)
    for name, args in defs:
        n = 0 if args.strip() in ("", "void") else len(args.split(","))
        assert n == len(sigs[name][1]), name


def _make(dt, da, dv, S, C, K, dataset, spk):
    return mmdfn_b200.DialogueGNNModel(
        "LSTM", dt, 150, 150, 100, 100, 100, 100, n_speakers=S, max_seq_len=200, window_past=10, window_future=10,
        n_classes=C, dropout=0.4, nodal_attention=True, no_cuda=True, graph_type="GDF", alpha=0.2, lamda=0.5,
        multiheads=6, graph_construct="direct", use_GCN=False, use_residue=True, D_m_v=dv, D_m_a=da, modals="avl",
        att_type="concat_subsequently", av_using_lstm=False, Deep_GCN_nlayers=K, dataset=dataset, use_speaker=False,
        use_modal=False, reason_flag=True, multi_modal=True, use_crn_speaker=True, speaker_weights=spk, modal_weight=1.0)


@pytest.mark.parametrize("tag,args", [("iemocap_k2", (100, 1582, 342, 2, 6, 2, "IEMOCAP", "3-0-1")),
                                      ("meld_k4", (600, 300, 342, 9, 7, 4, "MELD", "0.5-0.5-1.5"))])
def test_state_dict_matches_reference_manifest(tag, args):
    m = _make(*args)
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    ref = manifest(tag)
    assert list(mine) == list(ref)          # same keys in the same order
    assert mine == ref


def test_unsupported_configurations_raise():
    with pytest.raises(NotImplementedError):
        mmdfn_b200.DialogueGNNModel("DialogRNN", 100, 150, 150, 100, 100, 100, 100, 2, 200, 10, 10, graph_type="GDF")
    with pytest.raises(NotImplementedError):                      # the ctor's default att_type='gated' exists on 'relation' only
        mmdfn_b200.DialogueGNNModel("LSTM", 100, 150, 150, 100, 100, 100, 100, 2, 200, 10, 10, graph_type="GDF")
    with pytest.raises(NotImplementedError):
        mmdfn_b200.DialogueGNNModel("LSTM", 100, 150, 150, 100, 100, 100, 100, 2, 200, 10, 10, graph_type="relation",
                                    att_type="tfn_only")
    m = mmdfn_b200.DialogueGNNModel("LSTM", 100, 150, 150, 100, 100, 100, 100, 2, 200, 10, 10, graph_type="GDF", n_classes=6,
                                    att_type="mfn", Deep_GCN_nlayers=2)
    assert tuple(m.smax_fc.weight.shape) == (6, 400)          # code/model.py:991-994: 3 x 100 hidden states + the 100-d memory
    assert [k for k in m.state_dict() if k.startswith("mfn.")][:2] == ["mfn.lstm_l.weight_ih", "mfn.lstm_l.weight_hh"]
    m = mmdfn_b200.DialogueGNNModel("LSTM", 100, 150, 150, 100, 100, 100, 100, 2, 200, 10, 10, graph_type="relation",
                                    n_classes=6)              # defaults: att_type='gated' -> 300-wide classifier input
    assert tuple(m.smax_fc.weight.shape) == (6, 300)          # code/model.py:986-988


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    m = _make(100, 64, 32, 2, 6, 1, "IEMOCAP", "3-0-1")
    T, B = 5, 2
    with pytest.raises(mmdfn_b200.MMDFNError):
        m(torch.zeros(T, B, 100), torch.zeros(T, B, 2), torch.ones(B, T), [5, 5], torch.zeros(T, B, 64), torch.zeros(T, B, 32))
    with pytest.raises(mmdfn_b200.MMDFNError):
        mmdfn_b200.FocalLoss()(torch.zeros(3, 6), torch.zeros(3, dtype=torch.long))


def test_dropin_module_names_import():
    import importlib
    import sys
    d = os.path.join(ROOT, "mm-dfn_b200", "dropin")
    sys.path.insert(0, d)
    try:
        for name in ("model", "model_GCN", "model_mm", "loss"):
            sys.modules.pop(name, None)
            mod = importlib.import_module(name)
            assert mod.__file__.startswith(d)
        import model, loss
        assert model.DialogueGNNModel is mmdfn_b200.DialogueGNNModel
        assert loss.FocalLoss is mmdfn_b200.FocalLoss
        for name in ("LSTMModel", "GRUModel", "DialogRNNModel"):
            assert hasattr(model, name)
    finally:
        sys.path.remove(d)
        for name in ("model", "model_GCN", "model_mm", "loss", "_bootstrap"):
            sys.modules.pop(name, None)


def test_gru_tile_planner_fits_one_wave():
    """ops.plan_gru_tiles: both concurrently running encoders must be co-resident (2 * ceil(n / nb) CTAs each, 148 SMs),
    tiles come from the kernels' instantiations, and hopeless cases fall back to the library default (0, 0)."""
    from mmdfn_b200 import ops
    for T, n_a, n_b in ((100, 32, 192), (100, 31, 186), (24, 16, 432), (110, 32, 192), (30, 1, 6), (100, 8, 48)):
        a, b = ops.plan_gru_tiles(T, n_a, n_b)
        assert a in (2, 3, 4, 8) and b in (2, 3, 4, 8)
        assert 2 * -(-n_a // a) + 2 * -(-n_b // b) <= 148
    assert ops.plan_gru_tiles(100, 32, 192) == (4, 3)          # the bench shard: 16 + 128 CTAs
    assert ops.plan_gru_tiles(500, 64, 1536) == (0, 0)         # BASELINE config 5: more sequences than one wave holds


def test_trainer_parameter_selection_follows_the_configuration():
    """FlatAdamTrainer buckets the parameters the reference's Adam would update for each configuration."""
    from types import SimpleNamespace as NS
    from mmdfn_b200.dp import used_prefixes, USED_PREFIXES
    net = NS(reason_flag=True, convs=[0, 1])
    gdf = NS(graph_type="GDF", att_type="concat_subsequently", use_crn_speaker=True, graph_model=NS(graph_net=net))
    assert sorted(used_prefixes(gdf)) == sorted(USED_PREFIXES)
    gdf.use_crn_speaker = False
    assert "rnn_parties." not in used_prefixes(gdf)
    net.reason_flag = False
    assert "graph_model.graph_net.rnn." not in used_prefixes(gdf)
    net.reason_flag, net.convs = True, []
    assert "graph_model.graph_net.rnn." not in used_prefixes(gdf)
    rel = NS(graph_type="relation", att_type="gated", use_crn_speaker=True)
    pre = used_prefixes(rel)
    assert all(x in pre for x in ("graph_net_a.", "graph_net_v.", "graph_net_l.", "att_model.scalar.", "gatedatt."))
    rel.att_type = "concat_subsequently"
    assert "gatedatt." not in used_prefixes(rel)
