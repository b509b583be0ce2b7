"""BASELINE config 5 on one GPU shard (64 dialogues x 500 utterances, 8 speakers, 6 GCN layers, 100/512/1024-d): a few
eager training steps timed with CUDA events, then the graph-conv aggregate's roofline sweep over the dialogue length
(the L > 128 lengths take the FFMA kernel, the others the tcgen05 kernel).  One JSON object on stdout.  Not a bench line."""
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench as B
import mmdfn_b200
from mmdfn_b200.dp import FlatAdamTrainer

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
out = {}


def batch(lengths, S, C, seed):
    rs = np.random.RandomState(seed)
    nb, T = len(lengths), int(max(lengths))
    feats = [torch.from_numpy(rs.standard_normal((T, nb, d)).astype(np.float32)) for d in (B.D_T, B.D_A, B.D_V)]
    spk = rs.randint(0, S, size=(T, nb))
    q = np.zeros((T, nb, S), np.float32)
    q[np.arange(T)[:, None], np.arange(nb)[None, :], spk] = 1
    u = np.ones((nb, T), np.float32)
    lab = torch.from_numpy(rs.randint(0, C, size=sum(lengths)).astype(np.int64))
    return feats[0], feats[1], feats[2], torch.from_numpy(q), torch.from_numpy(u), lab


try:
    S, C, K, NB, L = 8, 6, 6, 64, 500
    lengths = [L] * NB
    torch.manual_seed(2021)
    model = mmdfn_b200.DialogueGNNModel(
        "LSTM", B.D_T, 150, 150, 100, 100, 100, 100, n_speakers=S, max_seq_len=L, window_past=10, window_future=10,
        n_classes=C, dropout=B.DROPOUT, graph_type="GDF", alpha=0.2, lamda=0.5, D_m_v=B.D_V, D_m_a=B.D_A, modals="avl",
        att_type="concat_subsequently", Deep_GCN_nlayers=K, use_speaker=False, reason_flag=True, use_crn_speaker=True,
        speaker_weights="1-1-1").to(dev).train()
    trainer = FlatAdamTrainer(model, mmdfn_b200.FocalLoss(gamma=1.0), lr=B.LR, weight_decay=B.L2)
    t, a, v, q, u, lab = (x.to(dev) for x in batch(lengths, S, C, 7))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss0 = float(trainer.step(t, q, u, lengths, a, v, lab))
    out["first_step_s"] = time.perf_counter() - t0
    for _ in range(2):
        trainer.step(t, q, u, lengths, a, v, lab)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    e0.record()
    for _ in range(steps):
        loss = trainer.step(t, q, u, lengths, a, v, lab)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out["config5_shard"] = {"dialogues": NB, "utterances_per_dialogue": L, "speakers": S, "gcn_layers": K,
                            "ms_per_step": ms, "utterances_per_s": NB * L / (ms * 1e-3), "loss_first": loss0,
                            "loss_last": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    del trainer, model, t, a, v, q, u, lab
    torch.cuda.empty_cache()
except Exception as e:  # keep going: the sweep below is independent
    out["config5_shard"] = {"error": repr(e)}

sweep = []
for L, nd in ((16, 2048), (50, 640), (100, 256), (128, 160), (200, 96), (500, 24)):
    try:
        B.UTT = L
        r = B.roofline_graph_conv(dev, nd)
        sweep.append({"L": L, "dialogues": nd, "us_per_launch": r["us_per_launch"], "achieved_gbs": r["achieved"], "frac": r["frac"],
                      "algorithmic_bytes_per_launch": r["algorithmic_bytes_per_launch"],
                      "copy_kernel_frac": r["same_bytes_copy_kernel"]["frac"]})
    except Exception as e:
        sweep.append({"L": L, "dialogues": nd, "error": repr(e)})
out["aggregate_roofline_sweep"] = sweep
print(json.dumps(out))
