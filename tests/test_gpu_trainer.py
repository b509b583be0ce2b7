"""Trainer-side pieces on the GPU: (1) the sync-free mirror of train_or_eval_graph_model gives the same losses and
metrics as a literal transcription of the reference loop (code/run_train_erc.py:149-238) driving the same model;
(2) the reference's UNCHANGED `code/run_train_erc.py` (staged byte for byte under baseline/_ref/code), started through
tools/run_reference_trainer.py with the drop-in modules AND the drop-in data path, reproduces the epoch losses of the
unmodified reference run on CPU (tests/golden/run_train_erc_small.json, made by tests/golden/make_golden_trainer.py)."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)
from helpers import write_small_iemocap_pickle  # noqa: E402


def _model(dropout):
    import mmdfn_b200
    torch.manual_seed(2021)
    m = mmdfn_b200.DialogueGNNModel("LSTM", 100, 150, 150, 100, 100, 100, 100, n_speakers=2, max_seq_len=200, window_past=10,
                                    window_future=10, n_classes=6, dropout=dropout, graph_type="GDF", alpha=0.2, lamda=0.5,
                                    D_m_v=342, D_m_a=1582, modals="avl", att_type="concat_subsequently", Deep_GCN_nlayers=2,
                                    use_speaker=False, reason_flag=True, use_crn_speaker=True, speaker_weights="3-0-1")
    return m.cuda()


def _literal_reference_loop(model, loss_f, dataloader, train_flag, optimizer, seed):
    """code/run_train_erc.py:149-238 transcribed line by line (cuda_flag=True, modals='avl', concat_subsequently)"""
    from sklearn.metrics import f1_score, accuracy_score
    from mmdfn_b200.trainer import seed_everything
    losses, preds, labels = [], [], []
    model.train() if train_flag else model.eval()
    seed_everything(seed)
    for data in dataloader:
        if train_flag:
            optimizer.zero_grad()
        textf, visuf, acouf, qmask, umask, label = [d.cuda() for d in data[:6]]
        lengths = [(umask[j] == 1).nonzero().tolist()[-1][0] + 1 for j in range(len(umask))]
        log_prob = model(textf, qmask, umask, lengths, acouf, visuf, False)[0]
        label = torch.cat([label[j][:lengths[j]] for j in range(len(label))])
        loss = loss_f(log_prob, label)
        preds.append(torch.argmax(log_prob, 1).cpu().numpy())
        labels.append(label.cpu().numpy())
        losses.append(loss.item())
        if train_flag:
            loss.backward()
            optimizer.step()
    preds, labels = np.concatenate(preds), np.concatenate(labels)
    return (round(np.sum(losses) / len(losses), 4), round(accuracy_score(labels, preds) * 100, 2),
            round(f1_score(labels, preds, average='weighted') * 100, 2), labels, preds)


def test_mirror_loop_equals_literal_reference_loop(tmp_path):
    import mmdfn_b200
    from mmdfn_b200.dataloader import IEMOCAPDataset
    from mmdfn_b200.trainer import train_or_eval_graph_model
    pkl = write_small_iemocap_pickle(str(tmp_path / "small.pkl"))
    if pkl is None:
        pytest.skip("IEMOCAP test dialogues not staged under baseline/_ref/data")
    train, test = IEMOCAPDataset(pkl, True), IEMOCAPDataset(pkl, False)
    results = []
    for which in ("literal", "mirror"):
        model = _model(0.4)
        loss_f = mmdfn_b200.FocalLoss(gamma=1.0)
        opt = torch.optim.Adam(model.parameters(), lr=3e-4, weight_decay=1e-4)
        out = []
        for epoch in range(2):
            for ds, flag in ((train, True), (test, False)):
                loader = torch.utils.data.DataLoader(ds, batch_size=8, collate_fn=ds.collate_fn, shuffle=flag)
                if which == "literal":
                    r = _literal_reference_loop(model, loss_f, loader, flag, opt if flag else None, 2021)
                else:
                    t = train_or_eval_graph_model(model, loss_f, loader, epoch, flag, opt if flag else None, True, "avl",
                                                  ['hap', 'sad', 'neu', 'ang', 'exc', 'fru'])
                    r = (t[2], t[3], t[6], t[4], t[5])
                    assert isinstance(t[0], str) and t[1][0] == "ACC" and len(t[7]) == 5
                out.append(r)
        results.append(out)
    for a, b in zip(*results):
        assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2]                 # same kernels, same seeds: identical numbers
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
    assert results[1][2][0] < results[1][0][0]                                # and it trains: epoch-1 train loss < epoch-0


def test_flat_adam_trainer_in_the_mirror_loop(tmp_path):
    import mmdfn_b200
    from mmdfn_b200.dataloader import IEMOCAPDataset
    from mmdfn_b200.dp import FlatAdamTrainer
    from mmdfn_b200.trainer import train_or_eval_graph_model
    pkl = write_small_iemocap_pickle(str(tmp_path / "small.pkl"))
    if pkl is None:
        pytest.skip("IEMOCAP test dialogues not staged under baseline/_ref/data")
    ds = IEMOCAPDataset(pkl, True)
    model = _model(0.4)
    loss_f = mmdfn_b200.FocalLoss(gamma=1.0)
    tr = FlatAdamTrainer(model, loss_f, lr=3e-4, weight_decay=1e-4)
    batches = ds.length_bucketed_batches(8, shuffle=True, seed=1)
    losses = []
    for epoch in range(3):
        loader = [ds.collate_indices(b) for b in batches]
        losses.append(train_or_eval_graph_model(model, loss_f, loader, epoch, True, tr, True, "avl")[2])
    assert losses[-1] < losses[0]


def test_unchanged_reference_trainer_runs_on_the_dropin_and_matches_the_reference(tmp_path):
    script = os.path.join(ROOT, "baseline", "_ref", "code", "run_train_erc.py")
    pkl = write_small_iemocap_pickle(str(tmp_path / "small.pkl"))
    if pkl is None or not os.path.exists(script):
        pytest.skip("reference trainer / IEMOCAP test dialogues not staged under baseline/_ref")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "run_train_erc_small.json")))
    assert hashlib.sha256(open(script, "rb").read()).hexdigest() == gold["run_train_erc_sha256"]   # byte-for-byte the reference's file
    from make_golden_trainer import parse_epochs
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_trainer.py"), script, "--data_dir", pkl] + gold["args"],
                       capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-3000:]
    assert "Running on GPU" in r.stdout and "MM-DFN with LSTM as base model" in r.stdout
    assert "The model have 1784415 paramerters in total" in r.stdout                # code/run_train_erc.py:493, IEMOCAP K=2
    epochs = parse_epochs(r.stdout)
    assert len(epochs) == len(gold["epochs"]) == 2
    for mine, ref in zip(epochs, gold["epochs"]):
        # fp32 re-association through two epochs of Adam: 1e-3 on the (rounded to 1e-4) mean losses, one utterance in
        # 1623 is 0.06 accuracy points
        assert abs(mine["train_loss"] - ref["train_loss"]) < 1e-3 * max(1.0, abs(ref["train_loss"])), (mine, ref)
        assert abs(mine["test_loss"] - ref["test_loss"]) < 1e-3 * max(1.0, abs(ref["test_loss"])), (mine, ref)
        assert abs(mine["train_acc"] - ref["train_acc"]) <= 0.5 and abs(mine["test_acc"] - ref["test_acc"]) <= 0.5, (mine, ref)
