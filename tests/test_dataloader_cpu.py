"""The drop-in data path (mm-dfn_b200/dataloader.py) against the reference's own `code/dataloader.py` (imported from the
staged, unmodified copy under baseline/_ref/code or from /root/reference) on pickles in the author's two formats:
items, collate output (bit-exact), key order, lengths, packed labels, bucketed batches, rank shards.  CPU only."""
import importlib.util
import os
import pickle
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_dataloader():
    for base in (os.path.join(ROOT, "baseline", "_ref", "code"), "/root/reference/code"):
        p = os.path.join(base, "dataloader.py")
        if os.path.exists(p):
            spec = importlib.util.spec_from_file_location("ref_dataloader_unmodified", p)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    return None


def _write_pickles(tmp):
    rs = np.random.RandomState(0)
    lens = [7, 1, 12, 5, 9, 3, 12, 4]
    vids_i = ["Ses%02d" % i for i in range(len(lens))]
    vids_m = list(range(100, 100 + len(lens)))

    def feats(vids, d):
        return {v: [rs.standard_normal(d) for _ in range(L)] for v, L in zip(vids, lens)}     # lists of float64 rows, like the author's files

    ie = ({v: list(range(L)) for v, L in zip(vids_i, lens)}, {v: [rs.choice(["M", "F"]) for _ in range(L)] for v, L in zip(vids_i, lens)},
          {v: [int(x) for x in rs.randint(0, 6, L)] for v, L in zip(vids_i, lens)}, feats(vids_i, 10), feats(vids_i, 16), feats(vids_i, 12),
          {v: [""] * L for v, L in zip(vids_i, lens)}, vids_i[:5], vids_i[5:])
    spk_m = {v: [list(np.eye(9)[rs.randint(0, 9)]) for _ in range(L)] for v, L in zip(vids_m, lens)}
    me = ({v: list(range(L)) for v, L in zip(vids_m, lens)}, spk_m, {v: [int(x) for x in rs.randint(0, 7, L)] for v, L in zip(vids_m, lens)},
          feats(vids_m, 6), feats(vids_m, 5), feats(vids_m, 4), {v: [""] * L for v, L in zip(vids_m, lens)}, set(vids_m[:5]), set(vids_m[5:]), None)
    pi, pm = os.path.join(tmp, "ie.pkl"), os.path.join(tmp, "me.pkl")
    pickle.dump(ie, open(pi, "wb"))
    pickle.dump(me, open(pm, "wb"))
    return pi, pm


@pytest.mark.parametrize("which", ["IEMOCAP", "MELD"])
@pytest.mark.parametrize("train", [True, False])
def test_dropin_dataset_matches_reference(tmp_path, which, train):
    ref = _ref_dataloader()
    if ref is None:
        pytest.skip("reference dataloader.py not staged")
    sys.path.insert(0, ROOT)
    from mmdfn_b200 import dataloader as mine
    pi, pm = _write_pickles(str(tmp_path))
    path = pi if which == "IEMOCAP" else pm
    R = getattr(ref, which + "Dataset")(path, train)
    M = getattr(mine, which + "Dataset")(path, train)
    M.pin_batches = False
    assert len(R) == len(M) and R.keys == M.keys
    for i in range(len(R)):
        r, m = R[i], M[i]
        assert r[6] == m[6]
        for a, b in zip(r[:6], m[:6]):
            assert a.dtype == b.dtype and torch.equal(a, b)
    order = list(range(len(R)))[::-1]
    rb = R.collate_fn([R[i] for i in order])
    mb = M.collate_fn([M[i] for i in order])
    assert len(rb) == len(mb) == 7 and rb[6] == mb[6]
    for a, b in zip(rb[:6], mb[:6]):
        assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)          # bit-exact collate
    # what the reference trainer recomputes per batch (code/run_train_erc.py:194,201)
    umask, label = rb[4], rb[5]
    lengths = [(umask[j] == 1).nonzero().tolist()[-1][0] + 1 for j in range(len(umask))]
    assert mb.lengths == lengths
    assert torch.equal(mb.label_packed, torch.cat([label[j][:lengths[j]] for j in range(len(label))]))
    if which == "MELD":
        assert R.return_labels() == M.return_labels()


def test_length_buckets_and_shards(tmp_path):
    sys.path.insert(0, ROOT)
    from mmdfn_b200 import dataloader as mine
    pi, _ = _write_pickles(str(tmp_path))
    D = mine.IEMOCAPDataset(pi, True)
    for shuffle in (False, True):
        batches = D.length_bucketed_batches(2, shuffle=shuffle, seed=3, bucket_mult=2)
        assert sorted(i for b in batches for i in b) == list(range(len(D)))              # every dialogue exactly once
        assert all(1 <= len(b) <= 2 for b in batches)
    plain = D.length_bucketed_batches(2)
    assert all(abs(D.lengths[b[0]] - D.lengths[b[-1]]) <= 4 for b in plain)              # similar lengths share a batch
    idx = list(range(7))
    parts = [D.shard(idx, r, 3) for r in range(3)]
    assert sum(parts, []) == idx and max(map(len, parts)) - min(map(len, parts)) <= 1
    with pytest.raises(NotImplementedError):
        mine.DailyDialogueDataset("train", "x")


def test_scores_from_confusion_match_sklearn():
    sys.path.insert(0, ROOT)
    from mmdfn_b200.trainer import scores_from_confusion
    sk = pytest.importorskip("sklearn.metrics")
    rs = np.random.RandomState(1)
    for C in (2, 6, 7):
        y, p = rs.randint(0, C, 500), rs.randint(0, C, 500)
        p[::3] = y[::3]
        conf = np.zeros((C, C), np.int64)
        np.add.at(conf, (y, p), 1)
        acc, f1 = scores_from_confusion(conf)
        assert acc == round(sk.accuracy_score(y, p) * 100, 2)
        assert f1 == round(sk.f1_score(y, p, average="weighted") * 100, 2)
