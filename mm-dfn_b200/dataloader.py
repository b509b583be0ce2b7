"""Data path of the MM-DFN hot path: drop-in for the reference's ``code/dataloader.py``.

``IEMOCAPDataset`` / ``MELDDataset`` keep the reference's constructor, ``__getitem__`` / ``__len__`` / ``collate_fn`` /
``return_labels`` contract (code/dataloader.py:9-68) -- the unchanged ``code/run_train_erc.py`` builds its
``DataLoader``s from them -- but the work is organised for a GPU consumer:

* the author's pickle (a 9- / 10-tuple of dicts keyed by dialogue id) is read ONCE and every dialogue is converted ONCE
  into contiguous fp32 / int64 arrays (the reference rebuilds ``torch.FloatTensor(list of arrays)`` on every
  ``__getitem__``: 28 % of its epoch time on IEMOCAP);
* ``collate_fn`` writes the padded time-major batch straight into freshly allocated (pinned, when a CUDA device is
  present) tensors -- no pandas ``DataFrame``, no ``pad_sequence`` -- and returns exactly the reference's list
  ``[textf (T,B,Dt), visuf (T,B,Dv), acouf (T,B,Da), qmask (T,B,S), umask (B,T), label (B,T), vids]`` (bit-identical
  values); the result is a ``Batch`` list that also carries what the reference trainer recomputes on the device with a
  sync per batch: ``lengths`` (host ints) and ``label_packed`` (the ragged ``torch.cat(label[j][:L_j])``);
* ``length_bucketed_batches`` groups dialogues of similar length (less padding: every dialogue's cost depends on the
  batch's max length T) and ``shard`` hands each data-parallel rank its contiguous slice.

This module is host-side logic only (numpy / torch CPU tensors): it has no device arithmetic and therefore nothing to
fall back from."""
import pickle

import numpy as np
import torch
from torch.utils.data import Dataset


class Batch(list):
    """The reference's collate output (a plain list of 7) plus host-side ragged metadata."""
    lengths = None          # list[int], L_j of every dialogue (code/run_train_erc.py:194 recomputes it on the device)
    label_packed = None     # int64 (N,) = torch.cat([label[j][:L_j]])   (code/run_train_erc.py:201)

    def pin(self):
        if torch.cuda.is_available():
            for i in range(6):
                self[i] = self[i].pin_memory()
            self.label_packed = self.label_packed.pin_memory()
        return self


def _alloc(shape, dtype, pinned):
    return torch.zeros(shape, dtype=dtype, pin_memory=bool(pinned and torch.cuda.is_available()))


class _DialogueDataset(Dataset):
    """Shared implementation: `self.rows[vid]` = (text, visual, audio, speakers one-hot, labels) as contiguous arrays."""

    pin_batches = True

    def _ingest(self, speakers_one_hot):
        self.len = len(self.keys)
        self.rows = {}
        for vid in self.keys:
            lab = np.asarray(self.videoLabels[vid], dtype=np.int64)
            self.rows[vid] = (np.ascontiguousarray(np.asarray(self.videoText[vid], dtype=np.float32)),
                              np.ascontiguousarray(np.asarray(self.videoVisual[vid], dtype=np.float32)),
                              np.ascontiguousarray(np.asarray(self.videoAudio[vid], dtype=np.float32)),
                              np.ascontiguousarray(speakers_one_hot(vid)), lab)
        self.lengths = [len(self.rows[v][4]) for v in self.keys]

    def __getitem__(self, index):
        vid = self.keys[index]
        t, v, a, q, lab = self.rows[vid]
        return (torch.from_numpy(t), torch.from_numpy(v), torch.from_numpy(a), torch.from_numpy(q),
                torch.ones(len(lab), dtype=torch.float32), torch.from_numpy(lab), vid)

    def __len__(self):
        return self.len

    def collate_fn(self, data):
        """data: list of __getitem__ tuples -> the reference's 7-list (code/dataloader.py:31-34), as a `Batch`."""
        B = len(data)
        lengths = [int(d[5].shape[0]) for d in data]
        T = max(lengths) if lengths else 0
        pinned = self.pin_batches
        out = Batch()
        for i in range(4):                                   # text, visual, audio, qmask: time-major (T, B, D)
            D = int(data[0][i].shape[1])
            buf = _alloc((T, B, D), torch.float32, pinned)
            for b, d in enumerate(data):
                buf[:lengths[b], b] = d[i]
            out.append(buf)
        umask = _alloc((B, T), torch.float32, pinned)
        label = _alloc((B, T), torch.int64, pinned)
        for b, d in enumerate(data):
            umask[b, :lengths[b]] = 1.0
            label[b, :lengths[b]] = d[5]
        out.append(umask)
        out.append(label)
        out.append([d[6] for d in data])
        out.lengths = lengths
        packed = _alloc((sum(lengths),), torch.int64, pinned)
        pos = 0
        for b, d in enumerate(data):
            packed[pos:pos + lengths[b]] = d[5]
            pos += lengths[b]
        out.label_packed = packed
        return out

    def collate_indices(self, indices):
        return self.collate_fn([self[i] for i in indices])

    def length_bucketed_batches(self, batch_size, shuffle=False, seed=0, bucket_mult=8):
        """Lists of dataset indices: dialogues are sorted by length inside windows of `bucket_mult * batch_size`
        (shuffled first when `shuffle`), then cut into batches -- batches hold similar lengths, so the padded length T
        (which every dialogue of the batch pays for) stays close to the real lengths.  Covers every index exactly once."""
        idx = np.arange(self.len)
        rs = np.random.RandomState(seed)
        if shuffle:
            rs.shuffle(idx)
        win = max(1, bucket_mult) * batch_size
        batches = []
        for s in range(0, self.len, win):
            w = sorted(idx[s:s + win].tolist(), key=lambda i: (self.lengths[i], i))
            batches += [w[j:j + batch_size] for j in range(0, len(w), batch_size)]
        if shuffle:
            rs.shuffle(batches)
        return batches

    @staticmethod
    def shard(indices, rank, world):
        """contiguous slice of one batch's dialogue indices for data-parallel rank `rank` (dialogues are independent)"""
        per, rem = divmod(len(indices), world)
        lo = rank * per + min(rank, rem)
        return indices[lo:lo + per + (1 if rank < rem else 0)]


class IEMOCAPDataset(_DialogueDataset):
    """code/dataloader.py:9-34.  Pickle: (videoIDs, videoSpeakers 'M'/'F', videoLabels, videoText, videoAudio,
    videoVisual, videoSentence, trainVid, testVid); 'M' -> [1, 0], else [0, 1]."""

    def __init__(self, path=None, train=True):
        self.videoIDs, self.videoSpeakers, self.videoLabels, self.videoText, \
            self.videoAudio, self.videoVisual, self.videoSentence, self.trainVid, \
            self.testVid = pickle.load(open(path, 'rb'), encoding='latin1')
        self.keys = [x for x in (self.trainVid if train else self.testVid)]
        self._ingest(lambda vid: np.array([[1, 0] if x == 'M' else [0, 1] for x in self.videoSpeakers[vid]],
                                          dtype=np.float32).reshape(-1, 2))


class MELDDataset(_DialogueDataset):
    """code/dataloader.py:37-68.  Pickle has a tenth entry; videoSpeakers already holds one-hot rows (9 speakers)."""

    def __init__(self, path=None, train=True):
        self.videoIDs, self.videoSpeakers, self.videoLabels, self.videoText, \
            self.videoAudio, self.videoVisual, self.videoSentence, self.trainVid, \
            self.testVid, self.aaa = pickle.load(open(path, 'rb'), encoding='latin1')
        self.keys = [x for x in (self.trainVid if train else self.testVid)]
        self._ingest(lambda vid: np.asarray(self.videoSpeakers[vid], dtype=np.float32))

    def return_labels(self):
        return_label = []
        for key in self.keys:
            return_label += list(self.videoLabels[key])
        return return_label


class DailyDialogueDataset(Dataset):
    """code/dataloader.py:71-110 feeds the DailyDialogue baselines, which are outside the MM-DFN hot path."""

    def __init__(self, *a, **k):
        raise NotImplementedError("DailyDialogueDataset belongs to baselines outside the MM-DFN hot path (SURVEY.md section 2)")
