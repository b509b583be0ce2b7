"""Trainer-side step pieces of the MM-DFN hot path: a mirror of the reference's
``train_or_eval_graph_model`` (code/run_train_erc.py:149-238) without its per-batch host synchronisations.

Same signature, same semantics, same return tuple:
  * ``model.train()`` / ``model.eval()``, then ``seed_everything(seed)`` at the start of EVERY call (:164) -- so, like the
    reference, every epoch replays the same shuffle and the same dropout stream (the kernels' mask counter is reset too);
  * per batch: ``optimizer.zero_grad()``, H2D, ``model(textf, qmask, umask, lengths, acouf, visuf, test_label)``,
    ``loss_f(log_prob, label)``, ``loss.backward()``, ``optimizer.step()``;
  * returns ``(all_each, all_acc, avg_loss, avg_accuracy, labels, preds, avg_fscore, [vids, ei, et, en, el])``.

What is different underneath:
  * ``lengths`` come from the host copy of ``umask`` BEFORE the H2D copy (or from ``Batch.lengths`` of the drop-in
    dataloader) -- the reference runs ``(umask[j] == 1).nonzero().tolist()`` on the device, one sync per dialogue (:194);
  * the ragged label vector is packed on the host (``Batch.label_packed``) instead of ``torch.cat`` of device slices (:201);
  * per-batch ``loss.item()`` / ``argmax(...).cpu()`` (:202-205) are gone: losses stay on the device until the epoch ends,
    predictions and the confusion matrix accumulate on the device (``mmdfn_confusion_accumulate``), ONE D2H at the end;
  * accuracy and weighted F1 are computed from the C x C confusion counts (same numbers as sklearn's
    ``accuracy_score`` / ``f1_score(average='weighted')``); sklearn is used for the text report only, if installed;
  * H2D copies are ``non_blocking`` from the pinned batches of ``mmdfn_b200.dataloader``.
``optimizer`` may be any ``torch.optim`` optimizer or a ``mmdfn_b200.dp.FlatAdamTrainer`` (flat bucket, one all-reduce,
fused Adam)."""
import random

import numpy as np
import torch

from . import ops
from ._lib import call, ptr, stream


def seed_everything(seed=2021):
    """code/run_train_erc.py:19-26"""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    ops._mask_counter[0] = 0              # the kernels' dropout stream restarts with the seed, like torch's generator


def scores_from_confusion(conf):
    """(accuracy %, weighted F1 %) from a (C, C) count matrix conf[target, pred]; rounded like the reference (:229-230)."""
    conf = np.asarray(conf, dtype=np.float64)
    total = conf.sum()
    if total == 0:
        return float('nan'), float('nan')
    tp = np.diag(conf)
    support, predicted = conf.sum(1), conf.sum(0)
    with np.errstate(divide='ignore', invalid='ignore'):
        prec = np.where(predicted > 0, tp / predicted, 0.0)
        rec = np.where(support > 0, tp / support, 0.0)
        f1 = np.where(prec + rec > 0, 2 * prec * rec / (prec + rec), 0.0)
    return round(float(tp.sum() / total) * 100, 2), round(float((f1 * support).sum() / total) * 100, 2)


def _host_lengths(data, umask):
    lengths = getattr(data, "lengths", None)
    if lengths is not None:
        return [int(x) for x in lengths]
    um = umask if not umask.is_cuda else umask.cpu()
    # (umask[j] == 1).nonzero()[-1] + 1 of the reference: index of the last real utterance + 1
    return [int((um[j] == 1).nonzero()[-1, 0]) + 1 for j in range(um.shape[0])]


def train_or_eval_graph_model(model, loss_f, dataloader, epoch=0, train_flag=False, optimizer=None, cuda_flag=False,
                              modals=None, target_names=None, test_label=False, tensorboard=False, seed=2021,
                              n_classes=None):
    assert not train_flag or optimizer is not None
    if train_flag:
        model.train()
    else:
        model.eval()
    seed_everything(seed)
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise ops.MMDFNError("train_or_eval_graph_model drives the B200 path: the model must live on a CUDA device")
    flat = optimizer if (optimizer is not None and hasattr(optimizer, "flat_g")) else None
    losses, preds, labels, vids = [], [], [], []
    conf = None
    last = None
    for data in dataloader:
        last = data
        textf, visuf, acouf, qmask, umask, label = data[:6]
        lengths = _host_lengths(data, umask)
        packed = getattr(data, "label_packed", None)
        if packed is None:
            lab_h = label if not label.is_cuda else label.cpu()
            packed = torch.cat([lab_h[j][:lengths[j]] for j in range(len(lengths))])
        textf, visuf, acouf, qmask, umask = (t.to(dev, non_blocking=True) for t in (textf, visuf, acouf, qmask, umask))
        packed = packed.to(dev, non_blocking=True)
        if train_flag and flat is not None:
            loss = flat.step(textf, qmask, umask, lengths, acouf, visuf, packed)
            log_prob = flat.last_log_prob
        else:
            if train_flag:
                optimizer.zero_grad()
            log_prob = model(textf, qmask, umask, lengths, acouf, visuf, test_label)[0]
            loss = loss_f(log_prob, packed)
            if train_flag:
                loss.backward()
                optimizer.step()
        N, C = log_prob.shape
        if conf is None:
            conf = torch.zeros(C * C, dtype=torch.int64, device=dev)
        pred = torch.empty(N, dtype=torch.int64, device=dev)
        call("mmdfn_confusion_accumulate", N, C, ptr(log_prob.detach().contiguous()), ptr(packed, torch.int64),
             ptr(pred, torch.int64), ptr(conf, torch.int64), stream())
        preds.append(pred)
        labels.append(packed)
        losses.append(loss.detach().reshape(1))
    if not preds:
        return [], [], float('nan'), float('nan'), [], [], float('nan'), []
    # the epoch's single device -> host transfer
    preds = torch.cat(preds).cpu().numpy()
    labels = torch.cat(labels).cpu().numpy()
    loss_h = torch.cat(losses).cpu().numpy().astype(np.float64)
    C = int(round(conf.numel() ** 0.5))
    conf_h = conf.cpu().numpy().reshape(C, C)
    vids += last[6] if len(last) > 6 else []
    avg_loss = round(float(np.sum(loss_h) / len(loss_h)), 4)
    avg_accuracy, avg_fscore = scores_from_confusion(conf_h)
    names = list(target_names) if target_names is not None else [str(i) for i in range(C)]
    try:
        from sklearn import metrics
        all_each = metrics.classification_report(labels, preds, labels=list(range(len(names))), target_names=names, digits=4,
                                                 zero_division=0)
    except Exception:  # pragma: no cover  (sklearn absent: the numbers above do not depend on it)
        all_each = "accuracy {:.4f}".format(avg_accuracy / 100)
    all_acc = ["ACC"]
    for i in range(len(names)):
        sup = conf_h[i].sum() if i < C else 0
        all_acc.append("{}: {:.4f}".format(names[i], float(conf_h[i, i] / sup) if sup > 0 else float('nan')))
    empty = np.empty(0)
    return all_each, all_acc, avg_loss, avg_accuracy, labels, preds, avg_fscore, [np.array(vids), empty, empty, empty, np.array([])]
