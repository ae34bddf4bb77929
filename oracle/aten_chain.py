"""CPU baseline: the reference's algorithm for the path, op for op, on torch CPU tensors.  TEST INFRASTRUCTURE ONLY.

The reference is pure Python over ATen + NumPy and cannot travel to the GPU box, so its CPU cost is measured with this
restatement: the same sequence of library calls the reference issues (same temporaries, same host loop, same
reductions), multi-threaded by ATen over all host cores exactly as the reference would be.  It is what
``bench.py --impl reference`` and the ``cpu_baseline`` leg time; ``tests/test_oracle_golden.py`` pins it to the
fixtures produced by the unmodified reference.  Nothing under ``morgana_b200/`` imports this file.

Each function cites the reference lines (relative to the reference root) whose op sequence it restates.
"""
import numpy as np
import torch
import torch.nn.functional as F


def upsample_chain(features, durations):
    """morgana/utils.py:193-228: append a zero item, build a (B, T) index table on the host, gather.

    Works on CPU tensors (the timed CPU baseline) and, like the reference, on CUDA tensors (the "stock ATen on the same
    GPU" bar): the index table is always built on the host and uploaded (utils.py:218-222).
    """
    device = features.device
    n_utts, n_items, dim = features.shape
    totals = durations.sum(dim=1)                                    # :198
    longest = int(totals.max())                                      # :199 (a sync on a GPU)
    per_item = durations.reshape(n_utts, -1).cpu()                   # :202, :218 (device -> host)
    totals_host = totals.cpu()
    zero_item = torch.zeros(n_utts, 1, dim, dtype=features.dtype).to(device)   # :206
    extended = torch.cat([features, zero_item], dim=1)               # :207 full copy of the input
    utt_index = torch.arange(n_utts)[:, None].repeat(1, longest)     # :210-211 (stays on the CPU, Q8)
    item_index = np.full((n_utts, longest), -1, dtype=np.int64)      # :214 (-1 -> the zero item)
    positions = np.arange(n_items)
    for u in range(n_utts):                                          # :218-220 Python loop over the batch
        item_index[u, :int(totals_host[u])] = np.repeat(positions, per_item[u].numpy())
    item_index = torch.tensor(item_index).to(device)                 # :222 (host -> device)
    return extended[utt_index, item_index]                           # :226 advanced-index gather


def lengths_mask(lengths, longest, dtype):
    """morgana/utils.py:134-144."""
    steps = torch.arange(longest).type(lengths.dtype).to(lengths.device)
    return (steps[None, :] < lengths[:, None])[:, :, None].type(dtype)


def normalise_minmax_chain(x, lo, hi):
    """morgana/data.py:579-583 (rebuilds `scale` every call)."""
    span = hi - lo
    span[abs(span) <= 1e-8] = 1.
    return (x - lo[..., None, :]) / span[..., None, :]


def normalise_mvn_chain(x, mean, std):
    """morgana/data.py:533-534."""
    return (x - mean[..., None, :]) / (std[..., None, :] + 1e-8)


def denormalise_mvn_chain(x, mean, std):
    """morgana/data.py:537-538."""
    return (x * std[..., None, :]) + mean[..., None, :]


def _sequence_loss(pointwise, lengths):
    """morgana/losses.py:29-44: mask, per-utterance normalisation, mean over (batch, feature)."""
    if lengths is None:
        per_utt = pointwise.sum(dim=1) / pointwise.shape[1]
    else:
        mask = lengths_mask(lengths, pointwise.shape[1], pointwise.dtype)
        per_utt = (pointwise * mask).sum(dim=1) / mask.sum(dim=1)
    return per_utt.mean()


def mse_chain(pred, target, lengths=None):
    """morgana/losses.py:49-51."""
    return _sequence_loss(F.mse_loss(pred, target, reduction='none'), lengths)


def bce_chain(prob, label, lengths=None):
    """morgana/losses.py:54-56."""
    return _sequence_loss(F.binary_cross_entropy(prob, label, reduction='none'), lengths)


def ce_chain(logits, classes, lengths=None):
    """morgana/losses.py:59-61."""
    return _sequence_loss(F.cross_entropy(logits.transpose(1, 2), classes, reduction='none').unsqueeze(dim=-1), lengths)


def _mean_increment(values, lengths):
    """metrics.Mean.accumulate, morgana/metrics.py:383-394 (with its .item() on the count)."""
    if lengths is None:
        return torch.sum(values), float(values.numel())
    mask = lengths_mask(lengths, values.shape[1], values.dtype)
    return torch.sum(values * mask), torch.sum(mask).item()


def rmse_increment(target, pred, lengths=None):
    """metrics.RMSE.accumulate, morgana/metrics.py:492-495."""
    return _mean_increment((target - pred) ** 2, lengths)


def melcep_increment(target, pred, lengths=None):
    """metrics.MelCepDistortion.accumulate, morgana/metrics.py:690-694."""
    return rmse_increment(target[..., 1:], pred[..., 1:], lengths)


def distortion_increment(target, pred, lengths=None):
    """metrics.Distortion.accumulate, morgana/metrics.py:657-665."""
    sq = torch.sum((target - pred) ** 2, keepdim=True, dim=-1)
    return _mean_increment(torch.sqrt(sq), lengths)


def lf0_increment(lf0_target, lf0_pred, voiced, lengths=None):
    """metrics.LF0Distortion.accumulate -> F0Distortion.accumulate, morgana/metrics.py:597-609, 630-634."""
    f0_target, f0_pred = torch.exp(lf0_target), torch.exp(lf0_pred)
    weight = voiced.type(f0_target.dtype)
    if lengths is not None:
        weight = weight * lengths_mask(lengths, f0_target.shape[1], f0_target.dtype)
    sq = (f0_target - f0_pred) ** 2
    return torch.sum(sq * weight), torch.sum(weight).item()


def ema_chain(shadows, params, decay):
    """utils.ExponentialMovingAverage.update_params, morgana/utils.py:443-456: two ATen ops + a temporary per tensor."""
    for s, x in zip(shadows, params):
        delta = s - x
        s -= (1.0 - decay) * delta


def acoustic_loss_and_metrics(pred, target, voiced_target, lengths, with_grad=True):
    """What LSTMAcousticModel.loss does per batch on the 187-dim layout lf0[0:3] | vuv[3] | mcep[4:184] | bap[184:187]
    (models/RNN_SPSS.py:120-139): four metric accumulations, then three mse terms and one bce term, averaged."""
    pred = pred.detach().requires_grad_(with_grad)
    vuv_prob = pred[..., 3:4]
    vuv = vuv_prob > 0.5
    increments = [
        lf0_increment(target[..., 0:1], pred[..., 0:1].detach(), vuv, lengths),
        _mean_increment((voiced_target == vuv).type(torch.float), lengths),
        melcep_increment(target[..., 4:64], pred[..., 4:64].detach(), lengths),
        distortion_increment(target[..., 184:185], pred[..., 184:185].detach(), lengths),
    ]
    loss = mse_chain(pred[..., 0:3], target[..., 0:3], lengths)
    loss = loss + mse_chain(pred[..., 4:184], target[..., 4:184], lengths)
    loss = loss + mse_chain(pred[..., 184:187], target[..., 184:187], lengths)
    loss = loss + bce_chain(vuv_prob, voiced_target.type(torch.float), lengths)
    loss = loss / 4.
    grad = None
    if with_grad:
        loss.backward()
        grad = pred.grad
    return loss.detach(), grad, increments
