"""Drop-in counterparts of the reference's streaming-metric accumulators (``morgana/metrics.py:359-694``).

Same class names, ``accumulate(*tensors, seq_len=None)`` / ``result()`` / ``reset_state()`` and the ``sum`` / ``count``
attributes.  What changes is where the arithmetic happens: each ``accumulate`` is ONE kernel launch that reduces only
the valid frames and adds straight into a 48-byte device record owned by the metric (``sum`` and ``count`` are views of
it), so there is no mask tensor, no temporary, and no ``.item()`` sync per metric per step.

``count`` follows the reference's convention: number of valid FRAMES when ``seq_len`` is given, ``numel`` otherwise (Q2).
The host-side bookkeeping classes of the reference (``Handler``, ``Print``, ``History``, ``TensorHistory``) are out of
scope and keep working unchanged on top of these.
"""
import numpy as np
import torch

from morgana_b200 import _lib, ops


class StatefulMetric(object):
    r"""Base class: ``reset_state`` / ``accumulate`` / ``result`` (morgana/metrics.py:9-49)."""
    def __init__(self, hidden=False):
        self._hidden = hidden
        self.hidden = True

    def reset_state(self, *args):
        self.hidden = True

    def accumulate(self, *args, **kwargs):
        self.hidden = self._hidden

    def result(self, *args):
        raise NotImplementedError

    def result_as_json(self, *args):
        tensor = self.result(*args)
        if isinstance(tensor, torch.Tensor):
            tensor = tensor.detach().cpu().numpy().tolist()
        return tensor


def _reset(metric):
    metric.hidden = True
    metric._record = None
    metric._integer = False
    metric.sum = 0.
    metric.count = 0.


def _accumulate(metric, kind, a, b=None, m=None, seq_len=None):
    """Add one batch into the metric's device record with a single launch; refresh the `sum` / `count` views."""
    metric.hidden = metric._hidden
    ops._require_cuda(a, 'metric input')
    if a.dim() != 3:
        raise ValueError('metric inputs must have shape (batch_size, seq_len, feat_dim), got {}'.format(tuple(a.shape)))
    if getattr(metric, '_record', None) is None or metric._record.device != a.device:
        metric._record = ops.new_result_records(1, a.device)[0]
    B, T, _ = a.shape
    if B == 0:
        return
    with ops._device_of(a):
        term = ops.make_term(kind, a, b, m=m, result=metric._record, accumulate=True)
        ops.masked_reduce([term], seq_len, B, T, a.device)
    metric._integer = term[0].ab_dtype == _lib.DT_U8 and kind != _lib.RED_EQ
    if metric._integer:
        metric.sum = metric._record.view(torch.int64)[ops.I64_ISUM]     # the reference's sum is int64 here too
    else:
        metric.sum = metric._record.view(torch.float32)[ops.F32_SUM]
    metric.count = metric._record.view(torch.float64)[ops.F64_COUNT]


def _mean(metric):
    """``sum / (count + 1e-8)`` (morgana/metrics.py:396-397), formed in float64 on the device, returned as float32."""
    record = getattr(metric, '_record', None)
    if record is None:
        return torch.as_tensor(metric.sum / (metric.count + 1e-8), dtype=torch.float32)
    f64 = record.view(torch.float64)
    return (f64[ops.F64_SUM] / (f64[ops.F64_COUNT] + 1e-8)).to(torch.float32)


def _as_float(t):
    return t if t.dtype == torch.float32 or t.dtype in (torch.uint8, torch.bool) else t.to(torch.float32)


class Mean(StatefulMetric):
    r"""Online mean (morgana/metrics.py:359-397)."""
    def __init__(self, hidden=False):
        StatefulMetric.__init__(self, hidden=hidden)
        self.reset_state()

    def reset_state(self):
        _reset(self)

    def accumulate(self, tensor, seq_len=None):
        _accumulate(self, _lib.RED_SUM, _as_float(tensor), seq_len=seq_len)

    def result(self, *args):
        return _mean(self)


class RMSE(Mean):
    r"""Online root-mean-squared error (morgana/metrics.py:474-499)."""
    def accumulate(self, target, pred, seq_len=None):
        _accumulate(self, _lib.RED_SQDIFF, target, pred, seq_len=seq_len)

    def result(self, *args):
        return _mean(self) ** 0.5


class MAE(Mean):
    r"""Online mean-absolute error (morgana/metrics.py:556-576)."""
    def accumulate(self, target, pred, seq_len=None):
        _accumulate(self, _lib.RED_ABSDIFF, target, pred, seq_len=seq_len)


class Accuracy(Mean):
    r"""Percentage of frames where ``target & pred`` (morgana/metrics.py:502-526)."""
    def accumulate(self, target, pred, seq_len=None):
        _accumulate(self, _lib.RED_AND, target, pred, seq_len=seq_len)

    def result(self, *args):
        return _mean(self) * 100.


class Error(Mean):
    r"""Percentage of frames where ``target ^ pred`` (morgana/metrics.py:529-553)."""
    def accumulate(self, target, pred, seq_len=None):
        _accumulate(self, _lib.RED_XOR, target, pred, seq_len=seq_len)

    def result(self, *args):
        return _mean(self) * 100.


class F0Distortion(RMSE):
    r"""F0 RMSE over frames that are voiced and inside the utterance (morgana/metrics.py:579-609).

    Unlike the reference, ``is_voiced`` is never modified in place.
    """
    def accumulate(self, f0_target, f0_pred, is_voiced, seq_len=None):
        _accumulate(self, _lib.RED_SQDIFF, f0_target, f0_pred, m=_as_float(is_voiced), seq_len=seq_len)


class LF0Distortion(F0Distortion):
    r"""F0 RMSE in Hz from log-F0 streams; the ``exp`` is fused into the reduction (morgana/metrics.py:612-634)."""
    def accumulate(self, lf0_target, lf0_pred, is_voiced, seq_len=None):
        _accumulate(self, _lib.RED_SQDIFF_EXP, lf0_target, lf0_pred, m=_as_float(is_voiced), seq_len=seq_len)


class Distortion(Mean):
    r"""Mean per-frame Euclidean distance, in dB (morgana/metrics.py:637-669)."""
    log_spec_dB_const = 10. / np.log(10.) * np.sqrt(2.)

    def accumulate(self, target, pred, seq_len=None):
        _accumulate(self, _lib.RED_ROOT_SQDIFF, target, pred, seq_len=seq_len)

    def result(self, *args):
        return _mean(self) * self.log_spec_dB_const


class MelCepDistortion(RMSE):
    r"""RMSE ignoring coefficient 0 -- a strided view, not a copy (morgana/metrics.py:672-694)."""
    def accumulate(self, target, pred, seq_len=None):
        _accumulate(self, _lib.RED_SQDIFF, target[..., 1:], pred[..., 1:], seq_len=seq_len)


ACCUMULATORS = (Mean, RMSE, MAE, Accuracy, Error, F0Distortion, LF0Distortion, Distortion, MelCepDistortion)
