"""Drop-in counterparts of the reference's streaming-metric accumulators (``morgana/metrics.py:359-694``).

Same class names, ``accumulate(*tensors, seq_len=None)`` / ``result()`` / ``reset_state()`` and the ``sum`` / ``count``
attributes.  What changes is where the arithmetic happens: each ``accumulate`` is ONE kernel launch that reduces only
the valid frames and adds straight into a 48-byte device record owned by the metric (``sum`` and ``count`` are views of
it), so there is no mask tensor, no temporary, and no ``.item()`` sync per metric per step.

``count`` follows the reference's convention: number of valid FRAMES when ``seq_len`` is given, ``numel`` otherwise (Q2).
The container and bookkeeping classes (``Handler``, ``Print``, ``History``, ``TensorHistory``, morgana/metrics.py:52-356)
are here too, so ``self.metrics.accumulate(self.mode, LF0_RMSE_Hz=(...), ...)`` (models/RNN_SPSS.py:124-129) and
``metrics.accumulate(mode, loss=batch_loss)`` (experiment_builder.py:484) run as written.
"""
from collections.abc import Iterable

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

    def __str__(self):
        return _format_float(self.result())


def _format_float(value):
    r"""``utils.format_float_tensor`` of the reference (morgana/utils.py:17-34): ``tqdm.format_num`` of a scalar, of up to
    four values, or of the first two and the last value of a longer vector."""
    from tqdm import tqdm
    if isinstance(value, torch.Tensor):
        value = value.detach().cpu()
    try:
        n = len(value)
    except TypeError:
        n = 0
    if n <= 1:
        return tqdm.format_num(value)
    if n <= 4:
        return '[{}]'.format(', '.join(tqdm.format_num(v) for v in value))
    return '[{}, {}, ..., {}]'.format(tqdm.format_num(value[0]), tqdm.format_num(value[1]), tqdm.format_num(value[-1]))


def _listify(obj):
    r"""``utils.listify`` (morgana/utils.py:10-14): tuples / lists pass through, anything else is wrapped."""
    if isinstance(obj, (list, tuple)):
        return list(obj)
    return [obj]


def _names(collections):
    """One collection name or several -> a list of names."""
    if isinstance(collections, str) or not isinstance(collections, Iterable):
        return [collections]
    return list(collections)


class Handler(StatefulMetric):
    r"""Named collections of metrics -- 'all', 'train', 'valid', 'test' and any added later -- driven as
    ``handler.accumulate(mode, name=(tensors..., seq_len), ...)`` (morgana/metrics.py:52-185).  A metric object registered
    in several collections is shared between them, exactly as in the reference."""
    def __init__(self, **metrics):
        StatefulMetric.__init__(self, hidden=False)
        self.collections = dict(all=metrics, train={}, valid={}, test={})
        self.metrics = self.collections['all']              # alias: every metric ever registered
        for mode in ('train', 'valid'):
            self.collections[mode].update(metrics)

    def __getitem__(self, name):
        try:
            return self.collections[name]
        except KeyError:
            raise ValueError("No collection found by the name {}".format(name)) from None

    def add_metrics(self, collections=('all',), **kwargs):
        targets = _names(collections)
        if 'all' in targets:                                # 'all' means every collection that exists
            targets = list(self.collections)
        for target in targets:
            self.collections[target].update(kwargs)
        self.metrics.update(kwargs)

    def add_collection(self, collection, from_collections=tuple()):
        merged = {}
        for source in _names(from_collections):
            merged.update(self[source])
        self.collections[collection] = merged

    def reset_state(self, collection, *args):
        for metric in self[collection].values():
            metric.reset_state()

    def accumulate(self, collection, **kwargs):
        r"""``name=(inputs..., [kwargs dict])`` per metric; each accumulate is one launch, none of them synchronises."""
        members = self[collection]
        for name, inputs in kwargs.items():
            args = _listify(inputs)
            options = args.pop() if isinstance(args[-1], dict) else {}
            members[name].accumulate(*args, **options)

    def result(self, collection='all', *args):
        return {name: metric.result(*args) for name, metric in self[collection].items()}

    def _visible(self, collection):
        return ((name, metric) for name, metric in self[collection].items() if not metric.hidden)

    def results_as_json_dict(self, collection='all', prefix=''):
        return {prefix + name: metric.result_as_json() for name, metric in self._visible(collection)}

    def results_as_str_dict(self, collection='all', prefix=''):
        return {prefix + name: str(metric) for name, metric in self._visible(collection)}

    def __str__(self):
        return ' | '.join('{} = {}'.format(name, text) for name, text in self.results_as_str_dict('all').items())


class Print(StatefulMetric):
    r"""Keeps the last reported value (morgana/metrics.py:188-213)."""
    def __init__(self, hidden=False):
        StatefulMetric.__init__(self, hidden=hidden)
        self.reset_state()

    def reset_state(self, *args):
        StatefulMetric.reset_state(self)
        self.value = None

    def accumulate(self, tensor):
        StatefulMetric.accumulate(self)
        self.value = tensor

    def result(self, *args):
        return self.value


class History(StatefulMetric):
    r"""The (up to ``max_len``) most recent items of the iterables handed to ``accumulate`` (morgana/metrics.py:216-260);
    summarised by its newest item."""
    def __init__(self, max_len=None, hidden=False):
        StatefulMetric.__init__(self, hidden=hidden)
        self.max_len = max_len
        self.reset_state()

    def reset_state(self):
        StatefulMetric.reset_state(self)
        self.history = []

    def accumulate(self, obj):
        StatefulMetric.accumulate(self)
        self.history.extend(obj)
        if self.max_len is not None:
            self.history = self.history[-self.max_len:]

    def result(self):
        return self.history

    def str_summary(self, result):
        return str(result[-1])

    def result_as_json(self):
        return str(self)

    def __str__(self):
        return self.str_summary(self.result())


class TensorHistory(StatefulMetric):
    r"""The last ``max_len`` valid feature vectors seen, ``(n, feat_dim)`` (morgana/metrics.py:263-356).  With ``seq_len``
    the valid rows are packed by ``mg_pack_rows`` (one scan + one row-copy kernel) instead of mask / nonzero / index."""
    def __init__(self, feat_dim, max_len=None, dtype=torch.float32, device=None, hidden=False):
        StatefulMetric.__init__(self, hidden=hidden)
        self.feat_dim = feat_dim
        self.max_len = max_len
        self.dtype = dtype
        self.device = device
        self.reset_state()

    def reset_state(self):
        StatefulMetric.reset_state(self)
        shape = (0,) if self.feat_dim == 0 else (0, self.feat_dim)
        self.history = torch.empty(shape, dtype=self.dtype)
        if self.device is not None:
            self.history = self.history.to(self.device)

    def accumulate(self, tensor, seq_len=None):
        StatefulMetric.accumulate(self)
        if self.device is None:
            self.device = tensor.device
            self.history = self.history.to(self.device)
        tensor = tensor.to(self.device)
        if seq_len is None:
            tensor = tensor.reshape(-1, self.feat_dim)
        else:
            tensor = ops.pack_rows(tensor, seq_len)
        self.history = torch.cat([self.history, tensor])
        if self.max_len is not None:
            self.history = self.history[-self.max_len:]

    def result(self):
        return self.history

    def str_summary(self, result):
        mean, std, mmin, mmax = torch.mean(result), torch.std(result), torch.min(result), torch.max(result)
        if torch.isnan(std):
            std = torch.zeros_like(std)
        return 'N({mean}, {std}) in range [{min}, {max}]'.format(
            mean=_format_float(mean), std=_format_float(std), min=_format_float(mmin), max=_format_float(mmax))

    def result_as_json(self):
        result = self.result()
        return result.item() if result.numel() == 1 else self.str_summary(result)

    def __str__(self):
        result = self.result()
        return _format_float(result.item()) if result.numel() == 1 else self.str_summary(result)


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
    a, b, m = _as_batch(kind, a, b, m, seq_len)
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


def _as_batch(kind, a, b, m, seq_len):
    r"""Without ``seq_len`` the reference reduces tensors of any shape (a 0-dim batch loss at experiment_builder.py:484,
    ``count += numel``); present them to the kernel as one utterance of ``numel`` single-feature frames.  With
    ``seq_len`` the mask only broadcasts against ``(batch_size, seq_len, feat_dim)``."""
    if a.dim() == 3:
        return a, b, m
    if seq_len is not None:
        raise ValueError('metric inputs must have shape (batch_size, seq_len, feat_dim), got {}'.format(tuple(a.shape)))
    if b is not None and b.shape != a.shape:
        a, b = torch.broadcast_tensors(a, b)
    if m is not None and m.shape != a.shape:
        m = m.expand(a.shape)
    width = a.shape[-1] if (kind == _lib.RED_ROOT_SQDIFF and a.dim() >= 1) else 1
    shape = (1, -1, width)
    return a.reshape(shape), None if b is None else b.reshape(shape), None if m is None else m.reshape(1, -1)


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


class Variance(StatefulMetric):
    r"""Online variance from the masked sum and sum of squares (morgana/metrics.py:400-446) -- both reduced by ONE launch
    (two terms, ``MG_RED_SUM`` and ``MG_RED_SQ``).  Unlike the reference (``tensor *= sequence_mask``, :436) the input
    is not modified."""
    def __init__(self, hidden=False):
        StatefulMetric.__init__(self, hidden=hidden)
        self.reset_state()

    def reset_state(self):
        StatefulMetric.reset_state(self)
        self._records = None
        self.sum = 0.
        self.sum_square = 0.
        self.count = 0.

    def accumulate(self, tensor, seq_len=None):
        self.hidden = self._hidden
        ops._require_cuda(tensor, 'metric input')
        tensor, _, _ = _as_batch(_lib.RED_SUM, _as_float(tensor), None, None, seq_len)
        if tensor.dtype != torch.float32:
            tensor = tensor.to(torch.float32)
        if self._records is None or self._records.device != tensor.device:
            self._records = ops.new_result_records(2, tensor.device)
        B, T, _ = tensor.shape
        if B == 0:
            return
        with ops._device_of(tensor):
            terms = [ops.make_term(_lib.RED_SUM, tensor, result=self._records[0], accumulate=True),
                     ops.make_term(_lib.RED_SQ, tensor, result=self._records[1], accumulate=True)]
            ops.masked_reduce(terms, seq_len, B, T, tensor.device)
        self.sum = self._records[0].view(torch.float32)[ops.F32_SUM]
        self.sum_square = self._records[1].view(torch.float32)[ops.F32_SUM]
        self.count = self._records[0].view(torch.float64)[ops.F64_COUNT]

    def result(self, *args):
        if self._records is None:
            count = self.count + 1e-8
            return torch.as_tensor((self.sum_square - (self.sum ** 2) / count) / count, dtype=torch.float32)
        f64 = self._records.view(torch.float64)                  # (2, 6): row 0 = sum record, row 1 = sum of squares
        count = f64[0, ops.F64_COUNT] + 1e-8
        total = f64[0, ops.F64_SUM]
        return ((f64[1, ops.F64_SUM] - total * total / count) / count).to(torch.float32)


class StandardDeviation(Variance):
    r"""Square root of :class:`Variance` (morgana/metrics.py:449-471)."""
    def result(self, *args):
        return Variance.result(self, *args) ** 0.5


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


ACCUMULATORS = (Mean, Variance, StandardDeviation, RMSE, MAE, Accuracy, Error, F0Distortion, LF0Distortion, Distortion,
                MelCepDistortion)
