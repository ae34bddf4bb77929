"""``torch.ops.morgana_b200.*``: the kernels registered as torch custom operators (CUDA backend only).

The functions in :mod:`morgana_b200.utils` / ``losses`` / ``metrics`` / ``data`` call the C ABI directly (lowest
overhead).  The same entry points are also registered with the dispatcher so that code which wants operator objects
(``torch.ops`` call sites, export, fake-tensor shape propagation) can have them.  Only a CUDA kernel is registered: a CPU
tensor fails in the dispatcher with ``NotImplementedError`` -- there is no CPU fallback to fall into.
"""
import torch

from morgana_b200 import ops

_lib = torch.library.Library('morgana_b200', 'DEF')

_lib.define('dur_scan(Tensor repeats) -> (Tensor, Tensor, Tensor)')
_lib.define('upsample_norm(Tensor x, Tensor repeats, Tensor? p0, Tensor? p1, str kind, int max_len) -> Tensor')
_lib.define('pad_collate(Tensor packed, Tensor lengths, int max_len) -> Tensor')
_lib.define('normalise(Tensor x, Tensor p0, Tensor p1, str kind, bool inverse) -> Tensor')
_lib.define('masked_loss(Tensor predictions, Tensor targets, Tensor? seq_len, str kind) -> Tensor')
_lib.define('ema_update(Tensor(a!)[] shadow, Tensor[] params, float one_minus_decay) -> ()')
_lib.define('linear_bf16(Tensor x, Tensor weight, Tensor? bias, str act, bool bf16_out) -> Tensor')
_lib.define('act_grad_bf16(Tensor grad_y, Tensor? y) -> (Tensor, Tensor)')
_lib.define('linear_wgrad_bf16(Tensor g, Tensor x, int out_features, int in_features) -> Tensor')
_lib.define('mlpg(Tensor means, Tensor variances, int padding_size, Tensor? seq_len) -> Tensor')


def _dur_scan(repeats):
    return ops.dur_scan(repeats)


def _upsample_norm(x, repeats, p0, p1, kind, max_len):
    norm = None if kind in ('', 'none') else (kind, p0, p1)
    return ops.upsample(x, repeats, norm=norm, max_len=max_len if max_len >= 0 else None)


def _pad_collate(packed, lengths, max_len):
    return ops.pad_collate(packed, lengths, max_len=max_len if max_len >= 0 else None)


def _normalise(x, p0, p1, kind, inverse):
    return ops.normalise(x, p0, p1, kind, inverse=inverse)


def _masked_loss(predictions, targets, seq_len, kind):
    return ops.masked_loss(predictions, targets, seq_len, kind)


def _ema_update(shadow, params, one_minus_decay):
    ops.ema_update(list(zip(shadow, params)), one_minus_decay)


def _linear_bf16(x, weight, bias, act, bf16_out):
    return ops.linear_bf16(x, weight, bias, act=None if act in ('', 'none') else act,
                           out_dtype=torch.bfloat16 if bf16_out else torch.float32)


def _act_grad_bf16(grad_y, y):
    return ops.act_grad_bf16(grad_y, y)


def _linear_wgrad_bf16(g, x, out_features, in_features):
    return ops.linear_wgrad_bf16(g, x, out_features=out_features, in_features=in_features)


def _mlpg(means, variances, padding_size, seq_len):
    return ops.mlpg(means, variances, padding_size=padding_size, seq_len=seq_len)


for _name, _fn in [('dur_scan', _dur_scan), ('upsample_norm', _upsample_norm), ('pad_collate', _pad_collate),
                   ('normalise', _normalise), ('masked_loss', _masked_loss), ('ema_update', _ema_update),
                   ('linear_bf16', _linear_bf16), ('act_grad_bf16', _act_grad_bf16),
                   ('linear_wgrad_bf16', _linear_wgrad_bf16), ('mlpg', _mlpg)]:
    _lib.impl(_name, _fn, 'CUDA')

OPERATORS = ('dur_scan', 'upsample_norm', 'pad_collate', 'normalise', 'masked_loss', 'ema_update', 'linear_bf16',
             'act_grad_bf16', 'linear_wgrad_bf16', 'mlpg')
