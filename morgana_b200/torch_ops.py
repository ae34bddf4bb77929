"""``torch.ops.morgana_b200.*``: the kernels registered as torch custom operators.

The functions in :mod:`morgana_b200.utils` / ``losses`` / ``metrics`` / ``data`` call the C ABI directly (lowest overhead).
The same entry points are also registered with the dispatcher, the way SURVEY.md section 7 step 1 asks, so that code which
wants operator objects can have them:

* **CUDA kernel** for every operator (``Library.impl(..., 'CUDA')``); there is no CPU kernel -- a CPU tensor fails in the
  dispatcher with ``NotImplementedError``, as everywhere else in this package there is nothing to fall back to;
* **fake / meta kernel** (``torch.library.register_fake``) for every operator: output shapes, dtypes and devices without
  touching memory, so ``FakeTensorMode``, ``torch.export`` and ``torch.compile`` tracing can propagate shapes through them
  (data-dependent sizes -- the longest utterance when ``max_len`` is not given -- become unbacked symbolic ints);
* **autograd formulas** (``torch.library.register_autograd``) for the differentiable operators of the path --
  ``upsample_norm``, ``normalise``, ``masked_loss``, ``linear_bf16`` -- written in terms of further registered operators
  (``upsample_norm_backward``, ``masked_loss_backward``, ``act_grad_bf16``, ``linear_wgrad_bf16``, ``cast_transpose_bf16``),
  so the backward graph traces too.
"""
import torch

from morgana_b200 import ops

_lib = torch.library.Library('morgana_b200', 'DEF')

_lib.define('dur_scan(Tensor repeats) -> (Tensor, Tensor, Tensor)')
_lib.define('upsample_norm(Tensor x, Tensor repeats, Tensor? p0, Tensor? p1, str kind, int max_len) -> Tensor')
_lib.define('upsample_norm_backward(Tensor grad_out, Tensor repeats, Tensor? p0, Tensor? p1, str kind) -> Tensor')
_lib.define('pad_collate(Tensor packed, Tensor lengths, int max_len) -> Tensor')
_lib.define('normalise(Tensor x, Tensor p0, Tensor p1, str kind, bool inverse) -> Tensor')
_lib.define('masked_loss(Tensor predictions, Tensor targets, Tensor? seq_len, str kind) -> Tensor')
_lib.define('masked_loss_backward(Tensor grad_output, Tensor predictions, Tensor targets, Tensor? seq_len, str kind) -> Tensor')
_lib.define('ema_update(Tensor(a!)[] shadow, Tensor[] params, float one_minus_decay) -> ()')
_lib.define('linear_bf16(Tensor x, Tensor weight, Tensor? bias, str act, bool bf16_out) -> Tensor')
_lib.define('act_grad_bf16(Tensor grad_y, Tensor? y) -> (Tensor, Tensor)')
_lib.define('linear_wgrad_bf16(Tensor g, Tensor x, int out_features, int in_features) -> Tensor')
_lib.define('cast_transpose_bf16(Tensor weight) -> Tensor')
_lib.define('mlpg(Tensor means, Tensor variances, int padding_size, Tensor? seq_len) -> Tensor')


def _norm(kind, p0, p1):
    return None if kind in ('', 'none') else (kind, p0, p1)


def _round8(n):
    return (n + 7) // 8 * 8


# ----------------------------------------------------------------------------------------------------------------------
# CUDA kernels
# ----------------------------------------------------------------------------------------------------------------------
def _dur_scan(repeats):
    return ops.dur_scan(repeats)


def _upsample_norm(x, repeats, p0, p1, kind, max_len):
    with torch.no_grad():
        return ops.upsample(x, repeats, norm=_norm(kind, p0, p1), max_len=max_len if max_len >= 0 else None)


def _upsample_norm_backward(grad_out, repeats, p0, p1, kind):
    B, T, D = grad_out.shape
    prepared = ops._prepare_repeats(repeats, B)
    ends, _, _ = ops.dur_scan(prepared, want_summary=False)
    mode, q0, q1, p_sb = ops._norm_args(_norm(kind, p0, p1), D, B, grad_out.device)
    grad_out = grad_out.contiguous()
    P = prepared.shape[1]
    grad_x = torch.empty((B, P, D), dtype=torch.float32, device=grad_out.device)
    with ops._device_of(grad_out):
        ops.check(ops.lib.mg_upsample_norm_bwd_f32(ops._ptr(grad_out), ops._ptr(ends), ops._ptr(q0), ops._ptr(q1), p_sb, mode,
                                                   ops._ptr(grad_x), B, P, D, T, ops._stream()), 'mg_upsample_norm_bwd_f32')
    return grad_x


def _pad_collate(packed, lengths, max_len):
    return ops.pad_collate(packed, lengths, max_len=max_len if max_len >= 0 else None)


def _normalise(x, p0, p1, kind, inverse):
    with torch.no_grad():
        return ops.normalise(x, p0, p1, kind, inverse=inverse)


def _masked_loss(predictions, targets, seq_len, kind):
    with torch.no_grad():     # (a fresh 0-dim tensor: the direct path returns a view into its 48-byte result record)
        return ops.masked_loss(predictions, targets, seq_len, kind).clone()


def _masked_loss_backward(grad_output, predictions, targets, seq_len, kind):
    B, T, D = predictions.shape
    seq_len = ops._seq_len_arg(seq_len, B, predictions.device)
    grad = torch.empty((B, T, D), dtype=torch.float32, device=predictions.device)
    record = ops.new_output_records(1, predictions.device)
    scale = grad_output.detach().to(torch.float32).contiguous()
    with ops._device_of(predictions):
        term = ops.make_term(ops._LOSS_KINDS[kind], predictions, targets, result=record[0], grad=grad, grad_scale=1.0,
                             grad_scale_dev=scale)
        ops.masked_reduce([term], seq_len, B, T, predictions.device)
    return grad


def _ema_update(shadow, params, one_minus_decay):
    ops.ema_update(list(zip(shadow, params)), one_minus_decay)


def _linear_bf16(x, weight, bias, act, bf16_out):
    return ops.linear_bf16(x, weight, bias, act=None if act in ('', 'none') else act,
                           out_dtype=torch.bfloat16 if bf16_out else torch.float32)


def _act_grad_bf16(grad_y, y):
    return ops.act_grad_bf16(grad_y, y)


def _linear_wgrad_bf16(g, x, out_features, in_features):
    return ops.linear_wgrad_bf16(g, x, out_features=out_features, in_features=in_features)


def _cast_transpose_bf16(weight):
    return ops.cast_transpose_bf16(weight.float() if weight.dtype != torch.float32 else weight)


def _mlpg(means, variances, padding_size, seq_len):
    return ops.mlpg(means, variances, padding_size=padding_size, seq_len=seq_len)


for _name, _fn in [('dur_scan', _dur_scan), ('upsample_norm', _upsample_norm), ('upsample_norm_backward', _upsample_norm_backward),
                   ('pad_collate', _pad_collate), ('normalise', _normalise), ('masked_loss', _masked_loss),
                   ('masked_loss_backward', _masked_loss_backward), ('ema_update', _ema_update), ('linear_bf16', _linear_bf16),
                   ('act_grad_bf16', _act_grad_bf16), ('linear_wgrad_bf16', _linear_wgrad_bf16),
                   ('cast_transpose_bf16', _cast_transpose_bf16), ('mlpg', _mlpg)]:
    _lib.impl(_name, _fn, 'CUDA')


# ----------------------------------------------------------------------------------------------------------------------
# fake kernels: shapes / dtypes / devices only
# ----------------------------------------------------------------------------------------------------------------------
def _data_dependent_length():
    """The longest utterance is read back from the device in the real kernel; under tracing it is an unbacked size."""
    ctx = torch.library.get_ctx()
    n = ctx.new_dynamic_size()
    return n


@torch.library.register_fake('morgana_b200::dur_scan')
def _(repeats):
    B = repeats.shape[0]
    P = repeats.numel() // B if B else 0
    return (repeats.new_empty((B, P), dtype=torch.int32), repeats.new_empty((B,), dtype=torch.int64),
            repeats.new_empty((4,), dtype=torch.int64))


@torch.library.register_fake('morgana_b200::upsample_norm')
def _(x, repeats, p0, p1, kind, max_len):
    B, P, D = x.shape
    T = max_len if max_len >= 0 else _data_dependent_length()
    return x.new_empty((B, T, D))


@torch.library.register_fake('morgana_b200::upsample_norm_backward')
def _(grad_out, repeats, p0, p1, kind):
    B, T, D = grad_out.shape
    return grad_out.new_empty((B, repeats.numel() // B if B else 0, D), dtype=torch.float32)


@torch.library.register_fake('morgana_b200::pad_collate')
def _(packed, lengths, max_len):
    T = max_len if max_len >= 0 else _data_dependent_length()
    return packed.new_empty((lengths.shape[0], T, packed.shape[1]))


@torch.library.register_fake('morgana_b200::normalise')
def _(x, p0, p1, kind, inverse):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


@torch.library.register_fake('morgana_b200::masked_loss')
def _(predictions, targets, seq_len, kind):
    return predictions.new_empty((), dtype=torch.float32)


@torch.library.register_fake('morgana_b200::masked_loss_backward')
def _(grad_output, predictions, targets, seq_len, kind):
    return predictions.new_empty(tuple(predictions.shape), dtype=torch.float32)


@torch.library.register_fake('morgana_b200::ema_update')
def _(shadow, params, one_minus_decay):
    return None


@torch.library.register_fake('morgana_b200::linear_bf16')
def _(x, weight, bias, act, bf16_out):
    return x.new_empty((x.shape[0], weight.shape[0]), dtype=torch.bfloat16 if bf16_out else torch.float32)


@torch.library.register_fake('morgana_b200::act_grad_bf16')
def _(grad_y, y):
    M, N = grad_y.shape
    return grad_y.new_empty((M, _round8(N)), dtype=torch.bfloat16), grad_y.new_empty((N,), dtype=torch.float32)


@torch.library.register_fake('morgana_b200::linear_wgrad_bf16')
def _(g, x, out_features, in_features):
    return g.new_empty((out_features, in_features), dtype=torch.float32)


@torch.library.register_fake('morgana_b200::cast_transpose_bf16')
def _(weight):
    N, K = weight.shape
    return weight.new_empty((K, _round8(N)), dtype=torch.bfloat16)


@torch.library.register_fake('morgana_b200::mlpg')
def _(means, variances, padding_size, seq_len):
    B, T, D3 = means.shape
    return means.new_empty((B, T, D3 // 3), dtype=torch.float32)


# ----------------------------------------------------------------------------------------------------------------------
# autograd formulas (in terms of registered operators, so the backward graph traces as well)
# ----------------------------------------------------------------------------------------------------------------------
def _upsample_setup(ctx, inputs, output):
    x, repeats, p0, p1, kind, max_len = inputs
    ctx.save_for_backward(repeats, p0, p1)
    ctx.kind = kind


def _upsample_backward(ctx, grad_out):
    repeats, p0, p1 = ctx.saved_tensors
    grad_x = torch.ops.morgana_b200.upsample_norm_backward(grad_out, repeats, p0, p1, ctx.kind)
    return grad_x, None, None, None, None, None


torch.library.register_autograd('morgana_b200::upsample_norm', _upsample_backward, setup_context=_upsample_setup)


def _normalise_setup(ctx, inputs, output):
    x, p0, p1, kind, inverse = inputs
    ctx.save_for_backward(p0, p1)
    ctx.kind, ctx.inverse = kind, inverse


def _normalise_backward(ctx, grad):
    p0, p1 = ctx.saved_tensors
    # d/dx of (x - p0) / scale is 1 / scale, of x * scale + p0 it is scale: the same kernel with a zero offset
    scale = ops._scale_vector(ops._NORM_MODES[ctx.kind], p0, p1)
    zero = torch.zeros_like(p0)
    # (mvn: the kernel adds the reference's 1e-8 to the std itself; minmax: mmax - mmin with mmin = 0 is the scale itself)
    return torch.ops.morgana_b200.normalise(grad.contiguous(), zero, scale, ctx.kind, ctx.inverse), None, None, None, None


torch.library.register_autograd('morgana_b200::normalise', _normalise_backward, setup_context=_normalise_setup)


def _loss_setup(ctx, inputs, output):
    predictions, targets, seq_len, kind = inputs
    ctx.save_for_backward(predictions, targets, seq_len)
    ctx.kind = kind


def _loss_backward(ctx, grad_output):
    predictions, targets, seq_len = ctx.saved_tensors
    grad = torch.ops.morgana_b200.masked_loss_backward(grad_output, predictions, targets, seq_len, ctx.kind)
    grad_targets = -grad if (ctx.needs_input_grad[1] and ctx.kind == 'mse') else None
    return grad, grad_targets, None, None


torch.library.register_autograd('morgana_b200::masked_loss', _loss_backward, setup_context=_loss_setup)


def _linear_setup(ctx, inputs, output):
    x, weight, bias, act, bf16_out = inputs
    ctx.save_for_backward(x, weight, output if act == 'sigmoid' else None)
    ctx.act, ctx.has_bias = act, bias is not None


def _linear_backward(ctx, grad_y):
    x, weight, y = ctx.saved_tensors
    n, k = weight.shape[0], min(x.shape[1], weight.shape[1])
    g16, grad_b = torch.ops.morgana_b200.act_grad_bf16(grad_y.contiguous(), y)
    grad_x = grad_w = None
    if ctx.needs_input_grad[0]:
        w_t = torch.ops.morgana_b200.cast_transpose_bf16(weight)
        grad_x = torch.ops.morgana_b200.linear_bf16(g16, w_t, None, 'none', x.dtype == torch.bfloat16)
        if grad_x.shape[1] != x.shape[1]:
            grad_x = torch.nn.functional.pad(grad_x, (0, x.shape[1] - grad_x.shape[1]))
    if ctx.needs_input_grad[1]:
        grad_w = torch.ops.morgana_b200.linear_wgrad_bf16(g16, x, n, k).to(weight.dtype)
        if grad_w.shape[1] != weight.shape[1]:
            grad_w = torch.nn.functional.pad(grad_w, (0, weight.shape[1] - grad_w.shape[1]))
    return grad_x, grad_w, (grad_b if ctx.has_bias and ctx.needs_input_grad[2] else None), None, None


torch.library.register_autograd('morgana_b200::linear_bf16', _linear_backward, setup_context=_linear_setup)

OPERATORS = ('dur_scan', 'upsample_norm', 'upsample_norm_backward', 'pad_collate', 'normalise', 'masked_loss',
             'masked_loss_backward', 'ema_update', 'linear_bf16', 'act_grad_bf16', 'linear_wgrad_bf16', 'cast_transpose_bf16',
             'mlpg')
