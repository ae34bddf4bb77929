"""Torch-facing layer over the C ABI: checks tensors, allocates outputs, passes raw device pointers + the current stream.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); every byte of arithmetic on the path is done
by the kernels in ``csrc/``.  There is no CPU or eager fallback: non-CUDA tensors raise.
"""
import ctypes

import torch

from morgana_b200 import _lib
from morgana_b200._lib import lib, check, Term

_NORM_MODES = {None: _lib.NORM_NONE, 'none': _lib.NORM_NONE, 'mvn': _lib.NORM_MVN, 'minmax': _lib.NORM_MINMAX}
_PATHS = {'auto': _lib.PATH_AUTO, 'bulk': _lib.PATH_BULK, 'direct': _lib.PATH_DIRECT}
_INT_DTYPES = (torch.int64, torch.int32, torch.int16, torch.int8, torch.uint8)


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)
_raw_device = getattr(torch._C, '_cuda_getDevice', None)     # the C call behind torch.cuda.current_device(), without its lazy-init check


def _current_device():
    return _raw_device() if _raw_device is not None else torch.cuda.current_device()


def _stream():
    """The current CUDA stream of the current device as the integer handle the C ABI takes (no Stream object is built:
    ~0.3 us instead of ~2 us per launch)."""
    if _raw_stream is not None:
        return _raw_stream(_current_device())
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(tensor, what):
    if not isinstance(tensor, torch.Tensor):
        raise TypeError('{} must be a torch.Tensor, got {}'.format(what, type(tensor).__name__))
    if not tensor.is_cuda:
        raise RuntimeError('morgana_b200 has no CPU path: {} is on {} (move it to a CUDA device, or use the '
                           'unpatched reference for CPU runs)'.format(what, tensor.device))


class _device_of(object):
    """Make `tensor`'s device current for the duration of a launch (no-op when it already is)."""
    def __init__(self, tensor):
        self.idx = tensor.device.index
        self.prev = None

    def __enter__(self):
        cur = _current_device()
        if self.idx is not None and self.idx != cur:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


def _ptr(tensor):
    return tensor.data_ptr() if tensor is not None else None


class KernelProbe(object):
    """CUDA events around named kernel launches made through the public ops, on the stream they are launched on.

        with ops.KernelProbe('K2', 'K4b') as probe:      # measurement only: a pair of event records costs ~3 us of the stream
            out = utils.upsample_to_repetitions(...)
        torch.cuda.synchronize()
        probe.ms('K2')                                   # list of launch durations in ms

    Names: ``K1`` (duration scan), ``K2`` (normalise + expansion), ``K4b`` (whole-row objective).  Without an active probe the
    ops record nothing.
    """
    active = None

    def __init__(self, *names):
        self.names = frozenset(names)
        self.events = dict((name, []) for name in names)
        self._open = {}

    def __enter__(self):
        self._previous, KernelProbe.active = KernelProbe.active, self
        return self

    def __exit__(self, *exc):
        KernelProbe.active = self._previous

    def begin(self, name):
        if name in self.names:
            event = torch.cuda.Event(enable_timing=True)
            event.record(torch.cuda.current_stream())
            self._open[name] = event

    def end(self, name):
        start = self._open.pop(name, None)
        if start is not None:
            stop = torch.cuda.Event(enable_timing=True)
            stop.record(torch.cuda.current_stream())
            self.events[name].append((start, stop))

    def ms(self, name):
        return [a.elapsed_time(b) for a, b in self.events[name]]


# ----------------------------------------------------------------------------------------------------------------------
# K1 + K2: duration scan and fused normalise + expansion
# ----------------------------------------------------------------------------------------------------------------------

def _prepare_repeats(repeats, batch_size):
    _require_cuda(repeats, 'repeats')
    if repeats.dtype.is_floating_point or repeats.dtype == torch.bool or repeats.dtype.is_complex:
        # The reference fails with TypeError for non-integer repeats (morgana/utils.py:211).
        raise TypeError('repeats must be an integer tensor, got {}'.format(repeats.dtype))
    repeats = repeats.reshape(batch_size, -1)
    if repeats.dtype not in (torch.int64, torch.int32):
        repeats = repeats.to(torch.int64)
    if repeats.shape[1] > 0 and repeats.stride(1) != 1:
        repeats = repeats.contiguous()
    return repeats


def dur_scan(repeats, want_summary=True):
    """K1.  Returns ``(ends int32 (B, P), n_frames int64 (B,), summary int64 (4,))``, all on the device; ``summary`` is
    None (and neither cleared nor accumulated) with ``want_summary=False``."""
    repeats = _prepare_repeats(repeats, repeats.shape[0])
    B, P = repeats.shape
    dev = repeats.device
    ends = torch.empty((B, P), dtype=torch.int32, device=dev)
    n_frames = torch.empty((B,), dtype=torch.int64, device=dev)
    summary = torch.empty((4,), dtype=torch.int64, device=dev) if want_summary else None
    with _device_of(repeats):
        check(lib.mg_dur_scan(_ptr(repeats), int(repeats.dtype == torch.int32), repeats.stride(0) if B else 0, B, P,
                              _ptr(ends), _ptr(n_frames), _ptr(summary), _stream()), 'mg_dur_scan')
    return ends, n_frames, summary


def _norm_args(norm, feat_dim, batch_size, device):
    """-> (mode, p0, p1, param_stride_b).  `norm` is None or (kind, p0, p1) with (D,) or (B, D) fp32 parameters."""
    if norm is None:
        return _lib.NORM_NONE, None, None, 0
    kind, p0, p1 = norm
    mode = _NORM_MODES[kind]
    if mode == _lib.NORM_NONE:
        return mode, None, None, 0
    for name, p in (('p0', p0), ('p1', p1)):
        _require_cuda(p, 'normaliser parameter ' + name)
        if p.dtype != torch.float32:
            raise TypeError('normaliser parameters must be float32, got {}'.format(p.dtype))
        if p.requires_grad:
            raise NotImplementedError('gradients w.r.t. normaliser parameters are not provided')
    if p0.shape != p1.shape or p0.shape[-1] != feat_dim or p0.dim() not in (1, 2):
        raise ValueError('normaliser parameters must both be (D,) or (B, D) with D={}; got {} and {}'.format(
            feat_dim, tuple(p0.shape), tuple(p1.shape)))
    p0, p1 = p0.contiguous(), p1.contiguous()
    if p0.dim() == 2:
        if p0.shape[0] != batch_size:
            raise ValueError('per-utterance parameters need one row per batch item')
        return mode, p0, p1, feat_dim
    return mode, p0, p1, 0


def _upsample_forward(x, repeats, norm, max_len, path, out_dtype=None):
    _require_cuda(x, 'sequence_feature')
    if x.dim() != 3:
        raise IndexError('sequence_feature must have shape (batch_size, max_seq_len, feat_dim)')  # utils.py:196
    B, P, D = x.shape
    repeats = _prepare_repeats(repeats, B)
    if repeats.shape[1] != P:
        raise ValueError('repeats has {} items per utterance, sequence_feature has {}'.format(repeats.shape[1], P))
    if repeats.device != x.device:
        raise RuntimeError('repeats and sequence_feature are on different devices')
    ends, n_frames, summary = dur_scan(repeats, want_summary=max_len is None)
    if max_len is None:
        # The one permitted device->host read: 32 bytes that size the output (the reference syncs here too,
        # morgana/utils.py:199) and carry the validity flags.
        max_frames, n_negative, _, n_overflow = summary.tolist()
        if n_negative:
            raise ValueError('repeats may not contain negative values.')  # np.repeat's message, utils.py:220
        if n_overflow:
            raise OverflowError('an utterance expands to more than 2**31 - 1 frames')
        T = int(max_frames)
    else:
        T = int(max_len)   # caller-supplied bound (e.g. features['n_frames'].max() known on the host): no sync

    mode, p0, p1, p_sb = _norm_args(norm, D, B, x.device)
    if x.stride(2) != 1 and D > 1:
        x = x.contiguous()
    if out_dtype is not None and out_dtype != x.dtype:
        if x.dtype != torch.float32 or out_dtype != torch.bfloat16:
            raise TypeError('out_dtype: only float32 features -> bfloat16 frames is provided')
        out = torch.empty((B, T, D), dtype=torch.bfloat16, device=x.device)
        with _device_of(x):
            check(lib.mg_upsample_norm_f32_bf16out(_ptr(x), x.stride(0), x.stride(1), _ptr(ends), _ptr(p0), _ptr(p1), p_sb, mode,
                                                   _ptr(out), B, P, D, T, _stream()), 'mg_upsample_norm_f32_bf16out')
        return out, ends, n_frames, (mode, p0, p1, p_sb)
    out = torch.empty((B, T, D), dtype=x.dtype, device=x.device)
    probe = KernelProbe.active
    with _device_of(x):
        if x.dtype == torch.float32:
            if probe is not None:
                probe.begin('K2')
            check(lib.mg_upsample_norm_f32(_ptr(x), x.stride(0), x.stride(1), _ptr(ends), _ptr(p0), _ptr(p1), p_sb, mode,
                                           _ptr(out), B, P, D, T, _PATHS[path], _stream()), 'mg_upsample_norm_f32')
            if probe is not None:
                probe.end('K2')
        else:
            if mode != _lib.NORM_NONE:
                raise TypeError('fused normalisation needs float32 features, got {}'.format(x.dtype))
            es = x.element_size()
            check(lib.mg_upsample_bytes(_ptr(x), x.stride(0) * es, x.stride(1) * es, _ptr(ends), _ptr(out), B, P, D * es,
                                        T, _PATHS[path], _stream()), 'mg_upsample_bytes')
    return out, ends, n_frames, (mode, p0, p1, p_sb)


class _UpsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, repeats, norm, max_len, path):
        out, ends, n_frames, norm_args = _upsample_forward(x, repeats, norm, max_len, path)
        ctx.ends, ctx.norm_args, ctx.in_shape, ctx.in_dtype = ends, norm_args, tuple(x.shape), x.dtype
        ctx.mark_non_differentiable(n_frames)
        return out, n_frames

    @staticmethod
    def backward(ctx, grad_out, _grad_n_frames):
        if ctx.in_dtype != torch.float32:
            raise NotImplementedError('backward of upsample_to_repetitions is provided for float32 features only')
        B, P, D = ctx.in_shape
        mode, p0, p1, p_sb = ctx.norm_args
        grad_out = grad_out.contiguous()
        grad_x = torch.empty((B, P, D), dtype=torch.float32, device=grad_out.device)
        with _device_of(grad_out):
            check(lib.mg_upsample_norm_bwd_f32(_ptr(grad_out), _ptr(ctx.ends), _ptr(p0), _ptr(p1), p_sb, mode, _ptr(grad_x),
                                               B, P, D, grad_out.shape[1], _stream()), 'mg_upsample_norm_bwd_f32')
        return grad_x, None, None, None, None


def upsample(x, repeats, norm=None, max_len=None, path='auto', return_lengths=False, out_dtype=None):
    """``out[b, t] = norm(x[b, item(b, t)])`` zero-padded to the longest utterance; optionally also ``n_frames``."""
    if out_dtype is None and isinstance(x, torch.Tensor) and x.requires_grad and torch.is_grad_enabled():
        out, n_frames = _UpsampleFn.apply(x, repeats, norm, max_len, path)
    else:
        out, _, n_frames, _ = _upsample_forward(x, repeats, norm, max_len, path, out_dtype)
    return (out, n_frames) if return_lengths else out


def upsample_packed(packed, packed_repeats, n_items, norm=None, max_len=None, max_items=None, return_lengths=False):
    """K1 + K2 on the packed (ragged) wire format: ``packed`` (sum(n_items), D) float32 items and ``packed_repeats``
    (sum(n_items),) durations, utterance after utterance; ``n_items`` (B,) item counts.  Equal to padding both to
    (B, max(n_items), ...) and calling :func:`upsample`, without reading or producing the padding.

    ``max_len`` / ``max_items``: the longest utterance in frames / items when the host knows them (no device->host read).
    """
    _require_cuda(packed, 'packed')
    _require_cuda(packed_repeats, 'packed_repeats')
    _require_cuda(n_items, 'n_items')
    if packed.dim() != 2 or packed.dtype != torch.float32:
        raise TypeError('packed must be a (total_items, feat_dim) float32 tensor')
    if packed_repeats.dtype.is_floating_point or packed_repeats.dtype == torch.bool:
        raise TypeError('repeats must be an integer tensor, got {}'.format(packed_repeats.dtype))
    packed_repeats = packed_repeats.reshape(-1)
    if packed_repeats.dtype not in (torch.int64, torch.int32):
        packed_repeats = packed_repeats.to(torch.int64)
    packed_repeats = packed_repeats.contiguous()
    if packed_repeats.shape[0] != packed.shape[0]:
        raise ValueError('{} durations for {} items'.format(packed_repeats.shape[0], packed.shape[0]))
    if packed.shape[1] > 1 and packed.stride(1) != 1:
        packed = packed.contiguous()
    B, D, dev = n_items.shape[0], packed.shape[1], packed.device
    item_ends, _, item_summary = dur_scan(n_items.reshape(1, B))
    item_ends = item_ends.reshape(B)
    ends = torch.empty((packed.shape[0],), dtype=torch.int32, device=dev)
    n_frames = torch.empty((B,), dtype=torch.int64, device=dev)
    summary = torch.empty((4,), dtype=torch.int64, device=dev)
    with _device_of(packed):
        check(lib.mg_dur_scan_packed(_ptr(packed_repeats), int(packed_repeats.dtype == torch.int32), _ptr(item_ends), B,
                                     packed.shape[0], _ptr(ends), _ptr(n_frames), _ptr(summary), _stream()), 'mg_dur_scan_packed')
    if max_len is None or max_items is None:
        max_frames, n_negative, _, n_overflow = summary.tolist()        # the one device->host read, as in upsample()
        largest, bad_counts, total_items, _ = item_summary.tolist()
        if bad_counts or total_items != packed.shape[0]:      # (the kernels clamp such counts to the arrays)
            raise ValueError('n_items sums to {} but {} items were given'.format(total_items, packed.shape[0]))
        if n_negative:
            raise ValueError('repeats may not contain negative values.')
        if n_overflow:
            raise OverflowError('an utterance expands to more than 2**31 - 1 frames')
        max_len = int(max_frames) if max_len is None else int(max_len)
        max_items = int(largest) if max_items is None else int(max_items)
    mode, p0, p1, p_sb = _norm_args(norm, D, B, dev)
    out = torch.empty((B, int(max_len), D), dtype=torch.float32, device=dev)
    with _device_of(packed):
        check(lib.mg_upsample_packed_norm_f32(_ptr(packed), packed.stride(0), _ptr(item_ends), packed.shape[0], _ptr(ends), _ptr(p0), _ptr(p1),
                                              p_sb, mode, _ptr(out), B, int(max_items), D, int(max_len), _stream()),
              'mg_upsample_packed_norm_f32')
    return (out, n_frames) if return_lengths else out


# ----------------------------------------------------------------------------------------------------------------------
# K0: on-device collate (packed rows -> zero-padded batch)
# ----------------------------------------------------------------------------------------------------------------------

def pad_collate(packed, lengths, max_len=None):
    """``(sum(lengths), D)`` packed rows -> ``(B, T, D)`` zero-padded batch on the device (any dtype).

    ``lengths``: (B,) integer tensor on the device.  ``max_len``: T, when known on the host (no sync); otherwise the
    32-byte summary of the length scan is read back.
    """
    _require_cuda(packed, 'packed')
    _require_cuda(lengths, 'lengths')
    if packed.dim() != 2:
        raise ValueError('packed must be (total_rows, feat_dim)')
    packed = packed.contiguous()
    B = lengths.shape[0]
    ends, _, summary = dur_scan(lengths.reshape(1, B))
    if max_len is None:
        _, n_negative, total, _ = summary.tolist()
        if n_negative:
            raise ValueError('lengths may not contain negative values.')
        if total != packed.shape[0]:
            raise ValueError('lengths sum to {} rows but packed has {}'.format(total, packed.shape[0]))
        max_len = int(lengths.max().item()) if B else 0
    D = packed.shape[1]
    out = torch.empty((B, int(max_len), D), dtype=packed.dtype, device=packed.device)
    with _device_of(packed):
        check(lib.mg_pad_collate(_ptr(packed), _ptr(ends), _ptr(out), B, D * packed.element_size(), int(max_len), packed.shape[0],
                                 _stream()), 'mg_pad_collate')
    return out


# ----------------------------------------------------------------------------------------------------------------------
# K3: standalone normalise / denormalise
# ----------------------------------------------------------------------------------------------------------------------

def _normalise_launch(x, p0, p1, mode, inverse, rows_per_param):
    x = x.contiguous()
    out = torch.empty_like(x)
    D = x.shape[-1]
    rows = x.numel() // D if D else 0
    with _device_of(x):
        check(lib.mg_normalise_f32(_ptr(x), _ptr(p0), _ptr(p1), mode, int(inverse), _ptr(out), rows, D, rows_per_param,
                                   _stream()), 'mg_normalise_f32')
    return out


def _scale_vector(mode, p0, p1):
    """The multiplier the (de)normaliser applies, for the backward pass only (D-sized; not on the forward path)."""
    if mode == _lib.NORM_MVN:
        return p1
    scale = p1 - p0
    return torch.where(scale.abs() <= 1e-8, torch.ones_like(scale), scale)


class _NormaliseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p0, p1, mode, inverse, rows_per_param):
        ctx.save_for_backward(p0, p1)
        ctx.mode, ctx.inverse, ctx.rows_per_param = mode, inverse, rows_per_param
        return _normalise_launch(x, p0, p1, mode, inverse, rows_per_param)

    @staticmethod
    def backward(ctx, grad):
        p0, p1 = ctx.saved_tensors
        scale = _scale_vector(ctx.mode, p0, p1).contiguous()
        zero = torch.zeros_like(scale)
        # d/dx (x - a) / s = 1 / s and d/dx (x * s + a) = s: the same kernel with a = 0 and the scale as p1.
        grad_x = _normalise_launch(grad, zero, scale, ctx.mode, ctx.inverse, ctx.rows_per_param)
        return grad_x, None, None, None, None, None


def normalise(feature, p0, p1, kind, inverse=False):
    """mvn / minmax (de)normalisation of ``(..., T, D)`` features with ``(D,)`` or ``(..., D)`` parameters."""
    _require_cuda(feature, 'feature')
    if feature.dtype != torch.float32:
        raise TypeError('morgana_b200 normalisers take float32 features, got {}'.format(feature.dtype))
    mode = _NORM_MODES[kind]
    if mode == _lib.NORM_NONE:
        raise ValueError('kind must be "mvn" or "minmax"')
    D = feature.shape[-1]
    for p in (p0, p1):
        _require_cuda(p, 'normaliser parameter')
        if p.dtype != torch.float32:
            raise TypeError('normaliser parameters must be float32, got {}'.format(p.dtype))
        if p.requires_grad:
            raise NotImplementedError('gradients w.r.t. normaliser parameters are not provided')
    if p0.shape != p1.shape or p0.shape[-1] != D:
        raise ValueError('parameter shapes {} / {} do not match feature dim {}'.format(tuple(p0.shape), tuple(p1.shape), D))
    if p0.dim() == 1:
        rows_per_param = 0
    else:
        # Parameters broadcast as p[..., None, :] (morgana/data.py:534): one row per leading index, shared over time.
        if feature.dim() < 2 or tuple(p0.shape[:-1]) != tuple(feature.shape[:-2]):
            raise ValueError('parameters of shape {} do not broadcast over features of shape {} as p[..., None, :]'
                             .format(tuple(p0.shape), tuple(feature.shape)))
        rows_per_param = feature.shape[-2]
        if rows_per_param == 0:
            return torch.empty_like(feature)
    p0, p1 = p0.contiguous(), p1.contiguous()
    if feature.requires_grad and torch.is_grad_enabled():
        return _NormaliseFn.apply(feature, p0, p1, mode, bool(inverse), rows_per_param)
    return _normalise_launch(feature, p0, p1, mode, bool(inverse), rows_per_param)


# ----------------------------------------------------------------------------------------------------------------------
# K4 / K5: masked reductions
# ----------------------------------------------------------------------------------------------------------------------

RESULT_BYTES = _lib.TERM_RESULT_BYTES
# Views of one 48-byte result record.
F64_SUM, F64_COUNT, F64_LOSS, I64_ISUM = 0, 1, 2, 3
F32_SUM, F32_COUNT, F32_LOSS, F32_TOTAL = 8, 9, 10, 11

_workspaces = {}


def _workspace(device, n_terms, batch_size, max_len):
    """Zero-initialised scratch per (device, stream); the kernel leaves it clean, so it is zeroed only when (re)made."""
    need = lib.mg_masked_reduce_workspace_bytes(n_terms, batch_size, max_len)
    key = (device.index, _stream())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros((max(need, 1 << 20),), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def new_result_records(n, device):
    """``n`` zeroed ``mg_term_result`` records; row i viewed as float32 / float64 / int64 gives its fields."""
    return torch.zeros((n, RESULT_BYTES), dtype=torch.uint8, device=device)


def new_output_records(n, device):
    """``n`` uninitialised records for results that are WRITTEN, not accumulated into (a loss value): the finisher stores
    every field, so no fill kernel is needed."""
    return torch.empty((n, RESULT_BYTES), dtype=torch.uint8, device=device)


def _view3(t):
    """(B, T, D) view with unit inner stride -> (tensor, stride_b, stride_t)."""
    if t.dim() != 3:
        raise ValueError('expected a (batch_size, seq_len, feat_dim) tensor, got shape {}'.format(tuple(t.shape)))
    if t.shape[2] > 1 and t.stride(2) != 1:
        t = t.contiguous()
    return t, t.stride(0), t.stride(1)


def _dtype_code(t):
    if t.dtype == torch.float32:
        return _lib.DT_F32
    if t.dtype in (torch.uint8, torch.bool):
        return _lib.DT_U8
    raise TypeError('morgana_b200 reductions take float32 or uint8/bool tensors, got {}'.format(t.dtype))


def make_term(kind, a, b=None, m=None, result=None, accumulate=False, grad=None, grad_scale=1.0, grad_scale_dev=None,
              flags=0):
    """Describe one reduction for :func:`masked_reduce`.  Returns ``(Term, keepalive)``."""
    _require_cuda(a, 'tensor')
    a, a_sb, a_st = _view3(a)
    B, T, D = a.shape
    keep = [a, result]
    term = Term()
    term.kind, term.D = kind, D
    term.a, term.a_sb, term.a_st = _ptr(a), a_sb, a_st
    term.ab_dtype = _dtype_code(a)
    if kind == _lib.RED_CE:
        _require_cuda(b, 'targets')
        if b.dtype != torch.int64 or tuple(b.shape) != (B, T):
            raise TypeError('cross-entropy targets must be int64 class indices of shape (batch_size, seq_len)')
        term.b, term.b_sb, term.b_st = _ptr(b), b.stride(0), b.stride(1)
        keep.append(b)
    elif b is not None:
        _require_cuda(b, 'tensor')
        if tuple(b.shape) != (B, T, D):
            raise RuntimeError('operand shapes differ: {} vs {}'.format(tuple(a.shape), tuple(b.shape)))
        b, b_sb, b_st = _view3(b)
        b_code = _dtype_code(b)
        if kind == _lib.RED_EQ:
            term.b_is_u8 = int(b_code == _lib.DT_U8)
        elif b_code != term.ab_dtype:
            raise TypeError('operands must share a dtype, got {} and {}'.format(a.dtype, b.dtype))
        term.b, term.b_sb, term.b_st = _ptr(b), b_sb, b_st
        keep.append(b)
    if m is not None:
        _require_cuda(m, 'weight')
        if m.dim() == 3 and m.shape[2] == 1:
            m = m[:, :, 0]
        if tuple(m.shape) != (B, T):
            raise RuntimeError('per-frame weight must have shape (B, T) or (B, T, 1), got {}'.format(tuple(m.shape)))
        term.m, term.m_sb, term.m_st, term.m_dtype = _ptr(m), m.stride(0), m.stride(1), _dtype_code(m)
        keep.append(m)
    if grad is not None:
        term.grad, term.g_sb, term.g_st = _ptr(grad), grad.stride(0), grad.stride(1)
        if grad_scale_dev is not None:
            term.grad_scale_dev = _ptr(grad_scale_dev)
            keep.append(grad_scale_dev)
        keep.append(grad)
    term.grad_scale = float(grad_scale)   # also the term's weight in the first record's weighted total
    term.result = _ptr(result)
    term.accumulate = int(bool(accumulate))
    term.flags = int(flags)
    return term, keep


def _seq_len_arg(seq_len, batch_size, device):
    if seq_len is None:
        return None
    if not isinstance(seq_len, torch.Tensor):
        seq_len = torch.as_tensor(seq_len, device=device)
    _require_cuda(seq_len, 'seq_len')
    if tuple(seq_len.shape) != (batch_size,):
        # The reference fails to broadcast anything but (batch_size,).
        raise RuntimeError('seq_len must have shape (batch_size,) = ({},), got {}'.format(batch_size, tuple(seq_len.shape)))
    if seq_len.dtype.is_floating_point:
        seq_len = torch.ceil(seq_len)   # arange(T) < seq_len keeps ceil(seq_len) frames (utils.py:140-142)
    return seq_len.to(torch.int64).contiguous()


def masked_reduce(terms, seq_len, batch_size, max_len, device):
    """One launch over up to 12 terms.  `terms` is a list of (Term, keepalive) from :func:`make_term`."""
    n = len(terms)
    if not 1 <= n <= _lib.MAX_TERMS:
        raise ValueError('between 1 and {} terms per launch, got {}'.format(_lib.MAX_TERMS, n))
    seq_len = _seq_len_arg(seq_len, batch_size, device)
    array = (Term * n)(*[t for t, _ in terms])
    ws = _workspace(device, n, batch_size, max_len)
    check(lib.mg_masked_reduce(array, n, _ptr(seq_len), batch_size, max_len, _ptr(ws), ws.numel(), _stream()),
          'mg_masked_reduce')


def column_table(columns, device):
    """Upload a list of :class:`_lib.Column` (one per feature column) as a device byte tensor."""
    array = (_lib.Column * len(columns))(*columns)
    raw = torch.frombuffer(bytearray(bytes(array)), dtype=torch.uint8)
    return raw.to(device)


def masked_objective(pred, target, seq_len, cols, slots, grad=None, grad_scale_dev=None):
    """K4b: one whole-row pass over (B, T, D) `pred` / `target`; `cols` from :func:`column_table`, `slots` a ctypes array."""
    _require_cuda(pred, 'pred')
    _require_cuda(target, 'target')
    if pred.dtype != torch.float32 or target.dtype != torch.float32 or pred.shape != target.shape:
        raise TypeError('masked_objective takes two float32 tensors of one shape')
    pred, p_sb, p_st = _view3(pred)
    target, t_sb, t_st = _view3(target)
    B, T, D = pred.shape
    seq_len = _seq_len_arg(seq_len, B, pred.device)
    n_slots = len(slots)
    ws = _workspace(pred.device, n_slots, B, T)
    g_sb, g_st = (grad.stride(0), grad.stride(1)) if grad is not None else (0, 0)
    probe = KernelProbe.active
    if probe is not None:
        probe.begin('K4b')
    check(lib.mg_masked_objective_f32(_ptr(pred), p_sb, p_st, _ptr(target), t_sb, t_st, _ptr(grad), g_sb, g_st,
                                      _ptr(grad_scale_dev), _ptr(cols), D, slots, n_slots, _ptr(seq_len), B, T, _ptr(ws),
                                      ws.numel(), _stream()), 'mg_masked_objective_f32')
    if probe is not None:
        probe.end('K4b')


_LOSS_KINDS = {'mse': _lib.RED_SQDIFF, 'l1': _lib.RED_ABSDIFF, 'bce': _lib.RED_BCE, 'ce': _lib.RED_CE, 'mean': _lib.RED_SUM}


class _MaskedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, predictions, targets, seq_len, kind):
        B, T, D = predictions.shape
        record = new_output_records(1, predictions.device)
        with _device_of(predictions):
            term = make_term(kind, predictions, targets, result=record[0])
            masked_reduce([term], seq_len, B, T, predictions.device)
        ctx.save_for_backward(predictions, targets)
        ctx.seq_len, ctx.kind = seq_len, kind
        ctx.set_materialize_grads(False)
        return record[0].view(torch.float32)[F32_LOSS]

    @staticmethod
    def backward(ctx, grad_output):
        predictions, targets = ctx.saved_tensors
        if grad_output is None:
            return None, None, None, None
        B, T, D = predictions.shape
        grad = torch.empty((B, T, D), dtype=torch.float32, device=predictions.device)
        scale = grad_output.detach().to(torch.float32).contiguous()
        record = new_output_records(1, predictions.device)
        with _device_of(predictions):
            term = make_term(ctx.kind, predictions, targets, result=record[0], grad=grad, grad_scale=1.0,
                             grad_scale_dev=scale)
            masked_reduce([term], ctx.seq_len, B, T, predictions.device)
        grad_targets = None
        if ctx.needs_input_grad[1]:
            if ctx.kind != _lib.RED_SQDIFF:
                raise NotImplementedError('gradient w.r.t. targets is provided for mse only')
            grad_targets = -grad
        return grad, grad_targets, None, None


def masked_loss(predictions, targets, seq_len=None, kind='mse'):
    """``mean_{b,d} [ sum_{t < n_b} l(p, y) / n_b ]`` as a 0-dim float32 tensor with autograd."""
    _require_cuda(predictions, 'predictions')
    if kind == 'mean':     # the masked mean of a per-element loss the caller computed (losses.sequence_loss around a custom loss_fn)
        if targets is not None:
            raise TypeError("kind 'mean' takes the per-element loss alone")
        if predictions.dtype != torch.float32 or predictions.dim() != 3:
            raise TypeError('sequence_loss needs a float32 (batch_size, seq_len, feat_dim) per-element loss, got {} {}'
                            .format(predictions.dtype, tuple(predictions.shape)))
        if predictions.shape[0] == 0 or predictions.shape[2] == 0:
            return torch.full((), float('nan'), dtype=torch.float32, device=predictions.device)
        seq_len = _seq_len_arg(seq_len, predictions.shape[0], predictions.device)
        return _MaskedLossFn.apply(predictions, None, seq_len, _LOSS_KINDS[kind])
    _require_cuda(targets, 'targets')
    if kind == 'ce':
        if predictions.dtype != torch.float32 or predictions.dim() != 3 or targets.dtype != torch.int64 or \
                tuple(targets.shape) != tuple(predictions.shape[:2]):
            raise RuntimeError('ce takes float32 logits (batch_size, seq_len, n_classes) and int64 targets (batch_size, '
                               'seq_len), got {} {} and {} {}'.format(predictions.dtype, tuple(predictions.shape),
                                                                      targets.dtype, tuple(targets.shape)))
    elif predictions.dtype != torch.float32 or targets.dtype != torch.float32:
        raise TypeError('morgana_b200 losses take float32 tensors, got {} and {}'.format(predictions.dtype, targets.dtype))
    elif predictions.dim() != 3 or predictions.shape != targets.shape:
        raise RuntimeError('predictions and targets must share a (batch_size, seq_len, feat_dim) shape, got {} and {}'
                           .format(tuple(predictions.shape), tuple(targets.shape)))
    if predictions.shape[0] == 0 or predictions.shape[2] == 0:
        return torch.full((), float('nan'), dtype=torch.float32, device=predictions.device)
    seq_len = _seq_len_arg(seq_len, predictions.shape[0], predictions.device)
    if not (torch.is_grad_enabled() and (predictions.requires_grad or (targets is not None and targets.requires_grad))):
        # validation / metrics: no graph to build, one launch and a view of the record (skips ~10 us of autograd bookkeeping)
        B, T, D = predictions.shape
        record = new_output_records(1, predictions.device)
        with _device_of(predictions):
            term = make_term(_LOSS_KINDS[kind], predictions, targets, result=record[0])
            masked_reduce([term], seq_len, B, T, predictions.device)
        return record[0].view(torch.float32)[F32_LOSS]
    return _MaskedLossFn.apply(predictions, targets, seq_len, _LOSS_KINDS[kind])


class _KldFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, log_variance):
        rows, latent = mean.numel() // mean.shape[-1], mean.shape[-1]
        loss = torch.empty((), dtype=torch.float32, device=mean.device)
        ws = torch.empty((max(lib.mg_kld_workspace_bytes(mean.numel()), 8) // 8,), dtype=torch.float64, device=mean.device)
        with _device_of(mean):
            check(lib.mg_kld_standard_normal_f32(_ptr(mean), _ptr(log_variance), rows, latent, _ptr(loss), None, None, None,
                                                 _ptr(ws), ws.numel() * 8, _stream()), 'mg_kld_standard_normal_f32')
        ctx.save_for_backward(mean, log_variance)
        ctx.set_materialize_grads(False)
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        if grad_output is None:
            return None, None
        mean, log_variance = ctx.saved_tensors
        rows, latent = mean.numel() // mean.shape[-1], mean.shape[-1]
        grad_mean, grad_lv = torch.empty_like(mean), torch.empty_like(log_variance)
        scale = grad_output.detach().to(torch.float32).contiguous()
        ws = torch.empty((max(lib.mg_kld_workspace_bytes(mean.numel()), 8) // 8,), dtype=torch.float64, device=mean.device)
        with _device_of(mean):
            check(lib.mg_kld_standard_normal_f32(_ptr(mean), _ptr(log_variance), rows, latent, None, _ptr(grad_mean), _ptr(grad_lv),
                                                 _ptr(scale), _ptr(ws), ws.numel() * 8, _stream()), 'mg_kld_standard_normal_f32')
        return grad_mean, grad_lv


def kld_standard_normal(mean, log_variance):
    """``mean over rows of -0.5 * sum_d (1 + log_variance - mean**2 - exp(log_variance))`` (reference losses.py:64-67)."""
    _require_cuda(mean, 'mean')
    _require_cuda(log_variance, 'log_variance')
    if mean.dtype != torch.float32 or log_variance.dtype != torch.float32:
        raise TypeError('morgana_b200 losses take float32 tensors, got {} and {}'.format(mean.dtype, log_variance.dtype))
    if mean.shape != log_variance.shape or mean.dim() < 1:
        raise RuntimeError('mean and log_variance must share a (..., latent_dim) shape, got {} and {}'
                           .format(tuple(mean.shape), tuple(log_variance.shape)))
    if mean.numel() == 0:
        return torch.full((), float('nan'), dtype=torch.float32, device=mean.device)      # torch.mean of nothing
    return _KldFn.apply(mean.contiguous(), log_variance.contiguous())


# ----------------------------------------------------------------------------------------------------------------------
# K6: multi-tensor EMA
# ----------------------------------------------------------------------------------------------------------------------

class EmaPlan(object):
    """Pointer tables for one (shadow, param) pairing, rebuilt only when a storage moves."""
    def __init__(self):
        self.key = None
        self.n = 0
        self.shadow = self.param = self.numel = None

    def update(self, pairs):
        key = tuple((s.data_ptr(), p.data_ptr(), s.numel()) for s, p in pairs)
        if key != self.key:
            n = len(key)
            self.shadow = (ctypes.c_void_p * n)(*[k[0] for k in key])
            self.param = (ctypes.c_void_p * n)(*[k[1] for k in key])
            self.numel = (ctypes.c_int64 * n)(*[k[2] for k in key])
            self.key, self.n = key, n
        return self


def ema_update(pairs, one_minus_decay, plan=None):
    """``shadow -= fl32(one_minus_decay) * (shadow - param)`` in place for every (shadow, param) pair, one launch per 64."""
    if not pairs:
        return
    for s, p in pairs:
        _require_cuda(s, 'EMA shadow')
        _require_cuda(p, 'EMA parameter')
        if s.dtype != torch.float32 or p.dtype != torch.float32:
            raise TypeError('morgana_b200 EMA takes float32 parameters, got {} / {}'.format(s.dtype, p.dtype))
        if s.shape != p.shape:
            raise RuntimeError('EMA shadow {} and parameter {} differ in shape'.format(tuple(s.shape), tuple(p.shape)))
        if not s.is_contiguous() or not p.is_contiguous():
            raise NotImplementedError('EMA needs contiguous parameters (the update is in place on their storage)')
    plan = (plan or EmaPlan()).update(pairs)
    with _device_of(pairs[0][0]):
        check(lib.mg_ema_update_f32(plan.shadow, plan.param, plan.numel, plan.n, ctypes.c_float(one_minus_decay),
                                    _stream()), 'mg_ema_update_f32')


def ema_update_tensors(shadows, params, one_minus_decay, plan):
    """:func:`ema_update` for a fixed list of tensors that is updated every step: the checks and the pointer tables are redone
    only when a storage moved (the per-step cost is one ``data_ptr()`` per tensor and the launch)."""
    key = tuple(t.data_ptr() for t in params) + tuple(t.data_ptr() for t in shadows)
    if getattr(plan, 'fast_key', None) != key or plan.n != len(params):
        ema_update([(s, p.detach()) for s, p in zip(shadows, params)], one_minus_decay, plan=plan)
        plan.fast_key = key
        return
    if not params:
        return
    with _device_of(shadows[0]):
        check(lib.mg_ema_update_f32(plan.shadow, plan.param, plan.numel, plan.n, ctypes.c_float(one_minus_decay),
                                    _stream()), 'mg_ema_update_f32')


def both_nonzero(features):
    """uint8 tensor of the features' shape: 1 where every feature is non-zero (reference utils.py:169-172)."""
    if not features:
        raise RuntimeError('both_voiced_mask needs at least one sequence feature')     # torch.stack([]) raises too
    if len(features) > 8:
        raise NotImplementedError('both_voiced_mask takes at most 8 sequence features')
    shape = features[0].shape
    tensors = []
    for f in features:
        _require_cuda(f, 'sequence_feature')
        if f.shape != shape:
            raise RuntimeError('stack expects each tensor to be equal size, but got {} and {}'.format(list(shape), list(f.shape)))
        tensors.append(f.to(torch.float32).contiguous())
    out = torch.empty(shape, dtype=torch.uint8, device=features[0].device)
    table = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    with _device_of(out):
        check(lib.mg_both_nonzero_u8(table, len(tensors), out.numel(), _ptr(out), _stream()), 'mg_both_nonzero_u8')
    return out


# ----------------------------------------------------------------------------------------------------------------------
# K7: dense layers
# ----------------------------------------------------------------------------------------------------------------------

def cast_pad_bf16(x, k_padded=None):
    """fp32 (rows, K) -> bf16 (rows, k_padded) with zero-filled tail columns."""
    _require_cuda(x, 'x')
    if x.dtype != torch.float32 or x.dim() != 2:
        raise TypeError('cast_pad_bf16 takes a 2-D float32 tensor')
    if x.shape[1] > 1 and x.stride(1) != 1:
        x = x.contiguous()
    rows, K = x.shape
    k_padded = k_padded or ((K + 7) // 8) * 8
    out = torch.empty((rows, k_padded), dtype=torch.bfloat16, device=x.device)
    with _device_of(x):
        check(lib.mg_cast_pad_bf16(_ptr(x), x.stride(0), _ptr(out), k_padded, rows, K, _stream()), 'mg_cast_pad_bf16')
    return out


def cast_transpose_bf16(weight):
    """fp32 ``(N, K)`` weight -> its transpose ``(K, round_up(N, 8))`` in bf16 (zero padding columns): the B operand of the
    input-gradient GEMM ``g @ W`` as :func:`linear_bf16` takes it."""
    _require_cuda(weight, 'weight')
    if weight.dim() != 2 or weight.dtype != torch.float32:
        raise TypeError('cast_transpose_bf16 takes a 2-D float32 tensor')
    if weight.shape[1] > 1 and weight.stride(1) != 1:
        weight = weight.contiguous()
    N, K = weight.shape
    ld_out = (N + 7) // 8 * 8
    out = torch.empty((K, ld_out), dtype=torch.bfloat16, device=weight.device)
    with _device_of(weight):
        check(lib.mg_cast_transpose_bf16(_ptr(weight), weight.stride(0) if N else K, _ptr(out), ld_out, N, K, _stream()),
              'mg_cast_transpose_bf16')
    return out


_ACTS = {None: _lib.ACT_NONE, 'none': _lib.ACT_NONE, 'sigmoid': _lib.ACT_SIGMOID}


def linear_bf16(x, weight, bias=None, act=None, out_dtype=torch.float32, in_features=None):
    """K7: ``act(x @ weight.T + bias)`` on the tcgen05 tensor cores.  bf16 operands, fp32 accumulation.

    x : (M, K) bfloat16, row stride a multiple of 8 (fp32 input is converted with :func:`cast_pad_bf16`).
    weight : (N, K) bfloat16 (``nn.Linear.weight`` layout), bias : (N,) float32 or None.
    in_features : the true K when the operands carry padding columns; by default both operands must have the same width up
    to the padding to a multiple of 8 (a real mismatch raises, as ``nn.Linear`` does).
    """
    _require_cuda(x, 'x')
    _require_cuda(weight, 'weight')
    if x.dim() != 2 or weight.dim() != 2:
        raise ValueError('linear_bf16 takes 2-D operands')
    if x.dtype == torch.float32:
        x = cast_pad_bf16(x)
    if weight.dtype == torch.float32:
        weight = cast_pad_bf16(weight)
    if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise TypeError('linear_bf16 takes bfloat16 (or float32, converted) operands')
    M, N = x.shape[0], weight.shape[0]
    if in_features is None:
        if (x.shape[1] + 7) // 8 != (weight.shape[1] + 7) // 8:
            raise RuntimeError('mat1 and mat2 shapes cannot be multiplied ({}x{} and {}x{})'
                               .format(M, x.shape[1], weight.shape[1], N))
        K = min(x.shape[1], weight.shape[1])   # a padded operand carries zeros beyond the true K
    else:
        K = int(in_features)
        if not 0 <= K <= min(x.shape[1], weight.shape[1]):
            raise ValueError('in_features={} exceeds the operands ({} and {} columns)'.format(K, x.shape[1], weight.shape[1]))
    for name, t in (('x', x), ('weight', weight)):
        if t.stride(1) != 1 or t.stride(0) % 8 != 0 or t.data_ptr() % 16 != 0:
            raise ValueError('{}: rows must be contiguous, 16-byte aligned, with a stride that is a multiple of 8'.format(name))
    if bias is not None:
        _require_cuda(bias, 'bias')
        bias = bias.to(torch.float32).contiguous()
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise TypeError('out_dtype must be float32 or bfloat16')
    y = torch.empty((M, N), dtype=out_dtype, device=x.device)
    with _device_of(x):
        check(lib.mg_linear_bf16(_ptr(x), x.stride(0), _ptr(weight), weight.stride(0), _ptr(bias), _ptr(y), y.stride(0),
                                 int(out_dtype == torch.bfloat16), M, N, K, _ACTS[act], _stream()), 'mg_linear_bf16')
    return y


def act_grad_bf16(grad_y, y=None, want_bias_grad=True):
    """K7 backward, first step: ``g = grad_y * (1 - y) * y`` (sigmoid; ``y=None``: ``g = grad_y``) as bf16 rows padded to a
    multiple of 8 columns, and ``g.sum(0)`` (float32, from the unrounded values) in the same pass.

    grad_y, y : (M, N) float32 or bfloat16.  Returns ``(g_bf16 (M, round_up(N, 8)), bias_grad (N,) or None)``.
    """
    _require_cuda(grad_y, 'grad_y')
    if grad_y.dim() != 2 or (y is not None and y.shape != grad_y.shape):
        raise ValueError('act_grad_bf16 takes 2-D tensors of one shape')
    for name, t in (('grad_y', grad_y), ('y', y)):
        if t is not None and t.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError('{}: float32 or bfloat16 expected, got {}'.format(name, t.dtype))
    if grad_y.shape[1] > 1 and grad_y.stride(1) != 1:
        grad_y = grad_y.contiguous()
    if y is not None and y.shape[1] > 1 and y.stride(1) != 1:
        y = y.contiguous()
    M, N = grad_y.shape
    ld_out = (N + 7) // 8 * 8
    dev = grad_y.device
    out = torch.empty((M, ld_out), dtype=torch.bfloat16, device=dev)
    bias_grad = torch.empty((N,), dtype=torch.float32, device=dev) if want_bias_grad else None
    ws_bytes = lib.mg_act_grad_workspace_bytes(M, N) if want_bias_grad else 0
    ws = torch.empty((max(ws_bytes, 8) // 8,), dtype=torch.float64, device=dev)
    with _device_of(grad_y):
        check(lib.mg_act_grad_bf16(_ptr(grad_y), int(grad_y.dtype == torch.bfloat16), grad_y.stride(0) if M else N,
                                   _ptr(y), int(y is not None and y.dtype == torch.bfloat16),
                                   y.stride(0) if (y is not None and M) else N, _ptr(out), ld_out, _ptr(bias_grad), M, N,
                                   _ptr(ws), ws.numel() * 8, _stream()), 'mg_act_grad_bf16')
    return out, bias_grad


def linear_wgrad_bf16(g, x, out_features=None, in_features=None):
    """K7 backward, weight gradient: ``g[:, :N].T @ x[:, :K]`` -> (N, K) float32 on the tcgen05 tensor cores.

    g : (M, >= N) bfloat16, x : (M, >= K) bfloat16, both row-major with row strides that are multiples of 8 -- the layouts
    :func:`act_grad_bf16` and :func:`cast_pad_bf16` produce; nothing is transposed in memory.
    """
    _require_cuda(g, 'g')
    _require_cuda(x, 'x')
    if g.dim() != 2 or x.dim() != 2 or g.shape[0] != x.shape[0]:
        raise ValueError('linear_wgrad_bf16 takes 2-D operands with one row per frame')
    if g.dtype != torch.bfloat16 or x.dtype != torch.bfloat16:
        raise TypeError('linear_wgrad_bf16 takes bfloat16 operands')
    M = g.shape[0]
    N = g.shape[1] if out_features is None else int(out_features)
    K = x.shape[1] if in_features is None else int(in_features)
    if not (0 < N <= g.shape[1] and 0 < K <= x.shape[1]):
        raise ValueError('out_features / in_features exceed the operands')
    for name, t in (('g', g), ('x', x)):
        if M and (t.stride(1) != 1 or t.stride(0) % 8 != 0 or t.data_ptr() % 16 != 0):
            raise ValueError('{}: rows must be contiguous, 16-byte aligned, with a stride that is a multiple of 8'.format(name))
    grad_w = torch.empty((N, K), dtype=torch.float32, device=g.device)
    ws_bytes = lib.mg_linear_wgrad_workspace_bytes(M, N, K)
    ws = torch.empty((max(ws_bytes, 16) // 4,), dtype=torch.float32, device=g.device)
    with _device_of(g):
        check(lib.mg_linear_wgrad_bf16(_ptr(g), g.stride(0) if M else 8, _ptr(x), x.stride(0) if M else 8, _ptr(grad_w), K, M, N, K,
                                       _ptr(ws), ws.numel() * 4, _stream()), 'mg_linear_wgrad_bf16')
    return grad_w


# ----------------------------------------------------------------------------------------------------------------------
# K8: batched MLPG
# ----------------------------------------------------------------------------------------------------------------------

_mlpg_workspaces = {}


def mlpg_window_table(windows):
    """The reference's ``windows`` argument -- a list of ``(l, u, win_coeff)`` with ``len(win_coeff) == l + u + 1``
    (viz/synthesis.py:8-29) -- as rows of coefficients at frame offsets (-1, 0, +1).  Windows reaching further than one frame
    to either side would widen the band of the system beyond the pentadiagonal solver: NotImplementedError."""
    rows = []
    for l, u, coeff in windows:
        l, u = int(l), int(u)
        coeff = [float(c) for c in coeff]
        if l < 0 or u < 0 or len(coeff) != l + u + 1:
            raise AssertionError('window (l, u, win_coeff) needs len(win_coeff) == l + u + 1')      # synthesis.py:31-32
        if l > 1 or u > 1:
            raise NotImplementedError('MLPG windows that reach more than one frame to either side are not provided '
                                      '(got l={}, u={})'.format(l, u))
        rows.append([coeff[l - 1] if l == 1 else 0., coeff[l], coeff[l + 1] if u == 1 else 0.])
    if not 1 <= len(rows) <= 4:
        raise NotImplementedError('MLPG takes 1 to 4 windows, got {}'.format(len(rows)))
    return rows


def mlpg(means, variances, padding_size=0, seq_len=None, windows=None):
    """Maximum-likelihood parameter generation for (B, T, W * F) means laid out window after window ([static | delta |
    delta-delta] for the default windows) on the device.

    ``variances``: (W * F,) global, (B, W * F) per utterance or (B, T, W * F) per frame.  ``windows``: None (the reference's
    three defaults) or the rows :func:`mlpg_window_table` makes.  Returns (B, T, F) float32.
    """
    _require_cuda(means, 'means')
    _require_cuda(variances, 'variances')
    if means.dtype != torch.float32 or variances.dtype != torch.float32:
        raise TypeError('mlpg takes float32 means and variances')
    n_windows = 3 if windows is None else len(windows)
    if means.dim() != 3 or means.shape[2] % n_windows != 0:
        raise ValueError('means must be (batch_size, seq_len, n_windows * feat_dim)')
    if means.stride(2) != 1:
        means = means.contiguous()
    B, T, D3 = means.shape
    F = D3 // n_windows
    table = None
    if windows is not None:
        table = (ctypes.c_double * (3 * n_windows))(*[c for row in windows for c in row])
    variances = variances.contiguous()
    if variances.dim() == 1 and variances.shape[0] == D3:
        v_sb, v_st = 0, 0
    elif variances.dim() == 2 and tuple(variances.shape) == (B, D3):
        v_sb, v_st = D3, 0
    elif variances.dim() == 3 and tuple(variances.shape) == (B, T, D3):
        v_sb, v_st = T * D3, D3
    else:
        raise ValueError('variances of shape {} do not match means {}'.format(tuple(variances.shape), tuple(means.shape)))
    if seq_len is not None:
        seq_len = _seq_len_arg(seq_len, B, means.device)
    out = torch.empty((B, T, F), dtype=torch.float32, device=means.device)
    if B == 0 or T == 0 or F == 0:
        return out
    need = lib.mg_mlpg_workspace_bytes(B, T, F, int(padding_size))
    key = (means.device.index, _stream())
    ws = _mlpg_workspaces.get(key)
    if ws is None or ws.numel() * 8 < need:
        ws = torch.empty(((need + 7) // 8,), dtype=torch.float64, device=means.device)
        _mlpg_workspaces[key] = ws
    with _device_of(means):
        check(lib.mg_mlpg_f32(_ptr(means), means.stride(0), means.stride(1), _ptr(variances), v_sb, v_st, _ptr(seq_len), _ptr(out),
                              out.stride(0), out.stride(1), B, T, F, int(padding_size), table, n_windows, _ptr(ws), ws.numel() * 8, _stream()),
              'mg_mlpg_f32')
    return out


# ----------------------------------------------------------------------------------------------------------------------
# sibling segment operations ("next" row 4)
# ----------------------------------------------------------------------------------------------------------------------

def _rows3(x, what):
    _require_cuda(x, what)
    if x.dim() != 3:
        raise ValueError('{} must have shape (batch_size, max_seq_len, feat_dim)'.format(what))
    if x.shape[2] > 1 and x.stride(2) != 1:
        x = x.contiguous()
    es = x.element_size()
    return x, x.stride(0) * es, x.stride(1) * es, x.shape[2] * es


def _pack_rows_launch(x, ends, total):
    x, sb, st, row_bytes = _rows3(x, 'sequence_feature')
    B, T, D = x.shape
    out = torch.empty((total, D), dtype=x.dtype, device=x.device)
    with _device_of(x):
        check(lib.mg_pack_rows(_ptr(x), sb, st, _ptr(ends), _ptr(out), B, T, row_bytes, _stream()), 'mg_pack_rows')
    return out


class _PackRowsFn(torch.autograd.Function):
    """Backward = the on-device collate: packed gradient rows back into a zero-padded (B, T, D) batch."""
    @staticmethod
    def forward(ctx, x, ends, total):
        ctx.save_for_backward(ends)
        ctx.shape = tuple(x.shape)
        return _pack_rows_launch(x, ends, total)

    @staticmethod
    def backward(ctx, grad):
        (ends,) = ctx.saved_tensors
        B, T, D = ctx.shape
        grad = grad.contiguous()
        grad_x = torch.empty((B, T, D), dtype=grad.dtype, device=grad.device)
        with _device_of(grad):
            check(lib.mg_pad_collate(_ptr(grad), _ptr(ends), _ptr(grad_x), B, D * grad.element_size(), T, grad.shape[0],
                                     _stream()), 'mg_pad_collate')
        return grad_x, None, None


def pack_rows(x, seq_len):
    """``(B, T, D)`` + lengths -> ``(sum(min(len, T)), D)``: the rows inside every utterance, back to back.  Differentiable."""
    _require_cuda(x, 'sequence_feature')
    if x.dim() != 3:
        raise ValueError('sequence_feature must have shape (batch_size, max_seq_len, feat_dim)')
    B, T, D = x.shape
    if not isinstance(seq_len, torch.Tensor):
        seq_len = torch.as_tensor(seq_len, device=x.device)
    _require_cuda(seq_len, 'seq_len')
    lengths = seq_len.reshape(B).to(torch.int64).clamp(0, T)
    ends, _, summary = dur_scan(lengths.reshape(1, B))
    total = int(summary[2].item())                       # the output size is data dependent: one 8-byte read, as the reference's nonzero()
    if x.requires_grad and torch.is_grad_enabled():
        return _PackRowsFn.apply(x, ends, total)
    return _pack_rows_launch(x, ends, total)


def _segment_scan(segment_lens, batch_size):
    lens = _prepare_repeats(segment_lens, batch_size)
    ends, _, summary = dur_scan(lens)
    return lens, ends, summary


def _segments_bwd(grad, ends, shape, L):
    B, T, D = shape
    S = ends.shape[1]
    grad = grad.contiguous()
    grad_x = torch.empty((B, T, D), dtype=grad.dtype, device=grad.device)
    with _device_of(grad):
        check(lib.mg_segments_bwd(_ptr(grad), _ptr(ends), _ptr(grad_x), B, S, L, T, D * grad.element_size(), _stream()),
              'mg_segments_bwd')
    return grad_x


def _segment_ends_launch(x, ends):
    x, sb, st, row_bytes = _rows3(x, 'sequence_feature')
    B, T, D = x.shape
    S = ends.shape[1]
    out = torch.empty((B, S, D), dtype=x.dtype, device=x.device)
    with _device_of(x):
        check(lib.mg_segment_ends(_ptr(x), sb, st, _ptr(ends), _ptr(out), B, S, T, row_bytes, _stream()), 'mg_segment_ends')
    return out


class _SegmentEndsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ends):
        ctx.save_for_backward(ends)
        ctx.shape = tuple(x.shape)
        return _segment_ends_launch(x, ends)

    @staticmethod
    def backward(ctx, grad):
        (ends,) = ctx.saved_tensors
        return _segments_bwd(grad, ends, ctx.shape, 0), None


def segment_ends(x, segment_lens):
    """``out[b, s] = x[b, cumsum(lens)[b, s] - 1]``, zero for empty segments.  Differentiable (clockwork RNNs train through it)."""
    _require_cuda(x, 'sequence_feature')
    if x.dim() != 3:
        raise ValueError('sequence_feature must have shape (batch_size, max_seq_len, feat_dim)')
    _, ends, _ = _segment_scan(segment_lens, x.shape[0])
    if x.requires_grad and torch.is_grad_enabled():
        return _SegmentEndsFn.apply(x, ends)
    return _segment_ends_launch(x, ends)


def _split_launch(x, ends, L):
    x, sb, st, row_bytes = _rows3(x, 'sequence_feature')
    B, T, D = x.shape
    S = ends.shape[1]
    out = torch.empty((B, S, L, D), dtype=x.dtype, device=x.device)
    with _device_of(x):
        check(lib.mg_split_to_segments(_ptr(x), sb, st, _ptr(ends), _ptr(out), B, S, L, T, row_bytes, _stream()),
              'mg_split_to_segments')
    return out


class _SplitToSegmentsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ends, L):
        ctx.save_for_backward(ends)
        ctx.shape, ctx.L = tuple(x.shape), L
        return _split_launch(x, ends, L)

    @staticmethod
    def backward(ctx, grad):
        (ends,) = ctx.saved_tensors
        if ctx.L == 0:
            return torch.zeros(ctx.shape, dtype=grad.dtype, device=grad.device), None, None
        return _segments_bwd(grad, ends, ctx.shape, ctx.L), None, None


def split_to_segments(x, segment_lens, max_segment_len=None):
    """``out[b, s, j] = x[b, begin_s + j]`` for ``j < len_s`` else zero; ``(B, S, longest segment, D)``.  Differentiable."""
    _require_cuda(x, 'sequence_feature')
    if x.dim() != 3:
        raise ValueError('sequence_feature must have shape (batch_size, max_seq_len, feat_dim)')
    lens, ends, summary = _segment_scan(segment_lens, x.shape[0])
    if max_segment_len is None:
        if int(summary[1].item()):
            raise ValueError('segment_lens may not contain negative values.')
        max_segment_len = int(lens.max().item()) if lens.numel() else 0     # the reference syncs here too (utils.py:256)
    L = int(max_segment_len)
    if x.requires_grad and torch.is_grad_enabled():
        return _SplitToSegmentsFn.apply(x, ends, L)
    return _split_launch(x, ends, L)
