"""Drop-in counterparts of the hot-path callables in the reference's ``morgana/utils.py`` (same names and arguments).

* :func:`upsample_to_repetitions` -- morgana/utils.py:175-228
* :func:`sequence_mask` -- morgana/utils.py:115-144 (kept for API completeness; the kernels never build a mask)
* :class:`ExponentialMovingAverage` -- morgana/utils.py:421-456
"""
import torch

from morgana_b200 import ops


def upsample_to_repetitions(sequence_feature, repeats, normaliser=None, deltas=False, max_len=None,
                            return_lengths=False, path='auto', out_dtype=None):
    r"""Copies sequence items according to ``repeats`` (per-utterance ``np.repeat``), zero-padded to the longest result.

    The first two arguments are the reference's (morgana/utils.py:175).  The keyword arguments are additive:

    normaliser : a normaliser exposing ``normalise``-style parameters (see :mod:`morgana_b200.data`), or a
        ``(kind, p0, p1)`` tuple.  The normalisation is fused into the gather: equal, bit for bit, to
        ``upsample_to_repetitions(normaliser.normalise(sequence_feature), repeats)`` (padding stays 0).
    max_len : int, known upper bound of the output length; skips the 32-byte device->host read (no sync at all).
    return_lengths : also return ``n_frames`` (``sum(repeats, dim=1)``, int64, on the device).
    out_dtype : ``torch.bfloat16`` writes the frames as bf16 (the exact fp32 result rounded to nearest-even) -- half the
        bytes, and the activation format of :class:`morgana_b200.nn.Linear`.  No gradient flows through this variant.

    Returns ``(batch_size, max_repeated_len, feat_dim)``, same dtype, contiguous.  Raises ``TypeError`` for non-integer
    ``repeats`` and ``ValueError`` for negative ones, like the reference.
    """
    norm = None
    if normaliser is not None:
        norm = normaliser if isinstance(normaliser, tuple) else normaliser.fused_params(deltas=deltas)
    return ops.upsample(sequence_feature, repeats, norm=norm, max_len=max_len, path=path, return_lengths=return_lengths,
                        out_dtype=out_dtype)


def sequence_mask(seq_len, max_len=None, dtype=torch.ByteTensor, device=None):
    r"""``mask[b, t, 0] = t < seq_len[b]`` as ``(batch_size, max_len, 1)`` (morgana/utils.py:115-144).

    Off the hot path: the masked kernels index rows below ``seq_len`` directly and never read a mask.  Provided so code
    that calls ``utils.sequence_mask`` keeps working; it is three small ATen ops, as in the reference.
    """
    if max_len is None:
        max_len = int(torch.max(seq_len).item())
    if device is None:
        device = seq_len.device
    positions = torch.arange(max_len, device=device).type(seq_len.dtype)
    mask = positions[None, :] < seq_len.to(device)[:, None]
    return mask[:, :, None].type(dtype)


def both_voiced_mask(*sequence_features, dtype=torch.ByteTensor):
    r"""Whether the sequence features are non-zero at the same time (reference ``utils.py:169-172``), computed on the device.

    As in the reference the result is cast with ``.type(dtype)``: the default ``torch.ByteTensor`` is a CPU tensor type, so
    the mask lands on the host exactly as the reference's does; pass ``torch.cuda.ByteTensor`` / ``torch.uint8`` to keep it
    on the device."""
    return ops.both_nonzero(list(sequence_features)).type(dtype)


class ExponentialMovingAverage(object):
    """EMA of a model's trainable parameters, updated by one multi-tensor kernel (morgana/utils.py:421-456).

    ``shadow[name]`` aliases ``model``'s own ``param.data`` and is updated in place, exactly like the reference, so the
    averaged model's ``state_dict()`` is always current.
    """
    def __init__(self, model, decay):
        self.model = model
        self.decay = decay
        self.shadow = {}
        for name, param in self.model.named_parameters():
            if param.requires_grad:
                self.shadow[name] = param.data
        self._plan = ops.EmaPlan()
        self._pairing = None      # (other_model, its parameters that have a shadow, their shadows)

    def _update_param(self, name, x):
        """One tensor: ``shadow -= (1 - decay) * (shadow - x)``."""
        assert name in self.shadow
        ops.ema_update([(self.shadow[name], x)], 1.0 - self.decay)

    def update_params(self, other_model):
        """All tensors of ``other_model`` that have a shadow, in one launch per 64 tensors."""
        assert other_model is not self.model
        pairing = self._pairing
        if pairing is None or pairing[0] is not other_model:
            named = [(name, param) for name, param in other_model.named_parameters() if name in self.shadow]
            pairing = self._pairing = (other_model, [param for _, param in named], [self.shadow[name] for name, _ in named])
        ops.ema_update_tensors(pairing[2], pairing[1], 1.0 - self.decay, self._plan)


def upsample_packed_to_repetitions(packed_feature, packed_repeats, n_items, normaliser=None, deltas=False, max_len=None,
                                  max_items=None, return_lengths=False):
    r"""``upsample_to_repetitions`` on the packed wire format (additive; "next" row 3 of the scope table): the items of the
    batch as one ``(sum(n_items), feat_dim)`` float32 tensor, utterance after utterance, their durations as
    ``(sum(n_items),)`` integers and the per-utterance item counts ``n_items (batch_size,)``.  Returns the same
    ``(batch_size, max_frames, feat_dim)`` tensor as padding the items on the host first (``collate_fn``,
    morgana/data.py:184-193) and calling :func:`upsample_to_repetitions` -- without the padding ever existing."""
    norm = None
    if normaliser is not None:
        norm = normaliser if isinstance(normaliser, tuple) else normaliser.fused_params(deltas=deltas)
    return ops.upsample_packed(packed_feature, packed_repeats, n_items, norm=norm, max_len=max_len, max_items=max_items,
                               return_lengths=return_lengths)


def detach_batched_seqs(*sequence_features, seq_len=None, squeeze=True):
    r"""Converts :class:`torch.Tensor` to `np.ndarray`: moves data to the host, detaches gradients and removes padding
    (``morgana/utils.py:66-102``; called on the outputs of ``predict`` by ``models/RNN_SPSS.py:149`` and ``viz/io.py:47``).

    Same return structure as the reference -- per feature a list of ``(seq_len_b, feat_dim)`` arrays (squeezed when
    ``squeeze``), or the whole padded array without ``seq_len`` -- but a CUDA feature is packed on the device first
    (``mg_pack_rows``): only the valid rows cross PCIe, in one copy, and the per-utterance arrays are views of that buffer.
    """
    import numpy as np
    lengths = seq_len
    if isinstance(lengths, torch.Tensor):
        lengths = lengths.detach().cpu().numpy()
    detached = []
    for feature in sequence_features:
        on_device = isinstance(feature, torch.Tensor) and feature.is_cuda
        if on_device and lengths is not None and feature.dim() > 2 and feature.shape[0] == len(lengths) and \
                all(0 <= int(n) for n in lengths):
            batch_size, max_len = feature.shape[0], feature.shape[1]
            rows = feature.detach().reshape(batch_size, max_len, -1)
            clipped = np.minimum(np.asarray(lengths, dtype=np.int64), max_len)
            packed = ops.pack_rows(rows, torch.as_tensor(clipped, device=feature.device)).cpu().numpy()
            packed = packed.reshape((packed.shape[0],) + tuple(feature.shape[2:]))
            stops = np.cumsum(clipped)
            items = [packed[stop - n:stop] for stop, n in zip(stops, clipped)]
            feature = [item.squeeze() if squeeze else item for item in items]
        else:
            if isinstance(feature, torch.Tensor):
                feature = feature.cpu().detach().numpy()
            if lengths is not None and feature[0].ndim > 1:
                feature = [item[:n].squeeze() if squeeze else item[:n] for item, n in zip(feature, lengths)]
        detached.append(feature)
    if len(detached) == 1:
        return detached[0]
    return detached


def batched_masked_select(sequence_feature, seq_len):
    r"""Feature vectors of all batch items that lie inside their sequence, as one ``(sum(seq_len), feat_dim)`` tensor
    (morgana/utils.py:147-166).  One scan + one row-copy kernel instead of mask / nonzero / advanced indexing."""
    return ops.pack_rows(sequence_feature, seq_len)


def get_segment_ends(sequence_feature, segment_lens):
    r"""Feature at the last position of each segment, ``(batch_size, max_num_segments, feat_dim)``; zero for empty
    segments (morgana/utils.py:287-330)."""
    return ops.segment_ends(sequence_feature, segment_lens)


def split_to_segments(sequence_feature, segment_lens, max_segment_len=None):
    r"""Splits sequences into zero-padded segments, ``(batch_size, max_num_segments, max_segment_len, feat_dim)``
    (morgana/utils.py:231-284; its double Python loop at :272-276 becomes one gather kernel).
    ``max_segment_len`` (additive): known longest segment, skips the device->host read."""
    return ops.split_to_segments(sequence_feature, segment_lens, max_segment_len=max_segment_len)
