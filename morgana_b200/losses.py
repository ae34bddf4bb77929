"""Drop-in counterparts of the reference's masked sequence losses (``morgana/losses.py:9-56``).

``mse(predictions, targets, seq_len=None)`` and ``bce(...)`` keep the reference's signature and semantics --
``mean over (batch, feature) of [ sum over valid frames / number of valid frames ]`` -- and return a 0-dim float32 tensor
that supports ``.backward()``.  ``l1`` is the same wrapper around the absolute error; ``ce`` takes logits and class
indices (``morgana/losses.py:59-61``); ``KLD_standard_normal`` is the unmasked latent-space term (``morgana/losses.py:64-67``).
"""
import functools

from morgana_b200 import ops


def sequence_loss(loss_fn):
    r"""Sequence-loss wrapper of the reference (``morgana/losses.py:9-47``): adds the optional ``seq_len`` argument that masks
    padded frames.  ``loss_fn(predictions, targets)`` is the caller's own per-element loss (any differentiable torch code
    returning ``(batch_size, seq_len, feat_dim)``); the masking, the per-utterance normalisation by the number of valid frames
    and the mean over batch items and feature dimensions are one kernel launch (forward) and one (backward)."""
    @functools.wraps(loss_fn)
    def wrapped_loss(predictions, targets, seq_len=None):
        return ops.masked_loss(loss_fn(predictions, targets), None, seq_len, 'mean')
    return wrapped_loss


def mse(predictions, targets, seq_len=None):
    """Masked mean-squared error (morgana/losses.py:49-51)."""
    return ops.masked_loss(predictions, targets, seq_len, 'mse')


def bce(predictions, targets, seq_len=None):
    """Masked binary cross-entropy on probabilities, logs clamped at -100 (morgana/losses.py:54-56)."""
    return ops.masked_loss(predictions, targets, seq_len, 'bce')


def l1(predictions, targets, seq_len=None):
    """Masked mean-absolute error: ``sequence_loss`` applied to ``F.l1_loss(reduction='none')``."""
    return ops.masked_loss(predictions, targets, seq_len, 'l1')


def ce(predictions, targets, seq_len=None):
    """Masked cross-entropy of ``(batch_size, seq_len, n_classes)`` logits against ``(batch_size, seq_len)`` int64 class
    indices (morgana/losses.py:59-61: ``F.cross_entropy`` over the transposed logits, one value per frame)."""
    return ops.masked_loss(predictions, targets, seq_len, 'ce')


def KLD_standard_normal(mean, log_variance):
    r"""KL-divergence of :math:`\mathbb{N}` (`mean`, `log_variance`) with :math:`\mathbb{N}(0, 1)` (morgana/losses.py:64-67)."""
    return ops.kld_standard_normal(mean, log_variance)
