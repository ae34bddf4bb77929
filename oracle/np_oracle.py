"""CPU oracle (numpy) for morgana's per-batch frame-rate feature path.  TEST INFRASTRUCTURE ONLY.

This file is the checker for the CUDA kernels in ``morgana_b200/csrc``.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it;
nothing under ``morgana_b200/`` does, and the product path raises if its CUDA library is missing rather than
falling back to anything here.

Parity pinning: the reference (ZackHodari/morgana) ships no tests, golden vectors or fixtures of its own
(``TODO.md:3``), so the oracle is pinned against outputs of the reference's own functions
executed in the build container: ``tests/golden/make_golden.py`` imports the unmodified reference from
``/root/reference`` and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` holds every function below
to those vectors (bit-exact for index/layout/elementwise work, <=1e-6 relative for reductions).

Conventions: integer and elementwise fp32 work is done in exactly the reference's operation order so results
are bit-identical to ATen's; reductions are carried in float64 (the reference's fp32 reductions sit 5e-9..7e-8
relative from these), and rounded to fp32 only where the reference stores fp32.
All citations are relative to the reference root.
"""
import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------------------------------------
# a1  utils.upsample_to_repetitions  (morgana/utils.py:175-228)
# ----------------------------------------------------------------------------------------------------------------

def _as_2d_repeats(repeats, batch_size):
    """The reference flattens a trailing singleton axis with ``reshape((batch_size, -1))`` (utils.py:201-202)."""
    repeats = np.asarray(repeats)
    if not np.issubdtype(repeats.dtype, np.integer):
        # utils.py:211 -- ``batch_idx.repeat(1, max_repeated_len)`` rejects a float length with TypeError.
        raise TypeError('repeats must be an integer array, got {}'.format(repeats.dtype))
    repeats = repeats.reshape(batch_size, -1).astype(np.int64)
    if (repeats < 0).any():
        # utils.py:220 -- np.repeat raises for negative counts.
        raise ValueError('repeats may not contain negative values.')
    return repeats


def dur_scan(repeats):
    """Inclusive running sum of durations per utterance, per-utterance frame count, and the padded length.

    utils.py:198-199: ``repeated_lens = sum(repeats, dim=1)``; ``max_repeated_len = max(repeated_lens)``.
    """
    repeats = np.asarray(repeats)
    repeats = _as_2d_repeats(repeats, repeats.shape[0])
    ends = np.cumsum(repeats, axis=1, dtype=np.int64)
    n_frames = ends[:, -1].copy() if repeats.shape[1] else np.zeros(repeats.shape[0], np.int64)
    max_frames = int(n_frames.max()) if n_frames.size else 0
    return ends, n_frames, max_frames


def upsample_index_map(repeats):
    """``(B, T)`` map from output frame to source item, -1 where the output is padding (utils.py:214-220)."""
    ends, n_frames, max_frames = dur_scan(repeats)
    batch_size, n_items = ends.shape
    index_map = np.full((batch_size, max_frames), -1, dtype=np.int64)
    frame = np.arange(max_frames, dtype=np.int64)
    for b in range(batch_size):
        # Item p covers frames [ends[p-1], ends[p]); searchsorted(side='right') is the np.repeat expansion.
        src = np.searchsorted(ends[b], frame[:n_frames[b]], side='right')
        index_map[b, :n_frames[b]] = src
    return index_map


def upsample_to_repetitions(sequence_feature, repeats):
    """Per-utterance ``np.repeat`` along the item axis, zero-padded to the longest utterance (utils.py:175-228).

    Output is ``(B, max_b sum_p repeats[b, p], D)``, same dtype, contiguous; frames past an utterance's own
    length are zero (the reference indexes an appended all-zero row for them, utils.py:206-207, 214).
    """
    sequence_feature = np.asarray(sequence_feature)
    if sequence_feature.ndim != 3:
        raise IndexError('sequence_feature must be (batch_size, max_seq_len, feat_dim)')  # utils.py:196
    batch_size, n_items, feat_dim = sequence_feature.shape
    repeats = _as_2d_repeats(repeats, batch_size)
    n_frames = repeats.sum(axis=1)
    max_frames = int(n_frames.max()) if batch_size else 0
    out = np.zeros((batch_size, max_frames, feat_dim), dtype=sequence_feature.dtype)
    for b in range(batch_size):
        out[b, :n_frames[b]] = np.repeat(sequence_feature[b], repeats[b], axis=0)
    return out


def upsample_backward(grad_out, repeats):
    """Gradient of a1 w.r.t. its input: each item receives the sum of the gradient rows of its own frames.

    The reference gets this from autograd of the advanced-index gather at utils.py:226 (``IndexBackward0``).
    Carried in float64, returned in grad_out's dtype.
    """
    grad_out = np.asarray(grad_out)
    batch_size, _, feat_dim = grad_out.shape
    repeats = _as_2d_repeats(repeats, batch_size)
    n_items = repeats.shape[1]
    ends = np.cumsum(repeats, axis=1)
    grad_in = np.zeros((batch_size, n_items, feat_dim), dtype=np.float64)
    for b in range(batch_size):
        start = 0
        for p in range(n_items):
            stop = int(ends[b, p])
            if stop > start:
                grad_in[b, p] = grad_out[b, start:stop].astype(np.float64).sum(axis=0)
            start = stop
    return grad_in.astype(grad_out.dtype)


# ----------------------------------------------------------------------------------------------------------------
# a2  utils.sequence_mask  (morgana/utils.py:115-144)
# ----------------------------------------------------------------------------------------------------------------

def sequence_mask(seq_len, max_len=None, dtype=np.uint8):
    """``mask[b, t, 0] = t < seq_len[b]`` with the positions first cast to seq_len's dtype (utils.py:134-144)."""
    seq_len = np.asarray(seq_len)
    if max_len is None:
        max_len = int(seq_len.max())
    positions = np.arange(max_len).astype(seq_len.dtype)
    return (positions[None, :] < seq_len[:, None])[:, :, None].astype(dtype)


# ----------------------------------------------------------------------------------------------------------------
# a3 / a4  normaliser arithmetic  (morgana/data.py:533-538, 579-590)
# ----------------------------------------------------------------------------------------------------------------

def normalise_mvn(feature, mean, std_dev):
    """``(x - mean) / (std_dev + 1e-8)``, every step rounded to fp32 as ATen does (data.py:533-534)."""
    feature, mean, std_dev = F32(feature), F32(mean), F32(std_dev)
    return (feature - mean[..., None, :]) / (std_dev[..., None, :] + F32(1e-8))


def denormalise_mvn(feature, mean, std_dev):
    """``x * std_dev + mean`` as a multiply then an add -- two roundings, not an FMA (data.py:537-538)."""
    feature, mean, std_dev = F32(feature), F32(mean), F32(std_dev)
    return (feature * std_dev[..., None, :]) + mean[..., None, :]


def minmax_scale(mmin, mmax):
    """``scale = mmax - mmin`` with near-constant dims (``|scale| <= 1e-8``) mapped to 1 (data.py:580-581)."""
    scale = (F32(mmax) - F32(mmin)).astype(F32).copy()
    scale[np.abs(scale) <= F32(1e-8)] = F32(1.)
    return scale


def normalise_minmax(feature, mmin, mmax):
    """``(x - mmin) / scale`` (data.py:579-583)."""
    scale = minmax_scale(mmin, mmax)
    return (F32(feature) - F32(mmin)[..., None, :]) / scale[..., None, :]


def denormalise_minmax(feature, mmin, mmax):
    """``x * scale + mmin`` as multiply then add (data.py:586-590)."""
    scale = minmax_scale(mmin, mmax)
    return (F32(feature) * scale[..., None, :]) + F32(mmin)[..., None, :]


def normalise_upsample(sequence_feature, repeats, kind, p0, p1):
    """The fused op of the north star: normalise at item rate, then a1.

    Equal to the reference's composition ``upsample_to_repetitions(normaliser.normalise(x), dur)`` -- normalise is
    elementwise so it commutes with the gather on valid frames, and padding frames are 0, never ``normalise(0)``
    (utils.py:206-207).  ``kind`` is 'mvn' (p0=mean, p1=std_dev) or 'minmax' (p0=mmin, p1=mmax).
    """
    if kind == 'mvn':
        normed = normalise_mvn(sequence_feature, p0, p1)
    elif kind == 'minmax':
        normed = normalise_minmax(sequence_feature, p0, p1)
    else:
        raise ValueError(kind)
    return upsample_to_repetitions(normed.astype(F32), repeats)


# ----------------------------------------------------------------------------------------------------------------
# a6 / a7  losses.sequence_loss -> mse / bce (+ an L1 sibling)  (morgana/losses.py:9-56)
# ----------------------------------------------------------------------------------------------------------------

def _pointwise_loss(kind, predictions, targets):
    p = np.asarray(predictions, dtype=np.float64)
    y = np.asarray(targets, dtype=np.float64)
    if kind == 'mse':      # losses.py:49-51, F.mse_loss(reduction='none')
        return (p - y) ** 2
    if kind == 'l1':       # same wrapper applied to F.l1_loss; the north star's "masked L1"
        return np.abs(p - y)
    if kind == 'bce':      # losses.py:54-56, F.binary_cross_entropy clamps each log at -100
        with np.errstate(divide='ignore', invalid='ignore'):
            log_p = np.maximum(np.log(p), -100.)
            log_1mp = np.maximum(np.log1p(-p), -100.)
        return -(y * log_p + (1. - y) * log_1mp)
    raise ValueError(kind)


def masked_loss(predictions, targets, seq_len=None, kind='mse'):
    """``mean over (b, d) of [ sum_{t < n_b} l(p, y) / n_b ]`` (losses.py:29-44); ``seq_len=None`` divides by T.

    A zero-length utterance gives 0/0 = nan, as in the reference.  Returns a python float (fp64).
    """
    loss = _pointwise_loss(kind, predictions, targets)
    batch_size, max_len, feat_dim = loss.shape
    if seq_len is None:
        per_utt = loss.sum(axis=1) / max_len
    else:
        seq_len = np.asarray(seq_len)
        if seq_len.shape != (batch_size,):
            raise RuntimeError('seq_len must have shape (batch_size,)')  # a (B, 1) seq_len fails to broadcast
        mask = sequence_mask(seq_len, max_len, dtype=np.float64)
        with np.errstate(divide='ignore', invalid='ignore'):
            per_utt = (loss * mask).sum(axis=1) / mask.sum(axis=1)
    return float(per_utt.mean()) if per_utt.size else float('nan')


def masked_loss_grad(predictions, targets, seq_len=None, kind='mse', grad_output=1.0):
    """d(masked_loss)/d(predictions): ``l'(p, y) [t < n_b] / (n_b B D)`` (autograd of losses.py:29-44)."""
    p = np.asarray(predictions, dtype=np.float64)
    y = np.asarray(targets, dtype=np.float64)
    batch_size, max_len, feat_dim = p.shape
    if kind == 'mse':
        dl = 2. * (p - y)
    elif kind == 'l1':
        dl = np.sign(p - y)
    elif kind == 'bce':   # ATen binary_cross_entropy_backward: (p - y) / max((1 - p) p, 1e-12)
        dl = (p - y) / np.maximum((1. - p) * p, 1e-12)
    else:
        raise ValueError(kind)
    if seq_len is None:
        weight = np.full((batch_size, 1, 1), 1. / max_len)
        mask = 1.
    else:
        seq_len = np.asarray(seq_len)
        mask = sequence_mask(seq_len, max_len, dtype=np.float64)
        with np.errstate(divide='ignore'):
            weight = (1. / seq_len.astype(np.float64))[:, None, None]
    grad = dl * mask * weight / (batch_size * feat_dim) * grad_output
    return grad.astype(np.asarray(predictions).dtype)


def cross_entropy_loss(logits, targets, seq_len=None):
    """losses.ce (losses.py:59-61): F.cross_entropy over the class axis gives one value per frame (feature axis of size
    1), then the masked sequence loss of :func:`masked_loss`.  Returns (loss, gradient w.r.t. the logits)."""
    x = np.asarray(logits, dtype=np.float64)
    t = np.asarray(targets)
    batch_size, max_len, n_classes = x.shape
    shifted = x - x.max(axis=-1, keepdims=True)
    log_soft = shifted - np.log(np.exp(shifted).sum(axis=-1, keepdims=True))
    onehot = np.eye(n_classes)[t]
    per_frame = -(log_soft * onehot).sum(axis=-1, keepdims=True)                 # (B, T, 1)
    if seq_len is None:
        mask, frames = np.ones((batch_size, max_len, 1)), np.full((batch_size, 1, 1), float(max_len))
    else:
        mask = sequence_mask(np.asarray(seq_len), max_len, dtype=np.float64)
        frames = np.asarray(seq_len, dtype=np.float64)[:, None, None]
    loss = float(((per_frame * mask).sum(axis=1) / frames[:, 0]).mean())
    grad = (np.exp(log_soft) - onehot) * mask / frames / batch_size
    return loss, grad.astype(np.asarray(logits).dtype)


# ----------------------------------------------------------------------------------------------------------------
# a8 - a12  streaming-metric accumulators  (morgana/metrics.py:359-694)
# Each returns the (sum, count) increment that one ``accumulate`` call adds to the metric's state.
# ----------------------------------------------------------------------------------------------------------------

def _masked_sum_count(values, seq_len):
    """Mean.accumulate (metrics.py:383-394): with seq_len the count is the number of valid FRAMES (Q2)."""
    values = np.asarray(values)
    if seq_len is None:
        return values.astype(np.float64).sum(), float(values.size)
    mask = sequence_mask(np.asarray(seq_len), values.shape[1], dtype=np.float64)
    return (values.astype(np.float64) * mask).sum(), float(mask.sum())


def mean_acc(tensor, seq_len=None):
    """Mean (metrics.py:383-397)."""
    return _masked_sum_count(tensor, seq_len)


def rmse_acc(target, pred, seq_len=None):
    """RMSE (metrics.py:492-495): squared difference into Mean."""
    diff = np.asarray(target, np.float64) - np.asarray(pred, np.float64)
    return _masked_sum_count(diff ** 2, seq_len)


def mae_acc(target, pred, seq_len=None):
    """MAE (metrics.py:574-576)."""
    diff = np.asarray(target, np.float64) - np.asarray(pred, np.float64)
    return _masked_sum_count(np.abs(diff), seq_len)


def melcep_acc(target, pred, seq_len=None):
    """MelCepDistortion (metrics.py:690-694): RMSE without coefficient 0."""
    return rmse_acc(np.asarray(target)[..., 1:], np.asarray(pred)[..., 1:], seq_len)


def distortion_acc(target, pred, seq_len=None):
    """Distortion (metrics.py:657-665): per-frame Euclidean distance (feature axis collapsed) into Mean."""
    diff = np.asarray(target, np.float64) - np.asarray(pred, np.float64)
    root = np.sqrt((diff ** 2).sum(axis=-1, keepdims=True))
    return _masked_sum_count(root, seq_len)


def f0_acc(f0_target, f0_pred, is_voiced, seq_len=None):
    """F0Distortion (metrics.py:597-609): squared error over frames that are voiced AND inside the utterance."""
    mask = np.asarray(is_voiced).astype(np.float64)
    if seq_len is not None:
        mask = mask * sequence_mask(np.asarray(seq_len), np.asarray(f0_target).shape[1], dtype=np.float64)
    diff = np.asarray(f0_target, np.float64) - np.asarray(f0_pred, np.float64)
    return (diff ** 2 * mask).sum(), float(mask.sum())


def lf0_acc(lf0_target, lf0_pred, is_voiced, seq_len=None):
    """LF0Distortion (metrics.py:630-634): ``exp`` of both streams in fp32 (as torch.exp does), then F0Distortion."""
    f0_target = np.exp(np.asarray(lf0_target, F32).astype(np.float64)).astype(F32)
    f0_pred = np.exp(np.asarray(lf0_pred, F32).astype(np.float64)).astype(F32)
    return f0_acc(f0_target, f0_pred, is_voiced, seq_len)


def error_acc(target, pred, seq_len=None):
    """Error (metrics.py:547-549): ``target ^ pred`` on bool/uint8 into Mean; the sum is an exact integer."""
    return _masked_sum_count(np.bitwise_xor(np.asarray(target), np.asarray(pred)), seq_len)


def accuracy_acc(target, pred, seq_len=None):
    """Accuracy (metrics.py:520-522): ``target & pred`` on bool/uint8 into Mean."""
    return _masked_sum_count(np.bitwise_and(np.asarray(target), np.asarray(pred)), seq_len)


def variance_acc(tensor, seq_len=None):
    """Variance.accumulate (metrics.py:427-442): increments of (sum, sum_square, count)."""
    total, count = _masked_sum_count(tensor, seq_len)
    total_sq, _ = _masked_sum_count(np.asarray(tensor, np.float64) ** 2, seq_len)
    return total, total_sq, count


def variance_result(total, total_sq, count):
    """Variance.result (metrics.py:444-446)."""
    count = count + 1e-8
    return (total_sq - total ** 2 / count) / count


def tensor_history(batches, feat_dim, max_len=None):
    """TensorHistory (metrics.py:302-321): valid rows of every accumulate call concatenated, last `max_len` kept.
    `batches` is a list of (tensor, seq_len or None)."""
    history = np.zeros((0, feat_dim), dtype=np.asarray(batches[0][0]).dtype)
    for tensor, seq_len in batches:
        rows = np.asarray(tensor).reshape(-1, feat_dim) if seq_len is None else batched_masked_select(tensor, seq_len)
        history = np.concatenate([history, rows])
        if max_len is not None:
            history = history[-max_len:]
    return history


def mean_result(total, count):
    """Mean.result (metrics.py:396-397)."""
    return total / (count + 1e-8)


def rmse_result(total, count):
    """RMSE.result (metrics.py:497-499)."""
    return (total / (count + 1e-8)) ** 0.5


DISTORTION_DB_CONST = 10. / np.log(10.) * np.sqrt(2.)   # metrics.py:652


# ----------------------------------------------------------------------------------------------------------------
# a13  utils.ExponentialMovingAverage._update_param  (morgana/utils.py:443-448)
# ----------------------------------------------------------------------------------------------------------------

def ema_update(shadow, param, decay):
    """``delta = s - x; s -= (1 - decay) * delta`` in place, each step rounded to fp32 (bit-exact with ATen).

    ``1 - decay`` is formed in double by Python and rounded to fp32 when it meets the fp32 tensor (utils.py:448).
    """
    assert shadow.dtype == F32 and param.dtype == F32
    one_minus_decay = F32(1.0 - decay)
    delta = shadow - param
    shadow -= one_minus_decay * delta
    return shadow


# ----------------------------------------------------------------------------------------------------------------
# a14  nn.Linear (+ Sigmoid)  (README.rst:65-73, models/RNN_SPSS.py:33-41)
# ----------------------------------------------------------------------------------------------------------------

def kld_standard_normal(mean, log_variance):
    """``mean over rows of -0.5 * sum_d (1 + lv - mean**2 - exp(lv))`` in float64, with its gradients w.r.t. both operands
    (reference morgana/losses.py:64-67).  Returns ``(loss, grad_mean, grad_log_variance)``."""
    m, lv = np.asarray(mean, np.float64), np.asarray(log_variance, np.float64)
    rows = m.size // m.shape[-1]
    loss = float((-0.5 * (1. + lv - m ** 2 - np.exp(lv)).sum(-1)).mean())
    return loss, m / rows, 0.5 * (np.exp(lv) - 1.) / rows


def both_voiced_mask(*sequence_features, dtype=np.uint8):
    """``prod_k (feature_k != 0)`` cast to ``dtype`` (reference morgana/utils.py:169-172; NaN != 0 is True)."""
    voiced = np.ones(np.asarray(sequence_features[0]).shape, dtype=bool)
    for feature in sequence_features:
        voiced &= ~(np.asarray(feature) == 0.)
    return voiced.astype(dtype)


def linear(x, weight, bias=None, act=None):
    """``y = x W^T + b`` with an optional sigmoid, carried in float64 (the GEMM parity target)."""
    y = np.asarray(x, np.float64) @ np.asarray(weight, np.float64).T
    if bias is not None:
        y = y + np.asarray(bias, np.float64)
    if act == 'sigmoid':
        y = 1. / (1. + np.exp(-y))
    elif act is not None:
        raise ValueError(act)
    return y



def sigmoid_grad(grad_y, y):
    """ATen's ``sigmoid_backward``: ``(grad_y * (1 - y)) * y`` with every operation rounded in the operands' dtype -- what
    autograd runs for nn.Sigmoid in the models' training step (reference README.rst:65-73, experiment_builder.py:470-479)."""
    grad_y, y = np.asarray(grad_y), np.asarray(y)
    one = np.asarray(1, y.dtype)
    return (grad_y * (one - y)) * y


def linear_wgrad(g, x):
    """Weight gradient of ``y = x W^T + b``: ``g^T x`` summed over the frame axis, in float64; the bias gradient is
    ``g.sum(0)`` (autograd of nn.Linear in the reference's training step, experiment_builder.py:470-479)."""
    return np.asarray(g, np.float64).T @ np.asarray(x, np.float64)

# ----------------------------------------------------------------------------------------------------------------
# "next" row 1: viz.synthesis.MLPG  (morgana/viz/synthesis.py:79-180)
# PARITY UNPINNED: the reference solves the system with `bandmat` (unpinned in setup.py:12, absent from this image and
# from /root/reference), so this restatement is checked against the definition only -- dense window matrices built as
# synthesis.py:8-36 describes them, numpy's dense fp64 solver -- not against outputs of the reference itself.
# ----------------------------------------------------------------------------------------------------------------

MLPG_WINDOWS = [(0, 0, np.array([1.0])), (1, 1, np.array([-0.5, 0.0, 0.5])), (1, 1, np.array([1.0, -2.0, 1.0]))]   # :122-127


def _window_matrix(left, right, coeffs, n_frames):
    """frames x frames Toeplitz band matrix whose row t holds `coeffs` at columns t-left .. t+right (synthesis.py:24-27)."""
    mat = np.zeros((n_frames, n_frames))
    for j, c in enumerate(coeffs):
        off = j - left
        idx = np.arange(max(0, -off), min(n_frames, n_frames - off))
        mat[idx, idx + off] = c
    return mat


def mlpg(means, variances, padding_size=0, seq_len=None):
    """Most probable static trajectory given means / variances of [static | delta | delta-delta] features.

    means: (B, T, 3 F); variances: (B, T, 3 F) or (3 F,) global; returns (B, T, F) float64, zeros past seq_len.
    """
    means = np.asarray(means, np.float64)
    batch_size, n_frames, dim3 = means.shape
    feat_dim = dim3 // 3
    variances = np.asarray(variances, np.float64)
    if variances.ndim == 1:
        variances = np.broadcast_to(variances, means.shape)
    if seq_len is None:
        seq_len = [n_frames] * batch_size
    out = np.zeros((batch_size, n_frames, feat_dim))
    idx_base = np.arange(3) * feat_dim

    def pad(x, n):   # synthesis.py:114-120: replicate the edge frames as burn-in
        return np.concatenate([np.repeat(x[:1], n, axis=0), x, np.repeat(x[-1:], n, axis=0)], axis=0)

    for i in range(batch_size):
        n = int(seq_len[i])
        if n == 0:
            continue
        m_i, v_i = pad(means[i, :n], padding_size), pad(variances[i, :n], padding_size)
        length = n + 2 * padding_size
        wins = [_window_matrix(l, u, c, length) for l, u, c in MLPG_WINDOWS]
        for d in range(feat_dim):
            mu, var = m_i[:, idx_base + d], v_i[:, idx_base + d]
            b = sum(w.T @ (mu[:, k] / var[:, k]) for k, w in enumerate(wins))            # synthesis.py:44
            prec = sum(w.T @ np.diag(1. / var[:, k]) @ w for k, w in enumerate(wins))    # synthesis.py:47
            traj = np.linalg.solve(prec, b)                                               # synthesis.py:168
            out[i, :n, d] = traj[padding_size:length - padding_size]
    return out


def mlpg_banded(means, variances, padding_size=0, seq_len=None):
    """The same trajectories through the reference's loop structure (synthesis.py:153-171: one banded Cholesky solve per
    utterance and static dimension) with scipy's `solveh_banded` standing in for the absent `bandmat`.  Global variances
    (3 F,) only.  This is the CPU cost model of MLPG inside predict(); `mlpg` above is the definition it is tested against.
    """
    import scipy.linalg as sl
    means = np.asarray(means, np.float64)
    batch_size, n_frames, dim3 = means.shape
    feat_dim = dim3 // 3
    tau = 1. / np.asarray(variances, np.float64)
    if seq_len is None:
        seq_len = [n_frames] * batch_size
    out = np.zeros((batch_size, n_frames, feat_dim))
    pad = padding_size
    for i in range(batch_size):
        n = int(seq_len[i])
        if n == 0:
            continue
        length = n + 2 * pad
        mu = np.pad(means[i, :n], ((pad, pad), (0, 0)), mode='edge')
        for d in range(feat_dim):
            t0, t1, t2 = tau[d], tau[feat_dim + d], tau[2 * feat_dim + d]
            b0, b1, b2 = mu[:, d] * t0, mu[:, feat_dim + d] * t1, mu[:, 2 * feat_dim + d] * t2
            rhs = b0 - 2 * b2                                   # W^T (mu / var) for windows [1], [-.5, 0, .5], [1, -2, 1]
            rhs[1:] += 0.5 * b1[:-1] + b2[:-1]
            rhs[:-1] += -0.5 * b1[1:] + b2[1:]
            ab = np.zeros((3, length))                          # upper band storage of W^T diag(1 / var) W
            ab[2] = t0 + 4 * t2
            ab[2, 1:] += 0.25 * t1 + t2
            ab[2, :-1] += 0.25 * t1 + t2
            ab[1, 1:] = -4 * t2
            ab[0, 2:] = -0.25 * t1 + t2
            out[i, :n, d] = sl.solveh_banded(ab, rhs)[pad:length - pad]
    return out


# ----------------------------------------------------------------------------------------------------------------
# "next" row 4: sibling segment operations  (morgana/utils.py:147-166, 231-330); pinned by tests/golden/segments.npz
# ----------------------------------------------------------------------------------------------------------------

def batched_masked_select(sequence_feature, seq_len):
    """Rows inside each utterance's length, utterance after utterance (utils.py:147-166)."""
    sequence_feature = np.asarray(sequence_feature)
    n = np.minimum(np.asarray(seq_len), sequence_feature.shape[1])
    return np.concatenate([sequence_feature[b, :max(int(n[b]), 0)] for b in range(sequence_feature.shape[0])], axis=0)


def get_segment_ends(sequence_feature, segment_lens):
    """Feature at the last frame of every segment; zero for empty segments (utils.py:287-330)."""
    sequence_feature = np.asarray(sequence_feature)
    batch_size, n_frames, feat_dim = sequence_feature.shape
    lens = np.asarray(segment_lens).reshape(batch_size, -1)
    ends = np.cumsum(lens, axis=1)
    out = np.zeros((batch_size, lens.shape[1], feat_dim), dtype=sequence_feature.dtype)
    for b in range(batch_size):
        for s in range(lens.shape[1]):
            if lens[b, s] > 0 and ends[b, s] - 1 < n_frames:
                out[b, s] = sequence_feature[b, ends[b, s] - 1]
    return out


def split_to_segments(sequence_feature, segment_lens):
    """(B, T, D) -> (B, S, longest segment, D), each segment's frames then zeros (utils.py:231-284)."""
    sequence_feature = np.asarray(sequence_feature)
    batch_size, n_frames, feat_dim = sequence_feature.shape
    lens = np.asarray(segment_lens).reshape(batch_size, -1)
    longest = int(lens.max()) if lens.size else 0
    out = np.zeros((batch_size, lens.shape[1], longest, feat_dim), dtype=sequence_feature.dtype)
    for b in range(batch_size):
        start = 0
        for s in range(lens.shape[1]):
            stop = start + int(lens[b, s])
            take = max(0, min(stop, n_frames) - start)
            out[b, s, :take] = sequence_feature[b, start:start + take]
            start = stop
    return out
