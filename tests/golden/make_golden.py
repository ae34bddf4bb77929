"""Generate the golden fixtures in this directory by running the UNMODIFIED reference on seeded inputs.

Run in the build container only (``/root/reference`` must exist):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference ships no golden vectors of its own, so these fixtures -- outputs of the
reference's own functions under torch 2.11.0+cu128 / numpy 2.3.5 on CPU -- are what pins the oracle and the
CUDA kernels.  Every array is stored with the inputs that produced it, so tests need nothing but the .npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference  # noqa: E402

morgana = import_reference()
utils, losses, metrics, data = morgana.utils, morgana.losses, morgana.metrics, morgana.data

SEED = 1234


def gen(seed_offset=0):
    g = torch.Generator()
    g.manual_seed(SEED + seed_offset)
    return g


def rand_durations(g, batch_size, n_items, max_dur, p_zero=0.2, ragged=True):
    dur = torch.randint(0, max_dur + 1, (batch_size, n_items), generator=g)
    dur[torch.rand(batch_size, n_items, generator=g) < p_zero] = 0
    if ragged:
        n_valid = torch.randint(0, n_items + 1, (batch_size,), generator=g)
        dur[torch.arange(n_items)[None, :] >= n_valid[:, None]] = 0
    return dur


def upsample_cases():
    out = {}
    specs = [  # name, B, P, D, max_dur, dtype
        ('tiny_f32', 3, 5, 4, 3, torch.float32),
        ('odd_f32', 5, 9, 7, 6, torch.float32),
        ('lab600_f32', 4, 12, 600, 6, torch.float32),
        ('wide609_f32', 2, 6, 609, 5, torch.float32),
        ('d1_f32', 6, 11, 1, 9, torch.float32),
        ('f64', 3, 7, 5, 4, torch.float64),
        ('i64', 3, 7, 3, 4, torch.int64),
        ('f16', 3, 7, 6, 4, torch.float16),
        ('u8', 2, 5, 10, 4, torch.uint8),
    ]
    for i, (name, B, P, D, max_dur, dtype) in enumerate(specs):
        g = gen(i)
        dur = rand_durations(g, B, P, max_dur)
        if dtype.is_floating_point:
            x = torch.randn(B, P, D, generator=g).to(dtype)
        else:
            x = torch.randint(0, 100, (B, P, D), generator=g).to(dtype)
        y = utils.upsample_to_repetitions(x, dur[:, :, None])
        assert y.is_contiguous()
        out['ups_%s_x' % name] = x.numpy()
        out['ups_%s_dur' % name] = dur.numpy()
        out['ups_%s_out' % name] = y.numpy()

    # Hand-written edge cases.
    x = torch.arange(24, dtype=torch.float32).reshape(2, 4, 3)
    dur = torch.tensor([[2, 0, 1, 0], [0, 0, 0, 0]])
    out['ups_emptyrow_x'], out['ups_emptyrow_dur'] = x.numpy(), dur.numpy()
    out['ups_emptyrow_out'] = utils.upsample_to_repetitions(x, dur[:, :, None]).numpy()
    dur0 = torch.zeros(2, 4, dtype=torch.int64)
    out['ups_allzero_x'], out['ups_allzero_dur'] = x.numpy(), dur0.numpy()
    out['ups_allzero_out'] = utils.upsample_to_repetitions(x, dur0[:, :, None]).numpy()
    assert out['ups_allzero_out'].shape == (2, 0, 3)
    # 2-D repeats and int32 repeats are accepted.
    dur32 = torch.tensor([[1, 2, 0, 3], [4, 0, 0, 1]], dtype=torch.int32)
    out['ups_int32dur_x'], out['ups_int32dur_dur'] = x.numpy(), dur32.numpy()
    out['ups_int32dur_out'] = utils.upsample_to_repetitions(x, dur32).numpy()

    # Backward: gradient w.r.t. the item-rate input (IndexBackward0).
    g = gen(50)
    dur = rand_durations(g, 4, 8, 5)
    x = torch.randn(4, 8, 6, generator=g, requires_grad=True)
    y = utils.upsample_to_repetitions(x, dur[:, :, None])
    grad_out = torch.randn(y.shape, generator=g)
    y.backward(grad_out)
    out['upsbwd_x'], out['upsbwd_dur'] = x.detach().numpy(), dur.numpy()
    out['upsbwd_grad_out'], out['upsbwd_grad_x'] = grad_out.numpy(), x.grad.numpy()
    return out


def sequence_mask_cases():
    out = {}
    seq_len = torch.tensor([3, 0, 5, 2])
    out['mask_seq_len'] = seq_len.numpy()
    out['mask_default'] = utils.sequence_mask(seq_len).numpy()
    out['mask_len7_f32'] = utils.sequence_mask(seq_len, max_len=7, dtype=torch.float32).numpy()
    out['mask_len4_u8'] = utils.sequence_mask(seq_len, max_len=4).numpy()
    return out


def voiced_mask_cases():
    """utils.both_voiced_mask (reference utils.py:169-172)."""
    g = gen(77)
    out = {}
    a = torch.randn(5, 23, 1, generator=g)
    b = torch.randn(5, 23, 1, generator=g)
    c = torch.randn(5, 23, 1, generator=g)
    a[torch.rand(5, 23, 1, generator=g) < 0.3] = 0.
    b[torch.rand(5, 23, 1, generator=g) < 0.3] = 0.
    c[torch.rand(5, 23, 1, generator=g) < 0.3] = 0.
    a[0, 0, 0], b[0, 1, 0] = float('nan'), -0.
    out['voiced_a'], out['voiced_b'], out['voiced_c'] = a.numpy(), b.numpy(), c.numpy()
    out['voiced_mask_ab'] = utils.both_voiced_mask(a, b).numpy()
    out['voiced_mask_abc_f32'] = utils.both_voiced_mask(a, b, c, dtype=torch.FloatTensor).numpy()
    out['voiced_mask_a'] = utils.both_voiced_mask(a).numpy()
    return out


def kld_cases():
    """losses.KLD_standard_normal (reference losses.py:64-67) with autograd gradients of both operands."""
    g = gen(78)
    out = {}
    for tag, shape in (('2d', (32, 16)), ('3d', (4, 37, 5)), ('row', (1, 3))):
        mean = torch.randn(*shape, generator=g).requires_grad_()
        lv = (0.5 * torch.randn(*shape, generator=g)).requires_grad_()
        loss = losses.KLD_standard_normal(mean, lv)
        (0.25 * loss).backward()
        out['kld_%s_mean' % tag], out['kld_%s_lv' % tag] = mean.detach().numpy(), lv.detach().numpy()
        out['kld_%s_loss' % tag] = loss.detach().numpy()
        out['kld_%s_grad_mean' % tag], out['kld_%s_grad_lv' % tag] = mean.grad.numpy(), lv.grad.numpy()
    return out


def sequence_loss_cases():
    """losses.sequence_loss (reference losses.py:9-47) around a loss the reference does not ship: smooth-L1 per element."""
    import torch.nn.functional as F

    @losses.sequence_loss
    def huber(predictions, targets):
        return F.smooth_l1_loss(predictions, targets, reduction='none')

    g = gen(79)
    out = {}
    p = (2. * torch.randn(6, 29, 7, generator=g)).requires_grad_()
    t = torch.randn(6, 29, 7, generator=g)
    n = torch.tensor([29, 1, 17, 5, 29, 11])
    out['seqloss_p'], out['seqloss_t'], out['seqloss_n'] = p.detach().numpy(), t.numpy(), n.numpy()
    loss = huber(p, t, seq_len=n)
    (3. * loss).backward()
    out['seqloss_masked'], out['seqloss_masked_grad'] = loss.detach().numpy(), p.grad.numpy().copy()
    p.grad = None
    loss = huber(p, t)
    loss.backward()
    out['seqloss_full'], out['seqloss_full_grad'] = loss.detach().numpy(), p.grad.numpy().copy()
    return out


def detach_cases():
    """utils.detach_batched_seqs (reference utils.py:66-102): padding removed per utterance, squeezed or not."""
    g = gen(80)
    out = {}
    x = torch.randn(5, 13, 4, generator=g)
    y = torch.randn(5, 13, 1, generator=g)
    n = torch.tensor([13, 1, 0, 7, 9])
    out['detach_x'], out['detach_y'], out['detach_n'] = x.numpy(), y.numpy(), n.numpy()
    xs, ys = utils.detach_batched_seqs(x, y, seq_len=n)
    raw = utils.detach_batched_seqs(y, seq_len=n, squeeze=False)
    for b in range(5):
        out['detach_x_%d' % b], out['detach_y_%d' % b], out['detach_y_raw_%d' % b] = xs[b], ys[b], raw[b]
    out['detach_full'] = utils.detach_batched_seqs(x)
    return out


def normaliser_cases():
    out = {}
    g = gen(100)
    for name, shape, D in [('btd', (3, 11, 600), 600), ('td', (13, 7), 7), ('btd187', (2, 9, 187), 187)]:
        x = torch.randn(*shape, generator=g) * 3 + 1
        mean = torch.randn(D, generator=g)
        std = torch.randn(D, generator=g).abs() + 0.1
        mmin = torch.randn(D, generator=g)
        mmax = mmin + torch.rand(D, generator=g) + 0.5
        mmax[::5] = mmin[::5]            # constant dims -> scale forced to 1 (data.py:581)
        mmax[1] = mmin[1] + 5e-9         # |scale| <= 1e-8 but non-zero
        out['norm_%s_x' % name] = x.numpy()
        out['norm_%s_mean' % name], out['norm_%s_std' % name] = mean.numpy(), std.numpy()
        out['norm_%s_mmin' % name], out['norm_%s_mmax' % name] = mmin.numpy(), mmax.numpy()
        out['norm_%s_mvn' % name] = data.normalise_mvn(x, mean, std).numpy()
        out['norm_%s_demvn' % name] = data.denormalise_mvn(x, mean, std).numpy()
        out['norm_%s_minmax' % name] = data.normalise_minmax(x, mmin, mmax).numpy()
        out['norm_%s_deminmax' % name] = data.denormalise_minmax(x, mmin, mmax).numpy()

    # Speaker-dependent parameters broadcast per utterance: params are (B, D) (data.py:460-501).
    x = torch.randn(4, 6, 5, generator=g)
    mean, std = torch.randn(4, 5, generator=g), torch.rand(4, 5, generator=g) + 0.2
    mmin = torch.randn(4, 5, generator=g)
    mmax = mmin + torch.rand(4, 5, generator=g) + 0.5
    out['norm_sd_x'] = x.numpy()
    out['norm_sd_mean'], out['norm_sd_std'] = mean.numpy(), std.numpy()
    out['norm_sd_mmin'], out['norm_sd_mmax'] = mmin.numpy(), mmax.numpy()
    out['norm_sd_mvn'] = data.normalise_mvn(x, mean, std).numpy()
    out['norm_sd_demvn'] = data.denormalise_mvn(x, mean, std).numpy()
    out['norm_sd_minmax'] = data.normalise_minmax(x, mmin, mmax).numpy()
    out['norm_sd_deminmax'] = data.denormalise_minmax(x, mmin, mmax).numpy()

    # The composition the fused kernel replaces: normalise at item rate, then upsample.
    g = gen(101)
    dur = rand_durations(g, 4, 10, 6)
    x = torch.rand(4, 10, 600, generator=g)
    mmin = torch.zeros(600)
    mmax = torch.rand(600, generator=g) + 0.5
    mmax[::7] = 0.
    mean, std = torch.randn(600, generator=g), torch.randn(600, generator=g).abs() + 0.1
    out['fused_x'], out['fused_dur'] = x.numpy(), dur.numpy()
    out['fused_mmin'], out['fused_mmax'] = mmin.numpy(), mmax.numpy()
    out['fused_mean'], out['fused_std'] = mean.numpy(), std.numpy()
    out['fused_minmax_out'] = utils.upsample_to_repetitions(data.normalise_minmax(x, mmin, mmax), dur[:, :, None]).numpy()
    out['fused_mvn_out'] = utils.upsample_to_repetitions(data.normalise_mvn(x, mean, std), dur[:, :, None]).numpy()
    return out


def loss_cases():
    out = {}
    g = gen(200)
    for name, B, T, D in [('small', 4, 9, 5), ('wide', 3, 17, 187), ('d1', 6, 23, 1)]:
        seq_len = torch.randint(1, T + 1, (B,), generator=g)
        seq_len[0] = T
        tgt = torch.randn(B, T, D, generator=g)
        pred = (tgt + 0.3 * torch.randn(B, T, D, generator=g)).requires_grad_()
        out['loss_%s_pred' % name], out['loss_%s_tgt' % name] = pred.detach().numpy(), tgt.numpy()
        out['loss_%s_seq_len' % name] = seq_len.numpy()
        for kind, fn in [('mse', losses.mse)]:
            for masked in (True, False):
                pred.grad = None
                value = fn(pred, tgt, seq_len if masked else None)
                value.backward()
                tag = 'loss_%s_%s_%s' % (name, kind, 'masked' if masked else 'full')
                out[tag], out[tag + '_grad'] = value.detach().numpy(), pred.grad.numpy().copy()

        prob = torch.sigmoid(torch.randn(B, T, D, generator=g))
        prob[0, 0, 0], prob[0, 1, 0] = 0., 1.          # exercise the -100 log clamp
        prob = prob.requires_grad_()
        label = (torch.rand(B, T, D, generator=g) < 0.6).float()
        label[0, 0, 0], label[0, 1, 0] = 1., 0.
        out['loss_%s_prob' % name], out['loss_%s_label' % name] = prob.detach().numpy(), label.numpy()
        for masked in (True, False):
            prob.grad = None
            value = losses.bce(prob, label, seq_len if masked else None)
            value.backward()
            tag = 'loss_%s_bce_%s' % (name, 'masked' if masked else 'full')
            out[tag], out[tag + '_grad'] = value.detach().numpy(), prob.grad.numpy().copy()

    # losses.ce (losses.py:59-61): logits (B, T, C), class indices (B, T)
    for name, B, T, C in [('ce_small', 4, 9, 5), ('ce_wide', 3, 21, 64)]:
        seq_len = torch.randint(1, T + 1, (B,), generator=g)
        seq_len[0] = T
        logits = (3. * torch.randn(B, T, C, generator=g)).requires_grad_()
        classes = torch.randint(0, C, (B, T), generator=g)
        out['loss_%s_logits' % name], out['loss_%s_classes' % name] = logits.detach().numpy(), classes.numpy()
        out['loss_%s_seq_len' % name] = seq_len.numpy()
        for masked in (True, False):
            logits.grad = None
            value = losses.ce(logits, classes, seq_len if masked else None)
            value.backward()
            tag = 'loss_%s_%s' % (name, 'masked' if masked else 'full')
            out[tag], out[tag + '_grad'] = value.detach().numpy(), logits.grad.numpy().copy()

    # Zero-length utterance -> nan.
    pred, tgt = torch.randn(2, 4, 3, generator=g), torch.randn(2, 4, 3, generator=g)
    out['loss_zero_len'] = losses.mse(pred, tgt, torch.tensor([0, 3])).numpy()
    return out


def metric_cases():
    out = {}
    g = gen(300)
    B, T, D = 5, 19, 6
    batches = []
    for _ in range(2):   # two accumulate calls: the state is a running sum
        seq_len = torch.randint(1, T + 1, (B,), generator=g)
        tgt = torch.randn(B, T, D, generator=g)
        pred = tgt + 0.2 * torch.randn(B, T, D, generator=g)
        lf0_t = 5 + 0.3 * torch.randn(B, T, 1, generator=g)
        lf0_p = lf0_t + 0.05 * torch.randn(B, T, 1, generator=g)
        voiced = torch.rand(B, T, 1, generator=g) < 0.6
        bits_t = torch.rand(B, T, 1, generator=g) < 0.5
        bits_p = torch.rand(B, T, 1, generator=g) < 0.5
        batches.append((seq_len, tgt, pred, lf0_t, lf0_p, voiced, bits_t, bits_p))
    for i, (seq_len, tgt, pred, lf0_t, lf0_p, voiced, bits_t, bits_p) in enumerate(batches):
        out['met_b%d_seq_len' % i] = seq_len.numpy()
        out['met_b%d_tgt' % i], out['met_b%d_pred' % i] = tgt.numpy(), pred.numpy()
        out['met_b%d_lf0_t' % i], out['met_b%d_lf0_p' % i] = lf0_t.numpy(), lf0_p.numpy()
        out['met_b%d_voiced' % i] = voiced.numpy()
        out['met_b%d_bits_t' % i], out['met_b%d_bits_p' % i] = bits_t.numpy(), bits_p.numpy()

    def run(metric, args_of_batch, tag):
        for masked in (True, False):
            metric.reset_state()
            for batch in batches:
                args = args_of_batch(batch)
                metric.accumulate(*args, seq_len=batch[0] if masked else None)
            key = 'met_%s_%s' % (tag, 'masked' if masked else 'full')
            out[key + '_sum'] = np.asarray(torch.as_tensor(metric.sum).numpy())
            out[key + '_count'] = np.asarray(float(metric.count))
            out[key + '_result'] = np.asarray(torch.as_tensor(metric.result()).numpy())

    run(metrics.Mean(), lambda b: (b[1],), 'mean')
    run(metrics.RMSE(), lambda b: (b[1], b[2]), 'rmse')
    run(metrics.MAE(), lambda b: (b[1], b[2]), 'mae')
    run(metrics.MelCepDistortion(), lambda b: (b[1], b[2]), 'melcep')
    run(metrics.Distortion(), lambda b: (b[1], b[2]), 'distortion')
    # F0 metrics mutate a float `is_voiced` in place (Q4): hand them a fresh clone every call.
    run(metrics.F0Distortion(), lambda b: (b[3].exp(), b[4].exp(), b[5].clone()), 'f0')
    run(metrics.LF0Distortion(), lambda b: (b[3], b[4], b[5].clone()), 'lf0')
    run(metrics.LF0Distortion(), lambda b: (b[3], b[4], b[5].float()), 'lf0_floatmask')
    run(metrics.Error(), lambda b: (b[6], b[7]), 'error')
    run(metrics.Accuracy(), lambda b: (b[6], b[7]), 'accuracy')
    run(metrics.Error(), lambda b: (b[6].to(torch.uint8), b[7].to(torch.uint8)), 'error_u8')
    # VUV accuracy as the acoustic model computes it (models/RNN_SPSS.py:127).
    run(metrics.Mean(), lambda b: ((b[6] == b[7]).type(torch.float),), 'vuvacc')
    return out


def speaker_dependent_cases():
    """SpeakerDependent{MeanVariance,MinMax}Normaliser (data.py:388-531, 567-576, 619-628) on a 4-utterance batch of three
    speakers, a one-speaker batch (parameters squeezed to (feat_dim,), :500-501) and the NumPy path."""
    out = {}
    g = gen(450)
    D, speakers = 7, ['spk_a', 'spk_b', 'spk_c']
    batch_ids = ['spk_b', 'spk_a', 'spk_b', 'spk_c']
    x = torch.randn(4, 9, D, generator=g)
    out['sdn_x'], out['sdn_batch_ids'], out['sdn_speakers'] = x.numpy(), np.asarray(batch_ids), np.asarray(speakers)
    mvn = data.SpeakerDependentMeanVarianceNormaliser('lf0', 'speakers.scp', use_deltas=True)
    mm = data.SpeakerDependentMinMaxNormaliser('lab', 'speakers.scp')
    for i, spk in enumerate(speakers):
        p = {'mean': torch.randn(D, generator=g).numpy(), 'std_dev': (torch.rand(D, generator=g) + 0.2).numpy()}
        dp = {'mean': torch.randn(D, generator=g).numpy(), 'std_dev': (torch.rand(D, generator=g) + 0.2).numpy()}
        q = {'mmin': torch.randn(D, generator=g).numpy(), 'mmax': (torch.randn(D, generator=g) + 3.).numpy()}
        q['mmax'][i] = q['mmin'][i]                                   # a constant dimension per speaker (scale forced to 1)
        for name, params in (('mvn', p), ('mvn_deltas', dp), ('minmax', q)):
            for k, v in params.items():
                out['sdn_%s_%s_%s' % (name, spk, k)] = v
        mvn.params[spk], mvn.params_torch[spk] = p, {k: torch.tensor(v) for k, v in p.items()}
        mvn.delta_params[spk], mvn.delta_params_torch[spk] = dp, {k: torch.tensor(v) for k, v in dp.items()}
        mm.params[spk], mm.params_torch[spk] = q, {k: torch.tensor(v) for k, v in q.items()}
    out['sdn_mvn_norm'] = mvn.normalise(x, batch_ids).numpy()
    out['sdn_mvn_denorm'] = mvn.denormalise(x, batch_ids).numpy()
    out['sdn_mvn_norm_deltas'] = mvn.normalise(x, batch_ids, deltas=True).numpy()
    out['sdn_minmax_norm'] = mm.normalise(x, batch_ids).numpy()
    out['sdn_minmax_denorm'] = mm.denormalise(x, batch_ids).numpy()
    out['sdn_mvn_norm_single'] = mvn.normalise(x[1], 'spk_c').numpy()            # (T, D) feature, one speaker
    # the DataLoader-worker path: one utterance, one speaker, NumPy (a NumPy batch of several speakers fails in the
    # reference: np.squeeze(0) on a (batch_size, feat_dim) array, data.py:500-501)
    out['sdn_minmax_norm_numpy'] = mm.normalise(x[2].numpy(), 'spk_b')
    return out


def metric_extra_cases():
    """Variance / StandardDeviation (metrics.py:400-471), TensorHistory (:263-356) and the Handler container (:52-185)."""
    out = {}
    g = gen(350)
    B, T, D = 4, 13, 5
    batches = []
    for i in range(2):
        seq_len = torch.randint(1, T + 1, (B,), generator=g)
        x = 2. + torch.randn(B, T, D, generator=g)
        y = x + 0.3 * torch.randn(B, T, D, generator=g)
        batches.append((seq_len, x, y))
        out['mx_b%d_seq_len' % i], out['mx_b%d_x' % i], out['mx_b%d_y' % i] = seq_len.numpy(), x.numpy(), y.numpy()
    for masked in (True, False):
        tag = 'masked' if masked else 'full'
        var, std = metrics.Variance(), metrics.StandardDeviation()
        hist, hist_short = metrics.TensorHistory(D), metrics.TensorHistory(D, max_len=7)
        for seq_len, x, _ in batches:
            # Variance zeroes the padding of its input in place (metrics.py:436): give it a clone.
            var.accumulate(x.clone(), seq_len=seq_len if masked else None)
            std.accumulate(x.clone(), seq_len=seq_len if masked else None)
            hist.accumulate(x, seq_len=seq_len if masked else None)
            hist_short.accumulate(x, seq_len=seq_len if masked else None)
        out['mx_var_%s_sum' % tag] = var.sum.numpy()
        out['mx_var_%s_sum_square' % tag] = var.sum_square.numpy()
        out['mx_var_%s_count' % tag] = np.asarray(float(var.count))
        out['mx_var_%s_result' % tag] = var.result().numpy()
        out['mx_std_%s_result' % tag] = std.result().numpy()
        out['mx_hist_%s' % tag] = hist.result().numpy()
        out['mx_hist_short_%s' % tag] = hist_short.result().numpy()
    # Handler: the container models use as `self.metrics` (base_models.py), driven as models/RNN_SPSS.py:124-129 does.
    handler = metrics.Handler(loss=metrics.Mean())
    handler.add_metrics('all', err=metrics.RMSE(), mae=metrics.MAE())
    handler.add_metrics('valid', spread=metrics.Variance())
    for mode in ('train', 'valid'):
        handler.reset_state(mode)
        for seq_len, x, y in batches:
            kwargs = dict(err=(x, y, seq_len), mae=(x, y, {'seq_len': seq_len}), loss=torch.mean(x))   # 0-dim, as experiment_builder.py:484
            if mode == 'valid':
                kwargs['spread'] = (x.clone(), seq_len)
            handler.accumulate(mode, **kwargs)
        for name, value in handler.results_as_json_dict(mode).items():
            out['mx_handler_%s_%s' % (mode, name)] = np.asarray(value)
    out['mx_handler_train_names'] = np.asarray(sorted(handler['train'].keys()))
    out['mx_handler_valid_names'] = np.asarray(sorted(handler['valid'].keys()))
    return out


def ema_cases():
    out = {}
    g = gen(400)
    torch.manual_seed(SEED)
    shapes = [(7,), (5, 3), (64, 33), (1,), (1025,)]
    model = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes])
    ema_model = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes])
    decay = 0.999
    ema = utils.ExponentialMovingAverage(ema_model, decay)
    for i, p in enumerate(ema_model):
        out['ema_shadow0_%d' % i] = p.detach().numpy().copy()
    for step in range(3):
        with torch.no_grad():
            for p in model:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
        for i, p in enumerate(model):
            out['ema_param%d_%d' % (step, i)] = p.detach().numpy().copy()
        ema.update_params(model)
        for i, p in enumerate(ema_model):
            out['ema_shadow%d_%d' % (step + 1, i)] = p.detach().numpy().copy()
    out['ema_decay'] = np.asarray(decay)
    out['ema_n'] = np.asarray(len(shapes))
    return out


def linear_cases():
    out = {}
    g = gen(500)
    for name, M, K, N in [('readme_l1', 70, 600, 64), ('rnn_in', 33, 609, 48), ('out187', 45, 256, 187),
                          ('out1', 19, 32, 1)]:
        x = torch.rand(M, K, generator=g)
        w = torch.randn(N, K, generator=g) / K ** 0.5
        b = torch.randn(N, generator=g) * 0.1
        y = torch.nn.functional.linear(x, w, b)
        out['lin_%s_x' % name], out['lin_%s_w' % name], out['lin_%s_b' % name] = x.numpy(), w.numpy(), b.numpy()
        out['lin_%s_y' % name] = y.numpy()
        out['lin_%s_sig' % name] = torch.sigmoid(y).numpy()
    return out


def segment_cases():
    out = {}
    g = gen(600)
    B, T, D, S = 4, 23, 5, 6
    x = torch.randn(B, T, D, generator=g)
    seq_len = torch.tensor([23, 0, 7, 15])
    out['seg_x'], out['seg_seq_len'] = x.numpy(), seq_len.numpy()
    out['seg_select'] = utils.batched_masked_select(x, seq_len).numpy()
    lens = torch.randint(0, 6, (B, S), generator=g)
    lens[1] = 0
    lens[2, 3] = 0
    out['seg_lens'] = lens.numpy()
    out['seg_ends'] = utils.get_segment_ends(x, lens[:, :, None]).numpy()
    out['seg_split'] = utils.split_to_segments(x, lens[:, :, None]).numpy()
    xi = torch.randint(0, 50, (3, 9, 4), generator=g)
    li = torch.tensor([[2, 3, 0, 4], [0, 0, 0, 0], [9, 0, 0, 0]])
    out['seg_int_x'], out['seg_int_lens'] = xi.numpy(), li.numpy()
    out['seg_int_ends'] = utils.get_segment_ends(xi, li[:, :, None]).numpy()
    out['seg_int_split'] = utils.split_to_segments(xi, li[:, :, None]).numpy()
    out['seg_int_select'] = utils.batched_masked_select(xi, torch.tensor([9, 1, 4])).numpy()
    return out


def signature_cases():
    """Call signatures of every reference callable that `morgana_b200` mirrors, as {dotted name: [[parameter, default], ...]}
    (default "<required>" when there is none): the drop-in boundary of SURVEY.md section 8b, pinned as data."""
    import inspect
    from morgana.viz import synthesis
    names = {
        'utils': ['upsample_to_repetitions', 'sequence_mask', 'batched_masked_select', 'get_segment_ends', 'split_to_segments',
                  'both_voiced_mask', 'detach_batched_seqs'],
        'losses': ['mse', 'bce', 'ce', 'KLD_standard_normal', 'sequence_loss'],
        'data': ['normalise_mvn', 'denormalise_mvn', 'normalise_minmax', 'denormalise_minmax'],
        'viz.synthesis': ['MLPG'],
    }
    classes = {
        'utils': {'ExponentialMovingAverage': ['__init__', 'update_params']},
        'data': {'MeanVarianceNormaliser': ['__init__', 'normalise', 'denormalise', 'fetch_params', 'load_params'],
                 'MinMaxNormaliser': ['__init__', 'normalise', 'denormalise', 'fetch_params', 'load_params'],
                 'SpeakerDependentMeanVarianceNormaliser': ['__init__', 'normalise', 'denormalise', 'fetch_params', 'load_params'],
                 'SpeakerDependentMinMaxNormaliser': ['__init__', 'normalise', 'denormalise', 'fetch_params', 'load_params'],
                 'Normalisers': ['__init__'], 'ToDeviceWrapper': ['__init__', 'to_device']},
        'metrics': {cls: ['__init__', 'reset_state', 'accumulate', 'result'] for cls in
                    ['Handler', 'Print', 'History', 'TensorHistory', 'Mean', 'Variance', 'StandardDeviation', 'RMSE', 'MAE',
                     'Accuracy', 'Error', 'F0Distortion', 'LF0Distortion', 'Distortion', 'MelCepDistortion']},
    }
    classes['metrics']['Handler'] += ['add_metrics', 'add_collection', 'results_as_json_dict', 'results_as_str_dict']
    modules = {'utils': utils, 'losses': losses, 'data': data, 'metrics': metrics, 'viz.synthesis': synthesis}

    def describe(fn):
        out = []
        for p in inspect.signature(fn, follow_wrapped=False).parameters.values():
            kind = {p.VAR_POSITIONAL: '*', p.VAR_KEYWORD: '**'}.get(p.kind, '')
            default = '<required>' if p.default is p.empty else repr(p.default)
            out.append([kind + p.name, default])
        return out
    table = {}
    for mod, fns in names.items():
        for fn in fns:
            table['%s.%s' % (mod, fn)] = describe(getattr(modules[mod], fn))
    for mod, cls_map in classes.items():
        for cls, methods in cls_map.items():
            for method in methods:
                table['%s.%s.%s' % (mod, cls, method)] = describe(getattr(getattr(modules[mod], cls), method))
    return table


def main():
    import json
    if '--only' in sys.argv:      # add one group without rewriting the other archives
        name = sys.argv[sys.argv.index('--only') + 1]
        if name == 'signatures':
            with open(os.path.join(HERE, 'signatures.json'), 'w') as f:
                json.dump(signature_cases(), f, indent=1, sort_keys=True)
            print('signatures.json written')
            return
        arrays = {'voiced_mask': voiced_mask_cases, 'kld': kld_cases, 'sequence_loss': sequence_loss_cases, 'detach': detach_cases}[name]()
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **arrays)
        print('%-14s %4d arrays' % (name, len(arrays)))
        return
    with open(os.path.join(HERE, 'signatures.json'), 'w') as f:
        json.dump(signature_cases(), f, indent=1, sort_keys=True)
    print('signatures.json written')
    groups = {
        'segments': segment_cases(),
        'upsample': upsample_cases(),
        'sequence_mask': sequence_mask_cases(),
        'voiced_mask': voiced_mask_cases(),
        'kld': kld_cases(),
        'sequence_loss': sequence_loss_cases(),
        'detach': detach_cases(),
        'normalise': normaliser_cases(),
        'losses': loss_cases(),
        'metrics': metric_cases(),
        'metrics_extra': metric_extra_cases(),
        'normalise_sd': speaker_dependent_cases(),
        'ema': ema_cases(),
        'linear': linear_cases(),
    }
    for name, arrays in groups.items():
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **arrays)
        print('%-14s %4d arrays  %8.1f KB' % (name, len(arrays), os.path.getsize(path) / 1024.))
    with open(os.path.join(HERE, 'VERSIONS.txt'), 'w') as f:
        f.write('generated by tests/golden/make_golden.py from the unmodified reference at /root/reference\n')
        f.write('torch %s\nnumpy %s\nseed %d\n' % (torch.__version__, np.__version__, SEED))


if __name__ == '__main__':
    main()
