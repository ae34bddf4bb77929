"""Parity of the CUDA path (through the C ABI) against the golden fixtures and the CPU oracle.  Needs a B200."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

REL = 1e-6   # north-star tolerance for fp32 reductions; index / layout / elementwise work is bit-exact


@pytest.fixture(scope='module')
def mg():
    import morgana_b200
    return morgana_b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel_err(got, want):
    want = float(want)
    return abs(float(got) - want) / max(abs(want), 1e-30)


# ----------------------------------------------------------------------------------------------------------------------
# a1 upsample_to_repetitions
# ----------------------------------------------------------------------------------------------------------------------
UPSAMPLE_CASES = ['tiny_f32', 'odd_f32', 'lab600_f32', 'wide609_f32', 'd1_f32', 'f64', 'i64', 'f16', 'u8',
                  'emptyrow', 'allzero', 'int32dur']


@pytest.mark.parametrize('case', UPSAMPLE_CASES)
@pytest.mark.parametrize('path', ['auto', 'direct'])
def test_upsample_golden_bit_exact(mg, golden, case, path):
    g = golden('upsample')
    x, dur, want = g['ups_%s_x' % case], g['ups_%s_dur' % case], g['ups_%s_out' % case]
    got = mg.utils.upsample_to_repetitions(dev(x), dev(dur)[:, :, None], path=path)
    assert got.is_contiguous() and tuple(got.shape) == want.shape
    assert got.cpu().numpy().dtype == want.dtype
    assert np.array_equal(got.cpu().numpy(), want)


def test_upsample_2d_repeats_and_lengths(mg, golden):
    g = golden('upsample')
    x, dur = g['ups_odd_f32_x'], g['ups_odd_f32_dur']
    got, n_frames = mg.utils.upsample_to_repetitions(dev(x), dev(dur), return_lengths=True)
    assert np.array_equal(got.cpu().numpy(), g['ups_odd_f32_out'])
    assert np.array_equal(n_frames.cpu().numpy(), dur.sum(axis=1))


def test_upsample_errors(mg):
    x = torch.zeros(2, 3, 4, device='cuda')
    with pytest.raises(TypeError):
        mg.utils.upsample_to_repetitions(x, torch.ones(2, 3, 1, device='cuda'))
    with pytest.raises(ValueError, match='negative'):
        mg.utils.upsample_to_repetitions(x, torch.tensor([[1, -1, 2], [0, 0, 0]], device='cuda'))
    with pytest.raises(IndexError):
        mg.utils.upsample_to_repetitions(torch.zeros(2, 3, device='cuda'), torch.ones(2, 3, dtype=torch.long, device='cuda'))
    with pytest.raises(RuntimeError, match='no CPU path'):
        mg.utils.upsample_to_repetitions(torch.zeros(2, 3, 4), torch.ones(2, 3, 1, dtype=torch.long))


def test_upsample_non_contiguous_input(mg):
    rng = np.random.default_rng(0)
    base = rng.standard_normal((4, 9, 2 * 8)).astype(np.float32)
    dur = rng.integers(0, 5, (4, 9))
    xt = dev(base)[:, ::2, :8]          # strided in the item axis, sliced in the feature axis
    want = O.upsample_to_repetitions(base[:, ::2, :8], dur[:, ::2])
    got = mg.utils.upsample_to_repetitions(xt, dev(dur[:, ::2].copy()))
    assert np.array_equal(got.cpu().numpy(), want)


def test_upsample_max_len_hint_pads_with_zeros(mg, golden):
    g = golden('upsample')
    x, dur, want = g['ups_lab600_f32_x'], g['ups_lab600_f32_dur'], g['ups_lab600_f32_out']
    T = want.shape[1]
    got = mg.utils.upsample_to_repetitions(dev(x), dev(dur), max_len=T + 5)
    assert tuple(got.shape) == (x.shape[0], T + 5, x.shape[2])
    assert np.array_equal(got[:, :T].cpu().numpy(), want)
    assert not got[:, T:].any()


@pytest.mark.parametrize('kind', ['minmax', 'mvn'])
@pytest.mark.parametrize('path', ['bulk', 'direct'])
def test_fused_normalise_upsample_golden(mg, golden, kind, path):
    g = golden('normalise')
    p0, p1 = (g['fused_mmin'], g['fused_mmax']) if kind == 'minmax' else (g['fused_mean'], g['fused_std'])
    got = mg.utils.upsample_to_repetitions(dev(g['fused_x']), dev(g['fused_dur']), normaliser=(kind, dev(p0), dev(p1)),
                                           path=path)
    assert np.array_equal(got.cpu().numpy(), g['fused_%s_out' % kind])


@pytest.mark.parametrize('B,P,D,max_dur', [(7, 33, 600, 30), (3, 70, 187, 12), (16, 5, 4, 300), (2, 300, 64, 3),
                                             (5, 20, 609, 9), (1, 1, 8, 1), (2, 9000, 4, 2)])   # last: more items than the smem scan row
@pytest.mark.parametrize('kind', [None, 'minmax', 'mvn'])
def test_fused_vs_oracle_random(mg, B, P, D, max_dur, kind):
    rng = np.random.default_rng(B * 1000 + P + D)
    x = rng.random((B, P, D), dtype=np.float32)
    dur = rng.integers(0, max_dur + 1, (B, P))
    dur[rng.random((B, P)) < 0.15] = 0
    n_items = rng.integers(0, P + 1, B)
    dur[np.arange(P)[None, :] >= n_items[:, None]] = 0
    p0 = rng.standard_normal(D).astype(np.float32)
    p1 = (np.abs(rng.standard_normal(D)) + 0.1).astype(np.float32)
    if kind == 'minmax':
        p1 = p0 + p1
        p1[::3] = p0[::3]
    if kind is None:
        want = O.upsample_to_repetitions(x, dur)
        norm = None
    else:
        want = O.normalise_upsample(x, dur, kind, p0, p1)
        norm = (kind, dev(p0), dev(p1))
    got = mg.utils.upsample_to_repetitions(dev(x), dev(dur), normaliser=norm)
    assert np.array_equal(got.cpu().numpy(), want)
    for path in (('bulk', 'direct') if D % 4 == 0 else ('direct',)):   # the bulk engine moves 16-byte multiples
        assert torch.equal(mg.utils.upsample_to_repetitions(dev(x), dev(dur), normaliser=norm, path=path), got), path
    if kind is None:   # the segment-sum backward against the oracle
        xg = dev(x).requires_grad_()
        out = mg.utils.upsample_to_repetitions(xg, dev(dur))
        upstream = rng.standard_normal(out.shape).astype(np.float32)
        out.backward(dev(upstream))
        np.testing.assert_allclose(xg.grad.cpu().numpy(), O.upsample_backward(upstream, dur), rtol=1e-5, atol=1e-5)


def test_fused_speaker_dependent_params(mg):
    rng = np.random.default_rng(5)
    B, P, D = 4, 6, 8
    x = rng.random((B, P, D), dtype=np.float32)
    dur = rng.integers(0, 4, (B, P))
    mean = rng.standard_normal((B, D)).astype(np.float32)
    std = (rng.random((B, D)) + 0.2).astype(np.float32)
    want = O.upsample_to_repetitions(O.normalise_mvn(x, mean, std), dur)
    got = mg.utils.upsample_to_repetitions(dev(x), dev(dur), normaliser=('mvn', dev(mean), dev(std)))
    assert np.array_equal(got.cpu().numpy(), want)


def test_upsample_backward_golden_and_deterministic(mg, golden):
    g = golden('upsample')
    x = dev(g['upsbwd_x']).requires_grad_()
    out = mg.utils.upsample_to_repetitions(x, dev(g['upsbwd_dur']))
    out.backward(dev(g['upsbwd_grad_out']))
    np.testing.assert_allclose(x.grad.cpu().numpy(), g['upsbwd_grad_x'], rtol=REL, atol=1e-6)
    first = x.grad.clone()
    x.grad = None
    mg.utils.upsample_to_repetitions(x, dev(g['upsbwd_dur'])).backward(dev(g['upsbwd_grad_out']))
    assert torch.equal(first, x.grad)


def test_upsample_backward_through_normaliser(mg):
    rng = np.random.default_rng(11)
    B, P, D = 3, 10, 600
    x = rng.random((B, P, D), dtype=np.float32)
    dur = rng.integers(0, 6, (B, P))
    mean = rng.standard_normal(D).astype(np.float32)
    std = (rng.random(D) + 0.2).astype(np.float32)
    xt = dev(x).requires_grad_()
    out = mg.utils.upsample_to_repetitions(xt, dev(dur), normaliser=('mvn', dev(mean), dev(std)))
    grad_out = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    out.backward(dev(grad_out))
    want = O.upsample_backward(grad_out, dur).astype(np.float64) / (std + np.float32(1e-8)).astype(np.float64)
    np.testing.assert_allclose(xt.grad.cpu().numpy(), want, rtol=2e-6, atol=1e-6)


# ----------------------------------------------------------------------------------------------------------------------
# a3 / a4 normalisers
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('case', ['btd', 'td', 'btd187', 'sd'])
def test_normalisers_golden_bit_exact(mg, golden, case):
    g = golden('normalise')
    x = dev(g['norm_%s_x' % case])
    mean, std = dev(g['norm_%s_mean' % case]), dev(g['norm_%s_std' % case])
    mmin, mmax = dev(g['norm_%s_mmin' % case]), dev(g['norm_%s_mmax' % case])
    D = mg.data
    for fn, args, key in [(D.normalise_mvn, (mean, std), 'mvn'), (D.denormalise_mvn, (mean, std), 'demvn'),
                          (D.normalise_minmax, (mmin, mmax), 'minmax'), (D.denormalise_minmax, (mmin, mmax), 'deminmax')]:
        got = fn(x, *args).cpu().numpy()
        assert np.array_equal(got, g['norm_%s_%s' % (case, key)]), key


def test_normaliser_classes_and_autograd(mg):
    rng = np.random.default_rng(2)
    D = 187
    norm = mg.data.MeanVarianceNormaliser('mcep', use_deltas=True).set_params(
        {'mean': rng.standard_normal(D), 'std_dev': rng.random(D) + 0.3},
        {'mean': rng.standard_normal(D), 'std_dev': rng.random(D) + 0.3})
    x = rng.standard_normal((3, 50, D)).astype(np.float32)
    for deltas in (False, True):
        p = norm.fetch_params(np.ndarray, deltas=deltas)
        want = O.denormalise_mvn(x, p['mean'], p['std_dev'])
        assert np.array_equal(norm.denormalise(dev(x), deltas=deltas).cpu().numpy(), want)
        assert np.allclose(norm.denormalise(x, deltas=deltas), want)   # NumPy inputs stay in NumPy
    xt = dev(x).requires_grad_()
    norm.normalise(xt).sum().backward()
    want_grad = np.broadcast_to(1. / (norm.params['std_dev'] + np.float32(1e-8)), x.shape)
    np.testing.assert_allclose(xt.grad.cpu().numpy(), want_grad, rtol=1e-6)


def test_normalise_large_odd_shape(mg):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((37, 211, 187)).astype(np.float32)
    mmin = rng.standard_normal(187).astype(np.float32)
    mmax = (mmin + rng.random(187) + 0.01).astype(np.float32)
    got = mg.data.normalise_minmax(dev(x), dev(mmin), dev(mmax)).cpu().numpy()
    assert np.array_equal(got, O.normalise_minmax(x, mmin, mmax))
    got = mg.data.denormalise_minmax(dev(x)[:, :, :], dev(mmin), dev(mmax)).cpu().numpy()
    assert np.array_equal(got, O.denormalise_minmax(x, mmin, mmax))


# ----------------------------------------------------------------------------------------------------------------------
# a6 / a7 losses
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('case', ['small', 'wide', 'd1'])
@pytest.mark.parametrize('masked', [True, False])
def test_losses_golden(mg, golden, case, masked):
    g = golden('losses')
    tag = 'masked' if masked else 'full'
    seq_len = dev(g['loss_%s_seq_len' % case]) if masked else None
    for kind, fn, a, b in [('mse', mg.losses.mse, 'pred', 'tgt'), ('bce', mg.losses.bce, 'prob', 'label')]:
        p = dev(g['loss_%s_%s' % (case, a)]).requires_grad_()
        y = dev(g['loss_%s_%s' % (case, b)])
        value = fn(p, y, seq_len)
        assert value.dim() == 0 and value.dtype == torch.float32
        want = g['loss_%s_%s_%s' % (case, kind, tag)]
        assert rel_err(value.item(), want) <= REL, (kind, value.item(), float(want))
        (value * 0.25).backward()     # a non-trivial upstream gradient, as in `loss / 4.` (models/RNN_SPSS.py:139)
        np.testing.assert_allclose(p.grad.cpu().numpy(), 0.25 * g['loss_%s_%s_%s_grad' % (case, kind, tag)],
                                   rtol=2e-6, atol=1e-10)


def test_loss_semantics(mg):
    pred = torch.randn(2, 4, 3, device='cuda')
    tgt = torch.randn(2, 4, 3, device='cuda')
    assert torch.isnan(mg.losses.mse(pred, tgt, torch.tensor([0, 3], device='cuda')))   # 0/0, SURVEY.md Q6
    with pytest.raises(RuntimeError):
        mg.losses.mse(pred, tgt, torch.tensor([[1], [3]], device='cuda'))
    with pytest.raises(RuntimeError, match='no CPU path'):
        mg.losses.mse(pred.cpu(), tgt.cpu())
    # seq_len longer than T behaves like T (mask is arange(T) < seq_len)
    a = mg.losses.mse(pred, tgt, torch.tensor([9, 4], device='cuda')).item()
    b = mg.losses.mse(pred, tgt).item()
    assert a == b


@pytest.mark.parametrize('B,T,D', [(8, 301, 187), (5, 1000, 3), (3, 77, 180), (64, 50, 1)])
@pytest.mark.parametrize('kind', ['mse', 'l1', 'bce'])
def test_losses_vs_oracle_random_and_strided(mg, B, T, D, kind):
    rng = np.random.default_rng(B + T + D)
    seq_len = rng.integers(1, T + 1, B)
    wide = rng.random((B, T, D + 7), dtype=np.float32) * 0.98 + 0.01
    y = (rng.random((B, T, D)) < 0.5).astype(np.float32) if kind == 'bce' else rng.standard_normal((B, T, D)).astype(np.float32)
    fn = getattr(mg.losses, kind)
    wide_t = dev(wide).requires_grad_()
    value = fn(wide_t[:, :, 3:3 + D], dev(y), dev(seq_len))     # a column slice: strided rows, no copy
    want = O.masked_loss(wide[:, :, 3:3 + D], y, seq_len, kind)
    assert rel_err(value.item(), want) <= REL
    value.backward()
    want_grad = np.zeros_like(wide)
    want_grad[:, :, 3:3 + D] = O.masked_loss_grad(wide[:, :, 3:3 + D], y, seq_len, kind)
    np.testing.assert_allclose(wide_t.grad.cpu().numpy(), want_grad, rtol=REL, atol=1e-12)   # measured: mse 1.7e-7, l1 0, bce 2.4e-7 (scripts/measure_grad_errors.py)
    # contiguous operands take the vector path; same answer to the bit across runs
    c = dev(np.ascontiguousarray(wide[:, :, 3:3 + D]))
    v1, v2 = fn(c, dev(y), dev(seq_len)), fn(c, dev(y), dev(seq_len))
    assert rel_err(v1.item(), want) <= REL and v1.item() == v2.item()


def test_speaker_dependent_normalisers_golden(mg, golden):
    """SpeakerDependent{MeanVariance,MinMax}Normaliser against the reference's outputs: (B, D) parameters gathered per
    batch item, bit-exact; also fused into the expansion (per-utterance parameters in K2)."""
    g = golden('normalise_sd')
    speakers = [str(s) for s in g['sdn_speakers']]
    batch_ids = [str(s) for s in g['sdn_batch_ids']]
    D = mg.data
    mvn = D.SpeakerDependentMeanVarianceNormaliser('lf0', 'speakers.scp', use_deltas=True).set_params(
        {s: {k: g['sdn_mvn_%s_%s' % (s, k)] for k in ('mean', 'std_dev')} for s in speakers},
        {s: {k: g['sdn_mvn_deltas_%s_%s' % (s, k)] for k in ('mean', 'std_dev')} for s in speakers})
    mm = D.SpeakerDependentMinMaxNormaliser('lab', 'speakers.scp').set_params(
        {s: {k: g['sdn_minmax_%s_%s' % (s, k)] for k in ('mmin', 'mmax')} for s in speakers})
    x = dev(g['sdn_x'])
    for got, key in [(mvn.normalise(x, batch_ids), 'sdn_mvn_norm'), (mvn.denormalise(x, batch_ids), 'sdn_mvn_denorm'),
                     (mvn.normalise(x, batch_ids, deltas=True), 'sdn_mvn_norm_deltas'),
                     (mm.normalise(x, batch_ids), 'sdn_minmax_norm'), (mm.denormalise(x, batch_ids), 'sdn_minmax_denorm'),
                     (mvn.normalise(x[1], 'spk_c'), 'sdn_mvn_norm_single')]:
        assert np.array_equal(got.cpu().numpy(), g[key]), key
    # fused: per-utterance parameters inside the expansion == normalise, then expand
    dur = dev(np.array([[2, 0, 1, 3, 1, 0, 2, 1, 1]] * 4) + np.arange(4)[:, None] % 2)
    fused = mg.utils.upsample_to_repetitions(x, dur, normaliser=mm.fused_params(batch_ids))
    want = mg.utils.upsample_to_repetitions(mm.normalise(x, batch_ids), dur)
    assert torch.equal(fused, want)


def test_to_device_wrapper_prefetches_one_batch_ahead(mg):
    """The feeder (data.py:648-663 counterpart): pinned staging + side-stream copies one batch ahead, same values."""
    g = torch.Generator().manual_seed(5)
    batches = [{'lab': torch.randn(3, 50, 600, generator=g), 'dur': torch.randint(0, 9, (3, 50, 1), generator=g),
                'name': ['u%d' % i], 'n_frames': [torch.tensor([7, 8, 9])]} for i in range(5)]
    seen = []
    for features in mg.data.ToDeviceWrapper(batches, 'cuda'):
        assert features['lab'].is_cuda and features['n_frames'][0].is_cuda and features['name'][0].startswith('u')
        seen.append(mg.utils.upsample_to_repetitions(features['lab'], features['dur']).sum(dim=(1, 2)).cpu())
    assert len(seen) == 5
    for got, b in zip(seen, batches):
        want = O.upsample_to_repetitions(b['lab'].numpy(), b['dur'].numpy()[:, :, 0]).astype(np.float64).sum(axis=(1, 2))
        np.testing.assert_allclose(got.numpy(), want, rtol=1e-4)


@pytest.mark.parametrize('case', ['ce_small', 'ce_wide'])
@pytest.mark.parametrize('masked', [True, False])
def test_cross_entropy_golden(mg, golden, case, masked):
    """losses.ce (losses.py:59-61) forward and backward against the reference's outputs."""
    g = golden('losses')
    logits = dev(g['loss_%s_logits' % case]).requires_grad_()
    classes = dev(g['loss_%s_classes' % case])
    seq_len = dev(g['loss_%s_seq_len' % case]) if masked else None
    tag = 'loss_%s_%s' % (case, 'masked' if masked else 'full')
    value = mg.losses.ce(logits, classes, seq_len)
    assert rel_err(value.item(), g[tag]) <= REL
    (value * 3.).backward()
    np.testing.assert_allclose(logits.grad.cpu().numpy(), 3. * g[tag + '_grad'], rtol=3e-6, atol=6e-8)   # softmax - onehot cancels near 1


def test_cross_entropy_vs_oracle_strided_and_ragged(mg):
    rng = np.random.default_rng(41)
    B, T, C = 7, 301, 40
    wide = (2. * rng.standard_normal((B, T, C + 6))).astype(np.float32)
    classes = rng.integers(0, C, (B, T))
    seq_len = rng.integers(1, T + 1, B)
    logits = dev(wide)[:, :, 3:3 + C].requires_grad_()            # a column slice: strided rows
    value = mg.losses.ce(logits, dev(classes), dev(seq_len))
    want, want_grad = O.cross_entropy_loss(wide[:, :, 3:3 + C], classes, seq_len)
    assert rel_err(value.item(), want) <= REL
    grad, = torch.autograd.grad(value, logits)
    np.testing.assert_allclose(grad.cpu().numpy(), want_grad, rtol=3e-6, atol=2e-8)
    with pytest.raises(RuntimeError):
        mg.losses.ce(logits, dev(classes).float(), dev(seq_len))


def test_reductions_reuse_their_workspace_across_batch_sizes(mg):
    """The per-stream workspace is zeroed once and reused: a small batch followed by a larger one (the last batch of an
    epoch, then the next epoch; validation after training) must not see the small batch's partial sums as tickets."""
    from morgana_b200.fused import AcousticObjective
    rng = np.random.default_rng(77)
    T = 90
    results = []
    for B in (6, 3, 48, 5, 200, 48):
        n = rng.integers(1, T + 1, B)
        tgt = rng.standard_normal((B, T, 187)).astype(np.float32)
        pred = (tgt + 0.1 * rng.standard_normal((B, T, 187))).astype(np.float32)
        tgt[:, :, 3] = rng.random((B, T)) < 0.6
        pred[:, :, 3] = 1. / (1. + np.exp(-rng.standard_normal((B, T))))
        got = mg.losses.mse(dev(pred[..., 4:184]), dev(tgt[..., 4:184]), dev(n)).item()
        want = O.masked_loss(pred[..., 4:184], tgt[..., 4:184], n)
        assert rel_err(got, want) <= REL, B
        total, _ = AcousticObjective()(dev(pred), dev(tgt), dev(n))
        want_total = (O.masked_loss(pred[..., 0:3], tgt[..., 0:3], n) + want + O.masked_loss(pred[..., 184:187], tgt[..., 184:187], n)
                      + O.masked_loss(pred[..., 3:4], tgt[..., 3:4], n, 'bce')) / 4.
        assert rel_err(total.item(), want_total) <= REL, B
        results.append(total.item())
    assert all(np.isfinite(results)) and min(results) > 0.


@pytest.mark.parametrize('D,max_items', [(600, 40), (187, 25), (4, 9000), (8, 1)])
@pytest.mark.parametrize('kind', [None, 'minmax'])
def test_packed_items_equal_padded_items(mg, D, max_items, kind):
    """"Next" row 3: K1 / K2 on the packed wire format (items utterance after utterance, no padding) give the tensor the
    padded path gives, bit for bit, with and without the host-side hints; per-utterance parameters included."""
    rng = np.random.default_rng(D + max_items)
    B = 7
    n_items = rng.integers(0, max_items + 1, B)
    n_items[2] = 0                                            # an utterance without items
    n_items[4] = max_items
    P = int(n_items.max())
    x = rng.random((B, P, D), dtype=np.float32)
    dur = rng.integers(0, 5, (B, P))
    valid = np.arange(P)[None, :] < n_items[:, None]
    dur[~valid] = 0
    x[~valid] = 0
    lo = rng.standard_normal((B, D)).astype(np.float32)
    hi = (lo + np.abs(rng.standard_normal((B, D))) + 0.1).astype(np.float32)
    norm = None if kind is None else (kind, dev(lo), dev(hi))
    padded, n_frames = mg.utils.upsample_to_repetitions(dev(x), dev(dur), normaliser=norm, return_lengths=True)
    packed_x, packed_dur = dev(x[valid]), dev(dur[valid])
    got, got_frames = mg.utils.upsample_packed_to_repetitions(packed_x, packed_dur, dev(n_items), normaliser=norm, return_lengths=True)
    assert torch.equal(got, padded) and torch.equal(got_frames, n_frames)
    hinted = mg.utils.upsample_packed_to_repetitions(packed_x, packed_dur.to(torch.int32), dev(n_items), normaliser=norm,
                                                     max_len=padded.shape[1], max_items=P)
    assert torch.equal(hinted, padded)
    want = O.upsample_to_repetitions(x, dur) if kind is None else O.normalise_upsample(x, dur, kind, lo, hi)
    assert np.array_equal(got.cpu().numpy(), want)
    with pytest.raises(ValueError):
        mg.utils.upsample_packed_to_repetitions(packed_x, packed_dur, dev(n_items + 1))      # counts do not match the items


# ----------------------------------------------------------------------------------------------------------------------
# a8 - a12 metrics
# ----------------------------------------------------------------------------------------------------------------------
def _metric_batches(g):
    keys = ['seq_len', 'tgt', 'pred', 'lf0_t', 'lf0_p', 'voiced', 'bits_t', 'bits_p']
    return [{k: dev(g['met_b%d_%s' % (i, k)]) for k in keys} for i in range(2)]


def _metric_specs(M):
    return {
        'mean': (M.Mean, lambda b: (b['tgt'],)),
        'rmse': (M.RMSE, lambda b: (b['tgt'], b['pred'])),
        'mae': (M.MAE, lambda b: (b['tgt'], b['pred'])),
        'melcep': (M.MelCepDistortion, lambda b: (b['tgt'], b['pred'])),
        'distortion': (M.Distortion, lambda b: (b['tgt'], b['pred'])),
        'f0': (M.F0Distortion, lambda b: (b['lf0_t'].exp(), b['lf0_p'].exp(), b['voiced'])),
        'lf0': (M.LF0Distortion, lambda b: (b['lf0_t'], b['lf0_p'], b['voiced'])),
        'lf0_floatmask': (M.LF0Distortion, lambda b: (b['lf0_t'], b['lf0_p'], b['voiced'].float())),
        'error': (M.Error, lambda b: (b['bits_t'], b['bits_p'])),
        'accuracy': (M.Accuracy, lambda b: (b['bits_t'], b['bits_p'])),
        'error_u8': (M.Error, lambda b: (b['bits_t'].to(torch.uint8), b['bits_p'].to(torch.uint8))),
        'vuvacc': (M.Mean, lambda b: ((b['bits_t'] == b['bits_p']).type(torch.float),)),
    }


@pytest.mark.parametrize('name', ['mean', 'rmse', 'mae', 'melcep', 'distortion', 'f0', 'lf0', 'lf0_floatmask', 'error',
                                  'accuracy', 'error_u8', 'vuvacc'])
@pytest.mark.parametrize('masked', [True, False])
def test_metrics_golden(mg, golden, name, masked):
    g = golden('metrics')
    cls, args_of = _metric_specs(mg.metrics)[name]
    metric = cls()
    metric.reset_state()
    voiced_before = [b['voiced'].clone() for b in _metric_batches(g)]
    batches = _metric_batches(g)
    for b in batches:
        metric.accumulate(*args_of(b), seq_len=b['seq_len'] if masked else None)
    key = 'met_%s_%s' % (name, 'masked' if masked else 'full')
    assert float(metric.count) == float(g[key + '_count'])
    if name in ('error', 'accuracy', 'error_u8'):
        assert int(metric.sum) == int(g[key + '_sum'])            # integer work: exact
    else:
        assert rel_err(metric.sum, g[key + '_sum']) <= REL
    assert rel_err(metric.result(), g[key + '_result']) <= REL
    for b, before in zip(batches, voiced_before):                 # inputs are never mutated
        assert torch.equal(b['voiced'], before)


@pytest.mark.parametrize('masked', [True, False])
def test_variance_and_tensor_history_golden(mg, golden, masked):
    """Variance / StandardDeviation (metrics.py:400-471) and TensorHistory (:263-356) against the reference's outputs."""
    g = golden('metrics_extra')
    tag = 'masked' if masked else 'full'
    M = mg.metrics
    var, std = M.Variance(), M.StandardDeviation()
    hist, hist_short = M.TensorHistory(5), M.TensorHistory(5, max_len=7)
    for i in range(2):
        x = dev(g['mx_b%d_x' % i])
        before = x.clone()
        seq_len = dev(g['mx_b%d_seq_len' % i]) if masked else None
        for metric in (var, std, hist, hist_short):
            metric.accumulate(x, seq_len=seq_len)
        assert torch.equal(x, before)                      # the reference zeroes the padding in place; we do not
    assert float(var.count) == float(g['mx_var_%s_count' % tag])
    assert rel_err(var.sum, g['mx_var_%s_sum' % tag]) <= REL
    assert rel_err(var.sum_square, g['mx_var_%s_sum_square' % tag]) <= REL
    # the reference forms sum_square - sum^2 / count in fp32 (cancellation ~5x here); ours forms it in fp64
    assert rel_err(var.result(), g['mx_var_%s_result' % tag]) <= 1e-5
    want_std = float(g['mx_std_%s_result' % tag])
    if np.isnan(want_std):                                 # frames-vs-elements count quirk (Q2), pinned as is
        assert np.isnan(float(std.result()))
    else:
        assert rel_err(std.result(), want_std) <= 1e-5
    assert np.array_equal(hist.result().cpu().numpy(), g['mx_hist_%s' % tag])            # row packing: bit-exact
    assert np.array_equal(hist_short.result().cpu().numpy(), g['mx_hist_short_%s' % tag])
    assert str(hist).startswith('N(')


def test_metric_handler_golden(mg, golden):
    """The container as a model drives it: `metrics.accumulate(mode, name=(tensors..., seq_len))`
    (models/RNN_SPSS.py:124-129) plus the 0-dim batch loss of experiment_builder.py:484."""
    g = golden('metrics_extra')
    M = mg.metrics
    handler = M.Handler(loss=M.Mean())
    handler.add_metrics('all', err=M.RMSE(), mae=M.MAE())
    handler.add_metrics('valid', spread=M.Variance())
    assert sorted(handler['train']) == list(g['mx_handler_train_names'])
    assert sorted(handler['valid']) == list(g['mx_handler_valid_names'])
    with pytest.raises(ValueError, match='No collection found'):
        handler.accumulate('', loss=torch.zeros((), device='cuda'))
    for mode in ('train', 'valid'):
        handler.reset_state(mode)
        assert handler.results_as_json_dict(mode) == {}                      # everything hidden until accumulated
        for i in range(2):
            x, y, seq_len = (dev(g['mx_b%d_%s' % (i, k)]) for k in ('x', 'y', 'seq_len'))
            kwargs = dict(err=(x, y, seq_len), mae=(x, y, {'seq_len': seq_len}), loss=torch.mean(x))
            if mode == 'valid':
                kwargs['spread'] = (x, seq_len)
            handler.accumulate(mode, **kwargs)
        results = handler.results_as_json_dict(mode)
        assert sorted(results) == sorted(handler[mode])
        for name, value in results.items():
            tol = 1e-5 if name == 'spread' else REL
            assert rel_err(value, g['mx_handler_%s_%s' % (mode, name)]) <= tol, (mode, name)
        assert set(handler.results_as_str_dict(mode)) == set(results)
    assert ' | ' in str(handler)


def test_metrics_accept_any_shape_without_seq_len(mg):
    """Without seq_len the reference reduces any shape (count += numel), e.g. the 0-dim batch loss."""
    rng = np.random.default_rng(3)
    M = mg.metrics
    loss = M.Mean()
    values = rng.standard_normal(5).astype(np.float32)
    for v in values:
        loss.accumulate(dev(np.asarray(v)))
    assert float(loss.count) == 5. and rel_err(loss.result(), values.astype(np.float64).mean()) <= REL
    a, b = rng.standard_normal((7, 11)).astype(np.float32), rng.standard_normal((7, 11)).astype(np.float32)
    rmse, dist = M.RMSE(), M.Distortion()
    rmse.reset_state()
    rmse.accumulate(dev(a), dev(b))
    assert float(rmse.count) == 77. and rel_err(rmse.sum, ((a.astype(np.float64) - b) ** 2).sum()) <= REL
    dist.accumulate(dev(a), dev(b))
    assert float(dist.count) == 7.
    assert rel_err(dist.sum, np.sqrt(((a.astype(np.float64) - b) ** 2).sum(-1)).sum()) <= REL
    with pytest.raises(ValueError):
        rmse.accumulate(dev(a), dev(b), seq_len=dev(np.array([3] * 7)))


def test_metrics_full_size_vs_oracle(mg):
    """Config 3 shape (reduced batch): 187-dim targets, static-column metrics of models/RNN_SPSS.py:124-129."""
    rng = np.random.default_rng(7)
    B, T, D = 48, 1200, 187
    seq_len = rng.integers(300, T + 1, B)
    tgt = rng.standard_normal((B, T, D)).astype(np.float32)
    pred = (tgt + 0.1 * rng.standard_normal((B, T, D))).astype(np.float32)
    tgt[:, :, 0] = 5 + 0.3 * tgt[:, :, 0]
    pred[:, :, 0] = tgt[:, :, 0] + 0.05 * pred[:, :, 0]
    voiced = rng.random((B, T, 1)) < 0.6
    tgt_t, pred_t, seq_t, voiced_t = dev(tgt), dev(pred), dev(seq_len), dev(voiced)
    M = mg.metrics
    checks = [
        (M.LF0Distortion(), (tgt_t[:, :, 0:1], pred_t[:, :, 0:1], voiced_t), O.lf0_acc(tgt[:, :, 0:1], pred[:, :, 0:1], voiced, seq_len)),
        (M.MelCepDistortion(), (tgt_t[:, :, 4:64], pred_t[:, :, 4:64]), O.melcep_acc(tgt[:, :, 4:64], pred[:, :, 4:64], seq_len)),
        (M.Distortion(), (tgt_t[:, :, 184:185], pred_t[:, :, 184:185]), O.distortion_acc(tgt[:, :, 184:185], pred[:, :, 184:185], seq_len)),
        (M.RMSE(), (tgt_t, pred_t), O.rmse_acc(tgt, pred, seq_len)),
    ]
    for metric, args, (want_sum, want_count) in checks:
        metric.reset_state()
        metric.accumulate(*args, seq_len=seq_t)
        assert float(metric.count) == want_count
        assert rel_err(metric.sum, want_sum) <= REL, type(metric).__name__


# ----------------------------------------------------------------------------------------------------------------------
# a13 EMA
# ----------------------------------------------------------------------------------------------------------------------
def test_ema_golden_bit_exact(mg, golden):
    g = golden('ema')
    n, decay = int(g['ema_n']), float(g['ema_decay'])
    model = torch.nn.ParameterList([torch.nn.Parameter(dev(g['ema_param0_%d' % i])) for i in range(n)])
    ema_model = torch.nn.ParameterList([torch.nn.Parameter(dev(g['ema_shadow0_%d' % i])) for i in range(n)])
    ema = mg.utils.ExponentialMovingAverage(ema_model, decay)
    for step in range(3):
        with torch.no_grad():
            for i, p in enumerate(model):
                p.copy_(dev(g['ema_param%d_%d' % (step, i)]))
        ema.update_params(model)
        for i, p in enumerate(ema_model):
            assert np.array_equal(p.detach().cpu().numpy(), g['ema_shadow%d_%d' % (step + 1, i)])
    with pytest.raises(AssertionError):
        ema.update_params(ema_model)


def test_ema_many_tensors_bit_exact(mg):
    """More than 64 tensors (several launches), odd sizes, a misaligned view; against the NumPy restatement."""
    rng = np.random.default_rng(9)
    sizes = [1, 3, 4, 5, 8191, 8192, 8193, 100003] + [int(s) for s in rng.integers(1, 3000, 70)]
    shadow = [rng.standard_normal(s).astype(np.float32) for s in sizes]
    param = [rng.standard_normal(s).astype(np.float32) for s in sizes]
    flat = torch.zeros(sum(sizes) + 1, device='cuda')
    views, off = [], 1                                   # offset by one float: most views are not 16-byte aligned
    for s, arr in zip(sizes, shadow):
        views.append(flat[off:off + s])
        views[-1].copy_(dev(arr))
        off += s
    pairs = [(v, dev(p)) for v, p in zip(views, param)]
    mg.ops.ema_update(pairs, 1.0 - 0.9999)
    for v, s, p in zip(views, shadow, param):
        assert np.array_equal(v.cpu().numpy(), O.ema_update(s.copy(), p, 0.9999))


# ----------------------------------------------------------------------------------------------------------------------
# K4b: the whole-row fused objective (additive API) equals the composition of the drop-in ops and the oracle
# ----------------------------------------------------------------------------------------------------------------------
def _objective_oracle(p, t, voiced_t, n, mcep_static=60, bap_static=1):
    total = (O.masked_loss(p[..., 0:3], t[..., 0:3], n) + O.masked_loss(p[..., 4:184], t[..., 4:184], n) +
             O.masked_loss(p[..., 184:187], t[..., 184:187], n) + O.masked_loss(p[..., 3:4], t[..., 3:4], n, 'bce')) / 4.
    grad = np.zeros_like(p)
    for sl, kind in [(slice(0, 3), 'mse'), (slice(4, 184), 'mse'), (slice(184, 187), 'mse'), (slice(3, 4), 'bce')]:
        grad[..., sl] = 0.25 * O.masked_loss_grad(p[..., sl], t[..., sl], n, kind)
    voiced_p = p[..., 3:4] > 0.5
    metrics = {'LF0_RMSE_Hz': O.lf0_acc(t[..., 0:1], p[..., 0:1], voiced_p, n),
               'VUV_accuracy': O.mean_acc((voiced_t == voiced_p).astype(np.float32), n),
               'MCEP_distortion': O.melcep_acc(t[..., 4:4 + mcep_static], p[..., 4:4 + mcep_static], n),
               'BAP_distortion': O.distortion_acc(t[..., 184:184 + bap_static], p[..., 184:184 + bap_static], n)}
    return total, grad, metrics


@pytest.mark.parametrize('B,min_p,max_p', [(6, 5, 12), (33, 20, 40)])
@pytest.mark.parametrize('bap_static', [1, 3])
def test_fused_objective_vs_oracle(mg, B, min_p, max_p, bap_static):
    from morgana_b200 import workloads
    from morgana_b200.fused import AcousticObjective
    ling = workloads.linguistic_batch(batch_size=B, min_phones=min_p, max_phones=max_p, max_dur=20, seed=B)
    ac = workloads.acoustic_batch(ling['n_frames'], seed=B)
    p, t, n = ac['pred'].numpy(), ac['target'].numpy(), ling['n_frames'].numpy()
    want_total, want_grad, want_metrics = _objective_oracle(p, t, ac['voiced'].numpy(), n, bap_static=bap_static)
    pred, target, n_frames = ac['pred'].cuda(), ac['target'].cuda(), ling['n_frames'].cuda()
    for which in ('__call__', 'call_with_terms'):
        objective = AcousticObjective(bap_static=bap_static)
        for repeat in range(2):          # metric state is a running sum; the loss is per batch
            total, grad = getattr(objective, which)(pred, target, n_frames)
        assert rel_err(total.item(), want_total) <= REL, which
        np.testing.assert_allclose(grad.cpu().numpy(), want_grad, rtol=REL, atol=1e-12)
        for name, (s, c) in want_metrics.items():
            got = objective.metrics[name]
            assert float(got.count) == 2 * c, (which, name)
            assert rel_err(got.sum, 2 * s) <= REL, (which, name)
        loss_only, no_grad = getattr(objective, which)(pred, target, n_frames, want_grad=False)
        assert no_grad is None and loss_only.item() == total.item()       # bit-reproducible run to run


@pytest.mark.parametrize('lengths,T', [
    ([61], 61),                       # 8 stages, 2 CTAs: the tensor's short last stage lies inside a head start
    ([5, 3, 5, 1, 4, 5, 2], 5),       # utterances shorter than a stage: every stage spans several of them
    ([40, 0, 17, 64, 64, 1], 64),     # an empty utterance (loss nan, everything else defined), full-length ones
    ([9] * 300, 13),                  # many short utterances, 4 padding rows each
    ([700, 3, 350], 701),             # T odd: no stage boundary ever coincides with an utterance boundary
    ([1200, 1187, 33, 640], 1200),    # a full-length row first
    ([16] * 8, 16),                   # no padding anywhere, every stage full
    ([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12], 40),   # mostly padding: padding-only stages in the head starts
])
def test_fused_objective_edge_shapes(mg, lengths, T):
    """The persistent objective kernel on shapes that exercise its stage classification (full / mixed / padding-only / short last
    stage), the fixed head starts and the cost partition behind them, against the fp64 oracle."""
    from morgana_b200 import workloads
    from morgana_b200.fused import AcousticObjective
    n = np.asarray(lengths, dtype=np.int64)
    ac = workloads.acoustic_batch(torch.from_numpy(n), max_len=T, seed=len(lengths) + T)
    p, t = ac['pred'].numpy(), ac['target'].numpy()
    has_empty = bool((n == 0).any())
    with np.errstate(invalid='ignore', divide='ignore'):
        want_total, want_grad, want_metrics = _objective_oracle(p, t, ac['voiced'].numpy(), n)
    pred, target, n_frames = ac['pred'].cuda(), ac['target'].cuda(), torch.from_numpy(n).cuda()
    objective = AcousticObjective()
    total, grad = objective(pred, target, n_frames)
    if has_empty:
        assert np.isnan(total.item())                          # 0 / 0 for the empty utterance, as the reference (losses.py:39)
        keep = n > 0
        np.testing.assert_allclose(grad.cpu().numpy()[keep], want_grad[keep], rtol=REL, atol=1e-12)
        assert (grad.cpu().numpy()[~keep] == 0).all()          # an empty utterance is padding only
    else:
        assert rel_err(total.item(), want_total) <= REL
        np.testing.assert_allclose(grad.cpu().numpy(), want_grad, rtol=REL, atol=1e-12)
    for name, (s_want, c_want) in want_metrics.items():
        got = objective.metrics[name]
        assert float(got.count) == c_want, name
        assert rel_err(got.sum, s_want) <= REL or abs(float(got.sum) - s_want) <= 1e-12, name
    again, _ = objective(pred, target, n_frames, want_grad=False)
    if not has_empty:
        assert again.item() == total.item()                    # forward-only form: same bits


def test_fused_objective_matches_drop_in_composition(mg):
    """What models/RNN_SPSS.py:120-139 computes through the drop-in ops == the one-launch objective."""
    from morgana_b200 import workloads
    from morgana_b200.fused import AcousticObjective
    ling = workloads.linguistic_batch(batch_size=9, min_phones=8, max_phones=15, max_dur=15, seed=21)
    ac = workloads.acoustic_batch(ling['n_frames'], seed=21)
    pred, target, n_frames = ac['pred'].cuda().requires_grad_(), ac['target'].cuda(), ling['n_frames'].cuda()
    L, M = mg.losses, mg.metrics
    loss = (L.mse(pred[..., 0:3], target[..., 0:3], n_frames) + L.mse(pred[..., 4:184], target[..., 4:184], n_frames) +
            L.mse(pred[..., 184:187], target[..., 184:187], n_frames) + L.bce(pred[..., 3:4], target[..., 3:4], n_frames)) / 4.
    loss.backward()
    vuv = pred.detach()[..., 3:4] > 0.5
    lf0, acc, mcd, bap = M.LF0Distortion(), M.Mean(), M.MelCepDistortion(), M.Distortion()
    lf0.accumulate(target[..., 0:1], pred.detach()[..., 0:1], vuv, seq_len=n_frames)
    acc.accumulate((target[..., 3:4] == vuv).type(torch.float), seq_len=n_frames)
    mcd.accumulate(target[..., 4:64], pred.detach()[..., 4:64], seq_len=n_frames)
    bap.accumulate(target[..., 184:185], pred.detach()[..., 184:185], seq_len=n_frames)
    objective = AcousticObjective()
    total, grad = objective(pred.detach(), target, n_frames)
    assert rel_err(total.item(), loss.item()) <= REL
    np.testing.assert_allclose(grad.cpu().numpy(), pred.grad.cpu().numpy(), rtol=REL, atol=1e-12)
    for got, want in zip(objective.metrics.values(), (lf0, acc, mcd, bap)):
        assert float(got.count) == float(want.count)
        assert rel_err(got.sum, float(want.sum)) <= REL
        assert rel_err(got.result(), float(want.result())) <= REL


def test_sequence_loss_wrapper_golden(mg, golden):
    """losses.sequence_loss around a caller-defined per-element loss (smooth-L1, which the reference does not ship) against
    the reference's wrapper (losses.py:9-47): the loss_fn is the caller's torch code, masking + normalisation + mean are ours."""
    g = golden('sequence_loss')

    @mg.losses.sequence_loss
    def huber(predictions, targets):
        return torch.nn.functional.smooth_l1_loss(predictions, targets, reduction='none')

    assert huber.__name__ == 'huber'
    p, t, n = dev(g['seqloss_p']).requires_grad_(), dev(g['seqloss_t']), dev(g['seqloss_n'])
    loss = huber(p, t, seq_len=n)
    assert abs(loss.item() - float(g['seqloss_masked'])) <= 1e-6 * abs(float(g['seqloss_masked']))
    (3. * loss).backward()
    np.testing.assert_allclose(p.grad.cpu().numpy(), g['seqloss_masked_grad'], rtol=REL, atol=1e-10)
    p.grad = None
    loss = huber(p, t)
    assert abs(loss.item() - float(g['seqloss_full'])) <= 1e-6 * abs(float(g['seqloss_full']))
    loss.backward()
    np.testing.assert_allclose(p.grad.cpu().numpy(), g['seqloss_full_grad'], rtol=REL, atol=1e-10)


@pytest.mark.parametrize('tag', ['2d', '3d', 'row'])
def test_kld_standard_normal_golden(mg, golden, tag):
    """losses.KLD_standard_normal against outputs and autograd gradients of the reference (losses.py:64-67): 1e-6 relative on
    the value, 3e-6 on the gradients (an upstream factor of 0.25 included)."""
    g = golden('kld')
    mean, lv = dev(g['kld_%s_mean' % tag]).requires_grad_(), dev(g['kld_%s_lv' % tag]).requires_grad_()
    loss = mg.losses.KLD_standard_normal(mean, lv)
    assert loss.dim() == 0 and loss.dtype == torch.float32
    want = float(g['kld_%s_loss' % tag])
    assert abs(loss.item() - want) <= 1e-6 * abs(want)
    (0.25 * loss).backward()
    np.testing.assert_allclose(mean.grad.cpu().numpy(), g['kld_%s_grad_mean' % tag], rtol=3e-6, atol=1e-9)
    np.testing.assert_allclose(lv.grad.cpu().numpy(), g['kld_%s_grad_lv' % tag], rtol=3e-6, atol=1e-9)


def test_kld_standard_normal_large_and_deterministic(mg):
    rng = np.random.default_rng(4)
    m, lv = rng.standard_normal((1000, 257)).astype(np.float32), (0.3 * rng.standard_normal((1000, 257))).astype(np.float32)
    md, lvd = dev(m).requires_grad_(), dev(lv).requires_grad_()
    loss = mg.losses.KLD_standard_normal(md, lvd)
    want, want_gm, want_glv = O.kld_standard_normal(m, lv)
    assert abs(loss.item() - want) <= 1e-6 * abs(want)
    assert mg.losses.KLD_standard_normal(md, lvd).item() == loss.item()
    loss.backward()
    np.testing.assert_allclose(md.grad.cpu().numpy(), want_gm, rtol=REL, atol=1e-12)       # measured 1.1e-7
    # 1 - exp(lv) cancels near lv = 0: relative error is meaningless there (the reference's fp32 path has the same property)
    np.testing.assert_allclose(lvd.grad.cpu().numpy(), want_glv, rtol=3e-6, atol=1e-10)


def test_detach_batched_seqs_golden(mg, golden):
    """utils.detach_batched_seqs on CUDA tensors (packed on the device, one copy of the valid rows) against the lists the
    reference returns (utils.py:66-102): shapes after squeeze included (length-1 and length-0 utterances)."""
    g = golden('detach')
    x, y, n = dev(g['detach_x']).requires_grad_(), dev(g['detach_y']), dev(g['detach_n'])
    xs, ys = mg.utils.detach_batched_seqs(x, y, seq_len=n)
    raw = mg.utils.detach_batched_seqs(y, seq_len=g['detach_n'], squeeze=False)
    assert len(xs) == len(ys) == len(raw) == 5
    for b in range(5):
        for got, want in ((xs[b], g['detach_x_%d' % b]), (ys[b], g['detach_y_%d' % b]), (raw[b], g['detach_y_raw_%d' % b])):
            assert isinstance(got, np.ndarray) and got.shape == want.shape and np.array_equal(got, want), b
    full = mg.utils.detach_batched_seqs(x)
    assert isinstance(full, np.ndarray) and np.array_equal(full, g['detach_full'])


def test_both_voiced_mask_golden(mg, golden):
    """utils.both_voiced_mask against outputs of the reference (utils.py:169-172): NaN and -0. included, dtype argument kept."""
    g = golden('voiced_mask')
    a, b, c = dev(g['voiced_a']), dev(g['voiced_b']), dev(g['voiced_c'])
    got = mg.utils.both_voiced_mask(a, b)
    assert got.dtype == torch.uint8 and np.array_equal(got.cpu().numpy(), g['voiced_mask_ab'])
    got = mg.utils.both_voiced_mask(a, b, c, dtype=torch.cuda.FloatTensor)
    assert got.dtype == torch.float32 and np.array_equal(got.cpu().numpy(), g['voiced_mask_abc_f32'])
    assert np.array_equal(mg.utils.both_voiced_mask(a).cpu().numpy(), g['voiced_mask_a'])
    big = torch.randn(64, 1200, 1, device='cuda')
    big[big.abs() < 0.3] = 0.
    other = torch.randn(64, 1200, 1, device='cuda').round()
    on_device = mg.utils.both_voiced_mask(big, other, dtype=torch.uint8)       # a dtype (not a CPU tensor type) keeps the device
    assert on_device.is_cuda and torch.equal(on_device.bool(), (big != 0) & (other != 0))
    assert not mg.utils.both_voiced_mask(big, other).is_cuda                 # the reference's default lands on the host


@pytest.mark.parametrize('lo,hi,width', [(4, 184, 187), (1, 60, 60), (3, 100, 101), (2, 9, 12), (0, 5, 9)])
@pytest.mark.parametrize('mode', ['0', '1'])
def test_wide_column_slices_stream_flat(mg, monkeypatch, lo, hi, width, mode):
    """The slices LSTMAcousticModel.loss takes of one (B, T, 187) tensor (models/RNN_SPSS.py:133-135) and MelCepDistortion's
    target[..., 1:] (metrics.py:690): at least half of the row -> the flat masked stream; same values as thread-per-column."""
    monkeypatch.setenv('MG_RED_SLICE_MODE', mode)
    rng = np.random.default_rng(lo + hi + width)
    B, T = 7, 61
    n = rng.integers(0, T + 1, B)
    n[0], n[1] = T, 1
    p = rng.standard_normal((B, T, width)).astype(np.float32)
    y = rng.standard_normal((B, T, width)).astype(np.float32)
    pd, yd = dev(p).requires_grad_(), dev(y)
    for kind in ('mse', 'l1'):
        n_pos = np.maximum(n, 1)
        loss = getattr(mg.losses, kind)(pd[..., lo:hi], yd[..., lo:hi], dev(n_pos))
        want = O.masked_loss(p[..., lo:hi], y[..., lo:hi], n_pos, kind)
        assert abs(loss.item() - want) <= 1e-6 * abs(want)
        grad, = torch.autograd.grad(loss, pd)
        want_grad = np.zeros_like(p)
        want_grad[..., lo:hi] = O.masked_loss_grad(p[..., lo:hi], y[..., lo:hi], n_pos, kind)
        np.testing.assert_allclose(grad.cpu().numpy(), want_grad, rtol=REL, atol=1e-12)
    rmse = mg.metrics.RMSE()
    rmse.reset_state()
    rmse.accumulate(yd[..., lo:hi], pd.detach()[..., lo:hi], seq_len=dev(n))
    s_, c_ = O.rmse_acc(y[..., lo:hi], p[..., lo:hi], n)
    assert float(rmse.count) == c_ and abs(float(rmse.sum) - s_) <= 2e-6 * s_


# ----------------------------------------------------------------------------------------------------------------------
# a14: dense layers on tcgen05 (bf16 operands, fp32 accumulate).  Tolerances: (1) against the fp64 product of the
# bf16-ROUNDED operands only accumulation order and the sigmoid approximation differ: 2e-3 absolute on O(1) activations;
# (2) against the full-precision layer the bf16 operand rounding dominates: <= 2% of the output range.
# ----------------------------------------------------------------------------------------------------------------------
def _bf16_round(a):
    return torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize('M,K,N', [(70, 600, 64), (33, 609, 48), (45, 256, 187), (19, 32, 1), (1000, 600, 512),
                                   (257, 512, 256), (128, 128, 32), (5, 8, 16), (4097, 640, 199)])
@pytest.mark.parametrize('act', [None, 'sigmoid'])
def test_linear_tcgen05_vs_oracle(mg, M, K, N, act):
    rng = np.random.default_rng(M + K + N)
    x = rng.random((M, K), dtype=np.float32)
    w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = (0.1 * rng.standard_normal(N)).astype(np.float32)
    y = mg.ops.linear_bf16(dev(x), dev(w), dev(b), act=act).cpu().numpy()
    assert y.shape == (M, N) and y.dtype == np.float32
    exact_on_rounded = O.linear(_bf16_round(x), _bf16_round(w), b, act)
    np.testing.assert_allclose(y, exact_on_rounded, rtol=2e-3, atol=2e-3)
    full = O.linear(x, w, b, act)
    assert np.abs(y - full).max() <= 2e-2 * max(1.0, np.abs(full).max())
    y16 = mg.ops.linear_bf16(dev(x), dev(w), dev(b), act=act, out_dtype=torch.bfloat16).float().cpu().numpy()
    np.testing.assert_allclose(y16, exact_on_rounded, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize('M,N', [(70, 64), (1000, 512), (45, 187), (19, 1), (700, 3), (4097, 199), (0, 16)])
@pytest.mark.parametrize('act', [None, 'sigmoid'])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_act_grad_bf16_vs_oracle(mg, M, N, act, dtype):
    """K7g: sigmoid backward + bf16 cast + bias gradient in one pass.  The bf16 rows are the round-to-nearest-even of the
    fp32 ATen formula (bit-exact); the bias gradient is the fp64 column sum of the unrounded values, rounded once."""
    rng = np.random.default_rng(M + N)
    grad = torch.from_numpy(rng.standard_normal((M, N)).astype(np.float32)).to(dtype)
    y = torch.from_numpy(rng.random((M, N), dtype=np.float32)).to(dtype) if act else None
    g16, bias_grad = mg.ops.act_grad_bf16(grad.cuda(), None if y is None else y.cuda())
    n_pad = (N + 7) // 8 * 8
    assert g16.shape == (M, n_pad) and g16.dtype == torch.bfloat16 and bias_grad.shape == (N,)
    g32 = grad.float().numpy() if y is None else O.sigmoid_grad(grad.float().numpy(), y.float().numpy())
    assert g32.dtype == np.float32
    want16 = torch.from_numpy(g32).to(torch.bfloat16)
    assert torch.equal(g16[:, :N].cpu().view(torch.int16), want16.view(torch.int16))
    assert not g16[:, N:].any()
    want_bias = g32.astype(np.float64).sum(0)
    np.testing.assert_allclose(bias_grad.cpu().numpy(), want_bias, rtol=1e-6, atol=1e-6 * np.abs(g32).sum(0).max() if M else 0)
    only, none = mg.ops.act_grad_bf16(grad.cuda(), None if y is None else y.cuda(), want_bias_grad=False)
    assert none is None and torch.equal(only, g16)


@pytest.mark.parametrize('M,N,K', [(70, 64, 600), (33, 48, 609), (45, 187, 256), (19, 1, 32), (5000, 512, 600), (257, 256, 512),
                                   (128, 32, 128), (5, 16, 8), (4097, 199, 640), (20001, 128, 512), (1, 3, 64), (9000, 130, 130)])
@pytest.mark.parametrize('pair,frames', [(None, None), ('0', None), ('1', None), ('1', '64'), ('0', '64')])
def test_linear_wgrad_tcgen05_vs_oracle(mg, monkeypatch, M, N, K, pair, frames):
    """K7w: g^T @ x over the frame axis with MN-major tcgen05 operands, as single CTAs and as CTA pairs (cta_group::2).  Against the fp64 product of the same bf16 operands
    only the fp32 accumulation differs: 1e-4 of the largest entry; and the split reduction has a fixed order (same bits twice)."""
    if pair is not None:
        monkeypatch.setenv('MG_WGRAD_PAIR', pair)
    if frames is not None:
        monkeypatch.setenv('MG_WGRAD_FRAMES', frames)
    rng = np.random.default_rng(M + N + K)
    g = torch.from_numpy(rng.standard_normal((M, N)).astype(np.float32))
    x = torch.from_numpy(rng.random((M, K), dtype=np.float32))
    g16 = mg.ops.act_grad_bf16(g.cuda(), None, want_bias_grad=False)[0]
    x16 = mg.ops.cast_pad_bf16(x.cuda())
    got = mg.ops.linear_wgrad_bf16(g16, x16, out_features=N, in_features=K)
    assert got.shape == (N, K) and got.dtype == torch.float32
    want = O.linear_wgrad(g16[:, :N].float().cpu().numpy(), x16[:, :K].float().cpu().numpy())
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-4, atol=1e-4 * max(1.0, np.abs(want).max()))
    again = mg.ops.linear_wgrad_bf16(g16, x16, out_features=N, in_features=K)
    assert torch.equal(got, again)


def test_linear_wgrad_empty_batch_is_zero(mg):
    g16 = torch.zeros((0, 16), dtype=torch.bfloat16, device='cuda')
    x16 = torch.zeros((0, 64), dtype=torch.bfloat16, device='cuda')
    got = mg.ops.linear_wgrad_bf16(g16, x16, out_features=12, in_features=60)
    assert got.shape == (12, 60) and not got.any()


def test_linear_golden(mg, golden):
    g = golden('linear')
    for case in ['readme_l1', 'rnn_in', 'out187', 'out1']:
        x, w, b = g['lin_%s_x' % case], g['lin_%s_w' % case], g['lin_%s_b' % case]
        y = mg.ops.linear_bf16(dev(x), dev(w), dev(b), act='sigmoid').cpu().numpy()
        assert np.abs(y - g['lin_%s_sig' % case]).max() <= 1e-2      # bf16 operands vs the reference's fp32 sgemm


# ----------------------------------------------------------------------------------------------------------------------
# K0: on-device collate
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('D,dtype', [(187, np.float32), (600, np.float32), (1, np.int64), (3, np.uint8), (8, np.float16)])
def test_pad_collate_matches_host_padding(mg, D, dtype):
    rng = np.random.default_rng(D)
    lengths = rng.integers(0, 40, 9)
    lengths[3] = 0
    rows = [(rng.random((n, D)) * 100).astype(dtype) for n in lengths]
    packed = np.concatenate(rows, axis=0)
    T = int(lengths.max())
    want = np.zeros((len(rows), T, D), dtype)                # what collate_fn builds on the host (data.py:184-193)
    for b, r in enumerate(rows):
        want[b, :len(r)] = r
    got = mg.data.pad_collate(dev(packed), dev(lengths))
    assert np.array_equal(got.cpu().numpy(), want)
    got = mg.data.pad_collate(dev(packed), dev(lengths), max_len=T + 3)
    assert np.array_equal(got[:, :T].cpu().numpy(), want) and not got[:, T:].any()


@pytest.mark.parametrize('M,K,N,act,out_dtype', [(300, 600, 512, 'sigmoid', torch.bfloat16), (1000, 40, 96, None, torch.float32),
                                                  (700, 136, 400, 'sigmoid', torch.float32), (513, 72, 1000, None, torch.bfloat16),
                                                  (517, 256, 187, None, torch.float32), (256, 64, 64, 'sigmoid', torch.float32)])
def test_linear_cta_pair_equals_single_cta(mg, monkeypatch, M, K, N, act, out_dtype):
    """The 2-CTA (cta_group::2, 256-row tile) form of K7 against the single-CTA form: same K order per output element, so
    the results are identical; ragged M (second CTA of the last pair partly / wholly outside the matrix) included."""
    from morgana_b200 import ops
    rng = np.random.default_rng(M + K + N)
    x = dev(rng.standard_normal((M, K)).astype(np.float32)).to(torch.bfloat16)
    w = dev((rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)).to(torch.bfloat16)
    if K % 8:
        x, w = torch.nn.functional.pad(x, (0, 8 - K % 8)), torch.nn.functional.pad(w, (0, 8 - K % 8))
    bias = dev(rng.standard_normal(N).astype(np.float32))
    monkeypatch.setenv('MG_GEMM_PAIR', '0')
    single = ops.linear_bf16(x, w, bias, act=act, out_dtype=out_dtype)
    monkeypatch.setenv('MG_GEMM_PAIR', '1')
    monkeypatch.setenv('MG_GEMM_WIDE', '0')
    paired = ops.linear_bf16(x, w, bias, act=act, out_dtype=out_dtype)
    assert torch.equal(single, paired)
    monkeypatch.setenv('MG_GEMM_WIDE', '1')                  # 256 x 512 pair tiles (taken when N > 256)
    wide = ops.linear_bf16(x, w, bias, act=act, out_dtype=out_dtype)
    assert torch.equal(single, wide)
    want = O.linear(x.float().cpu().numpy()[:, :K], w.float().cpu().numpy()[:, :K], bias.cpu().numpy(), act)
    np.testing.assert_allclose(paired.float().cpu().numpy(), want, rtol=2e-2, atol=2e-2)


def test_nn_linear_module_forward_backward(mg):
    """README MLP stack (600 -> 512 -> 128 -> 32 -> 1) through morgana_b200.nn.Linear against fp32 torch.nn layers."""
    from morgana_b200 import nn as mnn
    torch.manual_seed(0)
    dims = [600, 512, 128, 32, 1]
    ref = torch.nn.Sequential(*[m for i in range(4) for m in
                                ([torch.nn.Linear(dims[i], dims[i + 1])] + ([torch.nn.Sigmoid()] if i < 3 else []))]).cuda()
    ours = [mnn.Linear(dims[i], dims[i + 1], act='sigmoid' if i < 3 else None, device='cuda') for i in range(4)]
    ref_linears = [m for m in ref if isinstance(m, torch.nn.Linear)]
    for a, b in zip(ours, ref_linears):
        a.load_state_dict(b.state_dict())
    x = torch.rand(3, 50, 600, device='cuda')
    xo = x.clone().requires_grad_()
    h = xo
    for layer in ours:
        h = layer(h)
    want = ref(x)
    assert h.shape == want.shape == (3, 50, 1)
    assert (h - want).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())
    h.sum().backward()
    xr = x.clone().requires_grad_()
    ref(xr).sum().backward()
    for a, b in zip(ours, ref_linears):
        scale = b.weight.grad.abs().max().item()
        assert (a.weight.grad - b.weight.grad).abs().max().item() <= 5e-2 * scale + 1e-4
        assert (a.bias.grad - b.bias.grad).abs().max().item() <= 5e-2 * b.bias.grad.abs().max().item() + 1e-4
    # weight gradients: tcgen05 (MN-major operands); the input gradient of the 512-wide layer too, the narrower ones
    # through the library GEMM
    assert (xo.grad - xr.grad).abs().max().item() <= 5e-2 * xr.grad.abs().max().item() + 1e-6
    # the bf16 shadow follows parameter updates
    before = ours[0].weight_bf16().clone()
    with torch.no_grad():
        ours[0].weight.add_(1.0)
    assert not torch.equal(before, ours[0].weight_bf16())


@pytest.mark.parametrize('kind', [None, 'minmax', 'mvn'])
def test_fused_upsample_bf16_output_is_rounded_exact_result(mg, kind):
    rng = np.random.default_rng(17)
    B, P, D = 6, 25, 600
    x = rng.random((B, P, D), dtype=np.float32)
    dur = rng.integers(0, 9, (B, P))
    p0 = rng.standard_normal(D).astype(np.float32)
    p1 = (p0 + np.abs(rng.standard_normal(D)) + 0.1).astype(np.float32)
    norm = None if kind is None else (kind, dev(p0), dev(p1))
    exact = O.upsample_to_repetitions(x, dur) if kind is None else O.normalise_upsample(x, dur, kind, p0, p1)
    want = torch.from_numpy(exact).to(torch.bfloat16)                       # round-to-nearest-even of the exact result
    got = mg.utils.upsample_to_repetitions(dev(x), dev(dur), normaliser=norm, out_dtype=torch.bfloat16)
    assert got.dtype == torch.bfloat16 and tuple(got.shape) == exact.shape
    assert torch.equal(got.cpu().view(torch.int16), want.view(torch.int16))


def test_first_layer_commutes_with_the_expansion(mg):
    """Every frame row is a copy of an item row, so Linear(+Sigmoid) may run at item rate before the expansion: the same
    bits on every valid frame; padding frames hold 0 instead of sigmoid(bias)."""
    from morgana_b200 import ops
    rng = np.random.default_rng(23)
    B, P, D, N = 5, 17, 600, 512
    lab = dev(rng.random((B, P, D), dtype=np.float32))
    dur = dev(rng.integers(0, 12, (B, P)))
    mmin, mmax = dev(np.zeros(D, np.float32)), dev((rng.random(D) + 0.5).astype(np.float32))
    w = ops.cast_pad_bf16(dev((rng.standard_normal((N, D)) * 0.05).astype(np.float32)))
    bias = dev(rng.standard_normal(N).astype(np.float32))
    frames, n_frames = mg.utils.upsample_to_repetitions(lab, dur, normaliser=('minmax', mmin, mmax), out_dtype=torch.bfloat16,
                                                        return_lengths=True)
    T = frames.shape[1]
    at_frame_rate = ops.linear_bf16(frames.reshape(B * T, D), w, bias, act='sigmoid', out_dtype=torch.bfloat16).reshape(B, T, N)
    items = mg.data.normalise_minmax(lab, mmin, mmax).reshape(B * P, D)
    at_item_rate = ops.linear_bf16(items, w, bias, act='sigmoid', out_dtype=torch.bfloat16).reshape(B, P, N)
    expanded = mg.utils.upsample_to_repetitions(at_item_rate, dur, max_len=T)          # bf16 rows: the byte-copy path
    valid = torch.arange(T, device='cuda')[None] < n_frames[:, None]
    assert expanded.dtype == torch.bfloat16 and valid.any() and not valid.all()
    assert torch.equal(expanded[valid].view(torch.int16), at_frame_rate[valid].view(torch.int16))
    assert not expanded[~valid].any()


def test_nn_linear_follows_fused_optimizer_updates(mg):
    """Fused Adam updates parameters without bumping the autograd version counter: the bf16 shadow must still follow."""
    from morgana_b200 import nn as mnn
    torch.manual_seed(1)
    layer = mnn.Linear(64, 16, device='cuda')
    opt = torch.optim.Adam(layer.parameters(), lr=0.1, fused=True)
    x = torch.rand(9, 64, device='cuda')
    y0 = layer(x).detach().clone()
    layer(x).sum().backward()
    opt.step()
    y1 = layer(x).detach()
    want = torch.nn.functional.linear(x, layer.weight.detach(), layer.bias.detach())
    assert (y1 - y0).abs().max().item() > 0.05
    assert (y1 - want).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())


# ----------------------------------------------------------------------------------------------------------------------
# "next" row 1: batched MLPG.  Parity is against the oracle's restatement of the definition (dense window matrices, dense
# fp64 solve); the reference's own solver (bandmat) is absent here, so this row is "parity unpinned" (DESIGN.md).
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('B,T,F,pad', [(3, 40, 5, 0), (4, 57, 3, 10), (2, 120, 60, 100), (5, 33, 1, 3)])
@pytest.mark.parametrize('var_kind', ['global', 'frame'])
def test_mlpg_vs_oracle(mg, B, T, F, pad, var_kind):
    from morgana_b200.viz.synthesis import MLPG
    rng = np.random.default_rng(B * 100 + T + F + pad)
    means = rng.standard_normal((B, T, 3 * F)).astype(np.float32)
    seq_len = rng.integers(1, T + 1, B)
    seq_len[0] = T
    if var_kind == 'global':
        var = (rng.random(3 * F) + 0.3).astype(np.float32)
    else:
        var = (rng.random((B, T, 3 * F)) + 0.3).astype(np.float32)
    want = O.mlpg(means, var, padding_size=pad, seq_len=seq_len)
    got = MLPG(dev(means), dev(var), padding_size=pad, seq_len=dev(seq_len))
    assert got.dtype == torch.float32 and tuple(got.shape) == (B, T, F)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-5)       # fp64 solve, fp32 output
    for b in range(B):
        assert not got[b, seq_len[b]:].any()                                         # out-of-sequence frames are zero
    as_np = MLPG(means, var, padding_size=pad, seq_len=seq_len)                       # NumPy in -> NumPy out, as the reference
    assert isinstance(as_np, np.ndarray) and as_np.dtype == np.float64
    np.testing.assert_allclose(as_np, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('pad', [0, 100])
def test_mlpg_chunked_solve_at_every_split(mg, pad):
    """The solve cuts a sequence into 1..32 chunks depending on its padded length: lengths on both sides of every change
    of the chunk count, and long utterances, against the banded fp64 solve of the oracle."""
    from morgana_b200.viz.synthesis import MLPG
    rng = np.random.default_rng(99 + pad)
    lengths = np.array([1, 2, 3, 11, 12, 23, 24, 25, 35, 36, 37, 47, 48, 59, 60, 383, 384, 385, 777, 1500])
    B, T, F = len(lengths), 1500, 3
    means = rng.standard_normal((B, T, 3 * F)).astype(np.float32)
    var = (rng.random(3 * F) * 2 + 0.05).astype(np.float32)
    want = O.mlpg_banded(means, var, padding_size=pad, seq_len=lengths)
    got = MLPG(dev(means), dev(var), padding_size=pad, seq_len=dev(lengths)).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
    for b, n in enumerate(lengths):
        assert not got[b, n:].any()


def test_mlpg_static_only_limit(mg):
    """With enormous delta variances the deltas carry no information and the trajectory is the static mean."""
    from morgana_b200.viz.synthesis import MLPG
    rng = np.random.default_rng(3)
    means = rng.standard_normal((2, 50, 6)).astype(np.float32)
    var = np.array([1., 1., 1e12, 1e12, 1e12, 1e12], np.float32)
    got = MLPG(dev(means), dev(var), padding_size=0).cpu().numpy()
    np.testing.assert_allclose(got, means[..., :2], rtol=1e-4, atol=1e-4)


# ----------------------------------------------------------------------------------------------------------------------
# "next" row 4: sibling segment operations (bit-exact: pure row movers)
# ----------------------------------------------------------------------------------------------------------------------
def test_segment_ops_golden(mg, golden):
    g = golden('segments')
    U = mg.utils
    x, lens = dev(g['seg_x']), dev(g['seg_lens'])
    assert np.array_equal(U.batched_masked_select(x, dev(g['seg_seq_len'])).cpu().numpy(), g['seg_select'])
    assert np.array_equal(U.get_segment_ends(x, lens[:, :, None]).cpu().numpy(), g['seg_ends'])
    assert np.array_equal(U.split_to_segments(x, lens[:, :, None]).cpu().numpy(), g['seg_split'])
    xi, li = dev(g['seg_int_x']), dev(g['seg_int_lens'])
    assert np.array_equal(U.get_segment_ends(xi, li).cpu().numpy(), g['seg_int_ends'])
    assert np.array_equal(U.split_to_segments(xi, li).cpu().numpy(), g['seg_int_split'])
    assert np.array_equal(U.batched_masked_select(xi, torch.tensor([9, 1, 4], device='cuda')).cpu().numpy(), g['seg_int_select'])


@pytest.mark.parametrize('D', [1, 7, 600])
def test_segment_ops_vs_oracle_and_collate_round_trip(mg, D):
    rng = np.random.default_rng(D)
    B, T, S = 9, 61, 8
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    seq_len = rng.integers(0, T + 1, B)
    lens = rng.integers(0, 12, (B, S))
    lens[rng.random((B, S)) < 0.2] = 0
    while (lens.sum(axis=1) > T).any():
        lens[lens.sum(axis=1) > T] //= 2
    U = mg.utils
    packed = U.batched_masked_select(dev(x), dev(seq_len))
    assert np.array_equal(packed.cpu().numpy(), O.batched_masked_select(x, seq_len))
    # pack is the inverse of the on-device collate
    padded = mg.data.pad_collate(packed, dev(seq_len), max_len=T).cpu().numpy()
    mask = np.arange(T)[None, :] < seq_len[:, None]
    assert np.array_equal(padded[mask], x[mask]) and not padded[~mask].any()
    assert np.array_equal(U.get_segment_ends(dev(x), dev(lens)).cpu().numpy(), O.get_segment_ends(x, lens))
    assert np.array_equal(U.split_to_segments(dev(x), dev(lens)).cpu().numpy(), O.split_to_segments(x, lens))


def test_torch_custom_ops_match_the_direct_path(mg, golden):
    g = golden('upsample')
    x, dur, want = dev(g['ups_lab600_f32_x']), dev(g['ups_lab600_f32_dur']), g['ups_lab600_f32_out']
    got = torch.ops.morgana_b200.upsample_norm(x, dur, None, None, 'none', -1)
    assert np.array_equal(got.cpu().numpy(), want)
    gl = golden('losses')
    p, y, n = dev(gl['loss_wide_pred']), dev(gl['loss_wide_tgt']), dev(gl['loss_wide_seq_len'])
    value = torch.ops.morgana_b200.masked_loss(p, y, n, 'mse')
    assert rel_err(value.item(), gl['loss_wide_mse_masked']) <= REL
    s, q = torch.randn(100, device='cuda'), torch.randn(100, device='cuda')
    want_s = O.ema_update(s.cpu().numpy().copy(), q.cpu().numpy(), 0.99)
    torch.ops.morgana_b200.ema_update([s], [q], 1.0 - 0.99)
    assert np.array_equal(s.cpu().numpy(), want_s)


# ----------------------------------------------------------------------------------------------------------------------
# "next" row 2: the epoch loops on the device (single process; the two-rank logic is covered by tests/test_dp_gloo.py)
# ----------------------------------------------------------------------------------------------------------------------
def test_trainer_epochs_on_device(mg):
    """DataParallelTrainer with the real pieces: tcgen05 layers, masked mse through K4, device-resident metric records,
    the multi-tensor EMA, a prefetching feeder; against the same loop written out with stock torch ops in fp32 (built at the
    end of this test: same initial weights, same batches, `torch.nn.Linear`, `repeat_interleave`, the masked mean written
    out, un-fused Adam, the EMA formula).  Tolerance of the comparison: the two layers run on bf16 operands, so the epoch
    losses agree to 2 % and the trained / averaged weights to 2 % of their range -- not to fp32 round-off."""
    from morgana_b200 import nn as mnn, trainer as T

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(5)
            self.l1 = mnn.Linear(40, 64, act='sigmoid', device='cuda')
            self.l2 = mnn.Linear(64, 3, device='cuda')
            self.mode, self.step = '', 0
            self.metrics = mg.metrics.Handler(loss=mg.metrics.Mean())
            self.metrics.add_metrics('all', err=mg.metrics.RMSE())

        def forward(self, features):
            frames = mg.utils.upsample_to_repetitions(features['lab'], features['dur'], max_len=features['T'])
            pred = self.l2(self.l1(frames))
            self.metrics.accumulate(self.mode, err=(features['target'], pred.detach(), features['n_frames']))
            return mg.losses.mse(pred, features['target'], features['n_frames']), {'pred': pred}

    g = torch.Generator().manual_seed(9)
    batches = []
    for _ in range(4):
        dur = torch.randint(1, 6, (6, 11, 1), generator=g)
        n_frames = dur.sum(dim=(1, 2))
        lab = torch.rand(6, 11, 40, generator=g)
        frames = torch.from_numpy(O.upsample_to_repetitions(lab.numpy(), dur.numpy()[:, :, 0]))
        target = torch.stack([frames[:, :, :5].sum(-1), frames[:, :, 5:9].mean(-1), frames[:, :, 9] * 2.], dim=-1)
        batches.append({'lab': lab, 'dur': dur, 'n_frames': n_frames, 'target': target, 'T': int(n_frames.max()), 'name': ['x']})
    model, ema_model = Model(), Model()
    tr = T.DataParallelTrainer(model, ema_model=ema_model, ema_decay=0.9)
    before = [p.detach().clone() for p in ema_model.parameters()]
    opt = torch.optim.Adam(model.parameters(), lr=0.01, fused=True)
    losses = []
    for tr.epoch in range(1, 6):
        losses.append(tr.train_epoch(mg.data.ToDeviceWrapper(batches, 'cuda'), opt))
    assert model.step == 20 and losses[-1] < 0.6 * losses[0]
    train = model.metrics.results_as_json_dict('train')
    assert set(train) == {'loss', 'err'} and train['loss'] == pytest.approx(losses[-1], rel=1e-5)
    # count is frames, the sum runs over frames x 3 dims (Q2): err^2 ~ 3 x the frame-weighted mse of the epoch
    assert 0.5 * 3 * train['loss'] < train['err'] ** 2 < 2. * 3 * train['loss']
    assert all(p.grad.untyped_storage().data_ptr() == tr.bucket.flat.untyped_storage().data_ptr() for p in model.parameters())
    # the EMA model moved towards the trained weights and validates into its own handler
    assert all(not torch.equal(a, b.detach()) for a, b in zip(before, ema_model.parameters()))
    valid_loss = tr.valid_epoch(mg.data.ToDeviceWrapper(batches, 'cuda'), model=tr.ema.model)
    assert np.isfinite(valid_loss) and ema_model.metrics.results_as_json_dict('valid')['loss'] == pytest.approx(valid_loss, rel=1e-5)

    # ---- the same five epochs with stock torch ops in fp32 (what experiment_builder.py:464-490 does on the reference's ops) --
    torch.manual_seed(5)
    stock = torch.nn.Sequential(torch.nn.Linear(40, 64), torch.nn.Sigmoid(), torch.nn.Linear(64, 3)).cuda()
    fresh = Model()                                             # same seed: the weights `model` started from
    with torch.no_grad():
        stock[0].weight.copy_(fresh.l1.weight); stock[0].bias.copy_(fresh.l1.bias)
        stock[2].weight.copy_(fresh.l2.weight); stock[2].bias.copy_(fresh.l2.bias)
    stock_ema = [p.detach().clone() for p in stock.parameters()]
    stock_opt = torch.optim.Adam(stock.parameters(), lr=0.01)
    stock_losses = []
    for epoch in range(5):
        total = 0.
        for b in batches:
            lab, dur, n_frames, target = (b[k].cuda() for k in ('lab', 'dur', 'n_frames', 'target'))
            T_max = b['T']
            frames = torch.zeros(lab.shape[0], T_max, lab.shape[2], device='cuda')
            for i in range(lab.shape[0]):                        # np.repeat per utterance (utils.py:176)
                rows = torch.repeat_interleave(lab[i], dur[i, :, 0], dim=0)
                frames[i, :rows.shape[0]] = rows
            pred = stock(frames)
            mask = (torch.arange(T_max, device='cuda')[None, :] < n_frames[:, None])[:, :, None].float()
            loss = (((pred - target) ** 2 * mask).sum(dim=1) / n_frames[:, None].float()).mean()      # losses.py:34-42
            stock_opt.zero_grad()
            loss.backward()
            stock_opt.step()
            with torch.no_grad():
                for s_, p_ in zip(stock_ema, stock.parameters()):
                    s_ -= (1.0 - 0.9) * (s_ - p_)                  # utils.py:447-448
            total += loss.item()
        stock_losses.append(total / len(batches))
    for ours, theirs in zip(losses, stock_losses):
        assert abs(ours - theirs) <= 2e-2 * abs(theirs), (losses, stock_losses)
    for ours, theirs in zip(model.parameters(), stock.parameters()):
        assert float((ours - theirs).abs().max()) <= 2e-2 * float(theirs.abs().max())
    for ours, theirs in zip(ema_model.parameters(), stock_ema):
        assert float((ours - theirs).abs().max()) <= 2e-2 * float(theirs.abs().max())
