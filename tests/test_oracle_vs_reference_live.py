"""The CPU oracle against the UNMODIFIED reference, executed live on randomised shapes -- in the build container only.

The committed fixtures (tests/golden/*.npz) pin the oracle at a handful of shapes; here the same reference functions are
called on seeded random walks over shapes, lengths and dtypes and compared with ``oracle/np_oracle.py`` (and the timed op
chain ``oracle/aten_chain.py``).  ``/root/reference`` does not exist on the GPU box: the whole module is skipped there, and
nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` depends on it.
"""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get('MORGANA_REFERENCE_ROOT', '/root/reference')
if not os.path.isdir(os.path.join(REFERENCE_ROOT, 'morgana')):
    pytest.skip('the reference tree is not present (GPU box / clean checkout)', allow_module_level=True)

sys.path.insert(0, os.path.join(HERE, 'golden'))
sys.dont_write_bytecode = True              # never write into /root/reference
from ref_import import import_reference  # noqa: E402

morgana = import_reference()
from morgana import data as ref_data, losses as ref_losses, metrics as ref_metrics, utils as ref_utils  # noqa: E402
from oracle import aten_chain as A, np_oracle as O  # noqa: E402

SEEDS = [0, 1, 2, 3, 4, 5]
REL = 1e-6


def _shape(rng, max_b=9, max_t=40, dims=(1, 3, 7, 60)):
    return int(rng.integers(1, max_b)), int(rng.integers(1, max_t)), int(rng.choice(dims))


def _close(got, want, rel=REL):
    want = float(want)
    return abs(float(got) - want) <= rel * max(abs(want), 1e-30)


@pytest.mark.parametrize('seed', SEEDS)
def test_upsample_and_masks(seed):
    rng = np.random.default_rng(seed)
    B, P, D = _shape(rng)
    x = rng.standard_normal((B, P, D)).astype(np.float32)
    dur = rng.integers(0, 6, (B, P, 1))
    dur[rng.integers(0, B)] = 0                                            # an empty utterance
    want = ref_utils.upsample_to_repetitions(torch.from_numpy(x), torch.from_numpy(dur)).numpy()
    assert np.array_equal(O.upsample_to_repetitions(x, dur), want)
    assert np.array_equal(A.upsample_chain(torch.from_numpy(x), torch.from_numpy(dur)).numpy(), want)
    xi = rng.integers(-50, 50, (B, P, D))
    assert np.array_equal(O.upsample_to_repetitions(xi, dur), ref_utils.upsample_to_repetitions(torch.from_numpy(xi), torch.from_numpy(dur)).numpy())
    n = rng.integers(0, P + 1, B)
    for max_len, dtype, np_dtype in ((None, torch.ByteTensor, np.uint8), (P + 3, torch.float32, np.float32)):
        if max_len is None and n.max() == 0:
            continue
        want_mask = ref_utils.sequence_mask(torch.from_numpy(n), max_len=max_len, dtype=dtype).numpy()
        assert np.array_equal(O.sequence_mask(n, max_len, np_dtype), want_mask)
    a, b = rng.standard_normal((B, P, 1)).astype(np.float32), rng.standard_normal((B, P, 1)).astype(np.float32)
    a[rng.random((B, P, 1)) < 0.4] = 0.
    assert np.array_equal(O.both_voiced_mask(a, b), ref_utils.both_voiced_mask(torch.from_numpy(a), torch.from_numpy(b)).numpy())
    if n.sum() > 0:
        assert np.array_equal(O.batched_masked_select(x, n), ref_utils.batched_masked_select(torch.from_numpy(x), torch.from_numpy(n)).numpy())


@pytest.mark.parametrize('seed', SEEDS)
def test_normalisers_bit_exact(seed):
    rng = np.random.default_rng(100 + seed)
    B, T, D = _shape(rng)
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    mean, std = rng.standard_normal(D).astype(np.float32), (np.abs(rng.standard_normal(D)) + 0.05).astype(np.float32)
    mmin = rng.standard_normal(D).astype(np.float32)
    mmax = (mmin + np.abs(rng.standard_normal(D))).astype(np.float32)
    mmax[0] = mmin[0]                                                      # a constant dimension: scale forced to 1 (data.py:582)
    tx = torch.from_numpy(x)
    assert np.array_equal(O.normalise_mvn(x, mean, std), ref_data.normalise_mvn(tx, torch.from_numpy(mean), torch.from_numpy(std)).numpy())
    assert np.array_equal(O.denormalise_mvn(x, mean, std), ref_data.denormalise_mvn(tx, torch.from_numpy(mean), torch.from_numpy(std)).numpy())
    assert np.array_equal(O.normalise_minmax(x, mmin, mmax), ref_data.normalise_minmax(tx, torch.from_numpy(mmin), torch.from_numpy(mmax)).numpy())
    assert np.array_equal(O.denormalise_minmax(x, mmin, mmax), ref_data.denormalise_minmax(tx, torch.from_numpy(mmin), torch.from_numpy(mmax)).numpy())
    assert np.array_equal(A.normalise_minmax_chain(tx, torch.from_numpy(mmin), torch.from_numpy(mmax)).numpy(), O.normalise_minmax(x, mmin, mmax))


@pytest.mark.parametrize('seed', SEEDS)
def test_losses_and_gradients(seed):
    rng = np.random.default_rng(200 + seed)
    B, T, D = _shape(rng)
    n = rng.integers(1, T + 1, B)
    p = rng.standard_normal((B, T, D)).astype(np.float32)
    y = rng.standard_normal((B, T, D)).astype(np.float32)
    for seq_len in (n, None):
        tn = None if seq_len is None else torch.from_numpy(seq_len)
        tp = torch.from_numpy(p).requires_grad_()
        want = ref_losses.mse(tp, torch.from_numpy(y), seq_len=tn)
        want.backward()
        assert _close(O.masked_loss(p, y, seq_len, 'mse'), want.item())
        np.testing.assert_allclose(O.masked_loss_grad(p, y, seq_len, 'mse'), tp.grad.numpy(), rtol=2e-6, atol=1e-10)
        assert _close(A.mse_chain(torch.from_numpy(p), torch.from_numpy(y), tn).item(), want.item())
    prob = (1. / (1. + np.exp(-p))).astype(np.float32)
    label = (rng.random((B, T, D)) < 0.5).astype(np.float32)
    want = ref_losses.bce(torch.from_numpy(prob), torch.from_numpy(label), seq_len=torch.from_numpy(n)).item()
    assert _close(O.masked_loss(prob, label, n, 'bce'), want)
    C = int(rng.integers(2, 9))
    logits = rng.standard_normal((B, T, C)).astype(np.float32)
    classes = rng.integers(0, C, (B, T))
    tl = torch.from_numpy(logits).requires_grad_()
    want = ref_losses.ce(tl, torch.from_numpy(classes), seq_len=torch.from_numpy(n))
    want.backward()
    got, got_grad = O.cross_entropy_loss(logits, classes, n)
    assert _close(got, want.item(), 2e-6)
    np.testing.assert_allclose(got_grad, tl.grad.numpy(), rtol=3e-6, atol=1e-8)
    mean, lv = rng.standard_normal((B, D)).astype(np.float32), (0.5 * rng.standard_normal((B, D))).astype(np.float32)
    assert _close(O.kld_standard_normal(mean, lv)[0], ref_losses.KLD_standard_normal(torch.from_numpy(mean), torch.from_numpy(lv)).item(), 2e-6)


@pytest.mark.parametrize('seed', SEEDS)
def test_metric_accumulators(seed):
    rng = np.random.default_rng(300 + seed)
    B, T, D = _shape(rng, dims=(1, 5, 60))
    n = rng.integers(1, T + 1, B)
    tn = torch.from_numpy(n)
    y = rng.standard_normal((B, T, D)).astype(np.float32)
    p = (y + 0.3 * rng.standard_normal((B, T, D))).astype(np.float32)
    ty, tp = torch.from_numpy(y), torch.from_numpy(p)

    def state(metric, *args):
        metric.reset_state()
        metric.accumulate(*args, seq_len=tn)
        return float(metric.sum), float(metric.count)

    for name, acc in (('RMSE', O.rmse_acc), ('MAE', O.mae_acc), ('Distortion', O.distortion_acc)):
        want_sum, want_count = state(getattr(ref_metrics, name)(), ty, tp)
        got_sum, got_count = acc(y, p, n)
        assert got_count == want_count and _close(got_sum, want_sum, 2e-6), name
    if D > 1:
        want_sum, want_count = state(ref_metrics.MelCepDistortion(), ty, tp)
        got_sum, got_count = O.melcep_acc(y, p, n)
        assert got_count == want_count and _close(got_sum, want_sum, 2e-6)
    want_sum, want_count = state(ref_metrics.Mean(), ty)
    got_sum, got_count = O.mean_acc(y, n)
    assert got_count == want_count and abs(got_sum - want_sum) <= 2e-6 * np.abs(y).sum()
    lf0_t = (5 + 0.3 * rng.standard_normal((B, T, 1))).astype(np.float32)
    lf0_p = (lf0_t + 0.05 * rng.standard_normal((B, T, 1))).astype(np.float32)
    voiced = rng.random((B, T, 1)) < 0.6
    want_sum, want_count = state(ref_metrics.LF0Distortion(), torch.from_numpy(lf0_t), torch.from_numpy(lf0_p), torch.from_numpy(voiced.copy()))
    got_sum, got_count = O.lf0_acc(lf0_t, lf0_p, voiced, n)
    assert got_count == want_count and (want_count == 0 or _close(got_sum, want_sum, 2e-6))
    bits_a, bits_b = rng.random((B, T, 1)) < 0.5, rng.random((B, T, 1)) < 0.5
    for name, acc in (('Error', O.error_acc), ('Accuracy', O.accuracy_acc)):
        want_sum, want_count = state(getattr(ref_metrics, name)(), torch.from_numpy(bits_a), torch.from_numpy(bits_b))
        got_sum, got_count = acc(bits_a, bits_b, n)
        assert int(got_sum) == int(want_sum) and got_count == want_count, name


@pytest.mark.parametrize('seed', SEEDS)
def test_ema_and_segments(seed):
    rng = np.random.default_rng(400 + seed)
    model, other = torch.nn.Linear(int(rng.integers(1, 20)), int(rng.integers(1, 20))), None
    other = torch.nn.Linear(model.in_features, model.out_features)
    decay = float(rng.choice([0.9, 0.99, 0.999]))
    before = [prm.detach().numpy().copy() for prm in model.parameters()]
    ema = ref_utils.ExponentialMovingAverage(model, decay)
    ema.update_params(other)
    for old, new, src in zip(before, ema.model.parameters(), other.parameters()):
        assert np.array_equal(O.ema_update(old.copy(), src.detach().numpy(), decay), new.detach().numpy())
    B, T, D = _shape(rng)
    S = int(rng.integers(1, 6))
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    lens = rng.integers(0, max(1, T // S) + 1, (B, S))
    if lens.sum(1).max() > 0 and lens.max() > 0:
        tl = torch.from_numpy(lens)[:, :, None]
        assert np.array_equal(O.get_segment_ends(x, lens), ref_utils.get_segment_ends(torch.from_numpy(x), tl).numpy())
        assert np.array_equal(O.split_to_segments(x, lens), ref_utils.split_to_segments(torch.from_numpy(x), tl).numpy())


@pytest.mark.parametrize('kind', ['mvn', 'minmax'])
def test_normaliser_classes_numpy_path(kind, tmp_path):
    """The drop-in normaliser classes on NumPy inputs (what DataLoader workers call, data.py:119-127) against the reference's
    classes loaded from the same JSON files: same file names, same parameters, bit-identical results, deltas included."""
    import json
    mg_data = pytest.importorskip('morgana_b200.data')
    rng = np.random.default_rng(7)
    D = 9
    if kind == 'mvn':
        params = {'mean': rng.standard_normal(D).tolist(), 'std_dev': (np.abs(rng.standard_normal(D)) + 0.1).tolist()}
        ref_cls, our_cls = ref_data.MeanVarianceNormaliser, mg_data.MeanVarianceNormaliser
    else:
        lo = rng.standard_normal(D)
        hi = lo + np.abs(rng.standard_normal(D))
        hi[2] = lo[2]
        params = {'mmin': lo.tolist(), 'mmax': hi.tolist()}
        ref_cls, our_cls = ref_data.MinMaxNormaliser, mg_data.MinMaxNormaliser
    (tmp_path / 'norm').mkdir()
    for name in ('lab', 'lab_deltas'):
        with open(tmp_path / 'norm' / ('%s_%s.json' % (name, kind)), 'w') as f:
            json.dump(params if name == 'lab' else {k: [v * 0.5 + 0.25 for v in vals] for k, vals in params.items()}, f)
    ref_norm, our_norm = ref_cls('lab', use_deltas=True), our_cls('lab', use_deltas=True)
    ref_norm.load_params('norm', data_root=str(tmp_path))
    our_norm.load_params('norm', data_root=str(tmp_path))
    x = rng.standard_normal((23, D)).astype(np.float32)
    for deltas in (False, True):
        want = ref_norm.normalise(x, deltas=deltas)
        got = our_norm.normalise(x, deltas=deltas)
        assert isinstance(got, np.ndarray) and got.dtype == want.dtype and np.array_equal(got, want)
        assert np.array_equal(our_norm.denormalise(x, deltas=deltas), ref_norm.denormalise(x, deltas=deltas))
