"""Property tests on the GPU path (hypothesis): ragged / empty / odd shapes against the oracle, plus the size-independent
properties the domain offers (round trips, linearity, additivity of metric state, determinism)."""
import os

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu
COMMON = dict(deadline=None, max_examples=40, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@pytest.fixture(scope='module')
def mg():
    import morgana_b200
    return morgana_b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@st.composite
def ragged_batches(draw):
    B = draw(st.integers(1, 6))
    P = draw(st.integers(1, 14))
    D = draw(st.sampled_from([1, 2, 3, 4, 5, 8, 12, 187, 600, 609]))
    max_dur = draw(st.sampled_from([1, 3, 9, 40]))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    dur = rng.integers(0, max_dur + 1, (B, P))
    dur[rng.random((B, P)) < draw(st.sampled_from([0.0, 0.3, 0.9]))] = 0
    x = rng.standard_normal((B, P, D)).astype(np.float32)
    kind = draw(st.sampled_from([None, 'mvn', 'minmax']))
    p0 = rng.standard_normal(D).astype(np.float32)
    p1 = (p0 + np.abs(rng.standard_normal(D)) + 0.05).astype(np.float32) if kind == 'minmax' else \
        (np.abs(rng.standard_normal(D)) + 0.05).astype(np.float32)
    if kind == 'minmax' and D > 1:
        p1[0] = p0[0]
    return x, dur, kind, p0, p1


@settings(**COMMON)
@given(ragged_batches())
def test_upsample_any_ragged_batch_bit_exact(mg, batch):
    x, dur, kind, p0, p1 = batch
    want = O.upsample_to_repetitions(x, dur) if kind is None else O.normalise_upsample(x, dur, kind, p0, p1)
    norm = None if kind is None else (kind, dev(p0), dev(p1))
    got, n_frames = mg.utils.upsample_to_repetitions(dev(x), dev(dur)[:, :, None], normaliser=norm, return_lengths=True)
    assert tuple(got.shape) == want.shape
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(n_frames.cpu().numpy(), dur.sum(axis=1))
    # frame count conservation and zero padding, independent of the oracle
    assert int((got != 0).any(dim=2).sum()) <= int(dur.sum())


@settings(**COMMON)
@given(st.integers(1, 7), st.integers(1, 90), st.sampled_from([1, 3, 5, 60, 187]), st.integers(0, 2 ** 31 - 1),
       st.sampled_from(['mse', 'l1', 'bce']))
def test_losses_any_shape(mg, B, T, D, seed, kind):
    rng = np.random.default_rng(seed)
    seq_len = rng.integers(1, T + 1, B)
    p = (rng.random((B, T, D)) * 0.98 + 0.01).astype(np.float32)
    y = (rng.random((B, T, D)) < 0.5).astype(np.float32) if kind == 'bce' else rng.standard_normal((B, T, D)).astype(np.float32)
    pt = dev(p).requires_grad_()
    value = getattr(mg.losses, kind)(pt, dev(y), dev(seq_len))
    want = O.masked_loss(p, y, seq_len, kind)
    assert abs(value.item() - want) <= 1e-6 * abs(want) + 1e-12
    value.backward()
    np.testing.assert_allclose(pt.grad.cpu().numpy(), O.masked_loss_grad(p, y, seq_len, kind), rtol=3e-6, atol=1e-12)
    assert not pt.grad[torch.arange(T, device='cuda')[None, :] >= dev(seq_len)[:, None]].any()   # padding gradient is exactly 0


def test_normalise_denormalise_round_trip_full_size(mg):
    """Config-3 sized tensor: denormalise(normalise(x)) returns x to fp32 rounding (the reference's 2.4e-7)."""
    g = torch.Generator(device='cuda').manual_seed(0)
    x = torch.randn(1024, 1200, 187, device='cuda', generator=g)
    mean, std = torch.randn(187, device='cuda', generator=g), torch.rand(187, device='cuda', generator=g) + 0.1
    back = mg.data.denormalise_mvn(mg.data.normalise_mvn(x, mean, std), mean, std)
    assert (back - x).abs().max().item() <= 2e-6 * (x.abs().max().item() + 1)
    mmin = torch.randn(187, device='cuda', generator=g)
    mmax = mmin + torch.rand(187, device='cuda', generator=g) + 0.5
    back = mg.data.denormalise_minmax(mg.data.normalise_minmax(x, mmin, mmax), mmin, mmax)
    assert (back - x).abs().max().item() <= 2e-6 * (x.abs().max().item() + 1)


def test_upsample_full_size_properties(mg):
    """Config 2 at full size: layout properties that need no oracle (every utterance's rows are its items repeated in
    order, the tail is zero) plus bit-equality of the bulk and direct paths and run-to-run determinism."""
    from morgana_b200 import workloads
    ling = workloads.linguistic_batch(batch_size=256, seed=99)
    lab, dur = ling['lab'].cuda(), ling['dur'].cuda()
    out, n_frames = mg.utils.upsample_to_repetitions(lab, dur, return_lengths=True)
    assert torch.equal(n_frames.cpu(), ling['n_frames'])
    T = out.shape[1]
    assert T == int(ling['n_frames'].max())
    pad_mask = torch.arange(T, device='cuda')[None, :] >= n_frames[:, None]
    assert not out[pad_mask].any()
    # checksum of checksums: summing frames per utterance == summing items weighted by their durations
    frames_sum = out.double().sum(dim=1)
    items_sum = (lab.double() * dur.double()).sum(dim=1)
    assert torch.allclose(frames_sum, items_sum, rtol=1e-12, atol=1e-9)
    direct = mg.utils.upsample_to_repetitions(lab, dur, path='direct')
    again = mg.utils.upsample_to_repetitions(lab, dur)
    assert torch.equal(out, direct) and torch.equal(out, again)


def test_metric_state_is_additive_over_batches_and_shards(mg):
    """sum / count are additive: one pass over 64 utterances == two passes over 32 == four shards of 16."""
    from morgana_b200 import workloads
    n = workloads.acoustic_lengths(batch_size=64, min_frames=50, max_frames=200, seed=3)
    ac = workloads.acoustic_batch(n, seed=3)
    tgt, pred, nd = ac['target'].cuda(), ac['pred'].cuda(), n.cuda()
    whole = mg.metrics.RMSE()
    whole.reset_state()
    whole.accumulate(tgt, pred, seq_len=nd)
    for parts in (2, 4):
        split = mg.metrics.RMSE()
        split.reset_state()
        step = 64 // parts
        for i in range(parts):
            sl = slice(i * step, (i + 1) * step)
            split.accumulate(tgt[sl], pred[sl], seq_len=nd[sl])
        assert float(split.count) == float(whole.count)
        assert abs(float(split.sum) - float(whole.sum)) <= 1e-6 * float(whole.sum)


def test_ema_converges_to_the_parameters(mg):
    """Idempotence in the limit: repeated EMA updates with fixed parameters converge to them; decay 0 copies in one step."""
    p = [torch.randn(1000, device='cuda'), torch.randn(17, 3, device='cuda')]
    s = [torch.zeros_like(t) for t in p]
    mg.ops.ema_update(list(zip(s, p)), 1.0)          # decay 0  ->  shadow == param exactly
    for a, b in zip(s, p):
        assert torch.equal(a, b)
    s = [torch.zeros_like(t) for t in p]
    for _ in range(200):
        mg.ops.ema_update(list(zip(s, p)), 0.1)
    for a, b in zip(s, p):
        assert (a - b).abs().max().item() <= 1e-6 * b.abs().max().item() + 1e-6


def test_objective_full_size_properties(mg):
    """Config 3 at full size (1024 utterances, 300..1200 frames, 187 dims; ~0.9 GB per tensor), through properties that
    need no oracle: the fused objective equals the four separate losses, the gradient obeys sum(grad * (p - y)) = 2 * mse
    on the squared-error columns and is exactly zero in the padding, metric state is additive over shards, and two runs
    are bit-identical (deterministic reductions)."""
    from morgana_b200 import workloads
    from morgana_b200.fused import AcousticObjective
    B, T, D = 1024, 1200, 187
    n = workloads.acoustic_lengths(batch_size=B, min_frames=300, max_frames=T, seed=1234).cuda()
    g = torch.Generator(device='cuda').manual_seed(1234)
    target = torch.randn(B, T, D, generator=g, device='cuda')
    pred = target + 0.1 * torch.randn(B, T, D, generator=g, device='cuda')
    target[:, :, 0] = 5. + 0.3 * target[:, :, 0]
    pred[:, :, 0] = target[:, :, 0] + 0.05 * pred[:, :, 0]
    target[:, :, 3] = (torch.rand(B, T, generator=g, device='cuda') < 0.6).float()
    pred[:, :, 3] = torch.sigmoid(torch.randn(B, T, generator=g, device='cuda'))

    objective = AcousticObjective()
    total, grad = objective(pred, target, n)
    parts = (mg.losses.mse(pred[..., 0:3], target[..., 0:3], n) + mg.losses.mse(pred[..., 4:184], target[..., 4:184], n) +
             mg.losses.mse(pred[..., 184:187], target[..., 184:187], n) + mg.losses.bce(pred[..., 3:4], target[..., 3:4], n)) / 4.
    assert abs(total.item() - parts.item()) <= 1e-6 * abs(parts.item())

    # d/dp of mean_b mean_d sum_t (p - y)^2 / n_b, weighted by (p - y), gives back 2 x that loss (x 1/4 for the total)
    sq_cols = [c for c in range(D) if c != 3]
    inner = (grad[..., sq_cols].double() * (pred[..., sq_cols] - target[..., sq_cols]).double()).sum().item()
    three_mse = 4. * parts.item() - mg.losses.bce(pred[..., 3:4], target[..., 3:4], n).item()
    assert abs(inner - 2. * three_mse / 4.) <= 2e-6 * abs(three_mse)
    padding = torch.arange(T, device='cuda')[None, :] >= n[:, None]
    assert not grad[padding].any()

    whole = dict(objective.metrics)
    sums = {k: (float(m.sum), float(m.count)) for k, m in whole.items()}
    sharded = AcousticObjective()
    for i in range(4):
        sl = slice(i * 256, (i + 1) * 256)
        sharded(pred[sl], target[sl], n[sl], want_grad=False)
    for k, m in sharded.metrics.items():
        assert float(m.count) == sums[k][1], k
        assert abs(float(m.sum) - sums[k][0]) <= 1e-6 * abs(sums[k][0]), k

    again = AcousticObjective()
    total2, grad2 = again(pred, target, n)
    assert total2.item() == total.item() and torch.equal(grad, grad2)


_WALK_SEEDS = [int(v) for v in os.environ.get('MG_WALK_SEEDS', '2026,7,99').split(',')]   # more seeds for a soak run


@pytest.mark.parametrize('seed', _WALK_SEEDS)
def test_interleaved_calls_with_changing_shapes(mg, seed):
    """Every op keeps some state between calls (per-stream workspaces, pointer tables, cached attributes, metric records):
    a seeded random walk over ops and shapes, each result checked against the oracle, looks for state that leaks from one
    geometry into the next."""
    from oracle import np_oracle as O
    from morgana_b200 import ops
    from morgana_b200.fused import AcousticObjective
    from morgana_b200.viz.synthesis import MLPG
    rng = np.random.default_rng(seed)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()   # noqa: E731
    rel = lambda got, want: abs(float(got) - float(want)) / max(abs(float(want)), 1e-30)   # noqa: E731
    rmse, objective = mg.metrics.RMSE(), AcousticObjective()
    rmse.reset_state()
    rmse_sum = rmse_count = 0.
    obj_sum = obj_count = 0.
    for step in range(60):
        op = rng.integers(0, 15)
        B, T = int(rng.integers(1, 70)), int(rng.integers(1, 130))
        n = rng.integers(1, T + 1, B)
        if op == 0:      # K1 + K2, fused normalisation
            P, D = int(rng.integers(1, 40)), int(rng.choice([1, 5, 64, 187, 600]))
            x, dur = rng.random((B, P, D), dtype=np.float32), rng.integers(0, 7, (B, P))
            lo, hi = rng.standard_normal(D).astype(np.float32), (rng.standard_normal(D) + 3).astype(np.float32)
            got = mg.utils.upsample_to_repetitions(dev(x), dev(dur), normaliser=('minmax', dev(lo), dev(hi)))
            assert np.array_equal(got.cpu().numpy(), O.normalise_upsample(x, dur, 'minmax', lo, hi)), step
        elif op == 1:    # K4: a loss with its gradient
            D = int(rng.choice([1, 3, 60, 187]))
            kind = str(rng.choice(['mse', 'l1']))
            p = dev(rng.standard_normal((B, T, D)).astype(np.float32)).requires_grad_()
            y = rng.standard_normal((B, T, D)).astype(np.float32)
            value = getattr(mg.losses, kind)(p, dev(y), dev(n))
            grad, = torch.autograd.grad(value, p)
            assert rel(value.item(), O.masked_loss(p.detach().cpu().numpy(), y, n, kind)) <= 1e-6, step
            np.testing.assert_allclose(grad.cpu().numpy(), O.masked_loss_grad(p.detach().cpu().numpy(), y, n, kind), rtol=3e-6, atol=1e-12)
        elif op == 2:    # K5: a metric whose state spans the whole walk
            D = int(rng.choice([1, 7, 60]))
            a, b = rng.standard_normal((B, T, D)).astype(np.float32), rng.standard_normal((B, T, D)).astype(np.float32)
            rmse.accumulate(dev(a), dev(b), seq_len=dev(n))
            s, c = O.rmse_acc(a, b, n)
            rmse_sum, rmse_count = rmse_sum + s, rmse_count + c
            assert float(rmse.count) == rmse_count and rel(rmse.sum, rmse_sum) <= 2e-6, step
        elif op == 3:    # K4b
            y = rng.standard_normal((B, T, 187)).astype(np.float32)
            p = (y + 0.1 * rng.standard_normal((B, T, 187))).astype(np.float32)
            y[:, :, 3] = rng.random((B, T)) < 0.6
            p[:, :, 3] = 1. / (1. + np.exp(-rng.standard_normal((B, T))))
            total, grad = objective(dev(p), dev(y), dev(n))
            want = (O.masked_loss(p[..., 0:3], y[..., 0:3], n) + O.masked_loss(p[..., 4:184], y[..., 4:184], n) +
                    O.masked_loss(p[..., 184:187], y[..., 184:187], n) + O.masked_loss(p[..., 3:4], y[..., 3:4], n, 'bce')) / 4.
            assert rel(total.item(), want) <= 1e-6, step
            s, c = O.melcep_acc(y[..., 4:64], p[..., 4:64], n)
            obj_sum, obj_count = obj_sum + s, obj_count + c
            got = objective.metrics['MCEP_distortion']
            assert float(got.count) == obj_count and rel(got.sum, obj_sum) <= 2e-6, step
        elif op == 4:    # K6
            shapes = [tuple(int(v) for v in rng.integers(1, 40, int(rng.integers(1, 3)))) for _ in range(int(rng.integers(1, 6)))]
            shadow = [rng.standard_normal(s).astype(np.float32) for s in shapes]
            param = [rng.standard_normal(s).astype(np.float32) for s in shapes]
            s_dev, p_dev = [dev(s) for s in shadow], [dev(p) for p in param]
            ops.ema_update(list(zip(s_dev, p_dev)), 1. - 0.99)
            for got, s, p in zip(s_dev, shadow, param):
                assert np.array_equal(got.cpu().numpy(), O.ema_update(s.copy(), p, 0.99)), step
        elif op == 5:    # K7, all tile shapes, single-CTA and CTA-pair forms
            os.environ['MG_GEMM_PAIR'] = str(int(rng.integers(0, 2)))
            os.environ['MG_GEMM_WIDE'] = str(int(rng.integers(0, 2)))
            M, K, N = int(rng.integers(1, 700)), int(rng.choice([8, 40, 64, 600])), int(rng.choice([1, 3, 32, 187, 300, 512]))
            x = dev(rng.standard_normal((M, K)).astype(np.float32)).to(torch.bfloat16)
            w = dev((rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)).to(torch.bfloat16)
            bias = rng.standard_normal(N).astype(np.float32)
            act = [None, 'sigmoid'][int(rng.integers(0, 2))]
            got = ops.linear_bf16(x, w, dev(bias), act=act)
            want = O.linear(x.float().cpu().numpy(), w.float().cpu().numpy(), bias, act)
            np.testing.assert_allclose(got.cpu().numpy(), want, rtol=2e-3, atol=2e-3)
            del os.environ['MG_GEMM_PAIR'], os.environ['MG_GEMM_WIDE']
        elif op == 6:    # K8
            F, pad = int(rng.choice([1, 2, 5])), int(rng.choice([0, 3, 100]))
            means = rng.standard_normal((B, T, 3 * F)).astype(np.float32)
            var = (rng.random(3 * F) + 0.3).astype(np.float32)
            got = MLPG(dev(means), dev(var), padding_size=pad, seq_len=dev(n))
            np.testing.assert_allclose(got.cpu().numpy(), O.mlpg_banded(means, var, padding_size=pad, seq_len=n), rtol=1e-5, atol=1e-5)
        elif op == 7:    # K0 + the row packer: a round trip
            D = int(rng.choice([1, 9, 187]))
            x = rng.standard_normal((B, T, D)).astype(np.float32)
            packed = mg.utils.batched_masked_select(dev(x), dev(n))
            assert np.array_equal(packed.cpu().numpy(), O.batched_masked_select(x, n)), step
            padded = mg.data.pad_collate(packed, dev(n), max_len=T)
            assert np.array_equal(padded.cpu().numpy(), x * (np.arange(T)[None, :] < n[:, None])[:, :, None]), step
        elif op == 8:    # weighted / exp metric, Variance, integer metric
            lf0_t = (5 + 0.3 * rng.standard_normal((B, T, 1))).astype(np.float32)
            lf0_p = (lf0_t + 0.05 * rng.standard_normal((B, T, 1))).astype(np.float32)
            voiced = rng.random((B, T, 1)) < 0.6
            m = mg.metrics.LF0Distortion()
            m.reset_state()
            m.accumulate(dev(lf0_t), dev(lf0_p), dev(voiced), seq_len=dev(n))
            s_, c_ = O.lf0_acc(lf0_t, lf0_p, voiced, n)
            assert float(m.count) == c_ and (c_ == 0 or rel(m.sum, s_) <= 2e-6), step
            var = mg.metrics.Variance()
            var.accumulate(dev(lf0_t), seq_len=dev(n))
            s_, q_, c_ = O.variance_acc(lf0_t, n)
            assert float(var.count) == c_ and rel(var.sum, s_) <= 2e-6 and rel(var.sum_square, q_) <= 2e-6, step
            err = mg.metrics.Error()
            bits = rng.random((B, T, 1)) < 0.5
            err.accumulate(dev(voiced), dev(bits), seq_len=dev(n))
            s_, c_ = O.error_acc(voiced, bits, n)
            assert int(err.sum) == int(s_) and float(err.count) == c_, step
        elif op == 9:    # cross-entropy and bce with their gradients
            C = int(rng.choice([2, 5, 40]))
            logits = dev((2 * rng.standard_normal((B, T, C))).astype(np.float32)).requires_grad_()
            classes = rng.integers(0, C, (B, T))
            value = mg.losses.ce(logits, dev(classes), dev(n))
            want, want_grad = O.cross_entropy_loss(logits.detach().cpu().numpy(), classes, n)
            grad, = torch.autograd.grad(value, logits)
            assert rel(value.item(), want) <= 2e-6, step
            np.testing.assert_allclose(grad.cpu().numpy(), want_grad, rtol=3e-6, atol=3e-8)
            prob = (1. / (1. + np.exp(-rng.standard_normal((B, T, 1))))).astype(np.float32)
            label = (rng.random((B, T, 1)) < 0.5).astype(np.float32)
            assert rel(mg.losses.bce(dev(prob), dev(label), dev(n)).item(), O.masked_loss(prob, label, n, 'bce')) <= 2e-6, step
        elif op == 10:   # expansion: other dtypes (byte path), bf16 output, and the backward
            P, D = int(rng.integers(1, 30)), int(rng.choice([1, 8, 64, 600]))
            dur = rng.integers(0, 6, (B, P))
            xi = rng.integers(-1000, 1000, (B, P, D))
            assert np.array_equal(mg.utils.upsample_to_repetitions(dev(xi), dev(dur)).cpu().numpy(), O.upsample_to_repetitions(xi, dur)), step
            xf = rng.random((B, P, D), dtype=np.float32)
            xg = dev(xf).requires_grad_()
            out = mg.utils.upsample_to_repetitions(xg, dev(dur))
            up = rng.standard_normal(tuple(out.shape)).astype(np.float32)
            out.backward(dev(up))
            np.testing.assert_allclose(xg.grad.cpu().numpy(), O.upsample_backward(up, dur), rtol=1e-5, atol=1e-5)
            if D % 8 == 0:
                got16 = mg.utils.upsample_to_repetitions(dev(xf), dev(dur), out_dtype=torch.bfloat16)
                want16 = torch.from_numpy(O.upsample_to_repetitions(xf, dur)).to(torch.bfloat16)
                assert torch.equal(got16.cpu().view(torch.int16), want16.view(torch.int16)), step
        elif op == 11:   # segment ops
            S, D = int(rng.integers(1, 9)), int(rng.choice([1, 7, 64]))
            x = rng.standard_normal((B, T, D)).astype(np.float32)
            lens = rng.integers(0, max(1, T // S) + 1, (B, S))
            assert np.array_equal(mg.utils.get_segment_ends(dev(x), dev(lens)[:, :, None]).cpu().numpy(), O.get_segment_ends(x, lens)), step
            assert np.array_equal(mg.utils.split_to_segments(dev(x), dev(lens)[:, :, None]).cpu().numpy(), O.split_to_segments(x, lens)), step
        elif op == 12:   # per-utterance (speaker-dependent) parameters, standalone and fused
            P, D = int(rng.integers(1, 20)), int(rng.choice([4, 187, 600]))
            x, dur = rng.random((B, P, D), dtype=np.float32), rng.integers(0, 5, (B, P))
            lo, hi = rng.standard_normal((B, D)).astype(np.float32), (rng.standard_normal((B, D)) + 3).astype(np.float32)
            normed = mg.data.normalise_minmax(dev(x), dev(lo), dev(hi))
            assert np.array_equal(normed.cpu().numpy(), O.normalise_minmax(x, lo, hi)), step
            fused = mg.utils.upsample_to_repetitions(dev(x), dev(dur), normaliser=('minmax', dev(lo), dev(hi)))
            assert torch.equal(fused, mg.utils.upsample_to_repetitions(normed, dev(dur))), step
        elif op == 13:   # K7 backward: activation gradient + bias gradient (K7g), weight gradient as single CTAs / pairs (K7w)
            os.environ['MG_WGRAD_PAIR'] = str(int(rng.integers(0, 2)))
            M, K, N = int(rng.integers(1, 3000)), int(rng.choice([8, 40, 64, 256, 600])), int(rng.choice([1, 3, 32, 187, 300, 512]))
            grad_y = rng.standard_normal((M, N)).astype(np.float32)
            y = rng.random((M, N), dtype=np.float32) if rng.integers(0, 2) else None
            x16 = ops.cast_pad_bf16(dev(rng.random((M, K), dtype=np.float32)))
            g16, bias_grad = ops.act_grad_bf16(dev(grad_y), None if y is None else dev(y))
            g32 = grad_y if y is None else O.sigmoid_grad(grad_y, y)
            assert torch.equal(g16[:, :N].cpu().view(torch.int16), torch.from_numpy(g32).to(torch.bfloat16).view(torch.int16)), step
            np.testing.assert_allclose(bias_grad.cpu().numpy(), g32.astype(np.float64).sum(0), rtol=1e-6, atol=1e-6 * np.abs(g32).sum(0).max())
            got = ops.linear_wgrad_bf16(g16, x16, out_features=N, in_features=K)
            want = O.linear_wgrad(g16[:, :N].float().cpu().numpy(), x16[:, :K].float().cpu().numpy())
            np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-4, atol=1e-4 * max(1.0, np.abs(want).max()))
            del os.environ['MG_WGRAD_PAIR']
        else:            # K3
            D = int(rng.choice([1, 9, 187, 600]))
            x = rng.standard_normal((B, T, D)).astype(np.float32)
            mean, std = rng.standard_normal(D).astype(np.float32), (rng.random(D) + 0.1).astype(np.float32)
            assert np.array_equal(mg.data.normalise_mvn(dev(x), dev(mean), dev(std)).cpu().numpy(), O.normalise_mvn(x, mean, std)), step
            assert np.array_equal(mg.data.denormalise_mvn(dev(x), dev(mean), dev(std)).cpu().numpy(), O.denormalise_mvn(x, mean, std)), step


@pytest.mark.parametrize('pair', ['0', '1'])
def test_linear_backward_full_size_is_exact_on_integers(mg, monkeypatch, pair):
    """K7g / K7w at the config-2 frame count (348,928 frames, 600 -> 512): with one-hot gradient rows and small integer
    features every product and every partial sum is an integer below 2**24, so bf16 operands, fp32 accumulation in tensor
    memory and the split reduction must reproduce the integer result exactly -- any dropped, duplicated or misplaced frame,
    feature atom or slice shows up as a wrong integer."""
    from morgana_b200 import ops
    monkeypatch.setenv('MG_WGRAD_PAIR', pair)
    M, N, K = 256 * 1363, 512, 600
    g = torch.Generator(device='cuda').manual_seed(5)
    x_int = torch.randint(0, 256, (M, K), generator=g, device='cuda')
    hot = torch.randint(0, N, (M,), generator=g, device='cuda')
    grad_y = torch.zeros((M, N), device='cuda')
    grad_y[torch.arange(M, device='cuda'), hot] = 3.
    g16, bias_grad = ops.act_grad_bf16(grad_y, None)
    assert torch.equal(g16.float(), grad_y)
    assert torch.equal(bias_grad, 3. * torch.bincount(hot, minlength=N).float())
    got = ops.linear_wgrad_bf16(g16, x_int.to(torch.bfloat16), out_features=N, in_features=K)
    want = torch.zeros((N, K), dtype=torch.int64, device='cuda').index_add_(0, hot, 3 * x_int)
    assert int(want.max()) < 2 ** 24
    assert torch.equal(got.to(torch.int64), want) and torch.equal(got, want.float())


def test_path_replays_from_a_cuda_graph_on_a_side_stream(mg):
    """The kernels take the caller's stream, never synchronise and never allocate: the whole path (scan, expansion with a
    length hint, fused objective with its gradient, a streaming metric, the EMA update) captures into one CUDA graph on a
    side stream and replays on new inputs with the results of the eager calls."""
    from morgana_b200 import ops, workloads
    from morgana_b200.fused import AcousticObjective
    batches = []
    for seed in (11, 12, 13):
        ling = workloads.linguistic_batch(batch_size=12, min_phones=8, max_phones=16, max_dur=9, seed=seed)
        batches.append((ling, workloads.acoustic_batch(ling['n_frames'], max_len=150, seed=seed)))
    P = max(b[0]['lab'].shape[1] for b in batches)
    pad_items = lambda t: torch.nn.functional.pad(t, (0, 0, 0, P - t.shape[1]))   # noqa: E731
    static = {'lab': torch.empty(12, P, 600, device='cuda'), 'dur': torch.empty(12, P, 1, dtype=torch.int64, device='cuda'),
              'pred': torch.empty(12, 150, 187, device='cuda'), 'target': torch.empty(12, 150, 187, device='cuda')}
    mmin, mmax = batches[0][0]['mmin'].cuda(), batches[0][0]['mmax'].cuda()
    shadow, param = torch.zeros(5000, device='cuda'), torch.ones(5000, device='cuda')

    def load(ling, ac):
        static['lab'].copy_(pad_items(ling['lab']))
        static['dur'].copy_(pad_items(ling['dur']))
        static['pred'].copy_(ac['pred'])
        static['target'].copy_(ac['target'])

    def run(objective, rmse):
        frames, n_frames = mg.utils.upsample_to_repetitions(static['lab'], static['dur'], normaliser=('minmax', mmin, mmax),
                                                            max_len=150, return_lengths=True)
        total, grad = objective(static['pred'], static['target'], n_frames)
        rmse.accumulate(static['target'], static['pred'], seq_len=n_frames)
        ops.ema_update([(shadow, param)], 0.5)
        return frames, total, grad

    # eager reference on the default stream
    eager_obj, eager_rmse = AcousticObjective(), mg.metrics.RMSE()
    eager_rmse.reset_state()
    eager = []
    for ling, ac in batches:
        load(ling, ac)
        frames, total, grad = run(eager_obj, eager_rmse)
        eager.append((frames.clone(), total.clone(), grad.clone()))
    eager_shadow, eager_sum = shadow.clone(), float(eager_rmse.sum)
    shadow.zero_()

    side = torch.cuda.Stream()
    obj, rmse = AcousticObjective(), mg.metrics.RMSE()
    rmse.reset_state()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        load(*batches[0])
        run(obj, rmse)                       # warm-up on the side stream: workspaces and records exist before capture
        shadow.zero_()
        rmse.reset_state()
        rmse.accumulate(static['target'], static['pred'], seq_len=torch.zeros(12, dtype=torch.int64, device='cuda'))   # bind the record
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        g_frames, g_total, g_grad = run(obj, rmse)
    for (ling, ac), (frames, total, grad) in zip(batches, eager):
        load(ling, ac)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(g_frames, frames) and g_total.item() == total.item() and torch.equal(g_grad, grad)
    assert torch.equal(shadow, eager_shadow) and float(rmse.sum) == eager_sum


def test_back_to_back_launches_overlap_without_racing(mg):
    """The kernels of the path are launched with programmatic dependent launch: the CTAs of a launch may be resident while the launch
    before it still runs.  300 launches of the step's kernels back to back, alternating between two batches of different size (so
    grids, partitions and the shared reduction workspace change from launch to launch), no synchronisation in between: every result
    must equal, bit for bit, what the same call returns when it runs alone."""
    from morgana_b200 import workloads
    from morgana_b200.fused import AcousticObjective
    batches = []
    for B, seed in ((48, 5), (17, 6)):
        ling = workloads.linguistic_batch(batch_size=B, min_phones=20, max_phones=40, max_dur=20, seed=seed)
        ac = workloads.acoustic_batch(ling['n_frames'], seed=seed)
        batches.append({'lab': ling['lab'].cuda(), 'dur': ling['dur'].cuda(), 'mm': (ling['mmin'].cuda(), ling['mmax'].cuda()),
                        'T': int(ling['n_frames'].max()), 'pred': ac['pred'].cuda(), 'target': ac['target'].cuda()})

    def step(b, objective):
        out, n = mg.utils.upsample_to_repetitions(b['lab'], b['dur'], normaliser=('minmax',) + b['mm'], max_len=b['T'], return_lengths=True)
        loss, grad = objective(b['pred'], b['target'], n)
        return out, loss, grad

    alone = []
    for b in batches:
        torch.cuda.synchronize()
        out, loss, grad = step(b, AcousticObjective())
        torch.cuda.synchronize()
        alone.append((out.clone(), loss.clone(), grad.clone()))
    objectives = [AcousticObjective(), AcousticObjective()]
    kept = []
    for i in range(300):
        k = i % 2 if i % 7 else 1 - i % 2            # mostly alternating, sometimes the same batch twice in a row
        out, loss, grad = step(batches[k], objectives[k])
        if i % 25 == 0 or i >= 296:
            kept.append((k, out, loss, grad))
    torch.cuda.synchronize()
    for k, out, loss, grad in kept:
        assert torch.equal(out, alone[k][0])
        assert torch.equal(loss, alone[k][1]), (float(loss), float(alone[k][1]))
        assert torch.equal(grad, alone[k][2])
