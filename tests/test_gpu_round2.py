"""Round-2 GPU tests: the advisor's findings (bounds of the backward kernels, gradients of the segment ops, EMA launch
bookkeeping, collate bounds, in_features check), `sequence_mask` on CUDA lengths, and oracle parity at BASELINE.json sizes."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

REL = 1e-6


@pytest.fixture(scope='module')
def mg():
    import morgana_b200
    return morgana_b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ----------------------------------------------------------------------------------------------------------------------
# advisor findings
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('D', [600, 609, 1000])      # row-streaming form, odd rows, column-strip form
@pytest.mark.parametrize('kind', [None, 'mvn'])
def test_upsample_backward_with_truncating_max_len(mg, D, kind):
    """A caller-supplied max_len below an utterance's frame count truncates the forward; the backward must sum exactly the
    rows that exist (it used to read into the next utterance / past the end of grad_out)."""
    rng = np.random.default_rng(D)
    B, P = 4, 9
    x = rng.random((B, P, D), dtype=np.float32)
    dur = rng.integers(1, 7, (B, P))
    T = int(dur.sum(axis=1).min()) - 3                     # every utterance is cut, the last one included
    mean, std = rng.standard_normal(D).astype(np.float32), (rng.random(D) + 0.2).astype(np.float32)
    xt = dev(x).requires_grad_()
    norm = None if kind is None else ('mvn', dev(mean), dev(std))
    out = mg.utils.upsample_to_repetitions(xt, dev(dur), normaliser=norm, max_len=T)
    assert out.shape == (B, T, D)
    grad_out = rng.standard_normal((B, T, D)).astype(np.float32)
    # poison behind the buffer: an out-of-bounds read of the last utterance would pick this up
    backing = torch.full((B * T * D + 4 * D,), float('nan'), device='cuda')
    backing[:B * T * D] = dev(grad_out).reshape(-1)
    out.backward(backing[:B * T * D].view(B, T, D))
    full_T = int(dur.sum(axis=1).max())
    padded = np.zeros((B, full_T, D), dtype=np.float32)
    padded[:, :T] = grad_out
    want = O.upsample_backward(padded, dur).astype(np.float64)
    if kind is not None:
        want = want / (std + np.float32(1e-8)).astype(np.float64)
    got = xt.grad.cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-6)


def _reference_segment_ops(x, lens, seq_len):
    """The reference's advanced-indexing formulation (utils.py:147-166, 231-330) in stock torch ops, for autograd."""
    B, T, D = x.shape
    S = lens.shape[1]
    padded = torch.cat([x, torch.zeros(B, 1, D, dtype=x.dtype, device=x.device)], dim=1)
    batch = torch.arange(B, device=x.device)
    ends = torch.cumsum(lens, dim=1) * (lens > 0)
    seg_ends = padded[batch[:, None].expand(B, S), ends - 1]
    L = int(lens.max())
    begins = torch.cumsum(lens, dim=1) - lens
    j = torch.arange(L, device=x.device)
    idx = torch.where(j[None, None, :] < lens[:, :, None], begins[:, :, None] + j[None, None, :], torch.full((1,), -1, device=x.device))
    split = padded[batch[:, None, None].expand(B, S, L), idx]
    mask = torch.arange(T, device=x.device)[None, :] < seq_len[:, None]
    select = x[mask.nonzero(as_tuple=True)]
    return select, seg_ends, split


@pytest.mark.parametrize('D', [1, 7, 600])
def test_segment_ops_are_differentiable_like_the_reference(mg, D):
    rng = np.random.default_rng(100 + D)
    B, T, S = 7, 53, 6
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    seq_len = rng.integers(0, T + 1, B)
    lens = rng.integers(0, 11, (B, S))
    lens[rng.random((B, S)) < 0.25] = 0
    while (lens.sum(axis=1) > T).any():
        lens[lens.sum(axis=1) > T] //= 2
    lens[0, 0] = max(int(lens.max()), 1)           # keep the longest segment non-empty
    lens[0, 1:] = 0
    U = mg.utils
    ours_x, ref_x = dev(x).requires_grad_(), dev(x).requires_grad_()
    ours = (U.batched_masked_select(ours_x, dev(seq_len)), U.get_segment_ends(ours_x, dev(lens)[:, :, None]),
            U.split_to_segments(ours_x, dev(lens)[:, :, None]))
    ref = _reference_segment_ops(ref_x, dev(lens), dev(seq_len))
    weights = []
    for a, b in zip(ours, ref):
        assert a.requires_grad and torch.equal(a.detach(), b.detach())
        weights.append(torch.randn(b.shape, device='cuda', generator=torch.Generator(device='cuda').manual_seed(D)))
    for k in range(3):                             # one op at a time, then all three together (gradients accumulate)
        ours_x.grad = ref_x.grad = None
        (ours[k] * weights[k]).sum().backward(retain_graph=True)
        (ref[k] * weights[k]).sum().backward(retain_graph=True)
        assert torch.equal(ours_x.grad, ref_x.grad), k
    ours_x.grad = ref_x.grad = None
    sum((o * w).sum() for o, w in zip(ours, weights)).backward()
    sum((r * w).sum() for r, w in zip(ref, weights)).backward()
    np.testing.assert_allclose(ours_x.grad.cpu().numpy(), ref_x.grad.cpu().numpy(), rtol=1e-6, atol=1e-6)
    with torch.no_grad():                          # inference path: plain tensors out
        assert not U.get_segment_ends(ours_x, dev(lens)).requires_grad


def test_split_to_segments_gradient_with_short_max_segment_len(mg):
    rng = np.random.default_rng(4)
    x = dev(rng.standard_normal((2, 20, 8)).astype(np.float32)).requires_grad_()
    lens = torch.tensor([[5, 0, 9], [3, 3, 3]], device='cuda')
    out = mg.ops.split_to_segments(x, lens, max_segment_len=4)        # segments longer than 4 rows are cut
    out.sum().backward()
    want = torch.zeros(2, 20, 8)
    want[0, 0:4] = 1; want[0, 5:9] = 1
    want[1, 0:9] = 1
    assert torch.equal(x.grad.cpu(), want)


def test_ema_skips_empty_tensors_without_repeating_any(mg):
    """> 64 tensors with empty ones among the first 64: every tensor is updated exactly once."""
    rng = np.random.default_rng(2)
    sizes = [int(s) for s in rng.integers(1, 500, 150)]
    for i in (0, 5, 63, 64, 70, 149):
        sizes[i] = 0
    shadow = [rng.standard_normal(s).astype(np.float32) for s in sizes]
    param = [rng.standard_normal(s).astype(np.float32) for s in sizes]
    pairs = [(dev(s), dev(p)) for s, p in zip(shadow, param)]
    mg.ops.ema_update(pairs, 1.0 - 0.99)
    for (s_dev, _), s, p in zip(pairs, shadow, param):
        assert np.array_equal(s_dev.cpu().numpy(), O.ema_update(s.copy(), p, 0.99))


def test_pad_collate_truncates_lengths_that_overrun_the_packed_rows(mg):
    """The no-sync path (max_len given) cannot validate sum(lengths); the kernel must stay inside `packed`."""
    rows = torch.arange(10 * 4, dtype=torch.float32, device='cuda').reshape(10, 4)
    backing = torch.full((10 * 4 + 64,), float('nan'), device='cuda')
    backing[:40] = rows.reshape(-1)
    packed = backing[:40].view(10, 4)
    lengths = torch.tensor([4, 5, 6], device='cuda')                  # sums to 15 > 10 rows
    out = mg.data.pad_collate(packed, lengths, max_len=6).cpu()
    assert torch.isfinite(out).all()
    assert torch.equal(out[0, :4], rows[0:4].cpu()) and torch.equal(out[1, :5], rows[4:9].cpu())
    assert torch.equal(out[2, :1], rows[9:10].cpu()) and not out[2, 1:].any() and not out[0, 4:].any()
    with pytest.raises(ValueError):
        mg.data.pad_collate(packed, lengths)                          # the checked path still refuses


def test_linear_in_features_mismatch_raises_like_nn_linear(mg):
    x = torch.randn(16, 600, device='cuda').to(torch.bfloat16)
    w = torch.randn(32, 512, device='cuda').to(torch.bfloat16)
    with pytest.raises(RuntimeError):
        mg.ops.linear_bf16(x, w)
    y = mg.ops.linear_bf16(x, torch.randn(32, 600, device='cuda').to(torch.bfloat16))
    assert y.shape == (16, 32)
    xp = torch.zeros(16, 608, device='cuda', dtype=torch.bfloat16)      # explicitly padded operand
    xp[:, :600] = x
    assert torch.equal(mg.ops.linear_bf16(xp, w.new_zeros(32, 600).copy_(torch.randn(32, 600)), in_features=600).isfinite().all(),
                       torch.tensor(True, device='cuda'))


# ----------------------------------------------------------------------------------------------------------------------
# a2 sequence_mask on the device
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dtype', [torch.ByteTensor, torch.cuda.ByteTensor, torch.float32, torch.bool, torch.long])
def test_sequence_mask_on_cuda_lengths(mg, dtype):
    """utils.py:115-144 on device-resident lengths.  The reference's default ``dtype=torch.ByteTensor`` is a *CPU* tensor type:
    ``.type()`` with it brings the mask to the host, and the drop-in does the same (parity, not a preference)."""
    seq_len = torch.tensor([0, 3, 7, 1, 7], device='cuda')
    want_dtype = torch.uint8 if dtype in (torch.ByteTensor, torch.cuda.ByteTensor) else dtype
    for max_len in (None, 7, 10, 2):
        mask = mg.utils.sequence_mask(seq_len, max_len=max_len, dtype=dtype)
        T = 7 if max_len is None else max_len
        assert mask.shape == (5, T, 1) and mask.dtype == want_dtype
        assert mask.is_cuda == (dtype is not torch.ByteTensor)
        want = O.sequence_mask(seq_len.cpu().numpy(), max_len=T)
        assert np.array_equal(mask.cpu().numpy().astype(np.uint8), want.astype(np.uint8))
    # it multiplies into (B, T, D) tensors as the reference's callers use it (losses.py:34-37)
    x = torch.ones(5, 7, 3, device='cuda')
    assert (x * mg.utils.sequence_mask(seq_len, dtype=torch.float32)).sum().item() == 3. * 18


# ----------------------------------------------------------------------------------------------------------------------
# a14 backward: the input gradient of every layer width runs through the tcgen05 kernel (no library GEMM)
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('M,K,N', [(1000, 32, 1), (777, 128, 32), (2049, 512, 128), (300, 256, 187), (513, 64, 3), (4100, 512, 256),
                                   (64, 600, 512)])
def test_cast_transpose_and_short_reduction_dgrad(mg, M, K, N):
    rng = np.random.default_rng(M + K + N)
    w = (rng.standard_normal((N, K)) * 0.1).astype(np.float32)
    wt = mg.ops.cast_transpose_bf16(dev(w))
    n_pad = (N + 7) // 8 * 8
    assert wt.shape == (K, n_pad) and wt.dtype == torch.bfloat16
    want_wt = torch.zeros(K, n_pad, dtype=torch.bfloat16)
    want_wt[:, :N] = torch.from_numpy(w).t().to(torch.bfloat16)
    assert torch.equal(wt.cpu().view(torch.int16), want_wt.view(torch.int16)), 'cast + transpose is not the rounded transpose'

    layer = mg.nn.Linear(K, N, device='cuda')
    with torch.no_grad():
        layer.weight.copy_(dev(w))
    x = dev(rng.standard_normal((M, K)).astype(np.float32)).requires_grad_()
    grad_y = rng.standard_normal((M, N)).astype(np.float32)
    layer(x).backward(dev(grad_y))
    g16 = torch.from_numpy(grad_y).to(torch.bfloat16).float().numpy().astype(np.float64)
    w16 = torch.from_numpy(w).to(torch.bfloat16).float().numpy().astype(np.float64)
    want = g16 @ w16                                             # exact product of the bf16-rounded operands
    got = x.grad.cpu().numpy()
    # bf16 operands, fp32 accumulation over N <= 512 terms: 2e-3 of the result's range (the forward layer's stated tolerance)
    assert np.abs(got - want).max() <= 2e-3 * max(np.abs(want).max(), 1e-6)
    x16 = torch.from_numpy(x.detach().cpu().numpy()).to(torch.bfloat16).float().numpy().astype(np.float64)
    want_w = g16.T @ x16
    assert np.abs(layer.weight.grad.cpu().numpy() - want_w).max() <= 1e-3 * max(np.abs(want_w).max(), 1e-6)


def test_nn_linear_backward_makes_no_library_gemm(mg):
    """The whole backward of a layer is our kernels: K7g (activation gradient), cast + transpose, K7 (input gradient), K7w."""
    layer = mg.nn.Linear(128, 32, act='sigmoid', device='cuda')
    x = torch.randn(4096, 128, device='cuda', requires_grad=True)
    from torch.profiler import profile, ProfilerActivity
    y = layer(x)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        y.sum().backward()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages()]
    offenders = [n for n in names if any(word in n.lower() for word in ('gemm', 'cutlass', 'cublas', 'gemv'))
                 and 'tcgen05' not in n]          # (our kernels' parameter struct is called GemmParams)
    assert not offenders, offenders
    ours = [n for n in names if 'tcgen05' in n or 'act_grad' in n or 'cast_transpose' in n]
    assert any('linear_tcgen05' in n for n in ours) and any('wgrad' in n for n in ours) and any('cast_transpose' in n for n in ours), ours


# ----------------------------------------------------------------------------------------------------------------------
# torch custom operators: CUDA kernel + fake kernel + autograd formula agree (torch.library.opcheck), and a function made of
# them traces under torch.compile (aot_eager: dynamo + AOT autograd through the fake kernels; no code generation)
# ----------------------------------------------------------------------------------------------------------------------
def test_custom_ops_pass_opcheck_and_trace(mg):
    T = torch.ops.morgana_b200
    rng = np.random.default_rng(0)
    x = dev(rng.random((3, 7, 16), dtype=np.float32))
    dur = dev(rng.integers(0, 5, (3, 7, 1)))
    T_len = int(dur.sum(dim=(1, 2)).max())
    p0, p1 = dev(np.zeros(16, np.float32)), dev((rng.random(16) + 0.5).astype(np.float32))
    checks = ('test_schema', 'test_faketensor', 'test_autograd_registration')
    torch.library.opcheck(T.upsample_norm, (x.clone().requires_grad_(), dur, p0, p1, 'minmax', T_len), test_utils=checks)
    pred, tgt = dev(rng.standard_normal((3, T_len, 16)).astype(np.float32)), dev(rng.standard_normal((3, T_len, 16)).astype(np.float32))
    n = dur.sum(dim=(1, 2))
    torch.library.opcheck(T.masked_loss, (pred.clone().requires_grad_(), tgt, n, 'mse'), test_utils=checks)
    torch.library.opcheck(T.normalise, (pred.clone().requires_grad_(), p0, p1, 'mvn', False), test_utils=checks)

    def objective(x, pred):
        frames = T.upsample_norm(x, dur, p0, p1, 'minmax', T_len)
        return T.masked_loss(frames + pred, tgt, n, 'mse')

    xe, pe = x.clone().requires_grad_(), pred.clone().requires_grad_()
    eager = objective(xe, pe)
    eager.backward()
    xc, pc = x.clone().requires_grad_(), pred.clone().requires_grad_()
    compiled = torch.compile(objective, backend='aot_eager', fullgraph=True)(xc, pc)
    compiled.backward()
    assert compiled.item() == eager.item() and torch.equal(xc.grad, xe.grad) and torch.equal(pc.grad, pe.grad)
    # the same numbers as the direct (ctypes) path of the drop-in functions
    xd, pd = x.clone().requires_grad_(), pred.clone().requires_grad_()
    direct = mg.losses.mse(mg.utils.upsample_to_repetitions(xd, dur, normaliser=('minmax', p0, p1), max_len=T_len) + pd, tgt, n)
    direct.backward()
    assert direct.item() == eager.item() and torch.equal(xd.grad, xe.grad) and torch.equal(pd.grad, pe.grad)


# ----------------------------------------------------------------------------------------------------------------------
# oracle parity at BASELINE.json sizes (not self-consistency: the NumPy oracle on the full tensors)
# ----------------------------------------------------------------------------------------------------------------------
def test_config2_full_size_fused_upsample_equals_the_oracle(mg):
    """configs[1]: 256 utterances x ~60 phones, dur U{1..30}, 600-dim labels, seed 1234 -- the 837 MB output, bit for bit."""
    from morgana_b200 import workloads
    ling = workloads.linguistic_batch(batch_size=256, seed=1234)
    lab, dur = ling['lab'].numpy(), ling['dur'].numpy()
    want = O.normalise_upsample(lab, dur, 'minmax', ling['mmin'].numpy(), ling['mmax'].numpy())
    assert want.shape == (256, int(ling['n_frames'].max()), 600) and want.nbytes > 800e6
    for path in ('auto', 'direct'):
        got = mg.ops.upsample(ling['lab'].cuda(), ling['dur'].cuda(), norm=('minmax', ling['mmin'].cuda(), ling['mmax'].cuda()),
                              path=path)
        assert np.array_equal(got.cpu().numpy(), want), path
        del got
    # through the reference's signature + the standalone normaliser: the composition the fused kernel replaces
    norm_lab = mg.data.normalise_minmax(ling['lab'].cuda(), ling['mmin'].cuda(), ling['mmax'].cuda())
    valid_phone = (torch.arange(lab.shape[1])[None, :] < ling['n_phones'][:, None])[:, :, None].cuda()
    got, n_frames = mg.utils.upsample_to_repetitions(norm_lab * valid_phone, ling['dur'].cuda(), return_lengths=True)
    assert np.array_equal(n_frames.cpu().numpy(), ling['n_frames'].numpy())
    assert np.array_equal(got.cpu().numpy(), want)
    # packed wire format (no phone padding on the wire), same bits
    valid = valid_phone[:, :, 0].cpu()
    got = mg.utils.upsample_packed_to_repetitions(ling['lab'][valid].cuda(), ling['dur'][:, :, 0][valid].cuda(), ling['n_phones'].cuda(),
                                                  normaliser=('minmax', ling['mmin'].cuda(), ling['mmax'].cuda()))
    assert np.array_equal(got.cpu().numpy(), want)


def test_config3_full_size_objective_equals_the_oracle(mg):
    """configs[2]: 1024 utterances x U{300..1200} frames x 187 dims: loss, metric sums / counts (all of it) and the gradient
    (every element of the first 128 utterances) of the one-launch objective against the fp64 NumPy oracle; and the drop-in
    losses.mse / metrics.RMSE on the full tensors."""
    from morgana_b200 import workloads
    from morgana_b200.fused import AcousticObjective
    n = workloads.acoustic_lengths(batch_size=1024, seed=1234)
    ac = workloads.acoustic_batch(n, seed=1234)
    p, t, nn_ = ac['pred'].numpy(), ac['target'].numpy(), n.numpy()
    pred, target, n_frames = ac['pred'].cuda(), ac['target'].cuda(), n.cuda()
    objective = AcousticObjective()
    total, grad = objective(pred, target, n_frames)
    want_total = (O.masked_loss(p[..., 0:3], t[..., 0:3], nn_) + O.masked_loss(p[..., 4:184], t[..., 4:184], nn_) +
                  O.masked_loss(p[..., 184:187], t[..., 184:187], nn_) + O.masked_loss(p[..., 3:4], t[..., 3:4], nn_, 'bce')) / 4.
    assert abs(total.item() - want_total) <= REL * abs(want_total), (total.item(), want_total)
    voiced_p = p[..., 3:4] > 0.5
    want_metrics = {'LF0_RMSE_Hz': O.lf0_acc(t[..., 0:1], p[..., 0:1], voiced_p, nn_),
                    'VUV_accuracy': O.mean_acc((ac['voiced'].numpy() == voiced_p).astype(np.float32), nn_),
                    'MCEP_distortion': O.melcep_acc(t[..., 4:64], p[..., 4:64], nn_),
                    'BAP_distortion': O.distortion_acc(t[..., 184:185], p[..., 184:185], nn_)}
    for name, (s, c) in want_metrics.items():
        got = objective.metrics[name]
        assert float(got.count) == c, name
        assert abs(float(got.sum) - s) <= REL * abs(s), (name, float(got.sum), s)
    sub = slice(0, 128)
    g = grad[sub].cpu().numpy()
    for sl, kind in [(slice(0, 3), 'mse'), (slice(4, 184), 'mse'), (slice(184, 187), 'mse'), (slice(3, 4), 'bce')]:
        # the gradient of the mean over ALL 1024 utterances: the oracle on a sub-batch scales by its own batch size
        want = 0.25 * O.masked_loss_grad(p[sub][..., sl], t[sub][..., sl], nn_[sub], kind) * (128. / 1024.)
        np.testing.assert_allclose(g[..., sl], want, rtol=3e-6 if kind == 'bce' else REL, atol=1e-14)
    valid = np.arange(p.shape[1])[None, :] < nn_[:, None]
    assert not grad.cpu().numpy()[~valid].any(), 'gradient of padding frames must be zero'
    # the drop-in pieces on the full tensors
    loss = mg.losses.mse(pred, target, n_frames)
    want = O.masked_loss(p, t, nn_)
    assert abs(loss.item() - want) <= REL * abs(want)
    rmse = mg.metrics.RMSE()
    rmse.reset_state()
    rmse.accumulate(target, pred, seq_len=n_frames)
    s, c = O.rmse_acc(t, p, nn_)
    assert float(rmse.count) == c and abs(float(rmse.sum) - s) <= REL * abs(s)


# ----------------------------------------------------------------------------------------------------------------------
# f1 MLPG: the reference's OWN function (viz/synthesis.py:79-180, mirrored in oracle/_ref, on the bandmat stand-in) as checker
# ----------------------------------------------------------------------------------------------------------------------
MLPG_WINDOWS = {
    'default': None,
    'static+delta': [(0, 0, np.array([1.0])), (1, 1, np.array([-0.5, 0.0, 0.5]))],
    'one-sided': [(0, 0, np.array([1.0])), (1, 0, np.array([-1.0, 1.0])), (0, 1, np.array([-1.0, 1.0]))],
    'four': [(0, 0, np.array([1.0])), (1, 1, np.array([-0.5, 0.0, 0.5])), (1, 1, np.array([1.0, -2.0, 1.0])),
             (1, 1, np.array([0.25, 0.5, 0.25]))],
}


@pytest.mark.parametrize('name', sorted(MLPG_WINDOWS))
@pytest.mark.parametrize('var_kind', ['global', 'frame'])
def test_mlpg_equals_the_reference_function_for_any_three_tap_windows(mg, name, var_kind):
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.fail('the reference mirror oracle/_ref is missing (python oracle/make_ref.py in the build container)')
    ref_loader.import_reference()
    from morgana.viz.synthesis import MLPG as reference_mlpg
    from morgana_b200.viz.synthesis import MLPG
    windows = MLPG_WINDOWS[name]
    W = 3 if windows is None else len(windows)
    rng = np.random.default_rng(len(name))
    B, T, F, pad = 4, 70, 5, 12
    means = rng.standard_normal((B, T, W * F)).astype(np.float32)
    seq_len = np.array([70, 41, 9, 1])
    if var_kind == 'global':
        var = (rng.random(W * F) + 0.3).astype(np.float32)
    else:
        var = (rng.random((B, T, W * F)) + 0.3).astype(np.float32)
    want = reference_mlpg(means.astype(np.float64), var.astype(np.float64), windows=windows, padding_size=pad, seq_len=seq_len)
    got = MLPG(dev(means), dev(var), windows=windows, padding_size=pad, seq_len=dev(seq_len))
    assert got.shape == (B, T, F) and got.dtype == torch.float32
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    # NumPy in -> float64 NumPy out, as the reference's callers expect (models/RNN_SPSS.py:111-116)
    got_np = MLPG(means, var, windows=windows, padding_size=pad, seq_len=seq_len)
    assert isinstance(got_np, np.ndarray) and got_np.dtype == np.float64
    np.testing.assert_allclose(got_np, want, rtol=1e-5, atol=1e-5)


def test_mlpg_rejects_windows_wider_than_the_band(mg):
    from morgana_b200.viz.synthesis import MLPG
    means = torch.randn(1, 20, 2, device='cuda')
    with pytest.raises(NotImplementedError):
        MLPG(means, torch.ones(2, device='cuda'), windows=[(0, 0, np.array([1.0])), (2, 2, np.array([1., -8., 0., 8., -1.]) / 12.)])
