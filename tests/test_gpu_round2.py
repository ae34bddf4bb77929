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
