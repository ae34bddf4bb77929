"""CPU-side checks: the C-ABI library loads and exports what the header declares; host logic; no CPU path."""
import ctypes
import os
import re
import sys
import types

import numpy as np
import pytest
import torch

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def mg():
    sys.path.insert(0, REPO_ROOT)
    import __graft_entry__
    __graft_entry__.build()
    import morgana_b200
    return morgana_b200


def header_symbols():
    with open(os.path.join(REPO_ROOT, 'include', 'morgana_b200.h')) as f:
        text = re.sub(r'/\*.*?\*/', '', f.read(), flags=re.S)
    return sorted(set(re.findall(r'\b(mg_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol(mg):
    from morgana_b200 import _lib
    names = header_symbols()
    assert len(names) >= 13
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), 'libmorgana_b200.so does not export ' + name
    assert sorted(_lib.PROTOTYPES) == names       # the ctypes table binds exactly the header's entry points
    assert raw.mg_abi_version() == 1


def test_struct_layouts_match_header(mg):
    from morgana_b200 import _lib
    assert ctypes.sizeof(_lib.Term) == 144
    assert ctypes.sizeof(_lib.TermResult) == 48
    assert _lib.TermResult.sum_f32.offset == 32 and _lib.TermResult.weighted_loss_f32.offset == 44


def test_sass_uses_bulk_copy_engine(mg):
    """The upsample kernel's stores go through the TMA engine: SASS shows UBLKCP (B200_PROFILING.md)."""
    from morgana_b200 import _lib
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([cuobjdump, '-sass', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in sass
    assert 'UBLKCP' in sass
    # the dense layers run on the 5th-generation tensor cores, forward and weight gradient, single CTAs and CTA pairs
    for mnemonic in ('UTCHMMA', 'UTCHMMA.2CTA', 'UTMALDG.2D', 'LDTM'):
        assert mnemonic in sass, mnemonic


def test_no_cpu_path(mg):
    x = torch.zeros(2, 3, 4)
    with pytest.raises(RuntimeError, match='no CPU path'):
        mg.utils.upsample_to_repetitions(x, torch.ones(2, 3, 1, dtype=torch.long))
    with pytest.raises(RuntimeError, match='no CPU path'):
        mg.losses.mse(x, x)
    with pytest.raises(RuntimeError, match='no CPU path'):
        mg.data.normalise_mvn(x, torch.zeros(4), torch.ones(4))
    with pytest.raises(RuntimeError, match='no CPU path'):
        mg.metrics.RMSE().accumulate(x, x)
    with pytest.raises(RuntimeError, match='no CPU path'):
        mg.ops.ema_update([(torch.zeros(3), torch.ones(3))], 0.1)


def test_numpy_normalisers_stay_numpy(mg):
    """DataLoader workers normalise NumPy arrays (reference data.py:119-127); that path must not touch CUDA."""
    from oracle import np_oracle as O
    rng = np.random.default_rng(0)
    x = rng.standard_normal((7, 5)).astype(np.float32)
    mean, std = rng.standard_normal(5).astype(np.float32), (rng.random(5) + 0.1).astype(np.float32)
    norm = mg.data.MeanVarianceNormaliser('lf0').set_params({'mean': mean, 'std_dev': std}, device='cpu')
    np.testing.assert_allclose(norm.normalise(x), O.normalise_mvn(x, mean, std), rtol=1e-6)
    np.testing.assert_allclose(norm.denormalise(x), O.denormalise_mvn(x, mean, std), rtol=1e-6)
    mmin = rng.standard_normal(5).astype(np.float32)
    mmax = mmin.copy()
    mmax[1:] += 1.
    mm = mg.data.MinMaxNormaliser('lab').set_params({'mmin': mmin, 'mmax': mmax}, device='cpu')
    np.testing.assert_allclose(mm.normalise(x), O.normalise_minmax(x, mmin, mmax), rtol=1e-6)
    assert mm.fused_params()[0] == 'minmax'


def test_normaliser_json_loading(mg, tmp_path):
    import json
    (tmp_path / 'norm').mkdir()
    with open(tmp_path / 'norm' / 'lf0_mvn.json', 'w') as f:
        json.dump({'mean': [1., 2.], 'std_dev': [3., 4.]}, f)
    with open(tmp_path / 'norm' / 'lf0_deltas_mvn.json', 'w') as f:
        json.dump({'mean': [0., 0.], 'std_dev': [1., 1.]}, f)
    norm = mg.data.MeanVarianceNormaliser('lf0', use_deltas=True)
    norm.load_params('norm', data_root=str(tmp_path), device='cpu')
    assert norm.params['std_dev'].dtype == np.float32 and norm.delta_params_torch['mean'].shape == (2,)


def test_patch_and_unpatch_rebind_reference_names(mg):
    """patch() swaps exactly the hot-path callables of a `morgana`-shaped package and unpatch() restores them."""
    fake = types.ModuleType('morgana')
    for sub in ('utils', 'losses', 'data', 'metrics'):
        setattr(fake, sub, types.ModuleType('morgana.' + sub))
    sentinel = object()
    fake.utils.upsample_to_repetitions = fake.utils.ExponentialMovingAverage = sentinel
    fake.losses.mse = fake.losses.bce = sentinel
    for name in ('normalise_mvn', 'denormalise_mvn', 'normalise_minmax', 'denormalise_minmax'):
        setattr(fake.data, name, sentinel)

    def ref_accumulate(self, *a, **k):
        return 'reference'
    for cls in mg.metrics.ACCUMULATORS:
        setattr(fake.metrics, cls.__name__, type(cls.__name__, (object,), {'accumulate': ref_accumulate}))
    mg.patch(fake)
    assert fake.utils.upsample_to_repetitions is mg.utils.upsample_to_repetitions
    assert fake.losses.mse is mg.losses.mse and fake.data.normalise_minmax is mg.data.normalise_minmax
    assert fake.metrics.RMSE.accumulate is mg.metrics.RMSE.__dict__['accumulate']
    mg.unpatch()
    assert fake.utils.upsample_to_repetitions is sentinel and fake.losses.bce is sentinel
    assert fake.metrics.RMSE.accumulate is ref_accumulate
    assert not hasattr(fake.metrics.RMSE, 'result')


def _sd_normalisers(mg, g, device):
    speakers = [str(s) for s in g['sdn_speakers']]
    mvn = mg.data.SpeakerDependentMeanVarianceNormaliser('lf0', 'speakers.scp', use_deltas=True).set_params(
        {s: {k: g['sdn_mvn_%s_%s' % (s, k)] for k in ('mean', 'std_dev')} for s in speakers},
        {s: {k: g['sdn_mvn_deltas_%s_%s' % (s, k)] for k in ('mean', 'std_dev')} for s in speakers}, device=device)
    mm = mg.data.SpeakerDependentMinMaxNormaliser('lab', 'speakers.scp').set_params(
        {s: {k: g['sdn_minmax_%s_%s' % (s, k)] for k in ('mmin', 'mmax')} for s in speakers}, device=device)
    return mvn, mm


def test_speaker_dependent_normalisers_numpy_path_and_loading(mg, golden, tmp_path):
    """The DataLoader-worker path (NumPy, one utterance, one speaker; data.py:388-531) and the per-speaker JSON layout."""
    import json
    g = golden('normalise_sd')
    mvn, mm = _sd_normalisers(mg, g, 'cpu')
    got = mm.normalise(g['sdn_x'][2], 'spk_b')
    assert isinstance(got, np.ndarray) and np.array_equal(got, g['sdn_minmax_norm_numpy'])
    params = mvn.fetch_params(['spk_b', 'spk_a'], np.ndarray, deltas=True)
    assert params['mean'].shape == (2, 7) and np.array_equal(params['mean'][1], g['sdn_mvn_deltas_spk_a_mean'])
    assert mvn.fetch_params('spk_c')['std_dev'].shape == (7,)                   # one speaker: squeezed (data.py:500-501)
    # load_params: {data_dir}/{speaker_id}/{name}_mvn.json (+ _deltas), speakers from the id list
    (tmp_path / 'speakers.scp').write_text('spk_a\nspk_c\n')
    for spk in ('spk_a', 'spk_c'):
        d = tmp_path / 'train' / spk
        d.mkdir(parents=True)
        for suffix, key in (('lf0_mvn.json', 'sdn_mvn_%s_%s'), ('lf0_deltas_mvn.json', 'sdn_mvn_deltas_%s_%s')):
            (d / suffix).write_text(json.dumps({k: g[key % (spk, k)].tolist() for k in ('mean', 'std_dev')}))
    loaded = mg.data.SpeakerDependentMeanVarianceNormaliser('lf0', 'speakers.scp', use_deltas=True)
    loaded.load_params('train', data_root=str(tmp_path), device='cpu')
    assert loaded.speaker_ids == ['spk_a', 'spk_c']
    assert np.array_equal(loaded.delta_params['spk_c']['std_dev'], g['sdn_mvn_deltas_spk_c_std_dev'])
    assert np.array_equal(loaded.normalise(g['sdn_x'][1], 'spk_c'), g['sdn_mvn_norm_single'])


def test_normalisers_container_and_cpu_feeder(mg, tmp_path):
    """data.Normalisers loads every source's JSON (data.py:227-249); ToDeviceWrapper passes CPU batches through."""
    import json
    (tmp_path / 'stats').mkdir()
    (tmp_path / 'stats' / 'lab_minmax.json').write_text(json.dumps({'mmin': [0., 1.], 'mmax': [2., 1.]}))
    (tmp_path / 'stats' / 'dur_mvn.json').write_text(json.dumps({'mean': [3.], 'std_dev': [2.]}))
    norms = mg.data.Normalisers({'lab': mg.data.MinMaxNormaliser('lab'), 'dur': mg.data.MeanVarianceNormaliser('dur')},
                                'stats', data_root=str(tmp_path), device='cpu')
    assert set(norms) == {'lab', 'dur'} and norms['dur'].params_torch['mean'].device.type == 'cpu'
    x = np.array([[1., 5.], [2., 7.]], dtype=np.float32)
    assert np.array_equal(norms['lab'].normalise(x), np.array([[0.5, 4.], [1., 6.]], dtype=np.float32))
    batches = [{'x': torch.arange(4.), 'name': ['a'], 'nested': (torch.ones(2), 3)} for _ in range(3)]
    feeder = mg.data.ToDeviceWrapper(batches, 'cpu')
    assert len(feeder) == 3
    out = list(feeder)
    assert len(out) == 3 and out[0]['name'] == ['a'] and out[0]['nested'][1] == 3
    assert torch.equal(out[2]['x'], batches[2]['x']) and isinstance(out[1]['nested'], tuple)


def test_metric_containers_host_logic(mg):
    """Handler / Print / History bookkeeping (morgana/metrics.py:52-260) needs no device."""
    M = mg.metrics
    calls = []

    class Probe(M.StatefulMetric):
        def accumulate(self, *args, **kwargs):
            M.StatefulMetric.accumulate(self)
            calls.append((args, kwargs))

        def result(self, *args):
            return len(calls)

    handler = M.Handler(a=Probe(), hidden=Probe(hidden=True))
    handler.add_metrics('valid', v=Probe())
    handler.add_collection('extra', from_collections=('valid',))
    assert set(handler['train']) == {'a', 'hidden'} and set(handler['valid']) == {'a', 'hidden', 'v'}
    assert set(handler['extra']) == {'a', 'hidden', 'v'} and set(handler.metrics) == {'a', 'hidden', 'v'}
    handler.accumulate('valid', a=(1, 2, {'seq_len': 3}), v=7, hidden=(8,))
    assert calls == [((1, 2), {'seq_len': 3}), ((7,), {}), ((8,), {})]
    assert handler.results_as_json_dict('valid') == {'a': 3, 'v': 3}           # the hidden metric is not reported
    assert handler.results_as_str_dict('valid', prefix='p_') == {'p_a': '3', 'p_v': '3'}
    handler.reset_state('valid')
    assert handler.results_as_json_dict('valid') == {}
    with pytest.raises(ValueError):
        handler['nope']
    last, hist = M.Print(), M.History(max_len=2)
    for v in (1, 2, 3):
        last.accumulate(v)
        hist.accumulate([v, -v])                                               # History extends by an iterable
    assert last.result() == 3 and hist.result() == [3, -3] and str(hist) == '-3' and hist.result_as_json() == '-3'
    th = M.TensorHistory(3, max_len=4)
    th.accumulate(torch.arange(18.).reshape(2, 3, 3))                         # no seq_len: a plain reshape, CPU is fine
    assert th.result().shape == (4, 3) and th.result()[-1].tolist() == [15., 16., 17.]


def test_sequence_mask_matches_golden(mg, golden):
    g = golden('sequence_mask')
    seq_len = torch.from_numpy(g['mask_seq_len'])
    assert np.array_equal(mg.utils.sequence_mask(seq_len).numpy(), g['mask_default'])
    assert np.array_equal(mg.utils.sequence_mask(seq_len, max_len=7, dtype=torch.float32).numpy(), g['mask_len7_f32'])


def test_torch_ops_are_registered_for_cuda_only(mg):
    """The dispatcher knows the operators and has no CPU kernel for them: CPU tensors cannot fall back to anything."""
    from morgana_b200 import torch_ops
    for name in torch_ops.OPERATORS:
        assert hasattr(torch.ops.morgana_b200, name)
    x = torch.zeros(2, 3, 4)
    with pytest.raises(NotImplementedError):
        torch.ops.morgana_b200.upsample_norm(x, torch.ones(2, 3, dtype=torch.long), None, None, 'none', -1)
    with pytest.raises(NotImplementedError):
        torch.ops.morgana_b200.masked_loss(x, x, None, 'mse')


def test_detach_batched_seqs_host_inputs_match_golden(mg, golden):
    """utils.detach_batched_seqs moves data to the host and strips padding -- no arithmetic: tensors that already live on the
    host (and NumPy inputs) go through the same slicing as the reference (utils.py:66-102)."""
    g = golden('detach')
    x, y, n = torch.from_numpy(g['detach_x']).requires_grad_(), g['detach_y'], torch.from_numpy(g['detach_n'])
    xs, ys = mg.utils.detach_batched_seqs(x, y, seq_len=n)
    raw = mg.utils.detach_batched_seqs(torch.from_numpy(y), seq_len=g['detach_n'], squeeze=False)
    for b in range(5):
        for got, want in ((xs[b], g['detach_x_%d' % b]), (ys[b], g['detach_y_%d' % b]), (raw[b], g['detach_y_raw_%d' % b])):
            assert got.shape == want.shape and np.array_equal(got, want)
    assert np.array_equal(mg.utils.detach_batched_seqs(x), g['detach_full'])


def test_wgrad_plan_invariants(mg):
    """The frame-split plan of the tcgen05 weight gradient, over many shapes, on the host: every slice of frames is
    non-empty (an empty slice would leave a CTA waiting for an accumulator that never completes), the slices cover every
    frame block exactly once, pairs only where the x tile splits into whole 64-column atoms, and the workspace matches."""
    import ctypes
    from morgana_b200 import _lib
    rng = np.random.default_rng(0)
    shapes = [(1, 1, 1), (63, 1, 32), (64, 3, 8), (65, 187, 256), (348928, 512, 600), (348928, 128, 512), (30000, 32, 128),
              (2 ** 31 - 200, 2048, 4096), (129, 2048, 64), (10 ** 6, 199, 609), (128 * 148, 256, 512), (128 * 148 + 1, 130, 130)]
    shapes += [(int(rng.integers(1, 400000)), int(rng.integers(1, 2049)), int(rng.integers(1, 2049))) for _ in range(300)]
    out = (ctypes.c_int64 * 8)()
    for M, N, K in shapes:
        assert _lib.lib.mg_linear_wgrad_plan(M, N, K, out) == 0
        tile_k, n_tiles, k_tiles, splits, per_split, n_fblocks, tile_rows, pair = list(out)
        frames = 128
        assert n_fblocks == (M + frames - 1) // frames
        assert tile_k in (64, 128, 192, 256) and tile_rows == (256 if pair else 128)
        assert n_tiles * tile_rows >= N > (n_tiles - 1) * tile_rows and k_tiles * tile_k >= K > (k_tiles - 1) * tile_k
        assert splits >= 1 and per_split >= 1
        assert splits * per_split >= n_fblocks > (splits - 1) * per_split, (M, N, K)          # covered, and the last slice is not empty
        ctas = splits * n_tiles * k_tiles * (2 if pair else 1)
        assert ctas <= max(148, n_tiles * k_tiles * (2 if pair else 1)), (M, N, K)             # one wave when the tiles allow it
        if pair:
            assert N > 128 and tile_k % 128 == 0
        assert _lib.lib.mg_linear_wgrad_workspace_bytes(M, N, K) == 4 * splits * n_tiles * tile_rows * k_tiles * tile_k


def test_objective_stream_partition_invariants(mg):
    """The stage -> CTA partition of the persistent objective kernel (host-only mg_objective_stream_plan: the same CostModel /
    cost_to_stage functions the device runs): the ranges tile [first stage behind the head starts, number of stages) exactly and in
    order, the head starts fit the tensor, and no CTA's range costs more than an even share plus one utterance boundary's worth
    of rounding -- for ragged, empty, full-length and tiny batches, forward-only and with the gradient."""
    import ctypes
    from morgana_b200 import _lib
    rng = np.random.default_rng(7)
    cases = [(256, 1387, 187, None), (1, 61, 187, [61]), (7, 5, 187, [5, 3, 5, 1, 4, 5, 2]), (6, 64, 187, [40, 0, 17, 64, 64, 1]),
             (300, 13, 187, [9] * 300), (3, 701, 187, [700, 3, 350]), (12, 40, 187, list(range(1, 13))), (1024, 1200, 187, None),
             (8, 16, 4, [16] * 8), (64, 50, 1, None), (33, 333, 224, None), (2, 16, 187, [0, 0])]
    for _ in range(60):
        B, T = int(rng.integers(1, 1025)), int(rng.integers(1, 1500))
        cases.append((B, T, int(rng.integers(1, 225)), None))
    checked = 0
    for B, T, D, lengths in cases:
        if lengths is None:
            lengths = rng.integers(0, T + 1, B) if rng.random() < 0.8 else np.full(B, T)
        n = np.clip(np.asarray(lengths, dtype=np.int64), 0, T)
        for has_grad in (1, 0):
            out = (ctypes.c_int64 * (4 + 1024))()
            seq = (ctypes.c_int64 * B)(*[int(v) for v in n])
            rc = _lib.lib.mg_objective_stream_plan(seq, B, T, D, has_grad, 8, 148, out, len(out))
            if rc == 1:                      # not a shape of the stream form (fewer than 32 rows, costs beyond 31 bits ...)
                assert B * T < 32 or B * T * 6 >= 2 ** 31 - 2 ** 24, (B, T, D)
                continue
            assert rc == 0, (B, T, D, _lib.last_error() if hasattr(_lib, 'last_error') else rc)
            grid, first_stage, n_stages = out[0], out[1], out[2]
            bounds = np.array(out[3:3 + grid + 1])
            assert n_stages == (B * T + 7) // 8 and 1 <= grid <= 296 and grid <= max(n_stages // 4, 1)
            assert first_stage == min(4 * grid, n_stages)                       # four fixed stages per CTA
            assert bounds[0] == first_stage and bounds[-1] == n_stages, (B, T, D, bounds[:3], bounds[-3:])
            assert (np.diff(bounds) >= 0).all()                                  # contiguous ranges, in order: they tile the stages
            # cost of every range under the kernel's model (valid row 6 / 2, padding row 1 / 0), rows behind the head starts only
            cost_valid, cost_pad = (6, 1) if has_grad else (2, 0)
            row_cost = np.where(np.arange(T)[None, :] < n[:, None], cost_valid, cost_pad).reshape(-1).astype(np.int64)
            row_cost[:first_stage * 8] = 0
            prefix = np.concatenate([[0], np.cumsum(row_cost)])
            ends = np.minimum(bounds * 8, B * T)
            per_cta = prefix[ends[1:]] - prefix[ends[:-1]]
            assert per_cta.sum() == prefix[-1]
            share = prefix[-1] / grid
            # a boundary is rounded up to a stage (8 rows) and the cost position to one row: nobody is more than ~two stages over
            assert per_cta.max() <= share + 2 * 8 * cost_valid + cost_valid, (B, T, D, has_grad, per_cta.max(), share)
            checked += 1
    assert checked > 100


def test_custom_ops_have_fake_kernels_and_autograd_formulas(mg):
    """torch.ops.morgana_b200.*: shape / dtype propagation on meta tensors (no memory, no GPU) for every operator, and
    a backward pass traced through fake tensors (register_fake + register_autograd)."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    T = torch.ops.morgana_b200
    for name in mg.torch_ops.OPERATORS:
        assert hasattr(T, name), name

    def meta(*shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device='meta')
    ends, n_frames, summary = T.dur_scan(meta(4, 10, 1, dtype=torch.int64))
    assert ends.shape == (4, 10) and ends.dtype == torch.int32 and n_frames.shape == (4,) and summary.shape == (4,)
    assert T.upsample_norm(meta(4, 10, 600), meta(4, 10, 1, dtype=torch.int64), None, None, 'none', 77).shape == (4, 77, 600)
    assert T.upsample_norm_backward(meta(4, 77, 600), meta(4, 10, dtype=torch.int64), None, None, 'none').shape == (4, 10, 600)
    assert T.pad_collate(meta(100, 187), meta(4, dtype=torch.int64), 30).shape == (4, 30, 187)
    assert T.normalise(meta(4, 77, 187), meta(187), meta(187), 'mvn', False).shape == (4, 77, 187)
    loss = T.masked_loss(meta(4, 77, 187), meta(4, 77, 187), meta(4, dtype=torch.int64), 'mse')
    assert loss.shape == () and loss.dtype == torch.float32
    y = T.linear_bf16(meta(100, 600, dtype=torch.bfloat16), meta(512, 600, dtype=torch.bfloat16), meta(512), 'sigmoid', True)
    assert y.shape == (100, 512) and y.dtype == torch.bfloat16
    g16, grad_b = T.act_grad_bf16(meta(100, 187), None)
    assert g16.shape == (100, 192) and g16.dtype == torch.bfloat16 and grad_b.shape == (187,)
    assert T.linear_wgrad_bf16(meta(100, 192, dtype=torch.bfloat16), meta(100, 256, dtype=torch.bfloat16), 187, 256).shape == (187, 256)
    assert T.cast_transpose_bf16(meta(187, 256)).shape == (256, 192)
    assert T.mlpg(meta(3, 50, 9), meta(9), 10, None).shape == (3, 50, 3)
    with FakeTensorMode():
        x = torch.empty(4, 10, 600, device='meta', requires_grad=True)
        frames = T.upsample_norm(x, torch.empty(4, 10, 1, dtype=torch.int64, device='meta'), torch.empty(600, device='meta'),
                                 torch.empty(600, device='meta'), 'minmax', 77)
        h = T.linear_bf16(frames.reshape(-1, 600).to(torch.bfloat16), torch.empty(8, 600, dtype=torch.bfloat16, device='meta'),
                          None, 'sigmoid', False)
        T.masked_loss(h.reshape(4, 77, 8), torch.empty(4, 77, 8, device='meta'), None, 'mse').backward()
        assert x.grad.shape == (4, 10, 600)
