"""CPU checks of the machinery that carries the unmodified reference to the GPU box (oracle/make_ref.py, ref_loader.py,
ref_harness.py, bandmat_standin.py).  The GPU comparisons themselves are in tests/test_reference_models_gpu.py."""
import os

import numpy as np
import pytest
import torch

from oracle import bandmat_standin, make_ref, np_oracle as O, ref_harness as H, ref_loader

needs_reference = pytest.mark.skipif(not ref_loader.available(), reason='neither oracle/_ref nor /root/reference is present')


@pytest.mark.skipif(not os.path.isdir('/root/reference/morgana'), reason='build container only')
def test_mirror_is_a_byte_for_byte_copy_of_the_reference():
    make_ref.make()
    assert make_ref.is_current()
    for rel in ('morgana/utils.py', 'morgana/losses.py', 'morgana/metrics.py', 'models/RNN_SPSS.py'):
        with open(os.path.join('/root/reference', rel), 'rb') as a, open(os.path.join(make_ref.DEST, rel), 'rb') as b:
            assert a.read() == b.read(), rel


def test_mirror_is_ignored_by_git_but_not_by_gpurun():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, '.gitignore')) as f:
        assert 'oracle/_ref/' in f.read().split()
    ignore = os.path.join(root, '.gpurunignore')
    if os.path.exists(ignore):
        with open(ignore) as f:
            assert not any(line.strip().startswith('oracle') for line in f)


def test_bandmat_stand_in_matches_dense_algebra():
    rng = np.random.default_rng(0)
    n = 17
    w1 = bandmat_standin.band_c_bm(1, 1, np.tile(np.array([[-0.5], [0.], [0.5]]), n)).T
    w2 = bandmat_standin.band_c_bm(1, 1, np.tile(np.array([[1.], [-2.], [1.]]), n)).T
    dense1 = w1.full()
    assert np.allclose(np.diag(dense1, 1), 0.5) and np.allclose(np.diag(dense1, -1), -0.5)      # row j: [-.5, 0, .5] around j
    tau = rng.random(n) + 0.5
    prec = bandmat_standin.zeros(2, 2, n)
    bandmat_standin.dot_mm_plus_equals(w1.T, w1, target_bm=prec, diag=tau)
    bandmat_standin.dot_mm_plus_equals(w2.T, w2, target_bm=prec, diag=tau)
    want = dense1.T @ np.diag(tau) @ dense1 + w2.full().T @ np.diag(tau) @ w2.full()
    assert np.allclose(prec.full(), want, rtol=1e-13, atol=1e-13)
    v = rng.standard_normal(n)
    b = np.zeros(n)
    bandmat_standin.dot_mv_plus_equals(w2.T, v, target=b)
    assert np.allclose(b, w2.full().T @ v, rtol=1e-13, atol=1e-13)
    spd = bandmat_standin.zeros(2, 2, n)
    bandmat_standin.dot_mm_plus_equals(w2.T, w2, target_bm=spd, diag=tau)
    spd.add_to_diagonal(0, 0, np.ones(n))
    assert np.allclose(bandmat_standin.solveh(spd, v), np.linalg.solve(spd.full(), v), rtol=1e-10, atol=1e-12)


@needs_reference
def test_reference_mlpg_runs_on_the_stand_in_and_equals_the_oracle():
    ref_loader.import_reference()
    from morgana.viz.synthesis import MLPG
    rng = np.random.default_rng(1)
    means = rng.standard_normal((3, 41, 12)).astype(np.float32)
    var = (rng.random(12) + 0.3).astype(np.float32)
    n = np.array([41, 30, 5])
    got = MLPG(means, var, padding_size=7, seq_len=n)
    # float32 inputs: the reference forms mean / variance and 1 / variance in float32 before widening (synthesis.py:162-163),
    # the oracle restates the definition in float64 throughout -> agreement at float32 resolution
    assert np.allclose(got, O.mlpg_banded(means, var, padding_size=7, seq_len=n), rtol=1e-6, atol=1e-6)
    assert np.allclose(got, O.mlpg(means, var, padding_size=7, seq_len=n), rtol=1e-6, atol=1e-6)
    got64 = MLPG(means.astype(np.float64), var.astype(np.float64), padding_size=7, seq_len=n)
    assert np.allclose(got64, O.mlpg_banded(means, var, padding_size=7, seq_len=n), rtol=1e-9, atol=1e-9)


@needs_reference
def test_reference_models_and_train_epoch_run_through_the_harness_on_cpu():
    morgana = ref_loader.import_reference()
    module = ref_loader.load_model_module_as('RNN_SPSS', 'ref_models_cpu_check')
    params = H.normaliser_params(seed=2)
    features = H.make_features(batch_size=3, seed=2, params=params)
    model = H.build_model(morgana, module.LSTMAcousticModel, params, 'cpu', output_dims=H.OUTPUT_DIMS_187, num_layers=1)
    loss, outputs, sums, grads = H.forward_backward(model, features)
    n_frames = int(features['n_frames'].sum())
    assert torch.isfinite(loss) and outputs['mcep'].shape[-1] == 60
    assert sums['VUV_accuracy'][1] == n_frames and sums['MCEP_distortion'][1] == n_frames      # Q2: counts are frames
    # the loss the model computed is the oracle's formula on the model's own predictions (models/RNN_SPSS.py:131-139)
    want = sum(O.masked_loss(outputs[k].detach().numpy(), features[k].numpy(), features['n_frames'].numpy())
               for k in ('normalised_lf0_deltas', 'normalised_mcep_deltas', 'normalised_bap_deltas'))
    want += O.masked_loss(outputs['vuv'].detach().numpy(), features['vuv'].float().numpy(), features['n_frames'].numpy(), 'bce')
    assert abs(float(loss) - want / 4.) <= 1e-6 * abs(want / 4.)
    ema_model = H.build_model(morgana, module.LSTMAcousticModel, params, 'cpu', state_dict=model.state_dict(),
                              output_dims=H.OUTPUT_DIMS_187, num_layers=1)
    builder = H.make_experiment(morgana, model, ema_model, 0.999)
    before = {n: p.detach().clone() for n, p in ema_model.named_parameters()}
    epoch_loss = H.train_epoch(builder, [features], torch.optim.Adam(model.parameters(), lr=1e-3))
    assert np.isfinite(epoch_loss)
    assert any(not torch.equal(before[n], p) for n, p in ema_model.named_parameters())            # experiment_builder.py:484 ran
