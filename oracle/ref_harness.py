"""Harness that drives the UNMODIFIED reference models and training loop on synthetic features.  TEST INFRASTRUCTURE ONLY.

What it runs is the reference's own code, loaded by ``oracle/ref_loader.py``: ``models/RNN_SPSS.py:LSTMAcousticModel``
(``predict`` :72-105, ``_prepare_output`` :107-118, ``loss`` :120-139), ``models/f0_test_model.py:F0Model`` (:77-108) and
``morgana/experiment_builder.py:ExperimentBuilder.train_epoch`` (:431-505, EMA at :484).  What it replaces is only what the
tier rules put out of scope -- the file-backed Dataset / DataLoader (features are synthesised here, with the dictionary keys
``FilesDataset.collate_fn`` would produce) and ``ExperimentBuilder.__init__`` (CLI, directories, logging) -- plus three
adapters for defects of the reference itself, applied identically to every arm (unpatched CPU, unpatched CUDA, patched CUDA):

* Q1: ``SequentialWithRecurrent.forward`` returns ``(output, hiddens)`` (utils.py:401-418) but both shipped
  models use the return value as a tensor (RNN_SPSS.py:83-88) -> a forward hook on ``model.layers`` keeps element 0.
* Q14 (torch drift): ``RecurrentCuDNNWrapper`` hands device-resident lengths to ``pack_padded_sequence`` (utils.py:366-369),
  which torch >= 1.7 rejects ("lengths should be a 1D CPU int64 tensor") -> ``cpu_lengths()`` moves them to the host first.
  The cuDNN wrappers are outside the path; this only lets the shipped models run on a CUDA device at all.
* Q3: metric state exists only after ``reset_state()`` -> the harness calls ``model.metrics.reset_state(mode)`` like
  ``train_epoch`` does (:449).

Nothing under ``morgana_b200/`` imports this module.
"""
import contextlib

import numpy as np
import torch

LAB_DIM, COUNTER_DIM = 600, 9
OUTPUT_DIMS_187 = {'lf0': 3, 'vuv': 1, 'mcep': 180, 'bap': 3}       # BASELINE.json configs[3]: 187-dim WORLD targets
STATIC_DIMS = {'lf0': 1, 'mcep': 60, 'bap': 1}


# ----------------------------------------------------------------------------------------------------------------------
# synthetic features with the keys FilesDataset.collate_fn would produce (data.py:159-224), on the CPU
# ----------------------------------------------------------------------------------------------------------------------
def with_deltas(static):
    """[static | delta | delta-delta] under the windows MLPG assumes (viz/synthesis.py:121-126), zeros beyond the ends."""
    padded = torch.nn.functional.pad(static, (0, 0, 1, 1))
    prev, nxt = padded[:, :-2], padded[:, 2:]
    return torch.cat([static, 0.5 * (nxt - prev), prev - 2. * static + nxt], dim=-1)


def normaliser_params(seed=1234):
    """name -> (params, delta_params or None) as float32 NumPy arrays, the layout ``_FeatureNormaliser._from_json`` loads."""
    rng = np.random.default_rng(seed)
    f32 = np.float32
    mmax = (rng.random(LAB_DIM) + 0.5).astype(f32)
    mmax[::97] = 0.                                                   # constant dims: scale forced to 1 (data.py:581)
    params = {'lab': ({'mmin': np.zeros(LAB_DIM, f32), 'mmax': mmax}, None),
              'counters': ({'mmin': np.zeros(COUNTER_DIM, f32), 'mmax': np.full(COUNTER_DIM, 40., f32)}, None),
              'dur': ({'mean': np.full(1, 15., f32), 'std_dev': np.full(1, 8., f32)}, None)}
    for name, dim in STATIC_DIMS.items():
        mean = 5. if name == 'lf0' else 0.
        params[name] = ({'mean': np.full(dim, mean, f32), 'std_dev': (0.4 + 0.2 * rng.random(dim)).astype(f32)},
                        {'mean': np.concatenate([np.full(dim, mean), np.zeros(2 * dim)]).astype(f32),
                         'std_dev': np.concatenate([0.4 + 0.2 * rng.random(dim), 0.15 + 0.1 * rng.random(dim),
                                                    0.25 + 0.1 * rng.random(dim)]).astype(f32)})
    return params


def make_features(batch_size=4, min_phones=6, max_phones=12, max_dur=8, seed=1234, params=None):
    """One padded batch on the CPU: linguistic inputs at phone rate, acoustic targets at frame rate (zeros in the padding)."""
    params = params or normaliser_params(seed)
    g = torch.Generator().manual_seed(seed)
    n_phones = torch.randint(min_phones, max_phones + 1, (batch_size,), generator=g)
    P = int(n_phones.max())
    valid_phone = torch.arange(P)[None] < n_phones[:, None]
    dur = (torch.randint(1, max_dur + 1, (batch_size, P), generator=g) * valid_phone)[:, :, None]
    n_frames = dur.sum(dim=(1, 2))
    T = int(n_frames.max())
    valid_frame = (torch.arange(T)[None] < n_frames[:, None])[:, :, None]
    mmax = torch.from_numpy(params['lab'][0]['mmax'])
    lab = torch.rand(batch_size, P, LAB_DIM, generator=g) * torch.where(mmax > 0, mmax, torch.ones_like(mmax)) * valid_phone[:, :, None]
    scale = torch.where(mmax.abs() <= 1e-8, torch.ones_like(mmax), mmax)
    feats = {'name': ['utt_%03d' % i for i in range(batch_size)], 'n_frames': n_frames, 'n_phones': n_phones, 'dur': dur,
             'lab': lab, 'normalised_lab': (lab / scale) * valid_phone[:, :, None],
             'normalised_counters': torch.rand(batch_size, T, COUNTER_DIM, generator=g) * valid_frame}
    for name, dim in STATIC_DIMS.items():
        smooth = torch.cumsum(0.1 * torch.randn(batch_size, T, dim, generator=g), dim=1)
        static = (smooth + (5. if name == 'lf0' else 0.)) * valid_frame
        deltas = with_deltas(static) * valid_frame
        mean, std = (torch.from_numpy(params[name][1][k]) for k in ('mean', 'std_dev'))
        feats[name] = static
        feats[name + '_deltas'] = deltas
        feats['normalised_%s_deltas' % name] = ((deltas - mean) / (std + 1e-8)) * valid_frame
    feats['vuv'] = (torch.rand(batch_size, T, 1, generator=g) < 0.6) & valid_frame
    return feats


def to_device(features, device):
    return {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in features.items()}


# ----------------------------------------------------------------------------------------------------------------------
# models and the training loop of the reference
# ----------------------------------------------------------------------------------------------------------------------
def make_normalisers(morgana, params, device):
    """The reference's own normaliser objects (data.py:541-616) with parameters set as ``load_params`` would (:362-372)."""
    kinds = {'lab': morgana.data.MinMaxNormaliser, 'counters': morgana.data.MinMaxNormaliser}
    out = {}
    for name, (p, dp) in params.items():
        norm = kinds.get(name, morgana.data.MeanVarianceNormaliser)(name, use_deltas=dp is not None)
        norm.params = {k: v.copy() for k, v in p.items()}
        norm.params_torch = norm._to_torch(norm.params, device=device)
        if dp is not None:
            norm.delta_params = {k: v.copy() for k, v in dp.items()}
            norm.delta_params_torch = norm._to_torch(norm.delta_params, device=device)
        out[name] = norm
    return out


def first_of_tuple_adapter(model):
    """Q1: keep the tensor of ``SequentialWithRecurrent``'s ``(output, hiddens)`` return."""
    return model.layers.register_forward_hook(lambda module, args, output: output[0] if isinstance(output, tuple) else output)


def build_model(morgana, model_class, params, device, state_dict=None, seed=0, **model_kwargs):
    torch.manual_seed(seed)
    model = model_class(**model_kwargs)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    model.to(device)
    model.normalisers = make_normalisers(morgana, params, device)
    first_of_tuple_adapter(model)
    return model


@contextlib.contextmanager
def cpu_lengths():
    """Q14: lengths reach ``pack_padded_sequence`` on the host, as torch >= 1.7 demands (utils.py:366-369 passes device ones)."""
    rnn = torch.nn.utils.rnn
    original = rnn.pack_padded_sequence

    def pack_padded_sequence(input, lengths, *args, **kwargs):
        if isinstance(lengths, torch.Tensor) and lengths.device.type != 'cpu':
            lengths = lengths.cpu()
        return original(input, lengths, *args, **kwargs)
    rnn.pack_padded_sequence = pack_padded_sequence
    try:
        yield
    finally:
        rnn.pack_padded_sequence = original


def forward_backward(model, features, mode='train'):
    """``BaseSPSS.forward`` (base_models.py:317-321) + backward, the inner part of ``train_epoch`` (:464-470).

    Returns (loss tensor, outputs, {metric name: (sum, count)} as Python floats, {param name: grad}); the gradient of the loss
    with respect to every differentiable output (the boundary between the path and the layers) is kept in ``outputs[k].grad``.
    """
    model.mode = mode
    model.metrics.reset_state(mode)
    model.zero_grad()
    with cpu_lengths():
        output_features = model.predict(features)           # BaseSPSS.forward, with the outputs' gradients retained
        for value in output_features.values():
            if isinstance(value, torch.Tensor) and value.requires_grad:
                value.retain_grad()
        loss, outputs = model.loss(features, output_features), output_features
        loss.backward()
    sums = {}
    for name, metric in model.metrics[mode].items():
        if hasattr(metric, 'sum'):
            sums[name] = (float(metric.sum), float(metric.count))
    grads = {name: p.grad.detach().clone() for name, p in model.named_parameters() if p.grad is not None}
    return loss, outputs, sums, grads


class _Batches(list):
    """A list of feature dictionaries standing in for the DataLoader (``len()`` + iteration are all train_epoch uses)."""


def make_experiment(morgana, model, ema_model=None, ema_decay=0., epoch=1):
    """An ``ExperimentBuilder`` carrying exactly the attributes ``train_epoch`` / ``valid_epoch`` read, without running
    ``__init__`` (argument parsing, directories, data loading: control plane, out of scope)."""
    builder = object.__new__(morgana.experiment_builder.ExperimentBuilder)
    builder.model, builder.epoch, builder.ema_decay = model, epoch, ema_decay
    builder.lr_schedule_name, builder.analysis_kwargs = 'constant', {}
    if ema_decay:
        builder.ema = morgana.utils.ExponentialMovingAverage(model=ema_model, decay=ema_decay)   # experiment_builder.py:281
    return builder


def train_epoch(builder, batches, optimizer):
    with cpu_lengths():
        return builder.train_epoch(_Batches(batches), optimizer)
