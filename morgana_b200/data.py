"""Drop-in counterparts of the reference's normaliser arithmetic (``morgana/data.py:292-340, 533-616``).

The four free functions keep the reference's names and arguments.  Torch tensors go to the CUDA kernel (K3, or fused
into the upsampling gather through :meth:`_FeatureNormaliser.fused_params`); NumPy arrays -- which the reference
normalises per utterance inside DataLoader worker processes (``data.py:119-127``), where CUDA must not be touched --
take the same arithmetic in NumPy.  There is no CPU path for torch tensors.
"""
import json
import os

import numpy as np
import torch

from morgana_b200 import ops


def pad_collate(packed, lengths, max_len=None):
    """Zero-pad packed per-utterance rows to ``(B, T, D)`` on the device -- ``collate_fn``'s padding (reference
    morgana/data.py:184-193) moved after the host->device copy, so only valid rows cross PCIe."""
    return ops.pad_collate(packed, lengths, max_len=max_len)


def _np_scale(mmin, mmax):
    scale = mmax - mmin
    scale[abs(scale) <= 1e-8] = 1.
    return scale


def normalise_mvn(feature, mean, std_dev):
    """``(feature - mean) / (std_dev + 1e-8)`` (morgana/data.py:533-534)."""
    if isinstance(feature, np.ndarray):
        return (feature - mean[..., None, :]) / (std_dev[..., None, :] + 1e-8)
    return ops.normalise(feature, mean, std_dev, 'mvn', inverse=False)


def denormalise_mvn(feature, mean, std_dev):
    """``feature * std_dev + mean`` (morgana/data.py:537-538)."""
    if isinstance(feature, np.ndarray):
        return (feature * std_dev[..., None, :]) + mean[..., None, :]
    return ops.normalise(feature, mean, std_dev, 'mvn', inverse=True)


def normalise_minmax(feature, mmin, mmax):
    """``(feature - mmin) / scale`` with ``scale = mmax - mmin`` forced to 1 on constant dims (morgana/data.py:579-583)."""
    if isinstance(feature, np.ndarray):
        return (feature - mmin[..., None, :]) / _np_scale(mmin, mmax)[..., None, :]
    return ops.normalise(feature, mmin, mmax, 'minmax', inverse=False)


def denormalise_minmax(feature, mmin, mmax):
    """``feature * scale + mmin`` (morgana/data.py:586-590)."""
    if isinstance(feature, np.ndarray):
        return (feature * _np_scale(mmin, mmax)[..., None, :]) + mmin[..., None, :]
    return ops.normalise(feature, mmin, mmax, 'minmax', inverse=True)


class _FeatureNormaliser(object):
    r"""``normalise(feature, deltas=False)`` / ``denormalise(feature, deltas=False)`` on NumPy arrays or CUDA tensors.

    Mirrors morgana/data.py:252-385.  Parameters come from :meth:`set_params` (synthetic / in-memory) or
    :meth:`load_params` (the reference's ``{name}_{kind}.json`` files).
    """
    kind = None
    param_names = ()
    file_suffix = ''

    def __init__(self, name, use_deltas=False, file_pattern=None):
        self.name = name
        self.use_deltas = use_deltas
        self.file_pattern = file_pattern or '{name}_' + self.file_suffix + '.json'
        self.params = None
        self.params_torch = None
        if self.use_deltas:
            self.delta_params = None
            self.delta_params_torch = None

    # -- parameters -------------------------------------------------------------------------------------------------
    @staticmethod
    def _to_torch(params, device):
        return {k: torch.tensor(np.asarray(v, dtype=np.float32)).to(device) for k, v in params.items()}

    def set_params(self, params, delta_params=None, device='cuda'):
        self.params = {k: np.asarray(v, dtype=np.float32) for k, v in params.items()}
        self.params_torch = self._to_torch(self.params, device)
        if delta_params is not None:
            self.use_deltas = True
            self.delta_params = {k: np.asarray(v, dtype=np.float32) for k, v in delta_params.items()}
            self.delta_params_torch = self._to_torch(self.delta_params, device)
        return self

    @staticmethod
    def _from_json(file_path):
        with open(file_path) as f:
            feat_params = json.load(f)
        return {k: np.array(v, dtype=np.float32) for k, v in feat_params.items()}

    def load_params(self, data_dir, data_root='.', device='cpu'):
        """Loads ``{data_root}/{data_dir}/{name}_{kind}.json`` (and the ``_deltas`` file), morgana/data.py:362-385."""
        params_file = os.path.join(data_root, data_dir, self.file_pattern.format(name=self.name))
        delta = None
        if self.use_deltas:
            delta = self._from_json(os.path.join(data_root, data_dir, self.file_pattern.format(name=self.name + '_deltas')))
        return self.set_params(self._from_json(params_file), delta, device=device)

    def fetch_params(self, data_type=np.ndarray, deltas=False):
        if deltas:
            return self.delta_params_torch if data_type == torch.Tensor else self.delta_params
        return self.params_torch if data_type == torch.Tensor else self.params

    def fused_params(self, deltas=False):
        """``(kind, p0, p1)`` for fusing this normaliser into ``utils.upsample_to_repetitions``."""
        params = self.fetch_params(torch.Tensor, deltas=deltas)
        return (self.kind,) + tuple(params[n] for n in self.param_names)

    # -- arithmetic -------------------------------------------------------------------------------------------------
    def _args(self, feature, deltas):
        data_type = torch.Tensor if isinstance(feature, torch.Tensor) else np.ndarray
        params = self.fetch_params(data_type, deltas=deltas)
        return tuple(params[n] for n in self.param_names)

    def normalise(self, feature, deltas=False):
        return self._normalise(feature, *self._args(feature, deltas))

    def denormalise(self, feature, deltas=False):
        return self._denormalise(feature, *self._args(feature, deltas))


class MeanVarianceNormaliser(_FeatureNormaliser):
    """Zero mean, unit variance (morgana/data.py:541-564)."""
    kind, param_names, file_suffix = 'mvn', ('mean', 'std_dev'), 'mvn'
    _normalise = staticmethod(normalise_mvn)
    _denormalise = staticmethod(denormalise_mvn)


class MinMaxNormaliser(_FeatureNormaliser):
    """Minimum 0, maximum 1 (morgana/data.py:593-616)."""
    kind, param_names, file_suffix = 'minmax', ('mmin', 'mmax'), 'minmax'
    _normalise = staticmethod(normalise_minmax)
    _denormalise = staticmethod(denormalise_minmax)


class _SpeakerDependentNormaliser(_FeatureNormaliser):
    r"""One parameter set per speaker, selected per batch item by ``speaker_ids`` (morgana/data.py:388-531).

    The reference assembles the ``(batch_size, feat_dim)`` parameters of a batch with one ``torch.cat`` per batch item
    and parameter (``data.py:482-498``); here every parameter is one ``(n_speakers, feat_dim)`` table on the device and a
    batch's rows are gathered with a single ``index_select``.  The kernels (K3, and K2 when fused) take the
    ``(batch_size, feat_dim)`` parameters directly -- row ``b`` applies to every frame of utterance ``b``.
    """
    def __init__(self, name, speaker_id_list=None, use_deltas=False, file_pattern=None):
        _FeatureNormaliser.__init__(self, name, use_deltas=use_deltas,
                                    file_pattern=file_pattern or '{speaker_id}/{name}_' + self.file_suffix + '.json')
        self.speaker_id_list = speaker_id_list
        self.speaker_ids = None
        self.params, self.params_torch = {}, {}
        if self.use_deltas:
            self.delta_params, self.delta_params_torch = {}, {}
        self._tables = {}          # deltas flag -> (speaker -> row, {param name: (n_speakers, feat_dim) device tensor})

    # -- parameters -------------------------------------------------------------------------------------------------
    def set_params(self, params, delta_params=None, device='cuda'):
        """``params``: ``{speaker_id: {param name: (feat_dim,) array}}`` (and the same for the delta features)."""
        self.speaker_ids = list(params.keys())
        self.params = {s: {k: np.asarray(v, dtype=np.float32) for k, v in p.items()} for s, p in params.items()}
        self.params_torch = {s: self._to_torch(p, device) for s, p in self.params.items()}
        if delta_params is not None:
            self.use_deltas = True
            self.delta_params = {s: {k: np.asarray(v, dtype=np.float32) for k, v in p.items()} for s, p in delta_params.items()}
            self.delta_params_torch = {s: self._to_torch(p, device) for s, p in self.delta_params.items()}
        self._tables = {}
        return self

    def load_params(self, data_dir, data_root='.', device='cpu'):
        """Loads ``{data_root}/{data_dir}/{speaker_id}/{name}_{kind}.json`` for every speaker of the id list (one id per
        line), morgana/data.py:506-531."""
        if self.speaker_ids is None:
            with open(os.path.join(data_root, self.speaker_id_list)) as f:
                self.speaker_ids = [line.strip() for line in f if line.strip()]
        params, delta_params = {}, ({} if self.use_deltas else None)
        for speaker_id in self.speaker_ids:
            params[speaker_id] = self._from_json(os.path.join(
                data_root, data_dir, self.file_pattern.format(name=self.name, speaker_id=speaker_id)))
            if self.use_deltas:
                delta_params[speaker_id] = self._from_json(os.path.join(
                    data_root, data_dir, self.file_pattern.format(name=self.name + '_deltas', speaker_id=speaker_id)))
        return self.set_params(params, delta_params, device=device)

    def _table(self, deltas):
        if deltas not in self._tables:
            per_speaker = self.delta_params_torch if deltas else self.params_torch
            rows = {speaker_id: i for i, speaker_id in enumerate(per_speaker)}
            stacked = {name: torch.stack([per_speaker[s][name] for s in per_speaker]) for name in self.param_names}
            self._tables[deltas] = (rows, stacked)
        return self._tables[deltas]

    def fetch_params(self, speaker_ids, data_type=np.ndarray, deltas=False):
        """``{param name: (batch_size, feat_dim)}`` for a list of speakers, ``(feat_dim,)`` for a single speaker."""
        single = not isinstance(speaker_ids, (list, tuple))
        speaker_ids = [speaker_ids] if single else list(speaker_ids)
        if data_type == torch.Tensor:
            rows, stacked = self._table(deltas)
            index = torch.tensor([rows[s] for s in speaker_ids], device=next(iter(stacked.values())).device)
            out = {name: table.index_select(0, index) for name, table in stacked.items()}
        else:
            per_speaker = self.delta_params if deltas else self.params
            out = {name: np.stack([per_speaker[s][name] for s in speaker_ids]) for name in self.param_names}
        if len(speaker_ids) == 1:                 # the reference squeezes a one-speaker batch to (feat_dim,), data.py:500-501
            out = {name: value[0] for name, value in out.items()}
        return out

    def fused_params(self, speaker_ids, deltas=False):
        """``(kind, p0, p1)`` with per-utterance parameters, for ``utils.upsample_to_repetitions(..., normaliser=...)``."""
        params = self.fetch_params(speaker_ids, torch.Tensor, deltas=deltas)
        return (self.kind,) + tuple(params[n] for n in self.param_names)

    # -- arithmetic -------------------------------------------------------------------------------------------------
    def _sd_args(self, feature, speaker_ids, deltas):
        data_type = torch.Tensor if isinstance(feature, torch.Tensor) else np.ndarray
        params = self.fetch_params(speaker_ids, data_type, deltas=deltas)
        return tuple(params[n] for n in self.param_names)

    def normalise(self, feature, speaker_ids, deltas=False):
        return self._normalise(feature, *self._sd_args(feature, speaker_ids, deltas))

    def denormalise(self, feature, speaker_ids, deltas=False):
        return self._denormalise(feature, *self._sd_args(feature, speaker_ids, deltas))


class SpeakerDependentMeanVarianceNormaliser(_SpeakerDependentNormaliser):
    """Per-speaker zero mean, unit variance (morgana/data.py:567-576)."""
    kind, param_names, file_suffix = 'mvn', ('mean', 'std_dev'), 'mvn'
    _normalise = staticmethod(normalise_mvn)
    _denormalise = staticmethod(denormalise_mvn)


class SpeakerDependentMinMaxNormaliser(_SpeakerDependentNormaliser):
    """Per-speaker minimum 0, maximum 1 (morgana/data.py:619-628)."""
    kind, param_names, file_suffix = 'minmax', ('mmin', 'mmax'), 'minmax'
    _normalise = staticmethod(normalise_minmax)
    _denormalise = staticmethod(denormalise_minmax)


class Normalisers(dict):
    r"""Dictionary of normalisers with their parameters loaded (morgana/data.py:227-249).  ``device`` is honoured: the
    reference passes it in ``data_root``'s position and leaves every parameter on the CPU."""
    def __init__(self, normaliser_sources, normalisation_dir, data_root='.', device='cpu'):
        dict.__init__(self)
        self.normalisation_dir = os.path.join(data_root, normalisation_dir)
        self.device = device
        for name, normaliser in normaliser_sources.items():
            self[name] = normaliser
            normaliser.load_params(self.normalisation_dir, device=self.device)


def _map_nested(func, data):
    """Apply ``func`` to every tensor / array leaf of nested dicts, lists and tuples (morgana/utils.py:37-54)."""
    if isinstance(data, (np.ndarray, torch.Tensor)):
        return func(data)
    if isinstance(data, dict):
        return {k: _map_nested(func, v) for k, v in data.items()}
    if isinstance(data, (list, tuple)):
        return type(data)(_map_nested(func, v) for v in data)
    return data


class ToDeviceWrapper(object):
    r"""Iterates over a data loader and moves every batch to ``device`` (morgana/data.py:631-663), as a feeder for a GPU
    that consumes a batch in well under a millisecond: tensors are staged in pinned host memory and copied on a side
    stream ONE BATCH AHEAD, so the upload of batch i + 1 runs under the kernels of batch i.  The consumer's stream only
    waits on the copy's event; nothing synchronises with the host.  For a CPU ``device`` batches pass through untouched.
    """
    def __init__(self, data_loader, device, prefetch=True):
        self.data_loader = data_loader
        self.torch_device = torch.device(device)
        self.prefetch = prefetch and self.torch_device.type == 'cuda'

    def __getattr__(self, attr):            # attribute access falls through to the wrapped loader (data.py:636-642)
        return getattr(self.__dict__['data_loader'], attr)

    def __len__(self):
        return len(self.data_loader)

    def to_device(self, tensor):
        if not isinstance(tensor, torch.Tensor):
            return tensor
        if self.torch_device.type == 'cuda' and not tensor.is_cuda:
            if not tensor.is_pinned():
                tensor = tensor.pin_memory()
            return tensor.to(self.torch_device, non_blocking=True)
        return tensor.to(self.torch_device)

    def __iter__(self):
        if not self.prefetch:
            for features in self.data_loader:
                yield _map_nested(self.to_device, features)
            return
        copy_stream = torch.cuda.Stream(self.torch_device)

        def upload(features):
            with torch.cuda.stream(copy_stream):
                moved = _map_nested(self.to_device, features)
                done = torch.cuda.Event()
                done.record(copy_stream)
            return moved, done

        def hand_over(moved, done):
            consumer = torch.cuda.current_stream(self.torch_device)
            consumer.wait_event(done)
            _map_nested(lambda t: t.record_stream(consumer) if isinstance(t, torch.Tensor) and t.is_cuda else None, moved)
            return moved
        pending = None
        for features in self.data_loader:
            ahead = upload(features)
            if pending is not None:
                yield hand_over(*pending)
            pending = ahead
        if pending is not None:
            yield hand_over(*pending)
