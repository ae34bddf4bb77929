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
