"""Rebind the reference's hot-path callables to the sm_100a kernels, leaving everything else of morgana untouched.

    import morgana, morgana_b200
    morgana_b200.patch()        # models/*.py, BaseSPSS subclasses and ExperimentBuilder now run on the new kernels
    ...
    morgana_b200.unpatch()      # restore the reference (e.g. for a CPU oracle run)

What is rebound: ``morgana.utils.upsample_to_repetitions``, ``morgana.utils.ExponentialMovingAverage``,
``morgana.losses.mse`` / ``bce``, ``morgana.data.normalise_*`` / ``denormalise_*``, ``morgana.viz.synthesis.MLPG`` and the
``accumulate`` / ``result`` / ``reset_state`` methods of the metric accumulators.  NumPy inputs to the normalisers (DataLoader workers) keep going
through NumPy arithmetic; torch tensors must be on a CUDA device.
"""
from morgana_b200 import data as _data
from morgana_b200 import losses as _losses
from morgana_b200 import metrics as _metrics
from morgana_b200 import utils as _utils

_saved = {}
_METRIC_METHODS = ('reset_state', 'accumulate', 'result')


def _swap(owner, name, new):
    _saved[(owner, name)] = owner.__dict__.get(name, _saved.get((owner, name)))
    setattr(owner, name, new)


def patch(morgana=None):
    """Monkey-patch an imported ``morgana`` package (imported here if not given).  Idempotent."""
    if morgana is None:
        import morgana
    if _saved:
        return morgana
    _swap(morgana.utils, 'upsample_to_repetitions', _utils.upsample_to_repetitions)
    _swap(morgana.utils, 'ExponentialMovingAverage', _utils.ExponentialMovingAverage)
    for name in ('batched_masked_select', 'get_segment_ends', 'split_to_segments', 'both_voiced_mask', 'detach_batched_seqs'):
        if hasattr(morgana.utils, name):
            _swap(morgana.utils, name, getattr(_utils, name))
    _swap(morgana.losses, 'mse', _losses.mse)
    _swap(morgana.losses, 'bce', _losses.bce)
    for name in ('ce', 'KLD_standard_normal', 'sequence_loss'):
        if hasattr(morgana.losses, name):
            _swap(morgana.losses, name, getattr(_losses, name))
    for name in ('normalise_mvn', 'denormalise_mvn', 'normalise_minmax', 'denormalise_minmax'):
        _swap(morgana.data, name, getattr(_data, name))
    if hasattr(morgana, 'viz') and hasattr(morgana.viz, 'synthesis'):
        # Model scripts bind the name at import (`from morgana.viz.synthesis import MLPG`, models/RNN_SPSS.py:9), so
        # patch() must run before they are imported for this rebinding to reach them.
        from morgana_b200.viz import synthesis as _synthesis
        _swap(morgana.viz.synthesis, 'MLPG', _synthesis.MLPG)
    for cls in _metrics.ACCUMULATORS:
        ref_cls = getattr(morgana.metrics, cls.__name__)
        for method in _METRIC_METHODS:
            if method in cls.__dict__:
                _swap(ref_cls, method, cls.__dict__[method])
    return morgana


def unpatch():
    """Undo :func:`patch`."""
    for (owner, name), old in list(_saved.items()):
        if old is None:
            try:
                delattr(owner, name)
            except AttributeError:
                pass
        else:
            setattr(owner, name, old)
    _saved.clear()
