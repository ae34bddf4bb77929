"""Data-parallel ``train_epoch`` / ``valid_epoch`` ("next" row 2 of SURVEY.md section 8f).

The reference's loops (``morgana/experiment_builder.py:431-505`` and ``:562-620``) are single-device and synchronise with
the host several times per step (``batch_loss.item()``, the ``.item()`` inside every metric, the progress-bar strings).
:class:`DataParallelTrainer` keeps their order of operations and the model protocol they rely on --

    ``model(features) -> (loss, output_features)``, ``model.mode``, ``model.step``, ``model.metrics`` (a ``Handler``),

-- and changes what limits a multi-GPU step:

* one process per GPU (``torchrun``); every rank iterates over its own shard of the utterances
  (:func:`morgana_b200.dp.sharded_batches`), same number of steps on every rank;
* gradients live in ONE flat buffer (:class:`GradientBucket`: every ``param.grad`` is a view of it), so the exchange is
  a single all-reduce with no packing copies and ``zero_grad`` is a single fill;
* the epoch loss is accumulated on the device; nothing in the loop reads a value back unless ``log_every`` asks;
* the metric state (device records of :mod:`morgana_b200.metrics`) and the epoch loss are summed over the ranks with one
  small all-reduce at the end of the epoch (``sum`` and ``count`` are additive), and written back so
  ``model.metrics.results_as_json_dict(mode)`` reports global numbers on every rank;
* validation with the EMA model accumulates into the EMA model's own ``metrics`` (the reference mixes the two models'
  handlers at ``experiment_builder.py:602``).

Checkpoints, logging, output generation, the CLI and the LR-schedule zoo stay where they are (out of scope); the caller
passes any ``lr_schedule`` object with a ``step()`` and says whether it is batch-level.
"""
import torch
import torch.distributed as dist

from morgana_b200 import dp, ops


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class GradientBucket(object):
    """Flat gradient storage for a list of parameters: ``param.grad`` of each is a view into ``self.flat``.

    ``zero()`` is one fill, ``all_reduce()`` one collective on the whole buffer (mean over the ranks).  Optimisers must
    be stepped with ``zero_grad(set_to_none=False)`` semantics -- call :meth:`zero` instead of ``optimizer.zero_grad()``.
    """
    def __init__(self, parameters):
        self.params = [p for p in parameters if p.requires_grad]
        if not self.params:
            raise ValueError('GradientBucket needs at least one trainable parameter')
        first = self.params[0]
        if any(p.dtype != first.dtype or p.device != first.device for p in self.params):
            raise ValueError('all parameters of a GradientBucket must share a dtype and a device')
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=first.dtype, device=first.device)
        offset = 0
        for p in self.params:
            p.grad = self.flat[offset:offset + p.numel()].view_as(p)
            offset += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, group=None):
        _, world = _world(group)
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat /= world


def _metric_records(handler, mode):
    """The device records (48-byte ``mg_term_result`` rows) behind the metrics of one collection, in name order."""
    rows = []
    for name in sorted(handler[mode]):
        metric = handler[mode][name]
        record = getattr(metric, '_record', None)
        if record is not None:
            rows.append(record.reshape(1, -1))
        records = getattr(metric, '_records', None)
        if records is not None:
            rows.append(records.reshape(-1, records.shape[-1]))
    return rows


def all_reduce_metrics(handler, mode, extra=None, group=None):
    """Sum the metric state of ``handler[mode]`` over the ranks (one collective) and write the totals back.

    ``extra``: optional 1-D float64 device tensor of further additive values (e.g. ``[loss_sum, n_batches]``) carried in
    the same collective; returned summed.  Integer sums travel as float64 (exact below 2**53).
    """
    rows = _metric_records(handler, mode)
    _, world = _world(group)
    if world == 1 or (not rows and extra is None):
        return extra
    device = rows[0].device if rows else extra.device
    pieces = []
    for r in rows:
        f64, i64 = r.view(torch.float64), r.view(torch.int64)
        pieces.append(torch.stack([f64[:, ops.F64_SUM], f64[:, ops.F64_COUNT], i64[:, ops.I64_ISUM].to(torch.float64)],
                                  dim=1).reshape(-1))
    if extra is not None:
        pieces.append(extra.to(device=device, dtype=torch.float64).reshape(-1))
    packed = torch.cat(pieces)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    offset = 0
    for r in rows:
        n = r.shape[0]
        block = packed[offset:offset + 3 * n].reshape(n, 3)
        offset += 3 * n
        f64, i64, f32 = r.view(torch.float64), r.view(torch.int64), r.view(torch.float32)
        f64[:, ops.F64_SUM], f64[:, ops.F64_COUNT] = block[:, 0], block[:, 1]
        i64[:, ops.I64_ISUM] = block[:, 2].to(torch.int64)
        f32[:, ops.F32_SUM], f32[:, ops.F32_COUNT] = block[:, 0].to(torch.float32), block[:, 1].to(torch.float32)   # fp32 mirrors
    return packed[offset:] if extra is not None else None


class DataParallelTrainer(object):
    """The training / validation loops of ``ExperimentBuilder`` for one process per GPU.

    Parameters
    ----------
    model : torch.nn.Module following the reference's ``BaseModel`` protocol (see module docstring).
    ema_model : a second instance of the model holding the averaged weights, or None.
    ema_decay : float, as ``ExperimentBuilder(ema_decay=...)``; 0 disables the EMA update.
    group : torch.distributed process group (default: the world group; no process group = single process).
    """
    def __init__(self, model, ema_model=None, ema_decay=0., group=None):
        self.model = model
        self.group = group
        self.rank, self.world_size = _world(group)
        self.epoch = 1
        self.bucket = GradientBucket(model.parameters())
        self.ema_decay = ema_decay
        self.ema = None
        if ema_decay:
            if ema_model is None:
                raise ValueError('ema_decay needs an ema_model (a second instance of the model)')
            from morgana_b200.utils import ExponentialMovingAverage
            self.ema = ExponentialMovingAverage(ema_model, ema_decay)

    # ---- experiment_builder.py:431-505 --------------------------------------------------------------------------
    def train_epoch(self, data_loader, optimizer, lr_schedule=None, batch_level_schedule=False, log_every=0, log=None):
        r"""One pass over this rank's batches: zero -> forward -> backward -> gradient all-reduce -> optimiser step ->
        [LR schedule] -> [EMA] -> loss metric.  Returns the epoch's mean loss over all ranks (one read-back, at the end)."""
        model = self.model
        model.mode = 'train'
        model.metrics.reset_state('train')
        loss_sum, n_batches = None, 0
        for i, features in enumerate(data_loader):
            model.step = (self.epoch - 1) * len(data_loader) + i + 1
            self.bucket.zero()
            batch_loss, output_features = model(features)
            batch_loss.backward()
            self.bucket.all_reduce(self.group)
            optimizer.step()
            if lr_schedule is not None and batch_level_schedule:
                lr_schedule.step()
            detached = batch_loss.detach()
            loss_sum = detached.to(torch.float64) if loss_sum is None else loss_sum + detached
            n_batches += 1
            if self.ema is not None:
                self.ema.update_params(model)
            model.metrics.accumulate(model.mode, loss=detached)
            if log is not None and log_every and (i + 1) % log_every == 0:
                log('train', self.epoch, i + 1, detached, model.metrics.results_as_str_dict('train'))
        mean_loss = self._finish_epoch(model, 'train', loss_sum, n_batches)
        model.mode = ''
        return mean_loss

    # ---- experiment_builder.py:562-620 --------------------------------------------------------------------------
    @torch.no_grad()
    def valid_epoch(self, data_loader, model=None, log_every=0, log=None):
        r"""Evaluates ``model`` (default: the trained model; pass ``self.ema.model`` for the averaged one) on this rank's
        batches; metrics accumulate into that model's own handler."""
        if model is None:
            model = self.model
        model.mode = 'valid'
        model.metrics.reset_state('valid')
        loss_sum, n_batches = None, 0
        for i, features in enumerate(data_loader):
            model.step = (self.epoch - 1) * len(data_loader) + i + 1
            batch_loss, output_features = model(features)
            detached = batch_loss.detach()
            loss_sum = detached.to(torch.float64) if loss_sum is None else loss_sum + detached
            n_batches += 1
            model.metrics.accumulate(model.mode, loss=detached)
            if log is not None and log_every and (i + 1) % log_every == 0:
                log('valid', self.epoch, i + 1, detached, model.metrics.results_as_str_dict('valid'))
        mean_loss = self._finish_epoch(model, 'valid', loss_sum, n_batches)
        model.mode = ''
        return mean_loss

    def _finish_epoch(self, model, mode, loss_sum, n_batches):
        if loss_sum is None:
            device = next(model.parameters()).device
            loss_sum = torch.zeros((), dtype=torch.float64, device=device)
        extra = torch.stack([loss_sum.reshape(()), torch.tensor(float(n_batches), dtype=torch.float64, device=loss_sum.device)])
        extra = all_reduce_metrics(model.metrics, mode, extra=extra, group=self.group)
        total, count = extra.tolist()                       # the epoch's one device -> host read
        return total / count if count else float('nan')


def rank_batches(n_items, batch_size, group=None, drop_last=True):
    """Index ranges ``[(begin, end), ...]`` of this rank's batches; see :func:`morgana_b200.dp.sharded_batches`."""
    rank, world = _world(group)
    return dp.sharded_batches(n_items, batch_size, rank, world, drop_last=drop_last)
