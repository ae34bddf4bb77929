"""world_size-2 checks of the data-parallel plumbing on CPU (gloo): sharding, record all-reduce, gradient all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import np_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _make_record(total, count, loss):
    rec = torch.zeros(48, dtype=torch.uint8)
    rec.view(torch.float64)[:3] = torch.tensor([total, count, loss], dtype=torch.float64)
    return rec


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from morgana_b200 import dp, workloads
    # every rank generates the full synthetic set and keeps its shard: utterances are the unit of sharding
    ling = workloads.linguistic_batch(batch_size=10, min_phones=4, max_phones=9, max_dur=7, seed=5)
    ac = workloads.acoustic_batch(ling['n_frames'], seed=5)
    lo, hi = dp.shard_range(10, rank, world)
    p, t, n = ac['pred'].numpy()[lo:hi], ac['target'].numpy()[lo:hi], ling['n_frames'].numpy()[lo:hi]
    s, c = O.rmse_acc(t[..., 4:64], p[..., 4:64], n)                 # what the kernels would leave in a record
    loss = O.masked_loss(p[..., 4:184], t[..., 4:184], n)
    packed = dp.allreduce_records(torch.stack([_make_record(s, c, loss)]))
    pending = dp.allreduce_records(torch.stack([_make_record(s, c, loss)]), async_op=True)   # same totals when joined later
    assert torch.equal(pending.result(), packed) and pending.result() is pending.packed
    # gradient all-reduce: rank-dependent gradients must come back as their mean
    model = torch.nn.Linear(3, 2)
    for i, prm in enumerate(model.parameters()):
        prm.grad = torch.full_like(prm, float(rank + 1 + i))
    dp.allreduce_gradients(list(model.parameters()))
    grads = [prm.grad.clone() for prm in model.parameters()]
    if rank == 0:
        torch.save({'packed': packed, 'grads': grads, 'shard': (lo, hi), 'local_loss': loss}, os.path.join(out_dir, 'r0.pt'))
    else:
        torch.save({'shard': (lo, hi), 'local_loss': loss}, os.path.join(out_dir, 'r%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    from morgana_b200 import dp
    for n, world in [(10, 2), (4096, 8), (7, 3), (2, 4)]:
        spans = [dp.shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_two_rank_exchange_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(str(tmp_path), 'r0.pt'))
    r1 = torch.load(os.path.join(str(tmp_path), 'r1.pt'))
    assert r0['shard'] == (0, 5) and r1['shard'] == (5, 10)

    from morgana_b200 import workloads
    ling = workloads.linguistic_batch(batch_size=10, min_phones=4, max_phones=9, max_dur=7, seed=5)
    ac = workloads.acoustic_batch(ling['n_frames'], seed=5)
    p, t, n = ac['pred'].numpy(), ac['target'].numpy(), ling['n_frames'].numpy()
    want_sum, want_count = O.rmse_acc(t[..., 4:64], p[..., 4:64], n)
    total, count, loss = r0['packed'][0].tolist()
    assert count == want_count                                    # counts are integers: exact
    assert total == pytest.approx(want_sum, rel=1e-12)            # sums are additive over shards
    # equal shard sizes: the global loss (mean over all 10 utterances) is the mean of the two rank losses (Q6)
    assert loss == pytest.approx(O.masked_loss(p[..., 4:184], t[..., 4:184], n), rel=1e-12)
    assert loss == pytest.approx(0.5 * (r0['local_loss'] + r1['local_loss']), rel=1e-12)
    for i, g in enumerate(r0['grads']):
        assert torch.equal(g, torch.full_like(g, 1.5 + i))        # mean of (1 + i) and (2 + i)


# ----------------------------------------------------------------------------------------------------------------------
# DataParallelTrainer ("next" row 2): the epoch loops on two gloo ranks against one process on the whole data
# ----------------------------------------------------------------------------------------------------------------------
def test_sharded_batches_equal_steps():
    from morgana_b200 import dp
    for n, bs, world in [(4096, 32, 8), (100, 8, 3), (10, 4, 2)]:
        per_rank = [dp.sharded_batches(n, bs, r, world) for r in range(world)]
        assert len({len(b) for b in per_rank}) == 1                              # same number of steps everywhere
        assert all(e - b == bs for batches in per_rank for b, e in batches)      # equal batch sizes (Q6)
        flat = [i for batches in per_rank for b, e in batches for i in range(b, e)]
        assert len(flat) == len(set(flat))                                       # no utterance is seen twice
        loose = [dp.sharded_batches(n, bs, r, world, drop_last=False) for r in range(world)]
        assert len({len(b) for b in loose}) == 1
        covered = sorted(i for batches in loose for b, e in batches for i in range(b, e))
        assert covered == list(range(n))                                         # nothing dropped
    assert dp.sharded_batches(4096, 32, 3, 8)[0] == (1536, 1568)


def _cpu_mean_metric():
    """A stand-in with the same 48-byte record layout as the device metrics, reduced by torch on the CPU."""
    from morgana_b200 import metrics as M

    class CpuMean(M.StatefulMetric):
        def __init__(self):
            M.StatefulMetric.__init__(self)
            self.reset_state()

        def reset_state(self):
            M.StatefulMetric.reset_state(self)
            self._record = torch.zeros(48, dtype=torch.uint8)

        def accumulate(self, tensor, seq_len=None):
            M.StatefulMetric.accumulate(self)
            f64 = self._record.view(torch.float64)
            f64[0] += tensor.double().sum()
            f64[1] += tensor.numel()

        def result(self):
            f64 = self._record.view(torch.float64)
            return (f64[0] / (f64[1] + 1e-8)).float()
    return CpuMean


class _ToyModel(torch.nn.Module):
    """The reference's BaseModel protocol (base_models.py:27-34, :279-286) in miniature."""
    def __init__(self):
        super().__init__()
        from morgana_b200 import metrics as M
        CpuMean = _cpu_mean_metric()
        torch.manual_seed(11)
        self.net = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 2))
        self.mode, self.step = '', 0
        self.metrics = M.Handler(loss=CpuMean())
        self.metrics.add_metrics('all', abs_err=CpuMean())

    def forward(self, features):
        out = self.net(features['x'])
        self.metrics.accumulate(self.mode, abs_err=((out.detach() - features['y']).abs(),))
        return ((out - features['y']) ** 2).mean(), {'out': out}


def _toy_data(n=24):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, 6, generator=g)
    return x, torch.stack([x[:, :3].sum(1), x[:, 3:].prod(1)], dim=1)


def _run_trainer(rank, world, batch_ranges):
    from morgana_b200 import trainer as T
    x, y = _toy_data()
    model = _ToyModel()
    tr = T.DataParallelTrainer(model)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    batches = [{'x': x[b:e], 'y': y[b:e]} for b, e in batch_ranges]
    train_losses = []
    for tr.epoch in (1, 2):
        train_losses.append(tr.train_epoch(batches, opt))
    train_metrics = model.metrics.results_as_json_dict('train')
    valid_loss = tr.valid_epoch(batches)
    return {'train_losses': train_losses, 'train_metrics': train_metrics, 'valid_loss': valid_loss,
            'valid_metrics': model.metrics.results_as_json_dict('valid'), 'step': model.step,
            'weights': [p.detach().clone() for p in model.parameters()],
            'grad_is_view': all(p.grad.untyped_storage().data_ptr() == tr.bucket.flat.untyped_storage().data_ptr()
                                for p in model.parameters())}


def _trainer_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from morgana_b200 import trainer as T
    result = _run_trainer(rank, world, T.rank_batches(24, 4))
    torch.save(result, os.path.join(out_dir, 'trainer_r%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_trainer_two_ranks_match_one_process(tmp_path):
    world = 2
    mp.spawn(_trainer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(str(tmp_path), 'trainer_r0.pt'))
    r1 = torch.load(os.path.join(str(tmp_path), 'trainer_r1.pt'))
    # one process stepping on the union of the two ranks' batches (rank 0: items 0-11, rank 1: 12-23, 4 per step):
    # the mean of two equal-sized batch means is the mean over the union, so the averaged gradients are the same
    union = [(0, 4), (4, 8), (8, 12)]
    from morgana_b200 import trainer as T
    x, y = _toy_data()
    model = _ToyModel()
    tr = T.DataParallelTrainer(model)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    batches = [{'x': torch.cat([x[b:e], x[b + 12:e + 12]]), 'y': torch.cat([y[b:e], y[b + 12:e + 12]])} for b, e in union]
    losses = []
    for tr.epoch in (1, 2):
        losses.append(tr.train_epoch(batches, opt))
    single_train = model.metrics.results_as_json_dict('train')
    valid_loss = tr.valid_epoch(batches)
    single_valid = model.metrics.results_as_json_dict('valid')

    assert r0['grad_is_view'] and r0['step'] == 6                        # (epoch 2 - 1) * 3 + 3
    for a, b in zip(r0['weights'], r1['weights']):
        assert torch.equal(a, b)                                         # replicas stay identical
    for a, b in zip(r0['weights'], model.parameters()):
        assert torch.allclose(a, b.detach(), rtol=1e-5, atol=1e-6)
    assert r0['train_losses'] == r1['train_losses']                      # every rank reports the global loss
    assert r0['train_losses'] == pytest.approx(losses, rel=1e-5)
    assert r0['valid_loss'] == pytest.approx(valid_loss, rel=1e-5)
    for name in ('loss', 'abs_err'):                                     # metric state summed over ranks, on every rank
        assert r0['train_metrics'][name] == r1['train_metrics'][name]
        assert r0['train_metrics'][name] == pytest.approx(single_train[name], rel=1e-5)
        assert r0['valid_metrics'][name] == pytest.approx(single_valid[name], rel=1e-5)
