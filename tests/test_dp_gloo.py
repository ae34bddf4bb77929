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
