#!/usr/bin/env python
"""Host time per call of the public ops at the training batch size (32 utterances): what an eager Python loop pays per op
when the kernels are shorter than the call.  `host us` = wall time per call of a back-to-back loop without
synchronisation (the GPU keeps up at this size, so this is the enqueue cost); `gpu us` = CUDA events over the same loop."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import morgana_b200 as mg                                  # noqa: E402
from morgana_b200 import ops, workloads                    # noqa: E402
from morgana_b200.fused import AcousticObjective          # noqa: E402

B = int(os.environ.get('B', '32'))
ling = workloads.linguistic_batch(batch_size=B, seed=1234)
ac = workloads.acoustic_batch(ling['n_frames'], seed=1234)
pred, target, n = ac['pred'].cuda(), ac['target'].cuda(), ling['n_frames'].cuda()
lab, dur, T = ling['lab'].cuda(), ling['dur'].cuda(), int(ling['n_frames'].max())
mmin, mmax = ling['mmin'].cuda(), ling['mmax'].cuda()
mean, std = torch.randn(187, device='cuda'), torch.rand(187, device='cuda') + 0.1


def timeit(fn, n_iter=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    for _ in range(n_iter):
        fn()
    e.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return round((t1 - t0) / n_iter * 1e6, 1), round(s.elapsed_time(e) / n_iter * 1e3, 1)


obj = AcousticObjective()
rmse, lf0 = mg.metrics.RMSE(), mg.metrics.LF0Distortion()
rmse.reset_state()
lf0.reset_state()
voiced = pred[..., 3:4] > 0.5
pg = pred.clone().requires_grad_()
model = torch.nn.Sequential(torch.nn.Linear(609, 512), torch.nn.LSTM(512, 512), torch.nn.Linear(512, 187)).cuda()
ema_model = torch.nn.Sequential(torch.nn.Linear(609, 512), torch.nn.LSTM(512, 512), torch.nn.Linear(512, 187)).cuda()
ema = mg.utils.ExponentialMovingAverage(ema_model, 0.999)
layer = mg.nn.Linear(600, 512, act='sigmoid', device='cuda')
frames16 = torch.randn(B * T, 600, device='cuda').to(torch.bfloat16)


def mse_fwd_bwd():
    pg.grad = None
    mg.losses.mse(pg, target, n).backward()


rows = {}
for name, fn in [
        ('utils.upsample_to_repetitions (max_len hint)', lambda: mg.utils.upsample_to_repetitions(lab, dur, normaliser=('minmax', mmin, mmax), max_len=T)),
        ('utils.upsample_to_repetitions (reference signature, 32-byte read-back)', lambda: mg.utils.upsample_to_repetitions(lab, dur)),
        ('data.denormalise_mvn', lambda: mg.data.denormalise_mvn(pred, mean, std)),
        ('losses.mse forward', lambda: mg.losses.mse(pred, target, n)),
        ('losses.mse forward + backward', mse_fwd_bwd),
        ('metrics.RMSE.accumulate', lambda: rmse.accumulate(target, pred, seq_len=n)),
        ('metrics.LF0Distortion.accumulate', lambda: lf0.accumulate(target[..., 0:1], pred[..., 0:1], voiced, seq_len=n)),
        ('fused.AcousticObjective (loss + gradient + 4 metrics)', lambda: obj(pred, target, n)),
        ('ExponentialMovingAverage.update_params (8 tensors)', lambda: ema.update_params(model)),
        ('nn.Linear 600 -> 512 + sigmoid forward (bf16 frames)', lambda: layer(frames16)),
        ('torch.empty + one trivial ATen kernel (yardstick)', lambda: torch.add(mean, std))]:
    rows[name] = dict(zip(('host_us', 'gpu_us'), timeit(fn)))
print(json.dumps({'batch_utterances': B, 'ops': rows}, indent=1))
