#!/usr/bin/env python
"""cProfile of the host side of a few public ops at the training batch size (where the Python call, not the kernel, is the cost)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import morgana_b200 as mg                                  # noqa: E402
from morgana_b200 import nn as mnn, workloads              # noqa: E402

B = 32
ling = workloads.linguistic_batch(batch_size=B, seed=1234)
ac = workloads.acoustic_batch(ling['n_frames'], seed=1234)
pred, target, n = ac['pred'].cuda(), ac['target'].cuda(), ling['n_frames'].cuda()
lab, dur, T = ling['lab'].cuda(), ling['dur'].cuda(), int(ling['n_frames'].max())
mmin, mmax = ling['mmin'].cuda(), ling['mmax'].cuda()
layer = mnn.Linear(600, 512, act='sigmoid', out_dtype=torch.bfloat16, device='cuda')
frames = torch.randn(B * T, 600, device='cuda').to(torch.bfloat16)
pg = pred.clone().requires_grad_()
rmse = mg.metrics.RMSE()
rmse.reset_state()

CASES = {
    'upsample': lambda: mg.utils.upsample_to_repetitions(lab, dur, normaliser=('minmax', mmin, mmax), max_len=T),
    'mse_fwd': lambda: mg.losses.mse(pred, target, n),
    'mse_fwd_bwd': lambda: mg.losses.mse(pg, target, n).backward(),
    'rmse': lambda: rmse.accumulate(target, pred, seq_len=n),
    'linear_fwd': lambda: layer(frames),
}
for name, fn in CASES.items():
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    prof = cProfile.Profile()
    prof.enable()
    for _ in range(300):
        fn()
    prof.disable()
    torch.cuda.synchronize()
    print('=' * 30, name, '(300 calls)')
    stats = pstats.Stats(prof, stream=sys.stdout)
    stats.sort_stats('tottime').print_stats(14)
