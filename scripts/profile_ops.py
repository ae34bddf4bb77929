#!/usr/bin/env python
"""One launch of every kernel at its benchmark shape, for `ncu --set full` (see profiles/)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import morgana_b200 as mg                          # noqa: E402
from morgana_b200 import ops, workloads           # noqa: E402

dev = torch.device('cuda', 0)
M = 256 * 1363
x = torch.rand(M, 600, device=dev).to(torch.bfloat16)
w1 = (torch.randn(512, 600, device=dev) / 600 ** 0.5).to(torch.bfloat16)
b1 = torch.randn(512, device=dev) * 0.1
h = torch.rand(M, 512, device=dev).to(torch.bfloat16)
w2 = (torch.randn(128, 512, device=dev) / 512 ** 0.5).to(torch.bfloat16)
b2 = torch.randn(128, device=dev) * 0.1
n3 = workloads.acoustic_lengths(batch_size=1024, seed=1234)
ac = workloads.acoustic_batch(n3, seed=1234)
pred, target, n3d = ac['pred'].to(dev), ac['target'].to(dev), n3.to(dev)
shapes = [(512, 609), (512,)] + [(2048, 512), (2048, 512), (2048,), (2048,)] * 8 + [(256, 512), (256,), (187, 256), (187,)]
params = [torch.randn(*s, device=dev) for s in shapes]
shadow = [torch.randn(*s, device=dev) for s in shapes]
mean, std = torch.randn(187, device=dev), torch.rand(187, device=dev) + 0.1
# K8 at the config-4 shape: the mcep stream of 32 utterances (60 static dims), padding 100
l32 = workloads.linguistic_batch(batch_size=32, seed=1234)
n32, T32 = l32['n_frames'].to(dev), int(l32['n_frames'].max())
mlpg_means = torch.randn(32, T32, 180, device=dev)
mlpg_var = torch.rand(180, device=dev) + 0.3
for _ in range(2):
    ops.mlpg(mlpg_means, mlpg_var, padding_size=100, seq_len=n32)
    ops.linear_bf16(x, w1, b1, act='sigmoid', out_dtype=torch.bfloat16)
    ops.linear_bf16(h, w2, b2, act='sigmoid', out_dtype=torch.bfloat16)
    mg.losses.mse(pred, target, n3d)
    ops.ema_update(list(zip(shadow, params)), 0.001)
    mg.data.denormalise_mvn(pred, mean, std)
torch.cuda.synchronize()
print('ok')
