#!/usr/bin/env python
"""A handful of launches of the fused objective (K4b) at config 2, for `ncu -k regex:objective`."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morgana_b200 import workloads                      # noqa: E402
from morgana_b200.fused import AcousticObjective        # noqa: E402
B = int(os.environ.get('B', '256'))
ling = workloads.linguistic_batch(batch_size=B, seed=1234)
ac = workloads.acoustic_batch(ling['n_frames'], seed=1234)
pred, target, n = ac['pred'].cuda(), ac['target'].cuda(), ling['n_frames'].cuda()
obj = AcousticObjective()
for _ in range(4):
    loss, grad = obj(pred, target, n)
torch.cuda.synchronize()
print('loss', float(loss))
