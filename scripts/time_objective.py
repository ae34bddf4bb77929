#!/usr/bin/env python
"""Times the fused objective at config 2 (used for tuning experiments through MG_OBJ_* environment variables)."""
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


def timeit(fn, n_iter=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n_iter):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n_iter


print('grad %.4f nograd %.4f' % (timeit(lambda: obj(pred, target, n)), timeit(lambda: obj(pred, target, n, want_grad=False))))
