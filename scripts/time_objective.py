#!/usr/bin/env python
"""Times the fused objective (K4b) at config 2 -- B utterances of the bench's batch -- for tuning experiments through the
MG_OBJ_* / MG_OBJECTIVE_STREAM environment variables.  Prints ms and algorithmic GB/s (valid rows of pred + target read, the
whole gradient written) with and without the gradient."""
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
frames, T = int(ling['n_frames'].sum()), pred.shape[1]
# a second pair of tensors so consecutive launches never find their inputs in L2 (each pair is ~0.5 GB anyway)
pred2, target2 = pred.clone(), target.clone()
obj = AcousticObjective()


def timeit(fn, n_iter=40):
    for _ in range(5):
        fn(0)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(n_iter):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n_iter


def run(i, want_grad=True):
    p, t = (pred, target) if i % 2 == 0 else (pred2, target2)
    return obj(p, t, n, want_grad=want_grad)


g_ms = timeit(lambda i: run(i))
f_ms = timeit(lambda i: run(i, want_grad=False))
g_bytes = 4 * 187 * (2 * frames + B * T)
f_bytes = 4 * 187 * 2 * frames
print('B=%d grad %.4f ms (%.0f GB/s) nograd %.4f ms (%.0f GB/s)  env %s'
      % (B, g_ms, g_bytes / g_ms / 1e6, f_ms, f_bytes / f_ms / 1e6,
         {k: v for k, v in os.environ.items() if k.startswith('MG_OBJ')}))
