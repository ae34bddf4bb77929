"""losses.mse on the 180-of-187 column slice the reference's LSTMAcousticModel.loss takes (models/RNN_SPSS.py:134), config-3 scale.
MG_RED_SLICE_MODE: 0 thread-per-column, 1 flat masked stream in the forward pass (default)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import morgana_b200 as mg
from morgana_b200 import workloads
n = workloads.acoustic_lengths(batch_size=1024, min_frames=300, max_frames=1200, seed=1234)
ac = workloads.acoustic_batch(n, seed=1234)
pred, target, n = ac['pred'].cuda().requires_grad_(), ac['target'].cuda(), n.cuda()
def timeit(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
F = int(n.sum())
for mode in ('0', '1'):
    os.environ['MG_RED_SLICE_MODE'] = mode
    fwd = timeit(lambda: mg.losses.mse(pred[..., 4:184], target[..., 4:184], n))
    def both():
        sl = pred[..., 4:184]
        loss = mg.losses.mse(sl, target[..., 4:184], n)
        return torch.autograd.grad(loss, sl)[0]
    fb = timeit(both)
    val = mg.losses.mse(pred[..., 4:184], target[..., 4:184], n).item()
    print('mode %s: forward %.3f ms (%.0f GB/s of slice bytes), forward + backward w.r.t. the slice %.3f ms, loss %.8f'
          % (mode, fwd, 8 * 180 * F / fwd / 1e6, fb, val))
