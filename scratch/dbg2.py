import sys, time, torch
sys.path.insert(0, '.')
from morgana_b200 import workloads
from morgana_b200.fused import AcousticObjective
ling = workloads.linguistic_batch(batch_size=256, seed=1234)
ac = workloads.acoustic_batch(ling['n_frames'], seed=1234)
pred, target, n = ac['pred'].cuda(), ac['target'].cuda(), ling['n_frames'].cuda()
obj = AcousticObjective()
def timeit(fn, n_iter=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n_iter): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n_iter
print('grad %.4f nograd %.4f' % (timeit(lambda: obj(pred, target, n)), timeit(lambda: obj(pred, target, n, want_grad=False))))
