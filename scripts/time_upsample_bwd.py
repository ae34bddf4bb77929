"""K2 backward at config 2: the row-streaming kernel against the column-strip kernel (MG_UPSAMPLE_BWD_ROWS=0 in a fresh process)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import morgana_b200 as mg
from morgana_b200 import workloads
ling = workloads.linguistic_batch(batch_size=256, seed=1234)
lab, dur = ling['lab'].cuda().requires_grad_(), ling['dur'].cuda()
norm = ('minmax', ling['mmin'].cuda(), ling['mmax'].cuda())
out = mg.utils.upsample_to_repetitions(lab, dur, normaliser=norm)
up = torch.randn(out.shape, device='cuda', generator=torch.Generator(device='cuda').manual_seed(3))
def bwd():
    return torch.autograd.grad(out, lab, up, retain_graph=True)[0]
for _ in range(5): g = bwd()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): bwd()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 50
nbytes = 4 * 600 * (int(ling['n_frames'].sum()) + lab.shape[0] * lab.shape[1])   # valid rows of grad_out read, every item row written
print('rows_form=%s backward %.4f ms, %.0f GB/s, checksum %.6f' % (os.environ.get('MG_UPSAMPLE_BWD_ROWS', '1'), ms, nbytes / ms / 1e6, float(g.double().sum())))
