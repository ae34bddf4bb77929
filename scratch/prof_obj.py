import sys, torch
sys.path.insert(0, '.')
from morgana_b200 import workloads
from morgana_b200.fused import AcousticObjective
ling = workloads.linguistic_batch(batch_size=256, seed=1234)
ac = workloads.acoustic_batch(ling['n_frames'], seed=1234)
pred, target, n = ac['pred'].cuda(), ac['target'].cuda(), ling['n_frames'].cuda()
obj = AcousticObjective()
for g in (True, False, True, False):
    l, gr = obj(pred, target, n, want_grad=g)
torch.cuda.synchronize()
print(l.item())
