#!/usr/bin/env python
"""Per-CTA timeline of the persistent objective kernel (MG_OBJ_DEBUG=8 stamps), config 2."""
import os
import sys
os.environ['MG_OBJ_DEBUG'] = str(int(os.environ.get('MG_OBJ_DEBUG', '0')) | 8)
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morgana_b200 import ops, workloads                 # noqa: E402
from morgana_b200.fused import AcousticObjective        # noqa: E402
B = int(os.environ.get('B', '256'))
ling = workloads.linguistic_batch(batch_size=B, seed=1234)
ac = workloads.acoustic_batch(ling['n_frames'], seed=1234)
pred, target, n = ac['pred'].cuda(), ac['target'].cuda(), ling['n_frames'].cuda()
obj = AcousticObjective()
for _ in range(3):
    obj(pred, target, n)
torch.cuda.synchronize()
ws = list(ops._workspaces.values())[0]
off = 256 + 65536 * 4 + 512 * 1024
ws[off:off + 1024 * 128].zero_()
torch.cuda.synchronize()
obj(pred, target, n)
torch.cuda.synchronize()
st = ws[off:off + 1024 * 128].cpu().numpy().view(np.uint64).reshape(1024, 16)
st = st[st[:, 0] > 0].astype(np.int64)
t0 = st[:, 0].min()
rel = (st[:, :5] - t0) / 1e3
print('CTAs', len(st), 'stages per CTA min/mean/max', st[:, 5].min(), st[:, 5].mean(), st[:, 5].max())
for i, name in enumerate(['entry', 'static job published', 'first stage ready', 'stream done', 'ticket taken']):
    col = rel[:, i]
    print('%-18s us: min %7.2f  p50 %7.2f  p90 %7.2f  max %7.2f' % (name, col.min(), np.median(col), np.percentile(col, 90), col.max()))
valid, pad = st[:, 7].astype(float), (st[:, 5] - st[:, 7]).astype(float)
A = np.stack([valid, pad, np.ones_like(valid)], axis=1)
coef, res, _, _ = np.linalg.lstsq(A, dur_all := (rel[:, 3] - rel[:, 1]), rcond=None)
print('fit: stream time (us) = %.4f * loaded stages + %.4f * pad stages + %.2f ; residual std %.2f us' % (coef[0], coef[1], coef[2], np.std(dur_all - A @ coef)))
print('loaded stages per CTA min/mean/max', valid.min(), valid.mean(), valid.max(), ' pad', pad.min(), pad.mean(), pad.max())
smid = np.arange(len(st)) % 148
fin = st[:, 6].max()
print('result records written at %.2f us' % ((fin - t0) / 1e3))
print('jobs per CTA min/mean/max', st[:, 12].min(), st[:, 12].mean(), st[:, 12].max(), ' static job done us p50 %.2f max %.2f' % (np.median((st[:, 13] - t0) / 1e3), ((st[:, 13] - t0) / 1e3).max()))
pro = (st[:, 8:10] - st[:, 0:1]) / 1e3
print('prologue: loads + first barrier p50 %.2f us, scan + second barrier p50 %.2f us' % (np.median(pro[:, 0]), np.median(pro[:, 1])))
last = int(np.argmax(st[:, 6]))
print('last CTA %d: stream done %.2f, ticket taken %.2f, is-last known %.2f, partials loaded %.2f, records %.2f'
      % (last, rel[last, 3], rel[last, 4], (st[last, 10] - t0) / 1e3, (st[last, 11] - t0) / 1e3, (st[last, 6] - t0) / 1e3))
dur = rel[:, 3] - rel[:, 2]
print('stream duration per CTA us: min %.2f p50 %.2f max %.2f' % (dur.min(), np.median(dur), dur.max()))
order = np.argsort(rel[:, 3])
print('earliest finishers (cta, stages, done us):', [(int(i), int(st[i, 5]), round(float(rel[i, 3]), 1)) for i in order[:5]])
print('latest finishers:', [(int(i), int(st[i, 5]), round(float(rel[i, 3]), 1)) for i in order[-5:]])

smid = st[:, 14]
done = rel[:, 3]
import collections
by_sm = collections.defaultdict(list)
for i in range(len(st)):
    by_sm[int(smid[i])].append((i, float(done[i])))
pairs = [v for v in by_sm.values() if len(v) == 2]
if pairs:
    a = np.array([p[0][1] for p in pairs]); b = np.array([p[1][1] for p in pairs])
    print('SMs with two CTAs: %d; correlation of the two finishing times on one SM: %.3f; mean |diff| %.2f us' % (len(pairs), np.corrcoef(a, b)[0, 1], np.abs(a - b).mean()))
np.save(os.environ.get('STAMP_OUT', 'gpurun_out/stamps.npy'), st)

def rel_us(i):
    return (st[:, i] - st[:, 0]) / 1e3
for i, name in ((12, 'producer: head loads issued'), (8, 'first CTA barrier'), (13, 'producer: prefix done'), (1, 'producer: range published'), (9, 'consumer tid 0: set up, about to wait'), (2, 'consumer tid 0: first stage in hand')):
    v = rel_us(i)
    print('%-40s us after own entry: min %6.2f p50 %6.2f max %6.2f' % (name, v.min(), np.median(v), v.max()))

last = int(np.argmax(st[:, 6]))
g = lambda i: (st[last, i] - t0) / 1e3
print('last CTA %d: stream done %.2f | ticket %.2f | is-last known %.2f | records loaded (warp 0, first slot) %.2f | warp sums %.2f | all done %.2f'
      % (last, g(3), g(4), g(10), g(11), g(15), g(6)))
