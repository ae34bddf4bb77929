import sys, time, torch
sys.path.insert(0, '.')
import morgana_b200 as mg
from morgana_b200 import workloads, ops
from morgana_b200.fused import AcousticObjective
ling = workloads.linguistic_batch(batch_size=256, seed=1234)
ac = workloads.acoustic_batch(ling['n_frames'], seed=1234)
pred, target, n = ac['pred'].cuda(), ac['target'].cuda(), ling['n_frames'].cuda()
lab, dur = ling['lab'].cuda(), ling['dur'].cuda()
mmin, mmax = ling['mmin'].cuda(), ling['mmax'].cuda()
obj = AcousticObjective()
for g in (True, False, True, False):
    l, gr = obj(pred, target, n, want_grad=g)
    torch.cuda.synchronize()
    print('want_grad', g, 'loss', l.item(), obj.last_loss_records.view(torch.float32)[:, 8:12].tolist())
def timeit(fn, n_iter=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); s.record()
    for _ in range(n_iter): fn()
    e.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n_iter, (t1 - t0) / n_iter * 1e3
T = int(ling['n_frames'].max())
print('objective grad   gpu ms %.4f host ms %.4f' % timeit(lambda: obj(pred, target, n)))
print('objective nograd gpu ms %.4f host ms %.4f' % timeit(lambda: obj(pred, target, n, want_grad=False)))
print('terms grad       gpu ms %.4f host ms %.4f' % timeit(lambda: obj.call_with_terms(pred, target, n)))
print('upsample hint    gpu ms %.4f host ms %.4f' % timeit(lambda: mg.utils.upsample_to_repetitions(lab, dur, normaliser=('minmax', mmin, mmax), max_len=T)))
print('upsample sync    gpu ms %.4f host ms %.4f' % timeit(lambda: mg.utils.upsample_to_repetitions(lab, dur, normaliser=('minmax', mmin, mmax))))
print('upsample direct  gpu ms %.4f host ms %.4f' % timeit(lambda: mg.utils.upsample_to_repetitions(lab, dur, normaliser=('minmax', mmin, mmax), max_len=T, path='direct')))
print('dur_scan         gpu ms %.4f host ms %.4f' % timeit(lambda: ops.dur_scan(dur)))
print('mse mcep slice   gpu ms %.4f host ms %.4f' % timeit(lambda: mg.losses.mse(pred[..., 4:184], target[..., 4:184], n)))
pc, tc = pred[..., 4:184].contiguous(), target[..., 4:184].contiguous()
print('mse mcep contig  gpu ms %.4f host ms %.4f' % timeit(lambda: mg.losses.mse(pc, tc, n)))
print('mse full 187     gpu ms %.4f host ms %.4f' % timeit(lambda: mg.losses.mse(pred, target, n)))
x = torch.randn(256 * T, 187, device='cuda'); m_, s_ = torch.randn(187, device='cuda'), torch.rand(187, device='cuda') + .1
print('denorm mvn       gpu ms %.4f host ms %.4f (bytes %.0f MB)' % (timeit(lambda: mg.data.denormalise_mvn(x, m_, s_)) + (x.numel() * 8 / 1e6,)))
params = [torch.randn(s, device='cuda') for s in [609 * 512, 512] + [4 * 512 * 512, 4 * 512 * 512, 2048, 2048] * 8 + [512 * 256, 256, 256 * 187, 187]]
shadow = [torch.randn_like(p) for p in params]
pairs = list(zip(shadow, params)); plan = ops.EmaPlan()
npar = sum(p.numel() for p in params)
print('ema %d params  gpu ms %.4f host ms %.4f (bytes %.0f MB)' % ((npar,) + timeit(lambda: ops.ema_update(pairs, 0.001, plan=plan)) + (npar * 12 / 1e6,)))
gout = torch.randn(256, T, 600, device='cuda')
labg = lab.clone().requires_grad_()
def bwd():
    out = mg.utils.upsample_to_repetitions(labg, dur, max_len=T)
    out.backward(gout)
print('upsample fwd+bwd gpu ms %.4f host ms %.4f' % timeit(bwd, 20))
