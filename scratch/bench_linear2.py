import sys, torch
sys.path.insert(0, '.')
import morgana_b200 as mg
def timeit(fn, n_iter=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n_iter): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n_iter
M, K, N = 256 * 1363, 600, 512
x = torch.rand(M, K, device='cuda').to(torch.bfloat16)
w = (torch.randn(N, K, device='cuda') / K ** 0.5).to(torch.bfloat16)
b = torch.randn(N, device='cuda') * 0.1
print('f32 %.3f bf16 %.3f' % (timeit(lambda: mg.ops.linear_bf16(x, w, b, act='sigmoid')), timeit(lambda: mg.ops.linear_bf16(x, w, b, act='sigmoid', out_dtype=torch.bfloat16))))
