"""K7 backward against what autograd runs today: sigmoid backward + cast + bias sum (ATen) and the cuBLAS bf16 wgrad GEMM."""
import os, sys, torch
sys.path.insert(0, '.')
import morgana_b200 as mg
def timeit(fn, n_iter=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n_iter): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n_iter
M = 256 * 1363
for (K, N, act) in [(600, 512, 'sigmoid'), (512, 128, 'sigmoid'), (128, 32, 'sigmoid'), (32, 1, None), (512, 256, 'sigmoid'), (256, 187, None)]:
    x16 = torch.rand(M, K, device='cuda').to(torch.bfloat16)
    grad = torch.randn(M, N, device='cuda')
    y = torch.rand(M, N, device='cuda') if act else None
    t_act = timeit(lambda: mg.ops.act_grad_bf16(grad, y))
    def aten_act():
        g = grad * (1. - y) * y if act else grad
        return g.to(torch.bfloat16), g.sum(0)
    t_act_ref = timeit(aten_act)
    g16, _ = mg.ops.act_grad_bf16(grad, y)
    os.environ['MG_WGRAD_PAIR'] = '0'
    t_w1 = timeit(lambda: mg.ops.linear_wgrad_bf16(g16, x16, out_features=N, in_features=K))
    os.environ['MG_WGRAD_PAIR'] = '1'
    t_w2 = timeit(lambda: mg.ops.linear_wgrad_bf16(g16, x16, out_features=N, in_features=K))
    del os.environ['MG_WGRAD_PAIR']
    extra = []
    for name, value in (('MG_WGRAD_FRAMES', '64'), ('MG_WGRAD_TILE_K', '128')):
        os.environ[name] = value
        extra.append('%s=%s %.3f' % (name, value, timeit(lambda: mg.ops.linear_wgrad_bf16(g16, x16, out_features=N, in_features=K))))
        del os.environ[name]
    t_w = timeit(lambda: mg.ops.linear_wgrad_bf16(g16, x16, out_features=N, in_features=K))
    gT = g16[:, :N]
    t_w_ref = timeit(lambda: torch.matmul(gT.t(), x16).float())
    err = (mg.ops.linear_wgrad_bf16(g16, x16, out_features=N, in_features=K) - torch.matmul(gT.float().t(), x16.float())).abs().max().item()
    flops = 2.0 * M * N * K
    act_bytes = M * N * (4 + (4 if act else 0) + 2)
    w_bytes = M * (g16.shape[1] + K) * 2
    print('M=%d K=%d N=%d act=%s | K7g %.3f ms (%.2f TB/s) vs ATen %.3f ms | K7w %.3f ms (single %.3f, pair %.3f; %.0f TF/s, %.2f TB/s) vs cuBLAS bf16 %.3f ms | max diff vs fp32 matmul %.3g | %s'
          % (M, K, N, act, t_act, act_bytes / t_act / 1e9, t_act_ref, t_w, t_w1, t_w2, flops / t_w / 1e9, w_bytes / t_w / 1e9, t_w_ref, err, ', '.join(extra)))
