import sys, torch
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
    x = torch.rand(M, K, device='cuda').to(torch.bfloat16)
    w = (torch.randn(N, K, device='cuda') / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device='cuda') * 0.1
    ours32 = timeit(lambda: mg.ops.linear_bf16(x, w, b, act=act))
    ours16 = timeit(lambda: mg.ops.linear_bf16(x, w, b, act=act, out_dtype=torch.bfloat16))
    b16 = b.to(torch.bfloat16)
    def ref():
        y = torch.nn.functional.linear(x, w, b16)
        return torch.sigmoid(y) if act else y
    ref16 = timeit(ref)
    flops = 2.0 * M * N * K
    bytes32 = M * K * 2 + N * K * 2 + M * N * 4
    bytes16 = M * K * 2 + N * K * 2 + M * N * 2
    print('M=%d K=%d N=%d act=%s | ours f32-out %.3f ms (%.0f TF/s, %.2f TB/s) | ours bf16-out %.3f ms (%.0f TF/s, %.2f TB/s) | cuBLAS bf16 + sigmoid %.3f ms'
          % (M, K, N, act, ours32, flops / ours32 / 1e9, bytes32 / ours32 / 1e9, ours16, flops / ours16 / 1e9, bytes16 / ours16 / 1e9, ref16))
