import sys, torch
sys.path.insert(0, '.')
import morgana_b200 as mg
M, K, N = 256 * 1363, 600, 512
x = torch.rand(M, K, device='cuda').to(torch.bfloat16)
w = (torch.randn(N, K, device='cuda') / K ** 0.5).to(torch.bfloat16)
b = torch.randn(N, device='cuda') * 0.1
for _ in range(3):
    y = mg.ops.linear_bf16(x, w, b, act='sigmoid')
torch.cuda.synchronize()
print(float(y.float().mean()))
