"""One launch each of K7g / K7w at the README MLP's first layer (600 -> 512, config-2 frame count) for ncu."""
import sys, torch
sys.path.insert(0, '.')
import morgana_b200 as mg
M, K, N = 256 * 1363, 600, 512
x16 = torch.rand(M, K, device='cuda').to(torch.bfloat16)
grad = torch.randn(M, N, device='cuda')
y = torch.rand(M, N, device='cuda')
for _ in range(3):
    g16, db = mg.ops.act_grad_bf16(grad, y)
    dw = mg.ops.linear_wgrad_bf16(g16, x16, out_features=N, in_features=K)
torch.cuda.synchronize()
print('ok', float(dw.abs().max()), float(db.abs().max()))
