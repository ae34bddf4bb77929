"""Dense layers of the example models on the tcgen05 tensor cores (reference README.rst:65-73, models/RNN_SPSS.py:33-41).

:class:`Linear` is a drop-in for ``torch.nn.Linear`` followed (optionally) by ``torch.nn.Sigmoid``: fp32 master weights
(what the optimiser and the EMA kernel update), a bf16 shadow of the weight refreshed whenever the parameter changes, and
a forward pass that is one ``mg_linear_bf16`` launch with the bias and the sigmoid fused into the epilogue.

Forward tolerance (stated in tests/test_gpu_parity.py): bf16 operands, fp32 accumulation -> <= 2 % of the output range
against the fp32 layer, 2e-3 against the exact product of the bf16-rounded operands.
Backward: the input gradient ``g @ W`` runs through the same tcgen05 kernel for every layer width (``y = g @ (W^T)^T``; the
transposed bf16 copy of the weight is one ``mg_cast_transpose_bf16`` launch from the fp32 master weight); the sigmoid's
backward, the bf16 cast of the gradient and the bias gradient are one pass (``mg_act_grad_bf16``); the weight gradient ``g^T @ x`` reduces over the frame axis: MN-major tcgen05
operands straight from the row-major tensors, frames split over the SMs, slices summed in a fixed order
(``mg_linear_wgrad_bf16``).
"""
import torch

from morgana_b200 import ops


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, weight, bias, weight_bf16, act, out_dtype):
        x_bf16 = ops.cast_pad_bf16(x2d) if x2d.dtype == torch.float32 else x2d
        y = ops.linear_bf16(x_bf16, weight_bf16, bias, act=act, out_dtype=out_dtype)
        ctx.save_for_backward(x_bf16, weight_bf16, y if act == 'sigmoid' else None, weight)
        ctx.act, ctx.k, ctx.has_bias, ctx.x_dtype = act, weight.shape[1], bias is not None, x2d.dtype
        return y

    @staticmethod
    def backward(ctx, grad_y):
        x_bf16, weight_bf16, y, weight = ctx.saved_tensors
        k, n = ctx.k, weight_bf16.shape[0]
        if grad_y.dtype not in (torch.float32, torch.bfloat16):
            grad_y = grad_y.to(torch.float32)
        if grad_y.stride(0) < n:                 # expanded gradients (e.g. of a plain .sum())
            grad_y = grad_y.contiguous()
        # K7g: sigmoid backward, the bf16 operand of both GEMMs (rows padded to 8 columns) and the bias gradient, one pass
        g16, grad_b = ops.act_grad_bf16(grad_y, y if ctx.act == 'sigmoid' else None,
                                        want_bias_grad=ctx.has_bias and ctx.needs_input_grad[2])
        grad_x = None
        if ctx.needs_input_grad[0]:
            # dgrad on the tensor cores, every width: (M, N') @ (N', K) as the forward kernel sees it, x' = g16 (M, N'),
            # w' = W^T (K, N') cast + transposed from the fp32 weight by one small kernel (N' = N rounded up to 8, zero padded)
            w_t = ops.cast_transpose_bf16(weight.detach())
            grad_x = ops.linear_bf16(g16, w_t, None, act=None, out_dtype=torch.float32 if ctx.x_dtype == torch.float32
                                     else torch.bfloat16)
        grad_w = None
        if ctx.needs_input_grad[1]:
            # K7w: g^T @ x over the frame axis, both operands as they lie in memory (MN-major tcgen05 operands)
            grad_w = ops.linear_wgrad_bf16(g16, x_bf16, out_features=n, in_features=k)
        return grad_x, grad_w, grad_b, None, None, None


class Linear(torch.nn.Linear):
    r"""``act(x W^T + b)`` with ``act`` in ``{None, 'sigmoid'}``; input ``(..., in_features)`` fp32 or bf16."""
    def __init__(self, in_features, out_features, bias=True, act=None, out_dtype=torch.float32, device=None):
        super(Linear, self).__init__(in_features, out_features, bias=bias, device=device)
        self.act, self.out_dtype = act, out_dtype
        self._shadow, self._shadow_version = None, None

    def weight_bf16(self):
        """bf16 copy of the weight (rows padded to a multiple of 8).

        Rebuilt on every call in training mode -- fused optimisers update parameters without bumping the autograd version
        counter, so staleness cannot be detected reliably, and the cast is one tiny kernel -- and cached in eval mode until
        the parameter's version or storage changes.
        """
        version = (self.weight._version, self.weight.data_ptr())
        if self.training or self._shadow is None or self._shadow_version != version:
            self._shadow = ops.cast_pad_bf16(self.weight.detach())
            self._shadow_version = version
        return self._shadow

    def forward(self, x):
        lead = x.shape[:-1]
        x2d = x.reshape(-1, x.shape[-1])
        y = _LinearFn.apply(x2d, self.weight, self.bias, self.weight_bf16(), self.act, self.out_dtype)
        return y.reshape(*lead, self.out_features)
