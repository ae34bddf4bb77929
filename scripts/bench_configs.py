#!/usr/bin/env python
"""Per-configuration timings on one B200 (BASELINE.json configs 1-4): the new kernels next to the reference's own op
chain run on CUDA tensors ("stock ATen on the same GPU").  Prints one JSON line per row.

    python scripts/bench_configs.py            # needs a GPU
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import morgana_b200 as mg                                    # noqa: E402
from morgana_b200 import nn as mnn, ops, workloads          # noqa: E402
from morgana_b200.fused import AcousticObjective            # noqa: E402
from oracle import aten_chain as ref                        # noqa: E402  (baseline leg only)

PEAK = 6549.8
if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')):
    PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']


def timeit(fn, n_iter=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n_iter):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n_iter


def row(config, what, ms, alg_bytes=None, frames=None, baseline_ms=None, note=None):
    out = {'config': config, 'op': what, 'ms': round(ms, 4)}
    if alg_bytes is not None:
        out['algorithmic_GB_per_s'] = round(alg_bytes / ms / 1e6, 1)
        out['frac_of_measured_hbm_peak'] = round(alg_bytes / ms / 1e6 / PEAK, 3)
        out['frac_of_8000'] = round(alg_bytes / ms / 1e6 / 8000., 3)
    if frames is not None:
        out['valid_frames_per_s'] = round(frames / ms * 1e3)
    if baseline_ms is not None:
        out['stock_aten_on_gpu_ms'] = round(baseline_ms, 4)
        out['speedup_vs_stock_aten'] = round(baseline_ms / ms, 1)
    if note:
        out['note'] = note
    print(json.dumps(out), flush=True)


def main():
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)

    # ---- config 2: 256 utterances, 600-dim labels ---------------------------------------------------------------
    ling = workloads.linguistic_batch(batch_size=256, seed=1234)
    lab, dur = ling['lab'].to(dev), ling['dur'].to(dev)
    mmin, mmax = ling['mmin'].to(dev), ling['mmax'].to(dev)
    B, P, D = lab.shape
    T, F = int(ling['n_frames'].max()), int(ling['n_frames'].sum())
    n_items = int(ling['n_phones'].sum())
    k2_bytes = 4 * D * (B * T + n_items) + 4 * B * P + 8 * D
    stock = timeit(lambda: ref.upsample_chain(ref.normalise_minmax_chain(lab, mmin, mmax), dur), n_iter=5, warmup=1)
    ms = timeit(lambda: mg.utils.upsample_to_repetitions(lab, dur, normaliser=('minmax', mmin, mmax), max_len=T))
    row('C2', 'K1+K2 fused minmax-normalise + upsample (bulk path, max_len hint)', ms, k2_bytes, F, stock)
    ms = timeit(lambda: mg.utils.upsample_to_repetitions(lab, dur, normaliser=('minmax', mmin, mmax)))
    row('C2', 'K1+K2 through the reference signature (32-byte read-back sizes the output)', ms, k2_bytes, F, stock)
    ms = timeit(lambda: mg.utils.upsample_to_repetitions(lab, dur, normaliser=('minmax', mmin, mmax), max_len=T, path='direct'))
    row('C2', 'K1+K2 direct path (register-staged vector stores)', ms, k2_bytes, F, stock)
    ms = timeit(lambda: mg.utils.upsample_to_repetitions(lab, dur, normaliser=('minmax', mmin, mmax), max_len=T,
                                                         out_dtype=torch.bfloat16))
    row('C2', 'K1+K2 with bf16 frames', ms, 2 * D * B * T + 4 * D * n_items + 4 * B * P, F, stock)
    grad_out = torch.randn(B, T, D, device=dev)
    lab_g = lab.clone().requires_grad_()

    def fwd_bwd():
        mg.utils.upsample_to_repetitions(lab_g, dur, max_len=T).backward(grad_out)

    lab_r = lab.clone().requires_grad_()

    def ref_fwd_bwd():
        ref.upsample_chain(lab_r, dur).backward(grad_out)
    row('C2', 'upsample forward + backward (deterministic segment sum)', timeit(fwd_bwd, 10), 4 * D * (B * T + F) + 8 * D * n_items, F,   # the backward reads valid rows only
        timeit(ref_fwd_bwd, 5, 1))
    del grad_out, lab_g, lab_r

    # ---- config 3: 1024 utterances, 187-dim targets ---------------------------------------------------------------
    n3 = workloads.acoustic_lengths(batch_size=1024, seed=1234)
    ac = workloads.acoustic_batch(n3, seed=1234)
    pred, target, voiced, n3d = ac['pred'].to(dev), ac['target'].to(dev), ac['voiced'].to(dev), n3.to(dev)
    F3, T3 = int(n3.sum()), int(n3.max())
    stock = timeit(lambda: ref.mse_chain(pred, target, n3d), 5, 1)
    row('C3', 'losses.mse forward, (1024, T, 187)', timeit(lambda: mg.losses.mse(pred, target, n3d)), 8 * 187 * F3, F3, stock)
    pg = pred.clone().requires_grad_()

    def mse_fb():
        pg.grad = None
        mg.losses.mse(pg, target, n3d).backward()

    pr = pred.clone().requires_grad_()

    def ref_mse_fb():
        pr.grad = None
        ref.mse_chain(pr, target, n3d).backward()
    row('C3', 'losses.mse forward + backward', timeit(mse_fb, 10), 16 * 187 * F3 + 4 * 187 * 1024 * T3, F3, timeit(ref_mse_fb, 5, 1))
    rmse = mg.metrics.RMSE()
    rmse.reset_state()
    stock = timeit(lambda: ref.rmse_increment(target, pred, n3d), 5, 1)
    row('C3', 'metrics.RMSE.accumulate, 187 dims', timeit(lambda: rmse.accumulate(target, pred, seq_len=n3d)), 8 * 187 * F3, F3, stock)
    lf0 = mg.metrics.LF0Distortion()
    vuv = pred[..., 3:4] > 0.5
    stock = timeit(lambda: ref.lf0_increment(target[..., 0:1], pred[..., 0:1], vuv, n3d), 5, 1)
    row('C3', 'metrics.LF0Distortion.accumulate (B, T, 1)', timeit(lambda: lf0.accumulate(target[..., 0:1], pred[..., 0:1], vuv, seq_len=n3d)),
        None, F3, stock)
    objective = AcousticObjective()
    stock = timeit(lambda: ref.acoustic_loss_and_metrics(pred, target, voiced, n3d), 3, 1)
    row('C3', 'whole objective of models/RNN_SPSS.py:120-139 (3 mse + bce + grad + 4 metrics), ONE launch',
        timeit(lambda: objective(pred, target, n3d)), 8 * 187 * F3 + 4 * 187 * 1024 * T3, F3, stock)
    del pred, target, voiced, pg, pr, ac

    # ---- config 4: EMA of LSTMAcousticModel-sized parameters (187 outputs) ---------------------------------------------
    shapes = [(512, 609), (512,)] + [(2048, 512), (2048, 512), (2048,), (2048,)] * 8 + [(256, 512), (256,), (187, 256), (187,)]
    params = [torch.randn(*s, device=dev) for s in shapes]
    shadow = [torch.randn(*s, device=dev) for s in shapes]
    n_par = sum(p.numel() for p in params)
    plan = ops.EmaPlan()
    pairs = list(zip(shadow, params))
    stock = timeit(lambda: ref.ema_chain(shadow, params, 0.999), 10, 2)
    row('C4', 'EMA update, %d parameters in %d tensors' % (n_par, len(params)), timeit(lambda: ops.ema_update(pairs, 0.001, plan=plan)),
        12 * n_par, None, stock)

    # ---- "next" row 1: MLPG inside predict() (models/RNN_SPSS.py:108-118): mcep stream, 32 utterances, padding 100 ---------
    import time
    import numpy as np
    import scipy.linalg as sl
    from morgana_b200.viz.synthesis import MLPG
    l32 = workloads.linguistic_batch(batch_size=32, seed=1234)
    n32, T32, Fm = l32['n_frames'], int(l32['n_frames'].max()), 60
    g = torch.Generator().manual_seed(5)
    means = torch.randn(32, T32, 3 * Fm, generator=g)
    var = torch.rand(3 * Fm, generator=g) + 0.3
    means_d, var_d, n32_d = means.to(dev), var.to(dev), n32.to(dev)
    ms = timeit(lambda: MLPG(means_d, var_d, padding_size=100, seq_len=n32_d), 10)

    def cpu_mlpg():      # the reference's loop (synthesis.py:153-171) with scipy's banded Cholesky in place of bandmat
        pad, out = 100, np.zeros((32, T32, Fm))
        mu_all, tau = means.numpy().astype(np.float64), 1. / var.numpy().astype(np.float64)
        for i in range(32):
            n = int(n32[i])
            L = n + 2 * pad
            mu = np.pad(mu_all[i, :n], ((pad, pad), (0, 0)), mode='edge')
            for d in range(Fm):
                t0, t1, t2 = tau[d], tau[Fm + d], tau[2 * Fm + d]
                b0, b1, b2 = mu[:, d] * t0, mu[:, Fm + d] * t1, mu[:, 2 * Fm + d] * t2
                b = b0 - 2 * b2
                b[1:] += 0.5 * b1[:-1] + b2[:-1]
                b[:-1] += -0.5 * b1[1:] + b2[1:]
                ab = np.zeros((3, L))
                ab[2] = t0 + 4 * t2
                ab[2, 1:] += 0.25 * t1 + t2
                ab[2, :-1] += 0.25 * t1 + t2
                ab[1, 1:] = -4 * t2
                ab[0, 2:] = -0.25 * t1 + t2
                out[i, :n, d] = sl.solveh_banded(ab, b)[pad:L - pad]
        return out
    t0 = time.perf_counter()
    want = cpu_mlpg()
    cpu_ms = (time.perf_counter() - t0) * 1e3
    err = float(np.abs(MLPG(means_d, var_d, padding_size=100, seq_len=n32_d).cpu().numpy() - want).max())
    print(json.dumps({'config': 'C4 predict()', 'op': 'MLPG, mcep stream: 32 utterances x 60 dims, padding 100 (1920 banded solves)',
                      'ms': round(ms, 4), 'cpu_scipy_banded_ms': round(cpu_ms, 1), 'speedup_vs_cpu_loop': round(cpu_ms / ms, 1),
                      'max_abs_difference_vs_cpu': err,
                      'note': 'CPU leg = the reference loop with scipy.linalg.solveh_banded for bandmat (absent); excludes the '
                              'reference\'s device<->host copies'}), flush=True)

    # ---- config 1 shapes at config-2 scale: README MLP forward on the frame-rate features ------------------------------
    dims = [600, 512, 128, 32, 1]
    torch.manual_seed(0)
    layers = [mnn.Linear(dims[i], dims[i + 1], act='sigmoid' if i < 3 else None,
                         out_dtype=torch.bfloat16 if i < 3 else torch.float32, device=dev) for i in range(4)]
    stock_layers = torch.nn.Sequential(*[m for i in range(4) for m in
                                         ([torch.nn.Linear(dims[i], dims[i + 1])] + ([torch.nn.Sigmoid()] if i < 3 else []))]).to(dev)
    lf0_mean, lf0_std = torch.tensor([5.0], device=dev), torch.tensor([0.3], device=dev)
    tgt = torch.randn(B, T, 1, device=dev)
    n_frames = ling['n_frames'].to(dev)

    def ours():
        with torch.no_grad():
            h = mg.utils.upsample_to_repetitions(lab, dur, normaliser=('minmax', mmin, mmax), max_len=T, out_dtype=torch.bfloat16)
            h = h.reshape(B * T, 600)
            for layer in layers:
                h = ops.linear_bf16(h, layer.weight_bf16(), layer.bias, act=layer.act, out_dtype=layer.out_dtype)
            pred_norm = h.reshape(B, T, 1)
            pred_lf0 = mg.data.denormalise_mvn(pred_norm, lf0_mean, lf0_std)
            return mg.losses.mse(pred_norm, tgt, n_frames), pred_lf0

    def ours_commuted():
        """Layer 1 commutes with the expansion (every frame row is a copy of a phone row): run it at PHONE rate, then expand
        its 512-dim bf16 activations.  Same bits on every valid frame; padding frames hold 0 instead of sigmoid(bias)."""
        with torch.no_grad():
            P = lab.shape[1]
            x = mg.data.normalise_minmax(lab, mmin, mmax).reshape(B * P, 600)
            h = ops.linear_bf16(x, layers[0].weight_bf16(), layers[0].bias, act='sigmoid', out_dtype=torch.bfloat16)
            h = mg.utils.upsample_to_repetitions(h.reshape(B, P, 512), dur, max_len=T).reshape(B * T, 512)
            for layer in layers[1:]:
                h = ops.linear_bf16(h, layer.weight_bf16(), layer.bias, act=layer.act, out_dtype=layer.out_dtype)
            pred_norm = h.reshape(B, T, 1)
            pred_lf0 = mg.data.denormalise_mvn(pred_norm, lf0_mean, lf0_std)
            return mg.losses.mse(pred_norm, tgt, n_frames), pred_lf0

    def stock_chain():
        with torch.no_grad():
            h = ref.upsample_chain(ref.normalise_minmax_chain(lab, mmin, mmax), dur)
            pred_norm = stock_layers(h)
            pred_lf0 = ref.denormalise_mvn_chain(pred_norm, lf0_mean, lf0_std)
            return ref.mse_chain(pred_norm, tgt, n_frames), pred_lf0
    flops = 2.0 * B * T * sum(dims[i] * dims[i + 1] for i in range(4))
    ms = timeit(ours, 10)
    stock = timeit(stock_chain, 5, 1)
    row('C1@C2', 'README F0 MLP predict + loss: fused normalise/upsample (bf16) -> 4 tcgen05 layers -> denormalise -> mse', ms,
        None, F, stock, note='%.0f TFLOP/s over the four layers incl. the feature path; stock = fp32 cuBLAS + ATen' % (flops / ms / 1e9))
    (loss_a, lf0_a), (loss_b, lf0_b) = ours(), ours_commuted()
    valid = (torch.arange(T, device=dev)[None] < n_frames[:, None])[:, :, None]
    same = bool(torch.equal(lf0_a[valid], lf0_b[valid])) and loss_a.item() == loss_b.item()
    ms_c = timeit(ours_commuted, 10)
    row('C1@C2', 'same, first layer run at phone rate and its bf16 activations expanded (layer 1 commutes with the expansion)', ms_c,
        None, F, stock, note='bit-identical loss and valid-frame predictions: %s' % same)


if __name__ == '__main__':
    main()
