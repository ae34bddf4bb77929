#!/usr/bin/env python
"""Largest relative error of the loss gradients against the fp64 oracle / the reference's golden outputs, per loss kind: the numbers
behind the tolerances written in tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import morgana_b200 as mg                    # noqa: E402
from oracle import np_oracle as O            # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def worst(got, want, floor):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    big = np.abs(want) > floor
    return float((np.abs(got - want)[big] / np.abs(want)[big]).max()), float(np.abs(got - want)[~big].max() if (~big).any() else 0.)


for kind in ('mse', 'l1', 'bce'):
    out = []
    for B, T, D in [(8, 301, 187), (5, 1000, 3), (3, 77, 180), (64, 50, 1)]:
        rng = np.random.default_rng(B + T + D)
        seq_len = rng.integers(1, T + 1, B)
        p = rng.random((B, T, D), dtype=np.float32) * 0.98 + 0.01
        y = (rng.random((B, T, D)) < 0.5).astype(np.float32) if kind == 'bce' else rng.standard_normal((B, T, D)).astype(np.float32)
        pt = dev(p).requires_grad_()
        getattr(mg.losses, kind)(pt, dev(y), dev(seq_len)).backward()
        out.append(worst(pt.grad.cpu().numpy(), O.masked_loss_grad(p, y, seq_len, kind), 1e-12)[0])
    print('%-4s gradient: worst relative error vs fp64 oracle over 4 shapes: %.2e' % (kind, max(out)))

rng = np.random.default_rng(41)
B, T, C = 7, 301, 40
logits = (2. * rng.standard_normal((B, T, C))).astype(np.float32)
classes, seq_len = rng.integers(0, C, (B, T)), rng.integers(1, T + 1, B)
lt = dev(logits).requires_grad_()
mg.losses.ce(lt, dev(classes), dev(seq_len)).backward()
_, want = O.cross_entropy_loss(logits, classes, seq_len)
for floor in (1e-12, 1e-9, 1e-7):
    r, a = worst(lt.grad.cpu().numpy(), want, floor)
    print('ce   gradient: worst relative error where |g| > %.0e: %.2e ; worst absolute error below: %.2e (largest |g| %.2e)' % (floor, r, a, np.abs(want).max()))

m, lv = rng.standard_normal((1000, 257)).astype(np.float32), (0.3 * rng.standard_normal((1000, 257))).astype(np.float32)
md, lvd = dev(m).requires_grad_(), dev(lv).requires_grad_()
mg.losses.KLD_standard_normal(md, lvd).backward()
_, gm, glv = O.kld_standard_normal(m, lv)
print('kld  gradient wrt mean: %.2e ; wrt log-variance: rel %.2e where |g| > 1e-9, abs below %.2e' % ((worst(md.grad.cpu().numpy(), gm, 1e-12)[0],) + worst(lvd.grad.cpu().numpy(), glv, 1e-9)))
