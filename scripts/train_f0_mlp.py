#!/usr/bin/env python
"""Config 1 / config 5: the README's F0 MLP (600 -> 512 -> 128 -> 32 -> 1, reference README.rst:61-99) trained on
synthetic lab / dur / lf0, through the new path, single- or multi-GPU (utterances sharded, NCCL all-reduce of the
gradients and of the loss / metric records), with the reference's own op chain + fp32 torch.nn on the CPU beside it.

    python scripts/train_f0_mlp.py --steps 30                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/train_f0_mlp.py --steps 30

What one step does, as ExperimentBuilder.train_epoch would (reference experiment_builder.py:464-490):
  zero_grad -> predict: fused min-max normalise + upsample (bf16 frames) -> 4 tcgen05 layers -> loss: masked mse (one
  kernel; its backward is the same kernel) -> backward -> [gradient all-reduce] -> Adam -> EMA (one multi-tensor kernel)
  -> RMSE metric (one kernel, device-resident state).
Prints one JSON line on rank 0: step time, frames/s, and the loss trajectory next to the CPU reference's.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import morgana_b200 as mg                                    # noqa: E402
from morgana_b200 import dp, nn as mnn, workloads           # noqa: E402
from oracle import aten_chain as ref                        # noqa: E402  (CPU reference leg only)

DIMS = [600, 512, 128, 32, 1]


def make_batches(n_batches, batch_size, seed):
    out = []
    for i in range(n_batches):
        ling = workloads.linguistic_batch(batch_size=batch_size, seed=seed + i)
        g = torch.Generator().manual_seed(seed + 7777 + i)
        T = int(ling['n_frames'].max())
        # a smooth synthetic normalised log-F0 target that depends on the labels, so there is something to learn
        frame_lab = ref.upsample_chain(ref.normalise_minmax_chain(ling['lab'], ling['mmin'], ling['mmax']), ling['dur'])
        target = torch.tanh(frame_lab[:, :, :8].sum(dim=-1, keepdim=True) - 4.) + 0.05 * torch.randn(batch_size, T, 1, generator=g)
        out.append({'lab': ling['lab'], 'dur': ling['dur'], 'n_frames': ling['n_frames'], 'target': target, 'T': T,
                    'mmin': ling['mmin'], 'mmax': ling['mmax']})
    return out


def build_reference_model(seed):
    torch.manual_seed(seed)
    layers = []
    for i in range(4):
        layers.append(torch.nn.Linear(DIMS[i], DIMS[i + 1]))
        if i < 3:
            layers.append(torch.nn.Sigmoid())
    return torch.nn.Sequential(*layers)


def cpu_reference_trajectory(batches, steps, seed, lr):
    """The reference's path on the CPU: its op chain + fp32 nn.Linear + Adam, same init, same data."""
    model = build_reference_model(seed)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    losses, t0 = [], time.perf_counter()
    for step in range(steps):
        b = batches[step % len(batches)]
        opt.zero_grad()
        frames = ref.upsample_chain(ref.normalise_minmax_chain(b['lab'], b['mmin'], b['mmax']), b['dur'])
        loss = ref.mse_chain(model(frames), b['target'], b['n_frames'])
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return losses, (time.perf_counter() - t0) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--batch-size', type=int, default=32, help='utterances per rank per step (the reference default)')
    ap.add_argument('--lr', type=float, default=0.01)
    ap.add_argument('--ema-decay', type=float, default=0.999)
    ap.add_argument('--seed', type=int, default=1234)
    ap.add_argument('--no-cpu-reference', action='store_true')
    ap.add_argument('--graph', action='store_true', help='capture the whole training step in one CUDA graph (static shapes: '
                    'every batch is padded to the longest utterance of the run) and replay it')
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    # every rank draws the same global stream of utterances and keeps its shard (utterances are the unit of sharding)
    n_batches = 4
    global_batches = make_batches(n_batches, args.batch_size * world, args.seed)
    lo, hi = dp.shard_range(args.batch_size * world, rank, world)
    batches = []
    for gb in global_batches:
        T = int(gb['n_frames'][lo:hi].max())
        batches.append({'lab': gb['lab'][lo:hi].to(dev), 'dur': gb['dur'][lo:hi].to(dev), 'n_frames': gb['n_frames'][lo:hi].to(dev),
                        'target': gb['target'][lo:hi, :T].contiguous().to(dev), 'T': T, 'frames': int(gb['n_frames'][lo:hi].sum())})
    mmin, mmax = global_batches[0]['mmin'].to(dev), global_batches[0]['mmax'].to(dev)
    if args.graph:
        # static shapes for capture: one (P, T) for the whole run; the kernels take the padded length as `max_len` and
        # mask by n_frames, so padding changes no result
        P_max, T_max = max(b['lab'].shape[1] for b in batches), max(b['T'] for b in batches)
        for b in batches:
            pad_p = P_max - b['lab'].shape[1]
            b['lab'] = torch.nn.functional.pad(b['lab'], (0, 0, 0, pad_p))
            b['dur'] = torch.nn.functional.pad(b['dur'], (0, 0, 0, pad_p))
            b['target'] = torch.nn.functional.pad(b['target'], (0, 0, 0, T_max - b['T']))
            b['T'] = T_max
        static = {k: torch.empty_like(batches[0][k]) for k in ('lab', 'dur', 'n_frames', 'target')}
        static['T'] = T_max

    ref_model = build_reference_model(args.seed)          # same initial weights as the CPU leg
    linears = [m for m in ref_model if isinstance(m, torch.nn.Linear)]
    layers = torch.nn.ModuleList([mnn.Linear(DIMS[i], DIMS[i + 1], act='sigmoid' if i < 3 else None,
                                             out_dtype=torch.bfloat16 if i < 3 else torch.float32, device=dev) for i in range(4)])
    ema_layers = torch.nn.ModuleList([mnn.Linear(DIMS[i], DIMS[i + 1], device=dev) for i in range(4)])
    for mine, avg, src in zip(layers, ema_layers, linears):
        mine.load_state_dict(src.state_dict())
        avg.load_state_dict(src.state_dict())
    def fresh_state():
        for mine, avg, src in zip(layers, ema_layers, linears):
            mine.load_state_dict(src.state_dict())
            avg.load_state_dict(src.state_dict())
        optimiser = torch.optim.Adam(layers.parameters(), lr=args.lr, fused=True, capturable=args.graph)
        metric = mg.metrics.RMSE()
        metric.reset_state()
        return optimiser, mg.utils.ExponentialMovingAverage(ema_layers, args.ema_decay), metric

    opt, ema, rmse = fresh_state()
    params = list(layers.parameters())
    bucket = None

    def train_step(b):
        nonlocal bucket
        opt.zero_grad(set_to_none=False)
        h = mg.utils.upsample_to_repetitions(b['lab'], b['dur'], normaliser=('minmax', mmin, mmax), max_len=b['T'],
                                             out_dtype=torch.bfloat16)
        for layer in layers:
            h = layer(h)
        loss = mg.losses.mse(h, b['target'], b['n_frames'])
        loss.backward()
        if world > 1:
            bucket = dp.allreduce_gradients(params, bucket=bucket)
        opt.step()
        ema.update_params(layers)
        rmse.accumulate(b['target'], h.detach(), seq_len=b['n_frames'])
        return loss

    # gradient parity at the initial weights (same data, same weights): bf16 operands in the layers vs fp32 on the CPU
    grad_rel = None
    if rank == 0 and world == 1 and not args.no_cpu_reference:
        b0, g0 = batches[0], global_batches[0]
        train_step(b0)                                       # leaves .grad of the first step in place (Adam already stepped,
        ours_grads = [p.grad.detach().cpu().clone() for p in params]   # but .grad still holds step-0 gradients)
        cpu_model = build_reference_model(args.seed)
        frames_cpu = ref.upsample_chain(ref.normalise_minmax_chain(g0['lab'], g0['mmin'], g0['mmax']), g0['dur'])
        ref.mse_chain(cpu_model(frames_cpu), g0['target'], g0['n_frames']).backward()
        grad_rel = max(float((a - b.grad).abs().max() / (b.grad.abs().max() + 1e-12))
                       for a, b in zip(ours_grads, cpu_model.parameters()))
    for step in range(3):                                   # warm-up (library initialisation), then start again from the
        train_step(batches[step % n_batches])               # initial weights with a fresh optimiser / EMA / metric
    opt, ema, rmse = fresh_state()
    losses = []
    graph, static_loss = None, None
    if args.graph:
        # Capture once (after a few eager steps on a side stream, as torch.cuda.graphs asks), replay per step: the ~60
        # launches of a step (scan, expansion, 4 + 8 GEMMs, reductions, Adam, EMA, metric, all-reduce) become one submit.
        def load(b):
            for k in ('lab', 'dur', 'n_frames', 'target'):
                static[k].copy_(b[k])
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for step in range(3):
                load(batches[step % n_batches])
                train_step(static)
        torch.cuda.current_stream().wait_stream(side)
        # Back to the initial state IN PLACE: anything allocated or zero-initialised during capture would be re-zeroed by
        # every replay (Adam's moments and step counter, a metric record), so they are created by the warm-up above and only
        # reset here.
        for mine, avg, src in zip(layers, ema_layers, linears):
            mine.load_state_dict(src.state_dict())
            avg.load_state_dict(src.state_dict())
        for state in opt.state.values():
            for value in state.values():
                if isinstance(value, torch.Tensor):
                    value.zero_()
        rmse.reset_state()
        rmse.accumulate(static['target'], torch.zeros_like(static['target']), seq_len=torch.zeros_like(static['n_frames']))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_loss = train_step(static)
        # the capture pass itself did not execute: state is still the fresh one
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    frames = 0
    start.record()
    for step in range(args.steps):
        b = batches[step % n_batches]
        if graph is None:
            losses.append(train_step(b))
        else:
            load(b)
            graph.replay()
            losses.append(static_loss.clone())
        frames += b['frames']
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / args.steps
    loss_values = torch.stack([l.detach() for l in losses]).to(torch.float64)
    stats = torch.tensor([ms, float(frames)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(loss_values)                        # equal shard sizes: global loss = mean of the rank losses (Q6)
        loss_values /= world
        t = stats[:1].clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        f = stats[1:].clone()
        dist.all_reduce(f)
        ms, frames = t.item(), f.item()
    packed = dp.allreduce_records(rmse._record)             # global RMSE from additive (sum, count)
    if rank == 0:
        line = {'config': 'C1/C5 README F0 MLP training, %d utterances / rank / step%s' % (args.batch_size, ', CUDA graph' if args.graph else ''),
                'n_gpus': world,
                'steps': args.steps, 'ms_per_step': round(ms, 3), 'valid_frames_per_s': round(frames / (ms * args.steps) * 1e3),
                'loss_first_last': [round(loss_values[0].item(), 5), round(loss_values[-1].item(), 5)],
                'train_rmse': round(float((packed[0, 0] / (packed[0, 1] + 1e-8)) ** 0.5), 5),
                'ema_param_delta': round(float((ema_layers[0].weight.detach() - layers[0].weight.detach()).abs().max()), 5)}
        if not args.no_cpu_reference:
            ref_losses, ref_s = cpu_reference_trajectory(global_batches, args.steps, args.seed, args.lr)
            ours = loss_values.tolist()
            rel = [abs(a - b) / max(abs(b), 1e-6) for a, b in zip(ours, ref_losses)]
            line.update({'cpu_reference_ms_per_step': round(ref_s * 1e3, 1), 'cpu_reference_loss_first_last':
                         [round(ref_losses[0], 5), round(ref_losses[-1], 5)],
                         'rel_loss_difference_step0': round(rel[0], 6),
                         'median_rel_loss_difference': round(sorted(rel)[len(rel) // 2], 4),
                         'max_rel_gradient_difference_step0': None if grad_rel is None else round(grad_rel, 4),
                         'tolerance_note': 'bf16 operands in the four layers, fp32 everywhere else: losses within 1e-3 relative of '
                                           'the fp32 CPU trajectory, step-0 gradients within 1 % of their range'})
        print(json.dumps(line), flush=True)
    if graph is not None:               # a captured NCCL all-reduce must be released before the process group goes away
        torch.cuda.synchronize()
        graph.reset()
        del graph
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
