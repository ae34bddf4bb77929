"""Further timed sections of ``bench.py`` (loaded by path from there; also usable on their own, see ``__main__``).

``other_configs(dev, peaks)``  -- BASELINE.json configs 1-4 on one B200, every row with its own CUDA-event timing, and the
    reference's own functions run on CUDA tensors beside them ("stock ATen on the same GPU": what morgana does today).
``training_section(rank, world, dev, peaks)`` -- BASELINE.json configs[4]: data-parallel training of the README F0 MLP
    (600 -> 512 -> 128 -> 32 -> 1, README.rst:61-99), 32 utterances per rank per step, flat-bucket gradient all-reduce.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


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


def timeit_flushed(fn, n_iter=20, warmup=3, flush_mb=512):
    """Like :func:`timeit` for working sets that fit the 126 MB L2: a buffer larger than L2 is overwritten between two
    timed calls (outside the event pairs), so every call starts from HBM."""
    flush = torch.empty(flush_mb << 20, dtype=torch.uint8, device='cuda')
    for _ in range(warmup):
        fn()
    pairs = []
    for _ in range(n_iter):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        pairs.append((s, e))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in pairs) / n_iter


class _Stock(object):
    """The reference's own functions on CUDA tensors (oracle/_ref through oracle/ref_loader.py); the restated op chain
    (oracle/aten_chain.py) when the mirror is absent."""
    def __init__(self):
        from oracle import aten_chain, ref_loader
        self.port = aten_chain
        self.morgana = ref_loader.import_reference() if ref_loader.available() else None
        self.kind = 'reference functions on CUDA tensors' if self.morgana is not None else 'restated op chain on CUDA tensors'

    def upsample(self, lab, dur, mmin, mmax):
        if self.morgana is None:
            return self.port.upsample_chain(self.port.normalise_minmax_chain(lab, mmin, mmax), dur)
        return self.morgana.utils.upsample_to_repetitions(self.morgana.data.normalise_minmax(lab, mmin, mmax), dur)

    def mse(self, pred, target, n):
        return self.port.mse_chain(pred, target, n) if self.morgana is None else self.morgana.losses.mse(pred, target, n)

    def rmse_accumulate(self, target, pred, n):
        if self.morgana is None:
            return self.port.rmse_increment(target, pred, n)
        if not hasattr(self, '_rmse'):
            self._rmse = self.morgana.metrics.RMSE()
            self._rmse.reset_state()
        return self._rmse.accumulate(target, pred, n)

    def ema(self, ema_model, model):
        if self.morgana is None:
            return self.port.ema_chain([p.data for p in ema_model.parameters()], [p.data for p in model.parameters()], 0.999)
        if not hasattr(self, '_ema'):
            self._ema = self.morgana.utils.ExponentialMovingAverage(ema_model, 0.999)
        return self._ema.update_params(model)


def other_configs(dev, peaks, reference_path=None):
    import morgana_b200 as mg
    from morgana_b200 import nn as mnn, ops, workloads
    from morgana_b200.fused import AcousticObjective
    hbm, tflops = peaks['hbm_gbs'], peaks.get('bf16_tflops', 1590.0)
    stock = _Stock()
    rows = []

    def row(config, op, ms, alg_bytes=None, frames=None, stock_ms=None, flops=None, note=None):
        out = {'config': config, 'op': op, 'ms': round(ms, 4)}
        if alg_bytes is not None:
            out['algorithmic_gb_per_s'] = round(alg_bytes / ms / 1e6, 1)
            out['frac_of_measured_hbm_peak'] = round(alg_bytes / ms / 1e6 / hbm, 3)
        if flops is not None:
            out['tflop_per_s'] = round(flops / ms / 1e9, 1)
            out['frac_of_measured_bf16_burst'] = round(flops / ms / 1e9 / tflops, 3)
        if frames is not None:
            out['valid_frames_per_s'] = round(frames / ms * 1e3)
        if stock_ms is not None:
            out['stock_aten_same_gpu_ms'] = round(stock_ms, 4)
            out['speedup_vs_stock_aten'] = round(stock_ms / ms, 1)
        if note:
            out['note'] = note
        rows.append(out)

    # ---- config 2: the path's first half alone, next to what the reference does today on the same GPU -------------------
    ling = workloads.linguistic_batch(batch_size=256, seed=1234)
    lab, dur = ling['lab'].to(dev), ling['dur'].to(dev)
    mmin, mmax = ling['mmin'].to(dev), ling['mmax'].to(dev)
    B, P, D = lab.shape
    T, F = int(ling['n_frames'].max()), int(ling['n_frames'].sum())
    n_items = int(ling['n_phones'].sum())
    k2_bytes = 4 * D * (B * T + n_items) + 4 * B * P + 8 * D
    stock_ms = timeit(lambda: stock.upsample(lab, dur, mmin, mmax), n_iter=5, warmup=1)
    ms = timeit(lambda: mg.utils.upsample_to_repetitions(lab, dur, normaliser=('minmax', mmin, mmax), max_len=T))
    row('C2', 'normalise_minmax + upsample_to_repetitions (K1 + K2, max_len hint)', ms, k2_bytes, F, stock_ms)
    ms = timeit(lambda: mg.utils.upsample_to_repetitions(lab, dur, normaliser=('minmax', mmin, mmax)))
    row('C2', 'same through the reference signature (32-byte read-back sizes the output)', ms, k2_bytes, F, stock_ms)
    ac = workloads.acoustic_batch(ling['n_frames'], seed=1234)
    pred2, target2, voiced2, n2 = ac['pred'].to(dev), ac['target'].to(dev), ac['voiced'].to(dev), ling['n_frames'].to(dev)
    ref_path = reference_path
    if ref_path is None:
        import bench
        ref_path = bench.ReferencePath()
    stock_step = timeit(lambda: (stock.upsample(lab, dur, mmin, mmax), ref_path.objective(pred2, target2, n2, voiced2)), n_iter=3, warmup=1)
    objective = AcousticObjective()

    def whole_step():
        out, n = mg.utils.upsample_to_repetitions(lab, dur, normaliser=('minmax', mmin, mmax), max_len=T, return_lengths=True)
        return objective(pred2, target2, n)
    ms = timeit(whole_step)
    row('C2', 'the whole bench step: normalise + upsample + objective (loss, gradient, 4 metrics)', ms,
        k2_bytes + 4 * 187 * (2 * F + B * T), F, stock_step,
        note='stock = %s: data.normalise_minmax, utils.upsample_to_repetitions, 3 x losses.mse + losses.bce + backward, '
             '4 x metrics.accumulate' % stock.kind)
    # K4b alone: back-to-back launches alternating between two input pairs (each pair is 0.5 GB: nothing of one launch's inputs is
    # left in L2 for the next); consecutive launches overlap head and tail (programmatic dependent launch), no event in between
    pred2b, target2b = pred2.clone(), target2.clone()
    pairs, turn = ((pred2, target2), (pred2b, target2b)), [0]

    def objective_alone():
        turn[0] ^= 1
        return objective(pairs[turn[0]][0], pairs[turn[0]][1], n2)
    k4b_bytes = 4 * 187 * (2 * F + B * T)
    row('C2', 'the objective alone (K4b: 3 mse + bce + gradient + 4 metrics), back-to-back launches on two input pairs',
        timeit(objective_alone, 40, 5), k4b_bytes, F, None,
        note='inside the bench step the same kernel is timed between two event records, right after K2: see roofline_k4b')
    row('C2', 'the objective alone, forward only (no gradient)', timeit(lambda: objective(pred2, target2, n2, want_grad=False), 40, 5),
        4 * 187 * 2 * F, F, None)
    del pred2, target2, voiced2, ac, pred2b, target2b, pairs

    # ---- config 3: 1024 utterances, 187-dim targets -------------------------------------------------------------------
    n3 = workloads.acoustic_lengths(batch_size=1024, seed=1234)
    ac = workloads.acoustic_batch(n3, seed=1234)
    pred, target, voiced, n3d = ac['pred'].to(dev), ac['target'].to(dev), ac['voiced'].to(dev), n3.to(dev)
    F3, T3 = int(n3.sum()), int(n3.max())
    stock_ms = timeit(lambda: stock.mse(pred, target, n3d), 5, 1)
    row('C3', 'losses.mse forward, (1024, %d, 187)' % T3, timeit(lambda: mg.losses.mse(pred, target, n3d)), 8 * 187 * F3, F3, stock_ms)
    pg, pr = pred.clone().requires_grad_(), pred.clone().requires_grad_()

    def mse_fb():
        pg.grad = None
        mg.losses.mse(pg, target, n3d).backward()

    def stock_mse_fb():
        pr.grad = None
        stock.mse(pr, target, n3d).backward()
    row('C3', 'losses.mse forward + backward', timeit(mse_fb, 10), 16 * 187 * F3 + 4 * 187 * 1024 * T3, F3, timeit(stock_mse_fb, 5, 1))
    rmse = mg.metrics.RMSE()
    rmse.reset_state()
    stock_ms = timeit(lambda: stock.rmse_accumulate(target, pred, n3d), 5, 1)
    row('C3', 'metrics.RMSE.accumulate, 187 dims', timeit(lambda: rmse.accumulate(target, pred, seq_len=n3d)), 8 * 187 * F3, F3, stock_ms)
    objective3 = AcousticObjective()
    stock_ms = timeit(lambda: ref_path.objective(pred, target, n3d, voiced), 3, 1)
    row('C3', 'whole objective of models/RNN_SPSS.py:120-139 (3 mse + bce + gradient + 4 metrics), one launch (K4b)',
        timeit(lambda: objective3(pred, target, n3d)), 8 * 187 * F3 + 4 * 187 * 1024 * T3, F3, stock_ms)
    del pred, target, voiced, pg, pr, ac

    # ---- config 4: EMA of LSTMAcousticModel-sized parameters (187 outputs) ---------------------------------------------
    class _Shapes(torch.nn.Module):
        def __init__(self):
            super().__init__()
            shapes = [(512, 609), (512,)] + [(2048, 512), (2048, 512), (2048,), (2048,)] * 8 + [(256, 512), (256,), (187, 256), (187,)]
            self.p = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*s, device=dev)) for s in shapes])
    model, ema_ours, ema_stock = _Shapes(), _Shapes(), _Shapes()
    n_par = sum(p.numel() for p in model.parameters())
    ema = mg.utils.ExponentialMovingAverage(ema_ours, 0.999)
    # 138 MB of state is about the size of the L2: back-to-back calls would be served from it (8.9 TB/s "of HBM"), so the L2 is
    # flushed between calls; the back-to-back figure is kept in the note (in training the optimiser has just touched the parameters)
    stock_ms = timeit_flushed(lambda: stock.ema(ema_stock, model), 10, 2)
    warm_ms = timeit(lambda: ema.update_params(model))
    row('C4', 'ExponentialMovingAverage.update_params, %d parameters in %d tensors (K6), L2 flushed between calls' % (n_par, len(model.p)),
        timeit_flushed(lambda: ema.update_params(model)), 12 * n_par, None, stock_ms,
        note='back to back (state partly resident in the 126 MB L2): %.4f ms' % warm_ms)
    del model, ema_ours, ema_stock

    # ---- the tcgen05 layers (K7 forward, K7g, K7w) at the config-2 frame count -----------------------------------------
    M = B * T
    x = torch.randn(M, 600, device=dev).to(torch.bfloat16)
    w = (torch.randn(512, 600, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(512, device=dev)
    lin = torch.nn.Linear(600, 512, device=dev, dtype=torch.bfloat16)
    stock_ms = timeit(lambda: torch.sigmoid(lin(x)), 10)
    ms = timeit(lambda: ops.linear_bf16(x, w, bias, act='sigmoid', out_dtype=torch.bfloat16), 10)
    row('K7', 'Linear 600 -> 512 + Sigmoid forward over %d frames (tcgen05, bf16 out)' % M, ms, None, None, stock_ms,
        flops=2.0 * M * 600 * 512, note='stock = cuBLAS bf16 nn.Linear + sigmoid kernel')
    y = ops.linear_bf16(x, w, bias, act='sigmoid', out_dtype=torch.bfloat16)
    gy = torch.randn(M, 512, device=dev)
    ms = timeit(lambda: ops.act_grad_bf16(gy, y), 10)
    row('K7g', 'sigmoid backward + bf16 cast + bias gradient, (%d, 512)' % M, ms, (4 + 2 + 2) * M * 512)
    g16, _ = ops.act_grad_bf16(gy, y)
    stock_ms = timeit(lambda: torch.matmul(g16.t(), x), 10)
    ms = timeit(lambda: ops.linear_wgrad_bf16(g16, x, out_features=512, in_features=600), 10)
    row('K7w', 'weight gradient 512 x 600 over %d frames (tcgen05, MN-major operands)' % M, ms, None, None, stock_ms,
        flops=2.0 * M * 600 * 512, note='stock = cuBLAS bf16 matmul')
    del x, y, gy, g16

    # ---- config 1 shapes at config-2 scale: README MLP predict + loss ---------------------------------------------------
    dims = [600, 512, 128, 32, 1]
    torch.manual_seed(0)
    layers = [mnn.Linear(dims[i], dims[i + 1], act='sigmoid' if i < 3 else None,
                         out_dtype=torch.bfloat16 if i < 3 else torch.float32, device=dev) for i in range(4)]
    stock_layers = torch.nn.Sequential(*[m for i in range(4) for m in
                                         ([torch.nn.Linear(dims[i], dims[i + 1])] + ([torch.nn.Sigmoid()] if i < 3 else []))]).to(dev)
    tgt = torch.randn(B, T, 1, device=dev)
    n_frames = ling['n_frames'].to(dev)

    def ours():
        with torch.no_grad():
            h = mg.utils.upsample_to_repetitions(lab, dur, normaliser=('minmax', mmin, mmax), max_len=T, out_dtype=torch.bfloat16)
            for layer in layers:
                h = layer(h)
            return mg.losses.mse(h, tgt, n_frames)

    def stock_chain():
        with torch.no_grad():
            return stock.mse(stock_layers(stock.upsample(lab, dur, mmin, mmax)), tgt, n_frames)
    flops = 2.0 * B * T * sum(dims[i] * dims[i + 1] for i in range(4))
    row('C1@C2', 'README F0 MLP predict + loss: fused normalise / upsample (bf16) -> 4 tcgen05 layers -> losses.mse', timeit(ours, 10),
        None, F, timeit(stock_chain, 5, 1), flops=flops, note='stock = fp32 cuBLAS layers + the reference functions; TFLOP/s counts the '
        'four layers and includes the feature path')
    return rows


def training_section(rank, world, dev, peaks, steps=40, batch_size=32):
    """BASELINE.json configs[4]: README F0 MLP, data-parallel: every rank trains on its own 32 utterances per step, ONE NCCL
    all-reduce of the flat gradient buffer per step (trainer.GradientBucket), fused Adam, multi-tensor EMA, RMSE metric.
    Weak scaling; CUDA events, max over ranks.  Three timings of the same step: eager (what a Python training loop gets:
    host-bound at this batch size), replayed from a CUDA graph (device time), and the same graph without the all-reduce --
    the difference is the collective's exposed time."""
    import morgana_b200 as mg
    from morgana_b200 import nn as mnn, trainer, workloads
    dims = [600, 512, 128, 32, 1]
    torch.manual_seed(1234)
    layers = torch.nn.ModuleList([mnn.Linear(dims[i], dims[i + 1], act='sigmoid' if i < 3 else None,
                                             out_dtype=torch.bfloat16 if i < 3 else torch.float32, device=dev) for i in range(4)])
    ema_layers = torch.nn.ModuleList([mnn.Linear(dims[i], dims[i + 1], device=dev) for i in range(4)])
    if world > 1:
        for p in layers.parameters():
            dist.broadcast(p.data, 0)
    ema_layers.load_state_dict(layers.state_dict())
    initial = {k: v.clone() for k, v in layers.state_dict().items()}
    bucket = trainer.GradientBucket(layers.parameters())
    opt = torch.optim.Adam(layers.parameters(), lr=0.01, fused=True, capturable=True)
    ema = mg.utils.ExponentialMovingAverage(ema_layers, 0.999)
    rmse = mg.metrics.RMSE()
    rmse.reset_state()
    raw = []
    for i in range(4):
        ling = workloads.linguistic_batch(batch_size=batch_size, seed=4096 + 31 * i + 1009 * rank)     # this rank's shard
        g = torch.Generator().manual_seed(99 + i + 1009 * rank)
        raw.append((ling, torch.randn(batch_size, int(ling['n_frames'].max()), 1, generator=g)))
    # static shapes (one P and T for the run) so the step can be captured; the kernels mask by n_frames, padding changes nothing
    P_max = max(l['lab'].shape[1] for l, _ in raw)
    T_max = max(int(l['n_frames'].max()) for l, _ in raw)
    if world > 1:
        dims_t = torch.tensor([P_max, T_max], device=dev)
        dist.all_reduce(dims_t, op=dist.ReduceOp.MAX)
        P_max, T_max = (int(v) for v in dims_t.tolist())
    pad = torch.nn.functional.pad
    batches = []
    for ling, tgt in raw:
        batches.append({'lab': pad(ling['lab'], (0, 0, 0, P_max - ling['lab'].shape[1])).to(dev),
                        'dur': pad(ling['dur'], (0, 0, 0, P_max - ling['dur'].shape[1])).to(dev),
                        'n_frames': ling['n_frames'].to(dev), 'target': pad(tgt, (0, 0, 0, T_max - tgt.shape[1])).to(dev),
                        'frames': int(ling['n_frames'].sum())})
    mmin, mmax = raw[0][0]['mmin'].to(dev), raw[0][0]['mmax'].to(dev)
    static = {k: torch.empty_like(batches[0][k]) for k in ('lab', 'dur', 'n_frames', 'target')}
    stream = torch.cuda.current_stream()

    def train_step(b, exchange=True):
        bucket.zero()
        h = mg.utils.upsample_to_repetitions(b['lab'], b['dur'], normaliser=('minmax', mmin, mmax), max_len=T_max,
                                             out_dtype=torch.bfloat16)
        for layer in layers:
            h = layer(h)
        loss = mg.losses.mse(h, b['target'], b['n_frames'])
        loss.backward()
        if exchange:
            bucket.all_reduce()
        opt.step()
        ema.update_params(layers)
        rmse.accumulate(b['target'], h.detach(), seq_len=b['n_frames'])
        return loss

    def load(b):
        for k in static:
            static[k].copy_(b[k])

    def reset():
        layers.load_state_dict(initial)
        ema_layers.load_state_dict(initial)
        for state in opt.state.values():
            for value in state.values():
                if isinstance(value, torch.Tensor):
                    value.zero_()

    def timed(run_step, steps=steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        frames, loss = 0, None
        start.record(stream)
        for i in range(steps):
            loss = run_step(batches[i % 4])
            frames += batches[i % 4]['frames']
        stop.record(stream)
        torch.cuda.synchronize()
        stats = torch.tensor([start.elapsed_time(stop) / steps, float(frames)], dtype=torch.float64, device=dev)
        if world > 1:
            t = stats[:1].clone()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            f = stats[1:].clone()
            dist.all_reduce(f)
            stats = torch.cat([t, f])
        ms, frames = stats.tolist()
        return ms, frames, float(loss.detach())

    # eager
    for i in range(5):
        train_step(batches[i % 4])        # nothing keeps the loss: a live autograd graph pins the AccumulateGrad nodes to this stream
    reset()
    first_loss = float(train_step(batches[0]).detach())
    eager_ms, frames, eager_last = timed(train_step)

    # captured: warm up on a side stream as torch.cuda.graphs asks, reset the state in place, capture, replay
    graphs = {}
    side = torch.cuda.Stream(device=dev)
    for name, exchange in (('with_allreduce', True), ('without_allreduce', False)):
        side.wait_stream(stream)
        with torch.cuda.stream(side):
            for i in range(3):
                load(batches[i % 4])
                train_step(static, exchange)
        stream.wait_stream(side)
        reset()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_loss = train_step(static, exchange).detach()
        graphs[name] = (graph, static_loss)

    def replay(name):
        graph, static_loss = graphs[name]

        def run(b):
            load(b)
            graph.replay()
            return static_loss
        return run
    graph_ms, _, graph_last = timed(replay('with_allreduce'))
    reset()
    bare_ms, _, _ = timed(replay('without_allreduce'))
    # BASELINE.json configs[4] to the letter: ONE epoch over 4096 utterances sharded over the ranks, 32 per rank per step (strong
    # scaling of the epoch: 128 / 64 / 32 / 16 steps per rank at 1 / 2 / 4 / 8 GPUs); each rank's shard cycles over its 4 batches
    reset()
    epoch_steps = max(4096 // (batch_size * world), 1)
    epoch_step_ms, epoch_frames, _ = timed(replay('with_allreduce'), epoch_steps)
    torch.cuda.synchronize()
    for graph, _ in graphs.values():       # a captured NCCL all-reduce must be released before the process group goes away
        graph.reset()
    graphs.clear()
    n_par = bucket.flat.numel()
    exposed = max(graph_ms - bare_ms, 0.) if world > 1 else 0.
    return {'config': 'configs[4] / configs[0] model: README F0 MLP 600-512-128-32-1, %d utterances per rank per step '
                      '(one process per GPU; utterances sharded)' % batch_size,
            'n_gpus': world, 'steps': steps,
            'ms_per_step': round(graph_ms, 4), 'valid_frames_per_s': round(frames / (graph_ms * steps) * 1e3),
            'eager_ms_per_step': round(eager_ms, 4), 'eager_valid_frames_per_s': round(frames / (eager_ms * steps) * 1e3),
            'gradient_allreduce': {'bytes': 4 * n_par, 'exposed_ms': round(exposed, 4),
                                   'share_of_step': round(exposed / graph_ms, 4) if graph_ms else None,
                                   'step_without_it_ms': round(bare_ms, 4),
                                   'what': 'one NCCL all-reduce of the flat fp32 gradient buffer (%d KB) between backward and Adam, '
                                           'captured in the step\'s CUDA graph; exposed = step with - step without the collective'
                                           % (4 * n_par // 1024)},
            'epoch_of_4096_utterances': {'steps_per_rank': epoch_steps, 'ms': round(epoch_step_ms * epoch_steps, 3),
                                         'valid_frames_per_s': round(epoch_frames / (epoch_step_ms * epoch_steps) * 1e3),
                                         'what': 'configs[4] as written: 4096 utterances sharded over the ranks, 32 per rank per step, '
                                                 'the captured step replayed steps_per_rank times (CUDA events, max over ranks)'},
            'loss_first_last': [round(first_loss, 5), round(graph_last, 5)],
            'step': 'upsample_to_repetitions (fused minmax, bf16 frames) -> 4 x nn.Linear (tcgen05 forward, act-grad, weight-gradient, '
                    'input-gradient kernels) -> losses.mse -> backward -> all-reduce -> fused Adam -> EMA (K6) -> RMSE metric; '
                    'ms_per_step = the whole step replayed from one CUDA graph, eager_ms_per_step = the same Python loop un-captured'}


if __name__ == '__main__':
    import json
    torch.cuda.set_device(0)
    pk = {'hbm_gbs': 6549.8, 'bf16_tflops': 1670.5}
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        pk = json.load(open(path))
    for r in other_configs(torch.device('cuda', 0), pk):
        print(json.dumps(r), flush=True)
    print(json.dumps(training_section(0, 1, torch.device('cuda', 0), pk)), flush=True)
