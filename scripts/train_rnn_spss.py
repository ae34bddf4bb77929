#!/usr/bin/env python
"""Config 4 / config 5: the LSTM acoustic model's training step with EMA (reference models/RNN_SPSS.py:21-139 driven by
experiment_builder.py:464-490) on synthetic 187-dim WORLD targets, through the new path, on 1..N GPUs.

    python scripts/train_rnn_spss.py --steps 8                                    # config 4, one GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/train_rnn_spss.py \
        --utterances 4096 --epochs 1                                              # config 5 (utterances sharded)

Two arms on the same GPU, same data, same initial weights:

  ours   predict = fused min-max normalise + upsample (K1 + K2) | counters normalise (K3) -> Linear + Sigmoid (K7) ->
         8 cuDNN LSTMs (library, out of scope) -> Linear + Sigmoid, Linear (K7) -> denormalise (K3) + MLPG on the device
         (K8); loss = 3 x losses.mse + losses.bce and four metric accumulators (K4 / K5, no host syncs); then backward,
         flat-bucket gradient all-reduce, fused Adam, multi-tensor EMA (K6) -- `morgana_b200.trainer.DataParallelTrainer`.
  stock  the reference's op chain (oracle/aten_chain.py) with torch.nn layers in fp32, MLPG per utterance and dimension
         on the CPU with the device<->host round trip of RNN_SPSS.py:111-116, its per-metric `.item()` syncs, and the
         per-tensor EMA.  This is what running the unmodified reference on this GPU costs.

Prints one JSON line on rank 0: ms per step of both arms, the phase breakdown of ours, valid frames/s, loss trajectories.
"""
import argparse
import copy
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import morgana_b200 as mg                                          # noqa: E402
from morgana_b200 import data as mdata, losses, metrics, nn as mnn, trainer, utils   # noqa: E402
from morgana_b200.viz.synthesis import MLPG                        # noqa: E402

STATIC = {'lf0': 1, 'mcep': 60, 'bap': 1}                          # static dims; the model predicts static + 2 deltas
OUTPUT_DIMS = {'lf0': 3, 'vuv': 1, 'mcep': 180, 'bap': 3}          # 187 columns, as BASELINE.json config 4 names them
LAB_DIM, COUNTER_DIM, HIDDEN = 600, 9, 512


# ----------------------------------------------------------------------------------------------------------------------
# synthetic utterances (shapes and statistics only)
# ----------------------------------------------------------------------------------------------------------------------
def with_deltas(static):
    """[static | delta | delta-delta] with the windows MLPG assumes ([-.5, 0, .5] and [1, -2, 1], zero beyond the ends)."""
    padded = torch.nn.functional.pad(static, (0, 0, 1, 1))
    prev, nxt = padded[:, :-2], padded[:, 2:]
    return torch.cat([static, 0.5 * (nxt - prev), prev - 2. * static + nxt], dim=-1)


def make_normalisers(dev, seed):
    g = torch.Generator().manual_seed(seed)
    norm = {'lab': mdata.MinMaxNormaliser('lab').set_params(
                {'mmin': np.zeros(LAB_DIM), 'mmax': (torch.rand(LAB_DIM, generator=g) + 0.5).numpy()}, device=dev),
            'counters': mdata.MinMaxNormaliser('counters').set_params(
                {'mmin': np.zeros(COUNTER_DIM), 'mmax': np.full(COUNTER_DIM, 40.)}, device=dev)}
    for name, dim in STATIC.items():
        mean = 5. if name == 'lf0' else 0.
        norm[name] = mdata.MeanVarianceNormaliser(name, use_deltas=True).set_params(
            {'mean': np.full(dim, mean), 'std_dev': np.full(dim, 0.5)},
            {'mean': np.concatenate([np.full(dim, mean), np.zeros(2 * dim)]),
             'std_dev': np.concatenate([np.full(dim, 0.5), np.full(dim, 0.2), np.full(dim, 0.3)])}, device=dev)
    return norm


def make_batch(index, batch_size, normalisers, dev, proj):
    """One batch of utterances; the content depends only on `index`, so any sharding sees the same global data set."""
    g = torch.Generator(device=dev).manual_seed(100003 + index)
    n_phones = torch.randint(40, 81, (batch_size,), generator=g, device=dev)
    P = int(n_phones.max())
    valid_phone = torch.arange(P, device=dev)[None] < n_phones[:, None]
    dur = (torch.randint(1, 31, (batch_size, P), generator=g, device=dev) * valid_phone)[:, :, None]
    lab = torch.rand(batch_size, P, LAB_DIM, generator=g, device=dev) * normalisers['lab'].params_torch['mmax'] * valid_phone[:, :, None]
    n_frames = dur.sum(dim=(1, 2))
    lengths = n_frames.tolist()                                      # known on the host before upload, as in the reference
    T = max(lengths)
    valid_frame = (torch.arange(T, device=dev)[None] < n_frames[:, None])[:, :, None]
    counters = torch.rand(batch_size, T, COUNTER_DIM, generator=g, device=dev) * 40. * valid_frame
    # smooth targets that depend on the labels, so the loss has somewhere to go
    frames = utils.upsample_to_repetitions(lab, dur, normaliser=normalisers['lab'], max_len=T)
    drive = torch.tanh(frames[:, :, :64] @ proj - 2.)                # (B, T, 62)
    feats = {'n_frames': n_frames, 'lengths': lengths, 'T': T, 'dur': dur, 'lab': lab, 'counters': counters,
             'frames_total': int(sum(lengths))}
    column = 0
    for name, dim in STATIC.items():
        static = drive[:, :, column:column + dim] * 0.5 + (5. if name == 'lf0' else 0.)
        static = (static + 0.02 * torch.randn(batch_size, T, dim, generator=g, device=dev)) * valid_frame
        column += dim
        deltas = with_deltas(static) * valid_frame
        feats[name] = static
        feats['normalised_%s_deltas' % name] = normalisers[name].normalise(deltas, deltas=True) * valid_frame
    feats['vuv'] = (drive[:, :, :1] > -0.5) & valid_frame
    feats['normalised_lab'] = normalisers['lab'].normalise(lab) * valid_phone[:, :, None]      # what the stock arm's loader ships
    feats['normalised_counters'] = normalisers['counters'].normalise(counters) * valid_frame
    return feats


# ----------------------------------------------------------------------------------------------------------------------
# the model, on the new path
# ----------------------------------------------------------------------------------------------------------------------
class AcousticModel(torch.nn.Module):
    """609 -> 512 (sigmoid) -> 8 x LSTM(512) -> 256 (sigmoid) -> 187, the layer stack of models/RNN_SPSS.py:32-42, with the
    reference's BaseModel protocol: forward(features) -> (loss, outputs), .mode, .step, .metrics."""
    def __init__(self, normalisers, num_layers=8, pack=True, device=None):
        super().__init__()
        self.normalisers, self.pack = normalisers, pack
        self.mode, self.step = '', 0
        self.inp = mnn.Linear(LAB_DIM + COUNTER_DIM, HIDDEN, act='sigmoid', device=device)
        self.recurrent = torch.nn.ModuleList([torch.nn.LSTM(HIDDEN, HIDDEN, batch_first=True, device=device) for _ in range(num_layers)])
        self.mid = mnn.Linear(HIDDEN, 256, act='sigmoid', device=device)
        self.out = mnn.Linear(256, sum(OUTPUT_DIMS.values()), device=device)
        self.metrics = metrics.Handler(loss=metrics.Mean())
        self.metrics.add_metrics('all', LF0_RMSE_Hz=metrics.LF0Distortion(), VUV_accuracy=metrics.Mean(),
                                 MCEP_distortion=metrics.MelCepDistortion(), BAP_distortion=metrics.Distortion())
        self.phase = None                                            # optional callback(name) for the phase timer

    def _mark(self, name):
        if self.phase is not None:
            self.phase(name)

    def predict(self, features):
        T = features['T']
        lab_frames = utils.upsample_to_repetitions(features['lab'], features['dur'], normaliser=self.normalisers['lab'], max_len=T)
        counters = self.normalisers['counters'].normalise(features['counters'])
        x = torch.cat((lab_frames, counters), dim=-1)
        self._mark('features')
        h = self.inp(x)
        self._mark('dense')
        if self.pack:
            h = torch.nn.utils.rnn.pack_padded_sequence(h, torch.tensor(features['lengths']), batch_first=True, enforce_sorted=False)
        for lstm in self.recurrent:
            h, _ = lstm(h)
        if self.pack:
            h, _ = torch.nn.utils.rnn.pad_packed_sequence(h, batch_first=True, total_length=T)
        self._mark('lstm')
        y = self.out(self.mid(h))
        self._mark('dense')
        outputs = {}
        column = 0
        for name, width in OUTPUT_DIMS.items():
            block = y[:, :, column:column + width]
            column += width
            if name == 'vuv':
                outputs['vuv'] = torch.sigmoid(block)
                continue
            outputs['normalised_%s_deltas' % name] = block
            normaliser = self.normalisers[name]
            deltas = normaliser.denormalise(block.detach(), deltas=True)
            outputs[name] = MLPG(deltas, normaliser.delta_params_torch['std_dev'] ** 2, padding_size=100)
        self._mark('outputs')
        return outputs

    def loss(self, features, outputs):
        n_frames = features['n_frames']
        voiced = outputs['vuv'] > 0.5
        self.metrics.accumulate(
            self.mode,
            LF0_RMSE_Hz=(features['lf0'], outputs['lf0'], voiced, n_frames),
            VUV_accuracy=((features['vuv'] == voiced).type(torch.float), n_frames),
            MCEP_distortion=(features['mcep'], outputs['mcep'], n_frames),
            BAP_distortion=(features['bap'], outputs['bap'], n_frames))
        total = losses.bce(outputs['vuv'], features['vuv'].type(torch.float), n_frames)
        for name in STATIC:
            key = 'normalised_%s_deltas' % name
            total = total + losses.mse(outputs[key], features[key], n_frames)
        self._mark('loss')
        return total / 4.

    def forward(self, features):
        outputs = self.predict(features)
        return self.loss(features, outputs), outputs


# ----------------------------------------------------------------------------------------------------------------------
# the stock arm: reference op chain + torch.nn, on the same GPU
# ----------------------------------------------------------------------------------------------------------------------
class StockModel(torch.nn.Module):
    def __init__(self, ours, normalisers):
        super().__init__()
        self.normalisers = normalisers
        self.inp, self.mid, self.out = (torch.nn.Linear(m.in_features, m.out_features, device=m.weight.device)
                                        for m in (ours.inp, ours.mid, ours.out))
        self.recurrent = copy.deepcopy(ours.recurrent)
        for lstm in self.recurrent:
            lstm.flatten_parameters()                                 # deepcopy un-flattens the cuDNN weight buffer
        for mine, theirs in ((self.inp, ours.inp), (self.mid, ours.mid), (self.out, ours.out)):
            mine.load_state_dict(theirs.state_dict())
        self.sums = {}

    def forward(self, features):
        from oracle import aten_chain as C, np_oracle             # reference arm only
        n_frames, T = features['n_frames'], features['T']
        x = torch.cat((C.upsample_chain(features['normalised_lab'], features['dur']), features['normalised_counters']), dim=-1)
        h = torch.sigmoid(self.inp(x))
        order = torch.argsort(n_frames, descending=True)             # utils.py:366-369: sort, pack, run, unpack, unsort per layer
        for lstm in self.recurrent:
            packed = torch.nn.utils.rnn.pack_padded_sequence(h[order], n_frames[order].cpu(), batch_first=True)
            out, _ = lstm(packed)
            out, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=T)
            h = out[torch.argsort(order)]
        y = self.out(torch.sigmoid(self.mid(h)))
        outputs, column = {}, 0
        for name, width in OUTPUT_DIMS.items():
            block = y[:, :, column:column + width]
            column += width
            if name == 'vuv':
                outputs['vuv'] = torch.sigmoid(block)
                continue
            outputs['normalised_%s_deltas' % name] = block
            p = self.normalisers[name].delta_params_torch
            deltas = C.denormalise_mvn_chain(block, p['mean'], p['std_dev']).detach().cpu().numpy()     # RNN_SPSS.py:109-111
            traj = np_oracle.mlpg_banded(deltas, self.normalisers[name].delta_params['std_dev'] ** 2, padding_size=100)
            outputs[name] = torch.tensor(traj).type(block.dtype).to(block.device)                       # :116
        voiced = outputs['vuv'] > 0.5
        incs = [C.lf0_increment(features['lf0'], outputs['lf0'], voiced, n_frames),
                C._mean_increment((features['vuv'] == voiced).type(torch.float), n_frames),
                C.melcep_increment(features['mcep'], outputs['mcep'], n_frames),
                C.distortion_increment(features['bap'], outputs['bap'], n_frames)]
        for key, (s, c) in zip(('lf0', 'vuv', 'mcep', 'bap'), incs):
            old = self.sums.get(key, (0., 0.))
            self.sums[key] = (old[0] + s, old[1] + c)
        total = C.bce_chain(outputs['vuv'], features['vuv'].type(torch.float), n_frames)
        for name in STATIC:
            key = 'normalised_%s_deltas' % name
            total = total + C.mse_chain(outputs[key], features[key], n_frames)
        return total / 4., outputs


def stock_steps(model, ema_model, batches, steps, lr, decay):
    from oracle import aten_chain
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    shadows = [p.data for p in ema_model.parameters()]
    out = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for step in range(steps):
        opt.zero_grad()
        loss, _ = model(batches[step % len(batches)])
        loss.backward()
        opt.step()
        out.append(loss.item())                                      # experiment_builder.py:481
        aten_chain.ema_chain(shadows, [p.data for p in model.parameters()], decay)
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) / steps * 1e3


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=8, help='timed steps of the single-GPU comparison (config 4)')
    ap.add_argument('--batch-size', type=int, default=32)
    ap.add_argument('--utterances', type=int, default=0, help='config 5: size of the synthetic set sharded over the ranks; '
                    '0 = config 4 only')
    ap.add_argument('--epochs', type=int, default=1)
    ap.add_argument('--lr', type=float, default=0.002)
    ap.add_argument('--ema-decay', type=float, default=0.999)
    ap.add_argument('--num-layers', type=int, default=8)
    ap.add_argument('--no-pack', action='store_true', help='run the LSTMs on the padded batch (no packing)')
    ap.add_argument('--no-stock', action='store_true')
    ap.add_argument('--seed', type=int, default=1234)
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    normalisers = make_normalisers(dev, args.seed)
    proj = torch.randn(64, sum(STATIC.values()), generator=torch.Generator().manual_seed(args.seed)).to(dev) * 0.6

    def build():
        torch.manual_seed(args.seed)
        model = AcousticModel(normalisers, num_layers=args.num_layers, pack=not args.no_pack, device=dev)
        ema_model = copy.deepcopy(model)
        for lstm in ema_model.recurrent:
            lstm.flatten_parameters()                                 # deepcopy un-flattens the cuDNN weight buffer
        return model, ema_model

    line = {'config': 'C4/C5 LSTM acoustic model (609-512-8xLSTM512-256-187) training step with EMA, %d utterances / rank / step'
                      % args.batch_size, 'n_gpus': world}

    # ---- config 4: one GPU, the step and its phases, with the stock arm beside it --------------------------------------
    if world == 1:
        batches = [make_batch(i, args.batch_size, normalisers, dev, proj) for i in range(4)]
        model, ema_model = build()
        tr = trainer.DataParallelTrainer(model, ema_model=ema_model, ema_decay=args.ema_decay)
        opt = torch.optim.Adam(model.parameters(), lr=args.lr, fused=True)
        tr.train_epoch(batches[:3], opt)                              # warm-up: cuDNN plans, workspaces, allocator
        model, ema_model = build()
        tr = trainer.DataParallelTrainer(model, ema_model=ema_model, ema_decay=args.ema_decay)
        opt = torch.optim.Adam(model.parameters(), lr=args.lr, fused=True)
        stock = None if args.no_stock else StockModel(model, normalisers)
        stock_ema = copy.deepcopy(stock) if stock is not None else None
        steps = [batches[i % len(batches)] for i in range(args.steps)]
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ours_losses = []
        torch.cuda.synchronize()
        start.record()
        for i in range(args.steps):                                   # one-batch epochs so every step's loss is kept
            tr.epoch = i + 1
            model.mode = 'train'
            tr.bucket.zero()
            loss, _ = model(steps[i])
            loss.backward()
            tr.bucket.all_reduce()
            opt.step()
            tr.ema.update_params(model)
            model.metrics.accumulate('train', loss=loss.detach())
            ours_losses.append(loss.detach())
        stop.record()
        torch.cuda.synchronize()
        ms = start.elapsed_time(stop) / args.steps
        frames = sum(b['frames_total'] for b in steps)
        line.update({'ms_per_step': round(ms, 2), 'valid_frames_per_s': round(frames / (ms * args.steps) * 1e3),
                     'loss_first_last': [round(ours_losses[0].item(), 5), round(ours_losses[-1].item(), 5)],
                     'train_metrics': {k: round(float(v), 4) for k, v in model.metrics.results_as_json_dict('train').items()}})
        # phase breakdown of one more step (events between the phases; the sum is that step's device time)
        marks = []

        def phase(name):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))
        model.phase = phase
        model.mode = 'train'
        tr.bucket.zero()
        phase('start')
        loss, _ = model(batches[0])
        loss.backward()
        phase('backward')
        tr.bucket.all_reduce()
        opt.step()
        phase('adam')
        tr.ema.update_params(model)
        phase('ema')
        torch.cuda.synchronize()
        model.phase = None
        phases = {}
        for (_, a), (name, b) in zip(marks, marks[1:]):
            phases[name] = round(phases.get(name, 0.) + a.elapsed_time(b), 3)
        line['phase_ms'] = phases
        if stock is not None:
            stock_steps(stock, stock_ema, batches, 1, args.lr, args.ema_decay)   # warm-up
            stock = StockModel(build()[0], normalisers)
            stock_ema = copy.deepcopy(stock)
            for lstm in stock_ema.recurrent:
                lstm.flatten_parameters()
            stock_losses, stock_ms = stock_steps(stock, stock_ema, batches, args.steps, args.lr, args.ema_decay)
            rel = [abs(a.item() - b) / max(abs(b), 1e-6) for a, b in zip(ours_losses, stock_losses)]
            line.update({'stock_ms_per_step': round(stock_ms, 1), 'stock_loss_first_last': [round(stock_losses[0], 5), round(stock_losses[-1], 5)],
                         'rel_loss_difference_step0': round(rel[0], 6), 'max_rel_loss_difference': round(max(rel), 5),
                         'speedup_vs_stock': round(stock_ms / ms, 1),
                         'stock_note': 'reference op chain + fp32 torch.nn on the same GPU, MLPG on the host CPU (scipy banded '
                                       'solver for bandmat) with its device<->host copies; host cores: %d' % (os.cpu_count() or 1)})

    # ---- config 5: epochs over the sharded synthetic set through DataParallelTrainer --------------------------------
    if args.utterances:
        spans = trainer.rank_batches(args.utterances, args.batch_size)
        model, ema_model = build()
        tr = trainer.DataParallelTrainer(model, ema_model=ema_model, ema_decay=args.ema_decay)
        opt = torch.optim.Adam(model.parameters(), lr=args.lr, fused=True)
        # batch k of the global list = utterances [k * batch_size, (k + 1) * batch_size): the same data for any N
        loader = [make_batch(b // args.batch_size, args.batch_size, normalisers, dev, proj) for b, _ in spans]
        tr.train_epoch(loader[:2], opt)                               # warm-up, then start again
        model, ema_model = build()
        tr = trainer.DataParallelTrainer(model, ema_model=ema_model, ema_decay=args.ema_decay)
        opt = torch.optim.Adam(model.parameters(), lr=args.lr, fused=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        epoch_losses = []
        for tr.epoch in range(1, args.epochs + 1):
            epoch_losses.append(tr.train_epoch(loader, opt))
        stop.record()
        valid_loss = tr.valid_epoch(loader[:4], model=tr.ema.model)   # the averaged model, into its own metric handler
        torch.cuda.synchronize()
        stats = torch.tensor([start.elapsed_time(stop), float(sum(b['frames_total'] for b in loader) * args.epochs)],
                             dtype=torch.float64, device=dev)
        if world > 1:
            t, f = stats[:1].clone(), stats[1:].clone()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(f)
            stats = torch.cat([t, f])
        total_ms, frames = stats.tolist()
        line['dp'] = {'utterances': args.utterances, 'steps_per_epoch': len(loader), 'epochs': args.epochs,
                      'ms_per_step': round(total_ms / (len(loader) * args.epochs), 2),
                      'valid_frames_per_s': round(frames / total_ms * 1e3),
                      'epoch_losses': [round(x, 5) for x in epoch_losses], 'ema_valid_loss': round(valid_loss, 5),
                      'train_metrics': {k: round(float(v), 4) for k, v in model.metrics.results_as_json_dict('train').items()},
                      'ema_valid_metrics': {k: round(float(v), 4) for k, v in tr.ema.model.metrics.results_as_json_dict('valid').items()}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
