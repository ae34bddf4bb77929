#!/usr/bin/env python
"""Benchmark of morgana's per-batch frame-rate feature path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one synthetic batch of configs[1] (256 utterances x ~60 phones, durations
U{1..30}, 600-dim labels; SURVEY.md section 8d, seed 1234), through the package's public API:

    utils.upsample_to_repetitions(lab, dur, normaliser=minmax)  = K1 duration scan + K2 fused normalise + expansion -> (B, T, 600)
    fused.AcousticObjective(pred, target, n_frames)             = K4b: 3 x mse + bce of models/RNN_SPSS.py:131-139 with the
                                                                  gradient, and its four streaming metrics (:124-129)

`value`   whole-job valid frames/s with the inputs resident in HBM (CUDA events on the launching stream, max over ranks).
`e2e`     the same metric through the same API from pinned HOST buffers, host<->device copies inside the timed region.
`roofline` the kernel with the largest share of the step, from CUDA events around its launches inside the timed region
          (`ops.KernelProbe`); the step's other kernel is reported the same way beside it.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference (ZackHodari/morgana, mirrored into oracle/_ref by
          oracle/make_ref.py) on the box's host cores, same 256 utterances; that arm never imports the product package.
`other_configs` (N = 1): the remaining BASELINE.json configurations and the stock-ATen path on the same GPU, each timed with its
          own CUDA events inside this run.  `training` (every N): BASELINE.json configs[4], data-parallel training of the README
          F0 MLP with the flat-bucket gradient all-reduce.
Prints exactly one JSON line on rank 0.
"""
import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'valid frames/sec through normalise->upsample->masked-loss path'
UNIT = 'frames/s'
N_ROTATING_BATCHES = 3        # distinct input batches cycled between steps (each step's working set is ~1.9 GB >> L2)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=400)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch-size', type=int, default=256)
    ap.add_argument('--e2e-steps', type=int, default=0, help='steps of the host-buffer loop (0: min(steps, 30))')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip other_configs and the training section')
    return ap.parse_args()


def load_by_path(name, *relative):
    """Import a source file without importing the package around it (the reference arm must not load the product)."""
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, *relative))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback (B200_PROFILING.md)'


# ----------------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi polled in the background; samples are stamped with time.time() so they can be windowed."""
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index, period_ms=50):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits',
                 '-lms', str(period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) >= 7:
                self.samples.append((time.time(), parts))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, windows):
        def in_windows(ts):
            return any(a <= ts <= b for a, b in windows)
        chosen = [p for ts, p in self.samples if in_windows(ts)]
        scope = 'timed regions'
        if not chosen:
            chosen, scope = [p for _, p in self.samples], 'whole run (timed regions shorter than the sampling period)'
        if not chosen:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0, 'scope': 'nvidia-smi unavailable'}
        sm, sm_max, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for p in chosen:
            try:
                sm.append(float(p[0]))
                sm_max.append(float(p[1]))
            except ValueError:
                continue
            for name, flag in zip(names, p[3:7]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(sm_max) if sm_max else None,
                'reasons': sorted(reasons), 'samples': len(chosen), 'scope': scope}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference on the host cores
# ----------------------------------------------------------------------------------------------------------------------
WORLD_LAYOUT = ((0, 3), (3, 4), (4, 184), (184, 187))     # lf0 | vuv | mcep | bap after torch.split (models/RNN_SPSS.py:86-88)


class ReferencePath(object):
    """One pass of the path as the reference executes it, with the reference's own functions (CPU or CUDA tensors):
    ``data.normalise_minmax`` (data.py:579-584) -> ``utils.upsample_to_repetitions`` (utils.py:175-228) -> the body of
    ``LSTMAcousticModel.loss`` (models/RNN_SPSS.py:120-139): four metric accumulations, 3 x ``losses.mse`` + ``losses.bce``,
    ``/ 4``, and ``backward()`` for the gradient w.r.t. the prediction."""
    def __init__(self):
        from oracle import ref_loader
        self.available = ref_loader.available()
        self.kind = 'reference' if self.available else 'port'
        if self.available:
            self.morgana = ref_loader.import_reference()
            self.where = os.path.relpath(ref_loader.reference_root(), ROOT) if ref_loader.reference_root().startswith(ROOT) \
                else ref_loader.reference_root()
            m = self.morgana.metrics
            self.metrics = {'LF0_RMSE_Hz': m.LF0Distortion(), 'VUV_accuracy': m.Mean(), 'MCEP_distortion': m.MelCepDistortion(),
                            'BAP_distortion': m.Distortion()}
            for metric in self.metrics.values():
                metric.reset_state()
        else:    # the mirror did not travel: the restatement of the same op chain (oracle/aten_chain.py)
            from oracle import aten_chain
            self.port, self.where = aten_chain, 'oracle/aten_chain.py'

    def upsample(self, lab, dur, mmin, mmax):
        if not self.available:
            return self.port.upsample_chain(self.port.normalise_minmax_chain(lab, mmin, mmax), dur)
        return self.morgana.utils.upsample_to_repetitions(self.morgana.data.normalise_minmax(lab, mmin, mmax), dur)

    def objective(self, pred, target, n_frames, voiced_target):
        import torch
        if not self.available:
            return self.port.acoustic_loss_and_metrics(pred, target, voiced_target, n_frames)[:2]
        losses, (lf0, vuv, mcep, bap) = self.morgana.losses, WORLD_LAYOUT
        pred = pred.detach().requires_grad_()
        vuv_pred = pred[..., vuv[0]:vuv[1]] > 0.5
        with torch.no_grad():
            self.metrics['LF0_RMSE_Hz'].accumulate(target[..., 0:1], pred[..., 0:1], vuv_pred.clone(), n_frames)
            self.metrics['VUV_accuracy'].accumulate((voiced_target == vuv_pred).type(torch.float), n_frames)
            self.metrics['MCEP_distortion'].accumulate(target[..., 4:64], pred[..., 4:64], n_frames)
            self.metrics['BAP_distortion'].accumulate(target[..., 184:185], pred[..., 184:185], n_frames)
        loss = 0.
        for a, b in (lf0, mcep, bap):
            loss += losses.mse(pred[..., a:b], target[..., a:b], n_frames)
        loss += losses.bce(pred[..., vuv[0]:vuv[1]], target[..., vuv[0]:vuv[1]], n_frames)
        loss = loss / 4.
        loss.backward()
        return loss.detach(), pred.grad

    def step(self, sample):
        frames = self.upsample(sample['lab'], sample['dur'], sample['mmin'], sample['mmax'])
        loss, grad = self.objective(sample['pred'], sample['target'], sample['n_frames'], sample['voiced'])
        return frames, loss, grad


def make_cpu_sample(workloads, batch_size, n_utts):
    import torch
    ling = workloads.linguistic_batch(batch_size=batch_size, seed=1234)
    ac = workloads.acoustic_batch(ling['n_frames'], seed=1234)
    n = min(n_utts, batch_size)
    T = int(ling['n_frames'][:n].max())
    sample = {'lab': ling['lab'][:n].contiguous(), 'dur': ling['dur'][:n].contiguous(), 'mmin': ling['mmin'], 'mmax': ling['mmax'],
              'n_frames': ling['n_frames'][:n].contiguous(), 'pred': ac['pred'][:n, :T].contiguous(),
              'target': ac['target'][:n, :T].contiguous(), 'voiced': ac['voiced'][:n, :T].contiguous()}
    torch.set_num_threads(os.cpu_count() or 1)
    return sample, int(sample['n_frames'].sum())


def time_cpu(path, sample, steps, warmup):
    import torch
    for _ in range(warmup):
        path.step(sample)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        path.step(sample)
        times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def cpu_model_name():
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('model name'):
                    return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


def run_reference(args, rank, world):
    """`--impl reference`: rank 0 alone times the reference on the host cores; nothing of the product package is imported."""
    if rank != 0:
        return
    import torch
    workloads = load_by_path('_mg_workloads', 'morgana_b200', 'workloads.py')     # the file, not the package
    path = ReferencePath()
    # The whole batch of the GPU arm (256 utterances) per step; shrink only if K + W passes would not end within ~2.5 minutes.
    probe, _ = make_cpu_sample(workloads, args.batch_size, 32)
    path.step(probe)
    t0 = time.perf_counter()
    path.step(probe)
    per_utt = (time.perf_counter() - t0) / 32
    budget_s = 150.0
    n_utts = int(budget_s / max(args.steps + args.warmup, 1) / max(per_utt, 1e-6))
    n_utts = max(8, min(args.batch_size, n_utts))
    sample, frames = make_cpu_sample(workloads, args.batch_size, n_utts)
    times, threads = time_cpu(path, sample, args.steps, args.warmup)
    total = sum(times)
    value = frames * len(times) / total
    assert 'morgana_b200' not in sys.modules, 'the reference arm imported the product package'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, note=None if n_utts == args.batch_size else
                                  'each step = the first %d utterances of the batch (time budget)' % n_utts),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': path.kind,
                         'sample': '%d of %d utterances (%d valid frames) per step; %s, imported from %s, on %s, torch %s'
                                   % (n_utts, args.batch_size, frames,
                                      'the unmodified reference functions (data.normalise_minmax, utils.upsample_to_repetitions, '
                                      'losses.mse / bce + backward, metrics.*.accumulate)' if path.available else
                                      'restatement of the reference op chain', path.where, cpu_model_name(), torch.__version__)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def workload_config(args, note=None):
    cfg = {'workload': 'configs[1]: upsample_to_repetitions + minmax normalise, %d utts x ~60 phones, dur U{1..30}, '
                       '600-dim labels -> masked loss + metrics on 187-dim WORLD targets of the same batch' % args.batch_size,
           'batch_utterances_per_gpu': args.batch_size, 'label_dim': 600, 'target_dim': 187, 'seed': 1234,
           'l2': 'inputs larger than L2: each step streams ~1.5 GB (837 MB output alone) and rotates over %d input batches'
                 % N_ROTATING_BATCHES,
           'parallelism': 'utterance-sharded, one process per GPU'}
    if note:
        cfg['note'] = note
    return cfg


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import morgana_b200 as mg
    from morgana_b200 import dp, ops, workloads
    from morgana_b200.fused import AcousticObjective

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    windows = []

    # ---- inputs: each rank owns its own shard of utterances (different seeds), resident in HBM -------------------
    host_batches, dev_batches = [], []
    for i in range(N_ROTATING_BATCHES):
        seed = 1234 + 7919 * i + 104729 * rank
        ling = workloads.linguistic_batch(batch_size=args.batch_size, seed=seed)
        ac = workloads.acoustic_batch(ling['n_frames'], seed=seed)
        valid_phone = torch.arange(ling['dur'].shape[1])[None, :] < ling['n_phones'][:, None]
        valid_frame = torch.arange(int(ling['n_frames'].max()))[None, :] < ling['n_frames'][:, None]
        hb = {'lab': ling['lab'].pin_memory(), 'dur': ling['dur'].pin_memory(), 'pred': ac['pred'].pin_memory(),
              'target': ac['target'].pin_memory(), 'T': int(ling['n_frames'].max()), 'frames': int(ling['n_frames'].sum()),
              'n_phones': int(ling['n_phones'].sum()), 'P': ling['dur'].shape[1],
              # packed (ragged) wire format for the end-to-end loop: valid rows only, padded on the device (K0)
              'lab_packed': ling['lab'][valid_phone].contiguous().pin_memory(),
              'dur_packed': ling['dur'][:, :, 0][valid_phone].contiguous().pin_memory(),
              'pred_packed': ac['pred'][valid_frame].contiguous().pin_memory(),
              'target_packed': ac['target'][valid_frame].contiguous().pin_memory(),
              'phone_counts': ling['n_phones'].pin_memory(), 'frame_counts': ling['n_frames'].pin_memory()}
        host_batches.append(hb)
        dev_batches.append({k: (v.to(dev) if isinstance(v, torch.Tensor) and not k.endswith(('_packed', '_counts')) else v)
                            for k, v in hb.items()})
        if i == 0:
            mmin, mmax = ling['mmin'].to(dev), ling['mmax'].to(dev)
    normaliser = ('minmax', mmin, mmax)
    objective = AcousticObjective()
    stream = torch.cuda.current_stream()
    # one row of four 48-byte loss records per step: K4b writes step i's losses into row i, nothing is accumulated or copied
    # per step; the rows and the four running metric records cross the ranks in ONE all-reduce when the epoch ends, as the
    # reference only needs epoch sums (experiment_builder.py:499-501)
    loss_log = ops.new_result_records(4 * max(args.steps, 1), dev).reshape(max(args.steps, 1), 4, -1)
    scratch_records = ops.new_result_records(4, dev)

    def step(batch, loss_records):
        # max_len: the padded length is known on the host, as features['n_frames'] is in the reference pipeline
        out, n_frames = mg.utils.upsample_to_repetitions(batch['lab'], batch['dur'], normaliser=normaliser,
                                                         max_len=batch['T'], return_lengths=True)
        loss, grad = objective(batch['pred'], batch['target'], n_frames, loss_records=loss_records)
        return out, loss, grad

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ---------------------------------------------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        step(dev_batches[i % N_ROTATING_BATCHES], scratch_records)
    if world > 1:     # NCCL communicator and protocol set-up for this message size outside the timed region
        dp.allreduce_records(loss_log.reshape(-1, loss_log.shape[-1]), objective._records)
    barrier()

    # ---- timed region: exactly K steps, device-resident inputs ---------------------------------------------------------
    # A pair of event records costs ~3 us of the stream's time and keeps the bracketed launch from overlapping its neighbours
    # (programmatic dependent launch), so the probe brackets K2 on steps 0, 8, 16, ... and K4b on steps 4, 12, 20, ... (every 4th
    # step each in runs shorter than 32 steps); the averages are still taken live, inside the timed region, on the launching stream.
    k2_probe, k4b_probe = ops.KernelProbe('K2'), ops.KernelProbe('K4b')
    k2_steps, k4b_steps = [], []
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    frames_done = 0
    period = 8 if args.steps >= 32 else 4
    barrier()
    wall0 = time.time()
    start.record(stream)
    for i in range(args.steps):
        batch = dev_batches[i % N_ROTATING_BATCHES]
        if i % period == 0:
            k2_steps.append(i)
            with k2_probe:
                step(batch, loss_log[i])
        elif i % period == period // 2:
            k4b_steps.append(i)
            with k4b_probe:
                step(batch, loss_log[i])
        else:
            step(batch, loss_log[i])
        frames_done += batch['frames']
    # the path's one collective: SUM of the epoch's loss records and metric records over the ranks (NCCL over NVLink)
    epoch_records = dp.allreduce_records(loss_log.reshape(-1, loss_log.shape[-1]), objective._records) if world > 1 else None
    stop.record(stream)
    barrier()
    wall1 = time.time()
    windows.append((wall0, wall1))
    elapsed_ms = start.elapsed_time(stop)
    del epoch_records

    def kernel_line(name, probe, steps_timed, bytes_of):
        ms = probe.ms(name)
        avg_ms = sum(ms) / len(ms)
        avg_bytes = sum(bytes_of(host_batches[i % N_ROTATING_BATCHES]) for i in steps_timed) / len(steps_timed)
        return avg_ms, avg_bytes, len(ms)

    k2_ms, k2_bytes, k2_n = kernel_line('K2', k2_probe, k2_steps, lambda hb: 4 * 600 * (args.batch_size * hb['T'] + hb['n_phones']) +
                                        4 * args.batch_size * hb['P'] + 8 * 600)
    # valid rows of pred + target read, the whole gradient written
    k4b_ms, k4b_bytes, k4b_n = kernel_line('K4b', k4b_probe, k4b_steps or k2_steps, lambda hb: 4 * 187 * (2 * hb['frames'] + args.batch_size * hb['T'])) \
        if k4b_steps else (float('nan'), 0., 0)

    # ---- end-to-end: the public API from pinned host buffers, copies inside the timed region ------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 30)
    copy_stream = torch.cuda.Stream(device=dev)

    def run_e2e(keys, label):
        """Every step's inputs are uploaded inside the loop (one cudaMemcpyAsync per buffer, on a copy stream); the upload of
        step i + 1 overlaps the kernels of step i; the step's result records are read back to the host every step."""
        h2d = sum(host_batches[0][k].numel() * host_batches[0][k].element_size() for k in keys)

        def upload(hb):
            with torch.cuda.stream(copy_stream):
                up = {k: hb[k].to(dev, non_blocking=True) for k in keys}
                done = torch.cuda.Event()
                done.record(copy_stream)
            return up, done

        def compute(hb, db, up, done):
            # The zero padding of collate_fn (reference data.py:184-193) is produced on the device.  Lengths are host-side
            # knowledge (features['n_frames']), so nothing synchronises until the result records are read back.
            stream.wait_event(done)
            for t in up.values():
                t.record_stream(stream)
            if 'pred_packed' in up:
                pred = mg.data.pad_collate(up['pred_packed'], up['frame_counts'], max_len=hb['T'])
            else:
                pred = db['pred']            # produced on the device by the model in the real pipeline (RNN_SPSS.py:83)
            target = mg.data.pad_collate(up['target_packed'], up['frame_counts'], max_len=hb['T'])
            # the items are consumed as they arrive (packed): no phone padding is built or read
            out, n_frames = mg.utils.upsample_packed_to_repetitions(up['lab_packed'], up['dur_packed'], up['phone_counts'],
                                                                    normaliser=normaliser, max_len=hb['T'], max_items=hb['P'],
                                                                    return_lengths=True)
            loss, grad = objective(pred, target, n_frames)
            if world > 1:     # the step's result is read back right away, so this exchange is joined at once
                return dp.allreduce_records(objective.last_loss_records, objective._records).cpu()
            return torch.cat([objective.last_loss_records, objective._records]).cpu()   # device -> host read of the result

        def loop(n_steps):
            frames = 0
            pending = upload(host_batches[0])
            for i in range(n_steps):
                hb, db = host_batches[i % N_ROTATING_BATCHES], dev_batches[i % N_ROTATING_BATCHES]
                up, done = pending
                if i + 1 < n_steps:
                    pending = upload(host_batches[(i + 1) % N_ROTATING_BATCHES])
                compute(hb, db, up, done)
                frames += hb['frames']
            return frames

        loop(3)
        barrier()
        e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e_start.record(stream)
        copy_stream.wait_event(e_start)          # the first upload belongs to the timed region
        frames = loop(e2e_steps)
        e_stop.record(stream)
        barrier()
        windows.append((w0, time.time()))
        return {'ms': e_start.elapsed_time(e_stop), 'frames': frames, 'h2d': h2d, 'label': label}

    all_keys = ('lab_packed', 'dur_packed', 'pred_packed', 'target_packed', 'phone_counts', 'frame_counts')
    e2e_full = run_e2e(all_keys, 'lab + dur + target + pred from the host')
    # `pred` is produced on the device by the model in the real pipeline; the host-side inputs of the path are lab / dur /
    # target / lengths -- reported beside the headline number, which keeps everything on the wire
    e2e_inputs = run_e2e(tuple(k for k in all_keys if k != 'pred_packed'), 'lab + dur + target from the host, pred resident')
    d2h = 8 * ops.RESULT_BYTES

    # ---- max over ranks -------------------------------------------------------------------------------------------
    stats = torch.tensor([elapsed_ms, e2e_full['ms'], e2e_inputs['ms'], float(frames_done), float(e2e_full['frames']),
                          float(e2e_inputs['frames'])], dtype=torch.float64, device=dev)
    if world > 1:
        times = stats[:3].clone()
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        counts = stats[3:].clone()
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        stats = torch.cat([times, counts])
    elapsed_ms, e2e_ms, e2e_in_ms, frames_done, e2e_frames, e2e_in_frames = stats.tolist()

    pk, pk_src = peaks()
    sections = None
    extras = {}
    if not args.no_extras:
        sections = load_by_path('_mg_bench_sections', 'scripts', 'bench_sections.py')
        # the extra sections never cost the headline line: a failure is reported in place (N > 1: every rank must agree, so the
        # training section, which holds collectives, is only guarded at N = 1)
        t0 = time.time()
        if world == 1:
            try:
                extras['training'] = sections.training_section(rank, world, dev, pk)
            except Exception as exc:       # noqa: BLE001
                extras['training'] = {'error': '%s: %s' % (type(exc).__name__, str(exc).splitlines()[0][:300])}
        else:
            extras['training'] = sections.training_section(rank, world, dev, pk)
        windows.append((t0, time.time()))
        if world == 1:
            t0 = time.time()
            try:
                extras['other_configs'] = sections.other_configs(dev, pk, ReferencePath())
            except Exception as exc:       # noqa: BLE001
                extras['other_configs'] = {'error': '%s: %s' % (type(exc).__name__, str(exc).splitlines()[0][:300])}
            windows.append((t0, time.time()))
    if sampler is not None:
        sampler.stop()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    step_ms = elapsed_ms / args.steps
    traffic, traffic_src = {}, {}
    for name, fname in (('K2', 'k2_traffic.json'), ('K4b', 'k4b_traffic.json')):
        path = os.path.join(ROOT, 'profiles', fname)
        if os.path.exists(path):       # DRAM bytes of one launch from the committed `ncu --set full` capture
            with open(path) as f:
                t = json.load(f)
            traffic[name], traffic_src[name] = t['dram_bytes_read'] + t['dram_bytes_write'], t['source']

    def roofline(name, kernel, ms, nbytes, n):
        achieved = nbytes / (ms * 1e-3) / 1e9
        return {'bound': 'hbm', 'kernel': kernel, 'achieved': achieved, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                'frac': achieved / pk['hbm_gbs'], 'frac_of_8000_nominal': achieved / 8000.0, 'traffic': traffic.get(name),
                'traffic_source': traffic_src.get(name), 'peak_source': pk_src, 'algorithmic_bytes_per_launch': nbytes,
                'avg_launch_ms': ms, 'launches_timed': n, 'share_of_step': ms / step_ms}

    lines = {'K2': roofline('K2', 'upsample_bulk_kernel<MINMAX> (K2, fused normalise + expansion)', k2_ms, k2_bytes, k2_n)}
    if k4b_n:
        lines['K4b'] = roofline('K4b', 'objective_stream_kernel<GRAD> (K4b: 3 x mse + bce + gradient + 4 metrics)', k4b_ms, k4b_bytes, k4b_n)
    dominant = max(lines, key=lambda k: lines[k]['share_of_step'])
    step_bytes = k2_bytes + (k4b_bytes if k4b_n else 0.)
    line = {
        'metric': METRIC, 'value': frames_done / (elapsed_ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': step_ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args),
        'clocks': sampler.summary(windows),
        'e2e': {'value': e2e_frames / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': e2e_full['h2d'], 'd2h_bytes_per_step': d2h,
                'steps': e2e_steps, 'ms_per_step': e2e_ms / e2e_steps,
                'h2d_gb_per_s_per_gpu': e2e_full['h2d'] / (e2e_ms / e2e_steps * 1e-3) / 1e9,
                'wire': e2e_full['label'],
                'api': 'utils.upsample_packed_to_repetitions(packed lab, packed dur, n_phones, normaliser=..., max_len=T) | '
                       'data.pad_collate(packed pred / target) -> fused.AcousticObjective, from pinned host tensors',
                'inputs_only': {'value': e2e_in_frames / (e2e_in_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': e2e_inputs['h2d'],
                                'ms_per_step': e2e_in_ms / e2e_steps, 'wire': e2e_inputs['label']}},
        'gpu_launches': 3 * args.steps, 'e2e_gpu_launches_per_step': 8,
        'step': {'algorithmic_bytes': step_bytes, 'achieved_gb_per_s': step_bytes / (step_ms * 1e-3) / 1e9,
                 'frac_of_measured_hbm_peak': step_bytes / (step_ms * 1e-3) / 1e9 / pk['hbm_gbs'],
                 'frac_of_8000_nominal': step_bytes / (step_ms * 1e-3) / 1e9 / 8000.0,
                 'kernels': 'K1 dur_scan + K2 + K4b per step, through utils.upsample_to_repetitions and fused.AcousticObjective; '
                            'one all-reduce of the loss / metric records per epoch (N > 1)'},
        'roofline': lines[dominant],
    }
    for name, entry in lines.items():
        if name != dominant:
            line['roofline_' + name.lower()] = entry
    line.update(extras)
    if world == 1 and not args.no_cpu_baseline:
        path = ReferencePath()
        sample, frames = make_cpu_sample(workloads, args.batch_size, args.batch_size)
        times, threads = time_cpu(path, sample, steps=5, warmup=1)
        line['cpu_baseline'] = {'value': frames / min(times), 'unit': UNIT, 'cores': threads, 'kind': path.kind,
                                'sample': 'all %d utterances (%d valid frames), best of %d passes of %s (imported from %s) on %s'
                                          % (args.batch_size, frames, len(times),
                                             'the unmodified reference functions' if path.available else 'the restated op chain',
                                             path.where, cpu_model_name())}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_RESULT_FD = None


def protect_stdout():
    """Keep stdout for the ONE JSON line: libraries that print there (NCCL writes its version banner to stdout when
    NCCL_DEBUG is set in the environment) are pointed at stderr for the rest of the run."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + '\n').encode()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, data)


def main():
    args = parse_args()
    protect_stdout()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
