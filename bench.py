#!/usr/bin/env python
"""Benchmark of morgana's per-batch frame-rate feature path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one synthetic batch of configs[1] (256 utterances x ~60 phones, durations
U{1..30}, 600-dim labels; SURVEY.md section 8d, seed 1234):

    K1 duration scan -> K2 fused min-max normalise + phone->frame expansion to (B, T, 600)
    -> K4/K5 one-launch masked objective on the batch's (B, T, 187) prediction/target pair: 3 x mse + bce with the
       gradient w.r.t. the prediction, plus the four streaming metrics of models/RNN_SPSS.py:124-129.

`value` is whole-job valid frames/s with inputs resident in HBM (CUDA events, max over ranks); `e2e` is the same metric
through the public Python API from pinned HOST buffers with the host<->device copies inside the timed region;
`roofline` describes the dominant kernel (K2) from CUDA events around its launches inside the timed region
(`roofline_k4b`: the same for the step's other kernel);
`cpu_baseline` / `--impl reference` time the reference's own op chain (oracle/aten_chain.py) on the host cores.
Prints exactly one JSON line on rank 0.
"""
import argparse
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
CPU_SAMPLE_UTTS = 64          # utterances per CPU-baseline pass (a quarter of the batch)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch-size', type=int, default=256)
    ap.add_argument('--e2e-steps', type=int, default=0, help='steps of the host-buffer loop (0: min(steps, 30))')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


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

    def __init__(self, gpu_index, period_ms=100):
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
# reference arm / CPU baseline: the reference's op chain on the host cores
# ----------------------------------------------------------------------------------------------------------------------
def cpu_pass(sample):
    """One pass of the path as the reference executes it (oracle/aten_chain.py), on CPU tensors."""
    from oracle import aten_chain as ref
    norm_lab = ref.normalise_minmax_chain(sample['lab'], sample['mmin'], sample['mmax'])
    frames = ref.upsample_chain(norm_lab, sample['dur'])
    loss, grad, increments = ref.acoustic_loss_and_metrics(sample['pred'], sample['target'], sample['voiced'],
                                                           sample['n_frames'])
    return frames, loss, grad, increments


def make_cpu_sample(batch_size, n_utts):
    import torch
    from morgana_b200 import workloads
    ling = workloads.linguistic_batch(batch_size=batch_size, seed=1234)
    ac = workloads.acoustic_batch(ling['n_frames'], seed=1234)
    n = min(n_utts, batch_size)
    T = int(ling['n_frames'][:n].max())
    sample = {'lab': ling['lab'][:n].contiguous(), 'dur': ling['dur'][:n].contiguous(), 'mmin': ling['mmin'], 'mmax': ling['mmax'],
              'n_frames': ling['n_frames'][:n].contiguous(), 'pred': ac['pred'][:n, :T].contiguous(),
              'target': ac['target'][:n, :T].contiguous(), 'voiced': ac['voiced'][:n, :T].contiguous()}
    torch.set_num_threads(os.cpu_count() or 1)
    return sample, int(sample['n_frames'].sum())


def time_cpu(sample, steps, warmup):
    import torch
    for _ in range(warmup):
        cpu_pass(sample)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_pass(sample)
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
    if rank != 0:
        return
    import torch
    # Bounded sample: calibrate on 16 utterances, then size each step so that warm-up + K steps take about two minutes.
    probe, probe_frames = make_cpu_sample(args.batch_size, 16)
    cpu_pass(probe)
    t0 = time.perf_counter()
    cpu_pass(probe)
    per_utt = (time.perf_counter() - t0) / 16
    budget_s = 120.0
    n_utts = int(budget_s / max(args.steps + args.warmup, 1) / max(per_utt, 1e-6))
    n_utts = max(4, min(CPU_SAMPLE_UTTS, n_utts, args.batch_size))
    sample, frames = make_cpu_sample(args.batch_size, n_utts)
    times, threads = time_cpu(sample, args.steps, args.warmup)
    total = sum(times)
    value = frames * len(times) / total
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, note='each step = the first %d utterances of the batch on the host CPU' % n_utts),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': '%d of %d utterances (%d valid frames) per step; reference op chain restated in '
                                   'oracle/aten_chain.py (the Python reference cannot travel to the GPU box); %s, torch %s'
                                   % (n_utts, args.batch_size, frames, cpu_model_name(), torch.__version__)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def workload_config(args, note=None):
    cfg = {'workload': 'configs[1]: upsample_to_repetitions + minmax normalise, %d utts x ~60 phones, dur U{1..30}, '
                       '600-dim labels -> masked loss + metrics on 187-dim WORLD targets of the same batch' % args.batch_size,
           'batch_utterances_per_gpu': args.batch_size, 'label_dim': 600, 'target_dim': 187, 'seed': 1234,
           'l2': 'inputs larger than L2: each step streams ~1.9 GB (837 MB output alone) and rotates over %d input batches'
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

    def step(batch, time_k2=None, time_k4b=None):
        # K1 + K2 (max_len: the padded length is known on the host, as features['n_frames'] is in the reference pipeline)
        if time_k2 is not None:
            ends, n_frames, _ = ops.dur_scan(batch['dur'])   # keep K1 outside the bracket: the events time K2 alone
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            B, P, D = batch['lab'].shape
            out = torch.empty((B, batch['T'], D), dtype=torch.float32, device=dev)
            e0.record(stream)
            ops.check(ops.lib.mg_upsample_norm_f32(batch['lab'].data_ptr(), batch['lab'].stride(0), batch['lab'].stride(1),
                                                   ends.data_ptr(), mmin.data_ptr(), mmax.data_ptr(), 0, 2, out.data_ptr(),
                                                   B, P, D, batch['T'], 0, stream.cuda_stream), 'mg_upsample_norm_f32')
            e1.record(stream)
            time_k2.append((e0, e1))
        else:
            out, n_frames = mg.utils.upsample_to_repetitions(batch['lab'], batch['dur'], normaliser=normaliser,
                                                             max_len=batch['T'], return_lengths=True)
        if time_k4b is not None:   # K4b (+ the 2 us fill of its result records) between its own pair of events
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            o0.record(stream)
            loss, grad = objective(batch['pred'], batch['target'], n_frames)
            o1.record(stream)
            time_k4b.append((o0, o1))
        else:
            loss, grad = objective(batch['pred'], batch['target'], n_frames)
        return out, loss, grad

    pending = []

    def exchange():
        """The path's one collective: SUM of the packed loss / metric-sum records over ranks (NCCL over NVLink).  It is
        issued asynchronously -- the records are copied first, the all-reduce runs on NCCL's stream while the next step's
        kernels run on ours -- and joined one step later (and before the timed region closes)."""
        if world > 1:
            join()
            pending.append(dp.allreduce_records(objective.last_loss_records, objective._records, async_op=True))

    def join():
        while pending:
            pending.pop().result()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ---------------------------------------------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        step(dev_batches[i % N_ROTATING_BATCHES])
        exchange()
    barrier()

    # ---- timed region: exactly K steps, device-resident inputs ---------------------------------------------------------
    # A pair of event records costs ~3 us of the stream's time: K2 is bracketed on every 4th step and K4b on every 8th of a
    # long run (every step of a short one) -- the averages are still taken live, inside the timed region.
    k2_events, k4b_events, k2_steps, k4b_steps = [], [], [], []
    k2_every, k4b_every = (4, 8) if args.steps >= 64 else (1, 1)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    frames_done = 0
    barrier()
    wall0 = time.time()
    start.record(stream)
    for i in range(args.steps):
        batch = dev_batches[i % N_ROTATING_BATCHES]
        time_k2, time_k4b = i % k2_every == 0, i % k4b_every == 0
        if time_k2:
            k2_steps.append(i)
        if time_k4b:
            k4b_steps.append(i)
        step(batch, time_k2=k2_events if time_k2 else None, time_k4b=k4b_events if time_k4b else None)
        exchange()
        frames_done += batch['frames']
    join()
    stop.record(stream)
    barrier()
    wall1 = time.time()
    elapsed_ms = start.elapsed_time(stop)

    k2_ms = [a.elapsed_time(b) for a, b in k2_events]
    k2_avg_ms = sum(k2_ms) / len(k2_ms)
    k2_bytes = []
    for i in k2_steps:
        hb = host_batches[i % N_ROTATING_BATCHES]
        k2_bytes.append(4 * 600 * (args.batch_size * hb['T'] + hb['n_phones']) + 4 * args.batch_size * hb['P'] + 8 * 600)
    k2_avg_bytes = sum(k2_bytes) / len(k2_bytes)
    k4b_avg_ms = sum(a.elapsed_time(b) for a, b in k4b_events) / len(k4b_events)
    k4b_avg_bytes = sum(4 * 187 * (2 * host_batches[i % N_ROTATING_BATCHES]['frames'] +
                                   args.batch_size * host_batches[i % N_ROTATING_BATCHES]['T'])
                        for i in k4b_steps) / len(k4b_steps)   # valid rows of pred + target read, the whole gradient written

    # ---- end-to-end: the public API from pinned host buffers, copies inside the timed region ------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 30)
    e2e_keys = ('lab_packed', 'dur_packed', 'pred_packed', 'target_packed', 'phone_counts', 'frame_counts')
    h2d = sum(host_batches[0][k].numel() * host_batches[0][k].element_size() for k in e2e_keys)
    d2h = 8 * ops.RESULT_BYTES

    copy_stream = torch.cuda.Stream(device=dev)

    def upload(hb):
        """Host -> device on the copy stream: only valid rows cross PCIe (packed wire format)."""
        with torch.cuda.stream(copy_stream):
            up = {k: hb[k].to(dev, non_blocking=True) for k in e2e_keys}
            done = torch.cuda.Event()
            done.record(copy_stream)
        return up, done

    def e2e_compute(hb, up, done):
        # The zero padding of collate_fn (reference data.py:184-193) is produced on the device.  Lengths are host-side
        # knowledge (features['n_frames']), so nothing synchronises until the result records are read back.
        stream.wait_event(done)
        for t in up.values():
            t.record_stream(stream)
        pred = mg.data.pad_collate(up['pred_packed'], up['frame_counts'], max_len=hb['T'])
        target = mg.data.pad_collate(up['target_packed'], up['frame_counts'], max_len=hb['T'])
        # the items are consumed as they arrive (packed): no phone padding is built or read
        out, n_frames = mg.utils.upsample_packed_to_repetitions(up['lab_packed'], up['dur_packed'], up['phone_counts'],
                                                                normaliser=normaliser, max_len=hb['T'], max_items=hb['P'],
                                                                return_lengths=True)
        loss, grad = objective(pred, target, n_frames)
        if world > 1:     # the step's result is read back right away, so this exchange is joined at once
            return dp.allreduce_records(objective.last_loss_records, objective._records).cpu()
        return torch.cat([objective.last_loss_records, objective._records]).cpu()   # device -> host read of the result

    def e2e_loop(n_steps):
        """Every step's inputs are uploaded inside the loop; the upload of step i + 1 overlaps the kernels of step i."""
        frames = 0
        pending = upload(host_batches[0])
        for i in range(n_steps):
            hb = host_batches[i % N_ROTATING_BATCHES]
            up, done = pending
            if i + 1 < n_steps:
                pending = upload(host_batches[(i + 1) % N_ROTATING_BATCHES])
            e2e_compute(hb, up, done)
            frames += hb['frames']
        return frames

    e2e_loop(3)
    barrier()
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ewall0 = time.time()
    e_start.record(stream)
    copy_stream.wait_event(e_start)          # the first upload belongs to the timed region
    e2e_frames = e2e_loop(e2e_steps)
    e_stop.record(stream)
    barrier()
    ewall1 = time.time()
    e2e_ms = e_start.elapsed_time(e_stop)

    # ---- max over ranks -------------------------------------------------------------------------------------------
    stats = torch.tensor([elapsed_ms, e2e_ms, float(frames_done), float(e2e_frames)], dtype=torch.float64, device=dev)
    if world > 1:
        times = stats[:2].clone()
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        counts = stats[2:].clone()
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        elapsed_ms, e2e_ms = times.tolist()
        frames_done, e2e_frames = counts.tolist()
    if sampler is not None:
        sampler.stop()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    achieved = k2_avg_bytes / (k2_avg_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    traffic_path = os.path.join(ROOT, 'profiles', 'k2_traffic.json')
    if os.path.exists(traffic_path):       # DRAM bytes of one K2 launch from the committed `ncu --set full` capture
        with open(traffic_path) as f:
            t = json.load(f)
        traffic, traffic_src = t['dram_bytes_read'] + t['dram_bytes_write'], t['source']
    line = {
        'metric': METRIC, 'value': frames_done / (elapsed_ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args),
        'clocks': sampler.summary([(wall0, wall1), (ewall0, ewall1)]),
        'e2e': {'value': e2e_frames / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'steps': e2e_steps, 'ms_per_step': e2e_ms / e2e_steps,
                'api': 'utils.upsample_packed_to_repetitions(packed lab, packed dur, n_phones, normaliser=..., max_len=T) | data.pad_collate(packed pred / target) -> fused.AcousticObjective, from pinned host tensors'},
        'gpu_launches': 3 * args.steps, 'e2e_gpu_launches_per_step': 8,
        'roofline': {'bound': 'hbm', 'kernel': 'upsample_bulk_kernel<MINMAX> (K2, fused normalise + expansion)',
                     'achieved': achieved, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': achieved / pk['hbm_gbs'],
                     'frac_of_8000_nominal': achieved / 8000.0, 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': pk_src,
                     'algorithmic_bytes_per_launch': k2_avg_bytes, 'avg_launch_ms': k2_avg_ms, 'launches_timed': len(k2_ms),
                     'share_of_step': k2_avg_ms / (elapsed_ms / args.steps)},
        # the step's other kernel, for the record (same method: CUDA events on the launching stream, algorithmic bytes)
        'roofline_k4b': {'bound': 'hbm', 'kernel': 'masked_objective_kernel<GRAD> (K4b: 3 x mse + bce + gradient + 4 metrics)',
                         'achieved': k4b_avg_bytes / (k4b_avg_ms * 1e-3) / 1e9, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                         'frac': k4b_avg_bytes / (k4b_avg_ms * 1e-3) / 1e9 / pk['hbm_gbs'],
                         'algorithmic_bytes_per_launch': k4b_avg_bytes, 'avg_launch_ms': k4b_avg_ms, 'launches_timed': len(k4b_events),
                         'share_of_step': k4b_avg_ms / (elapsed_ms / args.steps)},
    }
    if world == 1 and not args.no_cpu_baseline:
        sample, frames = make_cpu_sample(args.batch_size, CPU_SAMPLE_UTTS)
        times, threads = time_cpu(sample, steps=8, warmup=1)
        line['cpu_baseline'] = {'value': frames / min(times), 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                'sample': '%d of %d utterances (%d valid frames), best of %d passes of the reference op '
                                          'chain (oracle/aten_chain.py) on %s' % (CPU_SAMPLE_UTTS, args.batch_size, frames,
                                                                                 len(times), cpu_model_name())}
    emit(line)
    if world > 1:
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
