#!/usr/bin/env python
"""Host -> device copy bandwidth per rank, alone and with all ranks copying at once (the limiter of `e2e` at N >= 4).

    python scripts/h2d_bandwidth.py                                              # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/h2d_bandwidth.py

Every rank owns a 256 MiB pinned host buffer and copies it to its GPU with one cudaMemcpyAsync per repetition (CUDA events on
the copy stream).  Phase 1: the ranks copy one after the other (everyone else idle) -> the link rate of each GPU.  Phase 2: all
ranks copy at the same time -> what the host side (memory controllers, root complexes, NUMA placement) sustains in aggregate.
Prints one JSON line on rank 0 with the NUMA node of every GPU and the CPU affinity of every rank.
"""
import json
import os

import torch
import torch.distributed as dist


def numa_node_of_gpu(index):
    try:
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev = torch.cuda.get_device_properties(index).pci_device_id
        path = '/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node' % (dom, bus, dev)
        with open(path) as f:
            return int(f.read().strip())
    except Exception:
        return None


def copy_rate(dst, src, stream, reps):
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        dst.copy_(src, non_blocking=True)
        start.record(stream)
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        stop.record(stream)
    stream.synchronize()
    return src.numel() * src.element_size() * reps / (start.elapsed_time(stop) * 1e-3) / 1e9


def main():
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    host = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    host.fill_(rank)
    device = torch.empty_like(host, device=dev)
    stream = torch.cuda.Stream(device=dev)
    alone = torch.zeros(world, dtype=torch.float64, device=dev)
    for r in range(world):                      # phase 1: one rank at a time
        if world > 1:
            dist.barrier()
        if r == rank:
            alone[r] = copy_rate(device, host, stream, 8)
    together = torch.zeros(world, dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
    together[rank] = copy_rate(device, host, stream, 16)     # phase 2: everybody at once
    d2h = torch.zeros(world, dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
    d2h[rank] = copy_rate(host, device, stream, 8)
    info = {'rank': rank, 'gpu_numa_node': numa_node_of_gpu(local), 'cpus': len(os.sched_getaffinity(0))}
    gathered = [None] * world
    if world > 1:
        for t in (alone, together, d2h):
            dist.all_reduce(t)
        dist.all_gather_object(gathered, info)
    else:
        gathered = [info]
    if rank == 0:
        print(json.dumps({'n_gpus': world, 'buffer_mib': 256,
                          'h2d_alone_gb_per_s': [round(v, 1) for v in alone.tolist()],
                          'h2d_all_ranks_at_once_gb_per_s': [round(v, 1) for v in together.tolist()],
                          'h2d_aggregate_gb_per_s': round(float(together.sum()), 1),
                          'd2h_all_ranks_at_once_gb_per_s': [round(v, 1) for v in d2h.tolist()],
                          'ranks': gathered, 'host_cpus': os.cpu_count()}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
