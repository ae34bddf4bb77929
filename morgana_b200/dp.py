"""Data-parallel plumbing for the path: utterance sharding and the path's two exchanges.

One process per GPU (``torchrun``); ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is
plumbing.  The map-style kernels (scan, expansion, normalise, EMA) need no communication: utterances are independent
and the EMA / normaliser state is replicated.  The reductions need exactly one exchange per step -- a SUM all-reduce of
the packed ``[sum, count, loss]`` triples of the result records -- and training needs the gradient all-reduce.
Both ``sum`` and ``count`` are additive, so N-GPU metric results equal the 1-GPU ones up to fp64 rounding;
with equal per-rank batch sizes the global loss is the mean of the per-rank losses (Q6).
"""
import torch
import torch.distributed as dist

RECORD_DOUBLES = 6   # a 48-byte mg_term_result viewed as float64: [sum, count, loss, isum (int64 bits), f32 mirrors x2]


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced shard ``[begin, end)`` of ``n_items`` utterances for ``rank``."""
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def sharded_batches(n_items, batch_size, rank, world_size, drop_last=True):
    """This rank's batches as ``[(begin, end), ...]`` index ranges into the global utterance list.

    Rank r owns the contiguous shard :func:`shard_range` gives it and walks it in order.  Every rank gets the SAME number
    of steps (a collective per step must not dead-lock): with ``drop_last`` all batches hold exactly ``batch_size``
    utterances -- so the mean over a batch composes into the global mean -- and the remainder of each
    shard is dropped; otherwise the shard is cut into ``ceil(largest shard / batch_size)`` near-equal batches.
    """
    if batch_size < 1 or world_size < 1 or not 0 <= rank < world_size:
        raise ValueError('bad batch_size / rank / world_size')
    begin, end = shard_range(n_items, rank, world_size)
    smallest, largest = n_items // world_size, -(-n_items // world_size)
    if drop_last:
        steps = smallest // batch_size
        return [(begin + i * batch_size, begin + (i + 1) * batch_size) for i in range(steps)]
    steps = -(-largest // batch_size)
    out = []
    for i in range(steps):
        lo, hi = shard_range(end - begin, i, steps)
        out.append((begin + lo, begin + hi))
    return out


def pack_records(*record_blocks):
    """(n_i, 48) uint8 record blocks -> one (sum_i n_i, 3) float64 tensor of [sum, count, loss] (a copy)."""
    rows = [block.reshape(-1, 48).view(torch.float64)[:, :3] for block in record_blocks]
    return torch.cat(rows)


class PendingRecords(object):
    """An all-reduce of packed records in flight on the collective's own stream; :meth:`result` joins it."""
    def __init__(self, packed, work, world):
        self.packed, self.work, self.world = packed, work, world

    def result(self):
        """Make the current stream wait for the collective (no host sync) and return the (n, 3) float64 totals."""
        if self.work is not None:
            self.work.wait()
            self.work = None
            self.packed[:, 2] /= self.world
        return self.packed


def allreduce_records(*record_blocks, group=None, async_op=False):
    """SUM of the packed records over the ranks; ``loss`` columns come back as the MEAN over ranks (equal batch sizes).

    Returns a (n, 3) float64 tensor: global ``sum``, global ``count``, rank-mean ``loss``.  A single small collective
    per step; deterministic (fixed reduction order inside NCCL / gloo for a fixed world size).  With ``async_op`` the
    collective runs on its own stream, so the next step's kernels overlap its latency, and a :class:`PendingRecords` is
    returned instead (the records are copied before the call returns, so the caller may overwrite them).
    """
    packed = pack_records(*record_blocks)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    work = None
    if world > 1:
        work = dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if not async_op:
            packed[:, 2] /= world
    if async_op:
        return PendingRecords(packed, work, world)
    return packed


def metric_results(packed):
    """``sum / (count + 1e-8)`` per record (the reference's Mean.result, morgana/metrics.py:396-397) from global sums."""
    return packed[:, 0] / (packed[:, 1] + 1e-8)


def allreduce_gradients(parameters, group=None, bucket=None):
    """Average the gradients of ``parameters`` over the ranks with ONE flat-bucket all-reduce.

    ``bucket`` (optional) is a reusable flat tensor; it is returned so the caller can keep it across steps.
    """
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return bucket
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    total = sum(g.numel() for g in grads)
    if bucket is None or bucket.numel() != total or bucket.device != grads[0].device or bucket.dtype != grads[0].dtype:
        bucket = torch.empty(total, dtype=grads[0].dtype, device=grads[0].device)
    offset = 0
    for g in grads:
        bucket[offset:offset + g.numel()].copy_(g.reshape(-1))
        offset += g.numel()
    if world > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
        bucket /= world
    offset = 0
    for g in grads:
        g.copy_(bucket[offset:offset + g.numel()].view_as(g))
        offset += g.numel()
    return bucket
