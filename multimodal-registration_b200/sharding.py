"""Rank sharding of independent work units (volume pairs, batch items, BIDS subjects).

The deformation path has no exchange step: every item is processed by
exactly one rank and only scalars (timings, fold counts) are gathered at the end.  The reference
shards the same way: one process per subject (`sct_run_batch -jobs N`, README.md:131) and an
evenly split batch under MirroredStrategy (`train_synthmorph.py:193-194,284-285`).
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))


def shard_items(n_items, rank, world):
    """Indices of the items rank `rank` owns: `rank, rank + world, ...` (round robin, so a ragged
    tail is spread over the first ranks and no rank holds more than ceil(n/world) items)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError('bad rank/world: %d/%d' % (rank, world))
    return list(range(rank, n_items, world))


def per_device_batch(batch_size, nb_devices):
    """Per-device batch of the synchronous data-parallel training step; the reference asserts that
    the batch divides evenly (`train_synthmorph.py:193-195`)."""
    if batch_size % nb_devices != 0:
        raise ValueError('Batch size (%d) should be a multiple of the nr of gpus (%d)' % (batch_size, nb_devices))
    return batch_size // nb_devices


def gather_scalars(values, device=None):
    """All-gather a small list of floats from every rank -> tensor [world, len(values)]."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t[None]
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out)


def allreduce_mean_(flat_grad):
    """The one collective of the reference: average the flat gradient bucket across ranks
    (MirroredStrategy's all-reduce, `train_synthmorph.py:284-285`)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        flat_grad /= dist.get_world_size()
    return flat_grad
