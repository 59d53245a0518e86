"""Multi-process sharding of jobs (one process per GPU, launched by torchrun).

The path shards naturally: a unit of work is a (clip, row interval) task and tasks are
independent, so ranks split the JOBS (clips) between them and every rank's engine shards its own
tasks over its pipeline instances.  There is no data-path collective; results are small (192 B per
frame for histograms) and are gathered on rank 0 with `all_gather_object` only if the caller wants
them in one place.  This mirrors the reference master handing (job, task) pairs to workers
(master.cpp:1567-1606, worker.cpp:1876-1889) with the partition made static.
"""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_indices(n_items, rank, world, weights=None):
    """Indices of the items rank `rank` owns.  Without weights: strided (i % world == rank), which
    balances clip lengths that grow with the index.  With weights (e.g. frame counts): greedy
    longest-first onto the least loaded rank; deterministic, identical on every rank."""
    if weights is None:
        return list(range(rank, n_items, world))
    order = sorted(range(n_items), key=lambda i: (-weights[i], i))
    load = [0] * world
    owner = [0] * n_items
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += weights[i]
    return [i for i in range(n_items) if owner[i] == rank]


def gather_rows(local, n_items, group=None):
    """local: {item index: python object}.  Returns the full list on every rank (rank order
    resolved by item index); uses torch.distributed (nccl or gloo) if initialised."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [local.get(i) for i in range(n_items)]
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local, group=group)
    merged = {}
    for p in parts:
        merged.update(p)
    return [merged.get(i) for i in range(n_items)]
