"""Frame-halo exchange for sliding-window (stencil) ops when ONE clip is split into contiguous row
intervals across ranks (BASELINE configs[3]: dense OpticalFlow, stencil {0,1}).

The reference gives every task its halo by re-loading and re-DECODING the extra rows
(derive_stencil_requirements adds row+s for every stencil offset, dag_analysis.cpp:1634-1657).
Here each rank decodes only its own interval and the |stencil| boundary frames travel between
neighbouring GPUs: `torch.distributed` send/recv, i.e. NCCL over NVLink 5 / NVSwitch with the nccl
backend (a 1080p RGB frame is 6.2 MB ~ 8 us at 770 GB/s, against >= 1 ms to decode it again).
This is the path's only collective; per-frame ops need none.  Works unchanged with gloo on CPU
tensors (tests/test_multirank_cpu.py).
"""
import torch
import torch.distributed as dist


def interval_of(n_rows, rank, world):
    """Contiguous [start, end) of `n_rows` owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_rows, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def exchange_halo(block, before, after, group=None):
    """block: (n, ...) rows this rank owns (n >= max(before, after)).  Returns (prev, next): the
    `before` rows preceding and the `after` rows following the block in the whole sequence, taken
    from the neighbouring ranks; at the ends of the sequence the edge row is repeated
    (REPEAT_EDGE, reference evaluate_worker.cpp:1078-1086)."""
    if not (dist.is_available() and dist.is_initialized()):
        rank, world = 0, 1
    else:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = block.shape[0]
    assert n >= max(before, after, 1), "a rank's interval must be at least as long as the halo"
    prev = block[:1].expand(before, *block.shape[1:]).clone() if before else block[:0]
    nxt = block[-1:].expand(after, *block.shape[1:]).clone() if after else block[:0]
    if world == 1:
        return prev, nxt
    ops = []
    send_l = block[:after].contiguous() if after and rank > 0 else None        # my head is rank-1's "next"
    send_r = block[n - before:].contiguous() if before and rank < world - 1 else None  # my tail is rank+1's "prev"
    if send_l is not None:
        ops.append(dist.P2POp(dist.isend, send_l, rank - 1, group))
    if send_r is not None:
        ops.append(dist.P2POp(dist.isend, send_r, rank + 1, group))
    if after and rank < world - 1:
        ops.append(dist.P2POp(dist.irecv, nxt, rank + 1, group))
    if before and rank > 0:
        ops.append(dist.P2POp(dist.irecv, prev, rank - 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return prev, nxt


def stencil_windows(block, stencil, group=None):
    """For a sorted stencil (e.g. [0, 1] or [-1, 0, 1]) returns one tensor per offset, each (n, ...),
    such that out[k][i] is row (i + stencil[k]) of the GLOBAL sequence, edge-clamped."""
    stencil = sorted(stencil)
    before, after = max(0, -stencil[0]), max(0, stencil[-1])
    prev, nxt = exchange_halo(block, before, after, group)
    ext = torch.cat([prev, block, nxt], dim=0)
    n = block.shape[0]
    return [ext[before + s: before + s + n] for s in stencil]


def sharded_optical_flow(engine, stream_id, gpu, group=None):
    """configs[3] for one clip: every rank decodes its contiguous interval on its own GPU, swaps one
    boundary frame with its neighbour over NCCL and computes the dense flow of its rows.
    Returns (start_row, flows (n,H,W,2) float32 CUDA tensor)."""
    from . import kernels
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    total = engine.stream_rows(stream_id)
    a, b = interval_of(total, rank, world)
    frames = engine.decode_to_device(stream_id, range(a, b), gpu)
    cur, nxt = stencil_windows(frames, [0, 1], group)
    return a, kernels.optical_flow(cur, nxt.contiguous())
