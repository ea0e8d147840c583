"""One-process-per-GPU sharding of the hot path over torch.distributed (NCCL on the GPU box,
gloo in the CPU tests).

Every stage of the path is independent per query keypoint / per feature, so the
only exchanges are the ones the north star names: broadcast the train-descriptor set and the
two images from rank 0, gather the per-shard matches and normals.  No collective sits inside a
kernel.  Shards are contiguous blocks of queries, so concatenating the gathered shards in rank
order reproduces the single-GPU output order (ascending query index).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous block [lo, hi) of `n` items owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world: int):
    return [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]


def broadcast_(tensors, src: int = 0):
    """In-place broadcast of the replicated inputs (train descriptors, images)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in tensors:
        dist.broadcast(t, src=src)


def gather_rows(local: torch.Tensor, count: int, max_count: int):
    """All-gather of a variable number of rows per rank.

    local: (>=count, ...) tensor whose first `count` rows are valid.  Every rank contributes a
    block padded to `max_count` rows; returns (rows of all ranks in rank order, counts per rank).
    """
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local[:count].clone(), [count]
    world = dist.get_world_size()
    cnt = torch.tensor([count], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    pad = torch.zeros((max_count,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:count] = local[:count]
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0), counts


def gather_matches(qidx_local: torch.Tensor, tidx: torch.Tensor, dist_: torch.Tensor, count: int,
                   query_offset: int, max_count: int):
    """Gathers NNDR matches of all shards; local query indices become global ones."""
    q, counts = gather_rows(qidx_local + query_offset, count, max_count)
    t, _ = gather_rows(tidx, count, max_count)
    d, _ = gather_rows(dist_, count, max_count)
    return q, t, d, counts


def gather_packed(parts, counts, max_count: int, unpack: bool = True):
    """ONE collective for all per-shard results of a step.

    parts:  tensors of shape (>= max_count, ...) (any dtypes); counts[i] = valid rows of parts[i].
    Every rank contributes a header with its counts plus the first `max_count` rows of every part
    as raw bytes.  With unpack=True returns (list with, per part, the valid rows of all ranks
    concatenated in rank order; per-rank counts, world x len(parts)).  With unpack=False returns the
    gathered (world x bytes) buffer and the layout needed by `unpack_packed` -- no host
    synchronisation, which is what a latency-sensitive caller wants inside its step.
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    dev = parts[0].device
    header = torch.tensor(list(counts), dtype=torch.int64, device="cpu").to(dev, non_blocking=True).view(torch.uint8)
    blobs = [header] + [p[:max_count].contiguous().view(torch.uint8).reshape(-1) for p in parts]
    layout = {"sizes": [b.numel() for b in blobs], "dtypes": [p.dtype for p in parts],
              "rows": [(max_count,) + tuple(p.shape[1:]) for p in parts], "world": world}
    buf = torch.cat(blobs)
    out = torch.empty((world, buf.numel()), dtype=torch.uint8, device=dev)
    if world == 1:
        out[0].copy_(buf)
    else:
        dist.all_gather(list(out.unbind(0)), buf)     # equal-sized blocks: one ncclAllGather on the GPU box, portable to gloo
    if not unpack:
        return out, layout
    return unpack_packed(out, layout)


def unpack_packed(out: torch.Tensor, layout):
    """Host-side view of a gather_packed(..., unpack=False) result (synchronises once for the counts)."""
    sizes, world, nparts = layout["sizes"], layout["world"], len(layout["dtypes"])
    offs = [0]
    for z in sizes:
        offs.append(offs[-1] + z)
    all_counts = out[:, :sizes[0]].clone().view(torch.int64).reshape(world, nparts).cpu().tolist()
    res = []
    for i in range(nparts):
        chunks = []
        for r in range(world):
            # clone: aligned storage for the wider view
            blk = out[r, offs[i + 1]:offs[i + 2]].clone().view(layout["dtypes"][i]).reshape(layout["rows"][i])
            chunks.append(blk[:all_counts[r][i]])
        res.append(torch.cat(chunks, dim=0))
    return res, all_counts
