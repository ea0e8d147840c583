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


# ------------------------------------------------------------------ one broadcast, one gather per step
def _align(n, a=256):
    return (n + a - 1) // a * a


class ReplicatedBuffer:
    """The inputs every rank needs in full (train descriptors, train keypoints, both frames) as views into ONE device
    allocation, so that a step replicates them with ONE broadcast instead of one per tensor."""

    def __init__(self, specs, device):
        """specs: list of (name, shape, torch dtype)."""
        self.views, off = {}, 0
        layout = []
        for name, shape, dtype in specs:
            nbytes = int(torch.tensor([], dtype=dtype).element_size())
            for s in shape:
                nbytes *= int(s)
            layout.append((name, shape, dtype, off, nbytes))
            off = _align(off + nbytes)
        self.buf = torch.zeros(max(off, 1), dtype=torch.uint8, device=device)
        for name, shape, dtype, o, nbytes in layout:
            self.views[name] = self.buf[o:o + nbytes].view(dtype).reshape(shape)

    def __getitem__(self, name):
        return self.views[name]

    def broadcast_(self, src: int = 0):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(self.buf, src=src)


class ShardGather:
    """All per-shard results of a step in ONE all-gather with nothing on the host: every rank copies its header (the two
    device-resident counts) and the first `cap` rows of every part into a preallocated send block with device-side copies,
    then `all_gather_into_tensor` fills a preallocated (world, block) buffer.  `unpack` (outside any timed region) gives the
    valid rows of all ranks in rank order = ascending global query index."""

    HEADER = 256

    def __init__(self, parts, cap, device, rank=0):
        """parts: list of (name, row shape tuple, torch dtype, index of the count it is valid up to)."""
        self.parts, self.cap = parts, int(cap)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        off = self.HEADER
        self.layout = []
        for name, row, dtype, which in parts:
            nbytes = int(torch.tensor([], dtype=dtype).element_size()) * self.cap
            for s in row:
                nbytes *= int(s)
            self.layout.append((name, row, dtype, which, off, nbytes))
            off = _align(off + nbytes)
        self.block = off
        self.send = torch.zeros(self.block, dtype=torch.uint8, device=device)
        self.recv = torch.zeros((self.world, self.block), dtype=torch.uint8, device=device)
        self._hdr = self.send[:16].view(torch.int32)        # 4 x int32: counts[0], counts[1], rank, cap
        self._hdr[2:4] = torch.tensor([rank, self.cap], dtype=torch.int32)      # once, at construction

    def gather(self, tensors, counts_dev):
        """tensors: dict name -> device tensor with >= cap rows; counts_dev: two 1-element int32 device tensors."""
        self._hdr[0:1].copy_(counts_dev[0])
        self._hdr[1:2].copy_(counts_dev[1])
        for name, row, dtype, which, off, nbytes in self.layout:
            self.send[off:off + nbytes].view(dtype).reshape((self.cap,) + tuple(row)).copy_(tensors[name][:self.cap])
        if self.world == 1:
            self.recv[0].copy_(self.send)
        else:
            dist.all_gather_into_tensor(self.recv.view(-1), self.send)
        return self.recv

    def unpack(self):
        hdr = self.recv[:, :16].contiguous().view(torch.int32).reshape(self.world, 4).cpu().numpy()
        out = {}
        for name, row, dtype, which, off, nbytes in self.layout:
            chunks = []
            for r in range(self.world):
                blk = self.recv[r, off:off + nbytes].clone().view(dtype).reshape((self.cap,) + tuple(row))
                chunks.append(blk[:int(hdr[r, which])])
            out[name] = torch.cat(chunks, 0)
        return out, hdr[:, :2].tolist()
