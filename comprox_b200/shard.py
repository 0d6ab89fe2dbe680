"""Multi-GPU plumbing: independent containers ("shards") are dealt round-robin to the ranks of one node, every rank
compresses its shards with no data-path collective, and the finished containers are gathered
in shard order.  torch.distributed is used for the rendezvous and the gather only (NCCL on the GPU box, gloo in the
CPU tests); the compression itself never touches it."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shards_of(rank: int, world: int, nshards: int) -> list[int]:
    """Shard indices owned by `rank`: round-robin, so consecutive shards land on different GPUs."""
    return list(range(rank, nshards, world))


def gather_containers(mine: dict[int, bytes], nshards: int, device: str = "cpu") -> list[bytes] | None:
    """All ranks call this with {shard index: container}.  Rank 0 returns the containers in shard order, the other
    ranks return None.  Sizes travel first, then one padded uint8 tensor per rank."""
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = torch.zeros(nshards, dtype=torch.int64, device=device)
    for i, c in mine.items():
        sizes[i] = len(c)
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM)
    per_rank = [int(sum(int(sizes[i]) for i in shards_of(r, world, nshards))) for r in range(world)]
    cap = max(per_rank) if per_rank else 0
    buf = torch.zeros(max(cap, 1), dtype=torch.uint8, device=device)
    pos = 0
    for i in shards_of(rank, world, nshards):
        c = mine[i]
        buf[pos:pos + len(c)] = torch.frombuffer(bytearray(c), dtype=torch.uint8).to(device)
        pos += len(c)
    gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0)
    if rank != 0:
        return None
    out: list[bytes] = [b""] * nshards
    for r in range(world):
        flat = gathered[r].cpu().numpy().tobytes()
        pos = 0
        for i in shards_of(r, world, nshards):
            n = int(sizes[i])
            out[i] = flat[pos:pos + n]
            pos += n
    return out


def max_over_ranks(value: float, device: str = "cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
