"""Sharding of reads over GPUs / ranks and the host-side gather of the result rows.

The reference parallelises over reads with worker processes pulling SAM lines from a queue
(`mt_dispatcher`, scripts/STRique.py:733-830) and a collector process appending rows.  Reads are
independent, so here every rank (one process per GPU, `torch.distributed` for the plumbing) takes a
cost-balanced share of the reads, and the rows are gathered on rank 0 and written in input order.
There is no collective on the data path -- only this gather of ~100 bytes per read.
"""
import heapq
import os


def lpt_partition(costs, world):
    """Greedy longest-processing-time partition: -> list (per rank) of item indices.  Deterministic:
    ties are broken by index, every rank computes the same assignment without communicating."""
    world = max(int(world), 1)
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    shards = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + max(costs[i], 1), r))
    for s in shards:
        s.sort()
    return shards


def rank_and_world():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def init_host_group(backend='gloo'):
    """Join the torchrun rendezvous (RANK / WORLD_SIZE / MASTER_* from the environment) with a host
    backend: the gather moves Python rows, not device tensors.  No-op for a single process."""
    rank, world = rank_and_world()
    if world <= 1:
        return rank, world
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world


def broadcast_object(obj, src=0):
    """Every rank gets rank `src`'s object (the CLI's work lists: only rank 0 reads the SAM)."""
    _, world = rank_and_world()
    if world <= 1:
        return obj
    import torch.distributed as dist
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def gather_rows(local_rows, dst=0):
    """local_rows: list of (input index, row).  Rank `dst` gets all rows sorted by input index, the
    other ranks None."""
    rank, world = rank_and_world()
    if world <= 1:
        return sorted(local_rows, key=lambda r: r[0])
    import torch.distributed as dist
    gathered = [None] * world if rank == dst else None
    dist.gather_object(local_rows, gathered, dst=dst)
    if rank != dst:
        return None
    rows = [r for part in gathered for r in part]
    rows.sort(key=lambda r: r[0])
    return rows


def finalize():
    _, world = rank_and_world()
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
