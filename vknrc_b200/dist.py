"""Host-side helpers of the data-parallel path. One process per GPU: inference queries are sharded by index
range with no communication; every training batch's records are sharded per GPU and the reduced gradient is all-reduced
inside the training kernel (``NrcState.comm_connect``). ``torch.distributed`` is plumbing only: rendezvous, the exchange
of the 64-byte IPC handles, barriers. Everything here also runs on CPU with the gloo backend (tests/test_dist_cpu.py)."""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int, align: int = 1) -> tuple[int, int]:
    """[begin, end) of rank's contiguous index range of n items: ceil(n / world) items per rank, rounded up
    to a multiple of ``align`` (128 keeps every rank's shard a whole number of MMA tiles); trailing ranks may be empty."""
    if not (0 <= rank < world) or n < 0 or align < 1:
        raise ValueError("shard_range: need 0 <= rank < world, n >= 0, align >= 1")
    per = -(-n // world)
    per = -(-per // align) * align
    begin = min(n, rank * per)
    return begin, min(n, begin + per)


def exchange_handles(handle: bytes, group=None) -> list[bytes]:
    """All-gathers one fixed-size byte string per rank, returned in rank order (works with nccl and gloo groups)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [bytes(t.cpu().tolist()) for t in out]


def allreduce_gradient_reference(gradient, group=None):
    """The exchange step written with the stock collective (SUM over the 20 736-float gradient buffer: dW, loss sum,
    record count). It is the baseline the in-kernel NVLink all-reduce is checked against, and what a caller without
    peer-to-peer access would use (followed by ``NrcState.adam_step``)."""
    import torch.distributed as dist
    dist.all_reduce(gradient, op=dist.ReduceOp.SUM, group=group)
    return gradient
