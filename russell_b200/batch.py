"""Sharding of independent linear systems over ranks: system i -> rank i mod world, one solver
handle per rank, no data-path collective; the solutions are gathered with ONE all_gather (NCCL on GPUs, gloo in
the CPU tests).  The reference has no distributed layer; its closest analogue is Radau5 driving two solver
handles from two threads (russell_ode/src/radau5.rs:270-296)."""
import torch
import torch.distributed as dist


def shard(nsys, world, rank):
    """indices of the systems owned by `rank` (round-robin, like `i mod G`)"""
    return list(range(rank, nsys, world))


def slots_per_rank(nsys, world):
    return (nsys + world - 1) // world


def gather_solutions(x_local, nsys, world, rank):
    """x_local: (len(shard), n) tensor of this rank's solutions.  Returns the (nsys, n) tensor of all solutions in
    system order on every rank.  Ranks with fewer systems pad their slot block (all_gather needs equal shapes)."""
    n = x_local.shape[1] if x_local.dim() == 2 and x_local.shape[0] > 0 else int(x_local.shape[-1])
    per = slots_per_rank(nsys, world)
    block = torch.zeros((per, n), dtype=x_local.dtype, device=x_local.device)
    k = len(shard(nsys, world, rank))
    if k:
        block[:k] = x_local.reshape(k, n)
    if world == 1:
        return block[:nsys].clone()
    parts = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(parts, block)
    out = torch.empty((nsys, n), dtype=x_local.dtype, device=x_local.device)
    for r in range(world):
        for slot, i in enumerate(shard(nsys, world, r)):
            out[i] = parts[r][slot]
    return out


def max_over_ranks(value, world, device="cpu"):
    """timing rule of bench.py: a multi-rank number is the MAX over ranks"""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
