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


class SolutionGatherer:
    """The per-step gather of bench.py without per-step allocations or Python loops (round-1 review: three fresh tensors,
    a list-form all_gather and 2N copies per step cost 0.33 ms of a 9 ms step at N = 8).  Buffers are allocated once;
    `block` is this rank's slot block -- a solver can write its solution straight into `block[slot]` -- and `gather()`
    issues ONE all_gather_into_tensor on torch's current stream (the C-ABI solve calls return synchronised, so the block is
    complete; a caller that times on another stream lets that stream wait for this one).  With one system per rank the gathered tensor is
    already in system order; otherwise a precomputed index restores it."""

    def __init__(self, nsys, world, rank, n, dtype=torch.float64, device="cpu"):
        self.nsys, self.world, self.rank, self.n = nsys, world, rank, n
        self.per = slots_per_rank(nsys, world)
        self.block = torch.zeros((self.per, n), dtype=dtype, device=device)
        self.out = torch.empty((world * self.per, n), dtype=dtype, device=device) if world > 1 else self.block
        rows = [0] * nsys
        for r in range(world):
            for slot, i in enumerate(shard(nsys, world, r)):
                rows[i] = r * self.per + slot
        self.identity = rows == list(range(nsys))
        self.index = None if self.identity else torch.tensor(rows, dtype=torch.long, device=device)

    def gather(self):
        """all ranks' blocks -> (nsys, n) in system order (a view of the preallocated buffer when possible)"""
        if self.world > 1:
            dist.all_gather_into_tensor(self.out, self.block)
        if self.identity:
            return self.out[: self.nsys]
        return self.out.index_select(0, self.index)


def max_over_ranks(value, world, device="cpu"):
    """timing rule of bench.py: a multi-rank number is the MAX over ranks"""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
