"""Ray-sharded data parallelism (one process per GPU, torch.distributed over NCCL / NVLink).

The reference is single-GPU; this is the multi-GPU layer BASELINE.json asks for.  Surfels and the
acceleration structure are replicated (every rank builds the same tree from the same parameters), the ray batch is
split contiguously so that the S secondary rays of one pixel stay on one rank, outputs and per-ray gradients stay on
the owning rank, and the only collective is ONE all-reduce (sum) of the fused per-surfel gradient buffer
[N, 64] at the end of backward (GaussianTracer.flush_grads).
"""
import os

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world, align=1):
    """Contiguous [begin, end) of `n_items` for `rank`; boundaries are multiples of `align` (e.g. the S rays of one
    pixel bundle) and the sizes differ by at most one aligned unit."""
    if n_items % align:
        raise ValueError("n_items must be a multiple of align")
    units = n_items // align
    base, rem = divmod(units, world)
    b = rank * base + min(rank, rem)
    e = b + base + (1 if rank < rem else 0)
    return b * align, e * align


def shard_interleaved(n_units, rank, world, block=32):
    """Load-balanced shard: the units (pixel bundles) are cut into blocks of `block` consecutive units that are dealt
    round-robin to the ranks, so every rank receives samples from the whole image (a contiguous split of the C3 image
    left the slowest of 8 ranks with 1.35x the mean work).  Returns the sorted int64 unit indices of `rank`; the shards
    of all ranks partition range(n_units) and their sizes differ by at most one block."""
    if block < 1 or world < 1 or not 0 <= rank < world:
        raise ValueError("bad shard arguments")
    idx = torch.arange(n_units, dtype=torch.int64)
    return idx[(idx // block) % world == rank]


def init_from_env(backend=None):
    """Initialise torch.distributed from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).
    Returns (rank, local_rank, world).  A single process (no WORLD_SIZE) is rank 0 of 1 and initialises nothing."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def allreduce_sum_(t, group=None):
    """The one collective of the path: in-place sum of the fused per-surfel gradient buffer over all ranks."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def max_over_ranks(value, device):
    """Max of a python float over ranks (device-side timing is reported as the slowest rank)."""
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
