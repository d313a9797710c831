"""Scene sharding across ranks (one process per GPU).  Scenes are independent units: every rank owns a
disjoint slice of the global batch and there is no data-path collective in the forward; only timings /
counters are reduced (the reference shards with DistributedSampler,
det3d/torchie/apis/train.py:297-303)."""
import torch
import torch.distributed as dist


def scenes_of_rank(global_batch, rank, world_size):
    """Indices of the global batch owned by `rank`: contiguous, sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0,{world_size})")
    base, extra = divmod(global_batch, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def scene_seeds(cfg, per_rank_batch, rank):
    """Weak-scaling synthetic workload: rank r generates scenes with seeds 1000*cfg + r*B .. + B-1."""
    return [1000 * cfg + rank * per_rank_batch + i for i in range(per_rank_batch)]


def max_over_ranks(values, device=None):
    """Element-wise MAX of a list of floats over all ranks (identity when not distributed)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def sum_over_ranks(values, device=None):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()
