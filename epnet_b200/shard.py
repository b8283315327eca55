"""Data-parallel sharding of scenes across ranks (one process per GPU).  Scenes are independent units -- every op has
the batch index as a pure grid dimension (reference: blockIdx.x in FPS, sampling_gpu.cu:105) -- so the forward has no
collective; the only exchange is the timing reduction (max over ranks) and, in training, the gradient all-reduce."""
import torch
import torch.distributed as dist


def scene_ids(total_scenes: int, world: int, rank: int):
    """Contiguous block partition of scene ids 0..total-1; the first `total % world` ranks get one extra scene."""
    base, extra = divmod(total_scenes, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def max_over_ranks(values, device=None):
    """Element-wise maximum of a list of floats over all ranks (identity when not initialised)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def aggregate_scenes_per_second(scenes_this_rank: int, elapsed_ms_this_rank: float, device=None):
    """Whole-job throughput: scenes of ALL ranks / slowest rank's time."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        n = torch.tensor([float(scenes_this_rank)], dtype=torch.float64, device=device)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        total = n.item()
    else:
        total = float(scenes_this_rank)
    (ms,) = max_over_ranks([elapsed_ms_this_rank], device)
    return total / (ms / 1e3), ms
