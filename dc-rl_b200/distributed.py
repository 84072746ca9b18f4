"""Multi-GPU: env instances are independent, so N envs shard across ranks as contiguous blocks with no data-path
collective; the only exchange is the episode-metric vector of the logger.  One process per GPU,
`torch.distributed` (nccl on GPUs, gloo in CPU tests) for the plumbing."""
import numpy as np

from ._lib import N_METRICS


def shard_range(n_total, rank, world):
    """Contiguous block [lo, hi) of env ids owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def make_sharded_env(env_args, n_total, seed=0, rank=None, world=None, device=None, lib=None):
    """This rank's slice of an n_total-env job.  Env ids (hence months and seeds, harl/utils/envs_tools.py:56-67)
    are global, so the union over ranks is identical to one n_total-env CudaShareVecEnv."""
    import torch.distributed as dist
    from .vec_env import CudaShareVecEnv
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(n_total, rank, world)
    return CudaShareVecEnv(env_args, hi - lo, seed=seed, device=rank if device is None else device, lib=lib, first_env_id=lo)


def gather_metrics(local_metrics, device=None):
    """All-gathers the [N_METRICS] float64 metric vector of every rank -> [world, N_METRICS] numpy array.
    This is the single collective of the path."""
    import torch
    import torch.distributed as dist
    m = np.asarray(local_metrics, np.float64).reshape(N_METRICS)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return m[None].copy()
    t = torch.from_numpy(m.copy())
    if device is not None:
        t = t.to(device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()
