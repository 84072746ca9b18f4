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


def reduce_hvac_histogram(counts, device=None):
    """Sums the ranks' HVAC power histograms (Engine.hvac_histogram): one all-reduce of 32 KB instead of gathering every
    sample for the logger's 90th percentile (harl/envs/sustaindc/sustaindc_logger.py:98-99,152-155)."""
    import torch
    import torch.distributed as dist
    c = np.asarray(counts, np.int64)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return c.astype(np.uint64)
    t = torch.from_numpy(c.copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().astype(np.uint64)


def histogram_percentile(counts, range_kw, q):
    """q-th percentile (0..100) of the histogrammed samples: linear interpolation inside the bin that holds the target
    rank (bin width = range_kw / len(counts): 0.02 % of the range with 4096 bins)."""
    c = np.asarray(counts, np.float64)
    total = c.sum()
    if total == 0:
        return 0.0
    target = q / 100.0 * total
    cum = np.cumsum(c)
    b = int(np.searchsorted(cum, target, side="left"))
    b = min(b, len(c) - 1)
    below = cum[b] - c[b]
    frac = (target - below) / c[b] if c[b] > 0 else 0.0
    return (b + frac) * range_kw / len(c)
