"""Chain sharding over GPUs: chains are independent, so rank r of W owns a contiguous block
of global chain ids; the Philox streams are keyed by the GLOBAL chain id (icp_chain_io.chain_id_offset), which
makes every chain's log independent of W. The only collective is the final gather of chain statistics
(NCCL on the GPUs; gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(n_chains: int, rank: int, world: int):
    """Contiguous block [lo, hi) of global chain ids owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_chains), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def chain_statistics(theta_final: np.ndarray, log_values: np.ndarray):
    """Per-shard sufficient statistics of the posterior sample: count, sum and sum of squares of the shape
    coefficients of the final states and the sum of the last product log-values (cf. apps/util/PosteriorVariability.scala)."""
    a = np.asarray(theta_final)[:, 10:]
    return np.concatenate([[len(a)], a.sum(0), (a * a).sum(0), [np.asarray(log_values)[-1, :, 0].sum()]])


def combine_statistics(stats: np.ndarray):
    """stats: (world, 2K + 2) gathered rows -> dict(n, mean, var, mean_logp)."""
    s = np.asarray(stats).sum(0)
    n = s[0]
    k = (len(s) - 2) // 2
    mean = s[1:1 + k] / n
    return dict(n=int(n), mean=mean, var=s[1 + k:1 + 2 * k] / n - mean ** 2, mean_logp=s[-1] / n)


def gather_statistics(local: np.ndarray, group=None, device=None):
    """all_gather of the per-shard statistics through torch.distributed (backend of the default group)."""
    import torch
    import torch.distributed as dist

    t = torch.as_tensor(local, dtype=torch.float64, device=device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return torch.stack(out).cpu().numpy()
