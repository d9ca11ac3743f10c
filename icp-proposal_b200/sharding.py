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


# ---- posterior variability maps over shards (apps/util/PosteriorVariability.scala; SURVEY 8e "posterior statistics") ----
def variability_partials(count: int, mean: np.ndarray, cov: np.ndarray):
    """Sufficient statistics of one shard's icp_posterior_variability output: [n, n * mean (N x 3), centred second moments
    (n - 1) * cov (N x 9)] flattened. Centred moments, not raw ones: raw sums of squares of ~200 mm coordinates would lose the
    sub-millimetre variances to cancellation."""
    mean = np.asarray(mean, float).reshape(-1, 3)
    cov = np.asarray(cov, float).reshape(len(mean), 9)
    if count <= 0:   # an empty shard (fewer samples than ranks) contributes nothing; its device outputs are undefined
        mean = np.zeros_like(mean)
    m2 = cov * (count - 1) if count > 1 else np.zeros_like(cov)
    return np.concatenate([[float(count)], (mean * count).ravel(), m2.ravel(), mean.ravel()])


def merge_variability(parts: np.ndarray, normals: np.ndarray | None = None):
    """parts: (world, 1 + 15 N) rows of variability_partials -> what one icp_posterior_variability call over all samples
    returns: dict(n, mean N x 3, cov N x 3 x 3, total_variance N[, normal_variance N along the given unit normals]).
    Pairwise merge of the centred moments (Chan et al.): M2 = sum_r M2_r + sum_r n_r (mean_r - mean)(mean_r - mean)^T."""
    parts = np.asarray(parts, float)
    nv = (parts.shape[1] - 1) // 15
    n_r = parts[:, 0]
    n = n_r.sum()
    mean = parts[:, 1:1 + 3 * nv].sum(0).reshape(nv, 3) / n
    m2 = parts[:, 1 + 3 * nv:1 + 12 * nv].sum(0).reshape(nv, 3, 3)
    mean_r = parts[:, 1 + 12 * nv:].reshape(len(parts), nv, 3)
    d = np.where((n_r > 0)[:, None, None], mean_r - mean[None], 0.0)
    m2 = m2 + np.einsum("r,rvi,rvj->vij", n_r, d, d)
    with np.errstate(divide="ignore", invalid="ignore"):
        cov = m2 * (np.float64(1.0) / np.float64(n - 1))
    out = dict(n=int(n), mean=mean, cov=cov, total_variance=np.trace(cov, axis1=1, axis2=2))
    if normals is not None:
        nn = np.asarray(normals, float).reshape(nv, 3)
        out["normal_variance"] = np.einsum("vi,vij,vj->v", nn, cov, nn)
    return out
