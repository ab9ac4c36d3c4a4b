"""Sharding of independent chains / grid models over the GPUs of one box (one process per GPU, torch.distributed).

The path has no exchange step: chains never interact and the models of a `p = 0:P` grid never interact (the reference
already treats them as independent `mclapply` tasks, R/draw_posterior.bvarmodel.R:54-59).  Ranks therefore own
contiguous blocks of GLOBAL chain ids -- the Philox key contains the global id, so the draws do not depend on the
number of ranks -- and the only collectives are at the end of a run: an all-reduce of posterior-moment accumulators
(sum x, sum x^2, count) or a gather of the retained draws to rank 0.  Backend: NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np

__all__ = ["shard", "shard_weighted", "model_cost", "run_sharded", "allreduce_moments", "gather_draws", "bind_to_gpu_numa",
           "allreduce_information_criteria", "allreduce_time_series_se"]


def shard(n_items: int, rank: int, world: int):
    """Contiguous block partition: (offset, count) of `rank`; the first n_items % world ranks get one more."""
    base, extra = divmod(int(n_items), int(world))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def model_cost(obj) -> float:
    """Per-sweep cost of one chain of a model, ~ n^3/3 + 2 k M T -- the weight used to
    balance a mixed-p grid over ranks."""
    T, k = np.asarray(obj["data"]["Y"]).shape
    Z = obj["data"].get("Z")
    M = 0 if Z is None else np.asarray(Z).shape[1]
    n = k * M
    return n ** 3 / 3.0 + 2.0 * k * M * T + 50.0 * (n + k * k)


def shard_weighted(costs, world: int):
    """Contiguous partition of a work list with per-item costs into `world` blocks of about equal total cost.
    Returns a list of (offset, count), one per rank (model-major work lists stay model-major)."""
    costs = np.asarray(costs, dtype=np.float64)
    total = float(costs.sum())
    cum = np.concatenate([[0.0], np.cumsum(costs)])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        idx = int(np.searchsorted(cum, target, side="left"))
        # pick the boundary closest to the target
        if idx > 0 and abs(cum[idx - 1] - target) <= abs(cum[min(idx, len(cum) - 1)] - target):
            idx -= 1
        bounds.append(max(bounds[-1], min(idx, len(costs))))
    bounds.append(len(costs))
    return [(bounds[r], bounds[r + 1] - bounds[r]) for r in range(world)]


def bind_to_gpu_numa(device: int = 0) -> bool:
    """Pin the calling process to the CPU cores nearest to CUDA device `device` (NVML's ideal affinity), so that the
    pinned host buffers the draws are copied into are first-touched on that GPU's NUMA node.  One rank per GPU: call it
    before allocating the output buffers.  Returns False (and changes nothing) when NVML is unavailable."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device)
        bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode()))
        return True
    except Exception:
        return False


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def run_sharded(obj, n_chains_total, seed, runner=None, **kw):
    """Run this rank's block of the `n_chains_total` chains of one model.  Returns (offset, dict of
    [local chain][iter][rows] arrays).  `runner(obj, n_chains=, seed=, chain_offset=, **kw)` defaults to the CUDA
    path (`run_chains`)."""
    if runner is None:
        from .sampler import run_chains as runner
    _, rank, world = _dist()
    offset, count = shard(n_chains_total, rank, world)
    if count == 0:
        return offset, {}
    return offset, runner(obj, n_chains=count, seed=seed, chain_offset=offset, **kw)


def allreduce_moments(draws, device=None):
    """Posterior mean and variance per output row over ALL ranks' chains: all-reduce of (sum x, sum x^2, count).
    `draws`: [chain][iter][rows] array of this rank (host) or None for an empty shard."""
    import torch
    dist, _, world = _dist()
    if draws is None or draws.size == 0:
        raise ValueError("allreduce_moments needs the number of rows even for an empty shard; pass a (0, 0, rows) array")
    rows = draws.shape[-1]
    flat = np.asarray(draws, dtype=np.float64).reshape(-1, rows)
    # two all-reduces: (sums, count) first, then the sums of squares CENTRED at the global mean -- the one-pass form
    # sum x^2 - n mean^2 cancels for coefficients with |mean| >> sd (intercepts, trends), cf. DESIGN.md section 3.5b
    acc = torch.empty(rows + 1, dtype=torch.float64, device=device or "cpu")
    acc[:rows] = torch.from_numpy(flat.sum(axis=0))
    acc[rows] = float(flat.shape[0])
    if dist is not None and world > 1:
        dist.all_reduce(acc)
    acc_h = acc.cpu().numpy()
    cnt = acc_h[rows]
    mean = acc_h[:rows] / max(cnt, 1.0)
    dev = flat - mean[None, :]
    ss = torch.from_numpy((dev * dev).sum(axis=0)).to(device or "cpu")
    if dist is not None and world > 1:
        dist.all_reduce(ss)
    var = ss.cpu().numpy() / max(cnt - 1.0, 1.0)
    return mean, var, int(cnt)


def allreduce_information_criteria(ll_sum, n_draws, tot_pars, tt, device=None):
    """Model-comparison row over ALL ranks' chains: all-reduce of this rank's log-likelihood sums --
    a scalar (`ChainSampler.loglik_total`) or the T per-period sums (`ChainSampler.loglik`) -- and draw count; every
    rank gets dict(LL, AIC, BIC, HQ) as in R/summary.bvarlist.R:181-184."""
    import torch
    from .posterior import information_criteria
    dist, _, world = _dist()
    v = np.atleast_1d(np.asarray(ll_sum, dtype=np.float64))
    acc = torch.empty(v.size + 1, dtype=torch.float64, device=device or "cpu")
    acc[:v.size] = torch.from_numpy(v)
    acc[v.size] = float(n_draws)
    if dist is not None and world > 1:
        dist.all_reduce(acc)
    acc = acc.cpu().numpy()
    ll = float(acc[:v.size].sum() / acc[v.size])
    return information_criteria(ll, tot_pars, tt)


def allreduce_time_series_se(ts_local, n_chains_local, device=None):
    """Time-series SE over ALL ranks' chains from each rank's `ChainSampler.time_series_se` (coda pools chains as
    sqrt(mean_c spec_c / (iters * chains)), so ts_r^2 * chains_r^2 = sum_c spec_c / iters adds up across ranks):
    all-reduce of (ts_r^2 chains_r^2, chains_r).  Pass zeros and 0 for an empty shard."""
    import torch
    dist, _, world = _dist()
    ts = np.asarray(ts_local, dtype=np.float64).reshape(-1)
    acc = torch.empty(ts.size + 1, dtype=torch.float64, device=device or "cpu")
    acc[:ts.size] = torch.from_numpy(ts * ts * float(n_chains_local) ** 2)
    acc[ts.size] = float(n_chains_local)
    if dist is not None and world > 1:
        dist.all_reduce(acc)
    acc = acc.cpu().numpy()
    return np.sqrt(acc[:ts.size]) / acc[ts.size]


def gather_draws(local, n_chains_total, dst=0):
    """Gather [local chain][iter][rows] blocks to rank `dst` in global chain order (None on the other ranks)."""
    import torch
    dist, rank, world = _dist()
    if dist is None or world == 1:
        return np.asarray(local)
    local = np.ascontiguousarray(local, dtype=np.float64)
    shapes = [(shard(n_chains_total, r, world)[1],) + local.shape[1:] for r in range(world)]
    # NCCL moves device tensors only: stage through this rank's GPU then (gloo takes the host tensors as they are)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(local).to(dev)
    if rank == dst:
        bufs = [torch.empty(s, dtype=torch.float64, device=dev) for s in shapes]
        bufs[dst] = t
        reqs = [dist.irecv(bufs[r], src=r) for r in range(world) if r != dst and shapes[r][0] > 0]
        for q in reqs:
            q.wait()
        return np.concatenate([b.cpu().numpy() for b in bufs], axis=0)
    if local.shape[0] > 0:
        dist.send(t, dst=dst)
    return None
