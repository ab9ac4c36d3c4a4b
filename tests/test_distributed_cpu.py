"""N > 1 host logic on CPU: world_size-2 (and 3) `gloo` process groups exercise the sharding of chains over ranks,
the all-reduce of posterior moments and the gather of draws.  The per-rank compute is the TEST-ONLY host emulation
of the kernel logic (tests/emu) standing in for the CUDA path; what is checked is that the sharded job reproduces the
single-process job bit for bit (global chain ids key the Philox stream) and that the collectives are wired correctly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from bvartools_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _emu_runner(obj, n_chains, seed, chain_offset=0, **kw):
    emu = helpers.load_emu()
    rc, out = helpers.run_abi(emu.emu_bvaralg, obj, n_chains, None, seed=seed, chain_offset=chain_offset)
    assert rc == 0
    return dict(out.arrays)


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        obj = helpers.model_c1(iterations=12, burnin=3)
        off, res = D.run_sharded(obj, total, seed=77, runner=_emu_runner)
        coef = res["coef"] if res else np.zeros((0, 12, 21))
        mean, var, cnt = D.allreduce_moments(coef)
        gathered = D.gather_draws(coef, total, dst=0)
        # model-comparison row: this rank's log-likelihood sums (oracle loop standing in for the device reduction)
        from oracle import blocks as ob
        Y, Z = obj["data"]["Y"], obj["data"]["Z"]
        ll_t = np.zeros(Y.shape[0])
        sig = res["sigma"] if res else np.zeros((0, 12, 9))
        for c in range(coef.shape[0]):
            for j in range(coef.shape[1]):
                u = Y.T - coef[c, j].reshape(3, -1, order="F") @ Z.T
                ll_t += ob.loglik_normal(u, sig[c, j].reshape(3, 3, order="F"))
        ic = D.allreduce_information_criteria(ll_t, coef.shape[0] * coef.shape[1], Z.shape[1], Y.shape[0])
        # time-series SE: this rank's pooled value (oracle restatement standing in for bvg_sampler_tsse)
        ts_local = ob.time_series_se(coef.reshape(-1, 21), coef.shape[0]) if coef.shape[0] else np.zeros(21)
        ts = D.allreduce_time_series_se(ts_local, coef.shape[0])
        q.put((rank, off, coef.shape[0], mean, var, cnt, gathered, ic, ts))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 7), (3, 4)])
def test_sharded_job_equals_single_process_job(world, total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    obj = helpers.model_c1(iterations=12, burnin=3)
    full_res = _emu_runner(obj, total, 77)
    full = full_res["coef"]
    from oracle import blocks as ob
    want_ic = ob.summary_bvarlist_row(obj["data"]["Y"], obj["data"]["Z"], full.reshape(-1, 21),
                                      full_res["sigma"].reshape(-1, 9))
    got.sort(key=lambda t: t[0])
    assert sum(g[2] for g in got) == total
    assert [g[1] for g in got] == [D.shard(total, r, world)[0] for r in range(world)]
    flat = full.reshape(-1, 21)
    want_ts = ob.time_series_se(flat, total)
    for rank, off, cnt_local, mean, var, cnt, gathered, ic, ts in got:
        assert cnt == flat.shape[0]
        np.testing.assert_allclose(ts, want_ts, rtol=1e-12)
        for key in ("LL", "AIC", "BIC", "HQ"):
            assert abs(ic[key] - want_ic[key]) <= 1e-10 * abs(want_ic[key]), key
        np.testing.assert_allclose(mean, flat.mean(axis=0), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(var, flat.var(axis=0, ddof=1), rtol=1e-9)
        if rank == 0:
            assert np.array_equal(gathered, full)            # bit-identical regardless of the number of ranks
        else:
            assert gathered is None


def test_shard_partitions():
    for n, w in ((4096, 8), (5120, 8), (7, 2), (3, 8), (0, 4)):
        parts = [D.shard(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        for (o1, c1), (o2, _) in zip(parts, parts[1:]):
            assert o1 + c1 == o2
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_weighted_partition_of_the_p_grid():
    """config c2-B: p = 0:4 grid, 1024 chains per model, model-major work list balanced by n^3/3 + 2kMT."""
    x, names = helpers.e1_readme_sample()
    models = helpers.add_priors(helpers.gen_var(x, p=range(0, 5), deterministic="const", iterations=10, burnin=2,
                                                names=names), coef=dict(v_i=0, v_i_det=0), sigma=dict(df="k", scale=1e-4))
    costs = np.repeat([D.model_cost(m) for m in models], 1024)
    parts = D.shard_weighted(costs, 8)
    assert parts[0][0] == 0 and sum(c for _, c in parts) == costs.size
    for (o1, c1), (o2, _) in zip(parts, parts[1:]):
        assert o1 + c1 == o2
    loads = np.array([costs[o:o + c].sum() for o, c in parts])
    assert loads.max() / loads.mean() < 1.02
    assert parts[0][1] > parts[-1][1]                       # cheap p = 0 chains: more of them per rank
