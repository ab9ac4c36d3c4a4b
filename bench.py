#!/usr/bin/env python3
"""bench.py -- retained Gibbs draws/sec (chains x iterations, whole box) of the B200 sampler, next to the
reference-faithful CPU restatement timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2a|c2b|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch: `chains` independent chains x (iterations + burnin) sweeps of the
named model (default: BASELINE.json configs[1], e1 VAR(2) normal-Wishart, 4096 chains, 5000 / 1000).  Chains are
sharded over ranks by global chain id with no data-path collective (weak scaling: `chains` PER GPU); the only
collective is the NCCL all-reduce of on-device posterior moments at the end of each step.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "retained_gibbs_draws_per_sec"
UNIT = "draws/s"


def build_workload(name, iterations=None, burnin=None):
    import helpers
    if name == "c2a":
        it, bi = iterations or 5000, 1000 if burnin is None else burnin
        obj = helpers.model_c1(iterations=it, burnin=bi)
        label = "c2-A: e1 VAR(2)+const, independent normal-Wishart (v_i=0, df=1, scale=1e-4), %d iter / %d burn-in" % (it, bi)
    elif name == "c3":
        it, bi = iterations or 5000, 1000 if burnin is None else burnin
        obj = helpers.model_c3(iterations=it, burnin=bi)
        label = "c3: synthetic VAR(4) K=6 T=191 SSVS (semiautomatic 0.01/10), %d iter / %d burn-in" % (it, bi)
    elif name == "c4":
        it, bi = iterations or 500, 500 if burnin is None else burnin
        obj = helpers.model_sv(K=3, p=2, nobs=202, iterations=it, burnin=bi, tvp=True)
        label = "c4: synthetic TVP-VAR(2) K=3 T=200 + KSC-1998 SV, %d iter / %d burn-in" % (it, bi)
    elif name == "c5":
        it, bi = iterations or 200, 50 if burnin is None else burnin
        obj = helpers.model_c5(iterations=it, burnin=bi)
        label = "c5: synthetic large BVAR K=20 p=4 T=200 (n=1620), Minnesota prior, %d iter / %d burn-in" % (it, bi)
    else:
        raise SystemExit(f"unknown workload {name}")
    return obj, label


def flops_per_sweep(obj, resid_moment):
    """Algorithmic FP64 flops of one sweep of one chain as EXECUTED ('Roofline accounting').
    Constant-parameter path: precision build + rhs + Cholesky + two triangular solves + residual cross product
    + k x k Wishart block.  `resid_moment`: u u' from cross moments (what the kernel does when T > M)."""
    T, k = obj["data"]["Y"].shape
    M = obj["data"]["Z"].shape[1]
    n = k * M
    if obj["model"].get("tvp"):
        # per period: Cholesky n^3/3, triangular inverse n^3/3, M'M n^3/3, block build + mat-vecs ~ 6 n^2;
        # plus residuals, J pdf evaluations per (series, t) and the tridiagonal solves of the SV step
        return T * (n ** 3 + 6 * n * n) + 2 * k * M * T + T * k * 7 * 30 + 12 * T * k
    base = n * (n + 1) / 2 + 2 * k * k * M + n + n ** 3 / 3 + 2 * n * n + 5 * k ** 3
    if resid_moment:
        return base + 2 * k * M * M + 2 * k * k * M
    return base + 2 * k * M * T + k * (k + 1) * T


def kron_flops_per_sweep(obj):
    """Linear-algebra flops of the flat-prior Kronecker sweep as executed: E L_X^-1 (k M (M+1)),
    B^-1 applied to it (k (k+1) M), the n additions of A_ols, E E' (k (k+1) M), two k x k triple products and the
    k x k inverse-Wishart block (7 k^3).  Variate generation is NOT included."""
    T, k = obj["data"]["Y"].shape
    M = obj["data"]["Z"].shape[1]
    return k * M * (M + 1) + 2 * k * (k + 1) * M + k * M + 7 * k ** 3


def survey_flops_per_sweep(obj):
    """SURVEY.md section 8d figure (direct residuals): 8424 flop for c1."""
    T, k = obj["data"]["Y"].shape
    M = obj["data"]["Z"].shape[1]
    n = k * M
    return n * (n + 1) / 2 + 2 * k * k * M + n + n ** 3 / 3 + 2 * n * n + 2 * k * M * T + k * (k + 1) * T + 5 * k ** 3


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.rows, self.stop_flag = device, [], threading.Event()

    def _run_nvml(self):
        """NVML in-process: a sample every ~2 ms, so that even a 50 ms timed region is covered by tens of samples."""
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(self.device))
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
        while not self.stop_flag.is_set():
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            try:
                pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                pw = float("nan")
            try:
                rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            except Exception:
                rs = 0
            self.rows.append([str(self.device), str(sm), str(mx), f"{pw:.2f}", hex(rs)] +
                             [("Active" if rs & bits[n] else "Not Active")
                              for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
            self.stop_flag.wait(0.002)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass                                   # no NVML bindings: fall back to polling nvidia-smi
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.device)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        def num(v):
            try:
                f = float(v)
                return None if f != f else f
            except ValueError:
                return None
        sm = [num(r[1]) for r in self.rows if len(r) > 8 and num(r[1]) is not None]
        mx = [num(r[2]) for r in self.rows if len(r) > 8 and num(r[2]) is not None]
        pw = [num(r[3]) for r in self.rows if len(r) > 8 and num(r[3]) is not None]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def _numpy_oracle_chain(args):
    obj, seed = args
    from oracle import samplers
    from oracle.variates import VariateSource
    fn = samplers.bvartvpalg if obj["model"].get("tvp") else samplers.bvaralg
    fn(obj, VariateSource(rng=np.random.default_rng(seed)), literal=True)
    return 0


def cpu_reference_run(obj, label, steps, warmup, sample_sweeps, as_line, n_gpus=1):
    """The reference's own algorithm (dense restatement: oracle/bvar_dense_ref.c, or the NumPy oracle for SV / TVP /
    BVS) on all host cores: one chain per core like parallel::mclapply(mc.cores = ncores)."""
    cores = os.cpu_count() or 1
    is_c = not (obj["model"].get("tvp") or obj["model"].get("sv") or "bvs" in obj["priors"].get("coefficients", {}))
    if obj["data"]["SUR"].shape[1] > 600:
        is_c = False            # c5: the dense products go through LAPACK/OpenBLAS (NumPy oracle), as Armadillo would
    times = []
    if is_c:
        from oracle.dense_ref import bvar_dense
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            bvar_dense(obj, n_chains=cores, seed=1 + i, n_threads=cores, dense_gemm=True, iterations=sample_sweeps,
                       burnin=0)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
        kind_note = "oracle/bvar_dense_ref.c (dense kron(I_T,Sigma^-1) + GEMM + LU solve + chol, plain loops as reference BLAS)"
    else:
        import copy
        from concurrent.futures import ProcessPoolExecutor
        small = copy.deepcopy(obj)
        small["model"]["iterations"], small["model"]["burnin"] = sample_sweeps, 0
        with ProcessPoolExecutor(max_workers=cores) as ex:
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                list(ex.map(_numpy_oracle_chain, [(small, 100 * i + c) for c in range(cores)]))
                dt = time.perf_counter() - t0
                if i >= warmup:
                    times.append(dt)
        kind_note = "oracle/samplers.py literal dense path (NumPy/SciPy LAPACK), one chain per process"
    t = float(np.mean(times))
    value = cores * sample_sweeps / t
    base = {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} chains x {sample_sweeps} sweeps per step; {kind_note}"}
    if not as_line:
        return base
    return {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "e1 data set (shipped fixture) / synthetic", "impl": "reference", "config": {"workload": label},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def run_grid(args, rank, world, local_rank):
    """Workload c2-B: the p = 0:4 model grid (vignettes/model-comparison.Rmd:44-60), `chains` chains per model and GPU,
    followed by the model-comparison table (summary.bvarlist) from the draws resident in HBM."""
    import torch
    import torch.distributed as dist
    import helpers
    from bvartools_b200 import ChainSampler, _abi, pinned_empty
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        from bvartools_b200.distributed import bind_to_gpu_numa
        bind_to_gpu_numa(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    it = args.iterations or 5000
    bi = 1000 if args.burnin < 0 else args.burnin
    models = helpers.model_grid_c2b(iterations=it, burnin=bi)
    chains = args.chains or 1024
    stream = torch.cuda.current_stream()
    seed = 20240113
    samplers = [ChainSampler(m, n_chains=chains, seed=seed + 1000 * i, chain_offset=rank * chains, device=local_rank)
                for i, m in enumerate(models)]
    rows = [s.spec.out_rows() for s in samplers]

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def table():
        # one log-likelihood sum and draw count per model; a single all-reduce for the whole grid
        from bvartools_b200.posterior import information_criteria
        acc = torch.zeros(2 * len(samplers), dtype=torch.float64)
        for i, s in enumerate(samplers):
            tot, nd = s.loglik_total(stream.cuda_stream)
            acc[2 * i], acc[2 * i + 1] = tot, float(nd)
        if world > 1:
            acc = acc.to(dev)
            dist.all_reduce(acc)
            acc = acc.cpu()
        out = []
        for i, m in enumerate(models):
            T_, M_ = m["data"]["Z"].shape
            out.append(information_criteria(float(acc[2 * i] / acc[2 * i + 1]), M_, T_))
        return out

    def step():
        for s in samplers:           # back to back on one stream: each launch already fills the SMs evenly (running the
            s.reset()                # models on separate streams measured no faster: 35.8 vs 36.3 ms per step)
            s.run(stream.cuda_stream)
        return table()

    for _ in range(max(args.warmup, 1)):
        step()
    fence()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    e0.record(stream)
    for _ in range(args.steps):
        tab = step()
    e1.record(stream)
    fence()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    if rank == 0:
        clocks.stop_flag.set()
        clocks.join(timeout=3)
    sweep_ms = sum(s.last_kernel_ms for s in samplers)
    launches = sum(int(s.launch_count) for s in samplers) + 2 * len(samplers)       # + total-loglik + reduce kernels
    kernels = [s.kernel_name for s in samplers]
    total_draws = world * chains * it * len(models)
    out_bytes = sum(8.0 * chains * it * sum(r.values()) for r in rows)
    # ---- end to end: host inputs -> sweeps -> all draws to pinned host memory -> table
    e2e = None
    if not args.no_e2e:
        hosts = [_abi.OutBuffers(s.spec, alloc=pinned_empty) for s in samplers]
        d2h = sum(sum(a.nbytes for a in h.arrays.values()) + h.status.nbytes for h in hosts)
        for s in samplers:
            s.close()
        times, h2d = [], 0
        for i in range(1 + args.steps):
            fence()
            t0 = time.perf_counter()
            tab_e = []
            for j, m in enumerate(models):
                s = ChainSampler(m, n_chains=chains, seed=seed + 1000 * j, chain_offset=rank * chains, device=local_rank)
                s.run_to_host(hosts[j], stream.cuda_stream)
                tot, nd = s.loglik_total(stream.cuda_stream)
                tab_e.append(tot / nd)
                h2d += sum(v.nbytes for v in s.spec.keep.values()) if i == 0 else 0
                s.close()
            fence()
            if i >= 1:
                times.append(time.perf_counter() - t0)
        te = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": total_draws / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world,
               "d2h_bytes_per_step": int(d2h) * world, "ms_per_step": float(te.item()) * 1e3,
               "api": "per model: ChainSampler(model).run_to_host(pinned out) + loglik_total", "checksum": float(tab_e[2])}
    else:
        for s in samplers:
            s.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs") or 6530.6
    gbps = out_bytes / (sweep_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": gbps, "peak": hbm_peak, "unit": "GB/s", "frac": gbps / hbm_peak,
                "traffic": None, "kernel": ",".join(sorted(set(kernels))),
                "algorithmic_bytes_per_step": out_bytes, "sweep_kernel_ms_per_step": sweep_ms,
                "sweep_share_of_step": sweep_ms / ms_per_step,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "B200_PROFILING.md fallback"}
    cpu = None
    if not args.no_cpu:
        cpu = cpu_reference_run(models[2], "c2-B, p = 2 member of the grid", 1, 0, args.cpu_sweeps or 1500, False)
    par = "1 GPU" if world == 1 else f"every model's chains sharded over {world} GPUs by global chain id; NCCL all-reduce of one log-likelihood sum per model"
    line = {"metric": METRIC, "value": total_draws / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "e1 data set (shipped fixture)",
            "config": {"workload": "c2-B: e1 VAR(p)+const grid p = 0:4 on the common sample (T = 71), v_i = 0, df = k, "
                                   "scale = 1e-4, %d iter / %d burn-in, + summary.bvarlist table" % (it, bi),
                       "chains_per_model_per_gpu": chains, "chains_total": chains * world * len(models),
                       "parallelism": par,
                       "l2": f"retained draws {out_bytes / 1e9:.2f} GB per step per GPU >> 126 MB L2 (no flush needed)"},
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches * args.steps, "roofline": roofline,
            "cpu_baseline": cpu,
            "model_comparison": [{"p": i, **{k: float(v) for k, v in r.items()}} for i, r in enumerate(tab)]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2a")
    ap.add_argument("--chains", type=int, default=0, help="chains PER GPU (default: 4096 for c2a, 1024 otherwise)")
    ap.add_argument("--iterations", type=int, default=0)
    ap.add_argument("--burnin", type=int, default=-1)
    ap.add_argument("--threads-per-chain", type=int, default=0)
    ap.add_argument("--cpu-sweeps", type=int, default=0, help="sweeps per chain of the CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-loglik", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "c2b" and args.impl != "reference":
        return run_grid(args, rank, world, local_rank)
    obj, label = build_workload("c2a" if args.workload == "c2b" else args.workload, args.iterations or None,
                                None if args.burnin < 0 else args.burnin)
    cpu_sweeps = args.cpu_sweeps or {"c2a": 1500, "c3": 8, "c4": 2, "c5": 1}.get(args.workload, 50)

    if args.impl == "reference":
        if rank == 0:
            line = cpu_reference_run(obj, label, args.steps, min(args.warmup, 1), cpu_sweeps, True, n_gpus=args.gpus)
            print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    from bvartools_b200 import ChainSampler, _abi, pinned_empty
    from bvartools_b200._lib import check, lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        from bvartools_b200.distributed import bind_to_gpu_numa
        bind_to_gpu_numa(local_rank)          # pinned output buffers on the GPU's own NUMA node
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = lib()
    chains = args.chains or {"c2a": 4096, "c5": 256}.get(args.workload, 1024)
    it, bi = obj["model"]["iterations"], obj["model"]["burnin"]
    seed = 20240113
    stream = torch.cuda.current_stream()

    sampler = ChainSampler(obj, n_chains=chains, seed=seed, chain_offset=rank * chains, device=local_rank,
                           threads_per_chain=args.threads_per_chain)
    rows = sampler.spec.out_rows()
    dout = sampler.device_out()
    n_out = rows["coef"]
    sums = torch.zeros(n_out, dtype=torch.float64, device=dev)
    sumsq = torch.zeros(n_out, dtype=torch.float64, device=dev)

    def step():
        sampler.reset()
        sampler.run(stream.cuda_stream)
        check(L.bvg_moments(C.cast(dout.coef, C.c_void_p), chains, it, n_out, C.c_void_p(sums.data_ptr()),
                            C.c_void_p(sumsq.data_ptr()), C.c_void_p(stream.cuda_stream)))
        if world > 1:
            dist.all_reduce(sums)
            dist.all_reduce(sumsq)

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        step()
    fence()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sweep_ms = []
    fence()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    fence()
    elapsed_ms = e0.elapsed_time(e1)
    # duration of the sweep kernels of the last step, from the library's CUDA events on the launching stream
    last_sweep_ms = sampler.last_kernel_ms
    kernel_name = sampler.kernel_name
    launches_per_step = int(sampler.launch_count) + 1          # + the moments kernel
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    if rank == 0:
        clocks.stop_flag.set()
        clocks.join(timeout=3)
    ms_per_step = elapsed_ms / args.steps
    total_draws = world * chains * it
    value = total_draws / (ms_per_step * 1e-3)
    post_mean = (sums / (world * chains * it)).cpu().numpy()
    tpc_used = sampler.spec.c.threads_per_chain

    # ---- next-row kernel: summary.bvarlist log-likelihood over the draws resident in HBM
    model_cmp = None
    if not args.no_loglik:
        try:
            ll_ms = []
            for i in range(4):
                fence()
                t0 = time.perf_counter()
                ll_tot, n_dr = sampler.loglik_total(stream.cuda_stream)
                dt = time.perf_counter() - t0
                if i > 0:
                    ll_ms.append(dt * 1e3)
            rd_bytes = 8.0 * chains * it * (rows["coef"] + rows["sigma"])
            ms = float(np.median(ll_ms))
            model_cmp = {"what": "bvg_sampler_loglik_total: log-likelihood of summary.bvarlist over all resident draws "
                                 "(R/summary.bvarlist.R:160-186), incl. the result D2H; bound: HBM (draws read once)", "ms": ms,
                         "draws_per_s": n_dr / (ms * 1e-3), "read_GBps": rd_bytes / (ms * 1e-3) / 1e9,
                         "LL": float(ll_tot / n_dr)}
        except Exception as exc:       # never let the auxiliary measurement break the bench line
            model_cmp = {"error": str(exc)}

    # ---- next-row kernel: summary.bvar statistics (mean, SD, exact 2.5 / 50 / 97.5 % quantiles) of
    #      the coefficient draws resident in HBM
    post_summary = None
    if not args.no_loglik:
        try:
            sm_ms = []
            for i in range(3):
                fence()
                t0 = time.perf_counter()
                sm = sampler.summary("coef", ts_se=False)
                dt = time.perf_counter() - t0
                if i > 0:
                    sm_ms.append(dt * 1e3)
            ms = float(np.median(sm_ms))
            ts_ms = []
            for i in range(3):
                fence()
                t0 = time.perf_counter()
                ts = sampler.time_series_se("coef")
                dt = time.perf_counter() - t0
                if i > 0:
                    ts_ms.append(dt * 1e3)
            nbytes = 8.0 * chains * it * rows["coef"]
            post_summary = {"what": "bvg_sampler_summary(coef): moments + 8-pass radix select + neighbour pass over all "
                                    "resident draws (R/summary.bvar.R:121-140)", "ms": ms, "values": chains * it * rows["coef"],
                            "passes": 11, "stream_GBps": 11 * nbytes / (ms * 1e-3) / 1e9,
                            "median_first3": [float(v) for v in sm["quantiles"][1][:3]],
                            "time_series_se": {"what": "bvg_sampler_tsse(coef): coda spectrum0.ar per chain and row "
                                                       "(means, autocovariances to lag floor(10 log10 iter), "
                                                       "Levinson-Durbin + AIC), pooled over chains",
                                               "ms": float(np.median(ts_ms)),
                                               "first3": [float(v) for v in ts[:3]],
                                               "naive_se_first3": [float(v) for v in
                                                                   (sm["sd"] / np.sqrt(chains * it))[:3]]}}
        except Exception as exc:
            post_summary = {"error": str(exc)}

    # ---- end to end through the public API: host inputs -> H2D -> sweeps -> D2H into pinned host memory
    e2e = None
    if not args.no_e2e:
        host = _abi.OutBuffers(sampler.spec, alloc=pinned_empty)
        d2h = sum(a.nbytes for a in host.arrays.values()) + host.status.nbytes
        sampler.close()
        e2e_times, h2d, checksum = [], 0, 0.0
        n_warm = 2
        for i in range(n_warm + args.steps):
            fence()
            t0 = time.perf_counter()
            s = ChainSampler(obj, n_chains=chains, seed=seed, chain_offset=rank * chains, device=local_rank,
                             threads_per_chain=args.threads_per_chain)          # validates + uploads the inputs
            s.run_to_host(host, stream.cuda_stream)                             # sweeps, D2H overlapped
            checksum = float(host.arrays["coef"][0, -1, 0])                     # host read of the result
            fence()
            dt = time.perf_counter() - t0
            h2d = sum(v.nbytes for v in s.spec.keep.values())
            s.close()
            if i >= n_warm:
                e2e_times.append(dt)
        te = torch.tensor([float(np.mean(e2e_times))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": total_draws / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world,
               "d2h_bytes_per_step": int(d2h) * world, "ms_per_step": float(te.item()) * 1e3,
               "api": "ChainSampler(model).run_to_host(pinned out) = bvg_sampler_create + bvg_sampler_run_to_host",
               "checksum": checksum}
    else:
        sampler.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (the sweep kernel): executed algorithmic FP64 flops / launch time
    moment = not (obj["model"].get("tvp") or obj["model"].get("sv")) and \
        obj["data"]["Y"].shape[0] > obj["data"]["Z"].shape[1]
    fl_sweep = kron_flops_per_sweep(obj) if kernel_name.startswith("kron") else flops_per_sweep(obj, moment)
    if kernel_name.startswith("large"):
        fl_sweep = survey_flops_per_sweep(obj)          # direct residuals, full n^3/3 Cholesky: 1.4246 Gflop for c5
    peak = C.c_double(0.0)
    check(L.bvg_fp64_peak(C.byref(peak), None))
    dmma_peak = C.c_double(0.0)
    check(L.bvg_dmma_peak(C.byref(dmma_peak), None))
    peak_source = ("FP64 FMA probe kernel run live by bench.py (bvg_fp64_peak); MEASURED_PEAKS.json has no FP64 entry")
    if kernel_name.startswith("large") and dmma_peak.value > peak.value:
        peak = dmma_peak
        peak_source = "FP64 DMMA (mma.sync m8n8k4) probe kernel run live by bench.py (bvg_dmma_peak)"
    sweep_kernel_s = last_sweep_ms * 1e-3
    achieved = fl_sweep * chains * (it + bi) / sweep_kernel_s / 1e12
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    out_bytes = 8.0 * chains * it * sum(rows.values())
    hbm_peak = peaks.get("hbm_gbs") or 6530.6
    # ncu-measured DRAM bytes per launch of this kernel in this launch shape (profiles/r01_traffic.json, written from
    # `ncu --set full` captures of this very command); null when there is no capture of this shape
    traffic, traffic_note = None, None
    try:
        tj = json.loads((ROOT / "profiles" / "r01_traffic.json").read_text())
        key = f"{args.workload}:{chains}:{it}:{bi}:{kernel_name}"
        if key in tj:
            traffic, traffic_note = tj[key]["dram_bytes_per_launch"], tj[key]["note"]
    except Exception:
        pass
    n_sweep_launches = max(int(launches_per_step) - 1, 1)
    fp64 = {"achieved_tflops": achieved, "peak_tflops": peak.value,
            "frac": achieved / peak.value if peak.value > 0 else None, "peak_source": peak_source,
            "fp64_dmma_peak_tflops": dmma_peak.value, "flops_per_sweep_executed": fl_sweep,
            "flops_per_sweep_survey": survey_flops_per_sweep(obj)}
    if kernel_name.startswith("kron"):
        # flat-prior Kronecker kernel: ~550 flop and 240 B of output per chain-sweep; the applicable roofline of the
        # two is HBM (algorithmic bytes = retained draws only, 8 (n + k^2) per chain and iteration)
        gbps = out_bytes / sweep_kernel_s / 1e9
        roofline = {"bound": "hbm", "achieved": gbps, "peak": hbm_peak, "unit": "GB/s", "frac": gbps / hbm_peak,
                    "traffic": traffic, "traffic_note": traffic_note, "kernel": kernel_name,
                    "algorithmic_bytes_per_launch": out_bytes / n_sweep_launches,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "B200_PROFILING.md fallback",
                    "fp64": fp64}
    elif kernel_name.startswith("tvp"):
        # TVP sweep: the block-bidiagonal factor does not fit on chip and is streamed through HBM (written forward, read
        # backward); SURVEY.md 8d algorithmic bytes = 2 * 8 * T * tri(n) per chain-sweep + the retained draws
        T_, k_ = obj["data"]["Y"].shape
        n_ = k_ * obj["data"]["Z"].shape[1]
        stream_bytes = 2.0 * 8.0 * T_ * (n_ * (n_ + 1) / 2) * chains * (it + bi)
        gbps = (stream_bytes + out_bytes) / sweep_kernel_s / 1e9
        roofline = {"bound": "hbm", "achieved": gbps, "peak": hbm_peak, "unit": "GB/s", "frac": gbps / hbm_peak,
                    "traffic": traffic, "traffic_note": traffic_note, "kernel": kernel_name,
                    "algorithmic_bytes_per_launch": (stream_bytes + out_bytes) / n_sweep_launches,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "B200_PROFILING.md fallback",
                    "fp64": fp64}
    else:
        tensor = kernel_name.startswith("large") or kernel_name.startswith("bvar_g0")
        tpeak = max(peak.value, dmma_peak.value) if tensor else peak.value
        roofline = {"bound": "tensor" if tensor else "fp64", "achieved": achieved, "peak": tpeak, "unit": "TFLOP/s",
                    "frac": achieved / tpeak if tpeak > 0 else None, "traffic": traffic, "traffic_note": traffic_note,
                    "kernel": kernel_name,
                    "peak_source": ("FP64 DMMA (mma.sync m8n8k4) probe kernel run live by bench.py (bvg_dmma_peak)"
                                    if tensor and dmma_peak.value >= peak.value else peak_source),
                    "fp64_dmma_peak_tflops": dmma_peak.value, "flops_per_sweep_executed": fl_sweep,
                    "flops_per_sweep_survey": survey_flops_per_sweep(obj)}
    roofline.update({"sweep_kernel_ms_per_step": last_sweep_ms, "sweep_launches_per_step": n_sweep_launches,
                     "sweep_share_of_step": last_sweep_ms / ms_per_step,
                     "hbm_write_GBps": out_bytes / sweep_kernel_s / 1e9, "hbm_peak_GBps": hbm_peak})
    if model_cmp and "read_GBps" in model_cmp:
        model_cmp["hbm_frac"] = model_cmp["read_GBps"] / hbm_peak
        if not args.no_cpu and not obj["model"].get("tvp") and not obj["model"].get("sv"):
            # the reference's per-draw loop (oracle restatement of R/summary.bvarlist.R:160-180) on a bounded sample
            from oracle import blocks as ob
            rng = np.random.default_rng(0)
            Yh, Zh = obj["data"]["Y"], obj["data"]["Z"]
            kk_ = Yh.shape[1]
            nd = 2000
            ols = np.linalg.lstsq(Zh, Yh, rcond=None)[0].T
            pars = np.stack([(ols + 0.01 * rng.standard_normal(ols.shape)).reshape(-1, order="F") for _ in range(nd)])
            sg = np.stack([np.cov((Yh - Zh @ ols.T).T).reshape(-1, order="F") for _ in range(nd)])
            t0 = time.perf_counter()
            ob.summary_bvarlist_row(Yh, Zh, pars, sg)
            dtc = time.perf_counter() - t0
            model_cmp["cpu_baseline"] = {"draws_per_s": nd / dtc, "cores": 1, "kind": "port",
                                         "sample": f"{nd} draws, oracle/blocks.py summary_bvarlist_row (per-draw loop)"}
    cpu = None
    if not args.no_cpu:
        cpu = cpu_reference_run(obj, label, 1, 0, cpu_sweeps, False)
    par = "1 GPU" if world == 1 else (f"chains sharded over {world} GPUs by global chain id, no data-path "
                                      "collective; NCCL all-reduce of posterior moments per step")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64",
            "data": "e1 data set (shipped fixture)" if args.workload == "c2a" else "synthetic",
            "config": {"workload": label, "chains_per_gpu": chains, "chains_total": chains * world,
                       "parallelism": par,
                       "l2": f"retained draws {out_bytes / 1e9:.2f} GB per step per GPU >> 126 MB L2 (no flush needed)",
                       "threads_per_chain": tpc_used or "auto"},
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "model_comparison": model_cmp, "posterior_summary": post_summary,
            "posterior_mean_first3": [float(v) for v in post_mean[:3]]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
