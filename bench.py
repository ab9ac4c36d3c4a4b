#!/usr/bin/env python3
"""bench.py -- retained Gibbs draws/sec (chains x iterations, whole box) of the B200 sampler, next to the
reference-faithful CPU restatement timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload all|c2a|c2b|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch: `chains` independent chains x (iterations + burnin) sweeps of the
named model.  The headline (top-level keys of the JSON line) is BASELINE.json configs[1] part A: e1 VAR(2) normal-Wishart,
4096 chains, 5000 / 1000.  The default run (--workload all) also measures the other BASELINE configs at their full
iteration counts -- c2-B (p = 0:4 model grid + model-comparison table), c3 (SSVS, K = 6), c4 (TVP + SV, T = 200),
c5 (K = 20, n = 1620) -- and reports them under `workloads`, each with its own roofline / e2e / cpu_baseline, plus the
strong-scaling point of the north star (4096 chains IN TOTAL over the N GPUs) under `strong`.
Chains are sharded over ranks by global chain id with no data-path collective (weak scaling: `chains` PER GPU); the
only collective is the NCCL all-reduce of on-device posterior moments / log-likelihood sums at the end of each step.
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

METRIC = "retained_gibbs_draws_per_sec"
UNIT = "draws/s"
CHAINS = {"c2a": 4096, "c2b": 1024, "c3": 1024, "c4": 1024, "c5": 256}     # per GPU (c2b: per model and GPU)
SEED = 20240113


# ------------------------------------------------------------------------------------------------ flop / byte models
def flops_per_sweep(obj, resid_moment):
    """Algorithmic FP64 flops of one sweep of one chain as EXECUTED ('Roofline accounting').
    Constant-parameter path: precision build + rhs + Cholesky + two triangular solves + residual cross product
    + k x k Wishart block.  `resid_moment`: u u' from cross moments (what the kernel does when T > M)."""
    T, k = obj["data"]["Y"].shape
    M = obj["data"]["Z"].shape[1]
    n = k * M
    if obj["model"].get("tvp"):
        # equation-decoupled state draw (diagonal Sigma_t^-1): per period and equation Cholesky M^3/3, two column
        # solves M^2 per lane-column (M^3 in total), build + carry ~ 4 M^2; plus residuals, J pdf evaluations per
        # (series, t) and the tridiagonal solves of the SV step
        return T * k * (4.0 * M ** 3 / 3.0 + 4 * M * M) + 2 * k * M * T + T * k * 7 * 30 + 12 * T * k
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
    """SM clocks + throttle reasons during the timed region (B200_PROFILING.md recipe), NVML in-process every ~2 ms."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.rows, self.stop_flag = device, [], threading.Event()

    def _run_nvml(self):
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


# ------------------------------------------------------------------------------------------------------ CPU baselines
def _numpy_worker_init():
    """Worker start-up outside the timed region: imports, and one BLAS thread per chain (like one R process per core)."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    from oracle import samplers  # noqa: F401
    return 0


def _numpy_oracle_chain(args):
    obj, seed, literal = args
    from oracle import samplers
    from oracle.variates import VariateSource
    fn = samplers.bvartvpalg if obj["model"].get("tvp") else samplers.bvaralg
    fn(obj, VariateSource(rng=np.random.default_rng(seed)), literal=literal)
    return 0


def _c_port_ok(obj):
    pc = obj["priors"].get("coefficients", {})
    return not (obj["model"].get("tvp") or obj["model"].get("sv") or "bvs" in pc or "df" not in obj["priors"]["sigma"]) \
        and obj["data"]["SUR"].shape[1] <= 600


def cpu_timed(obj, variant, sweeps, steps=1, warmup=0, max_procs=0):
    """One CPU arm on all host cores, one chain per core like parallel::mclapply(mc.cores = ncores)
    (R/draw_posterior.bvarmodel.R:57-58).  Variants:
      port_dense        oracle/bvar_dense_ref.c: the reference's dense algorithm (kron(I_T, Sigma^-1), GEMM, LU solve,
                        chol) with plain loops standing in for BLAS (Wishart prior +- SSVS)
      port_structured   the same C port using the Kronecker / cross-moment identities (structure-exploiting CPU version)
      numpy_dense       oracle/samplers.py literal=True: the same dense matrices through NumPy / SciPy (OpenBLAS, LAPACK),
                        one BLAS thread per chain
      numpy_structured  oracle/samplers.py literal=False: banded / Kronecker forms (band-detecting for SV / TVP)
    max_procs: use at most that many cores (the dense (nT)^2 / (kT)^2 arms of c4 / c5 are memory-bound: all cores at once
    would take a minute for one sweep each).  Returns draws/s (all sweeps counted) and the description of the sample."""
    cores = os.cpu_count() or 1
    if max_procs:
        cores = min(cores, max_procs)
    times = []
    if variant.startswith("port_"):
        from oracle.dense_ref import bvar_dense
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            bvar_dense(obj, n_chains=cores, seed=1 + i, n_threads=cores, dense_gemm=(variant == "port_dense"),
                       iterations=sweeps, burnin=0)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        note = "oracle/bvar_dense_ref.c " + ("(dense kron(I_T,Sigma^-1) + GEMM + LU solve + chol, plain loops as reference BLAS)"
                                             if variant == "port_dense" else "(Kronecker / cross-moment identities, n x n Cholesky)")
    else:
        import copy
        from concurrent.futures import ProcessPoolExecutor
        small = copy.deepcopy(obj)
        small["model"]["iterations"], small["model"]["burnin"] = sweeps, 0
        lit = variant == "numpy_dense"
        with ProcessPoolExecutor(max_workers=cores, initializer=_numpy_worker_init) as ex:
            list(ex.map(abs, range(4 * cores)))            # spin the workers up before the clock starts
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                list(ex.map(_numpy_oracle_chain, [(small, 100 * i + c, lit) for c in range(cores)]))
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
        note = "oracle/samplers.py " + ("literal dense path (NumPy / SciPy: OpenBLAS + LAPACK, one BLAS thread per chain)"
                                        if lit else "structured path (Kronecker / banded forms, NumPy / SciPy)") + \
               ", one chain per process"
    t = float(np.mean(times))
    return {"value": cores * sweeps / t, "unit": UNIT, "cores": cores, "kind": "port", "variant": variant,
            "ms_per_step": t * 1e3, "sample": f"{cores} chains x {sweeps} sweeps per step; {note}"}


CPU_SWEEPS = {      # bounded samples: a few seconds of CPU work per variant
    "c2a": {"port_dense": 1500, "port_structured": 20000, "numpy_dense": 1500, "numpy_structured": 3000},
    "c2b": {"port_dense": 1500, "port_structured": 20000},
    "c3": {"port_dense": 8, "port_structured": 300, "numpy_dense": 3, "numpy_structured": 40},
    "c4": {"numpy_dense": 1, "numpy_structured": 4},
    "c5": {"numpy_dense": 1, "numpy_structured": 1},
}


def cpu_baselines(name, obj, args):
    """The `cpu_baseline` object of one workload: the reference-faithful dense arm first (C port where it covers the
    model, NumPy / OpenBLAS otherwise), the other variants under `variants` (dense with OpenBLAS,
    band-detecting, structure-exploiting) so that algorithmic and hardware gains can be separated."""
    table = dict(CPU_SWEEPS[name])
    if args.cpu_sweeps:
        table = {k: args.cpu_sweeps for k in table}
    primary = "port_dense" if ("port_dense" in table and _c_port_ok(obj)) else "numpy_dense"
    mp = 4 if name in ("c4", "c5") else 0
    out = cpu_timed(obj, primary, table[primary], max_procs=mp)
    variants = {}
    if not args.cpu_primary_only:
        for v, sw in table.items():
            if v == primary or (v.startswith("port_") and not _c_port_ok(obj)):
                continue
            try:
                variants[v] = cpu_timed(obj, v, sw, max_procs=mp)
            except Exception as exc:
                variants[v] = {"error": str(exc)}
    out["variants"] = variants
    return out


# ------------------------------------------------------------------------------------------------------- GPU workloads
class Ctx:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            from bvartools_b200.distributed import bind_to_gpu_numa
            bind_to_gpu_numa(self.local_rank)      # pinned output buffers on the GPU's own NUMA node
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream()
        from bvartools_b200._lib import lib
        self.L = lib()
        self.peaks = {}
        try:
            self.peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        self.hbm_peak = self.peaks.get("hbm_gbs") or 6650.0
        self.hbm_src = "MEASURED_PEAKS.json hbm_gbs" if self.peaks.get("hbm_gbs") else "B200_PROFILING.md fallback"
        self.traffic = {}
        for f in sorted((ROOT / "profiles").glob("r0*_traffic.json")):
            try:
                self.traffic.update(json.loads(f.read_text()))
            except Exception:
                pass
        self._fp64 = None

    def fence(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def fp64_peaks(self):
        if self._fp64 is None:
            from bvartools_b200._lib import check
            a, b = C.c_double(0.0), C.c_double(0.0)
            check(self.L.bvg_fp64_peak(C.byref(a), None))
            check(self.L.bvg_dmma_peak(C.byref(b), None))
            self._fp64 = (a.value, b.value)
        return self._fp64

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def d2h_ceiling(ctx, nbytes):
    """Bare device -> pinned-host copy rate of this box: one cudaMemcpyAsync per copy on the bench stream, every rank at
    the same time (max over ranks), nothing else running.  The end-to-end arm cannot beat it; `frac_of_ceiling` in the
    e2e record says how close it gets."""
    torch = ctx.torch
    size = int(min(max(nbytes, 1 << 26), 2 << 30))
    try:
        host = torch.empty(size, dtype=torch.uint8, pin_memory=True)
        dev = torch.empty(size, dtype=torch.uint8, device=ctx.dev)
    except Exception as exc:
        return {"error": str(exc)}
    dev.zero_()
    with torch.cuda.stream(ctx.stream):
        host.copy_(dev, non_blocking=True)
    ctx.fence()
    reps = 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ctx.stream)
    with torch.cuda.stream(ctx.stream):
        for _ in range(reps):
            host.copy_(dev, non_blocking=True)
    e1.record(ctx.stream)
    ctx.fence()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1)) / reps
    del host, dev
    return {"GBps_per_gpu": size / ms / 1e6, "GBps_all_gpus": ctx.world * size / ms / 1e6, "bytes_per_copy": size,
            "what": "plain cudaMemcpyAsync device -> pinned host, one copy at a time per rank, all ranks concurrently"}


def blocks_throughput(ctx, args):
    """Stand-alone building blocks through their C-ABI entry points, batched: host arrays in,
    host arrays out (the reference's calling convention), so the timed region holds H2D + kernel + D2H.  The bound of each
    call is the PCIe transfer of its arguments / results (bytes listed); `cpu_baseline` is the oracle restatement of the
    same block on a bounded sample, one core."""
    import bvartools_b200.blocks as bb
    rng = np.random.default_rng(7)
    out = {}

    def timed(fn, reps=3):          # (rank 0 only: no collective in here -- ctx.fence() would wait for the other ranks)
        fn()
        ts = []
        for _ in range(reps):
            ctx.torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts))

    def cpu(fn, n):
        if args.no_cpu:
            return None
        t0 = time.perf_counter()
        for i in range(n):
            fn(i)
        return {"problems_per_s": n / (time.perf_counter() - t0), "cores": 1, "kind": "port",
                "sample": f"{n} problems, oracle/blocks.py"}

    try:
        from oracle import blocks as ob
        k, M, T = 3, 7, 73
        n = k * M
        x = rng.standard_normal((M, T))
        y = rng.standard_normal((k, T))
        # ---- post_normal (src/post_normal.cpp:61-93): B posterior draws of the coefficients, c1 shape
        B = 65536
        w = rng.standard_normal((B, k, k + 2))
        sig = np.linalg.inv(np.einsum("bij,bkj->bik", w, w) / (k + 2))
        eps = rng.standard_normal((B, n))
        a0, v0 = np.zeros(n), np.eye(n) * 0.1
        t = timed(lambda: bb.post_normal(y, x, sig, a0, v0, eps=eps, method="chol"))
        moved = 8.0 * B * (k * k + n + n)
        out["post_normal"] = {"what": f"bvg_post_normal, k={k} M={M} T={T}, batch {B}", "problems_per_s": B / t, "ms": t * 1e3,
                              "pcie_bytes": moved, "pcie_GBps": moved / t / 1e9,
                              "cpu_baseline": cpu(lambda i: ob.post_normal(y, x, sig[i], a0[:, None], v0, eps[i]), 300)}
        # ---- bvs (src/bvs.cpp:73-159)
        B = 16384
        z = np.kron(x.T, np.eye(k))
        a = rng.standard_normal((B, n)) * 0.3
        lam = (rng.random((B, n)) < 0.6).astype(np.float64)
        prior = np.full(n, 0.5)
        u1, u2 = rng.random((B, n)), rng.random((B, n))
        t = timed(lambda: bb.bvs(y, z, a, lam, sig[:B], prior, u1=u1, u2=u2))
        moved = 8.0 * B * (n + n + k * k + 2 * n + n)
        out["bvs"] = {"what": f"bvg_bvs, k={k} T={T} n={n}, batch {B}", "problems_per_s": B / t, "ms": t * 1e3,
                      "pcie_bytes": moved, "pcie_GBps": moved / t / 1e9,
                      "cpu_baseline": cpu(lambda i: ob.bvs(y, z, a[i], np.diag(lam[i]), sig[i], prior, u1[i], u2[i]), 40)}
        # ---- stochvol (src/stochvol_ksc1998.cpp:60-108), c4 shape
        B, T2 = 8192, 200
        ysv = rng.standard_normal((B, T2, k)) * 0.5
        h0 = np.log(np.broadcast_to(ysv.var(axis=1, keepdims=True), ysv.shape)).copy()
        u, e = rng.random((B, k, T2)), rng.standard_normal((B, k, T2))
        sg_h, cst = np.full(k, 0.05), np.full(k, 1e-4)
        t = timed(lambda: bb.stochvol_ksc1998(ysv, h0, sg_h, h0[:, 0, :], cst, u=u, eps=e))
        moved = 8.0 * B * (5 * k * T2)
        out["stochvol"] = {"what": f"bvg_stochvol (KSC 1998), T={T2} k={k}, batch {B}", "problems_per_s": B / t, "ms": t * 1e3,
                           "pcie_bytes": moved, "pcie_GBps": moved / t / 1e9,
                           "cpu_baseline": cpu(lambda i: ob.stochvol_mixture(ysv[i], h0[i], sg_h, h0[i, 0], cst, u[i], e[i]), 20)}
        # ---- kalman_dk (src/kalman_dk.cpp:91-184): TVP state path of a k = 3, p = 2 VAR, T = 200
        B = 2048
        xk = rng.standard_normal((M, T2))
        zk = np.vstack([np.kron(xk[:, tt][None, :], np.eye(k)) for tt in range(T2)])
        yk = rng.standard_normal((k, T2))
        su = np.linalg.inv(sig[:B])
        sv = np.broadcast_to(np.eye(n) * 0.01, (B, n, n)).copy()
        Bm = np.eye(n)
        ai, Pi = np.zeros((B, n)), np.broadcast_to(np.eye(n), (B, n, n)).copy()
        ek = rng.standard_normal((B, n + T2 * (k + n)))
        t = timed(lambda: bb.kalman_dk(yk, zk, su, sv, Bm, ai, Pi, eps=ek), reps=2)
        moved = 8.0 * B * (k * k + 2 * n * n + n + ek.shape[1] + n * (T2 + 1))
        sm = lambda i: ob.kalman_dk(yk, zk, su[i], sv[i], Bm, ai[i], Pi[i], ek[i, :n], ek[i, n:].reshape(T2, k + n)[:, :k],
                                    ek[i, n:].reshape(T2, k + n)[:, k:])
        out["kalman_dk"] = {"what": f"bvg_kalman_dk, k={k} n={n} T={T2}, batch {B}", "problems_per_s": B / t, "ms": t * 1e3,
                            "pcie_bytes": moved, "pcie_GBps": moved / t / 1e9, "cpu_baseline": cpu(sm, 30)}
    except Exception as exc:                     # auxiliary measurement: never break the bench line
        out["error"] = f"{type(exc).__name__}: {exc}"
    return out


def timed_steps(ctx, step, steps, warmup, clocks=None):
    """W untimed + K timed steps bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks."""
    torch = ctx.torch
    for _ in range(max(warmup, 1)):
        step()
    ctx.fence()
    if clocks is not None and ctx.rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.fence()
    e0.record(ctx.stream)
    for _ in range(steps):
        step()
    e1.record(ctx.stream)
    ctx.fence()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    if clocks is not None and ctx.rank == 0:
        clocks.stop_flag.set()
        clocks.join(timeout=3)
    return ms / steps


def roofline_of(ctx, name, obj, sampler_info, chains, it, bi, rows, ms_per_step):
    """Roofline record of the dominant (sweep) kernel: algorithmic bytes or flops per step / sweep-kernel time measured
    with CUDA events on the launching stream (bvg_sampler_last_kernel_ms)."""
    kernel_name, last_sweep_ms, n_launches = sampler_info
    sweep_s = last_sweep_ms * 1e-3
    peak64, dmma = ctx.fp64_peaks()
    T, k = obj["data"]["Y"].shape
    M = obj["data"]["Z"].shape[1]
    n = k * M
    out_bytes = 8.0 * chains * it * sum(rows.values())
    moment = not (obj["model"].get("tvp") or obj["model"].get("sv")) and T > M
    fl = kron_flops_per_sweep(obj) if kernel_name.startswith("kron") else flops_per_sweep(obj, moment)
    if kernel_name.startswith("large"):
        fl = survey_flops_per_sweep(obj)            # direct residuals, full n^3/3 Cholesky: 1.4246 Gflop for c5
    tflops = fl * chains * (it + bi) / sweep_s / 1e12
    tr = ctx.traffic.get(f"{name}:{chains}:{it}:{bi}:{kernel_name}") or ctx.traffic.get(f"{name}:{kernel_name}")
    traffic = traffic_note = None
    if tr and "chains" in tr:
        # captures are per launch of a stated number of sweeps: scaled to the launches of this run
        per_sweep_chain = tr["dram_bytes_per_launch"] / (tr["chains"] * tr["sweeps_per_launch"])
        traffic = per_sweep_chain * chains * (it + bi) / max(n_launches, 1)
        traffic_note = tr["note"]
    elif tr:
        traffic, traffic_note = tr["dram_bytes_per_launch"], tr.get("note")
    fp64 = {"achieved_tflops": tflops, "peak_tflops": peak64, "frac": tflops / peak64 if peak64 > 0 else None,
            "peak_source": "FP64 FMA probe kernel run live by bench.py (bvg_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
            "fp64_dmma_peak_tflops": dmma, "flops_per_sweep_executed": fl, "flops_per_sweep_survey": survey_flops_per_sweep(obj)}
    common = {"traffic": traffic, "traffic_note": traffic_note, "kernel": kernel_name,
              "sweep_kernel_ms_per_step": last_sweep_ms, "sweep_launches_per_step": n_launches,
              "sweep_share_of_step": last_sweep_ms / ms_per_step, "hbm_write_GBps": out_bytes / sweep_s / 1e9,
              "hbm_peak_GBps": ctx.hbm_peak}
    if kernel_name.startswith("kron"):
        # flat-prior Kronecker kernel: ~550 flop and 240 B of output per chain-sweep; of the two stated rooflines the
        # applicable one is HBM (algorithmic bytes = retained draws, 8 (n + k^2) per chain and iteration).
        # Neither binds: the kernel is instruction-issue bound (variate generation), see `issue`.
        gbps = out_bytes / sweep_s / 1e9
        r = {"bound": "hbm", "achieved": gbps, "peak": ctx.hbm_peak, "unit": "GB/s", "frac": gbps / ctx.hbm_peak,
             "algorithmic_bytes_per_launch": out_bytes / max(n_launches, 1), "peak_source": ctx.hbm_src, "fp64": fp64}
        if tr and tr.get("warp_inst_per_chain_sweep"):
            return dict(r, **common), tr["warp_inst_per_chain_sweep"]
        return dict(r, **common), None
    if kernel_name.startswith("tvp"):
        # TVP sweep, HBM view.  Algorithmic bytes per chain-sweep of the algorithm that runs (equation-decoupled state
        # draw): the carry stream written forward / read backward, 2 * 8 * T * k (tri(M) + M), + the retained draws.
        # `round1_definition`: the same time against the bytes of the un-decoupled block recursion the round-1 kernel
        # streamed (2 * 8 * T * tri(n) + draws), kept for continuity with BENCH_r01 / VERDICT.
        diag = bool(obj["model"].get("sv")) and "psi" not in obj["priors"]
        per = 2.0 * 8.0 * T * (k * (M * (M + 1) / 2 + M) if diag else n * (n + 1) / 2)
        stream_bytes = per * chains * (it + bi)
        gbps = (stream_bytes + out_bytes) / sweep_s / 1e9
        old = (2.0 * 8.0 * T * (n * (n + 1) / 2) * chains * (it + bi) + out_bytes) / sweep_s / 1e9
        r = {"bound": "hbm", "achieved": gbps, "peak": ctx.hbm_peak, "unit": "GB/s", "frac": gbps / ctx.hbm_peak,
             "algorithmic_bytes_per_launch": (stream_bytes + out_bytes) / max(n_launches, 1),
             "round1_definition": {"achieved": old, "frac": old / ctx.hbm_peak,
                                   "what": "2*8*T*tri(n) + draws per chain-sweep (block recursion on n x n blocks)"},
             "peak_source": ctx.hbm_src, "fp64": fp64}
        return dict(r, **common), None
    tensor = kernel_name.startswith("large") or kernel_name.startswith("bvar_g0")
    tpeak = max(peak64, dmma) if tensor else peak64
    r = {"bound": "tensor" if tensor else "fp64", "achieved": tflops, "peak": tpeak, "unit": "TFLOP/s",
         "frac": tflops / tpeak if tpeak > 0 else None,
         "peak_source": ("FP64 DMMA (mma.sync m8n8k4) probe kernel run live by bench.py (bvg_dmma_peak)"
                         if tensor and dmma >= peak64 else fp64["peak_source"]),
         "fp64_dmma_peak_tflops": dmma, "flops_per_sweep_executed": fl, "flops_per_sweep_survey": survey_flops_per_sweep(obj)}
    return dict(r, **common), None


def run_sampler_workload(ctx, args, name, steps, warmup, chains=None, headline=False, with_cpu=True, with_e2e=True,
                         iterations=None, burnin=None):
    """One single-model workload (c2a, c3, c4, c5): device-timed steps, roofline, e2e through the public API, CPU arm."""
    torch, dist = ctx.torch, ctx.dist
    from bvartools_b200 import ChainSampler, _abi, pinned_empty
    from bvartools_b200._lib import check
    from bvartools_b200.workloads import baseline_workload
    obj, label = baseline_workload(name, iterations, burnin)
    chains = chains or CHAINS[name]
    it, bi = obj["model"]["iterations"], obj["model"]["burnin"]
    sampler = ChainSampler(obj, n_chains=chains, seed=SEED, chain_offset=ctx.rank * chains, device=ctx.local_rank,
                           threads_per_chain=args.threads_per_chain if headline else 0)
    rows = sampler.spec.out_rows()
    dout = sampler.device_out()
    n_out = rows["coef"]
    sums = torch.zeros(n_out, dtype=torch.float64, device=ctx.dev)
    sumsq = torch.zeros(n_out, dtype=torch.float64, device=ctx.dev)

    def step():
        sampler.reset()
        sampler.run(ctx.stream.cuda_stream)
        check(ctx.L.bvg_moments(C.cast(dout.coef, C.c_void_p), chains, it, n_out, C.c_void_p(sums.data_ptr()),
                                C.c_void_p(sumsq.data_ptr()), C.c_void_p(ctx.stream.cuda_stream)))
        if ctx.world > 1:
            dist.all_reduce(sums)
            dist.all_reduce(sumsq)

    clocks = ClockSampler(ctx.local_rank)
    ms_per_step = timed_steps(ctx, step, steps, warmup, clocks)
    info = (sampler.kernel_name, sampler.last_kernel_ms, max(int(sampler.launch_count), 1))
    launches_per_step = int(sampler.launch_count) + 1          # + the moments kernel
    total_draws = ctx.world * chains * it
    value = total_draws / (ms_per_step * 1e-3)
    post_mean = (sums / (ctx.world * chains * it)).cpu().numpy()
    out_bytes = 8.0 * chains * it * sum(rows.values())
    extras = {}
    if headline and not args.no_loglik:
        extras = headline_extras(ctx, sampler, obj, chains, it, rows, args)

    e2e = None
    if with_e2e:
        # pinned host memory per rank is bounded (8 ranks share one host): when the retained draws of the workload exceed
        # the budget the end-to-end run uses fewer chains and says so (`chains_per_gpu`)
        budget = (32 << 30) if ctx.world == 1 else (6 << 30)
        ce = chains
        while ce > 148 and 8.0 * ce * it * sum(rows.values()) > budget:
            ce //= 2
        sampler.close()
        spec_e = _abi.spec_from_model(obj, n_chains=ce, seed=SEED, chain_offset=ctx.rank * ce, device=ctx.local_rank)
        host = _abi.OutBuffers(spec_e, alloc=pinned_empty)
        d2h = sum(a.nbytes for a in host.arrays.values()) + host.status.nbytes
        times, h2d, checksum = [], 0, 0.0
        n_warm = 1 if not headline else 2
        chains_full, chains = chains, ce
        for i in range(n_warm + steps):
            ctx.fence()
            t0 = time.perf_counter()
            s = ChainSampler(obj, n_chains=chains, seed=SEED, chain_offset=ctx.rank * chains, device=ctx.local_rank,
                             threads_per_chain=args.threads_per_chain if headline else 0)   # validates + uploads the inputs
            s.run_to_host(host, ctx.stream.cuda_stream)                         # sweeps, D2H overlapped
            checksum = float(host.arrays["coef"][0, -1, 0])                     # host read of the result
            ctx.fence()
            dt = time.perf_counter() - t0
            h2d = sum(v.nbytes for v in s.spec.keep.values())
            s.close()
            if i >= n_warm:
                times.append(dt)
        te = ctx.max_over_ranks(float(np.mean(times)))
        e2e = {"value": ctx.world * chains * it / te, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * ctx.world,
               "d2h_bytes_per_step": int(d2h) * ctx.world, "ms_per_step": te * 1e3,
               "d2h_GBps_per_gpu": d2h / te / 1e9, "chains_per_gpu": chains,
               "api": "ChainSampler(model).run_to_host(pinned out) = bvg_sampler_create + bvg_sampler_run_to_host",
               "checksum": checksum}
        chains = chains_full
        del host
        if headline:
            ceil = d2h_ceiling(ctx, d2h)
            e2e["d2h_ceiling"] = ceil
            if ceil.get("GBps_per_gpu"):
                e2e["frac_of_ceiling"] = e2e["d2h_GBps_per_gpu"] / ceil["GBps_per_gpu"]
    else:
        sampler.close()
    if ctx.rank != 0:
        return None
    roofline, winst = roofline_of(ctx, name, obj, info, chains, it, bi, rows, ms_per_step)
    clk = clocks.summary()
    if winst and clk.get("sm_mhz"):
        # issue-slot roofline: executed warp instructions per chain-sweep (ncu capture of this launch shape, profiles/)
        # x chain-sweeps / sweep-kernel time against 4 schedulers x SMs x the SM clock seen during the run
        sms = torch.cuda.get_device_properties(ctx.dev).multi_processor_count
        rate = winst * chains * (it + bi) / (info[1] * 1e-3)
        peak = 4.0 * sms * clk["sm_mhz"] * 1e6
        roofline["issue"] = {"bound": "issue", "achieved": rate / 1e12, "peak": peak / 1e12, "unit": "T warp-inst/s",
                             "frac": rate / peak, "warp_inst_per_chain_sweep": winst,
                             "what": "executed warp instructions (ncu smsp__inst_executed.sum of this launch shape) / "
                                     "(4 issue slots x SMs x SM clock under load)"}
    cpu = cpu_baselines(name, obj, args) if (with_cpu and ctx.world == 1 and not args.no_cpu) else None
    par = "1 GPU" if ctx.world == 1 else (f"chains sharded over {ctx.world} GPUs by global chain id, no data-path "
                                          "collective; NCCL all-reduce of posterior moments per step")
    rec = {"value": value, "unit": UNIT, "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
           "config": {"workload": label, "chains_per_gpu": chains, "chains_total": chains * ctx.world, "parallelism": par,
                      "l2": f"retained draws {out_bytes / 1e9:.2f} GB per step per GPU >> 126 MB L2 (no flush needed)",
                      "threads_per_chain": sampler.spec.c.threads_per_chain or "auto"},
           "clocks": clk, "e2e": e2e, "gpu_launches": launches_per_step * steps, "roofline": roofline,
           "cpu_baseline": cpu, "posterior_mean_first3": [float(v) for v in post_mean[:3]]}
    rec.update(extras)
    return rec


def headline_extras(ctx, sampler, obj, chains, it, rows, args):
    """Next-row kernels on the draws resident in HBM: model-comparison log-likelihood and
    summary.bvar statistics, timed end to end per call."""
    out = {}
    try:
        ll_ms = []
        for i in range(4):
            ctx.fence()
            t0 = time.perf_counter()
            ll_tot, n_dr = sampler.loglik_total(ctx.stream.cuda_stream)
            if i > 0:
                ll_ms.append((time.perf_counter() - t0) * 1e3)
        rd_bytes = 8.0 * chains * it * (rows["coef"] + rows["sigma"])
        ms = float(np.median(ll_ms))
        mc = {"what": "bvg_sampler_loglik_total: log-likelihood of summary.bvarlist over all resident draws "
                      "(R/summary.bvarlist.R:160-186), incl. the result D2H; bound: HBM (draws read once)", "ms": ms,
              "draws_per_s": n_dr / (ms * 1e-3), "read_GBps": rd_bytes / (ms * 1e-3) / 1e9, "LL": float(ll_tot / n_dr)}
        mc["hbm_frac"] = mc["read_GBps"] / ctx.hbm_peak
        if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu:
            from oracle import blocks as ob          # the reference's per-draw loop on a bounded sample (checker as baseline)
            rng = np.random.default_rng(0)
            Yh, Zh = obj["data"]["Y"], obj["data"]["Z"]
            nd = 2000
            ols = np.linalg.lstsq(Zh, Yh, rcond=None)[0].T
            pars = np.stack([(ols + 0.01 * rng.standard_normal(ols.shape)).reshape(-1, order="F") for _ in range(nd)])
            sg = np.stack([np.cov((Yh - Zh @ ols.T).T).reshape(-1, order="F") for _ in range(nd)])
            t0 = time.perf_counter()
            ob.summary_bvarlist_row(Yh, Zh, pars, sg)
            mc["cpu_baseline"] = {"draws_per_s": nd / (time.perf_counter() - t0), "cores": 1, "kind": "port",
                                  "sample": f"{nd} draws, oracle/blocks.py summary_bvarlist_row (per-draw loop)"}
        out["model_comparison"] = mc
    except Exception as exc:                     # never let the auxiliary measurement break the bench line
        out["model_comparison"] = {"error": str(exc)}
    try:
        sm_ms, ts_ms = [], []
        for i in range(3):
            ctx.fence()
            t0 = time.perf_counter()
            sm = sampler.summary("coef", ts_se=False)
            if i > 0:
                sm_ms.append((time.perf_counter() - t0) * 1e3)
        for i in range(3):
            ctx.fence()
            t0 = time.perf_counter()
            ts = sampler.time_series_se("coef")
            if i > 0:
                ts_ms.append((time.perf_counter() - t0) * 1e3)
        nbytes = 8.0 * chains * it * rows["coef"]
        ms = float(np.median(sm_ms))
        out["posterior_summary"] = {
            "what": "bvg_sampler_summary(coef): moments + 8-pass radix select + neighbour pass over all resident draws "
                    "(R/summary.bvar.R:121-140)", "ms": ms, "values": chains * it * rows["coef"], "passes": 11,
            "stream_GBps": 11 * nbytes / (ms * 1e-3) / 1e9, "median_first3": [float(v) for v in sm["quantiles"][1][:3]],
            "time_series_se": {"what": "bvg_sampler_tsse(coef): coda spectrum0.ar per chain and row, pooled over chains",
                               "ms": float(np.median(ts_ms)), "first3": [float(v) for v in ts[:3]],
                               "naive_se_first3": [float(v) for v in (sm["sd"] / np.sqrt(chains * it))[:3]]}}
    except Exception as exc:
        out["posterior_summary"] = {"error": str(exc)}
    return out


def run_summary_e2e(ctx, args, steps):
    """Second end-to-end record of the headline workload: host inputs -> sweeps -> posterior summary ON THE DEVICE (moments,
    exact quantiles, time-series SE, log-likelihood) -> only the tables cross PCIe (what summary() of the fitted object needs)."""
    from bvartools_b200 import ChainSampler
    from bvartools_b200.workloads import baseline_workload
    obj, _ = baseline_workload("c2a")
    chains, it = CHAINS["c2a"], obj["model"]["iterations"]
    times, d2h, chk = [], 0, 0.0
    for i in range(1 + steps):
        ctx.fence()
        t0 = time.perf_counter()
        s = ChainSampler(obj, n_chains=chains, seed=SEED, chain_offset=ctx.rank * chains, device=ctx.local_rank)
        s.run(ctx.stream.cuda_stream)
        sc = s.summary("coef", ts_se=True)
        ss = s.summary("sigma", ts_se=True)
        ll, _n = s.loglik_total(ctx.stream.cuda_stream)
        chk = float(sc["mean"][0])
        ctx.fence()
        dt = time.perf_counter() - t0
        d2h = sum(np.asarray(v).nbytes for d in (sc, ss) for v in d.values()) + 8
        h2d = sum(v.nbytes for v in s.spec.keep.values())
        s.close()
        if i >= 1:
            times.append(dt)
    te = ctx.max_over_ranks(float(np.mean(times)))
    return {"value": ctx.world * chains * it / te, "unit": UNIT, "ms_per_step": te * 1e3,
            "h2d_bytes_per_step": int(h2d) * ctx.world, "d2h_bytes_per_step": int(d2h) * ctx.world,
            "api": "ChainSampler(model).run() + summary(coef) + summary(sigma) + loglik_total: draws stay in HBM, tables to host",
            "checksum": chk}


def run_grid(ctx, args, steps, warmup, with_cpu=True, with_e2e=True):
    """Workload c2-B: the p = 0:4 model grid (vignettes/model-comparison.Rmd:44-60), `chains` chains per model and GPU,
    followed by the model-comparison table (summary.bvarlist) from the draws resident in HBM.  The log-likelihood sums
    stay on the device and are all-reduced once per step (one host read of 2 x models doubles)."""
    torch, dist = ctx.torch, ctx.dist
    from bvartools_b200 import ChainSampler, _abi, pinned_empty
    from bvartools_b200.posterior import information_criteria
    from bvartools_b200.workloads import model_grid_c2b
    it = 5000
    bi = 1000
    models = model_grid_c2b(iterations=it, burnin=bi)
    chains = CHAINS["c2b"]
    stream = ctx.stream
    samplers = [ChainSampler(m, n_chains=chains, seed=SEED + 1000 * i, chain_offset=ctx.rank * chains, device=ctx.local_rank)
                for i, m in enumerate(models)]
    rows = [s.spec.out_rows() for s in samplers]
    acc = torch.zeros(2 * len(samplers), dtype=torch.float64, device=ctx.dev)
    cnt = torch.tensor([float(chains * it)] * len(samplers), dtype=torch.float64, device=ctx.dev)
    tab_holder = {}

    def step():
        for s in samplers:           # back to back on one stream: each launch already fills the SMs evenly
            s.reset()
            s.run(stream.cuda_stream)
        for i, s in enumerate(samplers):
            s.loglik_total_device(acc.data_ptr() + 16 * i, stream.cuda_stream)
        acc[1::2] = cnt
        if ctx.world > 1:
            dist.all_reduce(acc)
        tab_holder["acc"] = acc.cpu()              # the one host read of the step

    clocks = ClockSampler(ctx.local_rank)
    ms_per_step = timed_steps(ctx, step, steps, warmup, clocks)
    a = tab_holder["acc"]
    tab = []
    for i, m in enumerate(models):
        T_, M_ = m["data"]["Z"].shape
        tab.append(information_criteria(float(a[2 * i] / a[2 * i + 1]), M_, T_))
    sweep_ms = sum(s.last_kernel_ms for s in samplers)
    launches = sum(int(s.launch_count) for s in samplers) + 2 * len(samplers)       # + total-loglik + reduce kernels
    kernels = [s.kernel_name for s in samplers]
    total_draws = ctx.world * chains * it * len(models)
    out_bytes = sum(8.0 * chains * it * sum(r.values()) for r in rows)
    e2e = None
    if with_e2e:
        hosts = [_abi.OutBuffers(s.spec, alloc=pinned_empty) for s in samplers]
        d2h = sum(sum(a_.nbytes for a_ in h.arrays.values()) + h.status.nbytes for h in hosts)
        for s in samplers:
            s.close()
        times, h2d, tab_e = [], 0, []
        for i in range(1 + steps):
            ctx.fence()
            t0 = time.perf_counter()
            tab_e = []
            for j, m in enumerate(models):
                s = ChainSampler(m, n_chains=chains, seed=SEED + 1000 * j, chain_offset=ctx.rank * chains, device=ctx.local_rank)
                s.run_to_host(hosts[j], stream.cuda_stream)
                tot, nd = s.loglik_total(stream.cuda_stream)
                tab_e.append(tot / nd)
                h2d += sum(v.nbytes for v in s.spec.keep.values()) if i == 0 else 0
                s.close()
            ctx.fence()
            if i >= 1:
                times.append(time.perf_counter() - t0)
        te = ctx.max_over_ranks(float(np.mean(times)))
        e2e = {"value": total_draws / te, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * ctx.world,
               "d2h_bytes_per_step": int(d2h) * ctx.world, "ms_per_step": te * 1e3,
               "api": "per model: ChainSampler(model).run_to_host(pinned out) + loglik_total", "checksum": float(tab_e[2])}
        del hosts
    else:
        for s in samplers:
            s.close()
    if ctx.rank != 0:
        return None
    gbps = out_bytes / (sweep_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": gbps, "peak": ctx.hbm_peak, "unit": "GB/s", "frac": gbps / ctx.hbm_peak,
                "traffic": None, "kernel": ",".join(sorted(set(kernels))), "algorithmic_bytes_per_step": out_bytes,
                "sweep_kernel_ms_per_step": sweep_ms, "sweep_share_of_step": sweep_ms / ms_per_step,
                "peak_source": ctx.hbm_src}
    cpu = cpu_baselines("c2b", models[2], args) if (with_cpu and ctx.world == 1 and not args.no_cpu) else None
    par = "1 GPU" if ctx.world == 1 else (f"every model's chains sharded over {ctx.world} GPUs by global chain id; one NCCL "
                                          "all-reduce of the log-likelihood sums of the whole grid per step")
    return {"value": total_draws / (ms_per_step * 1e-3), "unit": UNIT, "ms_per_step": ms_per_step, "steps": steps,
            "warmup": warmup,
            "config": {"workload": "c2-B: e1 VAR(p)+const grid p = 0:4 on the common sample (T = 71), v_i = 0, df = k, "
                                   "scale = 1e-4, %d iter / %d burn-in, + summary.bvarlist table" % (it, bi),
                       "chains_per_model_per_gpu": chains, "chains_total": chains * ctx.world * len(models), "parallelism": par,
                       "l2": f"retained draws {out_bytes / 1e9:.2f} GB per step per GPU >> 126 MB L2 (no flush needed)"},
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches * steps, "roofline": roofline,
            "cpu_baseline": cpu,
            "model_comparison": [{"p": i, **{k_: float(v) for k_, v in r.items()}} for i, r in enumerate(tab)]}


def reference_line(args):
    """--impl reference: the reference's own algorithm for the headline config on the host cores (the dense C port:
    oracle/bvar_dense_ref.c -- the reference itself cannot be built here), same warm-up and steps."""
    from bvartools_b200.workloads import baseline_workload
    name = "c2a" if args.workload in ("all", "c2b") else args.workload
    obj, label = baseline_workload(name, args.iterations or None, None if args.burnin < 0 else args.burnin)
    table = CPU_SWEEPS[name]
    variant = "port_dense" if ("port_dense" in table and _c_port_ok(obj)) else "numpy_dense"
    sweeps = args.cpu_sweeps or table[variant]
    r = cpu_timed(obj, variant, sweeps, steps=args.steps, warmup=args.warmup)
    cores = r["cores"]
    return {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "e1 data set (shipped fixture) / synthetic", "impl": "reference",
            "config": {"workload": label, "sample": f"{cores} chains x {sweeps} sweeps per step (bounded sample of the workload)"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "variant", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", help="all (headline c2a + the other BASELINE configs) or one of c2a c2b c3 c4 c5")
    ap.add_argument("--chains", type=int, default=0, help="chains PER GPU of a single-workload run")
    ap.add_argument("--iterations", type=int, default=0)
    ap.add_argument("--burnin", type=int, default=-1)
    ap.add_argument("--threads-per-chain", type=int, default=0)
    ap.add_argument("--cpu-sweeps", type=int, default=0, help="sweeps per chain of the CPU samples")
    ap.add_argument("--cpu-primary-only", action="store_true", help="skip the extra CPU variants")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-loglik", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(reference_line(args)), flush=True)
        return 0

    ctx = Ctx(args)
    single = args.workload != "all"
    head_name = args.workload if single and args.workload != "c2b" else "c2a"
    it = args.iterations or None
    bi = None if args.burnin < 0 else args.burnin
    if single and args.workload == "c2b":
        head = run_grid(ctx, args, args.steps, args.warmup, with_e2e=not args.no_e2e)
    else:
        head = run_sampler_workload(ctx, args, head_name, args.steps, args.warmup, chains=args.chains or None,
                                    headline=True, with_e2e=not args.no_e2e, iterations=it, burnin=bi)
    workloads, strong, summary_e2e, blocks_rec = {}, None, None, None
    if not single:
        sub_steps, sub_warm = min(args.steps, 2), 1
        workloads["c2b"] = run_grid(ctx, args, sub_steps, sub_warm, with_e2e=not args.no_e2e)
        for nm in ("c3", "c4", "c5"):
            workloads[nm] = run_sampler_workload(ctx, args, nm, sub_steps, sub_warm, with_e2e=not args.no_e2e)
        # strong scaling point of the north star: 4096 chains IN TOTAL over the N GPUs (512 per GPU at N = 8)
        if ctx.world > 1:
            per = max(CHAINS["c2a"] // ctx.world, 1)
            strong = run_sampler_workload(ctx, args, "c2a", min(args.steps, 5), 2, chains=per, with_cpu=False,
                                          with_e2e=not args.no_e2e)
            if strong is not None:
                strong = {k_: strong[k_] for k_ in ("value", "unit", "ms_per_step", "config", "e2e", "roofline")}
                strong["scaling"] = "strong"
        if ctx.rank == 0 and ctx.world == 1 and not args.no_e2e:
            blocks_rec = blocks_throughput(ctx, args)
        if not args.no_e2e:
            try:
                summary_e2e = run_summary_e2e(ctx, args, min(args.steps, 3))
            except Exception as exc:
                summary_e2e = {"error": str(exc)}
    if ctx.rank == 0:
        if strong is None and not single:
            strong = {"scaling": "strong", "value": head["value"], "ms_per_step": head["ms_per_step"],
                      "note": "N = 1: the strong-scaling point (4096 chains in total) is the headline run"}
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64",
                "data": "e1 data set (shipped fixture)" if head_name == "c2a" else "synthetic"}
        for k_ in ("config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "model_comparison",
                   "posterior_summary", "posterior_mean_first3"):
            if k_ in head:
                line[k_] = head[k_]
        if not single:
            line["workloads"] = workloads
            line["strong"] = strong
            line["e2e_summary_mode"] = summary_e2e
            line["blocks"] = blocks_rec
            line["gpu_launches"] = head["gpu_launches"] + sum((w or {}).get("gpu_launches", 0) for w in workloads.values())
        print(json.dumps(line), flush=True)
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
