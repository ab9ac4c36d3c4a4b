// bvg_kern_util.cu -- launch planning, posterior-moment reduction, FP64 peak probe.
#include "bvg_kernels.cuh"

namespace bvg {

static int pick_tile(int n_work) {
  if (n_work >= 16) return 32;
  if (n_work >= 8) return 16;
  if (n_work >= 4) return 8;
  return 4;
}

static cudaError_t finish_plan(LaunchPlan* plan, int per_chain_doubles, long long n_chains, int tpc, int shared_doubles = 0) {
  if (tpc <= 32) {
    if (tpc != 4 && tpc != 8 && tpc != 16 && tpc != 32) return cudaErrorInvalidValue;
    plan->threads_per_chain = tpc;
    plan->smem_per_chain = per_chain_doubles;
    plan->block_threads = 128;
    plan->chains_per_block = 128 / tpc;
    while (plan->chains_per_block > 1 && ((size_t)plan->chains_per_block * per_chain_doubles + shared_doubles) * 8 > 200 * 1024) {
      plan->chains_per_block >>= 1;
      plan->block_threads >>= 1;
    }
    plan->smem_bytes = ((size_t)plan->chains_per_block * per_chain_doubles + shared_doubles) * 8;
  } else {
    if (tpc % 32 != 0 || tpc > 1024) return cudaErrorInvalidValue;
    plan->threads_per_chain = tpc;
    plan->smem_per_chain = per_chain_doubles;
    plan->block_threads = tpc;
    plan->chains_per_block = 1;
    plan->smem_bytes = ((size_t)per_chain_doubles + BVG_BLOCK_SCRATCH + shared_doubles) * 8;
  }
  if (plan->smem_bytes > 227 * 1024) return cudaErrorInvalidConfiguration;
  plan->grid = (int)((n_chains + plan->chains_per_block - 1) / plan->chains_per_block);
  return cudaSuccess;
}

cudaError_t plan_bvar(const BvarModel& m, long long n_chains, int requested_tpc, LaunchPlan* plan) {
  plan->kron = 0;
  if (m.kron_Linv != nullptr) return plan_bvar_kron(m, n_chains, requested_tpc, plan);
  int per = bvar_smem_len(m.k, m.T, m.M, m.n, m.sigma_type, m.resid_mode, false, m.n_psi);
  int tpc = requested_tpc;
  if (tpc <= 0) {
    if ((size_t)per * 8 > 24 * 1024 || m.n > 64) tpc = (m.n > 96) ? 256 : 128;   // one CTA per chain (n = 150: two CTAs per SM)
    else tpc = pick_tile(m.n > m.k ? m.n : m.k);
  }
  if (tpc > 32) per = bvar_smem_len(m.k, m.T, m.M, m.n, m.sigma_type, m.resid_mode, true, m.n_psi);   // tile-packed factor
  return finish_plan(plan, per, n_chains, tpc, bvar_consts_len(m.k, m.M, tpc > 32));
}

cudaError_t plan_tvp(const TvpModel& m, long long n_chains, int requested_tpc, LaunchPlan* plan) {
  plan->kron = 0;
  int per = tvp_smem_len(m.b.k, m.b.T, m.b.n, m.b.sigma_type, TVP_GENERIC, m.b.n_psi);
  int tpc = requested_tpc;
  if (tpc <= 0) {
    if (tvp_eq8_ok(m.b.k, m.b.M, m.b.n_psi, m.b.diag_sigma)) tpc = 32;      // equation-decoupled path: one warp per chain
    else if ((size_t)per * 8 > 48 * 1024 || m.b.n > 64) tpc = (m.b.n > 96) ? 256 : 128;
    else tpc = pick_tile(m.b.n);
  }
  const int mode = tvp_mode(tpc, m.b.k, m.b.M, m.b.n, m.b.n_psi, m.b.diag_sigma);
  if (mode != TVP_GENERIC) per = tvp_smem_len(m.b.k, m.b.T, m.b.n, m.b.sigma_type, mode, m.b.n_psi);
  cudaError_t e = finish_plan(plan, per, n_chains, tpc);
  if (e == cudaSuccess && mode != TVP_GENERIC) {
    // one chain per CTA: the chains (each a latency-bound sequential recursion) spread evenly over the SMs
    plan->block_threads = 32;
    plan->chains_per_block = 1;
    plan->smem_bytes = (size_t)per * 8;
    plan->grid = (int)n_chains;
  }
  return e;
}

typedef cudaError_t (*bvar_launcher)(BVG_BVAR_ARGS);

cudaError_t launch_bvar(const LaunchPlan& plan, const BvarModel& m, const RngParams& rp, double* state,
                        int state_len, const BvarOut& out, long long n_chains, int s0, int s1, int burnin,
                        cudaStream_t stream) {
  if (plan.kron) return launch_bvar_kron(plan, m, rp, state, state_len, out, n_chains, s0, s1, burnin, stream);
  // smallest compiled feature variant that covers the model: 0 (lean), 2 (+ SSVS/BVS), 31 (everything)
  const int need = bvar_features_needed(m.sigma_type, m.varsel, m.resid_mode, m.vi_off != nullptr);
  const int fi = (need == 0) ? 0 : ((need & ~(int)F_VARSEL) == 0 ? 1 : 2);
  static const bvar_launcher table[5][3] = {
      {launch_bvar_g0_f0, launch_bvar_g0_f2, launch_bvar_g0_f31},
      {launch_bvar_g4_f0, launch_bvar_g4_f2, launch_bvar_g4_f31},
      {launch_bvar_g8_f0, launch_bvar_g8_f2, launch_bvar_g8_f31},
      {launch_bvar_g16_f0, launch_bvar_g16_f2, launch_bvar_g16_f31},
      {launch_bvar_g32_f0, launch_bvar_g32_f2, launch_bvar_g32_f31}};
  int gi = 0;
  switch (plan.threads_per_chain) {
    case 4: gi = 1; break;
    case 8: gi = 2; break;
    case 16: gi = 3; break;
    case 32: gi = 4; break;
    default: gi = 0;
  }
  return table[gi][fi](plan, m, rp, state, state_len, out, n_chains, s0, s1, burnin, stream);
}

cudaError_t launch_tvp(const LaunchPlan& plan, const TvpModel& m, const RngParams& rp, double* state,
                       int state_len, double* scratch, long long spc, const TvpOut& out, long long n_chains,
                       int s0, int s1, int burnin, cudaStream_t stream) {
  switch (plan.threads_per_chain) {
    case 4: return launch_tvp_g4(plan, m, rp, state, state_len, scratch, spc, out, n_chains, s0, s1, burnin, stream);
    case 8: return launch_tvp_g8(plan, m, rp, state, state_len, scratch, spc, out, n_chains, s0, s1, burnin, stream);
    case 16: return launch_tvp_g16(plan, m, rp, state, state_len, scratch, spc, out, n_chains, s0, s1, burnin, stream);
    case 32: return launch_tvp_g32(plan, m, rp, state, state_len, scratch, spc, out, n_chains, s0, s1, burnin, stream);
    default: return launch_tvp_g0(plan, m, rp, state, state_len, scratch, spc, out, n_chains, s0, s1, burnin, stream);
  }
}

// ---------------- posterior moments: sums[r] = sum over (chain, iter) of x, sumsq likewise ----------------
// draws is [chains*iters][rows]; HBM-bound streaming reduction (8 B read per element), grid = 8 CTAs per SM.
__global__ void __launch_bounds__(256) moments_kernel(const double* __restrict__ draws, long long n_draws,
                                                      long long rows, double* sums, double* sumsq) {
  const long long total = n_draws * rows;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // per-thread stride is a multiple of rows, so a thread always sees the same row and loads stay coalesced
  const long long step = (stride / rows) * rows;
  if (step == 0) {   // more rows than threads: one row at a time per thread
    for (long long rr = idx; rr < rows; rr += stride) {
      double s = 0.0, s2 = 0.0;
      for (long long d = 0; d < n_draws; ++d) { const double v = draws[d * rows + rr]; s += v; s2 = fma(v, v, s2); }
      sums[rr] = s;
      sumsq[rr] = s2;
    }
    return;
  }
  if (idx >= step) return;
  const long long r = idx % rows;
  double s = 0.0, s2 = 0.0;
  for (; idx < total; idx += step) {
    const double v = draws[idx];
    s += v;
    s2 = fma(v, v, s2);
  }
  atomicAdd(&sums[r], s);
  atomicAdd(&sumsq[r], s2);
}

cudaError_t launch_moments(const double* draws, long long chains, long long iters, long long rows, double* sums,
                           double* sumsq, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)rows * 8, stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(sumsq, 0, (size_t)rows * 8, stream);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  moments_kernel<<<sms * 8, 256, 0, stream>>>(draws, chains * iters, rows, sums, sumsq);
  return cudaGetLastError();
}

// ---------------- FP64 FMA peak probe (roofline denominator; MEASURED_PEAKS.json has no FP64 entry) --------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* sink, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0,
         a6 = a0 + 6.0, a7 = a0 + 7.0;
  const double b = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) sink[0] = s;
}

// FP64 tensor-core (DMMA m8n8k4) peak: 8 independent accumulator tiles per warp
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* sink, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9 + i; c[i][1] = 1.0 + i; }
  const double a = 1.0000001, b = 0.9999999;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) sink[0] = s;
}

cudaError_t launch_dmma_peak(double* tflops, cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* sink = nullptr;
  cudaError_t e = cudaMalloc(&sink, 8);
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 1 << 14, blocks = sms * 8, threads = 256;
  dmma_peak_kernel<<<blocks, threads, 0, stream>>>(sink, 256);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, stream);
    dmma_peak_kernel<<<blocks, threads, 0, stream>>>(sink, iters);
    cudaEventRecord(e1, stream);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8 * 8 * 4 * 8.0 * (double)iters * (double)blocks * (threads / 32);   // 512 flop per DMMA
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *tflops = best;
  return e;
}

cudaError_t launch_fp64_peak(double* tflops, cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* sink = nullptr;
  cudaError_t e = cudaMalloc(&sink, 8);
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 1 << 16, blocks = sms * 8, threads = 256;
  fp64_peak_kernel<<<blocks, threads, 0, stream>>>(sink, 1024);   // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, stream);
    fp64_peak_kernel<<<blocks, threads, 0, stream>>>(sink, iters);
    cudaEventRecord(e1, stream);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *tflops = best;
  return e;
}

}  // namespace bvg
