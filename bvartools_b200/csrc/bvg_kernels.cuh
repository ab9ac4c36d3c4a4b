// bvg_kernels.cuh -- launcher interface between the host API (bvg_api.cu) and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include "bvg_bvar_sweep.cuh"
#include "bvg_tvp_sweep.cuh"

namespace bvg {

struct LaunchPlan {
  int threads_per_chain;   // 4..32 (tile of a warp) or a multiple of 32 (CTA per chain)
  int block_threads;
  int chains_per_block;
  size_t smem_bytes;       // dynamic shared memory per block
  int smem_per_chain;      // doubles
  int grid;
  int kron;                // 1 = flat-prior Kronecker sweep kernel (bvg_kern_kron.cu)
  int kron_nb;             // its sweeps per round
  int kron_sa;             // tuning: variate slots on the A warp of a chunk (0 = planner's choice)
};

#define BVG_BLOCK_SCRATCH 40   // doubles reserved in front of the chain's shared memory for BlockGroup

cudaError_t plan_bvar(const BvarModel& m, long long n_chains, int requested_tpc, LaunchPlan* plan);
cudaError_t launch_bvar(const LaunchPlan& plan, const BvarModel& m, const RngParams& rp, double* state,
                        int state_len, const BvarOut& out, long long n_chains, int sweep_begin, int sweep_end,
                        int burnin, cudaStream_t stream);

#define BVG_BVAR_ARGS_NAMED const LaunchPlan& plan, const BvarModel& m, const RngParams& rp, double* state, int state_len, const BvarOut& out, long long n_chains, int s0, int s1, int burnin, cudaStream_t stream
cudaError_t plan_bvar_kron(const BvarModel& m, long long n_chains, int requested_tpc, LaunchPlan* plan);
cudaError_t launch_bvar_kron(BVG_BVAR_ARGS_NAMED);

// large-VAR path (bvg_kern_large.cu): per-chain n x n matrices in HBM, blocked DMMA Cholesky
size_t large_workspace_doubles(const BvarModel& m, long long n_chains);
cudaError_t launch_bvar_large(const BvarModel& m, const RngParams& rp, double* state, int state_len, double* work,
                              const BvarOut& out, long long n_chains, int s0, int s1, int burnin, cudaStream_t stream,
                              long long* launches);

cudaError_t plan_tvp(const TvpModel& m, long long n_chains, int requested_tpc, LaunchPlan* plan);
cudaError_t launch_tvp(const LaunchPlan& plan, const TvpModel& m, const RngParams& rp, double* state,
                       int state_len, double* scratch, long long scratch_per_chain, const TvpOut& out,
                       long long n_chains, int sweep_begin, int sweep_end, int burnin, cudaStream_t stream);

// log-likelihood of resident draws reduced to per-period sums (bvg_loglik.cu)
cudaError_t launch_loglik_draws(int k, int M, int T, int tvp, int tv_sigma, long long n_draws, const double* dY,
                                const double* dX, const double* d_coef, const double* d_sigma, double* d_ll_t,
                                cudaStream_t stream);
cudaError_t launch_loglik_total(int k, int M, int T, long long n_draws, const double* dXX, const double* dAols,
                                const double* dSSE, const double* d_coef, const double* d_sigma, double* d_total,
                                cudaStream_t stream);
// posterior summaries of resident draws (bvg_summary.cu): exact type-7 quantiles by radix select, centred sum of squares
cudaError_t launch_quantiles(const double* draws, long long n_draws, long long rows, const double* d_probs, int n_probs,
                             double* d_out, cudaStream_t stream);
cudaError_t launch_tsse(const double* draws, long long n_chains, long long iters, long long rows, double* d_out,
                        cudaStream_t stream);
cudaError_t launch_centered_ss(const double* draws, long long n_draws, long long rows, const double* d_sums, double* d_ss,
                               cudaStream_t stream);
// impulse responses per draw (bvg_postest.cu)
cudaError_t launch_irf_draws(const double* coef, long long coef_stride, const double* sigma, long long sigma_stride,
                             long long n_draws, int k, int p, int h, int type, int impulse, int response, int shock_kind,
                             double shock, int cumulative, double* d_out, cudaStream_t stream, const double* a0 = nullptr,
                             long long a0_stride = 0, int a0_mode = 0);
cudaError_t launch_fevd_draws(const double* coef, long long coef_stride, const double* sigma, long long sigma_stride,
                              long long n_draws, int k, int p, int h, int type, int response, double* d_out,
                              cudaStream_t stream, const double* a0 = nullptr, long long a0_stride = 0,
                              int a0_mode = 0);
cudaError_t launch_forecast_draws(const double* coef, long long coef_stride, const double* sigma, long long sigma_stride,
                                  long long n_draws, int k, int p, int tot, int h, int use_a, const double* d_pred,
                                  const double* d_eps, unsigned long long seed, double* d_out, cudaStream_t stream,
                                  const double* a0 = nullptr, long long a0_stride = 0, int a0_mode = 0);
cudaError_t launch_moments(const double* draws, long long chains, long long iters, long long rows, double* sums,
                           double* sumsq, cudaStream_t stream);
cudaError_t launch_fp64_peak(double* tflops, cudaStream_t stream);
cudaError_t launch_dmma_peak(double* tflops, cudaStream_t stream);

// per-instantiation launchers (one translation unit each, see Makefile): G = 4, 8, 16, 32 lanes per chain, 0 = one
// CTA per chain; BVAR kernels additionally per compile-time feature mask (0 lean, 2 + variable selection, 31 all)
#define BVG_BVAR_ARGS const LaunchPlan&, const BvarModel&, const RngParams&, double*, int, const BvarOut&, long long, int, int, int, cudaStream_t
#define BVG_TVP_ARGS const LaunchPlan&, const TvpModel&, const RngParams&, double*, int, double*, long long, const TvpOut&, long long, int, int, int, cudaStream_t
#define BVG_DECL_TILE(G)                          \
  cudaError_t launch_bvar_g##G##_f0(BVG_BVAR_ARGS);  \
  cudaError_t launch_bvar_g##G##_f2(BVG_BVAR_ARGS);  \
  cudaError_t launch_bvar_g##G##_f31(BVG_BVAR_ARGS); \
  cudaError_t launch_tvp_g##G(BVG_TVP_ARGS);
BVG_DECL_TILE(0)
BVG_DECL_TILE(4)
BVG_DECL_TILE(8)
BVG_DECL_TILE(16)
BVG_DECL_TILE(32)

template <typename K>
static inline cudaError_t set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return cudaSuccess;
}

}  // namespace bvg
