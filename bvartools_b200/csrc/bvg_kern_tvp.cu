// bvg_kern_tvp.cu -- sm_100a kernel of the TVP-BVAR Gibbs sampler (bvartvpalg.cpp:282-553).
// Compiled once per BVG_G (lanes per chain; 0 = one CTA per chain).  The block-bidiagonal Cholesky factor of the
// state precision is streamed through HBM (M_t written forward, read backward) with coalesced accesses.
#include "bvg_kernels.cuh"
#ifndef BVG_G
#error "compile with -DBVG_G=<0|4|8|16|32>"
#endif
#define BVG_CAT2(a, b) a##b
#define BVG_CAT(a, b) BVG_CAT2(a, b)

namespace bvg {

#if BVG_G == 32
#define BVG_TVP_BOUNDS __launch_bounds__(256, 1)   // shared memory allows ~7 warps per SM anyway: registers are free
#else
#define BVG_TVP_BOUNDS __launch_bounds__(128, 6)
#endif
#if BVG_G > 0
__global__ void BVG_TVP_BOUNDS BVG_CAT(tvp_tile_kernel_g, BVG_G)(
    const __grid_constant__ TvpModel m, const __grid_constant__ RngParams rp, double* state, int state_len,
    double* scratch, long long scratch_per_chain, const __grid_constant__ TvpOut out, long long n_chains, int s0,
    int s1, int burnin, int smem_per_chain) {
  extern __shared__ double smem[];
  constexpr int G = BVG_G;
  const int cpb = blockDim.x / G;
  const int cib = threadIdx.x / G;
  const long long chain = (long long)blockIdx.x * cpb + cib;
  if (chain >= n_chains) return;
  TileGroup<G> g;
  const ChainRng rng(&rp, chain);
#if defined(BVG_TVP_PROF) && BVG_G == 32
  if (threadIdx.x < 16) s_tvp_prof[threadIdx.x] = 0;
  __syncthreads();
#endif
  tvp_chain(g, m, rng, state + (size_t)chain * state_len, scratch + (size_t)chain * scratch_per_chain,
            smem + (size_t)cib * smem_per_chain, out, chain, s0, s1, burnin);
#if defined(BVG_TVP_PROF) && BVG_G == 32
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < 16) g_tvp_prof[threadIdx.x] += s_tvp_prof[threadIdx.x];
#endif
}
#else
__global__ void BVG_CAT(tvp_tile_kernel_g, BVG_G)(
    const __grid_constant__ TvpModel m, const __grid_constant__ RngParams rp, double* state, int state_len,
    double* scratch, long long scratch_per_chain, const __grid_constant__ TvpOut out, long long n_chains, int s0,
    int s1, int burnin, int smem_per_chain) {
  extern __shared__ double smem[];
  const long long chain = blockIdx.x;
  if (chain >= n_chains) return;
  BlockGroup g(smem);
  const ChainRng rng(&rp, chain);
  tvp_chain(g, m, rng, state + (size_t)chain * state_len, scratch + (size_t)chain * scratch_per_chain,
            smem + BVG_BLOCK_SCRATCH, out, chain, s0, s1, burnin);
}
#endif

cudaError_t BVG_CAT(launch_tvp_g, BVG_G)(const LaunchPlan& plan, const TvpModel& m, const RngParams& rp,
                                         double* state, int state_len, double* scratch,
                                         long long scratch_per_chain, const TvpOut& out, long long n_chains,
                                         int sweep_begin, int sweep_end, int burnin, cudaStream_t stream) {
  cudaError_t e = set_smem(BVG_CAT(tvp_tile_kernel_g, BVG_G), plan.smem_bytes);
  if (e != cudaSuccess) return e;
  BVG_CAT(tvp_tile_kernel_g, BVG_G)<<<plan.grid, plan.block_threads, plan.smem_bytes, stream>>>(
      m, rp, state, state_len, scratch, scratch_per_chain, out, n_chains, sweep_begin, sweep_end, burnin,
      plan.smem_per_chain);
  return cudaGetLastError();
}

}  // namespace bvg

#if defined(BVG_TVP_PROF) && BVG_G == 32
extern "C" int bvg_debug_tvp_prof(long long* out16) {   // measurement builds only (make EXTRA=-DBVG_TVP_PROF)
  return (int)cudaMemcpyFromSymbol(out16, bvg::g_tvp_prof, 16 * sizeof(long long));
}
#endif
