// bvg_kern_bvar.cu -- sm_100a kernel of the constant-parameter BVAR Gibbs sampler (bvaralg.cpp:292-560).
// Compiled once per BVG_G (lanes per chain; 0 = one CTA per chain).  One chain is owned by a tile of G <= 32
// lanes (several chains per warp / per CTA) or by a whole CTA; all sweeps of a launch run on-chip with the chain
// state in shared memory; model constants and injected-variate tables stay in kernel-parameter constant space.
#include "bvg_kernels.cuh"
#if !defined(BVG_G) || !defined(BVG_FEAT)
#error "compile with -DBVG_G=<0|4|8|16|32> -DBVG_FEAT=<0|2|31>"
#endif
#define BVG_CAT2(a, b) a##b
#define BVG_CAT(a, b) BVG_CAT2(a, b)
#define BVG_CAT4_(a, b, c, d) a##b##c##d
#define BVG_CAT4(a, b, c, d) BVG_CAT4_(a, b, c, d)
#define BVG_KNAME BVG_CAT4(bvar_kernel_g, BVG_G, _f, BVG_FEAT)
#define BVG_LNAME BVG_CAT4(launch_bvar_g, BVG_G, _f, BVG_FEAT)

namespace bvg {

// copy the read-only tables of the model into the CTA's shared memory (all threads; ends with a CTA barrier)
__device__ __forceinline__ BvarConsts stage_consts(const BvarModel& m, double* dst, bool lean = false) {
  BvarConsts c;
  const int k = m.k, M = m.M, n = m.n;
  const double* src[7] = {m.XX, m.YX, m.Aols, m.SSE, m.S0, m.mu, m.vdiag0};
  const int len[7] = {M * M, k * M, k * M, k * k, k * k, n, n};
  const double** out[7] = {&c.XX, &c.YX, &c.Aols, &c.SSE, &c.S0, &c.mu, &c.vdiag0};
#pragma unroll
  for (int a = 0; a < 7; ++a) {
    if (lean && a >= 5) { *out[a] = src[a]; continue; }           // mu, vdiag0 stay in global memory
    if (src[a] != nullptr)
      for (int e = threadIdx.x; e < len[a]; e += blockDim.x) dst[e] = src[a][e];
    *out[a] = (src[a] != nullptr) ? dst : nullptr;
    dst += len[a];
  }
  __syncthreads();
  return c;
}

#if BVG_G > 0
__global__ void __launch_bounds__(128, 7) BVG_KNAME(
    const __grid_constant__ BvarModel m, const __grid_constant__ RngParams rp, double* state, int state_len,
    const __grid_constant__ BvarOut out, long long n_chains, int s0, int s1, int burnin, int smem_per_chain) {
  extern __shared__ double smem[];
  constexpr int G = BVG_G;
  const int cpb = blockDim.x / G;
  const int cib = threadIdx.x / G;
  const long long chain = (long long)blockIdx.x * cpb + cib;
  const BvarConsts cs = stage_consts(m, smem);
  if (chain >= n_chains) return;
  TileGroup<G> g;
  const ChainRng rng(&rp, chain);
  double* mine = smem + bvar_consts_len(m.k, m.M) + (size_t)cib * smem_per_chain;
  bvar_chain<BVG_FEAT>(g, m, cs, rng, state + (size_t)chain * state_len, mine, out, chain, s0, s1, burnin);
}
#else
__global__ void __launch_bounds__(512, 1) BVG_KNAME(
    const __grid_constant__ BvarModel m, const __grid_constant__ RngParams rp, double* state, int state_len,
    const __grid_constant__ BvarOut out, long long n_chains, int s0, int s1, int burnin, int smem_per_chain) {
  extern __shared__ double smem[];
  const long long chain = blockIdx.x;
  if (chain >= n_chains) return;
  BlockGroup g(smem);
  const BvarConsts cs = stage_consts(m, smem + BVG_BLOCK_SCRATCH, true);
  const ChainRng rng(&rp, chain);
  double* mine = smem + BVG_BLOCK_SCRATCH + bvar_consts_len(m.k, m.M, true);
  bvar_chain<BVG_FEAT, BlockGroup, true>(g, m, cs, rng, state + (size_t)chain * state_len, mine, out, chain, s0, s1, burnin);
}
#endif

cudaError_t BVG_LNAME(const LaunchPlan& plan, const BvarModel& m, const RngParams& rp,
                                          double* state, int state_len, const BvarOut& out, long long n_chains,
                                          int sweep_begin, int sweep_end, int burnin, cudaStream_t stream) {
  cudaError_t e = set_smem(BVG_KNAME, plan.smem_bytes);
  if (e != cudaSuccess) return e;
  BVG_KNAME<<<plan.grid, plan.block_threads, plan.smem_bytes, stream>>>(
      m, rp, state, state_len, out, n_chains, sweep_begin, sweep_end, burnin, plan.smem_per_chain);
  return cudaGetLastError();
}

}  // namespace bvg
