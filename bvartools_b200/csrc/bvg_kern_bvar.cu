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
// CTA per chain.  sched == nullptr: CTA b runs chain b for the whole launch.  sched != nullptr: PERSISTENT mode for chain
// counts that do not fill the last wave (1024 chains on 296 CTA slots = 3.46 waves): the launch is cut into items
// (block of `spi` sweeps, chain), handed out block-major through an atomic counter; an item waits until the same chain's
// previous block has been published (done[chain]), so a chain migrates between CTAs with its state in global memory and
// the tail shrinks from a fraction of the whole launch to a fraction of one item.
__global__ void __launch_bounds__(512, 1) BVG_KNAME(
    const __grid_constant__ BvarModel m, const __grid_constant__ RngParams rp, double* state, int state_len,
    const __grid_constant__ BvarOut out, long long n_chains, int s0, int s1, int burnin, int smem_per_chain,
    int* sched, int spi) {
  extern __shared__ double smem[];
  __shared__ int item_s;
  BlockGroup g(smem);
  const BvarConsts cs = stage_consts(m, smem + BVG_BLOCK_SCRATCH, true);
  double* mine = smem + BVG_BLOCK_SCRATCH + bvar_consts_len(m.k, m.M, true);
  if (sched == nullptr) {
    const long long chain = blockIdx.x;
    if (chain >= n_chains) return;
    const ChainRng rng(&rp, chain);
    bvar_chain<BVG_FEAT, BlockGroup, true>(g, m, cs, rng, state + (size_t)chain * state_len, mine, out, chain, s0, s1, burnin);
    return;
  }
  int* counter = sched;
  volatile int* done = sched + 1;
  const int nblk = (s1 - s0 + spi - 1) / spi;
  const long long n_items = (long long)nblk * n_chains;
  for (;;) {
    if (threadIdx.x == 0) item_s = atomicAdd(counter, 1);
    __syncthreads();
    const long long item = item_s;
    if (item >= n_items) break;
    const int b = (int)(item / n_chains);
    const long long chain = item - (long long)b * n_chains;
    if (threadIdx.x == 0) {
      while (done[chain] < b) __nanosleep(200);
    }
    __syncthreads();
    __threadfence();                                   // acquire: the previous owner's state writes are visible
    const int a0 = s0 + b * spi, a1 = (a0 + spi < s1) ? a0 + spi : s1;
    const ChainRng rng(&rp, chain);
    bvar_chain<BVG_FEAT, BlockGroup, true>(g, m, cs, rng, state + (size_t)chain * state_len, mine, out, chain, a0, a1, burnin);
    __threadfence();                                   // release: state of this chain
    __syncthreads();
    if (threadIdx.x == 0) done[chain] = b + 1;
  }
}
#endif

cudaError_t BVG_LNAME(const LaunchPlan& plan, const BvarModel& m, const RngParams& rp,
                                          double* state, int state_len, const BvarOut& out, long long n_chains,
                                          int sweep_begin, int sweep_end, int burnin, cudaStream_t stream) {
  cudaError_t e = set_smem(BVG_KNAME, plan.smem_bytes);
  if (e != cudaSuccess) return e;
#if BVG_G == 0
  // CTA slots of the device for this launch shape; persistent scheduling only when the last wave would be partial
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, BVG_KNAME, plan.block_threads, plan.smem_bytes) != cudaSuccess ||
      per_sm < 1)
    per_sm = 1;
  const long long slots = (long long)sms * per_sm;
  const int sweeps = sweep_end - sweep_begin;
  const bool tail = n_chains > slots && (n_chains % slots) != 0 && sweeps >= 8;
  if (!tail) {
    BVG_KNAME<<<plan.grid, plan.block_threads, plan.smem_bytes, stream>>>(
        m, rp, state, state_len, out, n_chains, sweep_begin, sweep_end, burnin, plan.smem_per_chain, nullptr, 0);
    return cudaGetLastError();
  }
  int* sched = nullptr;
  if ((e = cudaMallocAsync((void**)&sched, (size_t)(n_chains + 1) * sizeof(int), stream)) != cudaSuccess) return e;
  cudaMemsetAsync(sched, 0, (size_t)(n_chains + 1) * sizeof(int), stream);
  int spi = sweeps / 8;                                // ~8 items per chain and launch
  if (spi < 4) spi = 4;
  BVG_KNAME<<<(int)slots, plan.block_threads, plan.smem_bytes, stream>>>(
      m, rp, state, state_len, out, n_chains, sweep_begin, sweep_end, burnin, plan.smem_per_chain, sched, spi);
  e = cudaGetLastError();
  cudaFreeAsync(sched, stream);
  return e;
#else
  BVG_KNAME<<<plan.grid, plan.block_threads, plan.smem_bytes, stream>>>(
      m, rp, state, state_len, out, n_chains, sweep_begin, sweep_end, burnin, plan.smem_per_chain);
  return cudaGetLastError();
#endif
}

}  // namespace bvg
