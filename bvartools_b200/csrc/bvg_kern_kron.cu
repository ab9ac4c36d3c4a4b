// bvg_kern_kron.cu -- sm_100a kernel of the flat-prior Kronecker sweep (bvg_kron_sweep.cuh): configs c1 / c2.
//
// Everything random in this sweep is independent of the chain state, and the state recursion is a handful of
// k x k products.  A CTA owns CH chains and works in rounds of NB sweeps:
//   A  variate generation for round r+1 -- Philox4x32-10 + FP64 Box-Muller pairs / Marsaglia-Tsang chi-squares -- as a
//      flat list of (chain, sweep, item) work items handed out to warps through a shared-memory counter
//      (throughput-bound: this is where the FP64 pipe time goes);
//   B  the state recursion W <- U(S0 + SSE + W E E' W') A_B^-1 of round r, one THREAD per chain with W in registers
//      (latency-bound, a few hundred dependent FP64 operations per sweep); its warp joins A when it is done, and
//      the other CTAs resident on the SM fill the issue slots meanwhile;
//   C  one thread per (chain, sweep): the state-independent products of round r+1 (G = E E', A_B^-1, T = E L_X^-1 in
//      place) and the retained draws of round r (a = vec(A_ols + W T), Sigma = W W') staged in place in the records;
//   D  a flat copy of the staged draws to the [chain][iter][row] output: a warp writes consecutive addresses.
// Variates live in shared memory (double-buffered), W per sweep in a small history buffer; HBM sees only the draws.
#include "bvg_kernels.cuh"
#include "bvg_kron_sweep.cuh"

namespace bvg {

#ifndef BVG_KRON_MINB
#define BVG_KRON_MINB 1
#endif
struct KronPlanDims {
  int CH;      // chains per CTA
  int NB;      // sweeps per round
  int VL;      // doubles per (sweep, chain) variate record (odd)
};

// x / d for x * d < 2^32 by one multiply-high (the flat work-item lists are decoded with runtime divisors)
struct FastDiv {
  unsigned d, mg;
  __host__ __device__ explicit FastDiv(unsigned dd) : d(dd), mg((unsigned)(((1ull << 32) + dd - 1) / dd)) {}
  __device__ __forceinline__ unsigned div(unsigned x) const { return d == 1 ? x : __umulhi(x, mg); }
  __device__ __forceinline__ void divmod(unsigned x, unsigned& q, unsigned& r) const { q = div(x); r = x - q * d; }
};

template <int K>
__global__ void __launch_bounds__(512, BVG_KRON_MINB) bvar_kron_batch_kernel(
    const __grid_constant__ BvarModel m, const __grid_constant__ RngParams rp, double* state, int state_len,
    const __grid_constant__ BvarOut out, long long n_chains, int s0, int s1, int burnin,
    const __grid_constant__ KronPlanDims pd) {
  extern __shared__ double smem[];
  __shared__ int work_ctr[2];
  constexpr int KU = (K * (K + 1)) / 2, NW = (K * (K - 1)) / 2, KK = K * K;
  const int M = m.M, n = K * M, CH = pd.CH, NB = pd.NB, VL = pd.VL;
  const int NT = blockDim.x, tid = threadIdx.x, lane = tid & 31;
  // ---- shared memory: constants | records [2][NB][CH][VL] | W history [NB + 1][CH][KU]
  KronConsts cs;
  double* p = smem;
  {
    const int l0 = tri(M), l1 = K * M, l2 = K * K;
    for (int e = tid; e < l0; e += NT) p[e] = m.kron_Linv[e];
    for (int e = tid; e < l1; e += NT) p[l0 + e] = m.Aols[e];
    for (int e = tid; e < l2; e += NT) p[l0 + l1 + e] = m.kron_SS[e];
    cs.Linv = p; cs.Aols = p + l0; cs.SS = p + l0 + l1;
    p += l0 + l1 + l2;
  }
  double* var = p;                           p += 2 * (size_t)NB * CH * VL;
  double* wh = p;
  const long long chain0 = (long long)blockIdx.x * CH;
  const int CHv = (int)((n_chains - chain0 < CH) ? (n_chains - chain0) : CH);   // valid chains of this CTA
  const int NP = kron_pairs(K, M);
  const int items_per_cs = NP + K;           // Box-Muller pairs + chi-squares per (chain, sweep)
  const FastDiv dCH((unsigned)CHv), dNP((unsigned)NP);
  const int o_wn = n, o_chi = n + NW, o_G = n + NW + K, o_Ai = n + NW + K + KU;   // record layout

  // stage 1a: variates of the sweeps [b0, b0 + nb) into record buffer `buf`; work handed out in chunks of 64 items.
  // Item order: chi-squares first (their rejection loop diverges: keep them in the same warps), then Box-Muller
  // pairs, two per lane and call so that two independent dependency chains are in flight.
  auto gen_variates = [&](int b0, int nb, int buf) {
    const int n_chi = (CHv * nb * K + 63) & ~63;             // padded: a 64-chunk holds only one kind of item
    const int n_chi_real = CHv * nb * K;
    const int n_pair = CHv * nb * NP;
    const int total = n_chi + n_pair;
    double* vb = var + (size_t)buf * NB * CH * VL;
    for (;;) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&work_ctr[buf], 64);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base >= total) break;
      if (base < n_chi) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int it = base + 32 * h + lane;
          if (it < n_chi_real) {
            unsigned c, sw;
            const unsigned q = (unsigned)it % K;
            dCH.divmod((unsigned)it / K, sw, c);
            const ChainRng rng(&rp, chain0 + c);
            vb[(sw * (unsigned)CH + c) * (unsigned)VL + o_chi + q] = kron_gen_chi(rng, b0 + sw, q, m.wish_df_post);
          }
        }
      } else {
        const int ia = base - n_chi + lane, ib = ia + 32;
        unsigned ca, swa, qa, ra, cb, swb, qb, rb;
        dNP.divmod((unsigned)(ia < n_pair ? ia : 0), ra, qa);
        dCH.divmod(ra, swa, ca);
        dNP.divmod((unsigned)(ib < n_pair ? ib : 0), rb, qb);
        dCH.divmod(rb, swb, cb);
        const ChainRng rga(&rp, chain0 + ca), rgb(&rp, chain0 + cb);
        double* va = vb + (swa * (unsigned)CH + ca) * (unsigned)VL;
        double* vc = vb + (swb * (unsigned)CH + cb) * (unsigned)VL;
        if (ib < n_pair) kron_gen_pair_x2(rga, b0 + swa, qa, va, rgb, b0 + swb, qb, vc, K, M);
        else if (ia < n_pair) kron_gen_pair(rga, b0 + swa, qa, K, M, va);
      }
    }
  };
  // stage 2a: state-independent products of the `nb` sweeps of buffer `buf`, one thread per (chain, sweep):
  // G = E E', A_B^-1, then T = E L_X^-1 in place
  auto prep = [&](int it, int buf) {
    unsigned sw, c;
    dCH.divmod((unsigned)it, sw, c);
    double* v = var + (size_t)buf * NB * CH * VL + (sw * (unsigned)CH + c) * (unsigned)VL;
    kron_prep<K>(v, v + o_wn, v + o_chi, M, v + o_G, v + o_Ai);
    kron_t_inplace<K>(v, cs.Linv, M);
  };

  // chain state (threads < CHv): W upper in registers; so is the constant S0 + SSE
  double W[K][K], SSu[K][K];
  int fail = 0;
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) SSu[i][j] = m.kron_SS[i + K * j];
  if (tid < CHv) {
    const double* st = state + (size_t)(chain0 + tid) * state_len;
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int j = 0; j < K; ++j) W[i][j] = (i <= j) ? st[i + K * j] : 0.0;
  }
  if (tid < 2) work_ctr[tid] = 0;
  __syncthreads();
  {
    const int nb0 = (s1 - s0 < NB) ? (s1 - s0) : NB;
    gen_variates(s0, nb0, 0);
    __syncthreads();
    for (int it = tid; it < CHv * nb0; it += NT) prep(it, 0);
  }

  int buf = 0;
  for (int b0 = s0; b0 < s1; b0 += NB, buf ^= 1) {
    const int nb = (s1 - b0 < NB) ? (s1 - b0) : NB;
    const int b1 = b0 + nb;
    const int nb_next = (b1 < s1) ? ((s1 - b1 < NB) ? (s1 - b1) : NB) : 0;
    if (tid == 0) work_ctr[buf ^ 1] = 0;     // counter of the buffer filled in this round (idle since the last barrier)
    __syncthreads();
    // ---- stage 1: state recursion of this round (threads < CHv) ...
    double* vb = var + (size_t)buf * NB * CH * VL;
    if (tid < CHv) {
      for (int sw = 0; sw < nb; ++sw) {
        double* wp = wh + (sw * CH + tid) * KU;
#pragma unroll
        for (int j = 0; j < K; ++j)
#pragma unroll
          for (int i = 0; i <= j; ++i) wp[tri(j) + i] = W[i][j];
        const double* v = vb + (sw * CH + tid) * VL;
        kron_step_core<K>(W, v + o_G, v + o_Ai, SSu, fail);
      }
      double* wp = wh + (nb * CH + tid) * KU;
#pragma unroll
      for (int j = 0; j < K; ++j)
#pragma unroll
        for (int i = 0; i <= j; ++i) wp[tri(j) + i] = W[i][j];
    }
    // ---- ... overlapped with the variates of the next round (all warps; the recursion warp joins late)
    if (nb_next > 0) gen_variates(b1, nb_next, buf ^ 1);
    __syncthreads();
    // ---- stage 2: one thread per (chain, sweep): products of the next round; draws of this round, staged in place
    const int k0 = (b0 > burnin) ? b0 : burnin;        // first retained sweep of the round
    const int nk = (b1 > k0) ? b1 - k0 : 0, sw0 = k0 - b0;
    const int n_prep = CHv * nb_next, n_out = CHv * nk;
    for (int it = tid; it < n_prep + n_out; it += NT) {
      if (it < n_prep) { prep(it, buf ^ 1); continue; }
      unsigned swu, c;
      dCH.divmod((unsigned)(it - n_prep), swu, c);
      const int sw = sw0 + (int)swu;
      double* v = vb + (sw * CH + c) * VL;
      kron_coef_inplace<K>(wh + (sw * CH + c) * KU, v, cs.Aols, M);       // a = vec(A_ols + W T), W used by this sweep
      kron_sigma<K>(wh + ((sw + 1) * CH + c) * KU, v + o_G);              // Sigma = W W', W after this sweep
    }
    if (nk > 0) {
      __syncthreads();
      // ---- stage 3: coalesced copy of the staged draws: per chain nk * n (resp. nk * K * K) consecutive doubles;
      //      one warp per chain, lanes over consecutive addresses, (sweep, element) index kept incrementally
      const long long pos0 = (long long)k0 - burnin;
      const int warp = tid >> 5, n_warps = NT >> 5;
      for (int c = warp; c < CHv; c += n_warps) {
        double* dst = out.coef + ((size_t)(chain0 + c) * (size_t)out.iters + (size_t)pos0) * (size_t)n;
        const double* src = vb + (sw0 * CH + c) * VL;
        const int step_e = 32 % n, step_sw = 32 / n;
        int e = lane % n, sw = lane / n;
        for (int rem = lane; rem < nk * n; rem += 32) {
          dst[rem] = src[sw * CH * VL + e];
          e += step_e; sw += step_sw;
          if (e >= n) { e -= n; ++sw; }
        }
        double* dsts = out.sigma + ((size_t)(chain0 + c) * (size_t)out.iters + (size_t)pos0) * (size_t)KK;
        for (int rem = lane; rem < nk * KK; rem += 32)
          dsts[rem] = src[(rem / KK) * CH * VL + o_G + (rem % KK)];
      }
    }
    // (the next round starts with a barrier; its variates go to the other buffer)
  }
  if (tid < CHv) {
    double* st = state + (size_t)(chain0 + tid) * state_len;
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int j = 0; j < K; ++j) st[i + K * j] = (i <= j) ? W[i][j] : 0.0;
    if (fail) out.status[chain0 + tid] = 1;
  }
}

template <int K>
static cudaError_t launch_k(const KronPlanDims& pd, BVG_BVAR_ARGS_NAMED) {
  cudaError_t e = set_smem(bvar_kron_batch_kernel<K>, plan.smem_bytes);
  if (e != cudaSuccess) return e;
  // several CTAs per SM only co-reside when the L1 / shared-memory split favours shared memory
  cudaFuncSetAttribute(bvar_kron_batch_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  bvar_kron_batch_kernel<K><<<plan.grid, plan.block_threads, plan.smem_bytes, stream>>>(
      m, rp, state, state_len, out, n_chains, s0, s1, burnin, pd);
  return cudaGetLastError();
}

cudaError_t launch_bvar_kron(BVG_BVAR_ARGS_NAMED) {
  KronPlanDims pd;
  pd.CH = plan.chains_per_block;
  pd.NB = plan.kron_nb;
  pd.VL = plan.smem_per_chain;
  switch (m.k) {
    case 1: return launch_k<1>(pd, plan, m, rp, state, state_len, out, n_chains, s0, s1, burnin, stream);
    case 2: return launch_k<2>(pd, plan, m, rp, state, state_len, out, n_chains, s0, s1, burnin, stream);
    case 3: return launch_k<3>(pd, plan, m, rp, state, state_len, out, n_chains, s0, s1, burnin, stream);
    case 4: return launch_k<4>(pd, plan, m, rp, state, state_len, out, n_chains, s0, s1, burnin, stream);
    default: return cudaErrorInvalidValue;
  }
}

// Launch geometry.  The kernel runs for the whole launch with a fixed set of CTAs, so what matters is (i) an even
// fill of the SMs -- r CTAs on every SM, CH = ceil(chains / (SMs r)) <= 32 chains each -- and (ii) about 16 warps per
// SM so that the variate phase hides the latency of the state recursion.  threads_per_chain (tuning knob) = CTA
// threads per chain, 0 = auto.
cudaError_t plan_bvar_kron(const BvarModel& m, long long n_chains, int requested_tpc, LaunchPlan* plan) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int CH = 32, NT = 512;
  if (requested_tpc > 0) {
    // tuning / tests: 4..15 -> 128-thread CTAs, 16..128 -> 512-thread CTAs; >= 1000: explicit geometry
    // 1000 * (NT / 128) + CH, e.g. 4028 = 512 threads, 28 chains per CTA
    int tpc = requested_tpc < 4 ? 4 : requested_tpc;
    if (tpc >= 1000) { NT = 128 * (tpc / 1000); CH = tpc % 1000; if (NT > 512) NT = 512; }
    else { if (tpc > 128) tpc = 128; NT = (tpc >= 16) ? 512 : 128; CH = NT / tpc; }
    if (CH > 32) CH = 32;
    if (CH < 1) CH = 1;
  } else {
    int r = 1;                                   // CTAs per SM
    while (r < 4 && (n_chains + (long long)sms * r - 1) / ((long long)sms * r) > 32) ++r;
    long long ch = (n_chains + (long long)sms * r - 1) / ((long long)sms * r);
    CH = (int)(ch > 32 ? 32 : (ch < 1 ? 1 : ch));
    NT = 512 / r;
    if (NT < 128) NT = 128;
    if (NT < ((CH + 31) / 32) * 32) NT = ((CH + 31) / 32) * 32;
  }
  const int k = m.k, M = m.M, KU = (k * (k + 1)) / 2;
  const int VL = kron_variates_len(k, M) | 1;            // odd stride: the threads of phase B hit different banks
  int NB = 8;
  while (NB < 64 && CH * NB < 448) NB <<= 1;      // few chains per CTA: longer rounds keep >= 448 (chain, sweep) items per round
  auto bytes = [&](int nb) {
    return ((size_t)kron_consts_len(k, M) + 2 * (size_t)nb * CH * VL + (size_t)(nb + 1) * CH * KU) * 8;
  };
  while (NB > 1 && bytes(NB) > 200 * 1024 / (512 / NT)) NB >>= 1;
  if (bytes(NB) > 227 * 1024) return cudaErrorInvalidConfiguration;
  plan->kron = 1;
  plan->kron_nb = NB;
  plan->threads_per_chain = NT / CH;
  plan->smem_per_chain = VL;
  plan->block_threads = NT;
  plan->chains_per_block = CH;
  plan->smem_bytes = bytes(NB);
  plan->grid = (int)((n_chains + CH - 1) / CH);
  return cudaSuccess;
}

}  // namespace bvg
