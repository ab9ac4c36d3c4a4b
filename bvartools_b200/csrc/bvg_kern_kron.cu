// bvg_kern_kron.cu -- sm_100a kernel of the flat-prior Kronecker sweep (bvg_kron_sweep.cuh): configs c1 / c2.
//
// Everything random in this sweep is independent of the chain state, and the state recursion is a handful of
// k x k products.  A CTA owns CH chains and works in rounds of NB sweeps; the CH x NB records of a round (one record =
// the variates / products of one sweep of one chain, in shared memory) are split into CHUNKS of 32 records (32 / NB
// chains x NB sweeps), and the warps have fixed roles:
//   warp 0     the state recursion W <- U(S0 + SSE + W E E' W') A_B^-1, one LANE per chain with W in registers:
//              latency-bound (~150 instructions per sweep; k <= 3: the reciprocal square roots of the k x k factor start
//              together, branch-free), so nothing else runs on it; the
//              W a sweep starts from and ends with are left in its record (over the G / A_B^-1 slots just consumed);
//   warp 1     one round behind: the retained draws a = vec(A_ols + W T), Sigma = W W', one lane per record, written
//              straight to the [chain][iter][row] output (consecutive lanes = consecutive sweeps of a chain = consecutive
//              n-vectors in memory; the partial sectors of neighbouring stores merge in L2 -- profiles/r02_c2a_*: DRAM
//              bytes = algorithmic bytes).
//              Staging the draws in shared memory for cp.async.bulk copies was measured and dropped: one bulk copy per
//              chain and array is ~30 instructions of address set-up + R2UR per 0.7 KB, slower than the stores;
//   warps 2 + 2 j, 3 + 2 j   the variates of chunk j, one lane per record, FIXED for the whole launch: chain id, Philox
//              counter words and record address are computed once; a round is a uniform loop over variate SLOTS (Philox
//              counter word, block, the two destinations of a Box-Muller pair), three slots in flight per lane:
//              Philox4x32-10 with constant-bank round keys + the FP64 transforms of bvg_common.cuh.  The A warp draws
//              the first SA slots; the B warp the rest and, after a 64-thread named barrier with A, the Bartlett
//              diagonal (Marsaglia-Tsang, constants precomputed, the attempt-0 test inline, rejections through the
//              generic generator; a rolled loop: code size counts, the generators stall on instruction fetch) and the
//              state-independent products of the chunk (G = E E', A_B^-1, T = E L_X^-1).  Which warp of a pair is A
//              alternates so that every scheduler hosts as many A as B warps.
// Rounds are pipelined over three record generations: the generators run up to two rounds ahead of the recursion.  All
// cross-role hand-offs are mbarriers in shared memory (arrive = release, try_wait = acquire, waiting warps are parked by
// the hardware); there is no CTA-wide barrier after the set-up.  HBM sees only the draws.
#include "bvg_kernels.cuh"
#include "bvg_kron_sweep.cuh"

namespace bvg {

#ifndef BVG_KRON_MINB
#define BVG_KRON_MINB 1
#endif
#ifndef BVG_KRON_NG
#define BVG_KRON_NG 3      // Box-Muller slots in flight per generator lane (measured 1, 2, 3, 4, 6: see the generator loop)
#endif
struct KronPlanDims {
  int CH;      // chains per CTA
  int NB;      // sweeps per round
  int VL;      // doubles per record (odd)
  int CS;      // record stride between chains (odd): record (c, sw) at c CS + sw VL
  int SA;      // variate slots drawn by the A warp of a chunk
  uint32_t rk[20];   // Philox round keys: (k0 + r W0, k1 + r W1), r = 0..9
};

// Philox4x32-10 with the round keys as constant-bank operands of the xor (no key schedule in the instruction stream)
__device__ __forceinline__ Philox4 philox_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&rk)[20]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ rk[2 * r], n2 = hi0 ^ c3 ^ rk[2 * r + 1];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

// state-independent products of one record with E in registers (MM = M at compile time): G = E E', A_B^-1, T = E L_X^-1
template <int K, int MM>
__device__ __forceinline__ void kron_prep_regs(double* rec, const double* wn, const double* chi, const double* Linv,
                                               double* Gp, double* Aip) {
  double E[MM][K];
#pragma unroll
  for (int j = 0; j < MM; ++j)
#pragma unroll
    for (int i = 0; i < K; ++i) E[j][i] = rec[i + K * j];
  double Ai[K][K];
#pragma unroll
  for (int c = 0; c < K; ++c) {
    Ai[c][c] = 1.0 / chi[c];
#pragma unroll
    for (int r = c - 1; r >= 0; --r) {
      double v = 0.0;
#pragma unroll
      for (int mm = r + 1; mm <= c; ++mm) v = fma(wn[tri(mm - 1) + r], Ai[mm][c], v);
      Ai[r][c] = -v * Ai[r][r];
    }
  }
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double g = 0.0;
#pragma unroll
      for (int c = 0; c < MM; ++c) g = fma(E[c][i], E[c][j], g);
      Gp[tri(i) + j] = g;
      Aip[tri(i) + j] = Ai[j][i];
    }
#pragma unroll
  for (int j = 0; j < MM; ++j) {
    double t[K];
#pragma unroll
    for (int i = 0; i < K; ++i) t[i] = 0.0;
#pragma unroll
    for (int j2 = j; j2 < MM; ++j2) {
      const double li = Linv[tri(j2) + j];
#pragma unroll
      for (int i = 0; i < K; ++i) t[i] = fma(E[j2][i], li, t[i]);
    }
#pragma unroll
    for (int i = 0; i < K; ++i) rec[i + K * j] = t[i];
  }
}
// a = vec(A_ols + W T) -> dst, M at compile time (MM) or at run time (MM = 0)
template <int K, int MM>
__device__ __forceinline__ void kron_coef_to(const double* Wp, const double* tv, const double* Aols, int M, double* dst) {
  double W[K][K];
#pragma unroll
  for (int j = 0; j < K; ++j)
#pragma unroll
    for (int i = 0; i <= j; ++i) W[i][j] = Wp[tri(j) + i];
  auto column = [&](int j) {
    double t[K];
#pragma unroll
    for (int i = 0; i < K; ++i) t[i] = tv[i + K * j];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      double d = Aols[i + K * j];
#pragma unroll
      for (int i2 = i; i2 < K; ++i2) d = fma(W[i][i2], t[i2], d);
      dst[i + K * j] = d;
    }
  };
  if (MM > 0) {
#pragma unroll
    for (int j = 0; j < MM; ++j) column(j);
  } else {
    for (int j = 0; j < M; ++j) column(j);
  }
}

// mbarrier hand-offs (shared memory, CTA scope): arrive = release, wait = acquire; waiting warps are parked by the hardware
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait parks the warp only for a short, implementation-defined time: the role warps that mostly wait (recursion,
// output) would otherwise spend a third of the SM's issued instructions on re-polls (profiles/r02_c2a_final_*)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity, unsigned sleep_ns = 0) {
  while (!mbar_try(bar, parity))
    if (sleep_ns) __nanosleep(sleep_ns);
}

template <int K>
__global__ void __launch_bounds__(512, BVG_KRON_MINB) bvar_kron_batch_kernel(
    const __grid_constant__ BvarModel m, const __grid_constant__ RngParams rp, double* state, int state_len,
    const __grid_constant__ BvarOut out, long long n_chains, int s0, int s1, int burnin,
    const __grid_constant__ KronPlanDims pd) {
  extern __shared__ __align__(16) double smem[];
  // hand-offs, one pair per record generation: prep_bar[g] completes when every chunk of a round in generation g is
  // prepared (one arrival per B warp), rec_bar[g] when the recursion is through it, out_bar[g] when its draws are written
  // and the records may be overwritten (one arrival each)
  __shared__ uint64_t prep_bar[3], rec_bar[3], out_bar[3];
  constexpr int KU = (K * (K + 1)) / 2, NW = (K * (K - 1)) / 2, KK = K * K;
  constexpr int PW = (NW + 1) / 2, PX = (K + 1) / 2;
  const int M = m.M, n = K * M, CH = pd.CH, NB = pd.NB, VL = pd.VL, CS = pd.CS;
  const int PC = (n + 1) >> 1;
  const int NT = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- record layout: eps (2 PC) | G (KU) | A_B^-1 (KU) [| pad]; until the chunk is prepared the tail holds the
  //      Bartlett normals (2 PW, at o_wn) and the attempt-0 normals of the Bartlett diagonal (2 PX, at o_x0)
  const int o_G = 2 * PC, o_Ai = o_G + KU, o_wn = o_G, o_x0 = o_G + 2 * PW;
  // ---- shared memory: records [3][CH CS] | constants | MT constants
  double* var = smem;
  double* p = var + 3 * (size_t)CH * CS;
  KronConsts cs;
  {
    const int l0 = tri(M), l1 = K * M, l2 = K * K;
    for (int e = tid; e < l0; e += NT) p[e] = m.kron_Linv[e];
    for (int e = tid; e < l1; e += NT) p[l0 + e] = m.Aols[e];
    for (int e = tid; e < l2; e += NT) p[l0 + l1 + e] = m.kron_SS[e];
    cs.Linv = p; cs.Aols = p + l0; cs.SS = p + l0 + l1;
    p += l0 + l1 + l2;
  }
  double* mt = p;      // [K][2]: d, c of Marsaglia-Tsang for chi2(df - i) = 2 Gamma((df - i) / 2)
  if (tid < K) {
    const double a = 0.5 * (m.wish_df_post - (double)tid);
    const double d = a - 1.0 / 3.0;
    mt[2 * tid] = d;
    mt[2 * tid + 1] = 1.0 / sqrt(9.0 * d);
  }
  const long long chain0 = (long long)blockIdx.x * CH;
  const int CHv = (int)((n_chains - chain0 < CH) ? (n_chains - chain0) : CH);   // valid chains of this CTA
  const int CPC = 32 / NB;                                                      // chains per chunk
  const int NCHv = (CHv + CPC - 1) / CPC;                                       // valid chunks
  if (tid == 0) {
    for (int g = 0; g < 3; ++g) {
      mbar_init(&prep_bar[g], (unsigned)NCHv);
      mbar_init(&rec_bar[g], 1u);
      mbar_init(&out_bar[g], 1u);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int n_rounds = (s1 - s0 + NB - 1) / NB;
  auto round_nb = [&](int q) { return (s1 - (s0 + q * NB) < NB) ? (s1 - (s0 + q * NB)) : NB; };

  if (warp == 0) {
    // ================= state recursion =================
    double W[K][K], SSu[K][K];
    int fail = 0;
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int j = 0; j < K; ++j) SSu[i][j] = m.kron_SS[i + K * j];
    const bool act = tid < CHv;
    if (act) {
      const double* st = state + (size_t)(chain0 + tid) * state_len;
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) W[i][j] = (i <= j) ? st[i + K * j] : 0.0;
    }
    for (int q = 0; q < n_rounds; ++q) {
      mbar_wait(&prep_bar[q % 3], (unsigned)((q / 3) & 1), 40);
      const int nb = round_nb(q);
      if (act) {
        double* rec = var + (size_t)(q % 3) * CH * CS + tid * CS;
        double Gp[KU], Aip[KU];
#pragma unroll
        for (int e = 0; e < KU; ++e) { Gp[e] = rec[o_G + e]; Aip[e] = rec[o_Ai + e]; }
        for (int sw = 0; sw < nb; ++sw, rec += VL) {
#pragma unroll
          for (int j = 0; j < K; ++j)
#pragma unroll
            for (int i = 0; i <= j; ++i) rec[o_Ai + tri(j) + i] = W[i][j];        // W this sweep starts from
          double Gn[KU], Ain[KU];                              // next sweep's inputs: loads in flight during this
          const double* nx = rec + ((sw + 1 < nb) ? VL : 0);   //   sweep's dependent chain
#pragma unroll
          for (int e = 0; e < KU; ++e) { Gn[e] = nx[o_G + e]; Ain[e] = nx[o_Ai + e]; }
          kron_step_core<K>(W, Gp, Aip, SSu, fail);
#pragma unroll
          for (int j = 0; j < K; ++j)
#pragma unroll
            for (int i = 0; i <= j; ++i) rec[o_G + tri(j) + i] = W[i][j];         // W after this sweep
#pragma unroll
          for (int e = 0; e < KU; ++e) { Gp[e] = Gn[e]; Aip[e] = Ain[e]; }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&rec_bar[q % 3]);
    }
    if (act) {
      double* st = state + (size_t)(chain0 + tid) * state_len;
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) st[i + K * j] = (i <= j) ? W[i][j] : 0.0;
      if (fail) out.status[chain0 + tid] = 1;
    }
    return;
  }

  if (warp == 1) {
    // ================= retained draws, one round behind the recursion: lane = record =================
    // (records c NB + sw in passes of 32: full lanes whatever the number of chains of the CTA; consecutive lanes are
    //  consecutive sweeps of a chain, i.e. consecutive n-vectors of the output)
    const int lnb = 31 - __clz(NB);
    for (int q = 0; q < n_rounds; ++q) {
      mbar_wait(&rec_bar[q % 3], (unsigned)((q / 3) & 1), 200);
      const int nb = round_nb(q), b0 = s0 + q * NB;
      if (b0 + nb > burnin) {
        const double* vb = var + (size_t)(q % 3) * CH * CS;
        for (int r = lane; r < CHv * NB; r += 32) {
          const int c = r >> lnb, sw = r & (NB - 1);
          const long long pos = (long long)b0 + sw - burnin;
          if (sw >= nb || pos < 0) continue;
          const double* rec = vb + c * CS + sw * VL;
          double* oc = out.coef + ((size_t)(chain0 + c) * (size_t)out.iters + (size_t)pos) * (size_t)n;
          double* os = out.sigma + ((size_t)(chain0 + c) * (size_t)out.iters + (size_t)pos) * (size_t)KK;
          if (M == 2 * K + 1) kron_coef_to<K, 2 * K + 1>(rec + o_Ai, rec, cs.Aols, M, oc);   // a = vec(A_ols + W T)
          else if (M == K + 1) kron_coef_to<K, K + 1>(rec + o_Ai, rec, cs.Aols, M, oc);
          else if (M == 3 * K + 1) kron_coef_to<K, 3 * K + 1>(rec + o_Ai, rec, cs.Aols, M, oc);
          else if (M == 4 * K + 1) kron_coef_to<K, 4 * K + 1>(rec + o_Ai, rec, cs.Aols, M, oc);
          else kron_coef_to<K, 0>(rec + o_Ai, rec, cs.Aols, M, oc);
          kron_sigma<K>(rec + o_G, os);                                                      // Sigma = W W', W after the sweep
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&out_bar[q % 3]);
    }
    return;
  }

  // ================= variates and state-independent products of chunk j: lane = record =================
  const int j = (warp - 2) >> 1;
  // A (11 slots of pure Box-Muller) is the FP64-heavier role: alternate which warp of the pair takes it so that every
  // scheduler (warp id mod 4) hosts as many A as B warps
  const bool isB = (((warp & 1) != 0) != (((j >> 1) & 1) != 0));
  if (j * CPC >= CHv) return;                                  // (both warps of an empty chunk)
  int c = j * CPC + lane / NB;
  const int sw_l = lane & (NB - 1);
  const bool dup_c = c >= CHv;                                 // lanes past the last chain repeat its records
  if (dup_c) c = CHv - 1;
  const uint64_t chain = rp.chain_offset + (uint64_t)(chain0 + c);
  const uint32_t c2 = (uint32_t)chain, c3 = (uint32_t)(chain >> 32);
  const bool inj_coef = rp.inj[VB_COEF_N] != nullptr, inj_wn = rp.inj[VB_WISH_N] != nullptr;
  const bool inj_chi = rp.inj[VB_WISH_C] != nullptr;
  const int NSL = PC + PW + PX;                                // Box-Muller slots: coefficient normals, Bartlett normals,
                                                               //   attempt-0 normals of the Bartlett diagonal
  const int SA = pd.SA < NSL ? pd.SA : NSL;                    // slots [0, SA) on the A warp, the rest on B
  // slot: x = Philox counter word 0 (pair index), y = block << 24, z / w = destinations of the pair members
  auto slot_of = [&](int s) {
    int4 d;
    if (s < PC) { d.x = s; d.y = VB_COEF_N << 24; d.z = 2 * s; d.w = 2 * s + 1; }
    else if (s < PC + PW) { d.x = s - PC; d.y = VB_WISH_N << 24; d.z = o_wn + 2 * (s - PC); d.w = d.z + 1; }
    else { d.x = s - PC - PW; d.y = VB_WISH_C << 24; d.z = o_x0 + 2 * (s - PC - PW); d.w = d.z + 1; }  // x0 of chi i at o_x0 + i
    return d;
  };
  for (int q = 0; q < n_rounds; ++q) {
    if (q >= 3) mbar_wait(&out_bar[q % 3], (unsigned)((q / 3 - 1) & 1));      // round q - 3 (same generation) is written out
    const int nb = round_nb(q);
    const bool dup = dup_c || sw_l >= nb;                      // (tail round: lanes past its last sweep repeat it)
    const int sw = (sw_l < nb) ? sw_l : nb - 1;
    const int sweep = s0 + q * NB + sw;
    double* rec = var + (size_t)(q % 3) * CH * CS + c * CS + sw * VL;
    const int sb = isB ? SA : 0, se = isB ? NSL : SA;
    if (inj_coef || inj_wn || inj_chi) {
      for (int s = sb; s < se; ++s) {
        const int4 d = slot_of(s);
        const int blk = d.y >> 24, cnt = (blk == VB_COEF_N) ? n : NW;
        if (rp.inj[blk] != nullptr) {
          if (blk == VB_WISH_C) continue;                      // (injected chi-squares are read below)
          const double* src = rp.inj[blk] + (size_t)(chain0 + c) * (size_t)rp.inj_chain_stride[blk] +
                              (size_t)sweep * (size_t)rp.inj_stride[blk];
          rec[d.z] = src[2 * d.x];
          rec[d.w] = (2 * d.x + 1 < cnt) ? src[2 * d.x + 1] : 0.0;
        } else {
          const Philox4 ph = philox_rk((uint32_t)d.x, (uint32_t)sweep, c2, c3 ^ (uint32_t)d.y, pd.rk);
          bvg_box_muller(u53(ph.x, ph.y), u53(ph.z, ph.w), &rec[d.z], &rec[d.w]);
        }
      }
    } else {
      // NG slots in flight per lane: NG independent Philox / log / sincos dependency chains.  One group size only (the
      // last group of a warp repeats its last slot): the 4 / 2 / 1 ladder that avoided the repeat tripled the code of the
      // loop and cost more in instruction-cache misses (12 % of the generators' stall samples) than the repeat does.
      // Measured on c2-A (ms per step): NG = 1: 5.49, 2: 5.45, 3: 5.41, 4: 5.58, 6: 6.78; an out-of-line slot function
      // (smallest code, pointers through the generic address space): 7.0
      constexpr int NG = BVG_KRON_NG;
      for (int s = sb; s < se; s += NG) {
        int4 d[NG];
        Philox4 ph[NG];
#pragma unroll
        for (int i = 0; i < NG; ++i) {
          d[i] = slot_of((s + i < se) ? s + i : se - 1);
          ph[i] = philox_rk((uint32_t)d[i].x, (uint32_t)sweep, c2, c3 ^ (uint32_t)d[i].y, pd.rk);
        }
        double za[NG], zb[NG];
#pragma unroll
        for (int i = 0; i < NG; ++i) bvg_box_muller(u53(ph[i].x, ph[i].y), u53(ph[i].z, ph[i].w), &za[i], &zb[i]);
#pragma unroll
        for (int i = 0; i < NG; ++i) { rec[d[i].z] = za[i]; rec[d[i].w] = zb[i]; }
      }
    }
    __threadfence_block();
    asm volatile("bar.sync %0, 64;" ::"r"(1 + j) : "memory");  // the chunk's normals are complete (A and B)
    if (isB) {
      // ---- Bartlett diagonal: sqrt(chi2(df - i)), i < K, in place over the attempt-0 normals.  Rolled loop: the body
      //      (Philox, two logs, the rare rejection call) is ~300 instructions, and instruction fetch is what the generator
      //      warps stall on next to the FP64 pipe (lanes that repeat a record skip the whole block)
      if (!dup) {
#pragma unroll 1
        for (int i = 0; i < K; ++i) {
          double ch;
          if (inj_chi) {
            ch = sqrt(rp.inj[VB_WISH_C][(size_t)(chain0 + c) * (size_t)rp.inj_chain_stride[VB_WISH_C] +
                                        (size_t)sweep * (size_t)rp.inj_stride[VB_WISH_C] + i]);
          } else {
            const Philox4 pu = philox_rk((uint32_t)i | (100u << 24), (uint32_t)sweep, c2, c3 ^ ((uint32_t)VB_WISH_C << 24), pd.rk);
            const double d = mt[2 * i], cc = mt[2 * i + 1], x = rec[o_x0 + i];
            const double u = u53(pu.x, pu.y);
            const double v1 = fma(cc, x, 1.0), v = v1 * v1 * v1, x2 = x * x;
            const double vs = (v1 > 0.0) ? v : 1.0;
            const bool ok = (v1 > 0.0) && (d > 2.0 / 3.0) &&
                            ((u < 1.0 - 0.0331 * x2 * x2) || (bvg_logpos(u) < 0.5 * x2 + d - d * v + d * bvg_logpos(vs)));
            double g = d * v;
            if (!ok) g = philox_gamma_from_x0(rp.seed, chain, VB_WISH_C, sweep, i, 0.5 * (m.wish_df_post - (double)i), x);
            ch = bvg_sqrtpos(2.0 * g);
          }
          rec[o_x0 + i] = ch;
        }
        double chi[K];
#pragma unroll
        for (int i = 0; i < K; ++i) chi[i] = rec[o_x0 + i];
        double wn[NW > 0 ? NW : 1];
#pragma unroll
        for (int i = 0; i < NW; ++i) wn[i] = rec[o_wn + i];
        if (K * M <= 40 && M == 2 * K + 1) kron_prep_regs<K, 2 * K + 1>(rec, wn, chi, cs.Linv, rec + o_G, rec + o_Ai);
        else if (K * M <= 40 && M == K + 1) kron_prep_regs<K, K + 1>(rec, wn, chi, cs.Linv, rec + o_G, rec + o_Ai);
        else if (K * M <= 40 && M == 4 * K + 1) kron_prep_regs<K, (K * (4 * K + 1) <= 40) ? 4 * K + 1 : 1>(rec, wn, chi, cs.Linv, rec + o_G, rec + o_Ai);
        else if (K * M <= 40 && M == 3 * K + 1) kron_prep_regs<K, (K * (3 * K + 1) <= 40) ? 3 * K + 1 : 1>(rec, wn, chi, cs.Linv, rec + o_G, rec + o_Ai);
        else {
          kron_prep<K>(rec, wn, chi, M, rec + o_G, rec + o_Ai);
          kron_t_inplace<K>(rec, cs.Linv, M);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&prep_bar[q % 3]);
    }
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

// record and staging geometry shared by the planner and the launcher
static void kron_dims(int k, int M, int CH, int NB, KronPlanDims* pd) {
  const int n = k * M, NW = (k * (k - 1)) / 2, KU = (k * (k + 1)) / 2;
  const int tail_raw = 2 * ((NW + 1) / 2) + 2 * ((k + 1) / 2), tail = tail_raw > 2 * KU ? tail_raw : 2 * KU;
  const int len = 2 * ((n + 1) / 2) + tail;
  pd->CH = CH;
  pd->NB = NB;
  pd->VL = len | 1;                         // odd strides: the lanes of a warp (consecutive records / chains) hit
  pd->CS = (NB * pd->VL) | 1;               //   different banks
  // split of the Box-Muller slots between the two warps of a chunk: B also draws the Bartlett diagonal (about five
  // slots' worth of instructions) and prepares the chunk (about two)
  const int nsl = (n + 1) / 2 + (NW + 1) / 2 + (k + 1) / 2;
  int sa = (nsl + 7 + 1) / 2;               // B: its slots + ~7 slots' worth of extra work (measured: c2-A, SA = 10..14)
  sa = ((sa + BVG_KRON_NG - 1) / BVG_KRON_NG) * BVG_KRON_NG;   // whole groups on A (a partial group repeats a slot)
  if (sa > nsl) sa = nsl;
  pd->SA = sa;
}
static size_t kron_smem_bytes(int k, int M, const KronPlanDims& pd) {
  return (3 * (size_t)pd.CH * pd.CS + (size_t)kron_consts_len(k, M) + 2 * k) * 8;
}

cudaError_t launch_bvar_kron(BVG_BVAR_ARGS_NAMED) {
  KronPlanDims pd;
  kron_dims(m.k, m.M, plan.chains_per_block, plan.kron_nb, &pd);
  if (plan.kron_sa > 0) pd.SA = plan.kron_sa;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t k0 = (uint32_t)rp.seed, k1 = (uint32_t)(rp.seed >> 32);
  for (int r = 0; r < 10; ++r) { pd.rk[2 * r] = k0; pd.rk[2 * r + 1] = k1; k0 += W0; k1 += W1; }
  switch (m.k) {
    case 1: return launch_k<1>(pd, plan, m, rp, state, state_len, out, n_chains, s0, s1, burnin, stream);
    case 2: return launch_k<2>(pd, plan, m, rp, state, state_len, out, n_chains, s0, s1, burnin, stream);
    case 3: return launch_k<3>(pd, plan, m, rp, state, state_len, out, n_chains, s0, s1, burnin, stream);
    case 4: return launch_k<4>(pd, plan, m, rp, state, state_len, out, n_chains, s0, s1, burnin, stream);
    default: return cudaErrorInvalidValue;
  }
}

// Launch geometry.  A CTA holds NCH <= 7 chunks of 32 records (32 / NB chains x NB sweeps) and runs 2 + 2 NCH warps; one
// CTA per SM (shared memory).  Chains per SM decide the chunk shape: the fewer chains, the more sweeps per round, so that
// a round always offers 32 NCH records to the generators.  threads_per_chain (tuning knob): >= 1000 is an explicit
// geometry 1000 NB + NCH (e.g. 8007); other values are ignored (the kernel has one layout).
cudaError_t plan_bvar_kron(const BvarModel& m, long long n_chains, int requested_tpc, LaunchPlan* plan) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int k = m.k, M = m.M;
  int NB = 0, NCH = 0;
  int sa_override = 0;
  if (requested_tpc >= 1000) {
    sa_override = requested_tpc / 100000;                      // 100000 SA + 1000 NB + NCH
    requested_tpc %= 100000;
    NB = requested_tpc / 1000;
    NCH = requested_tpc % 1000;
    if (NB < 1 || NB > 32 || (NB & (NB - 1)) != 0) return cudaErrorInvalidConfiguration;
    if (NCH < 1) NCH = 1;
    if (NCH > 7) NCH = 7;
  } else {
    long long cps = (n_chains + sms - 1) / sms;              // chains per SM
    if (cps > 28) cps = 28;
    if (cps < 1) cps = 1;
    int CPC = 1;
    while ((cps + CPC - 1) / CPC > 7) CPC <<= 1;
    NB = 32 / CPC;
    NCH = (int)((cps + CPC - 1) / CPC);
  }
  KronPlanDims pd;
  auto bytes = [&](int nch) {
    kron_dims(k, M, nch * (32 / NB), NB, &pd);
    return kron_smem_bytes(k, M, pd);
  };
  while (NCH > 1 && bytes(NCH) > 227 * 1024 - 256) --NCH;
  if (bytes(NCH) > 227 * 1024 - 256) return cudaErrorInvalidConfiguration;
  const int CH = NCH * (32 / NB);
  plan->kron = 1;
  plan->kron_nb = NB;
  plan->block_threads = 64 + 64 * NCH;
  plan->threads_per_chain = plan->block_threads / CH;
  plan->kron_sa = sa_override;
  plan->smem_per_chain = pd.VL;
  plan->chains_per_block = CH;
  plan->smem_bytes = bytes(NCH);
  plan->grid = (int)((n_chains + CH - 1) / CH);
  return cudaSuccess;
}

}  // namespace bvg
