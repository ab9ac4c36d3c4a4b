// bvg_kern_kron.cu -- sm_100a kernel of the flat-prior Kronecker sweep (bvg_kron_sweep.cuh): configs c1 / c2.
//
// Everything random in this sweep is independent of the chain state, and the state recursion is a handful of
// k x k products.  A CTA owns CH chains and works in rounds of NB sweeps; one round q does
//   1  (threads < CH, one THREAD per chain, W in registers) the state recursion W <- U(S0 + SSE + W E E' W') A_B^-1 of
//      the sweeps of round q: latency-bound, ~150 dependent instructions per sweep; W before / after the step are left
//      in the record of the sweep (over the G / A_B^-1 slots it has just consumed);
//   G  (all warps; the recursion warp joins when it is done) the variates of round q + 1, as warp-sized work units
//      handed out through a shared-memory counter.  A unit is (two variate SLOTS) x (32 records): the slot -- Philox
//      counter word, variate block, the two destinations of a Box-Muller pair inside the record -- is warp-uniform and
//      comes from a small descriptor table, so that a lane runs nothing but Philox4x32-10 with constant-bank round keys,
//      the FP64 transforms of bvg_common.cuh (two independent dependency chains per lane) and four STS.  Chi-square units
//      (one lane per record) draw the Bartlett diagonal: Marsaglia-Tsang with the constants d, c of each diagonal
//      position precomputed, the attempt-0 acceptance test inline, the rare rejection through the generic generator;
//   -- barrier --
//   2  one thread per record: state-independent products of round q + 1 (G = E E', A_B^-1, T = E L_X^-1 in place;
//      E in registers when M is one of the compile-time shapes) and the retained draws of round q
//      (a = vec(A_ols + W T), Sigma = W W') written DENSE per chain into a staging buffer;
//   -- barrier --
//   3  thread c hands the staged draws of chain c to the copy engine: two cp.async.bulk (shared -> global) per chain
//      and round, nk n and nk k^2 consecutive doubles (16-byte alignment permitting; a plain coalesced copy otherwise).
// Variates live in shared memory (double-buffered records), HBM sees only the draws.
#include "bvg_kernels.cuh"
#include "bvg_kron_sweep.cuh"

namespace bvg {

#ifndef BVG_KRON_MINB
#define BVG_KRON_MINB 1
#endif
struct KronPlanDims {
  int CH;      // chains per CTA
  int NB;      // sweeps per round
  int VL;      // doubles per record (odd)
  int CS;      // record stride between chains (odd): record (c, sw) at c CS + sw VL
  int SC, SG;  // staging strides per chain (even): coefficients, Sigma
  int skip0;   // 1: the warps 4, 8, 12 (scheduler of the recursion warp) do not work
  uint32_t rk[20];   // Philox round keys: (k0 + r W0, k1 + r W1), r = 0..9
};

// x / d for x * d < 2^32 by one multiply-high
struct FastDiv {
  unsigned d, mg;
  __host__ __device__ explicit FastDiv(unsigned dd) : d(dd), mg(dd > 1 ? 0xFFFFFFFFu / dd + 1u : 0u) {}   // ceil(2^32 / d)
  __device__ __forceinline__ unsigned div(unsigned x) const { return d == 1 ? x : __umulhi(x, mg); }
  __device__ __forceinline__ void divmod(unsigned x, unsigned& q, unsigned& r) const { q = div(x); r = x - q * d; }
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

__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}

template <int K>
__global__ void __launch_bounds__(512, BVG_KRON_MINB) bvar_kron_batch_kernel(
    const __grid_constant__ BvarModel m, const __grid_constant__ RngParams rp, double* state, int state_len,
    const __grid_constant__ BvarOut out, long long n_chains, int s0, int s1, int burnin,
    const __grid_constant__ KronPlanDims pd) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int work_ctr[1];
  __shared__ int flags[2];      // [0] rounds prepared, [1] rounds through the recursion (monotone; polled)
  constexpr int KU = (K * (K + 1)) / 2, NW = (K * (K - 1)) / 2, KK = K * K;
  constexpr int PW = (NW + 1) / 2, PX = (K + 1) / 2;
  const int M = m.M, n = K * M, CH = pd.CH, NB = pd.NB, VL = pd.VL, CS = pd.CS;
  const int PC = (n + 1) >> 1;
  const int NT = blockDim.x, tid = threadIdx.x, lane = tid & 31;
  // ---- record layout: eps (2 PC) | Bartlett normals (2 PW) | G (KU) | A_B^-1 (KU); the chi-square slots overlay G
  const int o_wn = 2 * PC, o_G = o_wn + 2 * PW, o_Ai = o_G + KU, o_chi = o_G;
  // ---- shared memory: staging (coef [CH][SC], Sigma [CH][SG]) | records [3][CH CS] | constants | MT constants
  double* st_coef = smem;
  double* st_sig = st_coef + (size_t)CH * pd.SC;
  double* var = st_sig + (size_t)CH * pd.SG;
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
  // slot descriptor: x = Philox counter word 0 (pair index), y = block << 24, z / w = destinations of the pair members
  const int NSL = PC + PW;                    // Box-Muller slots of the pair units
  const int NSLP = (NSL + 1) >> 1;            // slot pairs (the last one may repeat its first member)
  auto slot_of = [&](int s) {
    int4 d;
    if (s < PC) { d.x = s; d.y = VB_COEF_N << 24; d.z = 2 * s; d.w = 2 * s + 1; }
    else { d.x = s - PC; d.y = VB_WISH_N << 24; d.z = o_wn + 2 * (s - PC); d.w = d.z + 1; }
    return d;
  };
  const FastDiv dNB((unsigned)NB);
  auto div_of = [&](int x) { return (x == NB) ? dNB : FastDiv((unsigned)(x > 0 ? x : 1)); };
  const long long chain0 = (long long)blockIdx.x * CH;
  const int CHv = (int)((n_chains - chain0 < CH) ? (n_chains - chain0) : CH);   // valid chains of this CTA
  const uint64_t chain_g0 = rp.chain_offset + (uint64_t)chain0;
  const bool inj_coef = rp.inj[VB_COEF_N] != nullptr, inj_wn = rp.inj[VB_WISH_N] != nullptr;
  const bool inj_chi = rp.inj[VB_WISH_C] != nullptr;

  // records r = c nb + sw of a round of nb sweeps
  // ---- G: variates of the sweeps [b0, b0 + nb) into record buffer `buf`
  auto gen_variates = [&](int b0, int nb, int buf) {
    const int R = CHv * nb, RW = (R + 31) >> 5;
    const int n_units = RW * (1 + NSLP);                       // chi-square units first (they hold the divergent code)
    const FastDiv dnb = div_of(nb), dRW((unsigned)RW);
    double* vb = var + (size_t)buf * CH * CS;
    for (;;) {
      int unit = 0;
      if (lane == 0) unit = atomicAdd(&work_ctr[0], 1);
      unit = __shfl_sync(0xffffffffu, unit, 0);
      if (unit >= n_units) break;
      unsigned uq, ur;
      dRW.divmod((unsigned)unit, uq, ur);
      const int sp = (int)uq - 1;                              // -1: chi-square unit; else slot pair
      int r = (int)ur * 32 + lane;
      if (r >= R) r = R - 1;                                   // (idle lanes repeat the last record)
      unsigned c, sw;
      dnb.divmod((unsigned)r, c, sw);
      double* rec = vb + c * (unsigned)CS + sw * (unsigned)VL;
      const int sweep = b0 + (int)sw;
      const uint64_t chain = chain_g0 + c;
      const uint32_t c2 = (uint32_t)chain, c3 = (uint32_t)(chain >> 32);
      if (sp < 0) {
        // ---- Bartlett diagonal: sqrt(chi2(df - i)), i < K
        if (inj_chi) {
#pragma unroll
          for (int i = 0; i < K; ++i)
            rec[o_chi + i] = sqrt(rp.inj[VB_WISH_C][(size_t)(chain0 + c) * (size_t)rp.inj_chain_stride[VB_WISH_C] +
                                                    (size_t)sweep * (size_t)rp.inj_stride[VB_WISH_C] + i]);
          continue;
        }
        double x0[2 * PX];
#pragma unroll
        for (int q = 0; q < PX; ++q) {                          // attempt-0 normals: pairs q of block VB_WISH_C
          const Philox4 ph = philox_rk((uint32_t)q, (uint32_t)sweep, c2, c3 ^ ((uint32_t)VB_WISH_C << 24), pd.rk);
          bvg_box_muller(u53(ph.x, ph.y), u53(ph.z, ph.w), &x0[2 * q], &x0[2 * q + 1]);
        }
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const double d = mt[2 * i], cc = mt[2 * i + 1], x = x0[i];
          const Philox4 ph = philox_rk((uint32_t)i | (100u << 24), (uint32_t)sweep, c2, c3 ^ ((uint32_t)VB_WISH_C << 24), pd.rk);
          const double u = u53(ph.x, ph.y);
          const double v1 = fma(cc, x, 1.0), v = v1 * v1 * v1, x2 = x * x;
          bool ok = false;
          if (v1 > 0.0 && d > 2.0 / 3.0) {                      // (shape >= 1: no boost step)
            ok = u < 1.0 - 0.0331 * x2 * x2;
            if (!ok) ok = bvg_logpos(u) < 0.5 * x2 + d - d * v + d * bvg_logpos(v);
          }
          double g = d * v;
          if (!ok) g = philox_gamma_from_x0(rp.seed, chain, VB_WISH_C, sweep, i, 0.5 * (m.wish_df_post - (double)i), x);
          rec[o_chi + i] = bvg_sqrtpos(2.0 * g);
        }
        continue;
      }
      // ---- two Box-Muller slots of the record
      const int sa = 2 * sp, sb = (2 * sp + 1 < NSL) ? 2 * sp + 1 : 2 * sp;
      const int4 da = slot_of(sa), db = slot_of(sb);
      if (inj_coef || inj_wn) {
        auto one = [&](const int4& d) {
          const int blk = d.y >> 24, cnt = (blk == VB_COEF_N) ? n : NW;
          if (rp.inj[blk] != nullptr) {
            const double* src = rp.inj[blk] + (size_t)(chain0 + c) * (size_t)rp.inj_chain_stride[blk] +
                                (size_t)sweep * (size_t)rp.inj_stride[blk];
            rec[d.z] = src[2 * d.x];
            rec[d.w] = (2 * d.x + 1 < cnt) ? src[2 * d.x + 1] : 0.0;
          } else {
            const Philox4 ph = philox_rk((uint32_t)d.x, (uint32_t)sweep, c2, c3 ^ (uint32_t)d.y, pd.rk);
            bvg_box_muller(u53(ph.x, ph.y), u53(ph.z, ph.w), &rec[d.z], &rec[d.w]);
          }
        };
        one(da);
        one(db);
        continue;
      }
      const Philox4 pa = philox_rk((uint32_t)da.x, (uint32_t)sweep, c2, c3 ^ (uint32_t)da.y, pd.rk);
      const Philox4 pb = philox_rk((uint32_t)db.x, (uint32_t)sweep, c2, c3 ^ (uint32_t)db.y, pd.rk);
      double a0, a1, b0v, b1v;
      bvg_box_muller(u53(pa.x, pa.y), u53(pa.z, pa.w), &a0, &a1);
      bvg_box_muller(u53(pb.x, pb.y), u53(pb.z, pb.w), &b0v, &b1v);
      rec[da.z] = a0;
      rec[da.w] = a1;
      rec[db.z] = b0v;
      rec[db.w] = b1v;
    }
  };
  // ---- 2a: state-independent products of record `it` of buffer `buf` (round of nb sweeps)
  auto prep = [&](int it, const FastDiv& dnb, int buf) {
    unsigned c, sw;
    dnb.divmod((unsigned)it, c, sw);
    double* v = var + (size_t)buf * CH * CS + c * (unsigned)CS + sw * (unsigned)VL;
    double chi[K], wn[NW > 0 ? NW : 1];
#pragma unroll
    for (int i = 0; i < K; ++i) chi[i] = v[o_chi + i];
#pragma unroll
    for (int i = 0; i < NW; ++i) wn[i] = v[o_wn + i];
    if (K * M <= 40 && M == 2 * K + 1) kron_prep_regs<K, 2 * K + 1>(v, wn, chi, cs.Linv, v + o_G, v + o_Ai);
    else if (K * M <= 40 && M == K + 1) kron_prep_regs<K, K + 1>(v, wn, chi, cs.Linv, v + o_G, v + o_Ai);
    else if (K * M <= 40 && M == 4 * K + 1) kron_prep_regs<K, (K * (4 * K + 1) <= 40) ? 4 * K + 1 : 1>(v, wn, chi, cs.Linv, v + o_G, v + o_Ai);
    else if (K * M <= 40 && M == 3 * K + 1) kron_prep_regs<K, (K * (3 * K + 1) <= 40) ? 3 * K + 1 : 1>(v, wn, chi, cs.Linv, v + o_G, v + o_Ai);
    else {
      kron_prep<K>(v, wn, chi, M, v + o_G, v + o_Ai);
      kron_t_inplace<K>(v, cs.Linv, M);
    }
  };

  if (tid == 0) { work_ctr[0] = 0; flags[0] = 0; flags[1] = 0; }
  __syncthreads();
  // ---- roles: warp 0 runs the state recursion; the workers generate, prepare and write; with pd.skip0 the other warps of
  //      warp 0's scheduler (warp id % 4 == 0) stay out of it, so that the latency-bound recursion has its issue slots
  const int warp = tid >> 5, n_warps = NT >> 5;
  const int n_workers = (n_warps - 1) - (pd.skip0 ? (n_warps - 1) / 4 : 0);
  const int wi = warp - 1 - (pd.skip0 ? warp / 4 : 0);        // worker rank of this warp
  const int WT = n_workers * 32, wt = wi * 32 + lane;
  const int n_rounds = (s1 - s0 + NB - 1) / NB;
  auto round_b0 = [&](int q) { return s0 + q * NB; };
  auto round_nb = [&](int q) { return (q < 0 || q >= n_rounds) ? 0 : ((s1 - round_b0(q) < NB) ? (s1 - round_b0(q)) : NB); };

  if (warp == 0) {
    // ================= state recursion: round q as soon as its records are prepared =================
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
    for (int q = 0; q < n_rounds; ++q) {
      if (lane == 0)
        while (*(volatile int*)&flags[0] < q + 1) __nanosleep(20);
      __syncwarp();
      __threadfence_block();
      const int nb = round_nb(q);
      if (tid < CHv) {
        double* rec = var + (size_t)(q % 3) * CH * CS + tid * CS;
        double Gp[KU], Aip[KU];
#pragma unroll
        for (int e = 0; e < KU; ++e) { Gp[e] = rec[o_G + e]; Aip[e] = rec[o_Ai + e]; }
        for (int sw = 0; sw < nb; ++sw, rec += VL) {
#pragma unroll
          for (int j = 0; j < K; ++j)
#pragma unroll
            for (int i = 0; i <= j; ++i) rec[o_Ai + tri(j) + i] = W[i][j];          // W used by this sweep
          double Gn[KU], Ain[KU];                                                 // next sweep's inputs: loads in flight
          const double* nx = rec + ((sw + 1 < nb) ? VL : 0);                      //   during this sweep's dependent chain
#pragma unroll
          for (int e = 0; e < KU; ++e) { Gn[e] = nx[o_G + e]; Ain[e] = nx[o_Ai + e]; }
          kron_step_core<K>(W, Gp, Aip, SSu, fail);
#pragma unroll
          for (int j = 0; j < K; ++j)
#pragma unroll
            for (int i = 0; i <= j; ++i) rec[o_G + tri(j) + i] = W[i][j];           // W after this sweep
#pragma unroll
          for (int e = 0; e < KU; ++e) { Gp[e] = Gn[e]; Aip[e] = Ain[e]; }
        }
      }
      __threadfence_block();
      __syncwarp();
      if (lane == 0) *(volatile int*)&flags[1] = q + 1;
    }
    if (tid < CHv) {
      double* st = state + (size_t)(chain0 + tid) * state_len;
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) st[i + K * j] = (i <= j) ? W[i][j] : 0.0;
      if (fail) out.status[chain0 + tid] = 1;
    }
    return;
  }
  if (wi < 0 || (pd.skip0 && (warp & 3) == 0)) return;         // (idle warps of the recursion warp's scheduler)

  // ================= workers: iteration q generates and prepares round q + 2 and writes round q =================
  bool bulk_pending = false, bulk_any = false;
  for (int q = -2; q < n_rounds; ++q) {
    const int r2 = q + 2, nb2 = round_nb(r2);
    if (nb2 > 0) gen_variates(round_b0(r2), nb2, r2 % 3);
    if (bulk_pending) {                                        // the staging buffer is rewritten after the barrier
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      bulk_pending = false;
    }
    asm volatile("bar.sync 1, %0;" ::"r"(WT) : "memory");
    if (wt == 0) work_ctr[0] = 0;
    // ---- 2a: state-independent products of round q + 2
    {
      const int n_prep = CHv * nb2;
      const FastDiv d2 = div_of(nb2);
      for (int it = wt; it < n_prep; it += WT) prep(it, d2, r2 % 3);
    }
    // ---- 2b: draws of round q (its recursion has to be through -- also before round q + 3 may overwrite the records)
    const int b0 = round_b0(q), nb = round_nb(q), b1 = b0 + nb;
    const int k0 = (b0 > burnin) ? b0 : burnin;                // first retained sweep of the round
    const int nk = (q >= 0 && b1 > k0) ? b1 - k0 : 0, sw0 = k0 - b0;
    if (q >= 0) {
      if (lane == 0)
        while (*(volatile int*)&flags[1] < q + 1) __nanosleep(64);
      __syncwarp();
      __threadfence_block();
    }
    if (nk > 0) {
      const int n_out = CHv * nk;
      const FastDiv dnk = div_of(nk);
      const double* vb = var + (size_t)(q % 3) * CH * CS;
      for (int it = wt; it < n_out; it += WT) {
        unsigned c, swu;
        dnk.divmod((unsigned)it, c, swu);
        const double* v = vb + c * (unsigned)CS + (sw0 + swu) * (unsigned)VL;
        double* dc = st_coef + c * (unsigned)pd.SC + swu * (unsigned)n;
        if (M == 2 * K + 1) kron_coef_to<K, 2 * K + 1>(v + o_Ai, v, cs.Aols, M, dc);
        else if (M == K + 1) kron_coef_to<K, K + 1>(v + o_Ai, v, cs.Aols, M, dc);
        else kron_coef_to<K, 0>(v + o_Ai, v, cs.Aols, M, dc);
        kron_sigma<K>(v + o_G, st_sig + c * (unsigned)pd.SG + swu * (unsigned)KK);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged draws -> visible to the copy engine
    }
    asm volatile("bar.sync 1, %0;" ::"r"(WT) : "memory");
    if (wt == 0 && nb2 > 0) {
      __threadfence_block();
      *(volatile int*)&flags[0] = r2 + 1;                      // rounds <= q + 2 are prepared
    }
    // ---- 3: staged draws to the [chain][iter][row] output
    if (nk > 0) {
      const long long pos0 = (long long)k0 - burnin;
      // 16-byte alignment of every chain's destination and size (chains differ by iters * n doubles)
      const bool al_c = (((size_t)out.iters * (size_t)n) % 2 == 0 || CHv == 1) && ((nk * n) % 2 == 0) &&
                        ((reinterpret_cast<uintptr_t>(out.coef + ((size_t)chain0 * (size_t)out.iters + (size_t)pos0) * (size_t)n)) % 16 == 0);
      const bool al_s = (((size_t)out.iters * (size_t)KK) % 2 == 0 || CHv == 1) && ((nk * KK) % 2 == 0) &&
                        ((reinterpret_cast<uintptr_t>(out.sigma + ((size_t)chain0 * (size_t)out.iters + (size_t)pos0) * (size_t)KK)) % 16 == 0);
      for (int c = wi; c < CHv; c += n_workers) {              // one warp per chain: uniform addresses for the bulk copies
        double* dstc = out.coef + ((size_t)(chain0 + c) * (size_t)out.iters + (size_t)pos0) * (size_t)n;
        double* dsts = out.sigma + ((size_t)(chain0 + c) * (size_t)out.iters + (size_t)pos0) * (size_t)KK;
        const double* srcc = st_coef + c * pd.SC;
        const double* srcs = st_sig + c * pd.SG;
        if (lane == 0 && (al_c || al_s)) {
          if (al_c) bulk_store(dstc, srcc, (unsigned)(nk * n * 8));
          if (al_s) bulk_store(dsts, srcs, (unsigned)(nk * KK * 8));
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          bulk_pending = bulk_any = true;
        }
        if (!al_c)
          for (int e = lane; e < nk * n; e += 32) dstc[e] = srcc[e];
        if (!al_s)
          for (int e = lane; e < nk * KK; e += 32) dsts[e] = srcs[e];
      }
    }
  }
  if (bulk_any) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
  const int len = 2 * ((n + 1) / 2) + 2 * ((NW + 1) / 2) + 2 * KU;
  pd->CH = CH;
  pd->NB = NB;
  pd->skip0 = 0;
  pd->VL = len | 1;                         // odd strides: the lanes of a warp (consecutive records / chains) hit
  pd->CS = (NB * pd->VL) | 1;               //   different banks
  pd->SC = (NB * n + 1) & ~1;               // even: 16-byte aligned bulk-copy sources
  pd->SG = (NB * k * k + 1) & ~1;
}
static size_t kron_smem_bytes(int k, int M, const KronPlanDims& pd) {
  return ((size_t)pd.CH * (pd.SC + pd.SG) + 3 * (size_t)pd.CH * pd.CS + (size_t)kron_consts_len(k, M) + 2 * k) * 8;
}

cudaError_t launch_bvar_kron(BVG_BVAR_ARGS_NAMED) {
  KronPlanDims pd;
  kron_dims(m.k, m.M, plan.chains_per_block, plan.kron_nb & 0xffff, &pd);
  pd.skip0 = (plan.kron_nb >> 16) & 1;
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

// Launch geometry.  The kernel runs for the whole launch with a fixed set of CTAs, so what matters is (i) an even
// fill of the SMs -- r CTAs on every SM, CH = ceil(chains / (SMs r)) <= 32 chains each -- and (ii) enough warps per SM
// that the variate phase hides the latency of the state recursion.  threads_per_chain (tuning knob) = CTA threads per
// chain, 0 = auto.
cudaError_t plan_bvar_kron(const BvarModel& m, long long n_chains, int requested_tpc, LaunchPlan* plan) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int CH = 32, NT = 512, NB = 0, skip0 = 0;
  if (requested_tpc > 0) {
    // tuning / tests: 4..15 -> 128-thread CTAs, 16..128 -> 512-thread CTAs; >= 1000: explicit geometry
    // 10000000 * skip0 + 100000 * NB + 1000 * (NT / 128) + CH, e.g. 604028 = 6 sweeps per round, 512 threads, 28 chains
    int tpc = requested_tpc < 4 ? 4 : requested_tpc;
    if (tpc >= 1000) {
      skip0 = tpc / 10000000; tpc %= 10000000;
      NB = tpc / 100000; tpc %= 100000;
      NT = 128 * (tpc / 1000); CH = tpc % 1000;
      if (NT > 512) NT = 512;
      if (NT < 64) NT = 64;
    }
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
  }
  if (skip0 && NT < 256) skip0 = 0;
  const int k = m.k, M = m.M;
  KronPlanDims pd;
  auto bytes = [&](int nb) {
    kron_dims(k, M, CH, nb, &pd);
    return kron_smem_bytes(k, M, pd);
  };
  if (NB <= 0) {
    NB = 8;
    while (NB < 64 && CH * NB < 224) NB <<= 1;    // few chains per CTA: longer rounds keep >= 224 records per round
    while (NB > 1 && bytes(NB) > (size_t)216 * 1024 / (512 / NT)) --NB;
  }
  if (bytes(NB) > 227 * 1024 - 64) return cudaErrorInvalidConfiguration;
  plan->kron = 1;
  plan->kron_nb = NB | (skip0 << 16);
  plan->threads_per_chain = NT / CH;
  plan->smem_per_chain = pd.VL;
  plan->block_threads = NT;
  plan->chains_per_block = CH;
  plan->smem_bytes = bytes(NB);
  plan->grid = (int)((n_chains + CH - 1) / CH);
  return cudaSuccess;
}

}  // namespace bvg
