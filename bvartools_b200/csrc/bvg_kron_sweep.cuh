// bvg_kron_sweep.cuh -- flat-prior specialisation of the constant-parameter BVAR sweep (configs c1 / c2: README.md:147-149
// and vignettes/model-comparison.Rmd:58-60 both use v_i = 0).
//
// With V^-1 = 0 the posterior precision of bvaralg.cpp:305 is exactly Kronecker, P = (X X') (x) Sigma^-1, and so is its
// (unique) upper Cholesky factor: chol(P) = R_X (x) R_S with X X' = R_X' R_X and Sigma^-1 = R_S' R_S.  The reference's
// draw  a = P^-1 b + chol(P)^-1 eps  (bvaralg.cpp:306-307) therefore is, exactly,
//     A = A_ols + R_S^-1 E L_X^-1 ,        E = reshape(eps, k, M),   L_X = R_X' (lower factor of X X', constant)
// and the residual cross product needed by the Wishart block (bvaralg.cpp:364,511) collapses to
//     u u' = SSE_ols + (A - A_ols) X X' (A - A_ols)' = SSE_ols + R_S^-1 (E E') R_S^-T .
// The inverse-Wishart block (bvaralg.cpp:511, arma::wishrnd) draws Sigma^-1 = (A_B D)'(A_B D) with A_B the Bartlett
// factor (upper) and D = chol(S^-1) (upper), S = S0 + u u'.  Writing S = U U' with U UPPER (the Cholesky factorisation
// taken from the last row up), S^-1 = U^-T U^-1 and by uniqueness D = U^-1, so the new R_S = A_B U^-1 and
//     R_S^-1 = U A_B^-1 ,        Sigma = R_S^-1 R_S^-T                                      (bvaralg.cpp:528)
// -- no inverse of S and no second factorisation.  The chain state is W = R_S^-1 (k x k upper).  Same variates and the
// same mapping eps -> a as the reference (tier-1 parity to rounding).  Everything random in the sweep (eps, A_B) is
// independent of the chain state, so the kernels generate variates for a batch of sweeps in a throughput-bound phase
// and run the short state recursion  W <- U(S0 + SSE + W E E' W') A_B^-1  separately (bvg_kern_kron.cu).
// The general kernel (bvg_bvar_sweep.cuh) covers every other prior.
#pragma once
#include "bvg_bvar_sweep.cuh"

namespace bvg {

// constants of the model shared by all chains (staged in shared memory by the kernels)
struct KronConsts {
  const double* Linv;   // tri(M) packed lower row-major: L_X^-1
  const double* Aols;   // k x M col-major
  const double* SS;     // k x k col-major: S0 + SSE_ols
};
BVG_HD int kron_consts_len(int k, int M) { return tri(M) + k * M + k * k; }
// record of one sweep of one chain: eps (n) | Bartlett normals (k(k-1)/2) | sqrt(chi2) (k) | G = E E' (k(k+1)/2) |
// A_B^-1 (k(k+1)/2)
BVG_HD int kron_variates_len(int k, int M) { return k * M + (k * (k - 1)) / 2 + k + k * (k + 1); }
BVG_HD int kron_smem_len(int k, int M) { return kron_variates_len(k, M) + 1; }
BVG_HD int kron_pairs(int k, int M) { return ((k * M + 1) >> 1) + (((k * (k - 1)) / 2 + 1) >> 1); }

// Box-Muller pair `p` of the sweep's state-independent normals: coefficient normals first, then the Bartlett ones.
// v = variates of (chain, sweep): [eps (n) | wn (NW) | chi (K)]
BVG_HD void kron_gen_pair(const ChainRng& rng, int sweep, int p, int K, int M, double* v) {
  const int n = K * M, NW = (K * (K - 1)) / 2, p_coef = (n + 1) >> 1;
  int block = VB_COEF_N, pr = p, cnt = n;
  double* dst = v;
  if (p >= p_coef) { block = VB_WISH_N; pr = p - p_coef; cnt = NW; dst = v + n; }
  const Normal2 nn = rng.normal_pair(block, sweep, pr, cnt);
  dst[2 * pr] = nn.a;
  if (2 * pr + 1 < cnt) dst[2 * pr + 1] = nn.b;
}
// two pairs (of possibly different chains / sweeps) at once: see philox_normal_pair_x2
BVG_HD void kron_gen_pair_x2(const ChainRng& ra, int sweep_a, int pa, double* va, const ChainRng& rb, int sweep_b, int pb,
                             double* vb, int K, int M) {
  if (ra.injected(VB_COEF_N) || ra.injected(VB_WISH_N)) {
    kron_gen_pair(ra, sweep_a, pa, K, M, va);
    kron_gen_pair(rb, sweep_b, pb, K, M, vb);
    return;
  }
  const int n = K * M, NW = (K * (K - 1)) / 2, p_coef = (n + 1) >> 1;
  int blk_a = VB_COEF_N, pra = pa, cnt_a = n, blk_b = VB_COEF_N, prb = pb, cnt_b = n;
  double* da = va;
  double* db = vb;
  if (pa >= p_coef) { blk_a = VB_WISH_N; pra = pa - p_coef; cnt_a = NW; da = va + n; }
  if (pb >= p_coef) { blk_b = VB_WISH_N; prb = pb - p_coef; cnt_b = NW; db = vb + n; }
  const Normal4 nn = philox_normal_pair_x2(ra.seed, ra.chain, blk_a, sweep_a, pra, rb.chain, blk_b, sweep_b, prb);
  da[2 * pra] = nn.p.a;
  if (2 * pra + 1 < cnt_a) da[2 * pra + 1] = nn.p.b;
  db[2 * prb] = nn.q.a;
  if (2 * prb + 1 < cnt_b) db[2 * prb + 1] = nn.q.b;
}
// sqrt(chi2(df - i)) of the Bartlett diagonal (Marsaglia-Tsang; attempt 0 uses normal(VB_WISH_C, sweep, i))
BVG_HD double kron_gen_chi(const ChainRng& rng, int sweep, int i, double df_post) {
  if (rng.injected(VB_WISH_C)) return sqrt(rng.inj_at(VB_WISH_C, sweep, i));
  return sqrt(rng.chi2_from_x0(VB_WISH_C, sweep, i, df_post - (double)i, rng.normal(VB_WISH_C, sweep, i)));
}

// State-independent preparation of one sweep (flat, parallel work): G = E E' (lower packed, G_ij at tri(i) + j) and
// Ai = A_B^-1 (upper packed by column, Ai_ij at tri(j) + i) from the sweep's variates.
// eps: E col-major (i + K j), wn: strict-upper Bartlett normals column by column, chi: sqrt(chi2).
template <int K>
BVG_HD void kron_prep(const double* eps, const double* wn, const double* chi, int M, double* Gp, double* Aip) {
  double Gm[K][K];
BVG_UNROLL
  for (int i = 0; i < K; ++i)
BVG_UNROLL
    for (int j = 0; j <= i; ++j) Gm[i][j] = 0.0;
  for (int j2 = 0; j2 < M; ++j2) {
    double e[K];
BVG_UNROLL
    for (int i = 0; i < K; ++i) e[i] = eps[i + K * j2];
BVG_UNROLL
    for (int i = 0; i < K; ++i)
BVG_UNROLL
      for (int j = 0; j <= i; ++j) Gm[i][j] = fma(e[i], e[j], Gm[i][j]);
  }
  double Ai[K][K];
BVG_UNROLL
  for (int c = 0; c < K; ++c) {
    Ai[c][c] = 1.0 / chi[c];
BVG_UNROLL
    for (int r = c - 1; r >= 0; --r) {
      double v = 0.0;
BVG_UNROLL
      for (int mm = r + 1; mm <= c; ++mm) v = fma(wn[tri(mm - 1) + r], Ai[mm][c], v);
      Ai[r][c] = -v * Ai[r][r];
    }
  }
BVG_UNROLL
  for (int i = 0; i < K; ++i)
BVG_UNROLL
    for (int j = 0; j <= i; ++j) { Gp[tri(i) + j] = Gm[i][j]; Aip[tri(i) + j] = Ai[j][i]; }
}

// S = U U' with U upper (the Cholesky factorisation taken from the last column to the first); only the upper triangle of S
// is read.  fail is set on a non-positive pivot.
template <int K>
BVG_HD void kron_upper_factor(const double (&S)[K][K], double (&U)[K][K], int& fail) {
BVG_UNROLL
  for (int j = K - 1; j >= 0; --j) {
    double d = S[j][j];
BVG_UNROLL
    for (int c = j + 1; c < K; ++c) d = fma(-U[j][c], U[j][c], d);
    if (!(d > 0.0)) fail = 1;
    const double rs = bvg_rsqrt_fast(d);
    U[j][j] = d * rs;
BVG_UNROLL
    for (int i = j - 1; i >= 0; --i) {
      double v = S[i][j];
BVG_UNROLL
      for (int c = j + 1; c < K; ++c) v = fma(-U[i][c], U[j][c], v);
      U[i][j] = v * rs;
    }
  }
}
// k = 3 and k = 2: the pivots of the sequential factorisation are ratios of trailing minors, t_11 = n11 / s22,
// t_00 = det / n11, so the reciprocal square roots do not have to wait for one another: r22 = s22^-1/2, q11 = n11^-1/2,
// q00 = det^-1/2 start together and U follows by products.  The state recursion is the sequential part of the sweep: its
// latency, not its flop count, bounds launches with few chains per SM (c2-B, strong scaling).  The minors are formed with
// the same cancellation as the Cholesky pivots themselves (s11 - s12^2 / s22 against (s11 s22 - s12^2) / s22): same
// conditioning, results equal to rounding.
BVG_HD void kron_upper_factor(const double (&S)[3][3], double (&U)[3][3], int& fail) {
  const double s00 = S[0][0], s01 = S[0][1], s02 = S[0][2], s11 = S[1][1], s12 = S[1][2], s22 = S[2][2];
  const double n11 = fma(s11, s22, -(s12 * s12));
  const double c01 = fma(s01, s22, -(s02 * s12));
  const double c02 = fma(s01, s12, -(s02 * s11));
  const double det = fma(s02, c02, fma(s00, n11, -(s01 * c01)));
  if (!(s22 > 0.0) || !(n11 > 0.0) || !(det > 0.0)) fail = 1;
  const double r22 = bvg_rsqrt_fast(s22), q11 = bvg_rsqrt_fast(n11), q00 = bvg_rsqrt_fast(det);
  U[2][2] = s22 * r22; U[1][2] = s12 * r22; U[0][2] = s02 * r22;
  U[1][1] = (n11 * q11) * r22;
  U[0][1] = c01 * (q11 * r22);
  U[0][0] = (det * q00) * q11;
  U[1][0] = 0.0; U[2][0] = 0.0; U[2][1] = 0.0;
}
BVG_HD void kron_upper_factor(const double (&S)[2][2], double (&U)[2][2], int& fail) {
  const double s00 = S[0][0], s01 = S[0][1], s11 = S[1][1];
  const double n00 = fma(s00, s11, -(s01 * s01));
  if (!(s11 > 0.0) || !(n00 > 0.0)) fail = 1;
  const double r11 = bvg_rsqrt_fast(s11), q00 = bvg_rsqrt_fast(n00);
  U[1][1] = s11 * r11; U[0][1] = s01 * r11;
  U[0][0] = (n00 * q00) * r11;
  U[1][0] = 0.0;
}

// One step of the state recursion (the only state-dependent part of the sweep, a few dozen dependent FP64
// operations): W (upper, registers) <- U(SS + W G W') Ai, with S = U U', U upper.  SSu: upper triangle of S0 + SSE.
// fail is set on a non-positive pivot.
template <int K>
BVG_HD void kron_step_core(double (&W)[K][K], const double* Gp, const double* Aip, const double (&SSu)[K][K],
                           int& fail) {
  double Gm[K][K], Ai[K][K];
BVG_UNROLL
  for (int i = 0; i < K; ++i)
BVG_UNROLL
    for (int j = 0; j <= i; ++j) { Gm[i][j] = Gp[tri(i) + j]; Ai[j][i] = Aip[tri(i) + j]; }
  // S = SS + W G W'  (upper triangle: S[i][j], i <= j)
  double H[K][K], S[K][K];
BVG_UNROLL
  for (int i = 0; i < K; ++i)
BVG_UNROLL
    for (int j = 0; j < K; ++j) {
      double v = 0.0;
BVG_UNROLL
      for (int c = i; c < K; ++c) v = fma(W[i][c], (c >= j) ? Gm[c][j] : Gm[j][c], v);
      H[i][j] = v;
    }
BVG_UNROLL
  for (int i = 0; i < K; ++i)
BVG_UNROLL
    for (int j = i; j < K; ++j) {
      double v = SSu[i][j];
BVG_UNROLL
      for (int c = j; c < K; ++c) v = fma(H[i][c], W[j][c], v);
      S[i][j] = v;
    }
  // S = U U', U upper
  double U[K][K];
  kron_upper_factor(S, U, fail);
  // W = U A_B^-1
BVG_UNROLL
  for (int i = 0; i < K; ++i)
BVG_UNROLL
    for (int j = i; j < K; ++j) {
      double v = 0.0;
BVG_UNROLL
      for (int mm = i; mm <= j; ++mm) v = fma(U[i][mm], Ai[mm][j], v);
      W[i][j] = v;
    }
}

// T = E L_X^-1 in place (eps <- T, column by column: column j only needs the columns >= j)     [state independent]
template <int K>
BVG_HD void kron_t_inplace(double* eps, const double* Linv, int M) {
  for (int j = 0; j < M; ++j) {
    double t[K];
BVG_UNROLL
    for (int i = 0; i < K; ++i) t[i] = 0.0;
    for (int j2 = j; j2 < M; ++j2) {
      const double li = Linv[tri(j2) + j];
BVG_UNROLL
      for (int i = 0; i < K; ++i) t[i] = fma(eps[i + K * j2], li, t[i]);
    }
BVG_UNROLL
    for (int i = 0; i < K; ++i) eps[i + K * j] = t[i];
  }
}
// a = vec(A_ols + W T) in place (tvals <- a)                                          (bvaralg.cpp:306-307)
// Wp: the K(K+1)/2 upper entries of W packed column by column (W_ij at tri(j) + i)
template <int K>
BVG_HD void kron_coef_inplace(const double* Wp, double* tvals, const double* Aols, int M) {
  double W[K][K];
BVG_UNROLL
  for (int j = 0; j < K; ++j)
BVG_UNROLL
    for (int i = 0; i <= j; ++i) W[i][j] = Wp[tri(j) + i];
  for (int j = 0; j < M; ++j) {
    double t[K];
BVG_UNROLL
    for (int i = 0; i < K; ++i) t[i] = tvals[i + K * j];
BVG_UNROLL
    for (int i = 0; i < K; ++i) {
      double d = Aols[i + K * j];
BVG_UNROLL
      for (int i2 = i; i2 < K; ++i2) d = fma(W[i][i2], t[i2], d);
      tvals[i + K * j] = d;
    }
  }
}
// Sigma = W W' (full K x K, col-major) into dst                                        (bvaralg.cpp:528)
template <int K>
BVG_HD void kron_sigma(const double* Wp, double* dst) {
  double W[K][K];
BVG_UNROLL
  for (int j = 0; j < K; ++j)
BVG_UNROLL
    for (int i = 0; i <= j; ++i) W[i][j] = Wp[tri(j) + i];
BVG_UNROLL
  for (int i = 0; i < K; ++i)
BVG_UNROLL
    for (int j = 0; j <= i; ++j) {
      double v = 0.0;
BVG_UNROLL
      for (int mm = i; mm < K; ++mm) v = fma(W[i][mm], W[j][mm], v);
      dst[i + K * j] = v;
      dst[j + K * i] = v;
    }
}

// ---- group-generic chain (tile kernel for few sweeps per launch, and the host emulation of the arithmetic) --------
// Chain state between launches: W (k x k upper, col-major) in the Sigma^-1 slot of the generic state layout.
template <int K, class Grp>
BVG_HD void bvar_kron_chain(const Grp& g, const BvarModel& m, const KronConsts& cs, const ChainRng& rng,
                            double* state, double* sm, const BvarOut& out, long long chain_local, int sweep_begin,
                            int sweep_end, int burnin) {
  const int G = g.size(), tid = g.tid();
  const int M = m.M, n = K * M;
  constexpr int NW = (K * (K - 1)) / 2, KU = (K * (K + 1)) / 2;
  double* v = sm;                  // [eps | wn | chi]
  int fail = 0;
  double W[K][K];
BVG_UNROLL
  for (int i = 0; i < K; ++i)
BVG_UNROLL
    for (int j = 0; j < K; ++j) W[i][j] = (i <= j) ? state[i + K * j] : 0.0;
  const int n_pairs = kron_pairs(K, M);

  for (int sweep = sweep_begin; sweep < sweep_end; ++sweep) {
    const bool keep = sweep >= burnin;
    const long long pos = (long long)sweep - burnin;
    for (int it = tid; it < n_pairs + K; it += G) {
      if (it < n_pairs) kron_gen_pair(rng, sweep, it, K, M, v);
      else v[n + NW + (it - n_pairs)] = kron_gen_chi(rng, sweep, it - n_pairs, m.wish_df_post);
    }
    g.sync();
    double Wp[KU];
BVG_UNROLL
    for (int j = 0; j < K; ++j)
BVG_UNROLL
      for (int i = 0; i <= j; ++i) Wp[tri(j) + i] = W[i][j];
    double SSu[K][K];
BVG_UNROLL
    for (int i = 0; i < K; ++i)
BVG_UNROLL
      for (int j = 0; j < K; ++j) SSu[i][j] = cs.SS[i + K * j];
    double* Gp = v + n + NW + K;
    double* Aip = Gp + KU;
    if (tid == 0) {
      kron_prep<K>(v, v + n, v + n + NW, M, Gp, Aip);
      kron_t_inplace<K>(v, cs.Linv, M);
      if (keep) kron_coef_inplace<K>(Wp, v, cs.Aols, M);
    }
    g.sync();
    if (keep) {
      double* oc = out.coef + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)n;
      for (int e = tid; e < n; e += G) oc[e] = v[e];
    }
    kron_step_core<K>(W, Gp, Aip, SSu, fail);       // every lane, redundantly, on registers
    g.sync();
    if (keep) {
BVG_UNROLL
      for (int j = 0; j < K; ++j)
BVG_UNROLL
        for (int i = 0; i <= j; ++i) Wp[tri(j) + i] = W[i][j];
      if (tid == 0) kron_sigma<K>(Wp, Gp);          // staged in the G / A_B^-1 slots (no longer needed)
      g.sync();
      double* os = out.sigma + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)(K * K);
      for (int e = tid; e < K * K; e += G) os[e] = Gp[e];
    }
    g.sync();     // the variates are rewritten by the next sweep
  }
  if (tid == 0) {
BVG_UNROLL
    for (int i = 0; i < K; ++i)
BVG_UNROLL
      for (int j = 0; j < K; ++j) state[i + K * j] = (i <= j) ? W[i][j] : 0.0;
    if (fail) out.status[chain_local] = 1;
  }
  g.sync();
}

}  // namespace bvg
