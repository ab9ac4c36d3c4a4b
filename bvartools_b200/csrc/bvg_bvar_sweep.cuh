// bvg_bvar_sweep.cuh -- one chain of the constant-parameter BVAR Gibbs sampler, all sweeps on-chip.
//
// Restates the sweep of the reference's monolithic loop (src/bvaralg.cpp:292-560) for pure VAR regressors
// (z = kron(x_t', I_k), R/gen_var.R:281), exploiting the structure the reference materialises densely:
//   z' D z = (X X') (x) Sigma^-1               (const Sigma)     bvaralg.cpp:305   [post_normal.cpp:86 identity]
//          = interleave_i( sum_t e^{-h_ti} x_t x_t' )   (SV, diagonal Sigma_t)     bvaralg.cpp:494,305
//   z' D y = vec(Sigma^-1 Y X')                                    bvaralg.cpp:306
//   a = P^-1 b + chol(P)^-1 eps  ==  L^-T (L^-1 b + eps),  P = L L'                bvaralg.cpp:306-307
//   u u'  = SSE_ols + (A - A_ols) XX' (A - A_ols)'   (exact; "moment" residual mode)   bvaralg.cpp:364,511
// Sweep order is the reference's: coefficients (with last sweep's Sigma^-1) -> SSVS | BVS -> residuals ->
// SV | gamma | Wishart -> store.
#pragma once
#include "bvg_linalg.cuh"
#include "bvg_tiled.cuh"

namespace bvg {

enum SigmaType : int { SIG_WISHART = 0, SIG_GAMMA = 1, SIG_SV = 2 };
enum VarSel : int { VS_NONE = 0, VS_SSVS = 1, VS_BVS = 2 };
enum ResidMode : int { RES_DIRECT = 1, RES_MOMENT = 2 };
// compile-time feature mask of a kernel variant: code of features not in the mask is compiled out, which keeps
// the headline kernel (Wishart prior, moment residuals, no variable selection) small enough for the I-cache
enum Feat : int { F_SV = 1, F_VARSEL = 2, F_DIRECT = 4, F_GAMMA = 8, F_VIOFF = 16, F_ALL = 31 };
BVG_HD int bvar_features_needed(int sigma_type, int varsel, int resid_mode, bool has_vi_off) {
  int need = 0;
  if (sigma_type == SIG_SV) need |= F_SV | F_DIRECT;
  if (sigma_type == SIG_GAMMA) need |= F_GAMMA;
  if (varsel != VS_NONE) need |= F_VARSEL;
  if (resid_mode == RES_DIRECT) need |= F_DIRECT;
  if (has_vi_off) need |= F_VIOFF;
  return need;
}

// Model constants, shared by all chains of one model.  All pointers are device (or host, in tests/emu) memory.
struct BvarModel {
  int k, T, M, n;
  int sigma_type, varsel, resid_mode;
  int n_mix;                       // 7 (KSC-1998) or 10 (OCSN-2007)
  const double* Y;                 // k x T  col-major: y_it at i + k t            (bvaralg.cpp:10)
  const double* X;                 // M x T  col-major: x_jt at j + M t            (Z' of gen_var)
  const double* XX;                // M x M  X X'
  const double* YX;                // k x M  Y X'   col-major (i + k j)
  const double* Aols;              // k x M  OLS coefficients (moment mode)
  const double* SSE;               // k x k  OLS residual cross product (moment mode)
  const double* mu;                // n      prior mean
  const double* vi_off;            // n x n packed lower, off-diagonal part of V^-1 (diag slots zero) or nullptr
  const double* vmu_off;           // n      vi_off * mu (or nullptr)
  const double* vdiag0;            // n      initial diagonal of V^-1
  // SSVS / BVS
  const double* tau0; const double* tau1; const double* inprior;     // n each
  const unsigned char* include;    // n  (1 = position takes part in the selection)
  // error variance prior
  double wish_df_post;             // df + T                                         (bvaralg.cpp:234)
  const double* S0;                // k x k                                          (bvaralg.cpp:235)
  const double* g_shape_post;      // k   shape + T/2  (gamma and SV)                (bvaralg.cpp:223,239)
  const double* g_rate;            // k
  const double* sv_mu;             // k   prior mean of h_init
  const double* sv_vi;             // k x k prior precision of h_init (col-major)
  const double* sv_const;          // k   offset constant c                          (bvaralg.cpp:229)
  const double* mix_p; const double* mix_mu; const double* mix_s2;   // n_mix each
  const double* omega_off;         // k x k off-diagonal part of the initial Sigma^-1 (gamma prior), zero diagonal
  // covariance (psi) block (bvaralg.cpp:373-455); n_psi = 0: none
  int diag_sigma;                  // TVP: the measurement precision stays diagonal (SV, or gamma variances with a diagonal
                                   // initial value, no covariance block): equation-decoupled state draw (bvg_tvp_eq8.cuh)
  int a0_from;                     // structural models: first internal row of the -y block (k M_user); -1 otherwise
  const double* maskd;             // structural models with SSVS: n doubles, 1 = kept position, 0 = masked; nullptr otherwise
                                   // (without selection / with BVS the mask rides on lambda)
  int n_psi;                       // k (k - 1) / 2
  const double* psi_vi;            // n_psi x n_psi prior precision (col-major); its diagonal lives in the chain state
  const double* psi_vmu;           // n_psi  OFF-diagonal part of (prior precision times prior mean)
  const double* psi_mu;            // n_psi  prior mean
  int psi_varsel;                  // VS_* on the covariances (bvaralg.cpp:396-449)
  const double* psi_tau0; const double* psi_tau1; const double* psi_inprior;      // n_psi each
  const unsigned char* psi_include;                                               // n_psi
  // flat-prior Kronecker specialisation (bvg_kron_sweep.cuh); nullptr when the model does not qualify
  const double* kron_Linv;         // tri(M) packed lower: inverse of the lower Cholesky factor of X X'
  const double* kron_SS;           // k x k  S0 + SSE_ols
};

// The small read-only tables every sweep touches.  The CUDA kernels stage them once per CTA into shared memory
// (LDS instead of L1-cached global loads on the hot path); the host emulation points them at the model's arrays.
struct BvarConsts {
  const double* XX; const double* YX; const double* Aols; const double* SSE; const double* S0;
  const double* mu; const double* vdiag0;
};
// lean (CTA-per-chain kernels): mu and vdiag0 stay in global memory (read once per sweep and coefficient)
BVG_HD int bvar_consts_len(int k, int M, bool lean = false) { return M * M + 2 * k * M + 2 * k * k + (lean ? 0 : 2 * k * M); }
BVG_HD BvarConsts bvar_consts_of(const BvarModel& m) {
  BvarConsts c;
  c.XX = m.XX; c.YX = m.YX; c.Aols = m.Aols; c.SSE = m.SSE; c.S0 = m.S0; c.mu = m.mu; c.vdiag0 = m.vdiag0;
  return c;
}

// Per-chain persistent state in global memory (between kernel launches), doubles:
//   [0, k*k)            Sigma^-1 (col-major)           (const Sigma)          | unused for SV
//   then n              vdiag   (diagonal of V^-1, changed by SSVS)
//   then n              lambda  (0/1)
//   SV only: then T*k   h (col-major T x k: h_ti at t + T i), then k sigma_h, then k h_init
//   covariance block only: then k  diagonal of Omega^-1 (gamma prior), n_psi  psi (strict lower triangle of Psi by rows),
//                          n_psi  diagonal of the psi prior precision (SSVS), n_psi  psi inclusion indicators
BVG_HD int bvar_state_len(int k, int T, int n, int sigma_type, int n_psi = 0) {
  return k * k + 2 * n + (sigma_type == SIG_SV ? T * k + 2 * k : 0) + (n_psi > 0 ? k + 3 * n_psi : 0);
}

// Shared-memory doubles needed per chain.
BVG_HD int bvar_smem_len(int k, int T, int M, int n, int sigma_type, int resid_mode, bool tiled = false, int n_psi = 0) {
  int len = (tiled ? tiled_len(n) + 8 - n : tri(n)) + 5 * n   // L, w (padded when tiled), a, dinv (not tiled), vdiag, lam
            + 6 * k * k + 2 * k      // sig, S, 4 k x k work, kd, kv
            + 2 * k * M;             // Dm, Wm
  if (resid_mode == RES_DIRECT || sigma_type == SIG_SV) len += k * T;                 // U
  if (sigma_type == SIG_SV) len += 3 * k * T + k * tri(M) + 2 * k;                    // h, ew, sv work, GW, sigma_h, h_init
  if (n_psi > 0) len += tri(n_psi) + 6 * n_psi + k + k * k;   // psi system, rhs, psi, dinv, prior diag, lambda, work, omega diag, Psi
  return len;
}

struct BvarOut {
  double* coef;     // [chain][iter][n]
  double* sigma;    // [chain][iter][k*k]  or  [chain][iter][T][k*k] (SV)
  double* lambda;   // [chain][iter][n] or nullptr
  double* sigma_h;  // [chain][iter][k]  (SV) or nullptr
  int* status;      // [chain]
  long long iters;  // retained iterations (stride computation)
};

// ---- small k x k helpers (full col-major storage, group-parallel) -----------------------------------------
// lower Cholesky of a full symmetric k x k matrix A into packed Lp (+dinv), then true diagonal written.
template <class Grp>
BVG_HD void small_chol(const Grp& g, const double* A, double* Lp, double* dinv, int k, int& fail) {
  for (int e = g.tid(); e < tri(k); e += g.size()) {
    int i = 0; while (tri(i + 1) <= e) ++i;
    const int j = e - tri(i);
    Lp[e] = A[i + k * j];
  }
  g.sync();
  chol_crout_generic(g, Lp, dinv, (double*)nullptr, k, fail);   // k x k: keep the unrolled variant for the n x n factor
}

// The inverse-Wishart block (bvaralg.cpp:511-514,528): given S = S0 + u u' (full, col-major) draw
//   Sigma^-1 = (A D)' (A D),  D = chol(S^-1) upper,  A = Bartlett factor  (arma::wishrnd convention)
// and return Sigma = (Sigma^-1)^-1 when `want_sigma`.  Work arrays: W1..W4 (k*k each), kd (k).
template <class Grp>
BVG_HD void wishart_block(const Grp& g, const ChainRng& rng, int sweep, int k, double df_post, const double* S,
                          double* sig, double* sigma_out, bool want_sigma, double* W1, double* W2, double* W3,
                          double* W4, double* kd, int& fail) {
  const int G = g.size(), tid = g.tid();
  // 1. S = Ls Ls'   (W1 packed, kd)
  small_chol(g, S, W1, kd, k, fail);
  // 2. Minv = Ls^-1 (W2 packed);  Sbar = S^-1 = Minv' Minv (W3 packed lower)
  tri_inverse(g, W1, kd, W2, k);
  mtm_lower(g, W2, W3, k);
  // 3. Sbar = Ld Ld'  (in place in W3, kd);  D = Ld' upper
  chol_crout_generic(g, W3, kd, (double*)nullptr, k, fail);
  fix_diag(g, W3, kd, k);
  // 4. Bartlett factor A (upper, full col-major in W4): A_ii = sqrt(chi2(df - i)), strict upper N(0,1) filled
  //    column by column (column i holds i normals)
  for (int e = tid; e < k * k; e += G) {
    const int r = e % k, c = e / k;
    double v = 0.0;
    if (r == c) v = sqrt(rng.chi2(VB_WISH_C, sweep, r, df_post - (double)r));
    else if (r < c) v = rng.normal(VB_WISH_N, sweep, tri(c - 1) + r);
    W4[e] = v;
  }
  g.sync();
  // 5. B = A D (upper, full col-major in W1): B_rc = sum_{m=r..c} A_rm D_mc, D_mc = Ld_cm
  for (int e = tid; e < k * k; e += G) {
    const int r = e % k, c = e / k;
    double s = 0.0;
    if (r <= c) for (int m = r; m <= c; ++m) s = fma(W4[r + k * m], W3[tri(c) + m], s);
    W1[e] = s;
  }
  g.sync();
  // 6. Sigma^-1 = B' B
  for (int e = tid; e < k * k; e += G) {
    const int i = e % k, j = e / k;
    const int mm = i < j ? i : j;
    double s = 0.0;
    for (int m = 0; m <= mm; ++m) s = fma(W1[m + k * i], W1[m + k * j], s);
    sig[e] = s;
  }
  g.sync();
  if (want_sigma) {
    // Sigma = B^-1 B^-T.  Binv (upper) by columns: column c of B^-1 solves B x = e_c (back substitution).
    for (int c = tid; c < k; c += G) {
      for (int r = k - 1; r >= 0; --r) {
        double s = (r == c) ? 1.0 : 0.0;
        for (int m = r + 1; m <= c; ++m) s = fma(-W1[r + k * m], W2[m + k * c], s);
        W2[r + k * c] = (r <= c) ? s / W1[r + k * r] : 0.0;
      }
    }
    g.sync();
    for (int e = tid; e < k * k; e += G) {
      const int i = e % k, j = e / k;
      const int m0 = i > j ? i : j;
      double s = 0.0;
      for (int m = m0; m < k; ++m) s = fma(W2[i + k * m], W2[j + k * m], s);
      sigma_out[e] = s;
    }
    g.sync();
  }
}

// Register version of the same block for small k (K <= 6, compile time): every lane redundantly runs the whole
// K x K algebra on registers (no shared-memory traffic, no barriers); the variates were drawn before, in the
// sweep's single generator pass: wn[] = strict-upper normals, x0[] = attempt-0 normals of the chi-squares.
template <int K, class Grp>
BVG_HD void wishart_block_reg(const Grp& g, const ChainRng& rng, int sweep, double df_post, const double* S,
                              const double* wn, const double* x0, double* chis, double* sig, double* sigma_out,
                              int& fail) {
  const int tid = g.tid();
  for (int i = tid; i < K; i += g.size())
    chis[i] = sqrt(rng.chi2_from_x0(VB_WISH_C, sweep, i, df_post - (double)i, x0[i]));
  g.sync();
  double Ls[K][K], Mi[K][K], Sb[K][K], Ld[K][K], B[K][K], dinv[K];
  // 1. S = Ls Ls'
BVG_UNROLL
  for (int j = 0; j < K; ++j) {
    double d = S[j + K * j];
BVG_UNROLL
    for (int c = 0; c < j; ++c) d = fma(-Ls[j][c], Ls[j][c], d);
    if (!(d > 0.0)) fail = 1;
    const double rs = bvg_rsqrt(d);
    dinv[j] = rs;
BVG_UNROLL
    for (int i = j + 1; i < K; ++i) {
      double v = S[i + K * j];
BVG_UNROLL
      for (int c = 0; c < j; ++c) v = fma(-Ls[i][c], Ls[j][c], v);
      Ls[i][j] = v * rs;
    }
  }
  // 2. Mi = Ls^-1 (lower), Sbar = Mi' Mi
BVG_UNROLL
  for (int c = 0; c < K; ++c) {
    Mi[c][c] = dinv[c];
BVG_UNROLL
    for (int i = c + 1; i < K; ++i) {
      double v = 0.0;
BVG_UNROLL
      for (int mm = c; mm < i; ++mm) v = fma(Ls[i][mm], Mi[mm][c], v);
      Mi[i][c] = -v * dinv[i];
    }
  }
BVG_UNROLL
  for (int i = 0; i < K; ++i)
BVG_UNROLL
    for (int j = 0; j <= i; ++j) {
      double v = 0.0;
BVG_UNROLL
      for (int mm = i; mm < K; ++mm) v = fma(Mi[mm][i], Mi[mm][j], v);
      Sb[i][j] = v;
    }
  // 3. Sbar = Ld Ld' with its true diagonal; D = Ld'
BVG_UNROLL
  for (int j = 0; j < K; ++j) {
    double d = Sb[j][j];
BVG_UNROLL
    for (int c = 0; c < j; ++c) d = fma(-Ld[j][c], Ld[j][c], d);
    if (!(d > 0.0)) fail = 1;
    const double rs = bvg_rsqrt(d);
    Ld[j][j] = 1.0 / rs;
BVG_UNROLL
    for (int i = j + 1; i < K; ++i) {
      double v = Sb[i][j];
BVG_UNROLL
      for (int c = 0; c < j; ++c) v = fma(-Ld[i][c], Ld[j][c], v);
      Ld[i][j] = v * rs;
    }
  }
  // 4./5. B = A D (upper): A_rr = sqrt(chi2_r), A_rc = wn[tri(c-1)+r] (r < c), D_mc = Ld_cm
BVG_UNROLL
  for (int r = 0; r < K; ++r)
BVG_UNROLL
    for (int c = r; c < K; ++c) {
      double v = 0.0;
BVG_UNROLL
      for (int mm = r; mm <= c; ++mm) {
        const double arm = (mm == r) ? chis[r] : wn[tri(mm - 1) + r];
        v = fma(arm, Ld[c][mm], v);
      }
      B[r][c] = v;
    }
  // 6. Sigma^-1 = B'B ; lane 0 publishes
  if (tid == 0) {
BVG_UNROLL
    for (int i = 0; i < K; ++i)
BVG_UNROLL
      for (int j = 0; j < K; ++j) {
        double v = 0.0;
BVG_UNROLL
        for (int mm = 0; mm <= (i < j ? i : j); ++mm) v = fma(B[mm][i], B[mm][j], v);
        sig[i + K * j] = v;
      }
    if (sigma_out != nullptr) {
      // Sigma = B^-1 B^-T
      double Bi[K][K];
BVG_UNROLL
      for (int c = 0; c < K; ++c)
BVG_UNROLL
        for (int r = c; r >= 0; --r) {
          double v = (r == c) ? 1.0 : 0.0;
BVG_UNROLL
          for (int mm = r + 1; mm <= c; ++mm) v = fma(-B[r][mm], Bi[mm][c], v);
          Bi[r][c] = v / B[r][r];
        }
BVG_UNROLL
      for (int i = 0; i < K; ++i)
BVG_UNROLL
        for (int j = 0; j < K; ++j) {
          double v = 0.0;
BVG_UNROLL
          for (int mm = (i > j ? i : j); mm < K; ++mm) v = fma(Bi[i][mm], Bi[j][mm], v);
          sigma_out[i + K * j] = v;
        }
    }
  }
  g.sync();
}

// general symmetric positive definite k x k inverse (full col-major), used for storing Sigma under the gamma prior
template <class Grp>
BVG_HD void small_spd_inverse(const Grp& g, const double* A, double* out, double* W1, double* W2, double* W3,
                              double* kd, int k, int& fail) {
  small_chol(g, A, W1, kd, k, fail);
  tri_inverse(g, W1, kd, W2, k);
  mtm_lower(g, W2, W3, k);
  for (int e = g.tid(); e < k * k; e += g.size()) {
    const int i = e % k, j = e / k;
    out[e] = (i >= j) ? W3[tri(i) + j] : W3[tri(j) + i];
  }
  g.sync();
}

// ---- the KSC-1998 / OCSN-2007 step for all k series of one chain (stochvol_ksc1998.cpp:89-105) -----------------
// U (k x T col-major residuals) is overwritten by y* = log(u^2 + c).  h: T x k col-major (t + T i).
// Work: pr (k*T), rh (k*T).  The indicator step is time-parallel; the tridiagonal factor/solve is one thread
// per series.
template <class Grp>
BVG_HD void sv_step(const Grp& g, const ChainRng& rng, int sweep, const BvarModel& m, double* U, double* h,
                    const double* sigma_h, const double* h_init, double* pr, double* rh, int* s_out = nullptr) {
  const int G = g.size(), tid = g.tid(), k = m.k, T = m.T, J = m.n_mix;
  // mixture constants once per call (the reference recomputes sqrt(s2_j) and sd_j sqrt(2 pi) per element: same values)
  double c_sd[10], c_den[10], c_mu[10], c_p[10];
  for (int j = 0; j < J; ++j) {
    c_sd[j] = sqrt(m.mix_s2[j]);
    c_den[j] = c_sd[j] * 2.5066282746310002;
    c_mu[j] = m.mix_mu[j];
    c_p[j] = m.mix_p[j];
  }
  for (int e = tid; e < k * T; e += G) {
    const int i = e % k, t = e / k;
    const double u = U[e];
    const double ys = log(u * u + m.sv_const[i]);                                  // :91
    const double ht = h[t + T * i];
    double q[10];
    double qs = 0.0;
    for (int j = 0; j < J; ++j) {
      const double z = (ys - (ht + c_mu[j])) / c_sd[j];
      q[j] = c_p[j] * (exp(-0.5 * (z * z)) / c_den[j]);                            // :94 (arma::normpdf)
      qs += q[j];
    }
    const double uu = rng.uniform(VB_SV_U, sweep, i * T + t);
    double cum = 0.0;
    int cnt = 0;
    for (int j = 0; j < J; ++j) {
      cum += q[j] / qs;                                                            // :95-96
      cnt += (uu < cum) ? 1 : 0;
    }
    int s = J - cnt;
    if (s > J - 1) s = J - 1;                                                      // reference: out of bounds
    if (s_out != nullptr) s_out[t + T * i] = s;
    const double p = 1.0 / m.mix_s2[s];                                            // :101
    pr[i * T + t] = p;
    rh[i * T + t] = p * (ys - c_mu[s]);                                            // :103 sigs * (y - mu_s)
  }
  g.sync();
  // The draw  h = P^-1 rhs + chol(P)^-1 eps  with P = tridiag(diag = {2,...,2,1} si + p_t, off = -si)  (:99-104) through
  // the lower bidiagonal factor  l_t^2 = c_t,  c_t = dg_t - si^2 / c_{t-1},  m_t = -si / l_t.  Only three SHORT
  // recurrences stay sequential (one lane per series); everything else is parallel over (series, t):
  //   (a) c_t = N_t / D_t by the division-free continued-fraction form  N_{t+1} = dg_{t+1} N_t - si^2 D_t, D_{t+1} = N_t
  //   (b) w_t = b_t / l_t + (si / (l_{t-1} l_t)) w_{t-1}                 forward solve
  //   (c) h_t = z_t / l_t + (si / l_t^2) h_{t+1},  z = w + eps            back substitution
  // eps: the residual array U is free now and takes the normals, site (series, t) -> U[i T + t].
  for (int e = tid; e < k * T; e += G) U[e] = rng.normal(VB_SV_N, sweep, e);
  for (int i = tid; i < k; i += G) {                                               // (a): pr <- N_t, h <- D_t
    const double si = 1.0 / sigma_h[i], si2 = si * si;                             // hh / sigma(i), :99
    double* p = pr + i * T;
    double* hd = h + T * i;
    double N = ((0 < T - 1) ? 2.0 : 1.0) * si + p[0], D = 1.0;
    p[0] = N; hd[0] = D;
    for (int t = 1; t < T; ++t) {
      const double dg = ((t < T - 1) ? 2.0 : 1.0) * si + p[t];
      const double Nn = fma(dg, N, -si2 * D);
      D = N; N = Nn;
      if ((t & 3) == 3) {                                                          // joint power-of-two rescale: ratio exact
        int ex;
        (void)frexp(N, &ex);
        N = ldexp(N, -ex); D = ldexp(D, -ex);
      }
      p[t] = N; hd[t] = D;
    }
  }
  g.sync();
  for (int e = tid; e < k * T; e += G) {                                           // pr <- 1 / l_t ; non-positive pivot -> NaN
    const double c = pr[e] / h[e];
    pr[e] = 1.0 / sqrt(c);
  }
  g.sync();
  for (int i = tid; i < k; i += G) {                                               // (b) and (c)
    const double si = 1.0 / sigma_h[i];
    const double* rl = pr + i * T;
    double* r = rh + i * T;
    const double* ep = U + i * T;
    double w = (r[0] + si * h_init[i]) * rl[0];                                    // hh * 1 * h_init = e_1 h_init
    r[0] = w + ep[0];
    for (int t = 1; t < T; ++t) {
      w = fma(si * rl[t - 1] * rl[t], w, r[t] * rl[t]);
      r[t] = w + ep[t];                                                            // z_t = w_t + eps_t
    }
    double hn = r[T - 1] * rl[T - 1];
    h[T - 1 + T * i] = hn;
    for (int t = T - 2; t >= 0; --t) {
      hn = fma(si * rl[t] * rl[t], hn, r[t] * rl[t]);
      h[t + T * i] = hn;
    }
  }
  g.sync();
}

#if defined(__CUDACC__)
// The same step for a chain owned by one full warp (device only).  Differences to the group-generic version above, all
// about latency -- with ~200 KB of shared memory per SM the L1 is a few KB, so local-memory arrays and global tables miss
// to L2 on every touch:
//   * the mixture tables live in registers (J is a template parameter: the component loops unroll, no local arrays) and,
//     for the data-dependent look-up of the selected component, in 2 J doubles of shared memory (`tab`);
//   * divisions by per-component constants are multiplications by their reciprocals and the J normalisations q_j / sum
//     are one reciprocal (differences of an ulp in the cumulative probabilities: an indicator changes only if the uniform
//     falls within ~1e-16 of a boundary);
//   * the log-volatility normals are drawn as Box-Muller pairs (same sites, half the Philox / log calls);
//   * the power-of-two rescale of the continued fraction works on the exponent bits (same values as frexp / ldexp).
template <int J>
static __device__ __noinline__ void sv_step_w32_j(const ChainRng& rng, int sweep, const int k, const int T,
                                                  const double* __restrict__ mix_p, const double* __restrict__ mix_mu,
                                                  const double* __restrict__ mix_s2, const double* __restrict__ sv_const,
                                                  double* U, double* h, const double* sigma_h, const double* h_init,
                                                  double* pr, double* rh, double* tab) {
  const int lane = threadIdx.x & 31;
  double c_mu[J], c_isd[J], c_pd[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const double sd = sqrt(mix_s2[j]);
    c_isd[j] = 1.0 / sd;
    c_pd[j] = mix_p[j] / (sd * 2.5066282746310002);
    c_mu[j] = mix_mu[j];
  }
  if (lane < J) { tab[lane] = 1.0 / mix_s2[lane]; tab[J + lane] = mix_mu[lane]; }
  __syncwarp();
  const bool inj_u = rng.injected(VB_SV_U);
  for (int e = lane; e < k * T; e += 32) {
    const int t = e / k, i = e - t * k;
    const double u = U[e];
    const double ys = log(fma(u, u, sv_const[i]));                                 // :91
    const double ht = h[t + T * i];
    double q[J];
    double qs = 0.0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const double z = (ys - (ht + c_mu[j])) * c_isd[j];
      q[j] = c_pd[j] * exp(-0.5 * (z * z));                                        // :94 (arma::normpdf)
      qs += q[j];
    }
    const double uu = inj_u ? rng.inj_at(VB_SV_U, sweep, i * T + t) : philox_uniform(rng.seed, rng.chain, VB_SV_U, sweep, i * T + t, 0);
    const double rq = 1.0 / qs;
    double cum = 0.0;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      cum = fma(q[j], rq, cum);                                                    // :95-96
      cnt += (uu < cum) ? 1 : 0;
    }
    int s = J - cnt;
    if (s > J - 1) s = J - 1;                                                      // reference: out of bounds
    const double p = tab[s];                                                       // :101
    pr[i * T + t] = p;
    rh[i * T + t] = p * (ys - tab[J + s]);                                         // :103 sigs * (y - mu_s)
  }
  __syncwarp();
  // eps: the residual array U is free now and takes the normals, site (series, t) -> U[i T + t]
  {
    const int nsite = k * T, npair = (nsite + 1) >> 1;
    for (int p = lane; p < npair; p += 32) {
      const Normal2 nn = rng.normal_pair(VB_SV_N, sweep, p, nsite);
      U[2 * p] = nn.a;
      if (2 * p + 1 < nsite) U[2 * p + 1] = nn.b;
    }
  }
  if (lane < k) {                                                                  // (a): pr <- N_t, h <- D_t
    const int i = lane;
    const double si = 1.0 / sigma_h[i], si2 = si * si;                             // hh / sigma(i), :99
    double* p = pr + i * T;
    double* hd = h + T * i;
    double N = ((0 < T - 1) ? 2.0 : 1.0) * si + p[0], D = 1.0;
    p[0] = N; hd[0] = D;
#pragma unroll 4
    for (int t = 1; t < T; ++t) {
      const double dg = ((t < T - 1) ? 2.0 : 1.0) * si + p[t];
      const double Nn = fma(dg, N, -si2 * D);
      D = N; N = Nn;
      if ((t & 3) == 3) {                                                          // joint power-of-two rescale: ratio exact
        int f = 2045 - ((__double2hiint(N) >> 20) & 0x7ff);                        // N 2^(f - 1023) lies in [0.5, 1)
        f = f < 1 ? 1 : (f > 2046 ? 2046 : f);
        const double sc = __hiloint2double(f << 20, 0);
        N *= sc; D *= sc;
      }
      p[t] = N; hd[t] = D;
    }
  }
  __syncwarp();
  for (int e = lane; e < k * T; e += 32) {                                         // pr <- 1 / l_t ; non-positive pivot -> NaN
    const double c = pr[e] / h[e];
    pr[e] = 1.0 / sqrt(c);
  }
  __syncwarp();
  if (lane < k) {                                                                  // (b) and (c)
    const int i = lane;
    const double si = 1.0 / sigma_h[i];
    const double* rl = pr + i * T;
    double* r = rh + i * T;
    const double* ep = U + i * T;
    double w = (r[0] + si * h_init[i]) * rl[0];                                    // hh * 1 * h_init = e_1 h_init
    r[0] = w + ep[0];
#pragma unroll 4
    for (int t = 1; t < T; ++t) {
      w = fma(si * rl[t - 1] * rl[t], w, r[t] * rl[t]);
      r[t] = w + ep[t];                                                            // z_t = w_t + eps_t
    }
    double hn = r[T - 1] * rl[T - 1];
    h[T - 1 + T * i] = hn;
#pragma unroll 4
    for (int t = T - 2; t >= 0; --t) {
      hn = fma(si * rl[t] * rl[t], hn, r[t] * rl[t]);
      h[t + T * i] = hn;
    }
  }
  __syncwarp();
}

// tab: 20 doubles of shared memory
static __device__ __forceinline__ void sv_step_w32(const ChainRng& rng, int sweep, const BvarModel& m, double* U, double* h,
                                                   const double* sigma_h, const double* h_init, double* pr, double* rh,
                                                   double* tab) {
  if (m.n_mix == 7) sv_step_w32_j<7>(rng, sweep, m.k, m.T, m.mix_p, m.mix_mu, m.mix_s2, m.sv_const, U, h, sigma_h, h_init, pr, rh, tab);
  else sv_step_w32_j<10>(rng, sweep, m.k, m.T, m.mix_p, m.mix_mu, m.mix_s2, m.sv_const, U, h, sigma_h, h_init, pr, rh, tab);
}

// Tile-packed precision matrix for the CTA-per-chain kernels (bvg_tiled.cuh): P = V^-1 + Lambda (z'Dz) Lambda written
// half tile by half tile (a warp stores 32 consecutive doubles: no bank conflicts, no per-row imbalance), identity
// padding rows included.  GW != nullptr: SV form (block-diagonal weights), else XX' (x) Sigma^-1.   (bvaralg.cpp:303-305)
__device__ __forceinline__ void tiled_build_precision(double* L, int n, int k, int M, const double* XX, const double* sig,
                                                      const double* GW, const double* vdiag, const double* lam_bvs,
                                                      const double* vi_off, const double* PsiM = nullptr) {
  const int nt = tiled_nt(n), lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int li = lane & 7, lj = lane >> 3;
  const float kinv = 1.0f / (float)k;
  const int triM = tri(M);
  for (int I = warp; I < nt; I += nwarp) {
    const int r = 8 * I + li;
    const int jr = (int)(((float)r + 0.5f) * kinv), ir = r - jr * k;
    const double lr = (lam_bvs != nullptr && r < n) ? lam_bvs[r] : 1.0;
    double* row = L + (((I * (I + 1)) >> 1) << 6) + lane;
    for (int J = 0; J <= I; ++J) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = 8 * J + lj + 4 * h;
        double gv = 0.0;
        if (r >= n) gv = (c == r) ? 1.0 : 0.0;
        else if (c <= r) {
          const int jc = (int)(((float)c + 0.5f) * kinv), ic = c - jc * k;
          if (GW != nullptr && PsiM != nullptr) {                     // SV with the covariance block: Psi' Omega_t^-1 Psi
            for (int q = (ir > ic ? ir : ic); q < k; ++q) gv = fma(GW[q * triM + tri(jr) + jc], PsiM[q + k * ir] * PsiM[q + k * ic], gv);
          } else if (GW != nullptr) gv = (ir == ic) ? GW[ir * triM + tri(jr) + jc] : 0.0;
          else gv = XX[jr + M * jc] * sig[ir + k * ic];
          if (lam_bvs != nullptr) gv *= lr * lam_bvs[c];
          if (vi_off != nullptr) gv += vi_off[tri(r) + c];
          if (c == r) gv += vdiag[r];
        }
        row[J * 64 + 32 * h] = gv;
      }
    }
  }
}
#endif

// ---- the chain ---------------------------------------------------------------------------------------------
// TILED (CTA-per-chain device kernels only): L is 8 x 8 tile-packed and factorised with DMMA trailing updates
// (bvg_tiled.cuh) instead of the row-packed Crout factorisation.
template <int FEAT, class Grp, bool TILED = false>
BVG_HD void bvar_chain(const Grp& g, const BvarModel& m, const BvarConsts& cs, const ChainRng& rng, double* state,
                       double* sm, const BvarOut& out, long long chain_local, int sweep_begin, int sweep_end,
                       int burnin) {
  const int G = g.size(), tid = g.tid();
  const int k = m.k, T = m.T, M = m.M, n = m.n, kk = k * k;
  const bool sv = (FEAT & F_SV) && m.sigma_type == SIG_SV;
  const bool direct = (FEAT & F_DIRECT) && ((m.resid_mode == RES_DIRECT) || sv);
  const int varsel = (FEAT & F_VARSEL) ? m.varsel : (int)VS_NONE;
  const bool use_gamma = (FEAT & F_GAMMA) && m.sigma_type == SIG_GAMMA;
  const double* vi_off = (FEAT & F_VIOFF) ? m.vi_off : nullptr;
  const double* vmu_off = (FEAT & F_VIOFF) ? m.vmu_off : nullptr;
  // carve shared memory
  double* L = sm;           sm += TILED ? tiled_len(n) : tri(n);
  double* w = sm;           sm += TILED ? n + 8 : n;
  double* a = sm;           sm += n;
  double* dinv = sm;        sm += TILED ? 0 : n;
  double* vdiag = sm;       sm += n;
  double* lam = sm;         sm += n;
  double* sig = sm;         sm += kk;
  double* S = sm;           sm += kk;
  double* W1 = sm;          sm += kk;
  double* W2 = sm;          sm += kk;
  double* W3 = sm;          sm += kk;
  double* W4 = sm;          sm += kk;
  double* kd = sm;          sm += k;
  double* kv = sm;          sm += k;
  double* Dm = sm;          sm += k * M;
  double* Wm = sm;          sm += k * M;
  double* U = nullptr; double* h = nullptr; double* ew = nullptr; double* svw = nullptr; double* GW = nullptr;
  double* sigma_h = nullptr; double* h_init = nullptr;
  if (direct) { U = sm; sm += k * T; }
  if (sv) {
    h = sm; sm += k * T;  ew = sm; sm += k * T;  svw = sm; sm += k * T;  GW = sm; sm += k * tri(M);
    sigma_h = sm; sm += k;  h_init = sm; sm += k;
  }
  // covariance (psi) block (bvaralg.cpp:373-455): only compiled into the variants that know gamma / SV variances
  const int npsi = (FEAT & (F_GAMMA | F_SV)) ? m.n_psi : 0;
  const bool covar = npsi > 0;
  double* pL = nullptr; double* prhs = nullptr; double* psi = nullptr; double* pdinv = nullptr; double* omd = nullptr;
  double* PsiM = nullptr; double* pvd = nullptr; double* plam = nullptr; double* pwork = nullptr;
  if (covar) {
    pL = sm; sm += tri(npsi);  prhs = sm; sm += npsi;  psi = sm; sm += npsi;  pdinv = sm; sm += npsi;
    pvd = sm; sm += npsi;  plam = sm; sm += npsi;  pwork = sm; sm += npsi;
    omd = sm; sm += k;  PsiM = sm; sm += kk;
  }
  const int pvs = covar ? m.psi_varsel : VS_NONE;
  int fail = 0;

  // load state
  double* st_sig = state;
  double* st_vd = state + kk;
  double* st_lam = st_vd + n;
  double* st_h = st_lam + n;
  for (int e = tid; e < kk; e += G) sig[e] = st_sig[e];
  for (int e = tid; e < n; e += G) { vdiag[e] = st_vd[e]; lam[e] = st_lam[e]; }
  if (sv) {
    for (int e = tid; e < T * k; e += G) h[e] = st_h[e];
    for (int e = tid; e < k; e += G) { sigma_h[e] = st_h[T * k + e]; h_init[e] = st_h[T * k + k + e]; }
  }
  double* st_tail = state + bvar_state_len(k, T, n, m.sigma_type);       // [omega diagonal (k) | psi (npsi)]
  if (covar) {
    for (int e = tid; e < k; e += G) omd[e] = st_tail[e];
    for (int e = tid; e < npsi; e += G) { psi[e] = st_tail[k + e]; pvd[e] = st_tail[k + npsi + e]; plam[e] = st_tail[k + 2 * npsi + e]; }
    g.sync();
    for (int e = tid; e < kk; e += G) {                                  // Psi: unit lower, row i holds psi[tri(i-1) ..]
      const int i = e % k, a = e / k;
      PsiM[e] = (i == a) ? 1.0 : ((a < i) ? psi[(i * (i - 1)) / 2 + a] : 0.0);
    }
  }
  g.sync();

  for (int sweep = sweep_begin; sweep < sweep_end; ++sweep) {
    const bool keep = sweep >= burnin;
    const long long pos = (long long)sweep - burnin;

    // ---- variates of this sweep that do not depend on the chain state, drawn in ONE pass through one generator
    //      call site (no divergence between lanes): coefficient normals -> a[], and for the register Wishart block
    //      its strict-upper normals -> W4[] and the attempt-0 normals of its chi-squares -> kv[]
    const bool wish_reg = !sv && !use_gamma && k <= (Grp::kIsBlock ? 6 : 4);   // K = 5, 6: CTA kernels only
    {
      const int nwn = (k * (k - 1)) / 2;
      const int n_items = n + (wish_reg ? nwn + k : 0);
      for (int it = tid; it < n_items; it += G) {
        int block = VB_COEF_N, idx = it;
        double* dst = a + it;
        if (it >= n + nwn) { block = VB_WISH_C; idx = it - n - nwn; dst = kv + idx; }
        else if (it >= n) { block = VB_WISH_N; idx = it - n; dst = W4 + idx; }
        *dst = rng.normal(block, sweep, idx);
      }
    }
    if (n > 0) {
      // ---- precision P = V^-1 + Lambda (z'Dz) Lambda and rhs b = V^-1 mu + Lambda z'Dy  (bvaralg.cpp:303-306)
      if (sv) {
        // :494, ew[t + T i]; the very first sweep uses h.row(0) for every period (bvaralg.cpp:230-231,246)
        for (int e = tid; e < k * T; e += G) ew[e] = 1.0 / exp(sweep == 0 ? h[T * (e / T)] : h[e]);
        g.sync();
        for (int e = tid; e < k * tri(M); e += G) {
          const int i = e / tri(M), f = e % tri(M);
          int j = 0; while (tri(j + 1) <= f) ++j;
          const int j2 = f - tri(j);
          double s = 0.0;
          for (int t = 0; t < T; ++t) s = fma(m.X[j + M * t] * m.X[j2 + M * t], ew[t + T * i], s);
          GW[e] = s;
        }
        for (int r = tid; r < n; r += G) {
          const int j = r / k, i = r % k;
          double s = 0.0;
          if (covar) {
            // Sigma_t^-1 = Psi' Omega_t^-1 Psi:  (Sigma_t^-1 y_t)_i = sum_{q >= i} Psi_qi w_tq (Psi y_t)_q   (:495-497)
            for (int t = 0; t < T; ++t) {
              double v = 0.0;
              for (int q = i; q < k; ++q) {
                double py = 0.0;
                for (int b = 0; b <= q; ++b) py = fma(PsiM[q + k * b], m.Y[b + k * t], py);
                v = fma(PsiM[q + k * i] * ew[t + T * q], py, v);
              }
              s = fma(m.X[j + M * t], v, s);
            }
          } else {
            for (int t = 0; t < T; ++t) s = fma(m.X[j + M * t] * ew[t + T * i], m.Y[i + k * t], s);
          }
          w[r] = s;
        }
        g.sync();
      } else {
        for (int r = tid; r < n; r += G) {
          const int j = r / k, i = r % k;
          double s = 0.0;
          for (int i2 = 0; i2 < k; ++i2) s = fma(sig[i + k * i2], cs.YX[i2 + k * j], s);  // (Sigma^-1 Y X')_ij
          w[r] = s;
        }
      }
      // regressor mask of the coefficient draw: lambda (BVS), or the static mask of a structural model under SSVS
      const double* zmask = (varsel == VS_BVS) ? lam : (((FEAT & F_VARSEL) && varsel == VS_SSVS) ? m.maskd : nullptr);
      const bool bvs = zmask != nullptr;
#if defined(__CUDA_ARCH__)
      if constexpr (TILED) {
        for (int r = tid; r < tiled_nt(n) * 8; r += G) {
          double b = 0.0;                                          // zero padding of the rhs
          if (r < n) {
            b = w[r] * (bvs ? zmask[r] : 1.0) + vdiag[r] * cs.mu[r];
            if (vmu_off != nullptr) b += vmu_off[r];
          }
          w[r] = b;
        }
        tiled_build_precision(L, n, k, M, cs.XX, sig, sv ? GW : nullptr, vdiag, zmask, vi_off, (sv && covar) ? PsiM : nullptr);
      } else
#endif
      {
        for (int r = tid; r < n; r += G) {
          const int j = r / k, i = r % k;
          double* Lr = L + tri(r);
#define BVG_LSET(c, v) Lr[c] = (v)
          const double lr = bvs ? zmask[r] : 1.0;
          if (sv) {
            int j2 = 0, i2 = 0;
            for (int c = 0; c <= r; ++c) {
              double gv = 0.0;
              if (covar) {   // sum_q GW_q[j, j2] Psi_qi Psi_qi2
                for (int q = (i > i2 ? i : i2); q < k; ++q) gv = fma(GW[q * tri(M) + tri(j) + j2], PsiM[q + k * i] * PsiM[q + k * i2], gv);
              } else if (i == i2) gv = GW[i * tri(M) + tri(j) + j2];
              if (bvs) gv *= lr * zmask[c];
              if (vi_off != nullptr) gv += vi_off[tri(r) + c];
              if (c == r) gv += vdiag[r];
              BVG_LSET(c, gv);
              if (++i2 == k) { i2 = 0; ++j2; }
            }
          } else {
            // row (j,i) of (XX' (x) Sigma^-1): one XX entry per regressor block, k products each
            const double* sigrow = sig + i;
            int c = 0;
            for (int j2 = 0; j2 <= j; ++j2) {
              const double xv = cs.XX[j + M * j2];
              const int i2max = (j2 < j) ? k - 1 : i;
              for (int i2 = 0; i2 <= i2max; ++i2, ++c) {
                double gv = xv * sigrow[k * i2];
                if (bvs) gv *= lr * zmask[c];
                if (vi_off != nullptr) gv += vi_off[tri(r) + c];
                if (c == r) gv += vdiag[r];
                BVG_LSET(c, gv);
              }
            }
          }
#undef BVG_LSET
          double b = w[r] * lr + vdiag[r] * cs.mu[r];
          if (vmu_off != nullptr) b += vmu_off[r];
          w[r] = b;
        }
      }
      g.sync();
      // ---- a = L^-T (L^-1 b + eps)                                                  (bvaralg.cpp:306-307)
#if defined(__CUDA_ARCH__)
      if constexpr (TILED) {
        tiled_chol(L, w, n, fail);
        for (int r = tid; r < n; r += G) w[r] += a[r];
        g.sync();
        tiled_back_solve(L, w, a, n);
      } else
#endif
      {
        chol_crout_sel<true>(g, L, dinv, w, n, fail);
        for (int r = tid; r < n; r += G) w[r] += a[r];            // eps drawn at the top of the sweep
        g.sync();
        back_solve_lt(g, L, dinv, w, a, n);
      }

      // ---- SSVS (bvaralg.cpp:314-331)
      if (varsel == VS_SSVS) {
        if (zmask != nullptr) {                                    // structural model: masked positions are not coefficients
          for (int r = tid; r < n; r += G) a[r] *= zmask[r];
          g.sync();
        }
        for (int r = tid; r < n; r += G) {
          if (!m.include[r]) continue;
          const double t0 = m.tau0[r], t1 = m.tau1[r], pi = m.inprior[r];
          const double t0sq = t0 * t0, t1sq = t1 * t1, ar = a[r];
          const double u0 = 1.0 / t0 * exp(-((ar * ar) / (2.0 * t0sq))) * (1.0 - pi);     // :316
          const double u1 = 1.0 / t1 * exp(-((ar * ar) / (2.0 * t1sq))) * pi;             // :317
          const double p = u1 / (u0 + u1);
          const double ld = bernoulli_rbinom(p, rng.uniform(VB_SSVS_U, sweep, r));
          lam[r] = ld;
          vdiag[r] = (ld == 0.0) ? 1.0 / t0sq : 1.0 / t1sq;                               // :324-328
        }
        g.sync();
      }
    }

    // ---- residuals and BVS -------------------------------------------------------------------------------
    // Dm = A_AG - A_ols  (moment mode)  /  U = Y - A_AG X  (direct mode), with A_AG = Lambda a (stale, :335)
    const bool bvs = (varsel == VS_BVS) && n > 0;
    if (direct) {
      for (int e = tid; e < k * T; e += G) {
        const int i = e % k, t = e / k;
        double s = m.Y[e];
        for (int j = 0; j < M; ++j) {
          const double av = bvs ? lam[j * k + i] * a[j * k + i] : a[j * k + i];
          s = fma(-av, m.X[j + M * t], s);
        }
        U[e] = s;
      }
      g.sync();
    } else {
      for (int r = tid; r < n; r += G) Dm[r] = (bvs ? lam[r] * a[r] : a[r]) - cs.Aols[r];  // r = i + k j
      g.sync();
      for (int r = tid; r < n; r += G) {
        const int j = r / k, i = r % k;
        double s = 0.0;
        for (int j2 = 0; j2 < M; ++j2) s = fma(Dm[i + k * j2], cs.XX[j2 + M * j], s);
        Wm[r] = s;                                                                        // (D XX')_ij
      }
      g.sync();
    }
    if (bvs) {
      // Korobilis (2013) indicator update, coefficient-parallel rank-1 form:
      //   r_theta = r_base + z_j delta,  delta0 = lambda_j a_j, delta1 = (lambda_j - 1) a_j,  r_base = y - z a_AG
      //   q(delta) = q_base + 2 delta c_j + delta^2 g_j,   c_j = z_j' D r_base,  g_j = z_j' D z_j
      //   l1 - l0 = -(q1 - q0)/2 + (lp1 - lp0)                                          (bvaralg.cpp:342-350)
      for (int r = tid; r < n; r += G) {
        double newlam = lam[r];
        if (m.include[r]) {
          const int j = r / k, i = r % k;
          const double lp0 = log(1.0 - m.inprior[r]), lp1 = log(m.inprior[r]);            // :124-125
          const double g1 = log(rng.uniform(VB_BVS_U1, sweep, r));                        // :338
          const bool gate = (lam[r] == 1.0) ? (g1 < lp1) : (g1 < lp0);                    // :339-341
          if (gate) {
            double cj = 0.0, gj = 0.0;
            if (direct) {
              if (sv && covar) {
                // Sigma_t^-1 = Psi' Omega_t^-1 Psi (previous sweep's Psi):  c_j = sum_t x_jt (Sigma_t^-1 u_t)_i,
                // g_j = sum_t x_jt^2 (Sigma_t^-1)_ii
                for (int t = 0; t < T; ++t) {
                  double v = 0.0, s2 = 0.0;
                  for (int q = i; q < k; ++q) {
                    double pu = 0.0;
                    for (int b = 0; b <= q; ++b) pu = fma(PsiM[q + k * b], U[b + k * t], pu);
                    const double pw = PsiM[q + k * i] * ew[t + T * q];
                    v = fma(pw, pu, v);
                    s2 = fma(pw, PsiM[q + k * i], s2);
                  }
                  const double x = m.X[j + M * t];
                  cj = fma(x, v, cj);
                  gj = fma(x * x, s2, gj);
                }
              } else if (sv) {
                for (int t = 0; t < T; ++t) {
                  const double xw = m.X[j + M * t] * ew[t + T * i];
                  cj = fma(xw, U[i + k * t], cj);
                  gj = fma(xw, m.X[j + M * t], gj);
                }
              } else {
                for (int t = 0; t < T; ++t) {
                  double v = 0.0;
                  for (int i2 = 0; i2 < k; ++i2) v = fma(sig[i + k * i2], U[i2 + k * t], v);
                  cj = fma(m.X[j + M * t], v, cj);
                }
                gj = cs.XX[j + M * j] * sig[i + k * i];
              }
            } else {
              // c_j = [Sigma^-1 (Y - A_AG X) X']_ij = -[Sigma^-1 D XX']_ij
              for (int i2 = 0; i2 < k; ++i2) cj = fma(-sig[i + k * i2], Wm[i2 + k * j], cj);
              gj = cs.XX[j + M * j] * sig[i + k * i];
            }
            const double ar = a[r];
            const double d0 = lam[r] * ar, d1 = (lam[r] - 1.0) * ar;
            const double dq = 2.0 * (d1 - d0) * cj + (d1 * d1 - d0 * d0) * gj;           // q1 - q0
            const double bayes = -0.5 * dq + (lp1 - lp0);
            const double g2 = log(rng.uniform(VB_BVS_U2, sweep, r));                      // :351
            newlam = (bayes >= g2) ? 1.0 : 0.0;                                           // :352-356
          }
        }
        w[r] = newlam;      // w is free here; commit after everyone has read lam
      }
      g.sync();
      // a = Lambda a (:359); residuals must be refreshed for the new Lambda (:364)
      for (int r = tid; r < n; r += G) { lam[r] = w[r]; a[r] *= w[r]; }
      g.sync();
      if (direct) {
        for (int e = tid; e < k * T; e += G) {
          const int i = e % k, t = e / k;
          double s = m.Y[e];
          for (int j = 0; j < M; ++j) s = fma(-a[j * k + i], m.X[j + M * t], s);
          U[e] = s;
        }
        g.sync();
      } else {
        for (int r = tid; r < n; r += G) Dm[r] = a[r] - cs.Aols[r];
        g.sync();
        for (int r = tid; r < n; r += G) {
          const int j = r / k, i = r % k;
          double s = 0.0;
          for (int j2 = 0; j2 < M; ++j2) s = fma(Dm[i + k * j2], cs.XX[j2 + M * j], s);
          Wm[r] = s;
        }
        g.sync();
      }
    }

    // ---- covariance (psi) block (bvaralg.cpp:373-455) ----------------------------------------------------------
    // Row i of U regressed on -U[0:i]: the posterior precision is V_psi^-1 + blockdiag_i(sum_t w_ti u_{0:i,t} u_{0:i,t}'),
    // w_ti the Omega^-1 diagonal of the PREVIOUS sweep (:384).  Then U <- Psi U (:454).
    if (covar) {
      for (int e = tid; e < tri(npsi); e += G) {
        int r = 0; while (tri(r + 1) <= e) ++r;
        const int c = e - tri(r);
        int i = 1; while ((i * (i + 1)) / 2 <= r) ++i;                // psi index r = i (i - 1) / 2 + a  ->  row i, column a
        const int a = r - (i * (i - 1)) / 2;
        int i2 = 1; while ((i2 * (i2 + 1)) / 2 <= c) ++i2;
        const int b = c - (i2 * (i2 - 1)) / 2;
        double v = (r == c) ? pvd[r] : m.psi_vi[r + (size_t)npsi * c];
        if (i == i2) {
          double s = 0.0;
          for (int t = 0; t < T; ++t) {
            const double wt = sv ? ew[t + T * i] : omd[i];
            s = fma(U[a + k * t] * wt, U[b + k * t], s);
          }
          if (pvs == VS_BVS) s *= plam[r] * plam[c];                                     // psi_z = psi_z_bvs Lambda, :388-391
          v += s;
        }
        pL[e] = v;
      }
      for (int r = tid; r < npsi; r += G) {
        int i = 1; while ((i * (i + 1)) / 2 <= r) ++i;
        const int a = r - (i * (i - 1)) / 2;
        double s = 0.0;
        for (int t = 0; t < T; ++t) {
          const double wt = sv ? ew[t + T * i] : omd[i];
          s = fma(-U[a + k * t] * wt, U[i + k * t], s);
        }
        if (pvs == VS_BVS) s *= plam[r];
        prhs[r] = m.psi_vmu[r] + pvd[r] * m.psi_mu[r] + s;                                // :393
      }
      g.sync();
      chol_crout_generic(g, pL, pdinv, prhs, npsi, fail);
      for (int r = tid; r < npsi; r += G) prhs[r] += rng.normal(VB_PSI_N, sweep, r);      // :394
      g.sync();
      back_solve_lt_generic(g, pL, pdinv, prhs, psi, npsi);
      if (pvs == VS_SSVS) {                                                               // :401-418
        for (int r = tid; r < npsi; r += G) {
          if (!m.psi_include[r]) continue;
          const double t0 = m.psi_tau0[r], t1 = m.psi_tau1[r], pi = m.psi_inprior[r];
          const double t0sq = t0 * t0, t1sq = t1 * t1, pr = psi[r];
          const double u0 = 1.0 / t0 * exp(-((pr * pr) / (2.0 * t0sq))) * (1.0 - pi);
          const double u1 = 1.0 / t1 * exp(-((pr * pr) / (2.0 * t1sq))) * pi;
          const double p = u1 / (u0 + u1);
          const double ld = bernoulli_rbinom(p, rng.uniform(VB_PSI_U1, sweep, r));
          plam[r] = ld;
          pvd[r] = (ld == 0.0) ? 1.0 / t0sq : 1.0 / t1sq;
        }
        g.sync();
      }
      if (pvs == VS_BVS) {                                                                // :420-448
        // stale psi_AG = Lambda psi; equation i only: residual(delta) = rb_it + u_at delta with
        // rb_it = u_it + sum_a' u_a't lambda_ia' psi_ia',  delta0 = -lambda psi, delta1 = (1 - lambda) psi
        for (int r = tid; r < npsi; r += G) {
          double newlam = plam[r];
          if (m.psi_include[r]) {
            int i = 1; while ((i * (i + 1)) / 2 <= r) ++i;
            const int a = r - (i * (i - 1)) / 2;
            const double lp0 = log(1.0 - m.psi_inprior[r]), lp1 = log(m.psi_inprior[r]);
            const double g1 = log(rng.uniform(VB_PSI_U1, sweep, r));
            const bool gate = (plam[r] == 1.0) ? (g1 < lp1) : (g1 < lp0);
            if (gate) {
              double cj = 0.0, gj = 0.0;
              for (int t = 0; t < T; ++t) {
                double rb = U[i + k * t];
                for (int a2 = 0; a2 < i; ++a2) rb = fma(U[a2 + k * t], plam[(i * (i - 1)) / 2 + a2] * psi[(i * (i - 1)) / 2 + a2], rb);
                const double wt = sv ? ew[t + T * i] : omd[i];
                const double ua = U[a + k * t];
                cj = fma(ua * wt, rb, cj);
                gj = fma(ua * wt, ua, gj);
              }
              const double pr = psi[r];
              const double d0 = -plam[r] * pr, d1 = (1.0 - plam[r]) * pr;
              const double dq = 2.0 * (d1 - d0) * cj + (d1 * d1 - d0 * d0) * gj;          // q1 - q0
              const double bayes = -0.5 * dq + (lp1 - lp0);
              const double g2 = log(rng.uniform(VB_PSI_U2, sweep, r));
              newlam = (bayes >= g2) ? 1.0 : 0.0;
            }
          }
          pwork[r] = newlam;
        }
        g.sync();
        for (int r = tid; r < npsi; r += G) { plam[r] = pwork[r]; psi[r] *= pwork[r]; }   // :446
        g.sync();
      }
      for (int e = tid; e < kk; e += G) {
        const int i = e % k, a = e / k;
        PsiM[e] = (i == a) ? 1.0 : ((a < i) ? psi[(i * (i - 1)) / 2 + a] : 0.0);        // :451-453
      }
      g.sync();
      for (int t = tid; t < T; t += G)                                                    // u = Psi u, :454
        for (int i = k - 1; i >= 1; --i) {
          double s = U[i + k * t];
          for (int a = 0; a < i; ++a) s = fma(PsiM[i + k * a], U[a + k * t], s);
          U[i + k * t] = s;
        }
      g.sync();
    }

    // ---- error variances ---------------------------------------------------------------------------------
    double* out_sig = nullptr;
    if (keep) out_sig = out.sigma + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)(sv ? kk * T : kk);
    if (sv) {
      sv_step(g, rng, sweep, m, U, h, sigma_h, h_init, ew, svw);                          // :462
      // sigma_h (:465-471)
      for (int i = tid; i < k; i += G) {
        double ss = 0.0, prev = h_init[i];
        for (int t = 0; t < T; ++t) { const double d = h[t + T * i] - prev; ss = fma(d, d, ss); prev = h[t + T * i]; }
        const double scale = 1.0 / (m.g_rate[i] + ss * 0.5);
        kv[i] = 1.0 / (scale * rng.gamma(VB_SIGH_G, sweep, i, m.g_shape_post[i]));
      }
      g.sync();
      for (int i = tid; i < k; i += G) sigma_h[i] = kv[i];
      // h_init (:474-477): (V0^-1 + diag(1/sigma_h)) k x k
      for (int e = tid; e < kk; e += G) {
        const int i = e % k, j = e / k;
        S[e] = m.sv_vi[e] + ((i == j) ? 1.0 / kv[i] : 0.0);
      }
      for (int i = tid; i < k; i += G) {
        double s = (1.0 / kv[i]) * h[0 + T * i];
        for (int j = 0; j < k; ++j) s = fma(m.sv_vi[i + k * j], m.sv_mu[j], s);
        W4[i] = s;
      }
      g.sync();
      for (int e = tid; e < tri(k); e += G) {
        int i = 0; while (tri(i + 1) <= e) ++i;
        W1[e] = S[i + k * (e - tri(i))];
      }
      g.sync();
      chol_crout_generic(g, W1, kd, W4, k, fail);
      for (int i = tid; i < k; i += G) W4[i] += rng.normal(VB_HINIT_N, sweep, i);
      g.sync();
      back_solve_lt_generic(g, W1, kd, W4, h_init, k);
      if (keep) {
        // Sigma_t = diag(exp(h_t)) (solve of the diagonal block, :523-525); sigma_h (:526)
        if (covar) {
          // Sigma_t = (Psi' Omega_t^-1 Psi)^-1 = Pi diag(exp(h_t)) Pi',  Pi = Psi^-1 (unit lower), :495-497,523-525
          for (int c = tid; c < k; c += G)
            for (int i = 0; i < k; ++i) {
              double s = (i == c) ? 1.0 : 0.0;
              if (i > c) { s = 0.0; for (int a = c; a < i; ++a) s = fma(-PsiM[i + k * a], W1[a + k * c], s); }
              W1[i + k * c] = (i >= c) ? s : 0.0;
            }
          g.sync();
          for (int e = tid; e < kk * T; e += G) {
            const int t = e / kk, f = e % kk, i = f % k, j = f / k;
            double s = 0.0;
            for (int q = 0; q <= (i < j ? i : j); ++q) s = fma(W1[i + k * q] * (1.0 / (1.0 / exp(h[t + T * q]))), W1[j + k * q], s);
            out_sig[e] = s;
          }
        } else
        for (int e = tid; e < kk * T; e += G) {
          const int t = e / kk, f = e % kk, i = f % k, j = f / k;
          out_sig[e] = (i == j) ? 1.0 / (1.0 / exp(h[t + T * i])) : 0.0;
        }
        double* osh = out.sigma_h + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)k;
        for (int i = tid; i < k; i += G) osh[i] = kv[i];
      }
      g.sync();
    } else {
      // S = S0 + u u'  /  sse diagonal
      for (int e = tid; e < kk; e += G) {
        const int i = e % k, j = e / k;
        double s = 0.0;
        if (direct) {
          for (int t = 0; t < T; ++t) s = fma(U[i + k * t], U[j + k * t], s);
        } else {
          s = cs.SSE[e];
          for (int j2 = 0; j2 < M; ++j2) s = fma(Wm[i + k * j2], Dm[j + k * j2], s);
        }
        S[e] = s;
      }
      g.sync();
      if (use_gamma) {
        for (int i = tid; i < k; i += G) {
          const double gs = rng.gamma(VB_OMEGA_G, sweep, i, m.g_shape_post[i]);
          kv[i] = gs * (1.0 / (m.g_rate[i] + S[i + k * i] * 0.5));                        // :484
        }
        g.sync();
        for (int e = tid; e < kk; e += G) {
          const int i = e % k, j = e / k;
          sig[e] = (i == j) ? kv[i] : m.omega_off[e];                                     // :507
        }
        g.sync();
        if (covar) {                                                                      // :504-505
          for (int i = tid; i < k; i += G) omd[i] = kv[i];
          for (int e = tid; e < kk; e += G) W1[e] = sig[e];                               // omega_i
          g.sync();
          for (int e = tid; e < kk; e += G) {                                             // sigma_i = Psi' omega_i Psi
            const int a = e % k, b = e / k;
            double s = 0.0;
            for (int i = a; i < k; ++i) {
              double v = 0.0;
              for (int i2 = b; i2 < k; ++i2) v = fma(W1[i + k * i2], PsiM[i2 + k * b], v);
              s = fma(PsiM[i + k * a], v, s);
            }
            sig[e] = s;
          }
          g.sync();
        }
        if (keep) {
#if defined(__CUDA_ARCH__)
          if constexpr (Grp::kIsBlock) {
            if (tid < 32) {
              const TileGroup<32> wg;
              small_spd_inverse(wg, sig, W4, W1, W2, W3, kd, k, fail);
              for (int e = tid; e < kk; e += 32) out_sig[e] = W4[e];
            }
          } else
#endif
          {
            small_spd_inverse(g, sig, W4, W1, W2, W3, kd, k, fail);                       // :528
            for (int e = tid; e < kk; e += G) out_sig[e] = W4[e];
          }
        }
      } else {
        for (int e = tid; e < kk; e += G) S[e] += cs.S0[e];
        g.sync();
        // symmetrise exactly (moment mode computes both triangles independently)
        for (int e = tid; e < kk; e += G) { const int i = e % k, j = e / k; if (i < j) S[e] = S[j + k * i]; }
        g.sync();
        if (wish_reg) {
#define BVG_WISH_REG(grp)                                                                                                  \
          switch (k) {                                                                                                     \
            case 1: wishart_block_reg<1>(grp, rng, sweep, m.wish_df_post, S, W4, kv, kd, sig, out_sig, fail); break;       \
            case 2: wishart_block_reg<2>(grp, rng, sweep, m.wish_df_post, S, W4, kv, kd, sig, out_sig, fail); break;       \
            case 3: wishart_block_reg<3>(grp, rng, sweep, m.wish_df_post, S, W4, kv, kd, sig, out_sig, fail); break;       \
            default: wishart_block_reg<4>(grp, rng, sweep, m.wish_df_post, S, W4, kv, kd, sig, out_sig, fail); break;      \
          }
#if defined(__CUDA_ARCH__)
          if constexpr (Grp::kIsBlock) {
            if (tid < 32) {                                        // warp 0 alone; the CTA meets at the sync below
              const TileGroup<32> wg;
              if (k == 5) wishart_block_reg<5>(wg, rng, sweep, m.wish_df_post, S, W4, kv, kd, sig, out_sig, fail);
              else if (k == 6) wishart_block_reg<6>(wg, rng, sweep, m.wish_df_post, S, W4, kv, kd, sig, out_sig, fail);
              else { BVG_WISH_REG(wg) }
            }
          } else
#endif
          { BVG_WISH_REG(g) }
#undef BVG_WISH_REG
        } else {
#if defined(__CUDA_ARCH__)
          if constexpr (Grp::kIsBlock) {
            if (tid < 32) {      // one warp, warp-level barriers only; the CTA meets again at the sync below
              const TileGroup<32> wg;
              wishart_block(wg, rng, sweep, k, m.wish_df_post, S, sig, W4, keep, W1, W2, W3, W4, kd, fail);
              if (keep) for (int e = tid; e < kk; e += 32) out_sig[e] = W4[e];
            }
          } else
#endif
          {
            wishart_block(g, rng, sweep, k, m.wish_df_post, S, sig, W4, keep, W1, W2, W3, W4 /*A then Sigma*/, kd, fail);
            if (keep) for (int e = tid; e < kk; e += G) out_sig[e] = W4[e];
          }
        }
      }
      g.sync();
    }

    // ---- store (bvaralg.cpp:518-559)
    if (keep && n > 0 && m.a0_from >= 0) {
      // structural model: the kept rows [a | b | c | a0 (variable j, equation i > j)] of the masked augmented system
      const int a0 = m.a0_from, n_out = a0 + (k * (k - 1)) / 2;
      double* oc = out.coef + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)n_out;
      double* ol = (out.lambda != nullptr) ? out.lambda + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)n_out
                                           : nullptr;
      for (int r = tid; r < n; r += G) {
        int o = r;
        if (r >= a0) {
          const int j = (r - a0) / k, i = (r - a0) % k;
          if (i <= j) continue;
          o = a0 + j * (k - 1) - (j * (j - 1)) / 2 + (i - j - 1);
        }
        oc[o] = a[r];
        if (ol != nullptr) ol[o] = lam[r];
      }
    } else if (keep && n > 0) {
      double* oc = out.coef + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)n;
      for (int r = tid; r < n; r += G) oc[r] = a[r];
      if (out.lambda != nullptr && varsel != VS_NONE) {
        const int lrows = n + ((pvs != VS_NONE) ? npsi : 0);
        double* ol = out.lambda + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)lrows;
        for (int r = tid; r < n; r += G) ol[r] = lam[r];
      }
    }
    if (keep && pvs != VS_NONE && out.lambda != nullptr) {                                // draws_lambda_a0, :530-532
      const int l0 = (varsel != VS_NONE) ? n : 0;
      double* ol = out.lambda + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)(l0 + npsi) + l0;
      for (int r = tid; r < npsi; r += G) ol[r] = plam[r];
    }
    g.sync();
  }

  // save state
  for (int e = tid; e < kk; e += G) st_sig[e] = sig[e];
  for (int e = tid; e < n; e += G) { st_vd[e] = vdiag[e]; st_lam[e] = lam[e]; }
  if (sv) {
    for (int e = tid; e < T * k; e += G) st_h[e] = h[e];
    for (int e = tid; e < k; e += G) { st_h[T * k + e] = sigma_h[e]; st_h[T * k + k + e] = h_init[e]; }
  }
  if (covar) {
    for (int e = tid; e < k; e += G) st_tail[e] = omd[e];
    for (int e = tid; e < npsi; e += G) { st_tail[k + e] = psi[e]; st_tail[k + npsi + e] = pvd[e]; st_tail[k + 2 * npsi + e] = plam[e]; }
  }
  if (fail && tid == 0) out.status[chain_local] = 1;
  g.sync();
}

}  // namespace bvg
