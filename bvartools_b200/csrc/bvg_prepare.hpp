// bvg_prepare.hpp -- host-side validation of a bvg_spec and assembly of the constant tables / initial chain
// state the kernels consume.  Plain C++ (no CUDA): the API (bvg_api.cu) uploads the result to the device; the
// test-only emulator (tests/emu) binds it to host memory.
//
// What is restated here is the set-up part of the reference samplers, src/bvaralg.cpp:9-249 and
// src/bvartvpalg.cpp:11-279 (unpacking, feature flags, illegal combinations, initial values).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/bvargpu.h"
#include "bvg_tvp_sweep.cuh"
#include "bvg_kron_sweep.cuh"

namespace bvg {

struct Prepared {
  int k = 0, T = 0, M = 0, n = 0, tvp = 0, sigma_type = 0, varsel = 0, resid_mode = 0, n_mix = 7;
  int iterations = 0, burnin = 0;
  double wish_df_post = 0.0;
  std::vector<double> blob;
  std::vector<unsigned char> include;
  std::vector<double> state0;
  int state_len = 0;
  // offsets (doubles) into blob; -1 = absent
  long Y = -1, X = -1, XX = -1, YX = -1, Aols = -1, SSE = -1, mu = -1, vi_off = -1, vmu_off = -1, vdiag0 = -1,
       tau0 = -1, tau1 = -1, inprior = -1, S0 = -1, gshape = -1, grate = -1, svmu = -1, svvi = -1, svconst = -1,
       mixp = -1, mixmu = -1, mixs2 = -1, omega_off = -1, vshape = -1, vrate = -1, kLinv = -1, kSS = -1,
       psi_vshape = -1, psi_vrate = -1, maskd = -1;
  int kron = 0;   // 1 = flat-prior Kronecker sweep (bvg_kron_sweep.cuh)
  int n_psi = 0;  // k (k - 1) / 2 when the covariance (psi) block is active, else 0
  // structural models (A0): the sampler runs on the regressors [x; -y] with the positions (variable j, equation i <= j)
  // of the -y block permanently masked; n_user = k M_user + k (k - 1) / 2 rows are stored, varsel_user is what the
  // caller asked for (the mask itself rides on the BVS machinery), a0_from = k M_user (first row of the A0 block)
  int structural = 0, n_user = 0, varsel_user = 0, a0_from = -1;
  int diag_sigma = 0;   // Sigma_t^-1 stays diagonal over the whole run (SV / gamma variances, diagonal start, no psi block)
  int psi_varsel = 0;
  long psi_mu = -1, psi_vi = -1, psi_vmu = -1, psi_tau0 = -1, psi_tau1 = -1, psi_inprior = -1;
  std::vector<unsigned char> psi_include;
  std::string error;

  long push(const double* p, size_t cnt) {
    const long off = (long)blob.size();
    blob.insert(blob.end(), p, p + cnt);
    return off;
  }
  long push(const std::vector<double>& v) { return push(v.data(), v.size()); }
};

static inline bool has_nan(const double* p, size_t cnt) {
  for (size_t i = 0; i < cnt; ++i) if (isnan(p[i])) return true;
  return false;
}

// Cholesky solve in long double: returns false if not positive definite.  A (m x m, col-major) is destroyed.
static inline bool chol_ld(std::vector<long double>& A, int m) {
  for (int j = 0; j < m; ++j) {
    long double d = A[j + (size_t)m * j];
    for (int c = 0; c < j; ++c) d -= A[j + (size_t)m * c] * A[j + (size_t)m * c];
    if (!(d > 0.0L)) return false;
    d = sqrtl(d);
    A[j + (size_t)m * j] = d;
    for (int i = j + 1; i < m; ++i) {
      long double s = A[i + (size_t)m * j];
      for (int c = 0; c < j; ++c) s -= A[i + (size_t)m * c] * A[j + (size_t)m * c];
      A[i + (size_t)m * j] = s / d;
    }
  }
  return true;
}

static const double KSC_P[7] = {0.00730, 0.10556, 0.00002, 0.04395, 0.34001, 0.24566, 0.25750};
static const double KSC_MU[7] = {-10.12999, -3.97281, -8.56686, 2.77786, 0.61942, 1.79518, -1.08819};
static const double KSC_S2[7] = {5.79596, 2.61369, 5.17950, 0.16735, 0.64009, 0.34023, 1.26261};
static const double OCSN_P[10] = {0.00609, 0.04775, 0.13057, 0.20674, 0.22715, 0.18842, 0.12047, 0.05591, 0.01575, 0.00115};
static const double OCSN_MU[10] = {1.92677, 1.34744, 0.73504, 0.02266, -0.85173, -1.97278, -3.46788, -5.55246, -8.68384, -14.65000};
static const double OCSN_S2[10] = {0.11265, 0.17788, 0.26768, 0.40611, 0.62699, 0.98583, 1.57469, 2.54498, 4.16591, 7.33342};

#define BVG_FAIL(code, msg) do { P.error = (msg); return (code); } while (0)

static inline int prepare_impl(const bvg_spec* s, Prepared& P);

// internal coefficient index of user row r (reference order [a | b | c | a0], a0 by variable j then equation i > j)
static inline int structural_pos(int r, int k, int Mu) {
  if (r < k * Mu) return r;
  int q = r - k * Mu;
  for (int j = 0; j < k; ++j) {
    if (q < k - 1 - j) return k * (Mu + j) + j + 1 + q;
    q -= k - 1 - j;
  }
  return -1;
}

// Structural models (model$structural, R/gen_var.R:243-251; the A0 columns of data$SUR): equation i also has the
// regressors -y_jt, j < i.  z = [kron(x', I_k) | y_A0] is the column selection of kron([x; -y]', I_k) that drops
// (variable j, equation i <= j), so the reference's posterior precision is the kept block of the precision of the
// augmented VAR with those positions masked; masked positions decouple (prior identity) and are zeroed after the draw --
// exactly what BVS does with lambda = 0 on positions that are never re-drawn.  The model is therefore rewritten as a
// BVS model on M + k regressors (empty `include` when the caller uses no selection) and prepared as usual.
static inline int prepare(const bvg_spec* s, Prepared& P) {
  if (s == nullptr) BVG_FAIL(-1, "spec is NULL");
  if (!s->structural) return prepare_impl(s, P);
  const int k = s->k, T = s->T, Mu = s->M;
  if (k < 2) BVG_FAIL(-7, "Model must contain at least two endogenous variables for structural analysis.");
  if (k > 64 || T < 2 || Mu < 0 || !s->Y || (Mu > 0 && !s->X)) BVG_FAIL(-3, "invalid dimensions or missing data");
  if (s->tvp && (!s->tvp_shape || !s->tvp_rate || !s->a_init)) BVG_FAIL(-4, "TVP prior shape / rate or initial draw missing");
  if (s->varsel == BVG_VARSEL_SSVS && s->tvp) BVG_FAIL(-7, "SSVS is not defined for TVP models");
  if (s->sigma_type == BVG_SIGMA_WISHART) BVG_FAIL(-7, "Structural models may not use a Wishart prior.");
  if (s->psi_mu || s->psi_vi) BVG_FAIL(-7, "Error covariances and structural coefficients cannot be estimated at the same time.");
  if (!s->mu || !s->Vi) BVG_FAIL(-4, "priors$coefficients$mu / v_i missing");
  const int Ma = Mu + k, na = k * Ma, nu = k * Mu + k * (k - 1) / 2;
  std::vector<double> X((size_t)Ma * T), mu((size_t)na, 0.0), inprior((size_t)na, 0.5);
  for (int t = 0; t < T; ++t) {
    for (int j = 0; j < Mu; ++j) X[j + (size_t)Ma * t] = s->X[j + (size_t)Mu * t];
    for (int j = 0; j < k; ++j) X[Mu + j + (size_t)Ma * t] = -s->Y[j + (size_t)k * t];
  }
  std::vector<int> pos((size_t)nu);
  for (int r = 0; r < nu; ++r) pos[r] = structural_pos(r, k, Mu);
  std::vector<double> Vi((size_t)na * na, 0.0);
  for (int r = 0; r < na; ++r) Vi[r + (size_t)na * r] = 1.0;                 // masked positions: N(0, 1), never used
  for (int r = 0; r < nu; ++r) {
    mu[pos[r]] = s->mu[r];
    if (s->vi_dense) {
      for (int c = 0; c < nu; ++c) Vi[pos[r] + (size_t)na * pos[c]] = s->Vi[r + (size_t)nu * c];
    } else {
      Vi[pos[r] + (size_t)na * pos[r]] = s->Vi[r];
    }
  }
  std::vector<int32_t> include;
  std::vector<double> tau0, tau1;
  if (s->varsel == BVG_VARSEL_SSVS) {
    // SSVS on the kept positions (bvaralg.cpp:107-135 with the structural z): the static mask cannot ride on lambda (SSVS
    // records its indicators there and never masks the regressors), it travels as BvarModel::maskd instead
    if (!s->inprior || !s->tau0 || !s->tau1 || s->n_include < 0 || (s->n_include > 0 && !s->include))
      BVG_FAIL(-4, "SSVS prior (tau0 / tau1 / inprior / include) missing");
    tau0.assign((size_t)na, 1.0); tau1.assign((size_t)na, 1.0);
    for (int r = 0; r < nu; ++r) { inprior[pos[r]] = s->inprior[r]; tau0[pos[r]] = s->tau0[r]; tau1[pos[r]] = s->tau1[r]; }
    for (int e = 0; e < s->n_include; ++e) {
      if (s->include[e] < 0 || s->include[e] >= nu) BVG_FAIL(-8, "include index out of range");
      include.push_back(pos[s->include[e]]);
    }
  }
  if (s->varsel == BVG_VARSEL_BVS) {
    if (!s->inprior || s->n_include < 0 || (s->n_include > 0 && !s->include)) BVG_FAIL(-4, "inclusion prior / include index missing");
    for (int r = 0; r < nu; ++r) inprior[pos[r]] = s->inprior[r];
    for (int e = 0; e < s->n_include; ++e) {
      if (s->include[e] < 0 || s->include[e] >= nu) BVG_FAIL(-8, "include index out of range");
      include.push_back(pos[s->include[e]]);
    }
  }
  std::vector<double> tshape, trate, ainit;
  if (s->tvp) {                                  // state-variance priors and initial draw (masked positions: any valid value)
    tshape.assign((size_t)na, 1.0); trate.assign((size_t)na, 1.0); ainit.assign((size_t)na, 0.0);
    for (int r = 0; r < nu; ++r) { tshape[pos[r]] = s->tvp_shape[r]; trate[pos[r]] = s->tvp_rate[r]; ainit[pos[r]] = s->a_init[r]; }
  }
  bvg_spec eff = *s;
  eff.structural = 0;
  if (s->tvp) { eff.tvp_shape = tshape.data(); eff.tvp_rate = trate.data(); eff.a_init = ainit.data(); }
  eff.M = Ma; eff.X = X.data(); eff.mu = mu.data(); eff.Vi = Vi.data(); eff.vi_dense = 1;
  const bool ssvs = s->varsel == BVG_VARSEL_SSVS;
  eff.varsel = ssvs ? BVG_VARSEL_SSVS : BVG_VARSEL_BVS; eff.inprior = inprior.data();
  if (ssvs) { eff.tau0 = tau0.data(); eff.tau1 = tau1.data(); }
  eff.include = include.empty() ? nullptr : include.data(); eff.n_include = (int32_t)include.size();
  eff.resid_mode = BVG_RESID_DIRECT;                                          // [x; -y] fits y exactly: no OLS moments
  const int rc = prepare_impl(&eff, P);
  if (rc) return rc;
  P.structural = 1; P.n_user = nu; P.varsel_user = s->varsel; P.a0_from = k * Mu;
  if (ssvs) {
    std::vector<double> mk((size_t)na, 0.0);
    for (int r = 0; r < nu; ++r) mk[pos[r]] = 1.0;
    P.maskd = P.push(mk);
  }
  // lambda = 0 on the masked positions (state: Sigma^-1 (k k) | prior diagonal (n) | lambda (n) | ...;
  // TVP state: Sigma^-1 | state precisions (n) | initial state (n) | lambda (n) | ...)
  std::vector<unsigned char> kept((size_t)na, 0);
  for (int r = 0; r < nu; ++r) kept[pos[r]] = 1;
  const size_t lam_off = (size_t)k * k + (size_t)(s->tvp ? 2 : 1) * na;
  for (int r = 0; r < na; ++r) if (!kept[r]) P.state0[lam_off + r] = 0.0;
  return 0;
}

static inline int prepare_impl(const bvg_spec* s, Prepared& P) {
  if (s == nullptr) BVG_FAIL(-1, "spec is NULL");
  if (s->abi_version != BVG_ABI_VERSION) BVG_FAIL(-2, "bvg_spec.abi_version does not match this library");
  if (s->k < 1 || s->T < 2 || s->M < 0) BVG_FAIL(-3, "invalid dimensions (need k >= 1, T >= 2, M >= 0)");
  if (s->k > 64) BVG_FAIL(-3, "k > 64 is not supported");
  if (s->iterations < 1 || s->burnin < 0) BVG_FAIL(-3, "iterations must be >= 1 and burnin >= 0");
  if (s->n_chains < 1) BVG_FAIL(-3, "n_chains must be >= 1");
  if (s->Y == nullptr) BVG_FAIL(-4, "Y is NULL");
  if (s->M > 0 && s->X == nullptr) BVG_FAIL(-4, "X is NULL");
  const int k = s->k, T = s->T, M = s->M, n = k * M, kk = k * k;
  P.k = k; P.T = T; P.M = M; P.n = n; P.tvp = s->tvp ? 1 : 0;
  P.sigma_type = s->sigma_type; P.varsel = s->varsel;
  P.iterations = s->iterations; P.burnin = s->burnin;
  if (P.sigma_type < 0 || P.sigma_type > 2) BVG_FAIL(-5, "invalid sigma_type");
  if (P.varsel < 0 || P.varsel > 2) BVG_FAIL(-5, "invalid varsel");
  if (has_nan(s->Y, (size_t)k * T)) BVG_FAIL(-6, "Argument 'y' contains NAs.");
  if (M > 0 && has_nan(s->X, (size_t)M * T)) BVG_FAIL(-6, "Argument 'x' contains NAs.");
  // illegal combinations (bvaralg.cpp:116-118, R/ssvs_prior.R:39-45)
  if (P.varsel == BVG_VARSEL_SSVS && P.sigma_type == BVG_SIGMA_SV)
    BVG_FAIL(-7, "Not allowed to use SSVS with stochastic volatility.");
  if (P.varsel == BVG_VARSEL_SSVS && P.tvp) BVG_FAIL(-7, "SSVS cannot be used with models with time varying parameter.");
  if (P.tvp && n == 0) BVG_FAIL(-7, "TVP model without regressors");
  if (n == 0 && P.varsel != BVG_VARSEL_NONE) BVG_FAIL(-7, "variable selection needs regressors");

  P.Y = P.push(s->Y, (size_t)k * T);
  if (M > 0) P.X = P.push(s->X, (size_t)M * T);

  // ---- coefficient prior
  if (n > 0) {
    if (s->mu == nullptr || s->Vi == nullptr) BVG_FAIL(-4, "priors$coefficients$mu / v_i missing");
    if (has_nan(s->mu, n)) BVG_FAIL(-6, "prior mean contains NAs.");
    P.mu = P.push(s->mu, n);
    std::vector<double> vd(n);
    bool off_nonzero = false;
    if (s->vi_dense) {
      if (has_nan(s->Vi, (size_t)n * n)) BVG_FAIL(-6, "prior precision contains NAs.");
      for (int r = 0; r < n; ++r) vd[r] = s->Vi[r + (size_t)n * r];
      for (int c = 0; c < n && !off_nonzero; ++c)
        for (int r = 0; r < n; ++r)
          if (r != c && s->Vi[r + (size_t)n * c] != 0.0) { off_nonzero = true; break; }
    } else {
      if (has_nan(s->Vi, n)) BVG_FAIL(-6, "prior precision contains NAs.");
      for (int r = 0; r < n; ++r) vd[r] = s->Vi[r];
    }
    P.vdiag0 = P.push(vd);
    if (off_nonzero) {
      std::vector<double> vo((size_t)tri(n)), vm(n, 0.0);
      for (int r = 0; r < n; ++r)
        for (int c = 0; c <= r; ++c) vo[tri(r) + c] = (r == c) ? 0.0 : s->Vi[r + (size_t)n * c];
      for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c)
          if (r != c) vm[r] += s->Vi[r + (size_t)n * c] * s->mu[c];
      P.vi_off = P.push(vo);
      P.vmu_off = P.push(vm);
    }
    P.include.assign(n, 0);
    if (P.varsel != BVG_VARSEL_NONE) {
      if (s->inprior == nullptr) BVG_FAIL(-4, "inclusion prior missing");
      if (s->n_include < 0 || (s->n_include > 0 && s->include == nullptr)) BVG_FAIL(-4, "include index missing");
      for (int i = 0; i < s->n_include; ++i) {
        if (s->include[i] < 0 || s->include[i] >= n) BVG_FAIL(-8, "include index out of range");
        P.include[s->include[i]] = 1;
      }
      P.inprior = P.push(s->inprior, n);
      if (P.varsel == BVG_VARSEL_SSVS) {
        if (s->tau0 == nullptr || s->tau1 == nullptr) BVG_FAIL(-4, "SSVS tau0 / tau1 missing");
        P.tau0 = P.push(s->tau0, n);
        P.tau1 = P.push(s->tau1, n);
      }
    }
  }

  // ---- error variance prior + initial values
  std::vector<double> sig0((size_t)kk, 0.0), omega_off((size_t)kk, 0.0);
  if (P.sigma_type == BVG_SIGMA_SV) {
    if (!s->sv_mu || !s->sv_vi || !s->sigma_shape || !s->sigma_rate || !s->sv_sigma_h || !s->sv_constant || !s->sv_h)
      BVG_FAIL(-4, "Missing prior specifications for stochastic volatility prior.");
    P.n_mix = (s->sv_mixture == BVG_MIX_OCSN2007) ? 10 : 7;
    P.diag_sigma = 1;
    std::vector<double> gs(k);
    for (int i = 0; i < k; ++i) gs[i] = s->sigma_shape[i] + 0.5 * T;               // bvaralg.cpp:223
    P.gshape = P.push(gs);
    P.grate = P.push(s->sigma_rate, k);
    P.svmu = P.push(s->sv_mu, k);
    P.svvi = P.push(s->sv_vi, kk);
    P.svconst = P.push(s->sv_constant, k);
    if (P.n_mix == 7) {
      std::vector<double> mu7(7);
      for (int j = 0; j < 7; ++j) mu7[j] = KSC_MU[j] - 1.2704;                     // stochvol_ksc1998.cpp:77
      P.mixp = P.push(KSC_P, 7); P.mixmu = P.push(mu7); P.mixs2 = P.push(KSC_S2, 7);
    } else {
      P.mixp = P.push(OCSN_P, 10); P.mixmu = P.push(OCSN_MU, 10); P.mixs2 = P.push(OCSN_S2, 10);
    }
  } else {
    if (s->sigma_i == nullptr) BVG_FAIL(-4, "initial$sigma$sigma_i missing");
    if (has_nan(s->sigma_i, kk)) BVG_FAIL(-6, "Argument 'sigma_i' contains NAs.");
    for (int i = 0; i < k; ++i) sig0[i + (size_t)k * i] = s->sigma_i[i + (size_t)k * i];   // bvaralg.cpp:246 (diagonal only)
    if (P.sigma_type == BVG_SIGMA_WISHART) {
      if (s->wishart_scale == nullptr) BVG_FAIL(-4, "priors$sigma$scale missing");
      P.wish_df_post = s->wishart_df + T;                                          // bvaralg.cpp:234
      if (!(P.wish_df_post > k - 1)) BVG_FAIL(-5, "posterior degrees of freedom must exceed k - 1");
      P.S0 = P.push(s->wishart_scale, kk);
    } else {
      if (!s->sigma_shape || !s->sigma_rate) BVG_FAIL(-4, "priors$sigma$shape / rate missing");
      std::vector<double> gs(k);
      for (int i = 0; i < k; ++i) gs[i] = s->sigma_shape[i] + 0.5 * T;             // bvaralg.cpp:239
      P.gshape = P.push(gs);
      P.grate = P.push(s->sigma_rate, k);
      for (int e = 0; e < kk; ++e) omega_off[e] = (e % k == e / k) ? 0.0 : s->sigma_i[e];
      P.omega_off = P.push(omega_off);
      P.diag_sigma = 1;
      for (int e = 0; e < kk; ++e) if (omega_off[e] != 0.0) P.diag_sigma = 0;
    }
  }

  // ---- covariance (psi) block (bvaralg.cpp:157-204): prior mean / precision of the k (k - 1) / 2 free elements
  if (s->psi_mu != nullptr || s->psi_vi != nullptr) {
    if (!s->psi_mu || !s->psi_vi) BVG_FAIL(-4, "priors$psi needs both mu and v_i");
    if (P.sigma_type == BVG_SIGMA_WISHART) BVG_FAIL(-5, "the covariance (psi) block needs a gamma or SV error prior");
    if (k < 2) BVG_FAIL(-5, "the covariance (psi) block needs at least two endogenous variables");
    P.n_psi = k * (k - 1) / 2;
    if (has_nan(s->psi_mu, P.n_psi) || has_nan(s->psi_vi, (size_t)P.n_psi * P.n_psi)) BVG_FAIL(-6, "priors$psi contains NAs.");
    P.psi_mu = P.push(s->psi_mu, P.n_psi);
    P.psi_vi = P.push(s->psi_vi, (size_t)P.n_psi * P.n_psi);
    std::vector<double> vm((size_t)P.n_psi, 0.0);       // OFF-diagonal part of psi_prior_vi * psi_prior_mu (:393); the
    for (int r = 0; r < P.n_psi; ++r)                   // diagonal of the prior precision lives in the chain state (SSVS)
      for (int c = 0; c < P.n_psi; ++c)
        if (c != r) vm[r] += s->psi_vi[r + (size_t)P.n_psi * c] * s->psi_mu[c];
    P.psi_vmu = P.push(vm);
    P.psi_varsel = s->psi_varsel;
    if (P.psi_varsel != BVG_VARSEL_NONE && P.psi_varsel != BVG_VARSEL_SSVS && P.psi_varsel != BVG_VARSEL_BVS)
      BVG_FAIL(-5, "invalid psi_varsel");
    if (P.psi_varsel != BVG_VARSEL_NONE) {
      if (P.psi_varsel == BVG_VARSEL_SSVS && P.sigma_type == BVG_SIGMA_SV)
        BVG_FAIL(-11, "Not allowed to use SSVS with stochastic volatility.");      // bvaralg.cpp:173-175
      if (!s->psi_inprior || !s->psi_include || s->psi_n_include < 0) BVG_FAIL(-4, "priors$psi variable selection: inprior / include missing");
      if (P.psi_varsel == BVG_VARSEL_SSVS && (!s->psi_tau0 || !s->psi_tau1)) BVG_FAIL(-4, "priors$psi$ssvs: tau0 / tau1 missing");
      P.psi_inprior = P.push(s->psi_inprior, P.n_psi);
      if (P.psi_varsel == BVG_VARSEL_SSVS) { P.psi_tau0 = P.push(s->psi_tau0, P.n_psi); P.psi_tau1 = P.push(s->psi_tau1, P.n_psi); }
      P.psi_include.assign((size_t)P.n_psi, 0);
      for (int e = 0; e < s->psi_n_include; ++e) {
        const int pos = s->psi_include[e];
        if (pos < 0 || pos >= P.n_psi) BVG_FAIL(-5, "priors$psi: include position out of range");
        P.psi_include[pos] = 1;
      }
      if ((int)P.include.size() < n) P.include.resize((size_t)n, 0);              // device mask: [coefficients (n) | psi (n_psi)]
      P.include.insert(P.include.end(), P.psi_include.begin(), P.psi_include.end());
    }
    if (P.tvp) {
      // time-varying covariances (bvartvpalg.cpp:143-176): random-walk states with inverse-gamma variances
      if (P.psi_varsel == BVG_VARSEL_SSVS) BVG_FAIL(-7, "SSVS cannot be used with models with time varying parameter.");
      if (k > BVG_TVP_PSI_MAXK) BVG_FAIL(-5, "the covariance block of a TVP model supports at most 8 endogenous variables");
      if (P.n_psi > n) BVG_FAIL(-5, "the covariance block of a TVP model needs k (k - 1) / 2 <= number of coefficients");
      if (!s->psi_shape || !s->psi_rate || !s->psi_init) BVG_FAIL(-4, "priors$psi shape / rate or initial$psi$draw missing");
      if (has_nan(s->psi_shape, P.n_psi) || has_nan(s->psi_rate, P.n_psi) || has_nan(s->psi_init, P.n_psi))
        BVG_FAIL(-6, "priors$psi contains NAs.");
      std::vector<double> ps((size_t)P.n_psi);
      for (int r = 0; r < P.n_psi; ++r) {
        ps[r] = s->psi_shape[r] + 0.5 * T;                                           // bvartvpalg.cpp:149
        if (!(s->psi_rate[r] > 0.0)) BVG_FAIL(-5, "priors$psi$rate must be positive");
      }
      P.psi_vshape = P.push(ps);
      P.psi_vrate = P.push(s->psi_rate, P.n_psi);
    }
  }

  // ---- residual mode and cross moments
  int rm = s->resid_mode;
  if (P.n_psi > 0) {
    if (rm == BVG_RESID_MOMENT) BVG_FAIL(-5, "the covariance (psi) block needs the residuals themselves (direct mode)");
    rm = BVG_RESID_DIRECT;
  }
  if (rm != BVG_RESID_AUTO && rm != BVG_RESID_DIRECT && rm != BVG_RESID_MOMENT) BVG_FAIL(-5, "invalid resid_mode");
  if (P.tvp || P.sigma_type == BVG_SIGMA_SV || n == 0) {
    if (rm == BVG_RESID_MOMENT) BVG_FAIL(-5, "moment residual mode needs a constant-parameter, constant-variance VAR");
    rm = BVG_RESID_DIRECT;
  }
  if (M > 0) {
    std::vector<double> XX((size_t)M * M, 0.0), YX((size_t)k * M, 0.0);
    for (int t = 0; t < T; ++t) {
      const double* xt = s->X + (size_t)M * t;
      const double* yt = s->Y + (size_t)k * t;
      for (int j = 0; j < M; ++j) {
        for (int j2 = 0; j2 < M; ++j2) XX[j + (size_t)M * j2] += xt[j] * xt[j2];
        for (int i = 0; i < k; ++i) YX[i + (size_t)k * j] += yt[i] * xt[j];
      }
    }
    P.XX = P.push(XX);
    P.YX = P.push(YX);
    if (rm != BVG_RESID_DIRECT) {
      bool ok = T > M;
      std::vector<long double> C((size_t)M * M);
      std::vector<double> A((size_t)k * M, 0.0), SSE((size_t)kk, 0.0);
      if (ok) {
        for (size_t e = 0; e < C.size(); ++e) C[e] = XX[e];
        ok = chol_ld(C, M);
        if (ok) {   // reject numerically rank-deficient X (the identity needs the exact OLS projection)
          long double dmin = C[0], dmax = C[0];
          for (int j = 0; j < M; ++j) { const long double d = C[j + (size_t)M * j]; if (d < dmin) dmin = d; if (d > dmax) dmax = d; }
          if (dmin < 1e-7L * dmax) ok = false;
        }
      }
      if (ok) {
        // A_ols = YX (XX)^-1 : solve per row i in long double
        for (int i = 0; i < k; ++i) {
          std::vector<long double> b(M);
          for (int j = 0; j < M; ++j) b[j] = YX[i + (size_t)k * j];
          for (int j = 0; j < M; ++j) {
            long double v = b[j];
            for (int c = 0; c < j; ++c) v -= C[j + (size_t)M * c] * b[c];
            b[j] = v / C[j + (size_t)M * j];
          }
          for (int j = M - 1; j >= 0; --j) {
            long double v = b[j];
            for (int c = j + 1; c < M; ++c) v -= C[c + (size_t)M * j] * b[c];
            b[j] = v / C[j + (size_t)M * j];
          }
          for (int j = 0; j < M; ++j) A[i + (size_t)k * j] = (double)b[j];
        }
        std::vector<long double> sse((size_t)kk, 0.0L), u(k);
        for (int t = 0; t < T; ++t) {
          for (int i = 0; i < k; ++i) {
            long double v = s->Y[i + (size_t)k * t];
            for (int j = 0; j < M; ++j) v -= (long double)A[i + (size_t)k * j] * s->X[j + (size_t)M * t];
            u[i] = v;
          }
          for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) sse[i + (size_t)k * j] += u[i] * u[j];
        }
        for (int e = 0; e < kk; ++e) SSE[e] = (double)sse[e];
        P.Aols = P.push(A);
        P.SSE = P.push(SSE);
        // flat-prior Kronecker specialisation: V^-1 = 0, Wishart prior, no variable selection, k <= 4, and the
        // caller left the residual mode to the library (an explicit mode selects the general kernel)
        bool flat = (s->resid_mode == BVG_RESID_AUTO) && !P.tvp && P.sigma_type == BVG_SIGMA_WISHART &&
                    P.varsel == BVG_VARSEL_NONE && P.vi_off < 0 && k <= 4;
        for (int r = 0; r < n && flat; ++r) if (P.blob[P.vdiag0 + r] != 0.0) flat = false;
        if (flat) {
          // L_X^-1 in long double from the factor C (lower, col-major) computed above
          std::vector<long double> Li((size_t)M * M, 0.0L);
          for (int c = 0; c < M; ++c) {
            Li[c + (size_t)M * c] = 1.0L / C[c + (size_t)M * c];
            for (int i = c + 1; i < M; ++i) {
              long double v = 0.0L;
              for (int mm = c; mm < i; ++mm) v += C[i + (size_t)M * mm] * Li[mm + (size_t)M * c];
              Li[i + (size_t)M * c] = -v / C[i + (size_t)M * i];
            }
          }
          std::vector<double> lp((size_t)tri(M)), ss((size_t)kk);
          for (int i = 0; i < M; ++i)
            for (int c = 0; c <= i; ++c) lp[tri(i) + c] = (double)Li[i + (size_t)M * c];
          for (int e = 0; e < kk; ++e) ss[e] = (double)(sse[e] + (long double)s->wishart_scale[e]);
          P.kLinv = P.push(lp);
          P.kSS = P.push(ss);
          P.kron = 1;
        }
        rm = BVG_RESID_MOMENT;
      } else {
        if (rm == BVG_RESID_MOMENT) BVG_FAIL(-9, "moment residual mode needs T > M and X of full row rank");
        rm = BVG_RESID_DIRECT;
      }
    }
  }
  P.resid_mode = rm;

  // ---- TVP
  if (P.tvp) {
    if (!s->tvp_shape || !s->tvp_rate || !s->a_init) BVG_FAIL(-4, "TVP prior shape / rate or initial draw missing");
    std::vector<double> vs(n);
    for (int r = 0; r < n; ++r) {
      vs[r] = s->tvp_shape[r] + 0.5 * T;                                            // bvartvpalg.cpp:104
      if (!(s->tvp_rate[r] > 0.0)) BVG_FAIL(-5, "TVP prior rate must be positive");
    }
    P.vshape = P.push(vs);
    P.vrate = P.push(s->tvp_rate, n);
  }

  // ---- initial chain state
  if (P.tvp) {
    P.state_len = tvp_state_len(k, T, n, P.sigma_type, P.n_psi);
    P.state0.assign(P.state_len, 0.0);
    double* st = P.state0.data();
    for (int e = 0; e < kk; ++e) st[e] = sig0[e];
    for (int r = 0; r < n; ++r) { st[kk + r] = 1.0 / s->tvp_rate[r]; st[kk + n + r] = s->a_init[r]; st[kk + 2 * n + r] = 1.0; }
    double* sh = st + kk + 3 * n;
    if (P.sigma_type == BVG_SIGMA_SV) {
      for (int e = 0; e < T * k; ++e) sh[e] = s->sv_h[e];
      for (int i = 0; i < k; ++i) { sh[T * k + i] = s->sv_sigma_h[i]; sh[T * k + k + i] = s->sv_h[0 + (size_t)T * i]; }
      sh += T * k + 2 * k;
    }
    for (int r = 0; r < P.n_psi; ++r) { sh[r] = 1.0 / s->psi_rate[r]; sh[P.n_psi + r] = s->psi_init[r]; sh[2 * P.n_psi + r] = 1.0; }   // :175-176,169,180
  } else {
    P.state_len = bvar_state_len(k, T, n, P.sigma_type, P.n_psi);
    P.state0.assign(P.state_len, 0.0);
    double* st = P.state0.data();
    for (int e = 0; e < kk; ++e) st[e] = sig0[e];
    for (int r = 0; r < n; ++r) { st[kk + r] = P.blob[P.vdiag0 + r]; st[kk + n + r] = 1.0; }
    if (P.n_psi > 0) {   // tail of the state: diagonal of Omega^-1 (gamma prior) | psi = 0 (Psi = I, bvaralg.cpp:187) |
      //                      diagonal of the psi prior precision (changed by SSVS) | psi inclusion indicators (all 1)
      double* tail = st + P.state_len - (k + 3 * P.n_psi);
      for (int i = 0; i < k; ++i) tail[i] = sig0[i + (size_t)k * i];
      for (int r = 0; r < P.n_psi; ++r) {
        tail[k + P.n_psi + r] = s->psi_vi[r + (size_t)P.n_psi * r];
        tail[k + 2 * P.n_psi + r] = 1.0;
      }
    }
    if (P.kron)   // the Kronecker sweep keeps B^-1 = chol(Sigma^-1)^-1 (upper) in the Sigma^-1 slot; the initial Sigma^-1 is diagonal
      for (int i = 0; i < k; ++i) st[i + (size_t)k * i] = 1.0 / sqrt(sig0[i + (size_t)k * i]);
    if (P.sigma_type == BVG_SIGMA_SV) {
      double* sh = st + kk + 2 * n;
      for (int e = 0; e < T * k; ++e) sh[e] = s->sv_h[e];
      for (int i = 0; i < k; ++i) { sh[T * k + i] = s->sv_sigma_h[i]; sh[T * k + k + i] = s->sv_h[0 + (size_t)T * i]; }
    }
  }
  return 0;
}

static inline const double* at(const double* base, long off) { return off < 0 ? nullptr : base + off; }

static inline void bind_bvar(const Prepared& P, const double* base, const unsigned char* inc, BvarModel* m) {
  m->k = P.k; m->T = P.T; m->M = P.M; m->n = P.n;
  m->a0_from = P.a0_from;
  m->maskd = at(base, P.maskd);
  m->diag_sigma = (P.diag_sigma && P.n_psi == 0) ? 1 : 0;
  m->sigma_type = P.sigma_type; m->varsel = P.varsel; m->resid_mode = P.resid_mode; m->n_mix = P.n_mix;
  m->Y = at(base, P.Y); m->X = at(base, P.X); m->XX = at(base, P.XX); m->YX = at(base, P.YX);
  m->Aols = at(base, P.Aols); m->SSE = at(base, P.SSE); m->mu = at(base, P.mu);
  m->vi_off = at(base, P.vi_off); m->vmu_off = at(base, P.vmu_off); m->vdiag0 = at(base, P.vdiag0);
  m->tau0 = at(base, P.tau0); m->tau1 = at(base, P.tau1); m->inprior = at(base, P.inprior);
  m->include = inc;
  m->wish_df_post = P.wish_df_post;
  m->S0 = at(base, P.S0); m->g_shape_post = at(base, P.gshape); m->g_rate = at(base, P.grate);
  m->sv_mu = at(base, P.svmu); m->sv_vi = at(base, P.svvi); m->sv_const = at(base, P.svconst);
  m->mix_p = at(base, P.mixp); m->mix_mu = at(base, P.mixmu); m->mix_s2 = at(base, P.mixs2);
  m->omega_off = at(base, P.omega_off);
  m->n_psi = P.n_psi; m->psi_vi = at(base, P.psi_vi); m->psi_vmu = at(base, P.psi_vmu); m->psi_mu = at(base, P.psi_mu);
  m->psi_varsel = P.psi_varsel; m->psi_tau0 = at(base, P.psi_tau0); m->psi_tau1 = at(base, P.psi_tau1);
  m->psi_inprior = at(base, P.psi_inprior);
  m->psi_include = (inc != nullptr && P.psi_varsel != BVG_VARSEL_NONE) ? inc + P.n : nullptr;
  m->kron_Linv = P.kron ? at(base, P.kLinv) : nullptr;
  m->kron_SS = P.kron ? at(base, P.kSS) : nullptr;
}

static inline void bind_tvp(const Prepared& P, const double* base, const unsigned char* inc, TvpModel* m) {
  bind_bvar(P, base, inc, &m->b);
  m->v_shape_post = at(base, P.vshape);
  m->v_rate = at(base, P.vrate);
  m->psi_shape_post = at(base, P.psi_vshape);
  m->psi_rate = at(base, P.psi_vrate);
}

}  // namespace bvg
