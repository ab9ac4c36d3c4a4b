// bvg_glue.h -- shared part of the Rcpp glue: `bvarmodel` list <-> bvg_spec / bvg_out (include/bvargpu.h).
//
// What src/bvaralg.cpp:9-289 and src/bvartvpalg.cpp:11-279 of franzmohr/bvartools unpack by name, packed into the flat
// POD of the C ABI; and the slicing of the draws_* matrices into posteriors${a,b,c,a0,sigma} of
// src/bvaralg.cpp:562-627 / src/bvartvpalg.cpp:555-626.  Pure translation: no arithmetic happens here.
//
// Build inside the package (src/): add  PKG_LIBS += -L<prefix>/lib -lbvargpu  and  PKG_CPPFLAGS += -I<prefix>/include
// to src/Makevars.  The glue needs Rcpp only (no Armadillo).
#ifndef BVG_GLUE_H
#define BVG_GLUE_H

#include <Rcpp.h>
#include <algorithm>
#include <string>
#include <vector>
#include "bvargpu.h"

namespace bvgglue {

inline bool has_name(const Rcpp::List& l, const char* nm) {
  if (Rf_isNull(l.names())) return false;
  Rcpp::CharacterVector names = l.names();
  for (R_xlen_t i = 0; i < names.size(); ++i)
    if (std::string(names[i]) == nm) return true;
  return false;
}

// Owns every buffer a bvg_spec points to.
struct SpecHolder {
  bvg_spec c;
  std::vector<std::vector<double> > dbl;
  std::vector<std::vector<int32_t> > i32;
  int k, T, M, p, n_a, n_b, n_c, n_a0, n_tot;
  bool sv, tvp, structural, varsel, psi_varsel, covar;

  SpecHolder() : k(0), T(0), M(0), p(0), n_a(0), n_b(0), n_c(0), n_a0(0), n_tot(0), sv(false), tvp(false),
                 structural(false), varsel(false), psi_varsel(false), covar(false) { c = bvg_spec(); }

  const double* keep(const Rcpp::NumericVector& v) {            // R vectors / matrices are column-major already
    dbl.push_back(std::vector<double>(v.begin(), v.end()));
    return dbl.back().data();
  }
  const double* keep_transposed(const Rcpp::NumericMatrix& m) { // t(m), column-major
    std::vector<double> t((size_t)m.nrow() * m.ncol());
    for (int r = 0; r < m.nrow(); ++r)
      for (int cc = 0; cc < m.ncol(); ++cc) t[(size_t)cc + (size_t)m.ncol() * r] = m(r, cc);
    dbl.push_back(t);
    return dbl.back().data();
  }
  const int32_t* keep_index(const Rcpp::NumericVector& one_based, int32_t* count) {   // include - 1 (bvaralg.cpp:132)
    std::vector<int32_t> z((size_t)one_based.size());
    for (R_xlen_t i = 0; i < one_based.size(); ++i) z[(size_t)i] = (int32_t)one_based[i] - 1;
    *count = (int32_t)z.size();
    i32.push_back(z);
    return i32.back().data();
  }
};

// `object`: the bvarmodel list of gen_var() + add_priors() (R/gen_var.R:288-303, R/add_priors.bvarmodel.R:426-613)
inline void fill_spec(const Rcpp::List& object, SpecHolder& H, long long n_chains, long long chain_offset,
                      unsigned long long seed, int n_gpus, int device) {
  Rcpp::List data = object["data"], model = object["model"], priors = object["priors"], initial = object["initial"];
  Rcpp::NumericMatrix Y = Rcpp::as<Rcpp::NumericMatrix>(data["Y"]);            // T x k
  bvg_spec& c = H.c;
  c.abi_version = BVG_ABI_VERSION;
  c.device = device;
  c.n_gpus = n_gpus;
  H.T = c.T = Y.nrow();
  H.k = c.k = Y.ncol();
  H.tvp = Rcpp::as<bool>(model["tvp"]);
  H.sv = Rcpp::as<bool>(model["sv"]);
  H.structural = Rcpp::as<bool>(model["structural"]) && H.k > 1;
  c.tvp = H.tvp ? 1 : 0;
  c.structural = H.structural ? 1 : 0;
  c.iterations = Rcpp::as<int>(model["iterations"]);
  c.burnin = Rcpp::as<int>(model["burnin"]);
  c.n_chains = n_chains;
  c.chain_offset = chain_offset;
  c.seed = seed;
  c.Y = H.keep_transposed(Y);                                                   // y = t(Y), bvaralg.cpp:10
  // regressors: data$Z (T x M); SUR = kron(Z, I_k) plus the A0 columns (R/gen_var.R:281-286) is never materialised
  H.M = 0;
  if (!Rf_isNull(data["Z"])) {
    Rcpp::NumericMatrix Z = Rcpp::as<Rcpp::NumericMatrix>(data["Z"]);
    H.M = c.M = Z.ncol();
    c.X = H.keep_transposed(Z);
  }
  // blocks of the coefficient vector (bvaralg.cpp:28-72,256-263)
  Rcpp::List endogen = model["endogen"];
  H.p = Rcpp::as<int>(endogen["lags"]);
  H.n_a = H.k * H.k * H.p;
  if (has_name(model, "exogen")) {
    Rcpp::List exogen = model["exogen"];
    const int m = Rcpp::as<Rcpp::CharacterVector>(exogen["variables"]).size();
    H.n_b = H.k * m * (Rcpp::as<int>(exogen["lags"]) + 1);
  }
  if (has_name(model, "deterministic")) H.n_c = H.k * Rcpp::as<Rcpp::CharacterVector>(model["deterministic"]).size();
  H.n_a0 = H.structural ? H.k * (H.k - 1) / 2 : 0;
  H.n_tot = H.n_a + H.n_b + H.n_c + H.n_a0;
  if (H.n_a + H.n_b + H.n_c != H.k * H.M) Rcpp::stop("data$Z does not match the model description.");

  // priors$coefficients (+ ssvs / bvs), initial$coefficients$draw
  c.varsel = BVG_VARSEL_NONE;
  if (H.n_tot > 0) {
    Rcpp::List pc = priors["coefficients"];
    c.mu = H.keep(Rcpp::as<Rcpp::NumericVector>(pc["mu"]));
    Rcpp::NumericMatrix vi = Rcpp::as<Rcpp::NumericMatrix>(pc["v_i"]);          // n_tot x n_tot
    bool diagonal = true;
    for (int r = 0; r < vi.nrow() && diagonal; ++r)
      for (int cc = 0; cc < vi.ncol(); ++cc)
        if (r != cc && vi(r, cc) != 0.0) { diagonal = false; break; }
    if (diagonal) {
      Rcpp::NumericVector d(vi.nrow());
      for (int r = 0; r < vi.nrow(); ++r) d[r] = vi(r, r);
      c.Vi = H.keep(d);
      c.vi_dense = 0;
    } else {
      c.Vi = H.keep(vi);
      c.vi_dense = 1;
    }
    Rcpp::List sel;
    if (has_name(pc, "ssvs")) {
      c.varsel = BVG_VARSEL_SSVS;
      sel = pc["ssvs"];
      c.tau0 = H.keep(Rcpp::as<Rcpp::NumericVector>(sel["tau0"]));
      c.tau1 = H.keep(Rcpp::as<Rcpp::NumericVector>(sel["tau1"]));
    }
    if (has_name(pc, "bvs")) {
      c.varsel = BVG_VARSEL_BVS;
      sel = pc["bvs"];
    }
    if (c.varsel != BVG_VARSEL_NONE) {
      H.varsel = true;
      c.inprior = H.keep(Rcpp::as<Rcpp::NumericVector>(sel["inprior"]));
      c.include = H.keep_index(Rcpp::as<Rcpp::NumericVector>(sel["include"]), &c.n_include);
    }
    if (H.tvp) {                                                                // bvartvpalg.cpp:96-112
      Rcpp::List ic = initial["coefficients"];
      c.tvp_shape = H.keep(Rcpp::as<Rcpp::NumericVector>(pc["shape"]));
      c.tvp_rate = H.keep(Rcpp::as<Rcpp::NumericVector>(pc["rate"]));
      c.a_init = H.keep(Rcpp::as<Rcpp::NumericVector>(ic["draw"]));
    }
  }

  // priors$sigma, initial$sigma (bvaralg.cpp:205-249)
  Rcpp::List ps = priors["sigma"], is = initial["sigma"];
  if (H.sv) {
    c.sigma_type = BVG_SIGMA_SV;
    c.sv_mu = H.keep(Rcpp::as<Rcpp::NumericVector>(ps["mu"]));
    c.sv_vi = H.keep(Rcpp::as<Rcpp::NumericVector>(ps["v_i"]));
    c.sigma_shape = H.keep(Rcpp::as<Rcpp::NumericVector>(ps["shape"]));
    c.sigma_rate = H.keep(Rcpp::as<Rcpp::NumericVector>(ps["rate"]));
    c.sv_h = H.keep(Rcpp::as<Rcpp::NumericVector>(is["h"]));                    // T x k
    c.sv_sigma_h = H.keep(Rcpp::as<Rcpp::NumericVector>(is["sigma_h"]));
    c.sv_constant = H.keep(Rcpp::as<Rcpp::NumericVector>(is["constant"]));
    c.sv_mixture = BVG_MIX_KSC1998;
  } else {
    c.sigma_i = H.keep(Rcpp::as<Rcpp::NumericVector>(is["sigma_i"]));
    if (has_name(ps, "df")) {
      c.sigma_type = BVG_SIGMA_WISHART;
      c.wishart_df = Rcpp::as<double>(ps["df"]);
      c.wishart_scale = H.keep(Rcpp::as<Rcpp::NumericVector>(ps["scale"]));
    }
    if (has_name(ps, "shape")) {
      c.sigma_type = BVG_SIGMA_GAMMA;
      c.sigma_shape = H.keep(Rcpp::as<Rcpp::NumericVector>(ps["shape"]));
      c.sigma_rate = H.keep(Rcpp::as<Rcpp::NumericVector>(ps["rate"]));
    }
  }

  // priors$psi: covariance block (bvaralg.cpp:157-204, bvartvpalg.cpp:143-176)
  c.psi_varsel = BVG_VARSEL_NONE;
  if (has_name(priors, "psi")) {
    H.covar = true;
    Rcpp::List pp = priors["psi"];
    c.psi_mu = H.keep(Rcpp::as<Rcpp::NumericVector>(pp["mu"]));
    c.psi_vi = H.keep(Rcpp::as<Rcpp::NumericVector>(pp["v_i"]));
    if (H.tvp) {
      Rcpp::List ip = initial["psi"];
      c.psi_shape = H.keep(Rcpp::as<Rcpp::NumericVector>(pp["shape"]));
      c.psi_rate = H.keep(Rcpp::as<Rcpp::NumericVector>(pp["rate"]));
      c.psi_init = H.keep(Rcpp::as<Rcpp::NumericVector>(ip["draw"]));
    }
    Rcpp::List psel;
    if (has_name(pp, "ssvs")) {
      c.psi_varsel = BVG_VARSEL_SSVS;
      psel = pp["ssvs"];
      c.psi_tau0 = H.keep(Rcpp::as<Rcpp::NumericVector>(psel["tau0"]));
      c.psi_tau1 = H.keep(Rcpp::as<Rcpp::NumericVector>(psel["tau1"]));
    }
    if (has_name(pp, "bvs")) {
      c.psi_varsel = BVG_VARSEL_BVS;
      psel = pp["bvs"];
    }
    if (c.psi_varsel != BVG_VARSEL_NONE) {
      H.psi_varsel = true;
      c.psi_inprior = H.keep(Rcpp::as<Rcpp::NumericVector>(psel["inprior"]));
      c.psi_include = H.keep_index(Rcpp::as<Rcpp::NumericVector>(psel["include"]), &c.psi_n_include);
    }
  }
}

// rows [r0, r0 + nr) of every retained draw of chain `chain` -> nr x iterations matrix (constant-parameter models) or
// (nr T) x iterations (TVP: vectorise(a_mat.rows(..)), time-major, bvartvpalg.cpp:521-524)
inline Rcpp::NumericMatrix slice_rows(const std::vector<double>& buf, long long chain, int iters, long long rows_per_draw,
                                      int r0, int nr, int T_blocks, int rows_per_block) {
  Rcpp::NumericMatrix out(nr * T_blocks, iters);
  const double* base = buf.data() + (size_t)chain * iters * rows_per_draw;
  for (int it = 0; it < iters; ++it)
    for (int t = 0; t < T_blocks; ++t)
      for (int r = 0; r < nr; ++r)
        out(t * nr + r, it) = base[(size_t)it * rows_per_draw + (size_t)t * rows_per_block + r0 + r];
  return out;
}

// Result of ONE chain in the reference's shape: list(data, model, priors, posteriors = list(a0, a, b, c, sigma)).
struct Buffers {
  std::vector<double> coef, sigma, lambda, sigma_h, sigma_v;
  std::vector<int32_t> status;
  int64_t rows[5];
};

inline Rcpp::List pack_chain(const Rcpp::List& object, const SpecHolder& H, const Buffers& B, long long chain) {
  const int iters = H.c.iterations, k = H.k, T = H.T;
  const int nrow = H.n_tot;                                      // coefficient rows per draw (per period for TVP)
  const int Tb = H.tvp ? T : 1;
  const bool lam_coef = H.varsel, tv_sigma = H.sv || (H.tvp && H.covar);
  Rcpp::List posteriors = Rcpp::List::create(Rcpp::Named("a0") = R_NilValue, Rcpp::Named("a") = R_NilValue,
                                             Rcpp::Named("b") = R_NilValue, Rcpp::Named("c") = R_NilValue,
                                             Rcpp::Named("sigma") = R_NilValue);
  const char* names[4] = {"a", "b", "c", "a0"};
  const int cnt[4] = {H.n_a, H.n_b, H.n_c, H.n_a0};
  int r0 = 0;
  for (int b = 0; b < 4; ++b) {
    if (cnt[b] == 0) continue;
    Rcpp::List part = Rcpp::List::create(
        Rcpp::Named("coeffs") = slice_rows(B.coef, chain, iters, B.rows[0], r0, cnt[b], Tb, nrow));
    if (H.tvp)                                                   // state variances, bvartvpalg.cpp:525,562-621
      part["sigma"] = slice_rows(B.sigma_v, chain, iters, B.rows[4], r0, cnt[b], 1, nrow);
    if (lam_coef) part["lambda"] = slice_rows(B.lambda, chain, iters, B.rows[2], r0, cnt[b], 1, nrow);
    posteriors[names[b]] = part;
    r0 += cnt[b];
  }
  Rcpp::List sg = Rcpp::List::create(
      Rcpp::Named("coeffs") = slice_rows(B.sigma, chain, iters, B.rows[1], 0, k * k * (tv_sigma ? T : 1), 1, 0));
  if (H.sv) {                                                    // vec(diag(sigma_h)) (bvaralg.cpp:526)
    Rcpp::NumericMatrix sh(k * k, iters);
    const double* base = B.sigma_h.data() + (size_t)chain * iters * k;
    for (int it = 0; it < iters; ++it)
      for (int i = 0; i < k; ++i) sh(i + k * i, it) = base[(size_t)it * k + i];
    sg["sigma"] = sh;
  }
  if (H.psi_varsel)                                              // psi inclusion draws follow the coefficient ones
    sg["lambda"] = slice_rows(B.lambda, chain, iters, B.rows[2], lam_coef ? nrow : 0, k * (k - 1) / 2, 1, 0);
  posteriors["sigma"] = sg;
  return Rcpp::List::create(Rcpp::Named("data") = object["data"], Rcpp::Named("model") = object["model"],
                            Rcpp::Named("priors") = object["priors"], Rcpp::Named("posteriors") = posteriors);
}

// One call: all chains of one model.  Returns a list with one reference-shaped result per chain (n_chains = 1: that result).
inline Rcpp::List run(const Rcpp::List& object, bool want_tvp, int n_chains, int n_gpus, int device, double seed_in) {
  SpecHolder H;
  // seed: one 64-bit value from R's generator inside the RNGScope of the exported wrapper (set.seed reproducibility,
  // src/RcppExports.cpp:22), unless the caller passes one
  unsigned long long seed = 0;
  if (seed_in >= 0) seed = (unsigned long long)seed_in;
  else seed = ((unsigned long long)(unif_rand() * 4294967296.0) << 32) | (unsigned long long)(unif_rand() * 4294967296.0);
  fill_spec(object, H, n_chains, 0, seed, n_gpus, device);
  if (H.tvp != want_tvp) Rcpp::stop("model$tvp does not match the sampler.");
  Buffers B;
  if (bvg_out_rows(&H.c, &B.rows[0], &B.rows[1], &B.rows[2], &B.rows[3], &B.rows[4]) != 0) Rcpp::stop(bvg_last_error());
  const size_t per = (size_t)n_chains * H.c.iterations;
  B.coef.resize(per * B.rows[0]); B.sigma.resize(per * B.rows[1]); B.lambda.resize(per * B.rows[2]);
  B.sigma_h.resize(per * B.rows[3]); B.sigma_v.resize(per * B.rows[4]); B.status.assign((size_t)n_chains, 0);
  bvg_out out;
  out.coef = B.coef.empty() ? NULL : B.coef.data();
  out.sigma = B.sigma.data();
  out.lambda = B.lambda.empty() ? NULL : B.lambda.data();
  out.sigma_h = B.sigma_h.empty() ? NULL : B.sigma_h.data();
  out.sigma_v = B.sigma_v.empty() ? NULL : B.sigma_v.data();
  out.status = B.status.data();
  const int rc = want_tvp ? bvg_bvartvpalg(&H.c, &out) : bvg_bvaralg(&H.c, &out);
  if (rc != 0) Rcpp::stop(bvg_last_error());                     // -> R error, like Rcpp::stop in the reference
  for (int cidx = 0; cidx < n_chains; ++cidx)
    if (B.status[(size_t)cidx] != 0) Rcpp::warning("chain %d: a Cholesky pivot was not positive", cidx + 1);
  if (n_chains == 1) return pack_chain(object, H, B, 0);
  Rcpp::List all(n_chains);
  for (int cidx = 0; cidx < n_chains; ++cidx) all[cidx] = pack_chain(object, H, B, cidx);
  return all;
}

}  // namespace bvgglue
#endif
