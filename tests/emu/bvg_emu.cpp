// bvg_emu.cpp -- TEST-ONLY host build of the kernel logic.
//
// Compiles the very same group-generic headers the CUDA kernels are built from (bvartools_b200/csrc/*.cuh) with
// g++ and a single-thread group, so that the arithmetic, indexing and variate addressing of the kernels can be
// checked against the oracle on machines without a GPU (`pytest -m "not gpu"`).  It is NOT part of the product:
// nothing in bvartools_b200/ loads it, it is not a fallback, and it does not exercise the parallel decomposition
// (barriers, races) -- that is what the `-m gpu` parity tests and compute-sanitizer runs are for.
#include <string.h>
#include <barrier>
#include <memory>
#include <string>
#include <thread>
#include <vector>
#include "../../bvartools_b200/csrc/bvg_prepare.hpp"
#include "../../bvartools_b200/csrc/bvg_blocks.cuh"

using namespace bvg;

// SPMD emulation of a G-lane tile with real threads: sync() is a barrier, bcast()/sum() go through a shared
// exchange buffer with the same butterfly order as the device shuffles.  Slow, test-only.
struct ThreadShared {
  std::barrier<> bar;
  std::vector<double> slot, slot2;
  explicit ThreadShared(int G) : bar(G), slot(G), slot2(G) {}
};
struct ThreadGroup {
  static constexpr bool kCheapBcast = true;
  static constexpr bool kIsBlock = false;
  static constexpr bool kWarp32 = false;
  int lane, G;
  ThreadShared* sh;
  int tid() const { return lane; }
  int size() const { return G; }
  void sync() const { sh->bar.arrive_and_wait(); }
  double bcast(double v, int src) const {
    sh->slot[lane] = v;
    sync();
    const double r = sh->slot[src];
    sync();
    return r;
  }
  double sum(double v) const {
    for (int o = G / 2; o > 0; o >>= 1) {
      sh->slot[lane] = v;
      sync();
      v += sh->slot[lane ^ o];
      sync();
    }
    return v;
  }
};

template <class F>
static void run_lanes(int G, F&& body) {
  ThreadShared sh(G);
  std::vector<std::thread> th;
  for (int l = 0; l < G; ++l) th.emplace_back([&, l] { ThreadGroup g{l, G, &sh}; body(g); });
  for (auto& t : th) t.join();
}

static std::string g_err;

static void fill_rng(const bvg_spec* s, const Prepared& P, RngParams* rp) {
  memset(rp, 0, sizeof(*rp));
  rp->seed = s->seed;
  rp->chain_offset = (uint64_t)s->chain_offset;
  if (s->inject == nullptr) return;
  const long long sweeps = (long long)s->iterations + s->burnin;
  const int k = P.k, T = P.T, n = P.n;
  const double* ptr[VB_COUNT] = {s->inject->coef_normal, s->inject->ssvs_u, s->inject->bvs_u1, s->inject->bvs_u2,
                                 s->inject->sv_u, s->inject->sv_normal, s->inject->sigma_h_gamma,
                                 s->inject->h_init_normal, s->inject->omega_gamma, s->inject->wishart_chi2,
                                 s->inject->wishart_normal, s->inject->state_normal, s->inject->sigma_v_gamma,
                                 s->inject->a_init_normal, s->inject->psi_normal, s->inject->psi_u1, s->inject->psi_u2,
                                 s->inject->psi_sigma_v_gamma, s->inject->psi_init_normal};
  const int np = k * (k - 1) / 2;
  const int cnt[VB_COUNT] = {n, n, n, n, k * T, k * T, k, k, k, k, np, n * T, n, n, P.tvp ? np * T : np, np, np, np, np};
  for (int b = 0; b < VB_COUNT; ++b) {
    rp->inj[b] = ptr[b];
    rp->inj_stride[b] = cnt[b];
    rp->inj_chain_stride[b] = sweeps * cnt[b];
  }
}

extern "C" const char* emu_last_error(void) { return g_err.c_str(); }

extern "C" int emu_bvaralg(const bvg_spec* s, const bvg_out* o) {
  Prepared P;
  int rc = prepare(s, P);
  if (rc != 0) { g_err = P.error; return rc; }
  if (P.tvp) { g_err = "tvp spec passed to emu_bvaralg"; return -20; }
  BvarModel m;
  bind_bvar(P, P.blob.data(), P.include.data(), &m);
  RngParams rp;
  fill_rng(s, P, &rp);
  BvarOut out;
  out.coef = o->coef; out.sigma = o->sigma; out.lambda = o->lambda; out.sigma_h = o->sigma_h;
  std::vector<int> status(s->n_chains, 0);
  out.status = status.data();
  out.iters = s->iterations;
  std::vector<double> sm(bvar_smem_len(P.k, P.T, P.M, P.n, P.sigma_type, P.resid_mode, false, P.n_psi) + kron_smem_len(P.k, P.M) + 16);
  SerialGroup g;
  const int sweeps = s->iterations + s->burnin;
  const int chunk = s->sweeps_per_launch > 0 ? s->sweeps_per_launch : sweeps;
  const BvarConsts cs = bvar_consts_of(m);
  const int need = bvar_features_needed(m.sigma_type, m.varsel, m.resid_mode, m.vi_off != nullptr);
  const int G = s->threads_per_chain;          // 0/1: single-thread group; 2..64: threaded SPMD emulation
  for (long long c = 0; c < s->n_chains; ++c) {
    std::vector<double> st(P.state0);
    ChainRng rng(&rp, c);
    for (int s0 = 0; s0 < sweeps; s0 += chunk) {
      const int s1 = s0 + chunk < sweeps ? s0 + chunk : sweeps;
      auto go = [&](const auto& grp) {           // same variant choice as launch_bvar()
        if (P.kron) {
          KronConsts kc;
          kc.Linv = m.kron_Linv; kc.Aols = m.Aols; kc.SS = m.kron_SS;
          switch (P.k) {
            case 1: bvar_kron_chain<1>(grp, m, kc, rng, st.data(), sm.data(), out, c, s0, s1, s->burnin); break;
            case 2: bvar_kron_chain<2>(grp, m, kc, rng, st.data(), sm.data(), out, c, s0, s1, s->burnin); break;
            case 3: bvar_kron_chain<3>(grp, m, kc, rng, st.data(), sm.data(), out, c, s0, s1, s->burnin); break;
            default: bvar_kron_chain<4>(grp, m, kc, rng, st.data(), sm.data(), out, c, s0, s1, s->burnin); break;
          }
        }
        else if (need == 0) bvar_chain<0>(grp, m, cs, rng, st.data(), sm.data(), out, c, s0, s1, s->burnin);
        else if ((need & ~(int)F_VARSEL) == 0) bvar_chain<F_VARSEL>(grp, m, cs, rng, st.data(), sm.data(), out, c, s0, s1, s->burnin);
        else bvar_chain<F_ALL>(grp, m, cs, rng, st.data(), sm.data(), out, c, s0, s1, s->burnin);
      };
      if (G > 1) run_lanes(G, go); else go(g);
    }
    if (o->status) o->status[c] = status[c];
  }
  return 0;
}

extern "C" int emu_bvartvpalg(const bvg_spec* s, const bvg_out* o) {
  Prepared P;
  int rc = prepare(s, P);
  if (rc != 0) { g_err = P.error; return rc; }
  if (!P.tvp) { g_err = "non-tvp spec passed to emu_bvartvpalg"; return -20; }
  TvpModel m;
  bind_tvp(P, P.blob.data(), P.include.data(), &m);
  RngParams rp;
  fill_rng(s, P, &rp);
  TvpOut out;
  out.coef = o->coef; out.sigma = o->sigma; out.lambda = o->lambda; out.sigma_h = o->sigma_h; out.sigma_v = o->sigma_v;
  std::vector<int> status(s->n_chains, 0);
  out.status = status.data();
  out.iters = s->iterations;
  std::vector<double> sm(tvp_smem_len(P.k, P.T, P.n, P.sigma_type, TVP_GENERIC, P.n_psi) + 16);
  std::vector<double> scratch((size_t)tvp_scratch_len(P.T, P.k, P.M, P.n, P.n_psi, m.b.diag_sigma));
  SerialGroup g;
  const int sweeps = s->iterations + s->burnin;
  const int chunk = s->sweeps_per_launch > 0 ? s->sweeps_per_launch : sweeps;
  for (long long c = 0; c < s->n_chains; ++c) {
    std::vector<double> st(P.state0);
    ChainRng rng(&rp, c);
    for (int s0 = 0; s0 < sweeps; s0 += chunk) {
      const int s1 = s0 + chunk < sweeps ? s0 + chunk : sweeps;
      auto go = [&](const auto& grp) { tvp_chain(grp, m, rng, st.data(), scratch.data(), sm.data(), out, c, s0, s1, s->burnin); };
      if (s->threads_per_chain > 1) run_lanes(s->threads_per_chain, go); else go(g);
    }
    if (o->status) o->status[c] = status[c];
  }
  return 0;
}

// raw variate access, to pin the Philox stream / transforms from Python
extern "C" void emu_variates(uint64_t seed, uint64_t chain, int block, int sweep, int count, int kind, double shape,
                             double* out) {
  RngParams rp;
  memset(&rp, 0, sizeof(rp));
  rp.seed = seed;
  rp.chain_offset = chain;
  ChainRng rng(&rp, 0);
  for (int i = 0; i < count; ++i) {
    if (kind == 0) out[i] = rng.uniform(block, sweep, i);
    else if (kind == 1) out[i] = rng.normal(block, sweep, i);
    else if (kind == 2) out[i] = rng.gamma(block, sweep, i, shape);
    else out[i] = rng.chi2(block, sweep, i, shape);
  }
}

extern "C" void emu_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  const Philox4 p = philox4x32_10(c0, c1, c2, c3, k0, k1);
  out[0] = p.x; out[1] = p.y; out[2] = p.z; out[3] = p.w;
}

// FP64 transforms of the variate generators (host build of the same source): log, sqrt, sin / cos of 2 pi u, and the
// uniform built from two Philox words
extern "C" void emu_transforms(int n, const double* x, double* logx, double* sqrtx, double* s2pi, double* c2pi) {
  for (int i = 0; i < n; ++i) {
    logx[i] = bvg_logpos(x[i]);
    sqrtx[i] = bvg_sqrtpos(x[i]);
    bvg_sincos2pi(x[i], &s2pi[i], &c2pi[i]);
  }
}
extern "C" double emu_u53(uint32_t hi, uint32_t lo) { return u53(hi, lo); }

// ---- stand-alone building blocks: the same *_problem templates the CUDA kernels of bvg_blocks.cu instantiate,
//      one problem at a time on the host (single-thread group), same argument lists as the bvg_* entry points
static RngParams block_rp(uint64_t seed) {
  RngParams rp;
  memset(&rp, 0, sizeof(rp));
  rp.seed = seed;
  return rp;
}
static void block_inject(RngParams& rp, int block, const double* p, int count) {
  rp.inj[block] = p;
  rp.inj_stride[block] = count;
  rp.inj_chain_stride[block] = count;
}
static std::vector<unsigned char> block_mask(int n, const int32_t* include, int n_include) {
  std::vector<unsigned char> m(n, include == nullptr ? 1 : 0);
  if (include != nullptr) for (int i = 0; i < n_include; ++i) m[include[i]] = 1;
  return m;
}

extern "C" int emu_post_normal(int64_t B, int k, int T, int M, const double* y, const double* x, const double* sigma_i,
                               const double* a_prior, const double* v_i_prior, const double* eps, uint64_t seed,
                               int method, double* out) {
  const int n = k * M;
  std::vector<double> XX((size_t)M * M, 0.0), YX((size_t)k * M, 0.0), sm(post_normal_smem_len(n));
  for (int t = 0; t < T; ++t)
    for (int j = 0; j < M; ++j) {
      for (int j2 = 0; j2 < M; ++j2) XX[j + (size_t)M * j2] += x[j + (size_t)M * t] * x[j2 + (size_t)M * t];
      for (int i = 0; i < k; ++i) YX[i + (size_t)k * j] += y[i + (size_t)k * t] * x[j + (size_t)M * t];
    }
  RngParams rp = block_rp(seed);
  if (eps) block_inject(rp, VB_COEF_N, eps, n);
  SerialGroup g;
  int fl = 0;
  for (int64_t b = 0; b < B; ++b) {
    ChainRng rng(&rp, b);
    post_normal_problem(g, rng, k, M, XX.data(), YX.data(), sigma_i + (size_t)b * k * k, a_prior, v_i_prior, sm.data(),
                        out + (size_t)b * n, method, fl);
  }
  return fl;
}

extern "C" int emu_post_normal_sur(int64_t B, int k, int T, int n, const double* y, const double* z,
                                   const double* sigma_i, int sigma_tv, const double* a_prior, const double* v_i_prior,
                                   const double* eps, uint64_t seed, int method, double* out) {
  std::vector<double> sm(post_normal_smem_len(n));
  RngParams rp = block_rp(seed);
  if (eps) block_inject(rp, VB_COEF_N, eps, n);
  SerialGroup g;
  int fl = 0;
  const size_t ss = sigma_tv ? (size_t)k * T * k : (size_t)k * k;
  for (int64_t b = 0; b < B; ++b) {
    ChainRng rng(&rp, b);
    post_normal_sur_problem(g, rng, k, T, n, y, z, sigma_i + (size_t)b * ss, sigma_tv, a_prior, v_i_prior, sm.data(),
                            out + (size_t)b * n, method, fl);
  }
  return fl;
}

extern "C" int emu_bvs(int64_t B, int k, int T, int n, const double* y, const double* z, const double* a, int a_tv,
                       const double* lambda, const double* sigma_i, int sigma_tv, const double* prob_prior,
                       const int32_t* include, int n_include, const double* u1, const double* u2, uint64_t seed,
                       double* lambda_out) {
  std::vector<unsigned char> mask = block_mask(n, include, n_include);
  std::vector<double> sm(2 * (size_t)k * T);
  RngParams rp = block_rp(seed);
  if (u1) block_inject(rp, VB_BVS_U1, u1, n);
  if (u2) block_inject(rp, VB_BVS_U2, u2, n);
  SerialGroup g;
  const size_t as = a_tv ? (size_t)n * T : (size_t)n, ss = sigma_tv ? (size_t)k * T * k : (size_t)k * k;
  for (int64_t b = 0; b < B; ++b) {
    ChainRng rng(&rp, b);
    bvs_problem(g, rng, k, T, n, y, z, a + (size_t)b * as, a_tv, lambda + (size_t)b * n, sigma_i + (size_t)b * ss,
                sigma_tv, prob_prior, mask.data(), sm.data(), lambda_out + (size_t)b * n);
  }
  return 0;
}

extern "C" int emu_stochvol(int64_t B, int T, int k, const double* y, const double* h, const double* sigma,
                            const double* h_init, const double* constant, int mixture, const double* u,
                            const double* eps, uint64_t seed, double* h_out, int32_t* s_out) {
  const int J = (mixture == BVG_MIX_OCSN2007) ? 10 : 7;
  std::vector<double> tab(30, 0.0);
  for (int j = 0; j < J; ++j) {
    tab[j] = (J == 7) ? KSC_P[j] : OCSN_P[j];
    tab[10 + j] = (J == 7) ? KSC_MU[j] - 1.2704 : OCSN_MU[j];
    tab[20 + j] = (J == 7) ? KSC_S2[j] : OCSN_S2[j];
  }
  const size_t kt = (size_t)k * T;
  RngParams rp = block_rp(seed);
  if (u) block_inject(rp, VB_SV_U, u, (int)kt);
  if (eps) block_inject(rp, VB_SV_N, eps, (int)kt);
  BvarModel m;
  memset(&m, 0, sizeof(m));
  m.k = k; m.T = T; m.n_mix = J; m.sigma_type = SIG_SV; m.a0_from = -1;
  m.mix_p = tab.data(); m.mix_mu = tab.data() + 10; m.mix_s2 = tab.data() + 20;
  std::vector<double> U(kt), hh(kt), pr(kt), rh(kt);
  std::vector<int> sidx(kt);
  SerialGroup g;
  for (int64_t b = 0; b < B; ++b) {
    m.sv_const = constant + (size_t)b * k;
    for (size_t e = 0; e < kt; ++e) {
      const int i = (int)(e % k), t = (int)(e / k);
      U[e] = y[(size_t)b * kt + t + (size_t)T * i];
      hh[t + (size_t)T * i] = h[(size_t)b * kt + t + (size_t)T * i];
    }
    ChainRng rng(&rp, b);
    sv_step(g, rng, 0, m, U.data(), hh.data(), sigma + (size_t)b * k, h_init + (size_t)b * k, pr.data(), rh.data(),
            sidx.data());
    for (size_t e = 0; e < kt; ++e) { h_out[(size_t)b * kt + e] = hh[e]; if (s_out) s_out[(size_t)b * kt + e] = sidx[e]; }
  }
  return 0;
}

extern "C" int emu_kalman_dk(int64_t B, int k, int T, int n, const double* y, const double* z, const double* sigma_u,
                             const double* sigma_v, const double* Bmat, int tv, const double* a_init,
                             const double* P_init, const double* eps, uint64_t seed, double* out) {
  const int neps = n + T * (k + n);
  std::vector<double> sm(kalman_smem_len(k, n)), scratch((size_t)kalman_scratch_len(k, T, n));
  RngParams rp = block_rp(seed);
  if (eps) block_inject(rp, VB_STATE_N, eps, neps);
  SerialGroup g;
  int fl = 0;
  const size_t us = (tv & 1) ? (size_t)k * T * k : (size_t)k * k;
  const size_t vs = (tv & 2) ? (size_t)n * T * n : (size_t)n * n;
  const size_t bs = (tv & 4) ? (size_t)n * T * n : (size_t)n * n;
  for (int64_t b = 0; b < B; ++b) {
    ChainRng rng(&rp, b);
    kalman_dk_problem(g, rng, k, T, n, y, z, sigma_u + (size_t)b * us, sigma_v + (size_t)b * vs, Bmat + (size_t)b * bs,
                      tv, a_init + (size_t)b * n, P_init + (size_t)b * n * n, sm.data(), scratch.data(),
                      out + (size_t)b * n * (T + 1), fl);
  }
  return fl;
}

extern "C" int emu_ssvs(int64_t B, int n, const double* a, const double* tau0, const double* tau1,
                        const double* prob_prior, const int32_t* include, int n_include, const double* u,
                        uint64_t seed, double* lambda_out, double* vdiag_out) {
  std::vector<unsigned char> mask = block_mask(n, include, n_include);
  RngParams rp = block_rp(seed);
  if (u) block_inject(rp, VB_SSVS_U, u, n);
  for (int64_t b = 0; b < B; ++b) {
    ChainRng rng(&rp, b);
    for (int r = 0; r < n; ++r)
      ssvs_element(rng, r, a[(size_t)b * n + r], tau0[r], tau1[r], prob_prior[r], mask[r] != 0,
                   lambda_out + (size_t)b * n + r, vdiag_out + (size_t)b * n + r);
  }
  return 0;
}
