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
                                 s->inject->a_init_normal};
  const int cnt[VB_COUNT] = {n, n, n, n, k * T, k * T, k, k, k, k, k * (k - 1) / 2, n * T, n, n};
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
  std::vector<double> sm(bvar_smem_len(P.k, P.T, P.M, P.n, P.sigma_type, P.resid_mode) + 16);
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
        if (need == 0) bvar_chain<0>(grp, m, cs, rng, st.data(), sm.data(), out, c, s0, s1, s->burnin);
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
  std::vector<double> sm(tvp_smem_len(P.k, P.T, P.n, P.sigma_type) + 16);
  std::vector<double> scratch((size_t)tvp_scratch_len(P.T, P.n));
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
