// bvg_api.cu -- C ABI of libbvargpu.so (include/bvargpu.h): validation, device residency, launch sequencing.
//
// Host side of the drop-in boundary.  Replaces the body of the reference's exported samplers
// (_bvartools_bvaralg / _bvartools_bvartvpalg, src/RcppExports.cpp:19-38): inputs are uploaded once, all
// iterations + burnin sweeps run on the GPU without host round trips (chunked only so that device->host
// copies of retained draws overlap the next chunk and the interrupt callback can be polled, the analogue of
// Rcpp::checkUserInterrupt every 20 draws, bvaralg.cpp:294-296), and the draws_* matrices come back in the
// reference's layout.  There is no CPU fallback: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <cmath>
#include <future>
#include <string>
#include <thread>
#include <vector>
#include "bvg_kernels.cuh"
#include "bvg_prepare.hpp"

using namespace bvg;

static thread_local std::string g_last_error;
static int (*g_poll)(void) = nullptr;
// multi-GPU one-shot calls: the shards run on worker threads, which never call the interrupt callback themselves (it may
// touch the R API: calling thread only); the calling thread polls it and raises g_abort, which the workers check
// between kernel batches
static thread_local bool t_worker = false;
static std::atomic<int> g_abort{0};

static int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return (int)e > 0 ? (int)e : 999;
}
#define CK(call)                                              \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call);     \
  } while (0)

// device allocations of one entry point: released on every exit path
struct DevGuard {
  std::vector<void*> ptrs;
  ~DevGuard() { for (void* q : ptrs) cudaFree(q); }
  cudaError_t alloc(void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
};

struct bvg_sampler {
  Prepared P;
  int device = 0;
  long long n_chains = 0;
  long long chain_offset = 0;
  uint64_t seed = 0;
  int sweeps = 0;
  int chunk = 0;
  int requested_tpc = 0;
  // device memory
  double* d_blob = nullptr;
  unsigned char* d_include = nullptr;
  double* d_state = nullptr;
  double* d_state0 = nullptr;   // one chain's initial state
  double* d_scratch = nullptr;  // TVP
  double* d_work = nullptr;     // large-VAR path: per-chain matrices
  bool large = false;
  long long scratch_per_chain = 0;
  double* d_inj[VB_COUNT] = {nullptr};
  // outputs (device)
  double* d_coef = nullptr; double* d_sigma = nullptr; double* d_lambda = nullptr; double* d_sigma_h = nullptr;
  double* d_sigma_v = nullptr; int* d_status = nullptr;
  long long rows_coef = 0, rows_sigma = 0, rows_lambda = 0, rows_sigma_h = 0, rows_sigma_v = 0;
  // launch
  BvarModel bm; TvpModel tm; RngParams rp; LaunchPlan plan;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> chunk_ev;
  bool timed = false;
  long long launches = 0;
  char kernel_name[64] = {0};
};

// Sigma is stored per period with stochastic volatility and for TVP models with the covariance block (bvartvpalg.cpp:505)
static inline bool tv_sigma_of(const Prepared& P) { return P.sigma_type == BVG_SIGMA_SV || (P.tvp && P.n_psi > 0); }

static void rows_of(const Prepared& P, long long* coef, long long* sigma, long long* lambda, long long* sh,
                    long long* sv) {
  const long long k = P.k, T = P.T, n = P.n;
  const bool svm = P.sigma_type == BVG_SIGMA_SV;
  *coef = P.tvp ? n * T : n;
  *sigma = tv_sigma_of(P) ? k * k * T : k * k;
  *lambda = ((P.varsel != BVG_VARSEL_NONE) ? n : 0) + ((P.psi_varsel != BVG_VARSEL_NONE) ? k * (k - 1) / 2 : 0);
  *sh = svm ? k : 0;
  *sv = P.tvp ? n : 0;
  if (P.structural) {                       // stored rows: [a | b | c | a0]; the mask itself is not an output
    *coef = P.tvp ? (long long)P.n_user * T : P.n_user;
    *lambda = (P.varsel_user != BVG_VARSEL_NONE) ? P.n_user : 0;
    *sv = P.tvp ? P.n_user : 0;
  }
}

extern "C" {

const char* bvg_version(void) { return "bvargpu 0.2.0 (sm_100a, abi 2)"; }
const char* bvg_last_error(void) { return g_last_error.c_str(); }

int bvg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void* bvg_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void bvg_host_free(void* p) { if (p) cudaFreeHost(p); }

void bvg_set_interrupt_poll(int (*poll)(void)) { g_poll = poll; }

int bvg_out_rows(const bvg_spec* spec, int64_t* coef_rows, int64_t* sigma_rows, int64_t* lambda_rows,
                 int64_t* sigma_h_rows, int64_t* sigma_v_rows) {
  if (!spec) return fail(-1, "spec is NULL");
  Prepared P;                              // the fields rows_of() reads, derived as prepare() derives them
  P.k = spec->k; P.T = spec->T; P.M = spec->M; P.n = spec->k * spec->M; P.tvp = spec->tvp ? 1 : 0;
  P.sigma_type = spec->sigma_type; P.varsel = spec->varsel;
  P.n_psi = (spec->psi_mu != nullptr) ? spec->k * (spec->k - 1) / 2 : 0;
  P.psi_varsel = (spec->psi_mu != nullptr) ? spec->psi_varsel : BVG_VARSEL_NONE;
  if (spec->structural) {                  // masked BVS model on [x; -y]; the kept rows [a | b | c | a0] are stored
    P.structural = 1;
    P.n_user = spec->k * spec->M + spec->k * (spec->k - 1) / 2;
    P.varsel_user = spec->varsel;
    P.M = spec->M + spec->k; P.n = spec->k * P.M; P.varsel = BVG_VARSEL_BVS;
  }
  long long a, b, c, d, e;
  rows_of(P, &a, &b, &c, &d, &e);
  if (coef_rows) *coef_rows = a;
  if (sigma_rows) *sigma_rows = b;
  if (lambda_rows) *lambda_rows = c;
  if (sigma_h_rows) *sigma_h_rows = d;
  if (sigma_v_rows) *sigma_v_rows = e;
  return 0;
}

void bvg_sampler_destroy(bvg_sampler* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaFree(s->d_blob); cudaFree(s->d_include); cudaFree(s->d_state); cudaFree(s->d_state0); cudaFree(s->d_scratch); cudaFree(s->d_work);
  for (int b = 0; b < VB_COUNT; ++b) cudaFree(s->d_inj[b]);
  cudaFree(s->d_coef); cudaFree(s->d_sigma); cudaFree(s->d_lambda); cudaFree(s->d_sigma_h); cudaFree(s->d_sigma_v);
  cudaFree(s->d_status);
  if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  for (cudaEvent_t e : s->chunk_ev) cudaEventDestroy(e);
  delete s;
}

static int alloc_out(double** p, long long chains, long long iters, long long rows) {
  if (rows == 0) { *p = nullptr; return 0; }
  CK(cudaMalloc(p, (size_t)chains * (size_t)iters * (size_t)rows * sizeof(double)));
  return 0;
}

int bvg_sampler_reset(bvg_sampler* s) {
  if (!s) return fail(-1, "sampler is NULL");
  CK(cudaSetDevice(s->device));
  // replicate the single initial state to every chain (device-side strided copy via 2D memcpy with pitch 0 is not
  // allowed, so broadcast with a tiny kernel-free loop of async copies in doubling steps)
  const size_t len = (size_t)s->P.state_len * sizeof(double);
  CK(cudaMemcpy(s->d_state, s->d_state0, len, cudaMemcpyDeviceToDevice));
  long long have = 1;
  while (have < s->n_chains) {
    const long long cnt = (have < s->n_chains - have) ? have : s->n_chains - have;
    CK(cudaMemcpy((char*)s->d_state + (size_t)have * len, s->d_state, (size_t)cnt * len, cudaMemcpyDeviceToDevice));
    have += cnt;
  }
  CK(cudaMemset(s->d_status, 0, (size_t)s->n_chains * sizeof(int)));
  return 0;
}

int bvg_sampler_create(const bvg_spec* spec, bvg_sampler** out) {
  if (!out) return fail(-1, "sampler output pointer is NULL");
  *out = nullptr;
  if (spec && spec->n_gpus > 1)
    return fail(-54, "the handle API drives one device: n_gpus > 1 is served by bvg_bvaralg / bvg_bvartvpalg");
  bvg_sampler* s = new bvg_sampler();
  int rc = prepare(spec, s->P);
  if (rc != 0) { std::string m = s->P.error; delete s; return fail(rc, m); }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    delete s;
    return fail(-50, "no CUDA device available: libbvargpu has no CPU fallback");
  }
  if (spec->device >= 0) {
    if (spec->device >= ndev) { delete s; return fail(-51, "device ordinal out of range"); }
    s->device = spec->device;
  } else {
    cudaGetDevice(&s->device);
  }
#define CKS(call)                                                                     \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) { int r__ = cuda_fail(e__, #call); bvg_sampler_destroy(s); return r__; } \
  } while (0)
  CKS(cudaSetDevice(s->device));
  const Prepared& P = s->P;
  s->n_chains = spec->n_chains;
  s->chain_offset = spec->chain_offset;
  s->seed = spec->seed;
  s->sweeps = P.iterations + P.burnin;
  s->requested_tpc = spec->threads_per_chain;
  s->chunk = spec->sweeps_per_launch > 0 ? spec->sweeps_per_launch : 0;
  rows_of(P, &s->rows_coef, &s->rows_sigma, &s->rows_lambda, &s->rows_sigma_h, &s->rows_sigma_v);

  CKS(cudaMalloc(&s->d_blob, P.blob.size() * sizeof(double)));
  CKS(cudaMemcpy(s->d_blob, P.blob.data(), P.blob.size() * sizeof(double), cudaMemcpyHostToDevice));
  if (!P.include.empty()) {
    CKS(cudaMalloc(&s->d_include, P.include.size()));
    CKS(cudaMemcpy(s->d_include, P.include.data(), P.include.size(), cudaMemcpyHostToDevice));
  }
  CKS(cudaMalloc(&s->d_state0, (size_t)P.state_len * sizeof(double)));
  CKS(cudaMemcpy(s->d_state0, P.state0.data(), (size_t)P.state_len * sizeof(double), cudaMemcpyHostToDevice));
  CKS(cudaMalloc(&s->d_state, (size_t)s->n_chains * P.state_len * sizeof(double)));
  CKS(cudaMalloc(&s->d_status, (size_t)s->n_chains * sizeof(int)));
  if (P.tvp) {
    bind_tvp(P, s->d_blob, s->d_include, &s->tm);
    s->scratch_per_chain = tvp_scratch_len(P.T, P.k, P.M, P.n, P.n_psi, s->tm.b.diag_sigma);
    CKS(cudaMalloc(&s->d_scratch, (size_t)s->n_chains * (size_t)s->scratch_per_chain * sizeof(double)));
  } else {
    bind_bvar(P, s->d_blob, s->d_include, &s->bm);
  }
  // injected variates
  memset(&s->rp, 0, sizeof(s->rp));
  s->rp.seed = s->seed;
  s->rp.chain_offset = (uint64_t)s->chain_offset;
  if (spec->inject) {
    const int k = P.k, T = P.T, n = P.n;
    const double* ptr[VB_COUNT] = {spec->inject->coef_normal, spec->inject->ssvs_u, spec->inject->bvs_u1,
                                   spec->inject->bvs_u2, spec->inject->sv_u, spec->inject->sv_normal,
                                   spec->inject->sigma_h_gamma, spec->inject->h_init_normal,
                                   spec->inject->omega_gamma, spec->inject->wishart_chi2,
                                   spec->inject->wishart_normal, spec->inject->state_normal,
                                   spec->inject->sigma_v_gamma, spec->inject->a_init_normal,
                                   spec->inject->psi_normal, spec->inject->psi_u1, spec->inject->psi_u2,
                                   spec->inject->psi_sigma_v_gamma, spec->inject->psi_init_normal};
    const int np = k * (k - 1) / 2;
    const int cnt[VB_COUNT] = {n, n, n, n, k * T, k * T, k, k, k, k, np, n * T, n, n, P.tvp ? np * T : np, np, np, np, np};
    for (int b = 0; b < VB_COUNT; ++b) {
      if (!ptr[b] || cnt[b] == 0) continue;
      const size_t bytes = (size_t)s->n_chains * s->sweeps * cnt[b] * sizeof(double);
      CKS(cudaMalloc(&s->d_inj[b], bytes));
      CKS(cudaMemcpy(s->d_inj[b], ptr[b], bytes, cudaMemcpyHostToDevice));
      s->rp.inj[b] = s->d_inj[b];
      s->rp.inj_stride[b] = cnt[b];
      s->rp.inj_chain_stride[b] = (long long)s->sweeps * cnt[b];
    }
  }
  // outputs
  if (alloc_out(&s->d_coef, s->n_chains, P.iterations, s->rows_coef) ||
      alloc_out(&s->d_sigma, s->n_chains, P.iterations, s->rows_sigma) ||
      alloc_out(&s->d_lambda, s->n_chains, P.iterations, s->rows_lambda) ||
      alloc_out(&s->d_sigma_h, s->n_chains, P.iterations, s->rows_sigma_h) ||
      alloc_out(&s->d_sigma_v, s->n_chains, P.iterations, s->rows_sigma_v)) {
    std::string m = g_last_error;
    bvg_sampler_destroy(s);
    return fail(2, m);
  }
  cudaError_t pe = P.tvp ? plan_tvp(s->tm, s->n_chains, s->requested_tpc, &s->plan)
                         : plan_bvar(s->bm, s->n_chains, s->requested_tpc == -1 ? 0 : s->requested_tpc, &s->plan);
  // large-VAR path: the precision matrix does not fit on chip (or the caller forces it with threads_per_chain = -1)
  if (!P.tvp && (pe == cudaErrorInvalidConfiguration || s->requested_tpc == -1)) {
    const bool ok = P.n > 0 && P.sigma_type == BVG_SIGMA_WISHART && P.varsel == BVG_VARSEL_NONE && P.vi_off < 0 && P.k <= 32;
    if (!ok) {
      bvg_sampler_destroy(s);
      return fail(-53, "large models (n beyond the shared-memory kernels) are supported for a diagonal prior precision, "
                       "a Wishart error-variance prior and no variable selection");
    }
    if (s->plan.kron) {   // the state holds W = chol(Sigma^-1)^-1 for the Kronecker kernel: store Sigma^-1 itself
      for (int i = 0; i < P.k; ++i) s->P.state0[i + (size_t)P.k * i] = 1.0 / (s->P.state0[i + (size_t)P.k * i] * s->P.state0[i + (size_t)P.k * i]);
      CKS(cudaMemcpy(s->d_state0, s->P.state0.data(), (size_t)P.state_len * sizeof(double), cudaMemcpyHostToDevice));
    }
    s->large = true;
    s->plan.kron = 0;
    s->plan.threads_per_chain = 0;
    CKS(cudaMalloc(&s->d_work, large_workspace_doubles(s->bm, s->n_chains) * sizeof(double)));
    pe = cudaSuccess;
  }
  if (pe != cudaSuccess) {
    bvg_sampler_destroy(s);
    return fail(-52, "model does not fit the shared-memory kernels (n too large for this path) or invalid threads_per_chain");
  }
  {
    const int need = P.tvp ? 0 : bvar_features_needed(s->bm.sigma_type, s->bm.varsel, s->bm.resid_mode, s->bm.vi_off != nullptr);
    const int tpc = s->plan.threads_per_chain;
    if (s->large) snprintf(s->kernel_name, sizeof(s->kernel_name), "large_dmma_n%d", P.n);
    else if (P.tvp) snprintf(s->kernel_name, sizeof(s->kernel_name), "tvp_g%d", tpc <= 32 ? tpc : 0);
    else if (s->plan.kron) snprintf(s->kernel_name, sizeof(s->kernel_name), "kron_k%d_g%d", P.k, tpc);
    else snprintf(s->kernel_name, sizeof(s->kernel_name), "bvar_g%d_f%d", tpc <= 32 ? tpc : 0,
                  need == 0 ? 0 : ((need & ~(int)F_VARSEL) == 0 ? 2 : 31));
  }
  CKS(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
  CKS(cudaEventCreate(&s->ev0));
  CKS(cudaEventCreate(&s->ev1));
  rc = bvg_sampler_reset(s);
  if (rc != 0) { bvg_sampler_destroy(s); return rc; }
#undef CKS
  *out = s;
  return 0;
}

static int launch_chunk(bvg_sampler* s, int s0, int s1, cudaStream_t stream) {
  const Prepared& P = s->P;
  cudaError_t e;
  if (P.tvp) {
    TvpOut o;
    o.coef = s->d_coef; o.sigma = s->d_sigma; o.lambda = s->d_lambda; o.sigma_h = s->d_sigma_h;
    o.sigma_v = s->d_sigma_v; o.status = s->d_status; o.iters = P.iterations;
    e = launch_tvp(s->plan, s->tm, s->rp, s->d_state, P.state_len, s->d_scratch, s->scratch_per_chain, o,
                   s->n_chains, s0, s1, P.burnin, stream);
  } else {
    BvarOut o;
    o.coef = s->d_coef; o.sigma = s->d_sigma; o.lambda = s->d_lambda; o.sigma_h = s->d_sigma_h;
    o.status = s->d_status; o.iters = P.iterations;
    if (s->large) {
      e = launch_bvar_large(s->bm, s->rp, s->d_state, P.state_len, s->d_work, o, s->n_chains, s0, s1, P.burnin, stream,
                            &s->launches);
      if (e != cudaSuccess) return cuda_fail(e, "large-VAR kernel launch");
      return 0;
    }
    e = launch_bvar(s->plan, s->bm, s->rp, s->d_state, P.state_len, o, s->n_chains, s0, s1, P.burnin, stream);
  }
  if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
  s->launches += 1;
  return 0;
}

static int default_chunk(const bvg_sampler* s) {
  if (s->chunk > 0) return s->chunk;
  int c = (s->sweeps + 11) / 12;
  if (c < 20) c = 20;
  return c;
}

int bvg_sampler_run(bvg_sampler* s, void* cuda_stream) {
  if (!s) return fail(-1, "sampler is NULL");
  CK(cudaSetDevice(s->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int chunk = default_chunk(s);
  s->launches = 0;
  CK(cudaEventRecord(s->ev0, st));
  for (int s0 = 0; s0 < s->sweeps; s0 += chunk) {
    const int s1 = (s0 + chunk < s->sweeps) ? s0 + chunk : s->sweeps;
    int rc = launch_chunk(s, s0, s1, st);
    if (rc) return rc;
    if (t_worker) { if (g_abort.load()) return fail(-100, "interrupted"); }
    else if (g_poll) {   // the callback is only meaningful between finished batches
      CK(cudaStreamSynchronize(st));
      if (g_poll()) return fail(-100, "interrupted");
    }
  }
  CK(cudaEventRecord(s->ev1, st));
  s->timed = true;
  return 0;
}

static int copy_rows(double* host, const double* dev, long long chains, long long iters, long long rows,
                     long long it0, long long it1, cudaStream_t st) {
  if (!host || !dev || rows == 0 || it1 <= it0) return 0;
  const size_t pitch = (size_t)iters * rows * sizeof(double);
  const size_t width = (size_t)(it1 - it0) * rows * sizeof(double);
  const size_t off = (size_t)it0 * rows;
  CK(cudaMemcpy2DAsync(host + off, pitch, dev + off, pitch, width, (size_t)chains, cudaMemcpyDeviceToHost, st));
  return 0;
}

int bvg_sampler_run_to_host(bvg_sampler* s, const bvg_out* h, void* cuda_stream) {
  if (!s || !h) return fail(-1, "sampler / output is NULL");
  CK(cudaSetDevice(s->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const Prepared& P = s->P;
  const int chunk = default_chunk(s);
  const int nchunks = (s->sweeps + chunk - 1) / chunk;
  while ((int)s->chunk_ev.size() < nchunks) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    s->chunk_ev.push_back(e);
  }
  s->launches = 0;
  CK(cudaEventRecord(s->ev0, st));
  int ci = 0;
  for (int s0 = 0; s0 < s->sweeps; s0 += chunk, ++ci) {
    const int s1 = (s0 + chunk < s->sweeps) ? s0 + chunk : s->sweeps;
    int rc = launch_chunk(s, s0, s1, st);
    if (rc) return rc;
    CK(cudaEventRecord(s->chunk_ev[ci], st));
    const long long it0 = (s0 > P.burnin ? s0 : P.burnin) - P.burnin;
    const long long it1 = (long long)s1 - P.burnin;
    if (it1 > it0) {
      CK(cudaStreamWaitEvent(s->copy_stream, s->chunk_ev[ci], 0));
      if ((rc = copy_rows(h->coef, s->d_coef, s->n_chains, P.iterations, s->rows_coef, it0, it1, s->copy_stream))) return rc;
      if ((rc = copy_rows(h->sigma, s->d_sigma, s->n_chains, P.iterations, s->rows_sigma, it0, it1, s->copy_stream))) return rc;
      if ((rc = copy_rows(h->lambda, s->d_lambda, s->n_chains, P.iterations, s->rows_lambda, it0, it1, s->copy_stream))) return rc;
      if ((rc = copy_rows(h->sigma_h, s->d_sigma_h, s->n_chains, P.iterations, s->rows_sigma_h, it0, it1, s->copy_stream))) return rc;
      if ((rc = copy_rows(h->sigma_v, s->d_sigma_v, s->n_chains, P.iterations, s->rows_sigma_v, it0, it1, s->copy_stream))) return rc;
    }
    if (t_worker) { if (g_abort.load()) { cudaDeviceSynchronize(); return fail(-100, "interrupted"); } }
    else if (g_poll) {
      CK(cudaStreamSynchronize(st));
      if (g_poll()) { cudaDeviceSynchronize(); return fail(-100, "interrupted"); }
    }
  }
  CK(cudaEventRecord(s->ev1, st));
  s->timed = true;
  CK(cudaStreamSynchronize(st));
  if (h->status) CK(cudaMemcpyAsync(h->status, s->d_status, (size_t)s->n_chains * sizeof(int), cudaMemcpyDeviceToHost, s->copy_stream));
  CK(cudaStreamSynchronize(s->copy_stream));
  return 0;
}

int bvg_sampler_sync(bvg_sampler* s) {
  if (!s) return fail(-1, "sampler is NULL");
  CK(cudaSetDevice(s->device));
  CK(cudaDeviceSynchronize());
  return 0;
}

int bvg_sampler_fetch(bvg_sampler* s, const bvg_out* h) {
  if (!s || !h) return fail(-1, "sampler / output is NULL");
  CK(cudaSetDevice(s->device));
  CK(cudaDeviceSynchronize());
  const size_t per = (size_t)s->n_chains * s->P.iterations * sizeof(double);
  if (h->coef && s->d_coef) CK(cudaMemcpy(h->coef, s->d_coef, per * s->rows_coef, cudaMemcpyDeviceToHost));
  if (h->sigma && s->d_sigma) CK(cudaMemcpy(h->sigma, s->d_sigma, per * s->rows_sigma, cudaMemcpyDeviceToHost));
  if (h->lambda && s->d_lambda) CK(cudaMemcpy(h->lambda, s->d_lambda, per * s->rows_lambda, cudaMemcpyDeviceToHost));
  if (h->sigma_h && s->d_sigma_h) CK(cudaMemcpy(h->sigma_h, s->d_sigma_h, per * s->rows_sigma_h, cudaMemcpyDeviceToHost));
  if (h->sigma_v && s->d_sigma_v) CK(cudaMemcpy(h->sigma_v, s->d_sigma_v, per * s->rows_sigma_v, cudaMemcpyDeviceToHost));
  if (h->status) CK(cudaMemcpy(h->status, s->d_status, (size_t)s->n_chains * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

int bvg_sampler_device_out(bvg_sampler* s, bvg_out* d) {
  if (!s || !d) return fail(-1, "sampler / output is NULL");
  d->coef = s->d_coef; d->sigma = s->d_sigma; d->lambda = s->d_lambda; d->sigma_h = s->d_sigma_h;
  d->sigma_v = s->d_sigma_v; d->status = s->d_status;
  return 0;
}

double bvg_sampler_last_kernel_ms(bvg_sampler* s) {
  if (!s || !s->timed) return -1.0;
  cudaSetDevice(s->device);
  if (cudaEventSynchronize(s->ev1) != cudaSuccess) return -1.0;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, s->ev0, s->ev1) != cudaSuccess) return -1.0;
  return (double)ms;
}

int64_t bvg_sampler_launch_count(bvg_sampler* s) { return s ? s->launches : 0; }
const char* bvg_sampler_kernel_name(bvg_sampler* s) { return s ? s->kernel_name : ""; }

static int one_shot_single(const bvg_spec* spec, const bvg_out* out);

// n_gpus > 1: contiguous blocks of global chain ids, one host thread per shard, results straight into the caller's buffers
static int one_shot_sharded(const bvg_spec* spec, const bvg_out* out, int shards) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(-50, "no CUDA device available: libbvargpu has no CPU fallback");
  }
  int64_t rc_[5] = {0, 0, 0, 0, 0};
  int rc = bvg_out_rows(spec, &rc_[0], &rc_[1], &rc_[2], &rc_[3], &rc_[4]);
  if (rc) return rc;
  const int64_t n = spec->n_chains, iters = spec->iterations, sweeps = (int64_t)spec->iterations + spec->burnin;
  const int base = spec->device >= 0 ? spec->device : 0;
  const int k = spec->k, T = spec->T, nint = spec->structural ? k * (spec->M + k) : k * spec->M, np = k * (k - 1) / 2;
  // per-chain strides of the injected arrays ([chain][sweep][count], counts as in bvg_sampler_create)
  const int64_t icnt[19] = {nint, nint, nint, nint, (int64_t)k * T, (int64_t)k * T, k, k, k, k, np, (int64_t)nint * T, nint, nint,
                            spec->tvp ? (int64_t)np * T : np, np, np, np, np};
  struct Shard { bvg_spec sp; bvg_out o; bvg_inject inj; int rc; std::string msg; };
  std::vector<Shard> sh((size_t)shards);
  for (int g = 0; g < shards; ++g) {
    const int64_t c0 = n * g / shards, c1 = n * (g + 1) / shards;
    Shard& S = sh[(size_t)g];
    S.sp = *spec;
    S.sp.n_gpus = 1;
    S.sp.device = (base + g) % ndev;
    S.sp.n_chains = c1 - c0;
    S.sp.chain_offset = spec->chain_offset + c0;
    S.o = *out;
    const int64_t per = c0 * iters;
    if (out->coef) S.o.coef = out->coef + per * rc_[0];
    if (out->sigma) S.o.sigma = out->sigma + per * rc_[1];
    if (out->lambda) S.o.lambda = out->lambda + per * rc_[2];
    if (out->sigma_h) S.o.sigma_h = out->sigma_h + per * rc_[3];
    if (out->sigma_v) S.o.sigma_v = out->sigma_v + per * rc_[4];
    if (out->status) S.o.status = out->status + c0;
    if (spec->inject) {
      S.inj = *spec->inject;
      const double** f = reinterpret_cast<const double**>(&S.inj);
      for (int b = 0; b < 19; ++b)
        if (f[b]) f[b] += c0 * sweeps * icnt[b];
      S.sp.inject = &S.inj;
    }
    S.rc = 0;
  }
  g_abort.store(0);
  std::vector<std::future<void>> fut;
  for (int g = 0; g < shards; ++g)
    fut.push_back(std::async(std::launch::async, [&sh, g]() {
      t_worker = true;
      Shard& S = sh[(size_t)g];
      S.rc = (S.sp.n_chains > 0) ? one_shot_single(&S.sp, &S.o) : 0;
      if (S.rc) S.msg = g_last_error;
    }));
  bool interrupted = false;
  for (auto& f : fut) {
    while (f.wait_for(std::chrono::milliseconds(20)) != std::future_status::ready)
      if (g_poll && !interrupted && g_poll()) { interrupted = true; g_abort.store(1); }
  }
  g_abort.store(0);
  if (interrupted) return fail(-100, "interrupted");
  for (int g = 0; g < shards; ++g)
    if (sh[(size_t)g].rc) return fail(sh[(size_t)g].rc, "GPU shard " + std::to_string(g) + ": " + sh[(size_t)g].msg);
  return 0;
}

static int one_shot(const bvg_spec* spec, const bvg_out* out, int want_tvp) {
  if (!spec || !out) return fail(-1, "spec / out is NULL");
  if ((spec->tvp ? 1 : 0) != want_tvp)
    return fail(-10, want_tvp ? "bvg_bvartvpalg needs spec.tvp = 1" : "bvg_bvaralg needs spec.tvp = 0");
  if (spec->abi_version != BVG_ABI_VERSION) return fail(-2, "bvg_spec.abi_version does not match this library");
  int shards = spec->n_gpus;
  if (shards < 0) shards = bvg_device_count();
  if (shards > 1 && spec->n_chains > 1) {
    if ((int64_t)shards > spec->n_chains) shards = (int)spec->n_chains;
    return one_shot_sharded(spec, out, shards);
  }
  return one_shot_single(spec, out);
}

static int one_shot_single(const bvg_spec* spec, const bvg_out* out) {
  bvg_sampler* s = nullptr;
  int rc = bvg_sampler_create(spec, &s);
  if (rc) return rc;
  rc = bvg_sampler_run_to_host(s, out, nullptr);
  std::string keep = g_last_error;
  bvg_sampler_destroy(s);
  if (rc) g_last_error = keep;
  return rc;
}

int bvg_bvaralg(const bvg_spec* spec, const bvg_out* out) { return one_shot(spec, out, 0); }
int bvg_bvartvpalg(const bvg_spec* spec, const bvg_out* out) { return one_shot(spec, out, 1); }

int bvg_moments(const double* dev_draws, int64_t chains, int64_t iters, int64_t rows, double* dev_sums,
                double* dev_sumsq, void* cuda_stream) {
  if (!dev_draws || !dev_sums || !dev_sumsq || rows < 1) return fail(-1, "invalid arguments to bvg_moments");
  cudaError_t e = launch_moments(dev_draws, chains, iters, rows, dev_sums, dev_sumsq, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "moments kernel");
  return 0;
}

int bvg_loglik_draws(const double* dev_coef, const double* dev_sigma, int64_t n_draws, int32_t k, int32_t M, int32_t T,
                     int32_t tvp, int32_t tv_sigma, const double* dev_Y, const double* dev_X, double* dev_ll_t,
                     void* cuda_stream) {
  if (!dev_sigma || !dev_Y || !dev_ll_t || (M > 0 && (!dev_coef || !dev_X)))
    return fail(-1, "invalid arguments to bvg_loglik_draws");
  if (k < 1 || k > 32) return fail(-2, "bvg_loglik_draws supports 1 <= k <= 32");
  if (bvg_device_count() == 0) return fail(-50, "no CUDA device available");
  cudaError_t e = launch_loglik_draws(k, M, T, tvp, tv_sigma, n_draws, dev_Y, dev_X, dev_coef, dev_sigma, dev_ll_t,
                                      (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "log-likelihood kernel");
  return 0;
}

// Post-estimation calls may be given another stream than the run: order them after its last launch.
static inline cudaError_t after_run(bvg_sampler* s, void* cuda_stream) {
  if (s == nullptr || !s->timed) return cudaSuccess;
  return cudaStreamWaitEvent((cudaStream_t)cuda_stream, s->ev1, 0);
}

int bvg_sampler_loglik(bvg_sampler* s, double* host_ll_t, void* cuda_stream) {
  if (!s || !host_ll_t) return fail(-1, "sampler / output is NULL");
  if (s->P.structural) return fail(-6, "the log-likelihood kernels do not cover structural models");
  CK(cudaSetDevice(s->device));
  CK(after_run(s, cuda_stream));
  const int T = s->P.T;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  double* d_ll = nullptr;
  CK(cudaMalloc((void**)&d_ll, (size_t)T * sizeof(double)));
  const int tv_sigma = tv_sigma_of(s->P) ? 1 : 0;
  const BvarModel& bm = s->P.tvp ? s->tm.b : s->bm;          // the TVP sampler keeps its data in the TvpModel
  int rc = bvg_loglik_draws(s->d_coef, s->d_sigma, (int64_t)s->n_chains * s->P.iterations, s->P.k, s->P.M, T, s->P.tvp,
                            tv_sigma, bm.Y, bm.X, d_ll, cuda_stream);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(host_ll_t, d_ll, (size_t)T * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = cuda_fail(e, "log-likelihood copy");
  }
  cudaFree(d_ll);
  return rc;
}

int bvg_loglik_draws_host(const double* coef, const double* sigma, int64_t n_draws, int32_t k, int32_t M, int32_t T,
                          int32_t tvp, int32_t tv_sigma, const double* Y, const double* X, double* ll_t) {
  if (!sigma || !Y || !ll_t || (M > 0 && (!coef || !X)) || n_draws < 1)
    return fail(-1, "invalid arguments to bvg_loglik_draws_host");
  if (k < 1 || k > 32) return fail(-2, "bvg_loglik_draws_host supports 1 <= k <= 32");
  if (bvg_device_count() == 0) return fail(-50, "no CUDA device available");
  const size_t cs = (size_t)k * M * (tvp ? T : 1), ss = (size_t)k * k * (tv_sigma ? T : 1);
  // draws go through the device in chunks of <= 256 MB; the per-period sums of the chunks are added on the host
  int64_t chunk = (int64_t)((256u << 20) / ((cs + ss) * sizeof(double)));
  if (chunk < 1) chunk = 1;
  if (chunk > n_draws) chunk = n_draws;
  double *dY = nullptr, *dX = nullptr, *dc = nullptr, *ds = nullptr, *dl = nullptr;
  std::vector<double> part((size_t)T);
  int rc = 0;
  cudaError_t e = cudaSuccess;
  auto done = [&](int r) { cudaFree(dY); cudaFree(dX); cudaFree(dc); cudaFree(ds); cudaFree(dl); return r; };
  if ((e = cudaMalloc((void**)&dY, (size_t)k * T * sizeof(double))) != cudaSuccess) return done(cuda_fail(e, "cudaMalloc"));
  if (M > 0 && (e = cudaMalloc((void**)&dX, (size_t)M * T * sizeof(double))) != cudaSuccess) return done(cuda_fail(e, "cudaMalloc"));
  if (M > 0 && (e = cudaMalloc((void**)&dc, (size_t)chunk * cs * sizeof(double))) != cudaSuccess) return done(cuda_fail(e, "cudaMalloc"));
  if ((e = cudaMalloc((void**)&ds, (size_t)chunk * ss * sizeof(double))) != cudaSuccess) return done(cuda_fail(e, "cudaMalloc"));
  if ((e = cudaMalloc((void**)&dl, (size_t)T * sizeof(double))) != cudaSuccess) return done(cuda_fail(e, "cudaMalloc"));
  if ((e = cudaMemcpy(dY, Y, (size_t)k * T * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return done(cuda_fail(e, "cudaMemcpy"));
  if (M > 0 && (e = cudaMemcpy(dX, X, (size_t)M * T * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return done(cuda_fail(e, "cudaMemcpy"));
  for (int t = 0; t < T; ++t) ll_t[t] = 0.0;
  for (int64_t d0 = 0; d0 < n_draws && rc == 0; d0 += chunk) {
    const int64_t nd = (d0 + chunk <= n_draws) ? chunk : n_draws - d0;
    if (M > 0 && (e = cudaMemcpy(dc, coef + (size_t)d0 * cs, (size_t)nd * cs * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return done(cuda_fail(e, "cudaMemcpy"));
    if ((e = cudaMemcpy(ds, sigma + (size_t)d0 * ss, (size_t)nd * ss * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return done(cuda_fail(e, "cudaMemcpy"));
    rc = bvg_loglik_draws(dc, ds, nd, k, M, T, tvp, tv_sigma, dY, dX, dl, nullptr);
    if (rc) break;
    if ((e = cudaMemcpy(part.data(), dl, (size_t)T * sizeof(double), cudaMemcpyDeviceToHost)) != cudaSuccess) return done(cuda_fail(e, "cudaMemcpy"));
    for (int t = 0; t < T; ++t) ll_t[t] += part[t];
  }
  return done(rc);
}

// OLS moments of a VAR on the host: XX' (M x M), A_ols (k x M), SSE_ols (k x k); false when X'X is (numerically) singular
static bool host_ols_moments(const double* Y, const double* X, int k, int M, int T, std::vector<double>& XX,
                             std::vector<double>& A, std::vector<double>& SSE) {
  if (T <= M || M < 1) return false;
  XX.assign((size_t)M * M, 0.0); A.assign((size_t)k * M, 0.0); SSE.assign((size_t)k * k, 0.0);
  std::vector<double> YX((size_t)k * M, 0.0);
  for (int t = 0; t < T; ++t) {
    const double* xt = X + (size_t)M * t;
    const double* yt = Y + (size_t)k * t;
    for (int j = 0; j < M; ++j) {
      for (int j2 = 0; j2 < M; ++j2) XX[j + (size_t)M * j2] += xt[j] * xt[j2];
      for (int i = 0; i < k; ++i) YX[i + (size_t)k * j] += yt[i] * xt[j];
    }
  }
  std::vector<long double> Cm((size_t)M * M);
  for (size_t e = 0; e < Cm.size(); ++e) Cm[e] = XX[e];
  if (!chol_ld(Cm, M)) return false;
  long double dmin = Cm[0], dmax = Cm[0];
  for (int j = 0; j < M; ++j) { const long double dd = Cm[j + (size_t)M * j]; if (dd < dmin) dmin = dd; if (dd > dmax) dmax = dd; }
  if (dmin < 1e-7L * dmax) return false;
  std::vector<long double> b(M), u(k), sse((size_t)k * k, 0.0L);
  for (int i = 0; i < k; ++i) {
    for (int j = 0; j < M; ++j) b[j] = YX[i + (size_t)k * j];
    for (int j = 0; j < M; ++j) {
      long double v = b[j];
      for (int c = 0; c < j; ++c) v -= Cm[j + (size_t)M * c] * b[c];
      b[j] = v / Cm[j + (size_t)M * j];
    }
    for (int j = M - 1; j >= 0; --j) {
      long double v = b[j];
      for (int c = j + 1; c < M; ++c) v -= Cm[c + (size_t)M * j] * b[c];
      b[j] = v / Cm[j + (size_t)M * j];
    }
    for (int j = 0; j < M; ++j) A[i + (size_t)k * j] = (double)b[j];
  }
  for (int t = 0; t < T; ++t) {
    for (int i = 0; i < k; ++i) {
      long double v = Y[i + (size_t)k * t];
      for (int j = 0; j < M; ++j) v -= (long double)A[i + (size_t)k * j] * X[j + (size_t)M * t];
      u[i] = v;
    }
    for (int i = 0; i < k; ++i)
      for (int j = 0; j < k; ++j) sse[i + (size_t)k * j] += u[i] * u[j];
  }
  for (int e = 0; e < k * k; ++e) SSE[e] = (double)sse[e];
  return true;
}

int bvg_sampler_loglik_total(bvg_sampler* s, double* total, void* cuda_stream) {
  if (!s || !total) return fail(-1, "sampler / output is NULL");
  if (s->P.structural) return fail(-6, "the log-likelihood kernels do not cover structural models");
  CK(cudaSetDevice(s->device));
  CK(after_run(s, cuda_stream));
  const bool fast = !s->P.tvp && !tv_sigma_of(s->P) && s->P.M > 0 && s->bm.XX && s->bm.Aols && s->bm.SSE &&
                    s->P.k <= 32;
  if (!fast) {                           // time-varying coefficients / variances: per-period kernel, summed on the host
    std::vector<double> ll((size_t)s->P.T);
    int rc = bvg_sampler_loglik(s, ll.data(), cuda_stream);
    if (rc) return rc;
    double t = 0.0;
    for (double v : ll) t += v;
    *total = t;
    return 0;
  }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  double* d_tot = nullptr;
  CK(cudaMalloc((void**)&d_tot, sizeof(double)));
  cudaError_t e = launch_loglik_total(s->P.k, s->P.M, s->P.T, (long long)s->n_chains * s->P.iterations, s->bm.XX,
                                      s->bm.Aols, s->bm.SSE, s->d_coef, s->d_sigma, d_tot, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(total, d_tot, sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_tot);
  if (e != cudaSuccess) return cuda_fail(e, "total log-likelihood kernel");
  return 0;
}

int bvg_sampler_loglik_total_device(bvg_sampler* s, double* dev_total, void* cuda_stream) {
  if (!s || !dev_total) return fail(-1, "sampler / output is NULL");
  if (s->P.structural) return fail(-6, "the log-likelihood kernels do not cover structural models");
  CK(cudaSetDevice(s->device));
  CK(after_run(s, cuda_stream));
  const bool fast = !s->P.tvp && !tv_sigma_of(s->P) && s->P.M > 0 && s->bm.XX && s->bm.Aols && s->bm.SSE && s->P.k <= 32;
  if (!fast) return fail(-6, "bvg_sampler_loglik_total_device covers constant-parameter models with T > M (cross-moment kernel)");
  cudaError_t e = launch_loglik_total(s->P.k, s->P.M, s->P.T, (long long)s->n_chains * s->P.iterations, s->bm.XX,
                                      s->bm.Aols, s->bm.SSE, s->d_coef, s->d_sigma, dev_total, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "total log-likelihood kernel");
  return 0;
}

int bvg_loglik_total_host(const double* coef, const double* sigma, int64_t n_draws, int32_t k, int32_t M, int32_t T,
                          const double* Y, const double* X, double* total) {
  if (!coef || !sigma || !Y || !X || !total || n_draws < 1 || M < 1)
    return fail(-1, "invalid arguments to bvg_loglik_total_host");
  if (k < 1 || k > 32) return fail(-2, "bvg_loglik_total_host supports 1 <= k <= 32");
  if (bvg_device_count() == 0) return fail(-50, "no CUDA device available");
  std::vector<double> XX, A, SSE;
  if (!host_ols_moments(Y, X, k, M, T, XX, A, SSE)) {           // no OLS projection: per-period path
    std::vector<double> ll((size_t)T);
    int rc = bvg_loglik_draws_host(coef, sigma, n_draws, k, M, T, 0, 0, Y, X, ll.data());
    if (rc) return rc;
    double t = 0.0;
    for (double v : ll) t += v;
    *total = t;
    return 0;
  }
  const size_t cs = (size_t)k * M, ss = (size_t)k * k;
  int64_t chunk = (int64_t)((256u << 20) / ((cs + ss) * sizeof(double)));
  if (chunk < 1) chunk = 1;
  if (chunk > n_draws) chunk = n_draws;
  double *dXX = nullptr, *dA = nullptr, *dS = nullptr, *dc = nullptr, *ds = nullptr, *dt = nullptr;
  cudaError_t e = cudaSuccess;
  auto done = [&](int r) { cudaFree(dXX); cudaFree(dA); cudaFree(dS); cudaFree(dc); cudaFree(ds); cudaFree(dt); return r; };
#define BVG_TRY(call) do { if ((e = (call)) != cudaSuccess) return done(cuda_fail(e, #call)); } while (0)
  BVG_TRY(cudaMalloc((void**)&dXX, XX.size() * sizeof(double)));
  BVG_TRY(cudaMalloc((void**)&dA, A.size() * sizeof(double)));
  BVG_TRY(cudaMalloc((void**)&dS, SSE.size() * sizeof(double)));
  BVG_TRY(cudaMalloc((void**)&dc, (size_t)chunk * cs * sizeof(double)));
  BVG_TRY(cudaMalloc((void**)&ds, (size_t)chunk * ss * sizeof(double)));
  BVG_TRY(cudaMalloc((void**)&dt, sizeof(double)));
  BVG_TRY(cudaMemcpy(dXX, XX.data(), XX.size() * sizeof(double), cudaMemcpyHostToDevice));
  BVG_TRY(cudaMemcpy(dA, A.data(), A.size() * sizeof(double), cudaMemcpyHostToDevice));
  BVG_TRY(cudaMemcpy(dS, SSE.data(), SSE.size() * sizeof(double), cudaMemcpyHostToDevice));
  double sum = 0.0;
  for (int64_t d0 = 0; d0 < n_draws; d0 += chunk) {
    const int64_t nd = (d0 + chunk <= n_draws) ? chunk : n_draws - d0;
    BVG_TRY(cudaMemcpy(dc, coef + (size_t)d0 * cs, (size_t)nd * cs * sizeof(double), cudaMemcpyHostToDevice));
    BVG_TRY(cudaMemcpy(ds, sigma + (size_t)d0 * ss, (size_t)nd * ss * sizeof(double), cudaMemcpyHostToDevice));
    BVG_TRY(launch_loglik_total(k, M, T, nd, dXX, dA, dS, dc, ds, dt, nullptr));
    double part = 0.0;
    BVG_TRY(cudaMemcpy(&part, dt, sizeof(double), cudaMemcpyDeviceToHost));
    sum += part;
  }
#undef BVG_TRY
  *total = sum;
  return done(0);
}

int bvg_loglik_normal(const double* u, const double* sigma, int32_t k, int32_t T, int32_t sigma_rows, int64_t n_batch,
                      double* out) {
  if (!u || !sigma || !out || k < 1 || T < 1 || n_batch < 1) return fail(-1, "invalid arguments to bvg_loglik_normal");
  if (sigma_rows != k && sigma_rows != k * T) return fail(-3, "sigma must be k x k or (k T) x k");
  if (k > 32) return fail(-2, "bvg_loglik_normal supports k <= 32");
  if (bvg_device_count() == 0) return fail(-50, "no CUDA device available");
  for (int64_t e = 0; e < n_batch * (int64_t)k * T; ++e)
    if (u[e] != u[e]) return fail(-20, "Argument 'u' contains NAs.");
  const int tv = sigma_rows > k ? 1 : 0;
  const size_t sg_len = (size_t)k * k * (tv ? T : 1);
  std::vector<double> sg(sg_len * (size_t)n_batch);
  for (int64_t b = 0; b < n_batch; ++b) {
    const double* src = sigma + (size_t)b * sigma_rows * k;
    double* dst = sg.data() + (size_t)b * sg_len;
    if (!tv) { for (size_t e = 0; e < sg_len; ++e) dst[e] = src[e]; }
    else  // rows t k .. (t + 1) k - 1 of the (k T) x k column-major stack  (loglik_normal.cpp:47-48)
      for (int t = 0; t < T; ++t)
        for (int c = 0; c < k; ++c)
          for (int r = 0; r < k; ++r) dst[(size_t)t * k * k + r + (size_t)k * c] = src[(size_t)t * k + r + (size_t)sigma_rows * c];
  }
  double *d_u = nullptr, *d_sg = nullptr, *d_out = nullptr;
  DevGuard mem;
  CK(mem.alloc((void**)&d_u, (size_t)n_batch * k * T * sizeof(double)));
  CK(mem.alloc((void**)&d_sg, sg.size() * sizeof(double)));
  CK(mem.alloc((void**)&d_out, (size_t)n_batch * T * sizeof(double)));
  CK(cudaMemcpy(d_u, u, (size_t)n_batch * k * T * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_sg, sg.data(), sg.size() * sizeof(double), cudaMemcpyHostToDevice));
  int rc = 0;
  for (int64_t b = 0; b < n_batch && rc == 0; ++b)   // one "draw" whose residuals are the data: u = Y, no regressors
    rc = bvg_loglik_draws(nullptr, d_sg + (size_t)b * sg_len, 1, k, 0, T, 0, tv, d_u + (size_t)b * k * T, nullptr,
                          d_out + (size_t)b * T, nullptr);
  if (rc == 0) {
    cudaError_t e = cudaMemcpy(out, d_out, (size_t)n_batch * T * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "log-likelihood copy");
  }
  return rc;
}

int bvg_summary_draws(const double* dev_draws, int64_t n_draws, int64_t rows, const double* probs, int32_t n_probs,
                      double* mean, double* sd, double* quantiles, void* cuda_stream) {
  if (!dev_draws || n_draws < 1 || rows < 1 || n_probs < 0 || n_probs > 16 || (n_probs > 0 && (!probs || !quantiles)))
    return fail(-1, "invalid arguments to bvg_summary_draws");
  for (int q = 0; q < n_probs; ++q)
    if (!(probs[q] >= 0.0 && probs[q] <= 1.0)) return fail(-4, "probabilities must lie in [0, 1]");
  if (bvg_device_count() == 0) return fail(-50, "no CUDA device available");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  double *d_sums = nullptr, *d_sq = nullptr, *d_ss = nullptr, *d_p = nullptr, *d_q = nullptr;
  cudaError_t e = cudaSuccess;
  auto done = [&](int r) { cudaFree(d_sums); cudaFree(d_sq); cudaFree(d_ss); cudaFree(d_p); cudaFree(d_q); return r; };
#define BVG_TRY(call) do { if ((e = (call)) != cudaSuccess) return done(cuda_fail(e, #call)); } while (0)
  BVG_TRY(cudaMalloc((void**)&d_sums, (size_t)rows * 8));
  BVG_TRY(cudaMalloc((void**)&d_sq, (size_t)rows * 8));
  BVG_TRY(cudaMalloc((void**)&d_ss, (size_t)rows * 8));
  BVG_TRY(launch_moments(dev_draws, 1, n_draws, rows, d_sums, d_sq, st));
  BVG_TRY(launch_centered_ss(dev_draws, n_draws, rows, d_sums, d_ss, st));
  std::vector<double> hs((size_t)rows), hss((size_t)rows);
  BVG_TRY(cudaMemcpyAsync(hs.data(), d_sums, (size_t)rows * 8, cudaMemcpyDeviceToHost, st));
  BVG_TRY(cudaMemcpyAsync(hss.data(), d_ss, (size_t)rows * 8, cudaMemcpyDeviceToHost, st));
  if (n_probs > 0) {
    BVG_TRY(cudaMalloc((void**)&d_p, (size_t)n_probs * 8));
    BVG_TRY(cudaMalloc((void**)&d_q, (size_t)n_probs * rows * 8));
    BVG_TRY(cudaMemcpyAsync(d_p, probs, (size_t)n_probs * 8, cudaMemcpyHostToDevice, st));
    BVG_TRY(launch_quantiles(dev_draws, n_draws, rows, d_p, n_probs, d_q, st));
    BVG_TRY(cudaMemcpyAsync(quantiles, d_q, (size_t)n_probs * rows * 8, cudaMemcpyDeviceToHost, st));
  }
  BVG_TRY(cudaStreamSynchronize(st));
#undef BVG_TRY
  for (int64_t r = 0; r < rows; ++r) {
    if (mean) mean[r] = hs[r] / (double)n_draws;
    if (sd) sd[r] = (n_draws > 1) ? std::sqrt(hss[r] / (double)(n_draws - 1)) : 0.0;
  }
  return done(0);
}

int bvg_sampler_summary(bvg_sampler* s, int32_t which, const double* probs, int32_t n_probs, double* mean, double* sd,
                        double* quantiles, void* cuda_stream) {
  if (!s) return fail(-1, "sampler is NULL");
  CK(cudaSetDevice(s->device));
  CK(after_run(s, cuda_stream));
  const double* buf = nullptr;
  long long rows = 0;
  switch (which) {
    case 0: buf = s->d_coef; rows = s->rows_coef; break;
    case 1: buf = s->d_sigma; rows = s->rows_sigma; break;
    case 2: buf = s->d_lambda; rows = s->rows_lambda; break;
    case 3: buf = s->d_sigma_h; rows = s->rows_sigma_h; break;
    case 4: buf = s->d_sigma_v; rows = s->rows_sigma_v; break;
    default: return fail(-1, "which must be 0 (coef), 1 (sigma), 2 (lambda), 3 (sigma_h) or 4 (sigma_v)");
  }
  if (!buf || rows < 1) return fail(-6, "the sampler has no such output");
  return bvg_summary_draws(buf, (int64_t)s->n_chains * s->P.iterations, rows, probs, n_probs, mean, sd, quantiles,
                           cuda_stream);
}

int bvg_tsse_draws(const double* dev_draws, int64_t n_chains, int64_t iters, int64_t rows, double* tsse,
                   void* cuda_stream) {
  if (!dev_draws || !tsse || n_chains < 1 || iters < 1 || rows < 1) return fail(-1, "invalid arguments to bvg_tsse_draws");
  if (iters < 3 || iters >= 2500000) return fail(-2, "bvg_tsse_draws: need 3 <= iterations per chain < 2.5e6");
  if (bvg_device_count() == 0) return fail(-50, "no CUDA device available");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  double* d_out = nullptr;
  CK(cudaMalloc((void**)&d_out, (size_t)rows * 8));
  cudaError_t e = launch_tsse(dev_draws, n_chains, iters, rows, d_out, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(tsse, d_out, (size_t)rows * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_out);
  if (e != cudaSuccess) return cuda_fail(e, "time-series SE kernels");
  return 0;
}

int bvg_sampler_tsse(bvg_sampler* s, int32_t which, double* tsse, void* cuda_stream) {
  if (!s) return fail(-1, "sampler is NULL");
  CK(cudaSetDevice(s->device));
  CK(after_run(s, cuda_stream));
  const double* buf = nullptr;
  long long rows = 0;
  switch (which) {
    case 0: buf = s->d_coef; rows = s->rows_coef; break;
    case 1: buf = s->d_sigma; rows = s->rows_sigma; break;
    case 2: buf = s->d_lambda; rows = s->rows_lambda; break;
    case 3: buf = s->d_sigma_h; rows = s->rows_sigma_h; break;
    case 4: buf = s->d_sigma_v; rows = s->rows_sigma_v; break;
    default: return fail(-1, "which must be 0 (coef), 1 (sigma), 2 (lambda), 3 (sigma_h) or 4 (sigma_v)");
  }
  if (!buf || rows < 1) return fail(-6, "the sampler has no such output");
  return bvg_tsse_draws(buf, (int64_t)s->n_chains, (int64_t)s->P.iterations, rows, tsse, cuda_stream);
}

int bvg_irf_draws(const double* dev_coef, int64_t coef_stride, const double* dev_sigma, int64_t sigma_stride,
                  int64_t n_draws, int32_t k, int32_t p, int32_t n_ahead, int32_t type, int32_t impulse, int32_t response,
                  int32_t shock_kind, double shock, int32_t cumulative, double* dev_out, void* cuda_stream) {
  if (type == 3 || type == 4) return fail(-12, "Structural IR requires that draws of 'A0' are contained in the 'bvar' x.");
  return bvg_irf_draws_a0(dev_coef, coef_stride, dev_sigma, sigma_stride, nullptr, 0, 0, n_draws, k, p, n_ahead, type, impulse,
                          response, shock_kind, shock, cumulative, dev_out, cuda_stream);
}

int bvg_irf_draws_a0(const double* dev_coef, int64_t coef_stride, const double* dev_sigma, int64_t sigma_stride,
                     const double* dev_a0, int64_t a0_stride, int32_t a0_compact, int64_t n_draws, int32_t k, int32_t p,
                     int32_t n_ahead, int32_t type, int32_t impulse, int32_t response, int32_t shock_kind, double shock,
                     int32_t cumulative, double* dev_out, void* cuda_stream) {
  if (!dev_out || n_draws < 1 || (p > 0 && !dev_coef)) return fail(-1, "invalid arguments to bvg_irf_draws");
  if (type < 0 || type > 4) return fail(-7, "Argument 'type' not known.");      // feir / oir / gir / sir / sgir (R/irf.bvar.R:84)
  if (type >= 3 && !dev_a0)
    return fail(-12, "Structural IR requires that draws of 'A0' are contained in the 'bvar' x.");   // R/irf.bvar.R:101-104
  if (type >= 3 && !a0_compact && k > 16) return fail(-2, "structural impulse responses from full A0 matrices need k <= 16");
  if ((type == 1 || type == 2 || type == 4 || shock_kind != 0) && !dev_sigma)
    return fail(-8, "OIR, GIR, SGIR require that the 'bvar' x contains draws of 'Sigma'.");   // R/irf.bvar.R:112-115
  if (shock_kind < 0 || shock_kind > 2) return fail(-9, "Invalid specification of argument 'shock'.");
  if (k < 1 || k > 32 || impulse < 0 || impulse >= k || response < 0 || response >= k || n_ahead < 0 || p < 0)
    return fail(-2, "bvg_irf_draws: need 1 <= k <= 32, 0 <= impulse, response < k, n_ahead >= 0");
  if (bvg_device_count() == 0) return fail(-50, "no CUDA device available");
  cudaError_t e = launch_irf_draws(dev_coef, coef_stride, dev_sigma, sigma_stride, n_draws, k, p, n_ahead, type, impulse,
                                   response, shock_kind, shock, cumulative, dev_out, (cudaStream_t)cuda_stream, dev_a0,
                                   a0_stride, a0_compact ? 2 : 1);
  if (e != cudaSuccess) return cuda_fail(e, "impulse-response kernel");
  return 0;
}

int bvg_sampler_irf(bvg_sampler* s, int32_t p, int32_t n_ahead, int32_t type, int32_t impulse, int32_t response,
                    int32_t shock_kind, double shock, int32_t cumulative, int32_t period, const double* probs,
                    int32_t n_probs, double* quantiles, double* mean, double* host_draws, void* cuda_stream) {
  if (!s) return fail(-1, "sampler is NULL");
  CK(cudaSetDevice(s->device));
  CK(after_run(s, cuda_stream));
  const int k = s->P.k, T = s->P.T;
  if (p < 0 || (long long)k * k * p > (long long)s->P.n) return fail(-2, "p lags do not fit the coefficient vector");
  const bool tv_coef = s->P.tvp != 0, tv_sig = tv_sigma_of(s->P);
  if (tv_coef || tv_sig) {
    if (period < 0) period = T - 1;                                              // R/irf.bvar.R:127-133 (1-based there)
    if (period >= T) return fail(-11, "Implausible specification of argument 'period'.");
  }
  const long long nd = (long long)s->n_chains * s->P.iterations;
  const double* coef = s->d_coef + (tv_coef ? (size_t)period * (s->P.structural ? s->P.n_user : s->P.n) : 0);
  const double* sig = s->d_sigma + (tv_sig ? (size_t)period * k * k : 0);
  double* d_out = nullptr;
  CK(cudaMalloc((void**)&d_out, (size_t)nd * (n_ahead + 1) * sizeof(double)));
  const double* a0 = nullptr;                     // structural sampler: compact a0 rows at the end of every coefficient draw
  if (type >= 3) {
    if (!s->P.structural) { cudaFree(d_out); return fail(-12, "Structural IR requires that draws of 'A0' are contained in the 'bvar' x."); }
    a0 = coef + s->P.a0_from;                    // (period offset included for TVP draws)
  }
  int rc = bvg_irf_draws_a0(coef, s->rows_coef, sig, s->rows_sigma, a0, s->rows_coef, 1, nd, k, p, n_ahead, type, impulse,
                            response, shock_kind, shock, cumulative, d_out, cuda_stream);
  if (rc == 0 && (n_probs > 0 || mean))
    rc = bvg_summary_draws(d_out, nd, n_ahead + 1, probs, n_probs, mean, nullptr, quantiles, cuda_stream);
  if (rc == 0 && host_draws) {
    cudaError_t e = cudaMemcpyAsync(host_draws, d_out, (size_t)nd * (n_ahead + 1) * sizeof(double), cudaMemcpyDeviceToHost,
                                    (cudaStream_t)cuda_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)cuda_stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "impulse-response copy");
  }
  cudaFree(d_out);
  return rc;
}

int bvg_fevd_draws(const double* dev_coef, int64_t coef_stride, const double* dev_sigma, int64_t sigma_stride,
                   int64_t n_draws, int32_t k, int32_t p, int32_t n_ahead, int32_t type, int32_t response,
                   int32_t normalise_gir, double* host_out, void* cuda_stream) {
  if (type == 3 || type == 4)
    return fail(-12, "Structural FEVD requires that draws of 'A0' are contained in the 'bvar' object.");
  return bvg_fevd_draws_a0(dev_coef, coef_stride, dev_sigma, sigma_stride, nullptr, 0, 0, n_draws, k, p, n_ahead, type, response,
                           normalise_gir, host_out, cuda_stream);
}

int bvg_fevd_draws_a0(const double* dev_coef, int64_t coef_stride, const double* dev_sigma, int64_t sigma_stride,
                      const double* dev_a0, int64_t a0_stride, int32_t a0_compact, int64_t n_draws, int32_t k, int32_t p,
                      int32_t n_ahead, int32_t type, int32_t response, int32_t normalise_gir, double* host_out,
                      void* cuda_stream) {
  if (!host_out || n_draws < 1 || (p > 0 && !dev_coef)) return fail(-1, "invalid arguments to bvg_fevd_draws");
  if (!dev_sigma) return fail(-8, "Argument 'object' must include draws of the variance-covariance matrix Sigma.");
  if (type < 1 || type > 4) return fail(-7, "The specified type of the used impulse response is not known.");
  if (type >= 3 && !dev_a0)
    return fail(-12, "Structural FEVD requires that draws of 'A0' are contained in the 'bvar' object.");   // R/fevd.bvar.R:98-101
  if (type >= 3 && !a0_compact && k > 16) return fail(-2, "structural decompositions from full A0 matrices need k <= 16");
  if (k < 1 || k > 32 || response < 0 || response >= k || n_ahead < 0 || p < 0)
    return fail(-2, "bvg_fevd_draws: need 1 <= k <= 32, 0 <= response < k, n_ahead >= 0");
  if (bvg_device_count() == 0) return fail(-50, "no CUDA device available");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const size_t nout = (size_t)(n_ahead + 1) * k;
  double* d_out = nullptr;
  CK(cudaMalloc((void**)&d_out, nout * sizeof(double)));
  cudaError_t e = launch_fevd_draws(dev_coef, coef_stride, dev_sigma, sigma_stride, n_draws, k, p, n_ahead, type, response,
                                    d_out, st, dev_a0, a0_stride, a0_compact ? 2 : 1);
  if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, d_out, nout * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_out);
  if (e != cudaSuccess) return cuda_fail(e, "variance-decomposition kernel");
  if ((type == 2 || type == 4) && normalise_gir)                                // R/fevd.bvar.R:166-170
    for (int i = 0; i <= n_ahead; ++i) {
      double sum = 0.0;
      for (int c = 0; c < k; ++c) sum += host_out[(size_t)i * k + c];
      for (int c = 0; c < k; ++c) host_out[(size_t)i * k + c] /= sum;
    }
  return 0;
}

int bvg_sampler_fevd(bvg_sampler* s, int32_t p, int32_t n_ahead, int32_t type, int32_t response, int32_t normalise_gir,
                     int32_t period, double* host_out, void* cuda_stream) {
  if (!s) return fail(-1, "sampler is NULL");
  CK(cudaSetDevice(s->device));
  CK(after_run(s, cuda_stream));
  const int k = s->P.k, T = s->P.T;
  if (p < 0 || (long long)k * k * p > (long long)s->P.n) return fail(-2, "p lags do not fit the coefficient vector");
  const bool tv_coef = s->P.tvp != 0, tv_sig = tv_sigma_of(s->P);
  if (tv_coef || tv_sig) {
    if (period < 0) period = T - 1;
    if (period >= T) return fail(-11, "Implausible specification of argument 'period'.");
  }
  const double* coef = s->d_coef + (tv_coef ? (size_t)period * (s->P.structural ? s->P.n_user : s->P.n) : 0);
  const double* sig = s->d_sigma + (tv_sig ? (size_t)period * k * k : 0);
  const double* a0 = nullptr;
  if (type == 3 || type == 4) {
    if (!s->P.structural) return fail(-12, "Structural FEVD requires that draws of 'A0' are contained in the 'bvar' object.");
    a0 = coef + s->P.a0_from;                    // (period offset included for TVP draws)
  }
  return bvg_fevd_draws_a0(coef, s->rows_coef, sig, s->rows_sigma, a0, s->rows_coef, 1, (int64_t)s->n_chains * s->P.iterations,
                           k, p, n_ahead, type, response, normalise_gir, host_out, cuda_stream);
}

int bvg_forecast_draws(const double* dev_coef, int64_t coef_stride, const double* dev_sigma, int64_t sigma_stride,
                       int64_t n_draws, int32_t k, int32_t p, int32_t tot, int32_t n_ahead, const double* pred,
                       const double* dev_eps, uint64_t seed, double* dev_out, void* cuda_stream) {
  return bvg_forecast_draws_a0(dev_coef, coef_stride, dev_sigma, sigma_stride, nullptr, 0, 0, n_draws, k, p, tot, n_ahead, pred,
                               dev_eps, seed, dev_out, cuda_stream);
}

int bvg_forecast_draws_a0(const double* dev_coef, int64_t coef_stride, const double* dev_sigma, int64_t sigma_stride,
                          const double* dev_a0, int64_t a0_stride, int32_t a0_compact, int64_t n_draws, int32_t k, int32_t p,
                          int32_t tot, int32_t n_ahead, const double* pred, const double* dev_eps, uint64_t seed,
                          double* dev_out, void* cuda_stream) {
  if (dev_a0 && !a0_compact && k > 16) return fail(-2, "forecasts from full A0 matrices need k <= 16");
  if (!dev_sigma || !pred || !dev_out || n_draws < 1) return fail(-1, "invalid arguments to bvg_forecast_draws");
  if (k < 1 || k > 32 || p < 0 || n_ahead < 1 || tot < k || (p > 0 && tot < k * p))
    return fail(-2, "bvg_forecast_draws: need 1 <= k <= 32, n_ahead >= 1, tot >= k max(p, 1)");
  const int na = (p == 0) ? tot - k : tot;                     // columns of A = cbind(A, B, C) (R/predict.bvar.R:70-171)
  const int use_a = (na > 0 && dev_coef != nullptr) ? 1 : 0;
  if (na > 0 && !dev_coef) return fail(-1, "coefficient draws are missing");
  if (bvg_device_count() == 0) return fail(-50, "no CUDA device available");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  double* d_pred = nullptr;
  const size_t pb = (size_t)tot * (n_ahead + 1) * sizeof(double);
  CK(cudaMalloc((void**)&d_pred, pb));
  cudaError_t e = cudaMemcpyAsync(d_pred, pred, pb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = launch_forecast_draws(dev_coef, coef_stride, dev_sigma, sigma_stride, n_draws, k, p, tot, n_ahead, use_a, d_pred,
                              dev_eps, seed, dev_out, st, dev_a0, a0_stride, a0_compact ? 2 : 1);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_pred);
  if (e != cudaSuccess) return cuda_fail(e, "forecast kernel");
  return 0;
}

int bvg_sampler_forecast(bvg_sampler* s, int32_t p, int32_t tot, int32_t n_ahead, const double* pred, const double* host_eps,
                         uint64_t seed, const double* probs, int32_t n_probs, double* quantiles, double* mean,
                         double* host_draws, void* cuda_stream) {
  if (!s) return fail(-1, "sampler is NULL");
  CK(cudaSetDevice(s->device));
  CK(after_run(s, cuda_stream));
  const int k = s->P.k, T = s->P.T;
  const bool tv_coef = s->P.tvp != 0, tv_sig = tv_sigma_of(s->P);
  const long long nd = (long long)s->n_chains * s->P.iterations;
  const double* coef = s->d_coef + (tv_coef ? (size_t)(T - 1) * (s->P.structural ? s->P.n_user : s->P.n) : 0);      // last period (R/predict.bvar.R:73,85,146)
  const double* sig = s->d_sigma + (tv_sig ? (size_t)(T - 1) * k * k : 0);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const size_t ob = (size_t)nd * n_ahead * k * sizeof(double);
  double *d_out = nullptr, *d_eps = nullptr;
  CK(cudaMalloc((void**)&d_out, ob));
  int rc = 0;
  if (host_eps) {
    cudaError_t e = cudaMalloc((void**)&d_eps, ob);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_eps, host_eps, ob, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) rc = cuda_fail(e, "forecast eps upload");
  }
  if (rc == 0)
    rc = bvg_forecast_draws_a0(coef, s->rows_coef, sig, s->rows_sigma,
                               s->P.structural ? coef + s->P.a0_from : nullptr, s->rows_coef, 1, nd, k, p, tot, n_ahead,
                               pred, d_eps, seed, d_out, cuda_stream);    // structural: y = A0^-1 (A x + u), predict.bvar.R:200-209
  if (rc == 0 && (n_probs > 0 || mean))
    rc = bvg_summary_draws(d_out, nd, (int64_t)n_ahead * k, probs, n_probs, mean, nullptr, quantiles, cuda_stream);
  if (rc == 0 && host_draws) {
    cudaError_t e = cudaMemcpyAsync(host_draws, d_out, ob, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = cuda_fail(e, "forecast copy");
  }
  cudaFree(d_out);
  cudaFree(d_eps);
  return rc;
}

int bvg_fp64_peak(double* tflops_fma, void* cuda_stream) {
  if (!tflops_fma) return fail(-1, "output pointer is NULL");
  if (bvg_device_count() == 0) return fail(-50, "no CUDA device available");
  cudaError_t e = launch_fp64_peak(tflops_fma, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "fp64 peak probe");
  return 0;
}

int bvg_dmma_peak(double* tflops, void* cuda_stream) {
  if (!tflops) return fail(-1, "output pointer is NULL");
  if (bvg_device_count() == 0) return fail(-50, "no CUDA device available");
  cudaError_t e = launch_dmma_peak(tflops, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "dmma peak probe");
  return 0;
}

}  // extern "C"
