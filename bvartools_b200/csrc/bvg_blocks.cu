// bvg_blocks.cu -- CUDA kernels + C ABI of the batched stand-alone building blocks (include/bvargpu.h):
// bvg_post_normal, bvg_post_normal_sur, bvg_ssvs, bvg_bvs, bvg_stochvol, bvg_kalman_dk.
// One CTA per problem of the batch (ssvs: one thread per coefficient); host buffers in, host buffers out.
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "bvg_blocks.cuh"
#include "bvg_kernels.cuh"
#include "bvg_prepare.hpp"

using namespace bvg;

namespace {

thread_local std::string g_err;
int fail(int code, const char* msg) { g_err = msg; return code; }

struct DevBuf {   // RAII device buffer, optionally initialised from host
  void* p = nullptr;
  cudaError_t err = cudaSuccess;
  DevBuf() {}
  DevBuf(const void* host, size_t bytes) { alloc(host, bytes); }
  void alloc(const void* host, size_t bytes) {
    if (bytes == 0) return;
    err = cudaMalloc(&p, bytes);
    if (err == cudaSuccess && host != nullptr) err = cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice);
  }
  ~DevBuf() { if (p) cudaFree(p); }
  template <class T> T* as() const { return (T*)p; }
};

RngParams make_rp(uint64_t seed) {
  RngParams rp;
  memset(&rp, 0, sizeof(rp));
  rp.seed = seed;
  return rp;
}
void inject(RngParams& rp, int block, const double* dev, int count) {
  rp.inj[block] = dev;
  rp.inj_stride[block] = count;
  rp.inj_chain_stride[block] = count;
}

__global__ void post_normal_kernel(const __grid_constant__ RngParams rp, int k, int M, const double* XX,
                                   const double* YX, const double* sig, const double* mu, const double* Vi,
                                   double* out, int method, int* status) {
  extern __shared__ double smem[];
  BlockGroup g(smem);
  const long long b = blockIdx.x;
  const int n = k * M;
  const ChainRng rng(&rp, b);
  int fl = 0;
  post_normal_problem(g, rng, k, M, XX, YX, sig + (size_t)b * k * k, mu, Vi, smem + BVG_BLOCK_SCRATCH,
                      out + (size_t)b * n, method, fl);
  if (fl && threadIdx.x == 0) status[b] = 1;
}

__global__ void post_normal_sur_kernel(const __grid_constant__ RngParams rp, int k, int T, int n, const double* y,
                                       const double* z, const double* sig, int sigma_tv, const double* mu,
                                       const double* Vi, double* out, int method, int* status) {
  extern __shared__ double smem[];
  BlockGroup g(smem);
  const long long b = blockIdx.x;
  const ChainRng rng(&rp, b);
  int fl = 0;
  const size_t sstride = sigma_tv ? (size_t)k * T * k : (size_t)k * k;
  post_normal_sur_problem(g, rng, k, T, n, y, z, sig + (size_t)b * sstride, sigma_tv, mu, Vi,
                          smem + BVG_BLOCK_SCRATCH, out + (size_t)b * n, method, fl);
  if (fl && threadIdx.x == 0) status[b] = 1;
}

// ssvs (svss.cpp:78-108): element-parallel; ~48 B read + 16 B written per coefficient
__global__ void __launch_bounds__(256) ssvs_kernel(const __grid_constant__ RngParams rp, long long B, int n,
                                                   const double* __restrict__ a, const double* __restrict__ tau0,
                                                   const double* __restrict__ tau1, const double* __restrict__ prob,
                                                   const unsigned char* __restrict__ include, double* lam,
                                                   double* vdiag) {
  const long long total = B * n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / n;
    const int r = (int)(e % n);
    const ChainRng rng(&rp, b);
    double ld, v;
    ssvs_element(rng, r, a[e], tau0[r], tau1[r], prob[r], include[r] != 0, &ld, &v);
    lam[e] = ld;
    vdiag[e] = v;
  }
}

__global__ void bvs_kernel(const __grid_constant__ RngParams rp, int k, int T, int n, const double* y,
                           const double* z, const double* a, int a_tv, const double* lam, const double* sig,
                           int sigma_tv, const double* prob, const unsigned char* include, double* lam_out) {
  extern __shared__ double smem[];
  BlockGroup g(smem);
  const long long b = blockIdx.x;
  const ChainRng rng(&rp, b);
  const size_t astride = a_tv ? (size_t)n * T : (size_t)n;
  const size_t sstride = sigma_tv ? (size_t)k * T * k : (size_t)k * k;
  bvs_problem(g, rng, k, T, n, y, z, a + (size_t)b * astride, a_tv, lam + (size_t)b * n, sig + (size_t)b * sstride,
              sigma_tv, prob, include, smem + BVG_BLOCK_SCRATCH, lam_out + (size_t)b * n);
}

__global__ void stochvol_kernel(const __grid_constant__ RngParams rp, const __grid_constant__ BvarModel proto,
                                const double* y, const double* h_in, const double* sigma, const double* h_init,
                                const double* constant, double* h_out, int* s_out) {
  extern __shared__ double smem[];
  BlockGroup g(smem);
  const long long b = blockIdx.x;
  const int k = proto.k, T = proto.T;
  double* U = smem + BVG_BLOCK_SCRATCH;
  double* h = U + k * T;
  double* pr = h + k * T;
  double* rh = pr + k * T;
  BvarModel m = proto;
  m.sv_const = constant + (size_t)b * k;
  for (int e = threadIdx.x; e < k * T; e += blockDim.x) {
    const int i = e % k, t = e / k;
    U[e] = y[(size_t)b * k * T + t + (size_t)T * i];          // y is T x k col-major per problem
    h[t + T * i] = h_in[(size_t)b * k * T + t + (size_t)T * i];
  }
  __syncthreads();
  const ChainRng rng(&rp, b);
  sv_step(g, rng, 0, m, U, h, sigma + (size_t)b * k, h_init + (size_t)b * k, pr, rh,
          s_out ? s_out + (size_t)b * k * T : nullptr);
  for (int e = threadIdx.x; e < k * T; e += blockDim.x) h_out[(size_t)b * k * T + e] = h[e];
}

__global__ void kalman_dk_kernel(const __grid_constant__ RngParams rp, int k, int T, int n, const double* y,
                                 const double* z, const double* su, const double* sv, const double* Bm, int tv,
                                 const double* a_init, const double* P_init, double* scratch, long long spp,
                                 double* out, int* status) {
  extern __shared__ double smem[];
  BlockGroup g(smem);
  const long long b = blockIdx.x;
  const ChainRng rng(&rp, b);
  int fl = 0;
  const size_t us = (tv & 1) ? (size_t)k * T * k : (size_t)k * k;
  const size_t vs = (tv & 2) ? (size_t)n * T * n : (size_t)n * n;
  const size_t bs = (tv & 4) ? (size_t)n * T * n : (size_t)n * n;
  kalman_dk_problem(g, rng, k, T, n, y, z, su + (size_t)b * us, sv + (size_t)b * vs, Bm + (size_t)b * bs, tv,
                    a_init + (size_t)b * n, P_init + (size_t)b * n * n, smem + BVG_BLOCK_SCRATCH,
                    scratch + (size_t)b * spp, out + (size_t)b * n * (T + 1), fl);
  if (fl && threadIdx.x == 0) status[b] = 1;
}

template <typename K>
int prep_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) return fail(-52, "problem does not fit in shared memory (n too large for this block)");
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
  }
  return 0;
}

int finish(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { g_err = std::string(what) + ": " + cudaGetErrorString(e); return (int)e; }
  return 0;
}

int need_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(-50, "no CUDA device available: libbvargpu has no CPU fallback");
  }
  return 0;
}

#define CKB(buf) do { if ((buf).err != cudaSuccess) return fail((int)(buf).err, cudaGetErrorString((buf).err)); } while (0)

std::vector<unsigned char> include_mask(int n, const int32_t* include, int n_include, bool* ok) {
  std::vector<unsigned char> m(n, include == nullptr ? 1 : 0);
  *ok = true;
  if (include != nullptr)
    for (int i = 0; i < n_include; ++i) {
      if (include[i] < 0 || include[i] >= n) { *ok = false; break; }
      m[include[i]] = 1;
    }
  return m;
}

}  // namespace

extern "C" const char* bvg_blocks_last_error_internal(void) { return g_err.c_str(); }

extern "C" {

int bvg_post_normal(int64_t B, int k, int T, int M, const double* y, const double* x, const double* sigma_i,
                    const double* a_prior, const double* v_i_prior, const double* eps, uint64_t seed,
                    int method, double* out) {
  if (B < 1 || k < 1 || T < 1 || M < 1 || !y || !x || !sigma_i || !a_prior || !v_i_prior || !out)
    return fail(-1, "invalid arguments to bvg_post_normal");
  const int n = k * M;
  if (has_nan(sigma_i, (size_t)B * k * k)) return fail(-6, "Argument 'sigma_i' contains NAs.");   // post_normal.cpp:63-65
  int rc = need_device();
  if (rc) return rc;
  std::vector<double> XX((size_t)M * M, 0.0), YX((size_t)k * M, 0.0);
  for (int t = 0; t < T; ++t)
    for (int j = 0; j < M; ++j) {
      for (int j2 = 0; j2 < M; ++j2) XX[j + (size_t)M * j2] += x[j + (size_t)M * t] * x[j2 + (size_t)M * t];
      for (int i = 0; i < k; ++i) YX[i + (size_t)k * j] += y[i + (size_t)k * t] * x[j + (size_t)M * t];
    }
  DevBuf dXX(XX.data(), XX.size() * 8), dYX(YX.data(), YX.size() * 8), dS(sigma_i, (size_t)B * k * k * 8),
      dmu(a_prior, (size_t)n * 8), dVi(v_i_prior, (size_t)n * n * 8), dout(nullptr, (size_t)B * n * 8),
      deps(eps, eps ? (size_t)B * n * 8 : 0), dst(nullptr, (size_t)B * 4);
  CKB(dXX); CKB(dYX); CKB(dS); CKB(dmu); CKB(dVi); CKB(dout); CKB(deps); CKB(dst);
  cudaMemset(dst.p, 0, (size_t)B * 4);
  RngParams rp = make_rp(seed);
  if (eps) inject(rp, VB_COEF_N, deps.as<double>(), n);
  const size_t smem = ((size_t)post_normal_smem_len(n) + BVG_BLOCK_SCRATCH) * 8;
  if ((rc = prep_smem(post_normal_kernel, smem))) return rc;
  post_normal_kernel<<<(unsigned)B, 128, smem>>>(rp, k, M, dXX.as<double>(), dYX.as<double>(), dS.as<double>(),
                                                dmu.as<double>(), dVi.as<double>(), dout.as<double>(), method,
                                                dst.as<int>());
  if ((rc = finish("post_normal kernel"))) return rc;
  cudaMemcpy(out, dout.p, (size_t)B * n * 8, cudaMemcpyDeviceToHost);
  std::vector<int> st(B);
  cudaMemcpy(st.data(), dst.p, (size_t)B * 4, cudaMemcpyDeviceToHost);
  for (int v : st) if (v) return fail(1, "posterior precision is not positive definite");
  return 0;
}

int bvg_post_normal_sur(int64_t B, int k, int T, int n, const double* y, const double* z, const double* sigma_i,
                        int sigma_tv, const double* a_prior, const double* v_i_prior, const double* eps,
                        uint64_t seed, int method, double* out) {
  if (B < 1 || k < 1 || T < 1 || n < 1 || !y || !z || !sigma_i || !a_prior || !v_i_prior || !out)
    return fail(-1, "invalid arguments to bvg_post_normal_sur");
  const size_t sstride = sigma_tv ? (size_t)k * T * k : (size_t)k * k;
  if (has_nan(sigma_i, (size_t)B * sstride)) return fail(-6, "Argument 'sigma_i' contains NAs.");
  int rc = need_device();
  if (rc) return rc;
  DevBuf dy(y, (size_t)k * T * 8), dz(z, (size_t)k * T * n * 8), dS(sigma_i, (size_t)B * sstride * 8),
      dmu(a_prior, (size_t)n * 8), dVi(v_i_prior, (size_t)n * n * 8), dout(nullptr, (size_t)B * n * 8),
      deps(eps, eps ? (size_t)B * n * 8 : 0), dst(nullptr, (size_t)B * 4);
  CKB(dy); CKB(dz); CKB(dS); CKB(dmu); CKB(dVi); CKB(dout); CKB(deps); CKB(dst);
  cudaMemset(dst.p, 0, (size_t)B * 4);
  RngParams rp = make_rp(seed);
  if (eps) inject(rp, VB_COEF_N, deps.as<double>(), n);
  const size_t smem = ((size_t)post_normal_smem_len(n) + BVG_BLOCK_SCRATCH) * 8;
  if ((rc = prep_smem(post_normal_sur_kernel, smem))) return rc;
  post_normal_sur_kernel<<<(unsigned)B, 128, smem>>>(rp, k, T, n, dy.as<double>(), dz.as<double>(), dS.as<double>(),
                                                    sigma_tv, dmu.as<double>(), dVi.as<double>(), dout.as<double>(),
                                                    method, dst.as<int>());
  if ((rc = finish("post_normal_sur kernel"))) return rc;
  cudaMemcpy(out, dout.p, (size_t)B * n * 8, cudaMemcpyDeviceToHost);
  std::vector<int> st(B);
  cudaMemcpy(st.data(), dst.p, (size_t)B * 4, cudaMemcpyDeviceToHost);
  for (int v : st) if (v) return fail(1, "posterior precision is not positive definite");
  return 0;
}

int bvg_ssvs(int64_t B, int n, const double* a, const double* tau0, const double* tau1, const double* prob_prior,
             const int32_t* include, int n_include, const double* u, uint64_t seed, double* lambda_out,
             double* vdiag_out) {
  if (B < 1 || n < 1 || !a || !tau0 || !tau1 || !prob_prior || !lambda_out || !vdiag_out)
    return fail(-1, "invalid arguments to bvg_ssvs");
  bool ok;
  std::vector<unsigned char> mask = include_mask(n, include, n_include, &ok);
  if (!ok) return fail(-8, "include index out of range");
  int rc = need_device();
  if (rc) return rc;
  DevBuf da(a, (size_t)B * n * 8), dt0(tau0, (size_t)n * 8), dt1(tau1, (size_t)n * 8), dp(prob_prior, (size_t)n * 8),
      dm(mask.data(), (size_t)n), du(u, u ? (size_t)B * n * 8 : 0), dl(nullptr, (size_t)B * n * 8),
      dv(nullptr, (size_t)B * n * 8);
  CKB(da); CKB(dt0); CKB(dt1); CKB(dp); CKB(dm); CKB(du); CKB(dl); CKB(dv);
  RngParams rp = make_rp(seed);
  if (u) inject(rp, VB_SSVS_U, du.as<double>(), n);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = (B * n + 255) / 256;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  ssvs_kernel<<<(unsigned)blocks, 256>>>(rp, B, n, da.as<double>(), dt0.as<double>(), dt1.as<double>(),
                                         dp.as<double>(), dm.as<unsigned char>(), dl.as<double>(), dv.as<double>());
  if ((rc = finish("ssvs kernel"))) return rc;
  cudaMemcpy(lambda_out, dl.p, (size_t)B * n * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(vdiag_out, dv.p, (size_t)B * n * 8, cudaMemcpyDeviceToHost);
  return 0;
}

int bvg_bvs(int64_t B, int k, int T, int n, const double* y, const double* z, const double* a, int a_tv,
            const double* lambda, const double* sigma_i, int sigma_tv, const double* prob_prior,
            const int32_t* include, int n_include, const double* u1, const double* u2, uint64_t seed,
            double* lambda_out) {
  if (B < 1 || k < 1 || T < 1 || n < 1 || !y || !z || !a || !lambda || !sigma_i || !prob_prior || !lambda_out)
    return fail(-1, "invalid arguments to bvg_bvs");
  bool ok;
  std::vector<unsigned char> mask = include_mask(n, include, n_include, &ok);
  if (!ok) return fail(-8, "include index out of range");
  int rc = need_device();
  if (rc) return rc;
  const size_t kT = (size_t)k * T;
  const size_t astride = a_tv ? (size_t)n * T : (size_t)n, sstride = sigma_tv ? kT * k : (size_t)k * k;
  DevBuf dy(y, kT * 8), dz(z, kT * n * 8), da(a, (size_t)B * astride * 8), dl(lambda, (size_t)B * n * 8),
      dS(sigma_i, (size_t)B * sstride * 8), dp(prob_prior, (size_t)n * 8), dm(mask.data(), (size_t)n),
      du1(u1, u1 ? (size_t)B * n * 8 : 0), du2(u2, u2 ? (size_t)B * n * 8 : 0), dout(nullptr, (size_t)B * n * 8);
  CKB(dy); CKB(dz); CKB(da); CKB(dl); CKB(dS); CKB(dp); CKB(dm); CKB(du1); CKB(du2); CKB(dout);
  RngParams rp = make_rp(seed);
  if (u1) inject(rp, VB_BVS_U1, du1.as<double>(), n);
  if (u2) inject(rp, VB_BVS_U2, du2.as<double>(), n);
  const size_t smem = (2 * kT + BVG_BLOCK_SCRATCH) * 8;
  if ((rc = prep_smem(bvs_kernel, smem))) return rc;
  bvs_kernel<<<(unsigned)B, 128, smem>>>(rp, k, T, n, dy.as<double>(), dz.as<double>(), da.as<double>(), a_tv,
                                        dl.as<double>(), dS.as<double>(), sigma_tv, dp.as<double>(),
                                        dm.as<unsigned char>(), dout.as<double>());
  if ((rc = finish("bvs kernel"))) return rc;
  cudaMemcpy(lambda_out, dout.p, (size_t)B * n * 8, cudaMemcpyDeviceToHost);
  return 0;
}

int bvg_stochvol(int64_t B, int T, int k, const double* y, const double* h, const double* sigma,
                 const double* h_init, const double* constant, int mixture, const double* u, const double* eps,
                 uint64_t seed, double* h_out, int32_t* s_out) {
  if (B < 1 || T < 2 || k < 1 || !y || !h || !sigma || !h_init || !constant || !h_out)
    return fail(-1, "invalid arguments to bvg_stochvol");
  if (has_nan(y, (size_t)B * T * k)) return fail(-6, "Argument 'y' contains NAs.");     // stochvol_ksc1998.cpp:63-65
  int rc = need_device();
  if (rc) return rc;
  const int J = (mixture == BVG_MIX_OCSN2007) ? 10 : 7;
  std::vector<double> tab(3 * 10, 0.0);
  for (int j = 0; j < J; ++j) {
    tab[j] = (J == 7) ? KSC_P[j] : OCSN_P[j];
    tab[10 + j] = (J == 7) ? KSC_MU[j] - 1.2704 : OCSN_MU[j];
    tab[20 + j] = (J == 7) ? KSC_S2[j] : OCSN_S2[j];
  }
  const size_t kt = (size_t)k * T;
  DevBuf dtab(tab.data(), tab.size() * 8), dy(y, B * kt * 8), dh(h, B * kt * 8), dsg(sigma, (size_t)B * k * 8),
      dhi(h_init, (size_t)B * k * 8), dc(constant, (size_t)B * k * 8), du(u, u ? B * kt * 8 : 0),
      de(eps, eps ? B * kt * 8 : 0), dout(nullptr, B * kt * 8), ds(nullptr, s_out ? B * kt * 4 : 0);
  CKB(dtab); CKB(dy); CKB(dh); CKB(dsg); CKB(dhi); CKB(dc); CKB(du); CKB(de); CKB(dout); CKB(ds);
  RngParams rp = make_rp(seed);
  if (u) inject(rp, VB_SV_U, du.as<double>(), (int)kt);
  if (eps) inject(rp, VB_SV_N, de.as<double>(), (int)kt);
  BvarModel proto;
  memset(&proto, 0, sizeof(proto));
  proto.k = k; proto.T = T; proto.n_mix = J; proto.sigma_type = SIG_SV; proto.a0_from = -1;
  proto.mix_p = dtab.as<double>(); proto.mix_mu = dtab.as<double>() + 10; proto.mix_s2 = dtab.as<double>() + 20;
  const size_t smem = (4 * kt + BVG_BLOCK_SCRATCH) * 8;
  if ((rc = prep_smem(stochvol_kernel, smem))) return rc;
  stochvol_kernel<<<(unsigned)B, 128, smem>>>(rp, proto, dy.as<double>(), dh.as<double>(), dsg.as<double>(),
                                             dhi.as<double>(), dc.as<double>(), dout.as<double>(),
                                             s_out ? ds.as<int>() : nullptr);
  if ((rc = finish("stochvol kernel"))) return rc;
  cudaMemcpy(h_out, dout.p, B * kt * 8, cudaMemcpyDeviceToHost);
  if (s_out) cudaMemcpy(s_out, ds.p, B * kt * 4, cudaMemcpyDeviceToHost);
  return 0;
}

int bvg_kalman_dk(int64_t B, int k, int T, int n, const double* y, const double* z, const double* sigma_u,
                  const double* sigma_v, const double* Bmat, int tv_mask, const double* a_init,
                  const double* P_init, const double* eps, uint64_t seed, double* out) {
  if (B < 1 || k < 1 || T < 1 || n < 1 || !y || !z || !sigma_u || !sigma_v || !Bmat || !a_init || !P_init || !out)
    return fail(-1, "invalid arguments to bvg_kalman_dk");
  int rc = need_device();
  if (rc) return rc;
  const int neps = n + T * (k + n);
  const long long spp = kalman_scratch_len(k, T, n);
  const size_t us = (tv_mask & 1) ? (size_t)k * T * k : (size_t)k * k;
  const size_t vs = (tv_mask & 2) ? (size_t)n * T * n : (size_t)n * n;
  const size_t bs = (tv_mask & 4) ? (size_t)n * T * n : (size_t)n * n;
  DevBuf dy(y, (size_t)k * T * 8), dz(z, (size_t)k * T * n * 8), dsu(sigma_u, (size_t)B * us * 8),
      dsv(sigma_v, (size_t)B * vs * 8), dB(Bmat, (size_t)B * bs * 8), da(a_init, (size_t)B * n * 8),
      dP(P_init, (size_t)B * n * n * 8), de(eps, eps ? (size_t)B * neps * 8 : 0),
      dscr(nullptr, (size_t)B * spp * 8), dout(nullptr, (size_t)B * n * (T + 1) * 8), dst(nullptr, (size_t)B * 4);
  CKB(dy); CKB(dz); CKB(dsu); CKB(dsv); CKB(dB); CKB(da); CKB(dP); CKB(de); CKB(dscr); CKB(dout); CKB(dst);
  cudaMemset(dst.p, 0, (size_t)B * 4);
  RngParams rp = make_rp(seed);
  if (eps) inject(rp, VB_STATE_N, de.as<double>(), neps);
  const size_t smem = ((size_t)kalman_smem_len(k, n) + BVG_BLOCK_SCRATCH) * 8;
  if ((rc = prep_smem(kalman_dk_kernel, smem))) return rc;
  kalman_dk_kernel<<<(unsigned)B, 128, smem>>>(rp, k, T, n, dy.as<double>(), dz.as<double>(), dsu.as<double>(),
                                              dsv.as<double>(), dB.as<double>(), tv_mask, da.as<double>(),
                                              dP.as<double>(), dscr.as<double>(), spp, dout.as<double>(),
                                              dst.as<int>());
  if ((rc = finish("kalman_dk kernel"))) return rc;
  cudaMemcpy(out, dout.p, (size_t)B * n * (T + 1) * 8, cudaMemcpyDeviceToHost);
  std::vector<int> st(B);
  cudaMemcpy(st.data(), dst.p, (size_t)B * 4, cudaMemcpyDeviceToHost);
  for (int v : st) if (v) return fail(1, "a forecast-error covariance was not positive definite");
  return 0;
}

}  // extern "C"
