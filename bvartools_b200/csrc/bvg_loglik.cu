// bvg_loglik.cu -- log-likelihood of posterior draws for the information criteria of summary.bvarlist
// (R/summary.bvarlist.R:160-186: LL[t, j] = loglik_normal(u_j, Sigma_j)[t], ll = sum(rowMeans(LL)); AIC / BIC / HQ from
// ll) and the stand-alone block loglik_normal (src/loglik_normal.cpp:37-58).  In R this is a loop over every draw with
// one .Call per draw (det + solve of Sigma per period); here it is one pass over the draws where they already are --
// in HBM -- reduced to T per-period sums, the only thing the criteria need (and the only thing a multi-GPU run has to
// all-reduce).
//
// loglik_draws_kernel: one warp per draw (grid-stride), lanes over periods.  Per draw: A (k x M) staged in shared
// memory, Sigma = L L' factorised once (constant Sigma) or per period and lane (time-varying Sigma);
//   LL[t] = -k/2 ln(2 pi) - ln|Sigma_t|/2 - u_t' Sigma_t^-1 u_t / 2,  u_t = y_t - A_t x_t,  |Sigma| = prod L_ii^2,
//   u' Sigma^-1 u = |L^-1 u|^2.
// Per-period sums are kept in registers across the warp's draws, written once per warp to a partial buffer and reduced
// in a fixed order by loglik_reduce_kernel (bit-reproducible, no atomics).  HBM-bound by design: 8 (n + k^2) bytes per
// draw are read once.
#include "bvg_kernels.cuh"

namespace bvg {

static constexpr int LL_RMAX = 8;            // periods per lane kept in registers: T <= 256 per pass
static constexpr int LL_WARPS = 8;

struct LoglikDims {
  int k, M, T, n, tvp, tv_sigma, stage_data;
  long long n_draws, coef_stride, sigma_stride;
};

__global__ void __launch_bounds__(LL_WARPS * 32) loglik_draws_kernel(const __grid_constant__ LoglikDims d,
                                                                     const double* __restrict__ Y,
                                                                     const double* __restrict__ X,
                                                                     const double* __restrict__ coef,
                                                                     const double* __restrict__ sigma,
                                                                     double* __restrict__ partial, int t0) {
  extern __shared__ double sm[];
  const int k = d.k, M = d.M, T = d.T, n = d.n, trik = (k * (k + 1)) / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // shared layout: [Ys (k T) | Xs (M T)] (period-contiguous: lanes read consecutive addresses), then per warp:
  // Aw (n, constant-parameter models) | Lw (tri(k), or 32 tri(k) with time-varying Sigma) | uw (32 k)
  double* Ys = sm;
  double* Xs = Ys + (d.stage_data ? k * T : 0);
  double* per = Xs + (d.stage_data ? M * T : 0);
  const int per_len = (d.tvp ? 0 : n) + (d.tv_sigma ? 32 * trik : trik) + 32 * k;
  double* Aw = per + (size_t)warp * per_len;
  double* Lw = Aw + (d.tvp ? 0 : n);
  double* uw = Lw + (d.tv_sigma ? 32 * trik : trik);
  if (d.stage_data) {
    for (int e = threadIdx.x; e < k * T; e += blockDim.x) { const int i = e % k, t = e / k; Ys[i * T + t] = Y[e]; }
    for (int e = threadIdx.x; e < M * T; e += blockDim.x) { const int j = e % M, t = e / M; Xs[j * T + t] = X[e]; }
  }
  __syncthreads();
  const double part_a = -(double)k * log(2.0 * 3.14159265358979323846) / 2.0;      // loglik_normal.cpp:44
  double acc[LL_RMAX];
#pragma unroll
  for (int r = 0; r < LL_RMAX; ++r) acc[r] = 0.0;
  const int nwarp = blockDim.x >> 5;
  const long long wglobal = (long long)blockIdx.x * nwarp + warp, nw = (long long)gridDim.x * nwarp;
  for (long long dr = wglobal; dr < d.n_draws; dr += nw) {
    const double* cf = coef + dr * d.coef_stride;
    const double* sg = sigma + dr * d.sigma_stride;
    double half_logdet = 0.0;
    if (!d.tvp) for (int e = lane; e < n; e += 32) Aw[e] = cf[e];
    if (!d.tv_sigma) {
      // Sigma = L L' (lower, packed by rows), every lane the same values; lane 0 stores
      // (k is small: the factorisation is a few dozen dependent operations next to k M T products)
      if (lane == 0) {
        for (int i = 0; i < k; ++i)
          for (int j = 0; j <= i; ++j) {
            double s = sg[i + k * j];
            for (int c = 0; c < j; ++c) s = fma(-Lw[(i * (i + 1)) / 2 + c], Lw[(j * (j + 1)) / 2 + c], s);
            Lw[(i * (i + 1)) / 2 + j] = (i == j) ? sqrt(s) : s / Lw[(j * (j + 1)) / 2 + j];
          }
      }
      __syncwarp();
      for (int i = 0; i < k; ++i) half_logdet += log(Lw[(i * (i + 1)) / 2 + i]);
    } else {
      __syncwarp();
    }
#pragma unroll
    for (int r = 0; r < LL_RMAX; ++r) {
      const int t = t0 + lane + 32 * r;
      if (t < T) {
        const double* At = d.tvp ? cf + (size_t)t * n : Aw;
        double* u = uw + lane;                       // u_i at u[32 i]
        for (int i = 0; i < k; ++i) {
          double s = d.stage_data ? Ys[i * T + t] : Y[i + k * t];
          for (int j = 0; j < M; ++j) s = fma(-At[i + k * j], d.stage_data ? Xs[j * T + t] : X[j + M * t], s);
          u[32 * i] = s;                                                          // summary.bvarlist.R:165-169
        }
        const double* L = Lw;
        double hl = half_logdet;
        if (d.tv_sigma) {
          double* Lt = Lw + lane;                    // this lane's factor, element e at Lt[32 e]
          const double* st = sg + (size_t)t * k * k;
          hl = 0.0;
          for (int i = 0; i < k; ++i)
            for (int j = 0; j <= i; ++j) {
              double s = st[i + k * j];
              for (int c = 0; c < j; ++c) s = fma(-Lt[32 * ((i * (i + 1)) / 2 + c)], Lt[32 * ((j * (j + 1)) / 2 + c)], s);
              const double v = (i == j) ? sqrt(s) : s / Lt[32 * ((j * (j + 1)) / 2 + j)];
              Lt[32 * ((i * (i + 1)) / 2 + j)] = v;
              if (i == j) hl += log(v);
            }
          double quad = 0.0;
          for (int i = 0; i < k; ++i) {
            double s = u[32 * i];
            for (int c = 0; c < i; ++c) s = fma(-Lt[32 * ((i * (i + 1)) / 2 + c)], u[32 * c], s);
            s /= Lt[32 * ((i * (i + 1)) / 2 + i)];
            u[32 * i] = s;
            quad = fma(s, s, quad);
          }
          acc[r] += part_a - hl - 0.5 * quad;
        } else {
          double quad = 0.0;
          for (int i = 0; i < k; ++i) {
            double s = u[32 * i];
            for (int c = 0; c < i; ++c) s = fma(-L[(i * (i + 1)) / 2 + c], u[32 * c], s);
            s /= L[(i * (i + 1)) / 2 + i];
            u[32 * i] = s;
            quad = fma(s, s, quad);
          }
          acc[r] += part_a - hl - 0.5 * quad;                                     // loglik_normal.cpp:46-56
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < LL_RMAX; ++r) {
    const int t = t0 + lane + 32 * r;
    if (t < T) partial[wglobal * T + t] = acc[r];
  }
}

// ll_t[t] = sum over the warps' partial sums, in warp order (one thread per period)
__global__ void loglik_reduce_kernel(const double* __restrict__ partial, long long n_warps, int T, double* __restrict__ ll_t) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  double s = 0.0;
  for (long long w = 0; w < n_warps; ++w) s += partial[w * T + t];
  ll_t[t] = s;
}

cudaError_t launch_loglik_draws(int k, int M, int T, int tvp, int tv_sigma, long long n_draws, const double* dY,
                                const double* dX, const double* d_coef, const double* d_sigma, double* d_ll_t,
                                cudaStream_t stream) {
  if (k < 1 || k > 32 || M < 0 || T < 1 || n_draws < 1) return cudaErrorInvalidValue;
  LoglikDims d;
  d.k = k; d.M = M; d.T = T; d.n = k * M; d.tvp = tvp ? 1 : 0; d.tv_sigma = tv_sigma ? 1 : 0;
  d.n_draws = n_draws;
  d.coef_stride = (long long)d.n * (tvp ? T : 1);
  d.sigma_stride = (long long)k * k * (tv_sigma ? T : 1);
  const int trik = (k * (k + 1)) / 2;
  const size_t per_warp = (size_t)((tvp ? 0 : d.n) + (tv_sigma ? 32 * trik : trik) + 32 * k);
  const size_t data = (size_t)(k + M) * T;
  // warps per CTA: as many as the per-warp scratch allows (time-varying Sigma with large k needs 32 tri(k) doubles per warp)
  int warps = LL_WARPS;
  while (warps > 1 && (size_t)warps * per_warp * 8 > 200 * 1024) warps >>= 1;
  if ((size_t)warps * per_warp * 8 > 220 * 1024) return cudaErrorInvalidConfiguration;
  d.stage_data = ((data + warps * per_warp) * 8 <= 160 * 1024) ? 1 : 0;
  const size_t smem = ((d.stage_data ? data : 0) + warps * per_warp) * 8;
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(loglik_draws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  long long want = (n_draws + warps - 1) / warps;
  int grid = (int)((want < (long long)sms * per_sm) ? want : (long long)sms * per_sm);
  if (grid < 1) grid = 1;
  const long long n_warps = (long long)grid * warps;
  double* partial = nullptr;
  e = cudaMallocAsync((void**)&partial, (size_t)n_warps * T * sizeof(double), stream);
  if (e != cudaSuccess) return e;
  cudaMemsetAsync(partial, 0, (size_t)n_warps * T * sizeof(double), stream);
  for (int t0 = 0; t0 < T; t0 += 32 * LL_RMAX)     // T > 256: further passes over the draws
    loglik_draws_kernel<<<grid, warps * 32, smem, stream>>>(d, dY, dX, d_coef, d_sigma, partial, t0);
  loglik_reduce_kernel<<<(T + 127) / 128, 128, 0, stream>>>(partial, n_warps, T, d_ll_t);
  e = cudaGetLastError();
  cudaFreeAsync(partial, stream);
  return e;
}

}  // namespace bvg
