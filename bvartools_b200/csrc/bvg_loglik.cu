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

// ---- constant-parameter models: the TOTAL log-likelihood of a draw needs no pass over the periods -----------------
//   sum_t LL[t] = T part_a - T ln|Sigma| / 2 - tr(Sigma^-1 S) / 2,   S = U U' = SSE_ols + (A - A_ols) XX' (A - A_ols)'
// (the cancellation-free cross-moment form the sampler itself uses, bvg_bvar_sweep.cuh): k M (M + k) products per
// draw instead of k M T, which turns the reduction from FP64-bound into HBM-bound (8 (n + k^2) bytes per draw).
// One warp per draw (grid-stride); lanes over coefficients, then over the k x k entries, then over the k columns of the
// two triangular solves.  Partial sums per warp, reduced in fixed order by loglik_reduce_kernel (T = 1).
struct LoglikTotalDims { int k, M, T, n; long long n_draws; };

__global__ void __launch_bounds__(LL_WARPS * 32) loglik_total_kernel(const __grid_constant__ LoglikTotalDims d,
                                                                     const double* __restrict__ XX,
                                                                     const double* __restrict__ Aols,
                                                                     const double* __restrict__ SSE,
                                                                     const double* __restrict__ coef,
                                                                     const double* __restrict__ sigma,
                                                                     double* __restrict__ partial) {
  extern __shared__ double sm[];
  const int k = d.k, M = d.M, n = d.n, kk = k * k, trik = (k * (k + 1)) / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  double* XXs = sm;                      // M x M
  double* Aos = XXs + M * M;             // k x M
  double* SSs = Aos + n;                 // k x k
  double* per = SSs + kk;
  const int per_len = 2 * n + kk + trik;
  double* Dw = per + (size_t)warp * per_len;
  double* Ww = Dw + n;
  double* Sw = Ww + n;
  double* Lw = Sw + kk;
  for (int e = threadIdx.x; e < M * M; e += blockDim.x) XXs[e] = XX[e];
  for (int e = threadIdx.x; e < n; e += blockDim.x) Aos[e] = Aols[e];
  for (int e = threadIdx.x; e < kk; e += blockDim.x) SSs[e] = SSE[e];
  __syncthreads();
  const double part_a = -(double)k * log(2.0 * 3.14159265358979323846) / 2.0;
  double acc = 0.0;
  const long long wglobal = (long long)blockIdx.x * nwarp + warp, nw = (long long)gridDim.x * nwarp;
  for (long long dr = wglobal; dr < d.n_draws; dr += nw) {
    const double* cf = coef + dr * n;
    const double* sg = sigma + dr * kk;
    for (int e = lane; e < n; e += 32) Dw[e] = cf[e] - Aos[e];
    if (lane == 0) {                                            // Sigma = L L' (lower, packed by rows)
      for (int i = 0; i < k; ++i)
        for (int j = 0; j <= i; ++j) {
          double s = sg[i + k * j];
          for (int c = 0; c < j; ++c) s = fma(-Lw[(i * (i + 1)) / 2 + c], Lw[(j * (j + 1)) / 2 + c], s);
          Lw[(i * (i + 1)) / 2 + j] = (i == j) ? sqrt(s) : s / Lw[(j * (j + 1)) / 2 + j];
        }
    }
    __syncwarp();
    for (int e = lane; e < n; e += 32) {                        // W = D XX'
      const int i = e % k, j = e / k;
      double s = 0.0;
      for (int j2 = 0; j2 < M; ++j2) s = fma(Dw[i + k * j2], XXs[j2 + M * j], s);
      Ww[e] = s;
    }
    __syncwarp();
    for (int e = lane; e < kk; e += 32) {                       // S = SSE_ols + W D'
      const int a = e % k, b = e / k;
      double s = SSs[e];
      for (int j = 0; j < M; ++j) s = fma(Ww[a + k * j], Dw[b + k * j], s);
      Sw[e] = s;
    }
    __syncwarp();
    // tr(Sigma^-1 S) = sum_b [L^-T L^-1 S e_b]_b : lane b solves its column (k <= 32)
    double tr = 0.0;
    if (lane < k) {
      double z[32];
      const int b = lane;
      for (int i = 0; i < k; ++i) {
        double s = Sw[i + k * b];
        for (int c = 0; c < i; ++c) s = fma(-Lw[(i * (i + 1)) / 2 + c], z[c], s);
        z[i] = s / Lw[(i * (i + 1)) / 2 + i];
      }
      for (int i = k - 1; i >= b; --i) {
        double s = z[i];
        for (int c = i + 1; c < k; ++c) s = fma(-Lw[(c * (c + 1)) / 2 + i], z[c], s);
        z[i] = s / Lw[(i * (i + 1)) / 2 + i];
      }
      tr = z[b];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
    double half_logdet = 0.0;
    for (int i = 0; i < k; ++i) half_logdet += log(Lw[(i * (i + 1)) / 2 + i]);
    acc += (double)d.T * (part_a - half_logdet) - 0.5 * tr;
    __syncwarp();
  }
  if (lane == 0) partial[wglobal] = acc;
}

// Small models (k <= 4, n <= 64: the e1 VAR(p) grid of config c2): one THREAD per draw, k compile-time.  A CTA stages
// its 128 consecutive draws with coalesced loads (row stride padded to an odd number of doubles: conflict-free
// per-thread rows), D = A - A_ols on the fly; the k x k algebra runs in registers.  ~22 warp-instructions per draw
// instead of ~1500 for the generic warp-per-draw kernel: the pass over the draws becomes HBM-bound.
static constexpr int LLS_THREADS = 128;

template <int K>
__global__ void __launch_bounds__(LLS_THREADS) loglik_total_small_kernel(const __grid_constant__ LoglikTotalDims d,
                                                                         const double* __restrict__ XX,
                                                                         const double* __restrict__ Aols,
                                                                         const double* __restrict__ SSE,
                                                                         const double* __restrict__ coef,
                                                                         const double* __restrict__ sigma,
                                                                         double* __restrict__ partial) {
  extern __shared__ double sm[];
  const int M = d.M, n = d.n, KK = K * K;
  const int ns = n | 1, ks = KK | 1;                         // odd row strides
  double* XXs = sm;                                          // M x M
  double* Aos = XXs + M * M;                                 // n
  double* Ds = Aos + n;                                      // [128][ns]
  double* Sg = Ds + LLS_THREADS * ns;                        // [128][ks]
  double* red = Sg + LLS_THREADS * ks;                       // 128
  const int tid = threadIdx.x;
  for (int e = tid; e < M * M; e += LLS_THREADS) XXs[e] = XX[e];
  for (int e = tid; e < n; e += LLS_THREADS) Aos[e] = Aols[e];
  double S0[K][K];
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int b = 0; b < K; ++b) S0[a][b] = SSE[a + K * b];
  const double part_a = -(double)K * log(2.0 * 3.14159265358979323846) / 2.0;
  double acc = 0.0;
  const long long nblk = (d.n_draws + LLS_THREADS - 1) / LLS_THREADS;
  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const long long base = blk * LLS_THREADS;
    const int cnt = (int)((d.n_draws - base < LLS_THREADS) ? d.n_draws - base : LLS_THREADS);
    __syncthreads();                                         // previous block's rows are no longer read
    const double* cf = coef + base * n;
    for (int e = tid; e < cnt * n; e += LLS_THREADS) { const int r = e / n, c = e - r * n; Ds[r * ns + c] = cf[e] - Aos[c]; }
    const double* sg = sigma + base * KK;
    for (int e = tid; e < cnt * KK; e += LLS_THREADS) { const int r = e / KK, c = e - r * KK; Sg[r * ks + c] = sg[e]; }
    __syncthreads();
    if (tid < cnt) {
      const double* Dr = Ds + tid * ns;
      double S[K][K];
#pragma unroll
      for (int a = 0; a < K; ++a)
#pragma unroll
        for (int b = 0; b < K; ++b) S[a][b] = S0[a][b];
      for (int j = 0; j < M; ++j) {                          // S += (D XX')[:, j] D[:, j]'
        double w[K];
#pragma unroll
        for (int a = 0; a < K; ++a) w[a] = 0.0;
        const double* xc = XXs + M * j;
        for (int j2 = 0; j2 < M; ++j2) {
          const double x = xc[j2];
#pragma unroll
          for (int a = 0; a < K; ++a) w[a] = fma(Dr[a + K * j2], x, w[a]);
        }
#pragma unroll
        for (int b = 0; b < K; ++b) {
          const double db = Dr[b + K * j];
#pragma unroll
          for (int a = 0; a < K; ++a) S[a][b] = fma(w[a], db, S[a][b]);
        }
      }
      // Sigma = L L'
      const double* sr = Sg + tid * ks;
      double L[K][K], li[K];
      double half_logdet = 0.0;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        double dj = sr[j + K * j];
#pragma unroll
        for (int c = 0; c < j; ++c) dj = fma(-L[j][c], L[j][c], dj);
        const double l = sqrt(dj);
        L[j][j] = l;
        li[j] = 1.0 / l;
        half_logdet += log(l);
#pragma unroll
        for (int i = j + 1; i < K; ++i) {
          double v = sr[i + K * j];
#pragma unroll
          for (int c = 0; c < j; ++c) v = fma(-L[i][c], L[j][c], v);
          L[i][j] = v * li[j];
        }
      }
      // tr(Sigma^-1 S) = sum_b [L^-T L^-1 S e_b]_b
      double tr = 0.0;
#pragma unroll
      for (int b = 0; b < K; ++b) {
        double z[K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
          double v = S[i][b];
#pragma unroll
          for (int c = 0; c < i; ++c) v = fma(-L[i][c], z[c], v);
          z[i] = v * li[i];
        }
#pragma unroll
        for (int i = K - 1; i >= b; --i) {
          double v = z[i];
#pragma unroll
          for (int c = i + 1; c < K; ++c) v = fma(-L[c][i], z[c], v);
          z[i] = v * li[i];
        }
        tr += z[b];
      }
      acc += (double)d.T * (part_a - half_logdet) - 0.5 * tr;
    }
  }
  __syncthreads();
  red[tid] = acc;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int e = 0; e < LLS_THREADS; ++e) s += red[e];      // fixed order: reproducible
    partial[blockIdx.x] = s;
  }
}

template <int K>
static cudaError_t launch_small(const LoglikTotalDims& d, const double* dXX, const double* dAols, const double* dSSE,
                                const double* d_coef, const double* d_sigma, double* d_total, cudaStream_t stream) {
  const size_t smem = ((size_t)d.M * d.M + d.n + (size_t)LLS_THREADS * ((d.n | 1) + ((K * K) | 1)) + LLS_THREADS) * 8;
  cudaError_t e = cudaFuncSetAttribute(loglik_total_small_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 16) per_sm = 16;
  const long long nblk = (d.n_draws + LLS_THREADS - 1) / LLS_THREADS;
  int grid = (int)((nblk < (long long)sms * per_sm) ? nblk : (long long)sms * per_sm);
  double* partial = nullptr;
  e = cudaMallocAsync((void**)&partial, (size_t)grid * sizeof(double), stream);
  if (e != cudaSuccess) return e;
  loglik_total_small_kernel<K><<<grid, LLS_THREADS, smem, stream>>>(d, dXX, dAols, dSSE, d_coef, d_sigma, partial);
  loglik_reduce_kernel<<<1, 32, 0, stream>>>(partial, grid, 1, d_total);
  e = cudaGetLastError();
  cudaFreeAsync(partial, stream);
  return e;
}

cudaError_t launch_loglik_total(int k, int M, int T, long long n_draws, const double* dXX, const double* dAols,
                                const double* dSSE, const double* d_coef, const double* d_sigma, double* d_total,
                                cudaStream_t stream) {
  if (k < 1 || k > 32 || M < 1 || T < 1 || n_draws < 1) return cudaErrorInvalidValue;
  LoglikTotalDims d;
  d.k = k; d.M = M; d.T = T; d.n = k * M; d.n_draws = n_draws;
  if (k <= 4 && d.n <= 64) {
    switch (k) {
      case 1: return launch_small<1>(d, dXX, dAols, dSSE, d_coef, d_sigma, d_total, stream);
      case 2: return launch_small<2>(d, dXX, dAols, dSSE, d_coef, d_sigma, d_total, stream);
      case 3: return launch_small<3>(d, dXX, dAols, dSSE, d_coef, d_sigma, d_total, stream);
      default: return launch_small<4>(d, dXX, dAols, dSSE, d_coef, d_sigma, d_total, stream);
    }
  }
  const size_t shared = (size_t)M * M + d.n + (size_t)k * k;
  const size_t per_warp = 2 * (size_t)d.n + (size_t)k * k + (size_t)(k * (k + 1)) / 2;
  int warps = LL_WARPS;
  while (warps > 1 && (shared + warps * per_warp) * 8 > 200 * 1024) warps >>= 1;
  const size_t smem = (shared + warps * per_warp) * 8;
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(loglik_total_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
  e = cudaMallocAsync((void**)&partial, (size_t)n_warps * sizeof(double), stream);
  if (e != cudaSuccess) return e;
  cudaMemsetAsync(partial, 0, (size_t)n_warps * sizeof(double), stream);
  loglik_total_kernel<<<grid, warps * 32, smem, stream>>>(d, dXX, dAols, dSSE, d_coef, d_sigma, partial);
  loglik_reduce_kernel<<<1, 32, 0, stream>>>(partial, n_warps, 1, d_total);
  e = cudaGetLastError();
  cudaFreeAsync(partial, stream);
  return e;
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
