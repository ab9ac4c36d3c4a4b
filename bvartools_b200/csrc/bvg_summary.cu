// bvg_summary.cu -- posterior summaries of a draw buffer that stays in HBM: what summary.bvar takes
// from coda::summary.mcmc for every coefficient (R/summary.bvar.R:121-140) -- mean, SD, naive SE and the quantiles
// (R's default type 7) -- pooled over all draws of all chains, without shipping the draws to the host.
//
// Quantiles are EXACT order statistics found by an 8-pass MSB radix select on order-preserving 64-bit keys:
//   pass p: every row keeps one 256-bin histogram per requested probability, restricted to the elements whose top p
//           bytes equal the prefix selected so far (shared-memory histograms, warp-aggregated atomics, merged into a
//           global table); a one-CTA kernel then picks the bin that holds the wanted rank and extends the prefix.
//   after 8 passes the prefix IS the key of x_(lo), lo = floor((N - 1) p); one more pass finds its upper neighbour
//           (the smallest element above it, or itself when it is duplicated), and Q = x_lo + (h - lo)(x_hi - x_lo).
// Every pass streams the buffer once (HBM-bound: 8 bytes per element and pass); the multi-GPU version all-reduces the
// histogram table between the two kernels of a pass.
#include "bvg_kernels.cuh"

namespace bvg {

static constexpr int SUM_THREADS = 256;

__device__ __forceinline__ unsigned long long dkey(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

struct SelDims {
  long long n_draws, rows;
  int n_probs, pass;           // pass 0..7
  int row0, row_cnt;           // row chunk of this launch (shared-memory histograms: row_cnt * n_probs * 256 counters)
};

// prefix[row][q]: key bytes selected so far (top `pass` bytes valid); hist[row][q][256] global counters
template <int NQ>   // NQ >= histograms per row of this pass (1 in pass 0, n_probs afterwards): the target loop unrolls exactly
__global__ void __launch_bounds__(SUM_THREADS) select_hist_kernel(const __grid_constant__ SelDims d,
                                                                   const double* __restrict__ draws,
                                                                   const unsigned long long* __restrict__ prefix,
                                                                   unsigned int* __restrict__ hist) {
  extern __shared__ unsigned int sh[];                       // [row_cnt][nq][256]
  const int nq = (d.pass == 0) ? 1 : d.n_probs;              // pass 0: all probabilities share the empty prefix
  const int nbin = d.row_cnt * nq * 256;
  for (int e = threadIdx.x; e < nbin; e += blockDim.x) sh[e] = 0;
  __syncthreads();
  const int shift = 56 - 8 * d.pass;
  const unsigned long long himask = (d.pass == 0) ? 0ull : (~0ull << (64 - 8 * d.pass));
  // every thread stays on ONE row of the chunk (its targets' prefixes live in registers) and walks the draws with a
  // grid-wide stride; consecutive threads read consecutive rows of a draw (contiguous in memory)
  const long long nthr = (long long)gridDim.x * blockDim.x;
  const long long per_pass = nthr / d.row_cnt;               // draws covered by one grid-wide step
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (per_pass > 0 && gt < per_pass * d.row_cnt) {
    const int rl = (int)(gt % d.row_cnt);
    unsigned long long pf[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) pf[q] = (d.pass > 0 && q < nq) ? prefix[(size_t)(d.row0 + rl) * d.n_probs + q] : 0ull;
    unsigned int* myh = sh + (size_t)rl * nq * 256;
    for (long long dr = gt / d.row_cnt; dr < d.n_draws; dr += per_pass) {
      const unsigned long long key = dkey(draws[dr * d.rows + d.row0 + rl]);
      const unsigned int bin = (unsigned int)(key >> shift) & 255u;
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        if ((NQ == 1 || q < nq) && ((key ^ pf[q]) & himask) == 0ull) atomicAdd(&myh[q * 256 + bin], 1u);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nbin; e += blockDim.x) {
    const unsigned int v = sh[e];
    if (v) {
      const int rl = e / (nq * 256), rem = e % (nq * 256), q = rem / 256, b = rem % 256;
      if (d.pass == 0) { for (int qq = 0; qq < d.n_probs; ++qq) atomicAdd(&hist[((size_t)(d.row0 + rl) * d.n_probs + qq) * 256 + b], v); }
      else atomicAdd(&hist[((size_t)(d.row0 + rl) * d.n_probs + q) * 256 + b], v);
    }
  }
}

// one thread per (row, probability): find the bin holding the wanted rank, extend the prefix, clear the histogram
__global__ void select_pick_kernel(long long rows, int n_probs, int pass, unsigned long long* __restrict__ prefix,
                                   long long* __restrict__ rank, unsigned int* __restrict__ hist) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * n_probs) return;
  unsigned int* h = hist + (size_t)t * 256;
  long long r = rank[t], cum = 0;
  int b = 0;
  for (; b < 255; ++b) {
    if (cum + (long long)h[b] > r) break;
    cum += h[b];
  }
  rank[t] = r - cum;
  prefix[t] |= (unsigned long long)b << (56 - 8 * pass);
  for (int e = 0; e < 256; ++e) h[e] = 0;
}

// upper neighbour of x_lo: count of elements <= x_lo and the smallest key above it, per (row, probability)
template <int NQ>
__global__ void __launch_bounds__(SUM_THREADS) select_next_kernel(const __grid_constant__ SelDims d,
                                                                   const double* __restrict__ draws,
                                                                   const unsigned long long* __restrict__ prefix,
                                                                   unsigned long long* __restrict__ cnt_le,
                                                                   unsigned long long* __restrict__ min_gt) {
  // same mapping over all rows: a thread stays on one row, keeps count / minimum per probability in registers and
  // publishes them once
  const long long nthr = (long long)gridDim.x * blockDim.x;
  const long long per_pass = nthr / d.rows;
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (per_pass == 0) {                                        // more rows than threads: one row at a time per thread
    for (long long r = gt; r < d.rows; r += nthr)
      for (int q = 0; q < d.n_probs; ++q) {
        const size_t t = (size_t)r * d.n_probs + q;
        const unsigned long long lo = prefix[t];
        unsigned long long c = 0ull, mn = ~0ull;
        for (long long dr = 0; dr < d.n_draws; ++dr) {
          const unsigned long long key = dkey(draws[dr * d.rows + r]);
          if (key == lo) ++c; else if (key > lo && key < mn) mn = key;
        }
        cnt_le[t] = c; min_gt[t] = mn;
      }
    return;
  }
  if (gt >= per_pass * d.rows) return;
  const long long r = gt % d.rows;
  unsigned long long lo[NQ], c[NQ], mn[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) { lo[q] = (q < d.n_probs) ? prefix[(size_t)r * d.n_probs + q] : 0ull; c[q] = 0ull; mn[q] = ~0ull; }
  for (long long dr = gt / d.rows; dr < d.n_draws; dr += per_pass) {
    const unsigned long long key = dkey(draws[dr * d.rows + r]);
#pragma unroll
    for (int q = 0; q < NQ; ++q)
      if (q < d.n_probs) {
        if (key == lo[q]) ++c[q];
        else if (key > lo[q] && key < mn[q]) mn[q] = key;
      }
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q)
    if (q < d.n_probs) {
      const size_t t = (size_t)r * d.n_probs + q;
      if (c[q]) atomicAdd(&cnt_le[t], c[q]);
      if (mn[q] != ~0ull) atomicMin(&min_gt[t], mn[q]);
    }
}

__global__ void select_finish_kernel(long long rows, int n_probs, long long n_draws, const double* __restrict__ probs,
                                     const unsigned long long* __restrict__ prefix, const long long* __restrict__ rank_in_eq,
                                     const unsigned long long* __restrict__ cnt_eq, const unsigned long long* __restrict__ min_gt,
                                     double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * n_probs) return;
  const int q = (int)(t % n_probs);
  const long long r = t / n_probs;
  const double h = (double)(n_draws - 1) * probs[q];
  const double lo = floor(h);
  const double xlo = dunkey(prefix[t]);
  // rank_in_eq: 0-based position of x_(lo) among the elements equal to it; the next order statistic is the same
  // value when more duplicates follow, else the smallest element above it (the maximum has no successor: h == lo)
  double xhi = xlo;
  if (h > lo) {
    const bool dup_follows = (unsigned long long)(rank_in_eq[t] + 1) < cnt_eq[t];
    if (!dup_follows && min_gt[t] != ~0ull) xhi = dunkey(min_gt[t]);
  }
  out[(size_t)q * rows + r] = xlo + (h - lo) * (xhi - xlo);              // quantile type 7
}

__global__ void select_init_kernel(long long rows, int n_probs, long long n_draws, const double* __restrict__ probs,
                                   unsigned long long* __restrict__ prefix, long long* __restrict__ rank,
                                   unsigned long long* __restrict__ cnt_eq, unsigned long long* __restrict__ min_gt) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * n_probs) return;
  const double h = (double)(n_draws - 1) * probs[t % n_probs];
  prefix[t] = 0ull;
  rank[t] = (long long)floor(h);
  cnt_eq[t] = 0ull;
  min_gt[t] = ~0ull;
}

// second pass of the standard deviation: ss[r] = sum (x - mean_r)^2 (no E[x^2] - mean^2 cancellation)
__global__ void __launch_bounds__(256) centered_ss_kernel(const double* __restrict__ draws, long long n_draws, long long rows,
                                                          const double* __restrict__ sums, double* __restrict__ ss) {
  const long long nthr = (long long)gridDim.x * blockDim.x;
  const long long per_pass = nthr / rows;
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (per_pass == 0) {
    for (long long r = gt; r < rows; r += nthr) {
      const double mu = sums[r] / (double)n_draws;
      double acc = 0.0;
      for (long long dr = 0; dr < n_draws; ++dr) { const double v = draws[dr * rows + r] - mu; acc = fma(v, v, acc); }
      ss[r] = acc;
    }
    return;
  }
  if (gt >= per_pass * rows) return;
  const long long r = gt % rows;
  const double mu = sums[r] / (double)n_draws;
  double acc = 0.0;
  for (long long dr = gt / rows; dr < n_draws; dr += per_pass) { const double v = draws[dr * rows + r] - mu; acc = fma(v, v, acc); }
  atomicAdd(&ss[r], acc);
}

cudaError_t launch_centered_ss(const double* draws, long long n_draws, long long rows, const double* d_sums, double* d_ss,
                               cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(d_ss, 0, (size_t)rows * 8, stream);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  centered_ss_kernel<<<sms * 8, 256, 0, stream>>>(draws, n_draws, rows, d_sums, d_ss);
  return cudaGetLastError();
}

// quantiles[q][row] (device) of draws [n_draws][rows] (device); probs (device, n_probs)
cudaError_t launch_quantiles(const double* draws, long long n_draws, long long rows, const double* d_probs, int n_probs,
                             double* d_out, cudaStream_t stream) {
  if (n_draws < 1 || rows < 1 || n_probs < 1 || n_probs > 16) return cudaErrorInvalidValue;
  const size_t nt = (size_t)rows * n_probs;
  unsigned long long *prefix = nullptr, *cnt_eq = nullptr, *min_gt = nullptr;
  long long* rank = nullptr;
  unsigned int* hist = nullptr;
  cudaError_t e;
  if ((e = cudaMallocAsync((void**)&prefix, nt * 8, stream)) != cudaSuccess ||
      (e = cudaMallocAsync((void**)&rank, nt * 8, stream)) != cudaSuccess ||
      (e = cudaMallocAsync((void**)&cnt_eq, nt * 8, stream)) != cudaSuccess ||
      (e = cudaMallocAsync((void**)&min_gt, nt * 8, stream)) != cudaSuccess ||
      (e = cudaMallocAsync((void**)&hist, nt * 256 * 4, stream)) != cudaSuccess) {
    if (prefix) cudaFreeAsync(prefix, stream);
    if (rank) cudaFreeAsync(rank, stream);
    if (cnt_eq) cudaFreeAsync(cnt_eq, stream);
    if (min_gt) cudaFreeAsync(min_gt, stream);
    return e;
  }
  cudaMemsetAsync(hist, 0, nt * 256 * 4, stream);
  const int tb = 128, gb = (int)((nt + tb - 1) / tb);
  select_init_kernel<<<gb, tb, 0, stream>>>(rows, n_probs, n_draws, d_probs, prefix, rank, cnt_eq, min_gt);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // rows per launch so that the shared histograms fit (<= 96 KB: two CTAs per SM)
  int row_chunk = (int)((96 * 1024) / ((size_t)n_probs * 256 * 4));
  if (row_chunk < 1) row_chunk = 1;
  if (row_chunk > rows) row_chunk = (int)rows;
  // compile-time target counts: 1 (pass 0), then n_probs rounded up to 2 / 3 / 4 / 8 / 16
  auto hist_launch = [&](const SelDims& d, int grid_, size_t smem_) {
    const int nq = (d.pass == 0) ? 1 : n_probs;
#define BVG_HL(N) do { cudaFuncSetAttribute(select_hist_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); \
                       select_hist_kernel<N><<<grid_, SUM_THREADS, smem_, stream>>>(d, draws, prefix, hist); } while (0)
    if (nq == 1) BVG_HL(1); else if (nq == 2) BVG_HL(2); else if (nq == 3) BVG_HL(3); else if (nq == 4) BVG_HL(4);
    else if (nq <= 8) BVG_HL(8); else BVG_HL(16);
#undef BVG_HL
  };
  int grid = sms * 2;
  for (int pass = 0; pass < 8; ++pass) {
    for (int r0 = 0; r0 < rows; r0 += row_chunk) {
      SelDims d;
      d.n_draws = n_draws; d.rows = rows; d.n_probs = n_probs; d.pass = pass;
      d.row0 = r0; d.row_cnt = (int)((r0 + row_chunk <= rows) ? row_chunk : rows - r0);
      const size_t smem = (size_t)d.row_cnt * ((pass == 0) ? 1 : n_probs) * 256 * 4;
      hist_launch(d, grid, smem);
    }
    select_pick_kernel<<<gb, tb, 0, stream>>>(rows, n_probs, pass, prefix, rank, hist);
  }
  // after the last pass rank[t] = position of x_(lo) among its duplicates; count them and find the next larger key
  {
    SelDims d;
    d.n_draws = n_draws; d.rows = rows; d.n_probs = n_probs; d.pass = 8; d.row0 = 0; d.row_cnt = (int)rows;
#define BVG_NL(N) select_next_kernel<N><<<sms * 8, SUM_THREADS, 0, stream>>>(d, draws, prefix, cnt_eq, min_gt)
    if (n_probs == 1) BVG_NL(1); else if (n_probs == 2) BVG_NL(2); else if (n_probs == 3) BVG_NL(3);
    else if (n_probs == 4) BVG_NL(4); else if (n_probs <= 8) BVG_NL(8); else BVG_NL(16);
#undef BVG_NL
  }
  select_finish_kernel<<<gb, tb, 0, stream>>>(rows, n_probs, n_draws, d_probs, prefix, rank, cnt_eq, min_gt, d_out);
  e = cudaGetLastError();
  cudaFreeAsync(prefix, stream); cudaFreeAsync(rank, stream); cudaFreeAsync(cnt_eq, stream);
  cudaFreeAsync(min_gt, stream); cudaFreeAsync(hist, stream);
  return e;
}

// ---------------------------------------------------------------------------------------------------------------------
// Time-series SE of coda::summary.mcmc (the "Time-series SE" column summary.bvar copies: R/summary.bvar.R:133,141,195,
// 203).  coda is an un-vendored CRAN dependency; its spectrum0.ar fits an AR(m) by Yule-Walker (stats::ar.yw, AIC order
// selection, order.max = min(n - 1, floor(10 log10 n))) to every series and returns var.pred / (1 - sum ar)^2; the SE is
// sqrt(spec / n), and for several chains (summary.mcmc.list) sqrt(mean_c spec_c / (n chains)).
// draws: [chain][iter][row].  Three passes: per-chain means (+ the slope sum of the lm(x ~ t) degeneracy check),
// autocovariances for lags 0..L from shared-memory tiles of 32 rows, and one thread per series for Levinson-Durbin.
static constexpr int TS_ROWS = 32;        // rows per CTA tile (one 256-byte line per period)
static constexpr int TS_SEG = 128;        // periods per shared-memory segment
static constexpr int TS_LG = 8;           // lags per work item
static constexpr int TS_MAXLAG = 63;

// sums[chain][row] = sum_t x, tsum[chain][row] = sum_t t x  (t = 0..iters-1)
__global__ void __launch_bounds__(256) tsse_mean_kernel(const double* __restrict__ draws, long long iters, long long rows,
                                                        double* __restrict__ sums, double* __restrict__ tsum) {
  __shared__ double red[2][8][TS_ROWS];
  const long long chain = blockIdx.x;
  const int r = threadIdx.x % TS_ROWS, sl = threadIdx.x / TS_ROWS;
  const long long row = (long long)blockIdx.y * TS_ROWS + r;
  const double* base = draws + (size_t)chain * (size_t)iters * (size_t)rows;
  double s = 0.0, st = 0.0;
  if (row < rows)
    for (long long t = sl; t < iters; t += 8) {
      const double v = base[(size_t)t * (size_t)rows + row];
      s += v;
      st = fma((double)t, v, st);
    }
  red[0][sl][r] = s; red[1][sl][r] = st;
  __syncthreads();
  if (sl == 0 && row < rows) {
    double a = 0.0, b = 0.0;
    for (int j = 0; j < 8; ++j) { a += red[0][j][r]; b += red[1][j][r]; }
    sums[(size_t)chain * rows + row] = a;
    tsum[(size_t)chain * rows + row] = b;
  }
}

// acov[chain][row][l] = (1/n) sum_{t < n - l} (x_t - m)(x_{t+l} - m),  l = 0..L
__global__ void __launch_bounds__(256) tsse_acov_kernel(const double* __restrict__ draws, long long iters, long long rows,
                                                        int L, const double* __restrict__ sums, double* __restrict__ acov) {
  extern __shared__ double ts_sm[];                      // [(TS_SEG + halo)][TS_ROWS], then the slice reduction buffer
  const int nlg = (L + TS_LG) / TS_LG;                   // lag groups of 8
  const int halo = nlg * TS_LG;
  const int items = TS_ROWS * nlg;                       // <= 256
  const int slices = (nlg <= 1) ? 8 : (nlg == 2) ? 4 : (nlg <= 4) ? 2 : 1;   // power of two, slices * items <= 256
  const long long chain = blockIdx.x;
  const long long row0 = (long long)blockIdx.y * TS_ROWS;
  const double* base = draws + (size_t)chain * (size_t)iters * (size_t)rows;
  const int item = threadIdx.x % items, sl = threadIdx.x / items;
  const int r = item % TS_ROWS, lg = item / TS_ROWS;
  const bool worker = sl < slices;
  const int lr = threadIdx.x % TS_ROWS;
  const double mean = (row0 + lr < rows) ? sums[(size_t)chain * rows + row0 + lr] / (double)iters : 0.0;
  double acc[TS_LG];
#pragma unroll
  for (int j = 0; j < TS_LG; ++j) acc[j] = 0.0;
  for (long long t0 = 0; t0 < iters; t0 += TS_SEG) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < (TS_SEG + halo) * TS_ROWS; idx += 256) {
      const long long t = t0 + idx / TS_ROWS;            // idx % TS_ROWS == lr (256 is a multiple of 32)
      double v = 0.0;
      if (t < iters && row0 + lr < rows) v = base[(size_t)t * (size_t)rows + row0 + lr] - mean;
      ts_sm[idx] = v;
    }
    __syncthreads();
    if (worker) {
      // a slice owns span consecutive periods (a multiple of 8): per block of 8 periods the 8 values x_t and the 15
      // lagged values x_{t+l0 .. t+l0+14} are read once into registers and feed 64 FMAs
      const int span = TS_SEG / slices;
      const double* xr = ts_sm + r;
      const int l0 = lg * TS_LG;
      for (int t = sl * span; t < (sl + 1) * span; t += 8) {
        double a[8], w[15];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = xr[(t + i) * TS_ROWS];
#pragma unroll
        for (int i = 0; i < 15; ++i) w[i] = xr[(t + l0 + i) * TS_ROWS];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < TS_LG; ++j) acc[j] = fma(a[i], w[i + j], acc[j]);
      }
    }
  }
  __syncthreads();
  double* red = ts_sm;                                   // [slices][items][TS_LG]
  if (worker)
#pragma unroll
    for (int j = 0; j < TS_LG; ++j) red[((size_t)sl * items + item) * TS_LG + j] = acc[j];
  __syncthreads();
  if (sl == 0 && row0 + r < rows) {
#pragma unroll
    for (int j = 0; j < TS_LG; ++j) {
      const int l = lg * TS_LG + j;
      if (l > L) break;
      double s = 0.0;
      for (int q = 0; q < slices; ++q) s += red[((size_t)q * items + item) * TS_LG + j];
      acov[((size_t)chain * rows + row0 + r) * (size_t)(L + 1) + l] = s / (double)iters;
    }
  }
}

// spec[chain][row] = var.pred / (1 - sum ar)^2 of the AIC-selected Yule-Walker fit (0 for a degenerate series)
__global__ void tsse_spec_kernel(long long n_series, long long iters, int L, const double* __restrict__ sums,
                                 const double* __restrict__ tsum, const double* __restrict__ acov,
                                 double* __restrict__ spec) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_series) return;
  const double n = (double)iters;
  const double* r = acov + (size_t)e * (size_t)(L + 1);
  // residual sd of lm(x ~ t): SSR = n c0 - Sxy^2 / Sxx
  const double tbar = 0.5 * (n - 1.0);
  const double sxy = tsum[e] - tbar * sums[e];
  const double sxx = n * (n * n - 1.0) / 12.0;
  double ssr = n * r[0] - ((sxx > 0.0) ? sxy * sxy / sxx : 0.0);
  if (ssr < 0.0) ssr = 0.0;
  if (iters < 2 || sqrt(ssr / (n - 1.0)) < 1.5e-8) { spec[e] = 0.0; return; }
  double phi[TS_MAXLAG + 1], prev[TS_MAXLAG + 1];
  double v = r[0];
  double best_aic = n * log(v) + 2.0, best_v = v, best_sum = 0.0;
  int best_m = 0;
  for (int m = 1; m <= L; ++m) {
    double num = r[m];
    for (int j = 1; j < m; ++j) num = fma(-prev[j], r[m - j], num);
    const double kappa = num / v;
    for (int j = 1; j < m; ++j) phi[j] = fma(-kappa, prev[m - j], prev[j]);
    phi[m] = kappa;
    v *= (1.0 - kappa * kappa);
    double sum = 0.0;
    for (int j = 1; j <= m; ++j) { prev[j] = phi[j]; sum += phi[j]; }
    const double aic = n * log(v) + 2.0 * m + 2.0;
    if (aic < best_aic) { best_aic = aic; best_v = v; best_sum = sum; best_m = m; }   // NaN never wins (nanargmin)
  }
  const double var_pred = best_v * n / (n - (double)(best_m + 1));
  const double d = 1.0 - best_sum;
  spec[e] = var_pred / (d * d);
}

// one CTA of 128 threads per row; fixed summation order (strided partial sums, then a tree): deterministic
__global__ void __launch_bounds__(128) tsse_combine_kernel(long long n_chains, long long iters, long long rows,
                                                           const double* __restrict__ spec, double* __restrict__ out) {
  __shared__ double part[128];
  const long long row = blockIdx.x;
  double s = 0.0;
  for (long long c = threadIdx.x; c < n_chains; c += 128) s += spec[(size_t)c * rows + row];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int w = 64; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) part[threadIdx.x] += part[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[row] = sqrt(part[0] / (double)n_chains / ((double)iters * (double)n_chains));
}

// d_out[rows] (device) <- time-series SE of draws [n_chains][iters][rows] (device)
cudaError_t launch_tsse(const double* draws, long long n_chains, long long iters, long long rows, double* d_out,
                        cudaStream_t stream) {
  if (n_chains < 1 || iters < 1 || rows < 1) return cudaErrorInvalidValue;
  int L = (int)floor(10.0 * log10((double)iters));
  if ((long long)L > iters - 1) L = (int)(iters - 1);
  if (L > TS_MAXLAG) return cudaErrorInvalidValue;
  const long long tiles = (rows + TS_ROWS - 1) / TS_ROWS;
  if (n_chains > 0x7fffffffLL || tiles > 65535) return cudaErrorInvalidValue;
  const size_t ns = (size_t)n_chains * (size_t)rows;
  double *sums = nullptr, *tsum = nullptr, *acov = nullptr, *spec = nullptr;
  auto release = [&]() {
    if (sums) cudaFreeAsync(sums, stream);
    if (tsum) cudaFreeAsync(tsum, stream);
    if (spec) cudaFreeAsync(spec, stream);
    if (acov) cudaFreeAsync(acov, stream);
  };
  cudaError_t e;
  if ((e = cudaMallocAsync((void**)&sums, ns * 8, stream)) != cudaSuccess ||
      (e = cudaMallocAsync((void**)&tsum, ns * 8, stream)) != cudaSuccess ||
      (e = cudaMallocAsync((void**)&spec, ns * 8, stream)) != cudaSuccess ||
      (e = cudaMallocAsync((void**)&acov, ns * (size_t)(L + 1) * 8, stream)) != cudaSuccess) {
    release();
    return e;
  }
  const dim3 grid((unsigned)n_chains, (unsigned)tiles);
  tsse_mean_kernel<<<grid, 256, 0, stream>>>(draws, iters, rows, sums, tsum);
  const int nlg = (L + TS_LG) / TS_LG;
  size_t smem = (size_t)(TS_SEG + nlg * TS_LG) * TS_ROWS * 8;
  const size_t red = (size_t)256 * TS_LG * 8;
  if (red > smem) smem = red;
  cudaFuncSetAttribute(tsse_acov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);   // per device: cheap
  tsse_acov_kernel<<<grid, 256, smem, stream>>>(draws, iters, rows, L, sums, acov);
  tsse_spec_kernel<<<(unsigned)((ns + 127) / 128), 128, 0, stream>>>((long long)ns, iters, L, sums, tsum, acov, spec);
  tsse_combine_kernel<<<(unsigned)rows, 128, 0, stream>>>(n_chains, iters, rows, spec, d_out);
  e = cudaGetLastError();
  release();
  return e;
}

}  // namespace bvg
