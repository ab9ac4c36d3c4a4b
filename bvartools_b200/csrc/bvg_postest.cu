// bvg_postest.cu -- per-draw post-estimation on draws resident in HBM: impulse responses.
//
// irf.bvar (R/irf.bvar.R:78-216) loops over every posterior draw in R and calls .ir (src/ir.cpp:4-72), which builds the
// forecast-error responses Phi_i = sum_{j=1..min(i,p)} Phi_{i-j} A_j (k x k each) and returns
// theta_i = (Phi_i P)[response, impulse] for the transformation P of the requested type; the bands are quantiles over
// the draws.  Only row `response` of Phi_i is ever used, so one THREAD per draw propagates that row vector
// (phi_i = sum_j phi_{i-j} A_j: k^2 p products per horizon instead of k^3 p) against column `impulse` of P:
//   feir  P = I shock                      oir  P = L diag(L)^-1 shock, L = chol(Sigma) lower      (ir.cpp:28-34)
//   gir   P = Sigma / Sigma_jj shock                                                                 (ir.cpp:38-41)
// shock = number, or "sd" / "nsd": +- L_jj for oir, +- sqrt(Sigma_jj) otherwise (R/irf.bvar.R:176-188).
// Output [draw][h + 1] (cumulated on request, R/irf.bvar.R:204-206) stays on the device; the bands come from the radix
// select of bvg_summary.cu.  Structural types (ir.cpp:35-37,42-46): sir P = A0^-1 shock, sgir P = A0^-1 Sigma / Sigma_jj
// shock -- column `impulse` of P is one k x k solve per draw (Gaussian elimination with partial pivoting on a full A0,
// forward substitution when A0 arrives as the sampler's compact strict lower triangle).
#include "bvg_kernels.cuh"
#include "bvg_blocks.cuh"

namespace bvg {

struct IrfDims {
  long long n_draws, coef_stride, sigma_stride, a0_stride;
  int k, p, h, type, impulse, response, shock_kind, cumulative;
  int a0_mode;        // 0 none, 1 full k x k column-major, 2 compact strict lower triangle (by column), unit diagonal
  double shock;
};
static constexpr int IRF_A0_KMAX = 16;

static constexpr int IRF_KMAX = 32;

__global__ void __launch_bounds__(128) irf_draws_kernel(const __grid_constant__ IrfDims d, const double* __restrict__ coef,
                                                        const double* __restrict__ sigma, const double* __restrict__ a0,
                                                        double* __restrict__ out) {
  extern __shared__ double sm[];                       // per thread: ring of p row vectors (k each) + column of P (k)
  const int k = d.k, p = d.p, h = d.h;
  const int per_thread = (p + 1) * k;
  double* ring = sm + (size_t)threadIdx.x * per_thread;     // phi_{i-1..i-p}, slot (i mod p)
  double* pc = ring + p * k;                                 // P[:, impulse]
  for (long long dr = (long long)blockIdx.x * blockDim.x + threadIdx.x; dr < d.n_draws;
       dr += (long long)gridDim.x * blockDim.x) {
    const double* A = coef + dr * d.coef_stride;             // k x (k p), column-major: A_j[a][b] at a + k (b + k (j - 1))
    const double* S = sigma + dr * d.sigma_stride;           // k x k
    const int jm = d.impulse, rs = d.response;
    // ---- column `impulse` of P and the size of the shock
    double shock = d.shock;
    if (d.type == 1) {
      // L = chol(Sigma) lower, column jm: only rows >= jm are non-zero; needs the leading jm + 1 columns of L
      double Lc[IRF_KMAX * (IRF_KMAX + 1) / 2];
      for (int c = 0; c <= jm; ++c)
        for (int i = c; i < k; ++i) {
          double s = S[i + k * c];
          for (int m = 0; m < c; ++m) s = fma(-Lc[(i * (i + 1)) / 2 + m], Lc[(c * (c + 1)) / 2 + m], s);
          Lc[(i * (i + 1)) / 2 + c] = (i == c) ? sqrt(s) : s / Lc[(c * (c + 1)) / 2 + c];
        }
      const double ljj = Lc[(jm * (jm + 1)) / 2 + jm];
      if (d.shock_kind != 0) shock = (d.shock_kind == 2) ? -ljj : ljj;
      for (int i = 0; i < k; ++i) pc[i] = (i >= jm) ? Lc[(i * (i + 1)) / 2 + jm] / ljj * shock : 0.0;
    } else {
      if (d.shock_kind != 0) { const double sdv = sqrt(S[jm + k * jm]); shock = (d.shock_kind == 2) ? -sdv : sdv; }
      if (d.type == 2 || d.type == 4) { const double sjj = S[jm + k * jm]; for (int i = 0; i < k; ++i) pc[i] = S[i + k * jm] / sjj * shock; }
      else for (int i = 0; i < k; ++i) pc[i] = (i == jm) ? shock : 0.0;
      if (d.type >= 3) {
        // pc <- A0^-1 pc  (sir: A0^-1 e_j shock;  sgir: A0^-1 Sigma[:, j] / Sigma_jj shock)
        const double* A0 = a0 + dr * d.a0_stride;
        if (d.a0_mode == 2) {
          // unit lower, element (i, c), i > c, at c (k - 1) - c (c - 1) / 2 + (i - c - 1)
          for (int i = 1; i < k; ++i) {
            double sv = pc[i];
            for (int c = 0; c < i; ++c) sv = fma(-A0[c * (k - 1) - (c * (c - 1)) / 2 + (i - c - 1)], pc[c], sv);
            pc[i] = sv;
          }
        } else {
          double W[IRF_A0_KMAX * IRF_A0_KMAX];
          for (int e = 0; e < k * k; ++e) W[e] = A0[e];
          for (int c = 0; c < k; ++c) {
            int pv = c;
            for (int i = c + 1; i < k; ++i) if (fabs(W[i + k * c]) > fabs(W[pv + k * c])) pv = i;
            if (pv != c) {
              for (int e = 0; e < k; ++e) { const double tw = W[c + k * e]; W[c + k * e] = W[pv + k * e]; W[pv + k * e] = tw; }
              const double tb = pc[c]; pc[c] = pc[pv]; pc[pv] = tb;
            }
            const double piv = W[c + k * c];
            for (int i = c + 1; i < k; ++i) {
              const double f = W[i + k * c] / piv;
              for (int e = c + 1; e < k; ++e) W[i + k * e] = fma(-f, W[c + k * e], W[i + k * e]);
              pc[i] = fma(-f, pc[c], pc[i]);
            }
          }
          for (int i = k - 1; i >= 0; --i) {
            double sv = pc[i];
            for (int e = i + 1; e < k; ++e) sv = fma(-W[i + k * e], pc[e], sv);
            pc[i] = sv / W[i + k * i];
          }
        }
      }
    }
    // ---- phi_0 = e_response'
    double* o = out + dr * (h + 1);
    double acc = 0.0;
    {
      const double th = pc[rs];
      acc = th;
      o[0] = th;
    }
    if (p > 0) for (int b = 0; b < k; ++b) ring[b] = (b == rs) ? 1.0 : 0.0;     // slot 0 = phi_0
    for (int i = 1; i <= h; ++i) {
      double* cur = (p > 0) ? ring + (i % p) * k : nullptr;
      double th = 0.0;
      if (p > 0) {
        // phi_i[b] = sum_{j=1..min(i,p)} sum_a phi_{i-j}[a] A_j[a][b]; slot (i mod p) still holds phi_{i-p}: read it last
        double nw[IRF_KMAX];
        for (int b = 0; b < k; ++b) nw[b] = 0.0;
        const int jmax = (i < p) ? i : p;
        for (int j = 1; j <= jmax; ++j) {
          const double* ph = ring + ((i - j) % p) * k;
          const double* Aj = A + (size_t)k * k * (j - 1);
          for (int b = 0; b < k; ++b) {
            double s = nw[b];
            for (int a = 0; a < k; ++a) s = fma(ph[a], Aj[a + k * b], s);
            nw[b] = s;
          }
        }
        for (int b = 0; b < k; ++b) { cur[b] = nw[b]; th = fma(nw[b], pc[b], th); }
      }
      acc += th;
      o[i] = d.cumulative ? acc : th;
    }
  }
}

cudaError_t launch_irf_draws(const double* coef, long long coef_stride, const double* sigma, long long sigma_stride,
                             long long n_draws, int k, int p, int h, int type, int impulse, int response, int shock_kind,
                             double shock, int cumulative, double* d_out, cudaStream_t stream, const double* a0,
                             long long a0_stride, int a0_mode) {
  if (k < 1 || k > IRF_KMAX || p < 0 || h < 0 || type < 0 || type > 4 || impulse < 0 || impulse >= k || response < 0 ||
      response >= k || n_draws < 1)
    return cudaErrorInvalidValue;
  if (type >= 3 && (a0 == nullptr || a0_mode < 1 || a0_mode > 2 || (a0_mode == 1 && k > IRF_A0_KMAX))) return cudaErrorInvalidValue;
  IrfDims d;
  d.n_draws = n_draws; d.coef_stride = coef_stride; d.sigma_stride = sigma_stride;
  d.a0_stride = a0_stride; d.a0_mode = (type >= 3) ? a0_mode : 0;
  d.k = k; d.p = p; d.h = h; d.type = type; d.impulse = impulse; d.response = response; d.shock_kind = shock_kind;
  d.cumulative = cumulative; d.shock = shock;
  int threads = 128;
  size_t smem = (size_t)threads * (p + 1) * k * 8;
  while (smem > 96 * 1024 && threads > 32) { threads >>= 1; smem = (size_t)threads * (p + 1) * k * 8; }
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(irf_draws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long want = (n_draws + threads - 1) / threads;
  int grid = (int)((want < (long long)sms * 8) ? want : (long long)sms * 8);
  irf_draws_kernel<<<grid, threads, smem, stream>>>(d, coef, sigma, a0, d_out);
  return cudaGetLastError();
}

// ---- forecast error variance decomposition (fevd.bvar, R/fevd.bvar.R:72-186; .vardecomp, src/vardecomp.cpp:4-87) ------
// Per draw, with phi_i = row `response` of Phi_i:  numerator_i[c] = sum_{l<=i} (phi_l P)[c]^2,
// mse_i = sum_{l<=i} phi_l Sigma phi_l',  result_i[c] = numerator_i[c] / sigmajj / mse_i  (sigmajj = sqrt(Sigma_rr) for gir,
// 1 for oir -- exactly as the reference normalises).  fevd.bvar returns the MEAN over the draws of the (h + 1) x k table:
// a thread per draw, warp-shuffle sums, per-warp slots in shared memory, per-CTA partials reduced in a fixed order.
// Structural types (vardecomp.cpp:32-36,41-45): sir P = A0^-1, Sigma_mse = A0^-1 A0^-T; sgir P = A0^-1 Sigma, Sigma_mse =
// A0^-1 Sigma A0^-T -- with w = phi A0^-1 these are the oir-free / gir formulas applied to w instead of phi.
struct FevdDims {
  long long n_draws, coef_stride, sigma_stride, a0_stride;
  int k, p, h, type, response;                    // type 1 oir, 2 gir, 3 sir, 4 sgir
  int a0_mode;                                    // 1 full k x k, 2 compact strict lower triangle by column (unit diagonal)
};

__global__ void __launch_bounds__(128) fevd_draws_kernel(const __grid_constant__ FevdDims d, const double* __restrict__ coef,
                                                         const double* __restrict__ sigma, const double* __restrict__ a0,
                                                         double* __restrict__ partial) {
  extern __shared__ double sm[];
  const int k = d.k, p = d.p, h = d.h, nout = (h + 1) * k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  double* wacc = sm;                                         // [nwarp][nout] per-warp sums
  double* ring0 = wacc + (size_t)nwarp * nout;
  const int per_thread = p * k + 2 * k;
  double* ring = ring0 + (size_t)threadIdx.x * per_thread;   // phi ring (p x k) | numerator (k) | v = phi P (k)
  double* num = ring + p * k;
  double* v = num + k;
  for (int e = threadIdx.x; e < nwarp * nout; e += blockDim.x) wacc[e] = 0.0;
  __syncthreads();
  const long long nthr = (long long)gridDim.x * blockDim.x;
  const long long rounds = (d.n_draws + nthr - 1) / nthr;    // every thread of a warp runs the same number of rounds
  for (long long rd = 0; rd < rounds; ++rd) {
    const long long dr = rd * nthr + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = dr < d.n_draws;
    const double* A = coef + (live ? dr : 0) * d.coef_stride;
    const double* S = sigma + (live ? dr : 0) * d.sigma_stride;
    double Lc[IRF_KMAX * (IRF_KMAX + 1) / 2];
    double sigmajj = 1.0;
    if (d.type == 1) {                                       // P = chol(Sigma)' = L (lower)
      for (int c = 0; c < k; ++c)
        for (int i = c; i < k; ++i) {
          double s = S[i + k * c];
          for (int m = 0; m < c; ++m) s = fma(-Lc[(i * (i + 1)) / 2 + m], Lc[(c * (c + 1)) / 2 + m], s);
          Lc[(i * (i + 1)) / 2 + c] = (i == c) ? sqrt(s) : s / Lc[(c * (c + 1)) / 2 + c];
        }
    } else if (d.type != 3) {
      sigmajj = sqrt(S[d.response + k * d.response]);
    }
    const double* A0 = (d.type >= 3) ? a0 + (live ? dr : 0) * d.a0_stride : nullptr;
    double Ainv[IRF_A0_KMAX * IRF_A0_KMAX];
    if (d.type >= 3 && d.a0_mode == 1) {
      // full A0: Gauss-Jordan inverse with partial pivoting, once per draw
      double W[IRF_A0_KMAX * IRF_A0_KMAX];
      for (int e = 0; e < k * k; ++e) { W[e] = A0[e]; Ainv[e] = (e % k == e / k) ? 1.0 : 0.0; }
      for (int c = 0; c < k; ++c) {
        int pv = c;
        for (int i = c + 1; i < k; ++i) if (fabs(W[i + k * c]) > fabs(W[pv + k * c])) pv = i;
        if (pv != c)
          for (int e = 0; e < k; ++e) {
            double tw = W[c + k * e]; W[c + k * e] = W[pv + k * e]; W[pv + k * e] = tw;
            tw = Ainv[c + k * e]; Ainv[c + k * e] = Ainv[pv + k * e]; Ainv[pv + k * e] = tw;
          }
        const double piv = 1.0 / W[c + k * c];
        for (int e = 0; e < k; ++e) { W[c + k * e] *= piv; Ainv[c + k * e] *= piv; }
        for (int i = 0; i < k; ++i) {
          if (i == c) continue;
          const double f = W[i + k * c];
          for (int e = 0; e < k; ++e) { W[i + k * e] = fma(-f, W[c + k * e], W[i + k * e]); Ainv[i + k * e] = fma(-f, Ainv[c + k * e], Ainv[i + k * e]); }
        }
      }
    }
    for (int b = 0; b < k; ++b) num[b] = 0.0;
    double mse = 0.0;
    for (int i = 0; i <= h; ++i) {
      // phi_i into a local row
      double ph[IRF_KMAX];
      if (i == 0) { for (int b = 0; b < k; ++b) ph[b] = (b == d.response) ? 1.0 : 0.0; }
      else {
        for (int b = 0; b < k; ++b) ph[b] = 0.0;
        const int jmax = (i < p) ? i : p;
        for (int j = 1; j <= jmax; ++j) {
          const double* pv = ring + ((i - j) % p) * k;
          const double* Aj = A + (size_t)k * k * (j - 1);
          for (int b = 0; b < k; ++b) {
            double s = ph[b];
            for (int a = 0; a < k; ++a) s = fma(pv[a], Aj[a + k * b], s);
            ph[b] = s;
          }
        }
      }
      if (p > 0) for (int b = 0; b < k; ++b) ring[(i % p) * k + b] = ph[b];
      if (d.type >= 3) {
        // ph <- w = phi A0^-1 (the ring keeps phi itself)
        if (d.a0_mode == 2) {
          // w A0 = phi with A0 unit lower: w_c = phi_c - sum_{i > c} A0[i][c] w_i, from the last column to the first
          for (int c = k - 2; c >= 0; --c) {
            double sv = ph[c];
            for (int i2 = c + 1; i2 < k; ++i2) sv = fma(-A0[c * (k - 1) - (c * (c - 1)) / 2 + (i2 - c - 1)], ph[i2], sv);
            ph[c] = sv;
          }
        } else {
          double pw[IRF_A0_KMAX];
          for (int c = 0; c < k; ++c) {
            double sv = 0.0;
            for (int a = 0; a < k; ++a) sv = fma(ph[a], Ainv[a + k * c], sv);
            pw[c] = sv;
          }
          for (int c = 0; c < k; ++c) ph[c] = pw[c];
        }
      }
      // v = phi P ;  mse += phi Sigma phi'
      for (int c = 0; c < k; ++c) {
        double s = 0.0;
        if (d.type == 1) { for (int a = c; a < k; ++a) s = fma(ph[a], Lc[(a * (a + 1)) / 2 + c], s); }
        else if (d.type == 3) s = ph[c];
        else { for (int a = 0; a < k; ++a) s = fma(ph[a], S[a + k * c], s); }
        v[c] = s;
      }
      double q = 0.0;
      for (int c = 0; c < k; ++c) {
        double s = 0.0;
        if (d.type == 3) s = ph[c];
        else for (int a = 0; a < k; ++a) s = fma(ph[a], S[a + k * c], s);
        q = fma(s, ph[c], q);
      }
      mse += q;
      for (int c = 0; c < k; ++c) {
        num[c] = fma(v[c], v[c], num[c]);
        double r = live ? num[c] / sigmajj / mse : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (lane == 0) wacc[(size_t)warp * nout + i * k + c] += r;
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nout; e += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nwarp; ++w) s += wacc[(size_t)w * nout + e];
    partial[(size_t)blockIdx.x * nout + e] = s;
  }
}

__global__ void fevd_reduce_kernel(const double* __restrict__ partial, int n_part, int nout, double inv_n, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nout) return;
  double s = 0.0;
  for (int b = 0; b < n_part; ++b) s += partial[(size_t)b * nout + e];
  out[e] = s * inv_n;
}

// d_out: (h + 1) x k row-major = result[i][c], the mean over the draws
cudaError_t launch_fevd_draws(const double* coef, long long coef_stride, const double* sigma, long long sigma_stride,
                              long long n_draws, int k, int p, int h, int type, int response, double* d_out,
                              cudaStream_t stream, const double* a0, long long a0_stride, int a0_mode) {
  if (k < 1 || k > IRF_KMAX || p < 0 || h < 0 || type < 1 || type > 4 || response < 0 || response >= k || n_draws < 1)
    return cudaErrorInvalidValue;
  if (type >= 3 && (a0 == nullptr || a0_mode < 1 || a0_mode > 2 || (a0_mode == 1 && k > IRF_A0_KMAX))) return cudaErrorInvalidValue;
  FevdDims d;
  d.n_draws = n_draws; d.coef_stride = coef_stride; d.sigma_stride = sigma_stride;
  d.a0_stride = a0_stride; d.a0_mode = (type >= 3) ? a0_mode : 0;
  d.k = k; d.p = p; d.h = h; d.type = type; d.response = response;
  const int nout = (h + 1) * k;
  int threads = 128;
  auto bytes = [&](int t) { return ((size_t)(t / 32) * nout + (size_t)t * (p * k + 2 * k)) * 8; };
  while (bytes(threads) > 96 * 1024 && threads > 32) threads >>= 1;
  if (bytes(threads) > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(fevd_draws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes(threads));
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long want = (n_draws + threads - 1) / threads;
  int grid = (int)((want < (long long)sms * 4) ? want : (long long)sms * 4);
  double* partial = nullptr;
  if ((e = cudaMallocAsync((void**)&partial, (size_t)grid * nout * 8, stream)) != cudaSuccess) return e;
  fevd_draws_kernel<<<grid, threads, bytes(threads), stream>>>(d, coef, sigma, a0, partial);
  fevd_reduce_kernel<<<(nout + 127) / 128, 128, 0, stream>>>(partial, grid, nout, 1.0 / (double)n_draws, d_out);
  e = cudaGetLastError();
  cudaFreeAsync(partial, stream);
  return e;
}

// ---- forecasts (predict.bvar, R/predict.bvar.R:36-260; .draw_forecast, src/draw_forecast.cpp:6-51) ------------------
// Per draw: n_ahead simulated periods  y_{T+j+1} = A pred_j + u_j,  u_j = Sigma^{1/2} eps_j  with the SYMMETRIC square root
// (eig_sym in the reference, :27-28; cyclic Jacobi here -- the root of a positive definite matrix is unique), the lag
// block of pred shifted every period (:41-45).  pred is the reference's prediction matrix (rows: y lags | exogenous |
// deterministic; column 0 = data at T, later columns = the future exogenous / deterministic values), shared by all draws.
// One thread per draw, its matrices in a private slice of shared memory.  eps: injected [draw][n_ahead][k] or Philox.
struct FcstDims {
  long long n_draws, coef_stride, sigma_stride, a0_stride;
  int k, p, tot, h, use_a;
  int a0_mode;                                    // 0 none; 1 full k x k; 2 compact strict lower triangle (unit diagonal)
  unsigned long long seed;
};

// x <- A0^-1 x: forward substitution on the compact unit-lower rows, Gaussian elimination with partial pivoting otherwise
__device__ void a0_solve(const double* A0, int mode, int k, double* x) {
  if (mode == 2) {
    for (int i = 1; i < k; ++i) {
      double sv = x[i];
      for (int c = 0; c < i; ++c) sv = fma(-A0[c * (k - 1) - (c * (c - 1)) / 2 + (i - c - 1)], x[c], sv);
      x[i] = sv;
    }
    return;
  }
  double W[IRF_A0_KMAX * IRF_A0_KMAX];
  for (int e = 0; e < k * k; ++e) W[e] = A0[e];
  for (int c = 0; c < k; ++c) {
    int pv = c;
    for (int i = c + 1; i < k; ++i) if (fabs(W[i + k * c]) > fabs(W[pv + k * c])) pv = i;
    if (pv != c) {
      for (int e = 0; e < k; ++e) { const double tw = W[c + k * e]; W[c + k * e] = W[pv + k * e]; W[pv + k * e] = tw; }
      const double tb = x[c]; x[c] = x[pv]; x[pv] = tb;
    }
    const double piv = W[c + k * c];
    for (int i = c + 1; i < k; ++i) {
      const double f = W[i + k * c] / piv;
      for (int e = c + 1; e < k; ++e) W[i + k * e] = fma(-f, W[c + k * e], W[i + k * e]);
      x[i] = fma(-f, x[c], x[i]);
    }
  }
  for (int i = k - 1; i >= 0; --i) {
    double sv = x[i];
    for (int e = i + 1; e < k; ++e) sv = fma(-W[i + k * e], x[e], sv);
    x[i] = sv / W[i + k * i];
  }
}
static constexpr int VB_FORECAST = 40;            // variate block id of the forecast errors (site = draw, period, variable)

__global__ void __launch_bounds__(128) forecast_draws_kernel(const __grid_constant__ FcstDims d, const double* __restrict__ coef,
                                                             const double* __restrict__ sigma, const double* __restrict__ pred0,
                                                             const double* __restrict__ eps, const double* __restrict__ a0,
                                                             double* __restrict__ out) {
  extern __shared__ double sm[];
  const int k = d.k, p = d.p, tot = d.tot, h = d.h, kk = k * k;
  const int per_thread = 3 * kk + tot + k;
  double* Ae = sm + (size_t)threadIdx.x * per_thread;       // Sigma -> eigenvalues on the diagonal
  double* V = Ae + kk;
  double* Rt = V + kk;                                       // Sigma^{1/2}
  double* col = Rt + kk;                                     // current column of pred
  double* u = col + tot;
  const SerialGroup g;
  const int na = d.use_a ? ((p == 0) ? tot - k : tot) : 0;   // columns of A
  for (long long dr = (long long)blockIdx.x * blockDim.x + threadIdx.x; dr < d.n_draws;
       dr += (long long)gridDim.x * blockDim.x) {
    const double* A = coef + dr * d.coef_stride;             // k x na column-major
    const double* S = sigma + dr * d.sigma_stride;
    for (int e = 0; e < kk; ++e) Ae[e] = S[e];
    jacobi_eig_sym(g, Ae, V, k);
    sym_func(g, Ae, V, Rt, k, 0);
    for (int r = 0; r < tot; ++r) col[r] = pred0[r];         // column 0
    double* o = out + dr * (size_t)h * k;
    for (int j = 0; j < h; ++j) {
      for (int i = 0; i < k; ++i) {
        double s = 0.0;
        for (int c = 0; c < k; ++c) {
          const double e = (eps != nullptr) ? eps[(dr * h + j) * k + c]
                                            : philox_normal(d.seed, (unsigned long long)dr, VB_FORECAST, j, c, 0);
          s = fma(Rt[i + k * c], e, s);
        }
        u[i] = s;
      }
      // y_next = A0^-1 (A x + u)   (x = whole column, or its rows k.. when p = 0); without coefficients a random walk
      // y + A0^-1 u (draw_forecast.cpp:33-41; A0 = I for non-structural models)
      for (int i = 0; i < k; ++i) {
        double s = u[i];
        if (d.use_a) { for (int c = 0; c < na; ++c) s = fma(A[i + k * c], col[(p == 0 ? k : 0) + c], s); }
        u[i] = s;
      }
      if (d.a0_mode != 0) a0_solve(a0 + dr * d.a0_stride, d.a0_mode, k, u);
      for (int i = 0; i < k; ++i) o[(size_t)j * k + i] = d.use_a ? u[i] : u[i] + col[i];
      // next column: new y on top, lags shifted down, exogenous / deterministic rows from pred
      for (int l = p - 1; l >= 1; --l)
        for (int i = 0; i < k; ++i) col[l * k + i] = col[(l - 1) * k + i];
      for (int i = 0; i < k; ++i) col[i] = o[(size_t)j * k + i];
      const int ylen = (p > 0) ? k * p : k;
      for (int r = ylen; r < tot; ++r) col[r] = pred0[(size_t)(j + 1) * tot + r];
    }
  }
}

cudaError_t launch_forecast_draws(const double* coef, long long coef_stride, const double* sigma, long long sigma_stride,
                                  long long n_draws, int k, int p, int tot, int h, int use_a, const double* d_pred,
                                  const double* d_eps, unsigned long long seed, double* d_out, cudaStream_t stream,
                                  const double* a0, long long a0_stride, int a0_mode) {
  if (k < 1 || k > IRF_KMAX || p < 0 || h < 1 || tot < k || n_draws < 1 || ((p > 0) && tot < k * p)) return cudaErrorInvalidValue;
  if (a0 != nullptr && (a0_mode < 1 || a0_mode > 2 || (a0_mode == 1 && k > IRF_A0_KMAX))) return cudaErrorInvalidValue;
  FcstDims d;
  d.n_draws = n_draws; d.coef_stride = coef_stride; d.sigma_stride = sigma_stride;
  d.a0_stride = a0_stride; d.a0_mode = (a0 != nullptr) ? a0_mode : 0;
  d.k = k; d.p = p; d.tot = tot; d.h = h; d.use_a = use_a; d.seed = seed;
  int threads = 128;
  auto bytes = [&](int t) { return (size_t)t * (3 * k * k + tot + k) * 8; };
  while (bytes(threads) > 96 * 1024 && threads > 32) threads >>= 1;
  if (bytes(threads) > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(forecast_draws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes(threads));
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long want = (n_draws + threads - 1) / threads;
  int grid = (int)((want < (long long)sms * 8) ? want : (long long)sms * 8);
  forecast_draws_kernel<<<grid, threads, bytes(threads), stream>>>(d, coef, sigma, d_pred, d_eps, a0, d_out);
  return cudaGetLastError();
}

}  // namespace bvg
