// bvg_kern_large.cu -- large-VAR path of the constant-parameter sweep (config c5: K = 20, p = 4, n = 1620): the n x n
// posterior precision of bvaralg.cpp:305 does not fit on chip, so every chain owns an n_pad x n_pad FP64 matrix in HBM
// (n_pad = n rounded up to 64; the padding is an identity block, which leaves the factor of the leading n x n block
// unchanged) and the sweep is a sequence of kernels batched over all chains:
//
//   large_build   P = diag(V^-1) + (X X') (x) Sigma^-1 written tile by tile (never the dense kron(I_T, Sigma^-1) of
//                 bvaralg.cpp:219,514), rhs b = V^-1 mu + vec(Sigma^-1 Y X')                       [HBM-bound, 8 B / entry]
//   blocked left-looking Cholesky, 64-wide block columns J = 0 .. n_pad/64 - 1:                      (bvaralg.cpp:307)
//     large_gemm    A[I,J] -= sum_{K<J} L[I,K] L[J,K]'  for every row tile I >= J: FP64 tensor-core contraction
//                   (mma.sync m8n8k4 f64 = DMMA) on 64 x 64 CTA tiles, operands staged global -> shared by TMA bulk
//                   copies (cp.async.bulk + mbarrier, 3-stage ring)                                  [FP64-bound, n^3/3 flops]
//     large_potrf   L[J,J] = chol(A[J,J]) in shared memory, and its inverse (kept per block column)
//     large_trsm    L[I,J] = A[I,J] L[J,J]^-T  for I > J: one more DMMA tile product
//   large_solve   w = L^-1 b, w += eps (Philox), a = L^-T w  -- the draw a = P^-1 b + chol(P)^-1 eps of
//                 bvaralg.cpp:306-307 --, streaming L once per direction                             [HBM-bound, 8 n^2 B]
//   large_post    residuals u = y - A x, S = S0 + u u', inverse-Wishart block, store                 (bvaralg.cpp:364,511-528)
#include <cuda_runtime.h>
#include <stdint.h>
#include "bvg_kernels.cuh"
#include "bvg_tiled.cuh"

namespace bvg {

#ifndef BVG_LARGE_KC
#define BVG_LARGE_KC 8       // swept on B200 (c5): (16,3) 479 ms, (32,2) 542, (32,3) 692, (8,3) 473, (8,4) 472, (8,5) 469, (4,6) 525
#endif
#ifndef BVG_LARGE_NSTAGE
#define BVG_LARGE_NSTAGE 4
#endif
static constexpr int TB = 64;        // tile / block-column width
static constexpr int KC = BVG_LARGE_KC;        // k-chunk of the GEMM pipeline
static constexpr int LDS_ = 68;      // shared-memory column stride (doubles): conflict-free DMMA fragment loads
static constexpr int NSTAGE = BVG_LARGE_NSTAGE;

// ---- PTX helpers: mbarrier + TMA 1-D bulk copy + FP64 MMA ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy through the TMA unit (UBLKCP in SASS); size and addresses multiples of 16 B
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// D (8x8) += A (8x4, row) * B (4x8, col); thread t holds A[t/4][t%4], B[t%4][t/4], C[t/4][2(t%4) + {0,1}]
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct LargeDims {
  int k, T, M, n, npad, nt;          // nt = npad / TB block columns
  long long mat_stride;              // doubles per chain matrix (npad * npad)
  long long linv_stride;             // doubles per chain: nt * TB * TB
};

// ---- P and b ---------------------------------------------------------------------------------------------------
// Only block column 0 of P is materialised here (+ the right-hand side): every other tile is generated inside the
// epilogue of the trailing-update kernel that first touches it (large_gemm_kernel), which saves one write and one
// read of the whole matrix per sweep.
// grid: (nt row tiles, chains); block 256.  sig = state (k x k col-major, Sigma^-1).
__global__ void __launch_bounds__(256) large_build_kernel(const __grid_constant__ BvarModel m, const __grid_constant__ LargeDims d,
                                                          const double* __restrict__ state, int state_len,
                                                          double* __restrict__ mats, double* __restrict__ rhs) {
  __shared__ double sig[1024];                       // k <= 32 on this path (checked by the host)
  const long long chain = blockIdx.y;
  const int I = blockIdx.x, J = 0;
  const int k = d.k, kk = k * k, n = d.n, npad = d.npad;
  const double* st = state + (size_t)chain * state_len;
  for (int e = threadIdx.x; e < kk; e += blockDim.x) sig[e] = st[e];
  __syncthreads();
  double* A = mats + (size_t)chain * d.mat_stride;
  for (int e = threadIdx.x; e < TB * TB; e += blockDim.x) {
    const int rr = e % TB, cc = e / TB;
    const int r = I * TB + rr, c = J * TB + cc;
    double v = 0.0;
    if (r < n && c < n) {
      if (c <= r) {
        const int j = r / k, i = r % k, j2 = c / k, i2 = c % k;
        v = m.XX[j + (size_t)d.M * j2] * sig[i + k * i2];
        if (r == c) v += st[kk + r];                 // vdiag lives behind Sigma^-1 in the chain state
      }
    } else if (r == c) {
      v = 1.0;                                       // identity padding
    }
    A[(size_t)c * npad + r] = v;
  }
  {                                                  // rhs rows of this row tile
    double* b = rhs + (size_t)chain * npad;
    for (int rr = threadIdx.x; rr < TB; rr += blockDim.x) {
      const int r = I * TB + rr;
      double s = 0.0;
      if (r < n) {
        const int j = r / k, i = r % k;
        for (int i2 = 0; i2 < k; ++i2) s = fma(sig[i + k * i2], m.YX[i2 + k * j], s);
        s += st[kk + r] * m.mu[r];
      }
      b[r] = s;
    }
  }
}

// ---- A[I,J] -= sum_{K<J} L[I,K] L[J,K]'   (I >= J) ---------------------------------------------------------------
// grid: (nt - J row tiles, chains); block 128 = 4 warps, warp (wr, wc) owns the 32 x 32 quadrant of the 64 x 64 tile.
__global__ void __launch_bounds__(128, 5) large_gemm_kernel(const __grid_constant__ LargeDims d, int J, double* __restrict__ mats,
                                                         const double* __restrict__ XX, const double* __restrict__ state,
                                                         int state_len) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double sig[1024];                             // Sigma^-1 of the chain (k <= 32): P is generated in the epilogue
  double* sA = (double*)smraw;                             // [NSTAGE][KC][LDS_]
  double* sB = sA + NSTAGE * KC * LDS_;
  uint64_t* bars = (uint64_t*)(sB + NSTAGE * KC * LDS_);
  const long long chain = blockIdx.y;
  const int I = J + blockIdx.x, npad = d.npad;
  double* A = mats + (size_t)chain * d.mat_stride;
  const int kdim = J * TB, nchunk = kdim / KC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wr = warp >> 1, wc = warp & 1;
  const bool diag = (I == J);
  const double* st = state + (size_t)chain * state_len;
  for (int e = tid; e < d.k * d.k; e += blockDim.x) sig[e] = st[e];
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // warp 0 issues the 2 * KC bulk copies (512 B each) of a chunk, one per lane and pass: a single elected thread
  // looping over them kept its warp -- and, through the next CTA barrier, the whole CTA -- waiting
  auto issue = [&](int chunk) {
    const int s = chunk % NSTAGE;
    const uint32_t bytes = (uint32_t)(TB * sizeof(double));
    if (lane == 0) mbar_expect_tx(&bars[s], bytes * KC * (diag ? 1 : 2));
    __syncwarp();
    for (int cc = lane; cc < 2 * KC; cc += 32) {
      const int c = cc % KC, isB = cc / KC;
      const size_t col = (size_t)(chunk * KC + c) * npad;
      if (!isB) tma_bulk_g2s(sA + (s * KC + c) * LDS_, A + col + (size_t)I * TB, bytes, &bars[s]);
      else if (!diag) tma_bulk_g2s(sB + (s * KC + c) * LDS_, A + col + (size_t)J * TB, bytes, &bars[s]);
    }
  };
  if (warp == 0) for (int c = 0; c < NSTAGE - 1 && c < nchunk; ++c) issue(c);
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  const int fr = lane >> 2, fk = lane & 3;
  for (int chunk = 0; chunk < nchunk; ++chunk) {
    const int s = chunk % NSTAGE;
    mbar_wait(&bars[s], (uint32_t)((chunk / NSTAGE) & 1));
    const double* tA = sA + s * KC * LDS_;
    const double* tB = diag ? tA : sB + s * KC * LDS_;
#pragma unroll
    for (int k4 = 0; k4 < KC; k4 += 4) {
      double fa[4], fb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) fa[a] = tA[(k4 + fk) * LDS_ + wr * 32 + a * 8 + fr];
#pragma unroll
      for (int b = 0; b < 4; ++b) fb[b] = tB[(k4 + fk) * LDS_ + wc * 32 + b * 8 + fr];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
    }
    __syncthreads();                                        // everyone is done with stage s (and the older ones)
    if (warp == 0 && chunk + NSTAGE - 1 < nchunk) issue(chunk + NSTAGE - 1);
  }
  // epilogue: A[I,J] = P[I,J] - acc with P = XX' (x) Sigma^-1 + diag(V^-1) generated on the fly (identity padding rows);
  // the strict upper triangle of a diagonal tile is never read
  {
    const int k = d.k, kk = k * k, n = d.n, M = d.M;
    int jr[4], ir[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = I * TB + wr * 32 + a * 8 + fr;
      jr[a] = r / k; ir[a] = r - jr[a] * k;
    }
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = J * TB + wc * 32 + b * 8 + 2 * fk + h;
        const int jc = c / k, ic = c - jc * k;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int r = I * TB + wr * 32 + a * 8 + fr;
          double v = 0.0;
          if (r < n && c < n) {
            if (c <= r) {
              v = XX[jr[a] + (size_t)M * jc] * sig[ir[a] + k * ic];
              if (r == c) v += st[kk + r];
            }
          } else if (r == c) {
            v = 1.0;
          }
          A[(size_t)c * npad + r] = v - acc[a][b][h];
        }
      }
  }
}

// ---- L[J,J] = chol(A[J,J]) and its inverse --------------------------------------------------------------------------
// grid: chains; block 256.  The 64 x 64 diagonal block is factorised as 8 x 8 tiles in shared memory by the CTA-wide
// DMMA Cholesky of bvg_tiled.cuh (diagonal tiles factorised + inverted in registers, panels and updates as tile
// products), then inverted tile row by tile row:  M_II = T_I,  M_IJ = -T_I sum_{J <= q < I} L_Iq M_qJ  (one tile per warp).
// Only the inverse is written out (linv): the trailing kernels and the solve never read the diagonal block itself.
__global__ void __launch_bounds__(256) large_potrf_kernel(const __grid_constant__ LargeDims d, int J, double* __restrict__ mats,
                                                          double* __restrict__ linv, int* __restrict__ status) {
  extern __shared__ double potrf_sm[];
  constexpr int NTT = TB / 8;                                  // tile rows of the block
  double* Lt = potrf_sm;                                       // tiled_len(TB) doubles, tile (I, Jt) at (I (I + 1) / 2 + Jt) * 64
  double* rhs = Lt + tiled_len(TB);                            // TB zeros: tiled_chol carries a right-hand side along
  const long long chain = blockIdx.x;
  const int npad = d.npad, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int li = lane & 7, lj = lane >> 3, fr = lane >> 2, fk = lane & 3;
  const int fo0 = fk * 8 + fr, fo1 = fo0 + 32, co = (2 * fk) * 8 + fr, to0 = fr * 8 + fk, to1 = to0 + 4;
  double* A = mats + (size_t)chain * d.mat_stride + (size_t)J * TB * npad + (size_t)J * TB;
  // load the lower tiles: a warp fetches half a tile per pass (8 consecutive rows of 4 columns)
  for (int t = warp; t < (NTT * (NTT + 1)) / 2; t += nwarp) {
    int I = 0;
    while (((I + 1) * (I + 2)) / 2 <= t) ++I;
    const int Jt = t - (I * (I + 1)) / 2;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = 8 * I + li, c = 8 * Jt + lj + 4 * h;
      Lt[t * 64 + lane + 32 * h] = (c <= r) ? A[(size_t)c * npad + r] : 0.0;
    }
  }
  for (int e = tid; e < TB; e += blockDim.x) rhs[e] = 0.0;
  __syncthreads();
  int fail = 0;
  tiled_chol(Lt, rhs, TB, fail);
  if (fail && tid == 0) status[chain] = 1;                      // warp 0 sees every pivot
#define BVG_LT(I, Jt) (Lt + ((((I) * ((I) + 1)) >> 1) + (Jt)) * 64)
  for (int I = 1; I < NTT; ++I) {
    double x0 = 0.0, x1 = 0.0;
    const int Jt = warp;                                        // one tile of row I per warp (I <= 7 < nwarp = 8)
    if (Jt < I) {
      double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
      for (int q = Jt; q < I; ++q) {
        const double* LA = BVG_LT(I, q);
        const double* MB = BVG_LT(q, Jt);
        tiled_dmma(c0, c1, LA[fo0], MB[to0]);
        tiled_dmma(e0, e1, LA[fo1], MB[to1]);
      }
      x0 = c0 + e0; x1 = c1 + e1;
    }
    __syncthreads();                                            // every warp has read row I of L
    if (Jt < I) {
      double* X = BVG_LT(I, Jt);
      X[co] = x0; X[co + 8] = x1;
      __syncwarp();
      const double* TI = BVG_LT(I, I);
      double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
      tiled_dmma(c0, c1, -TI[fo0], X[to0]);
      tiled_dmma(e0, e1, -TI[fo1], X[to1]);
      X[co] = c0 + e0; X[co + 8] = c1 + e1;
    }
    __syncthreads();
  }
  // Minv, col-major TB x TB with an exactly zero upper triangle
  double* Lo = linv + (size_t)chain * d.linv_stride + (size_t)J * TB * TB;
  for (int t = warp; t < NTT * NTT; t += nwarp) {
    const int I = t % NTT, Jt = t / NTT;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = 8 * I + li, c = 8 * Jt + lj + 4 * h;
      Lo[(size_t)c * TB + r] = (Jt <= I) ? BVG_LT(I, Jt)[lane + 32 * h] : 0.0;
    }
  }
#undef BVG_LT
}

// ---- L[I,J] = A[I,J] L[J,J]^-T = A[I,J] Minv'   (I > J) -------------------------------------------------------------
// grid: (nt - J - 1, chains); block 128.  out[r][c] = sum_k A[r][k] Minv[c][k]  (Minv lower: k <= c).
__global__ void __launch_bounds__(128) large_trsm_kernel(const __grid_constant__ LargeDims d, int J, double* __restrict__ mats,
                                                         const double* __restrict__ linv) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* sA = (double*)smraw;                // [TB cols][LDS_]  A tile, col-major (k = column)
  double* sM = sA + TB * LDS_;                // [TB k][LDS_]     Minv, col-major: sM[k * LDS_ + c] = Minv[c][k]
  uint64_t* bar = (uint64_t*)(sM + TB * LDS_);
  const long long chain = blockIdx.y;
  const int I = J + 1 + blockIdx.x, npad = d.npad;
  double* A = mats + (size_t)chain * d.mat_stride + (size_t)J * TB * npad + (size_t)I * TB;
  const double* Mg = linv + (size_t)chain * d.linv_stride + (size_t)J * TB * TB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wr = warp >> 1, wc = warp & 1;
  {
    const uint32_t bytes = (uint32_t)(TB * sizeof(double));
    if (tid == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect_tx(bar, bytes * 2 * TB);
    }
    __syncthreads();
    for (int cc = tid; cc < 2 * TB; cc += blockDim.x) {      // 128 bulk copies, one per thread
      const int c = cc % TB;
      if (cc < TB) tma_bulk_g2s(sA + c * LDS_, A + (size_t)c * npad, bytes, bar);
      else tma_bulk_g2s(sM + c * LDS_, Mg + (size_t)c * TB, bytes, bar);
    }
  }
  mbar_wait(bar, 0);
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  const int fr = lane >> 2, fk = lane & 3;
  for (int k4 = 0; k4 < TB; k4 += 4) {
    double fa[4], fb[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) fa[a] = sA[(k4 + fk) * LDS_ + wr * 32 + a * 8 + fr];
#pragma unroll
    for (int b = 0; b < 4; ++b) fb[b] = sM[(k4 + fk) * LDS_ + wc * 32 + b * 8 + fr];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int r = wr * 32 + a * 8 + fr, c = wc * 32 + b * 8 + 2 * fk;
      A[(size_t)c * npad + r] = acc[a][b][0];
      A[(size_t)(c + 1) * npad + r] = acc[a][b][1];
    }
}

// ---- the draw: w = L^-1 b, w += eps, a = L^-T w ------------------------------------------------------------------------
// grid: chains; block 512.  Block forward / backward substitution with the stored diagonal-block inverses; L is
// streamed from HBM once per direction with coalesced column segments.
__global__ void __launch_bounds__(512) large_solve_kernel(const __grid_constant__ LargeDims d, const __grid_constant__ RngParams rp,
                                                          int sweep, const double* __restrict__ mats,
                                                          const double* __restrict__ linv, const double* __restrict__ rhs,
                                                          double* __restrict__ avec) {
  extern __shared__ double sm[];
  double* w = sm;                   // npad
  double* part = w + d.npad;        // 8 x TB partial sums
  double* tmp = part + 8 * TB;      // TB
  double* part2 = tmp + TB;         // TB x (TB + 1)
  const long long chain = blockIdx.x;
  const int npad = d.npad, nt = d.nt, n = d.n, tid = threadIdx.x;
  const double* A = mats + (size_t)chain * d.mat_stride;
  const double* Lv = linv + (size_t)chain * d.linv_stride;
  const ChainRng rng(&rp, chain);
  for (int r = tid; r < npad; r += blockDim.x) w[r] = rhs[(size_t)chain * npad + r];
  __syncthreads();
  const int rr = tid % TB, seg = tid / TB;           // 8 segments of columns
  // forward: w_J = Minv_J (b_J - sum_{K<J} L[J,K] w_K)
  for (int J = 0; J < nt; ++J) {
    double s = 0.0;
    for (int c = seg; c < J * TB; c += 8) s = fma(A[(size_t)c * npad + J * TB + rr], w[c], s);
    part[seg * TB + rr] = s;
    __syncthreads();
    if (tid < TB) {
      double v = w[J * TB + tid];
      for (int q = 0; q < 8; ++q) v -= part[q * TB + tid];
      tmp[tid] = v;
    }
    __syncthreads();
    if (tid < TB) {
      double v = 0.0;
      const double* Mj = Lv + (size_t)J * TB * TB;
      for (int c = 0; c <= tid; ++c) v = fma(Mj[tid + c * TB], tmp[c], v);
      w[J * TB + tid] = v;
    }
    __syncthreads();
  }
  // eps (bvaralg.cpp:307): N(0,1) per coefficient id; the padding rows stay untouched (they are zero)
  for (int r = tid; r < n; r += blockDim.x) w[r] += rng.normal(VB_COEF_N, sweep, r);
  __syncthreads();
  // backward: a_J = Minv_J' (w_J - sum_{K>J} L[K,J]' a_K).  Lanes run along the rows of a column (contiguous in
  // memory), 8 column groups; the 64 partial sums of every column are reduced through shared memory.
  const int rl = tid % TB, cg = tid / TB;
  for (int J = nt - 1; J >= 0; --J) {
    for (int cc = cg; cc < TB; cc += 8) {
      double s = 0.0;
      const double* col = A + (size_t)(J * TB + cc) * npad;
      for (int r = (J + 1) * TB + rl; r < npad; r += TB) s = fma(col[r], w[r], s);
      part2[cc * (TB + 1) + rl] = s;
    }
    __syncthreads();
    if (tid < TB) {
      double v = w[J * TB + tid];
      const double* pp = part2 + tid * (TB + 1);
      for (int q = 0; q < TB; ++q) v -= pp[q];
      tmp[tid] = v;
    }
    __syncthreads();
    if (tid < TB) {
      double v = 0.0;
      const double* Mj = Lv + (size_t)J * TB * TB;
      for (int r = tid; r < TB; ++r) v = fma(Mj[r + tid * TB], tmp[r], v);       // Minv'[tid][r] = Minv[r][tid]
      w[J * TB + tid] = v;
    }
    __syncthreads();
  }
  for (int r = tid; r < n; r += blockDim.x) avec[(size_t)chain * npad + r] = w[r];
}

// ---- residuals, Wishart block, store ------------------------------------------------------------------------------------
// grid: chains; block 256.  Shared memory: U (k*T) + 6 k*k + 2 k + a (n).
__global__ void __launch_bounds__(256) large_post_kernel(const __grid_constant__ BvarModel m, const __grid_constant__ LargeDims d,
                                                         const __grid_constant__ RngParams rp, int sweep, int burnin,
                                                         const double* __restrict__ avec, double* __restrict__ state,
                                                         int state_len, const __grid_constant__ BvarOut out) {
  extern __shared__ double sm[];
  BlockGroup g(sm);
  const long long chain = blockIdx.x;
  const int k = d.k, T = d.T, M = d.M, n = d.n, kk = k * k, tid = threadIdx.x, G = blockDim.x;
  double* p = sm + BVG_BLOCK_SCRATCH;
  double* a = p;        p += n;
  double* U = p;        p += k * T;
  double* S = p;        p += kk;
  double* sig = p;      p += kk;
  double* W1 = p;       p += kk;
  double* W2 = p;       p += kk;
  double* W3 = p;       p += kk;
  double* W4 = p;       p += kk;
  double* kd = p;       p += k;
  const ChainRng rng(&rp, chain);
  int fail = 0;
  for (int r = tid; r < n; r += G) a[r] = avec[(size_t)chain * d.npad + r];
  __syncthreads();
  for (int e = tid; e < k * T; e += G) {                                            // bvaralg.cpp:364
    const int i = e % k, t = e / k;
    double s = m.Y[e];
    for (int j = 0; j < M; ++j) s = fma(-a[j * k + i], m.X[j + M * t], s);
    U[e] = s;
  }
  __syncthreads();
  for (int e = tid; e < kk; e += G) {
    const int i = e % k, j = e / k;
    double s = m.S0[e];
    for (int t = 0; t < T; ++t) s = fma(U[i + k * t], U[j + k * t], s);
    S[e] = s;
  }
  __syncthreads();
  const bool keep = sweep >= burnin;
  wishart_block(g, rng, sweep, k, m.wish_df_post, S, sig, W4, keep, W1, W2, W3, W4, kd, fail);   // bvaralg.cpp:511
  double* st = state + (size_t)chain * state_len;
  for (int e = tid; e < kk; e += G) st[e] = sig[e];
  if (keep) {
    const long long pos = (long long)sweep - burnin;
    double* oc = out.coef + ((size_t)chain * (size_t)out.iters + (size_t)pos) * (size_t)n;
    for (int r = tid; r < n; r += G) oc[r] = a[r];
    double* os = out.sigma + ((size_t)chain * (size_t)out.iters + (size_t)pos) * (size_t)kk;
    for (int e = tid; e < kk; e += G) os[e] = W4[e];
  }
  if (fail && tid == 0) out.status[chain] = 1;
}

// ---- host orchestration ---------------------------------------------------------------------------------------------------
size_t large_workspace_doubles(const BvarModel& m, long long n_chains) {
  const int npad = ((m.n + TB - 1) / TB) * TB, nt = npad / TB;
  return (size_t)n_chains * ((size_t)npad * npad + (size_t)nt * TB * TB + 2 * (size_t)npad);
}

cudaError_t launch_bvar_large(const BvarModel& m, const RngParams& rp, double* state, int state_len, double* work,
                              const BvarOut& out, long long n_chains, int s0, int s1, int burnin, cudaStream_t stream,
                              long long* launches) {
  LargeDims d;
  d.k = m.k; d.T = m.T; d.M = m.M; d.n = m.n;
  d.npad = ((m.n + TB - 1) / TB) * TB;
  d.nt = d.npad / TB;
  d.mat_stride = (long long)d.npad * d.npad;
  d.linv_stride = (long long)d.nt * TB * TB;
  double* mats = work;
  double* linv = mats + (size_t)n_chains * d.mat_stride;
  double* rhs = linv + (size_t)n_chains * d.linv_stride;
  double* avec = rhs + (size_t)n_chains * d.npad;
  const size_t gemm_smem = (size_t)2 * NSTAGE * KC * LDS_ * 8 + NSTAGE * 8;
  const size_t trsm_smem = (size_t)2 * TB * LDS_ * 8 + 8;
  const size_t potrf_smem = ((size_t)tiled_len(TB) + TB + 8) * 8;
  const size_t solve_smem = ((size_t)d.npad + 9 * TB + TB * (TB + 1)) * 8;
  const size_t post_smem = ((size_t)BVG_BLOCK_SCRATCH + d.n + (size_t)d.k * d.T + 6 * d.k * d.k + 2 * d.k) * 8;
  cudaError_t e;
  if ((e = set_smem(large_gemm_kernel, gemm_smem)) != cudaSuccess) return e;
  if ((e = set_smem(large_trsm_kernel, trsm_smem)) != cudaSuccess) return e;
  if ((e = set_smem(large_potrf_kernel, potrf_smem)) != cudaSuccess) return e;
  if ((e = set_smem(large_solve_kernel, solve_smem)) != cudaSuccess) return e;
  if ((e = set_smem(large_post_kernel, post_smem)) != cudaSuccess) return e;
  const unsigned nc = (unsigned)n_chains;
  for (int sweep = s0; sweep < s1; ++sweep) {
    large_build_kernel<<<dim3(d.nt, nc), 256, 0, stream>>>(m, d, state, state_len, mats, rhs);
    *launches += 1;
    for (int J = 0; J < d.nt; ++J) {
      if (J > 0) {
        large_gemm_kernel<<<dim3(d.nt - J, nc), 128, gemm_smem, stream>>>(d, J, mats, m.XX, state, state_len);
        *launches += 1;
      }
      large_potrf_kernel<<<nc, 256, potrf_smem, stream>>>(d, J, mats, linv, out.status);
      *launches += 1;
      if (J + 1 < d.nt) {
        large_trsm_kernel<<<dim3(d.nt - J - 1, nc), 128, trsm_smem, stream>>>(d, J, mats, linv);
        *launches += 1;
      }
    }
    large_solve_kernel<<<nc, 512, solve_smem, stream>>>(d, rp, sweep, mats, linv, rhs, avec);
    large_post_kernel<<<nc, 256, post_smem, stream>>>(m, d, rp, sweep, burnin, avec, state, state_len, out);
    *launches += 2;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace bvg
