// bvg_tiled.cuh -- CTA-wide FP64 Cholesky / triangular solves on an 8 x 8 tile-packed lower triangle in shared memory,
// trailing updates on the FP64 tensor pipe (mma.sync m8n8k4 f64 = DMMA).  Used by the CTA-per-chain variant of the
// constant-parameter sweep (config c3: n = 150) in place of the row-cyclic Crout factorisation, whose n^2/2 dependent
// dot products per sweep left the SM idle.  Device only.
//
// Layout: the matrix is padded to nt = ceil(n / 8) tile rows plus ONE extra tile row that carries the right-hand side
// b' in its first row (augmented system: after the factorisation that row holds w' = (L^-1 b)').  Tile (I, J), J <= I,
// lives at ((I (I + 1) / 2 + J) * 64, column-major inside the tile (element (i, j) at j * 8 + i): a warp reads a DMMA
// A/B fragment (32 doubles) from one contiguous, conflict-free half tile.  The padding rows are an identity block.
// Reference call sites: arma::chol + arma::solve in bvaralg.cpp:306-307.
#pragma once
#include "bvg_group.cuh"

namespace bvg {

BVG_HD int tiled_nt(int n) { return (n + 7) >> 3; }
BVG_HD int tiled_len(int n) { const int r = tiled_nt(n) + 1; return ((r * (r + 1)) >> 1) * 64; }
BVG_HD int tiled_idx(int i, int j) {      // element (i, j), j <= i (tile-wise: J <= I)
  const int I = i >> 3, J = j >> 3;
  return ((((I * (I + 1)) >> 1) + J) << 6) + ((j & 7) << 3) + (i & 7);
}

#if defined(__CUDACC__)
__device__ __forceinline__ void tiled_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// identity padding + rhs row; call after the n x n lower triangle has been written, before tiled_chol
__device__ __forceinline__ void tiled_finish_build(double* L, const double* b, int n) {
  const int nt = tiled_nt(n), npad = nt * 8, tid = threadIdx.x, G = blockDim.x;
  // padding rows n .. npad-1 (inside the last tile row) and the upper part of diagonal tiles are never read except
  // through whole-tile DMMA operands: zero the padding rows, unit diagonal
  for (int e = tid; e < (npad - n) * npad; e += G) {
    const int i = n + e / npad, j = e % npad;
    if (j <= i) L[tiled_idx(i, j)] = (i == j) ? 1.0 : 0.0;
  }
  // strict upper triangle inside the diagonal tiles (read by the panel solve of nothing, but keep it clean)
  for (int e = tid; e < nt * 64; e += G) {
    const int I = e >> 6, jj = (e >> 3) & 7, ii = e & 7;
    if (jj > ii) L[(((I * (I + 1)) >> 1) + I) * 64 + jj * 8 + ii] = 0.0;
  }
  // augmented tile row nt: row 0 = b', rows 1..7 = 0
  for (int e = tid; e < npad * 8; e += G) {
    const int j = e >> 3, ii = e & 7;
    L[tiled_idx(npad + ii, j)] = (ii == 0 && j < n) ? b[j] : 0.0;
  }
  __syncthreads();
}

// In-place Cholesky of the tile-packed matrix (true diagonal stored, dinv[j] = 1 / L_jj), rhs carried in the
// augmented row: on exit w[j] = (L^-1 b)_j.  fail is set on a non-positive pivot.
static __device__ __noinline__ void tiled_chol(double* L, double* dinv, double* w, int n, int& fail) {
  const int nt = tiled_nt(n), npad = nt * 8, tid = threadIdx.x, G = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = G >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  for (int p = 0; p < nt; ++p) {
    double* D = L + (((p * (p + 1)) >> 1) + p) * 64;             // diagonal tile
    // 1. factor the 8 x 8 diagonal tile (warp 0): lane i owns row i
    if (warp == 0) {
      for (int j = 0; j < 8; ++j) {
        const double d = D[j * 8 + j];
        __syncwarp();                                             // every lane has read the pivot before it is replaced
        if (!(d > 0.0)) fail = 1;
        const double rs = rsqrt(d);
        if (lane == j) { D[j * 8 + j] = d * rs; dinv[p * 8 + j] = rs; }
        else if (lane > j && lane < 8) D[j * 8 + lane] *= rs;
        __syncwarp();
        // trailing 8 x 8 update: 64 elements over 32 lanes
        for (int e = lane; e < 64; e += 32) {
          const int i = e & 7, c = e >> 3;
          if (c > j && i >= c) D[c * 8 + i] = fma(-D[j * 8 + i], D[j * 8 + c], D[c * 8 + i]);
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // 2. panel: rows below the diagonal tile (incl. the rhs row), L(I,p) = A(I,p) L(p,p)^-T by forward substitution
    const int rows_below = (npad + 8) - (p + 1) * 8;
    for (int rr = tid; rr < rows_below; rr += G) {
      const int i = (p + 1) * 8 + rr;
      double* R = L + tiled_idx(i, p * 8);                        // element (i, 8p); next column is +8
      double x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        double s = R[j * 8];
#pragma unroll
        for (int c = 0; c < j; ++c) s = fma(-x[c], D[c * 8 + j], s);
        x[j] = s * dinv[p * 8 + j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) R[j * 8] = x[j];
    }
    __syncthreads();
    // 3. trailing update C(I,J) -= L(I,p) L(J,p)' for p < J <= I <= nt (the rhs tile row takes part, its own
    //    diagonal tile (nt, nt) does not exist), one tile per warp and pass
    const int m = nt - p;                                         // tile rows p+1 .. nt  -> m rows
    const int ntile = ((m * (m + 1)) >> 1) - 1;                   // minus the (nt, nt) corner
    for (int t = warp; t < ntile; t += nwarp) {
      int a = 0;
      while (((a + 1) * (a + 2)) >> 1 <= t) ++a;
      const int bcol = t - ((a * (a + 1)) >> 1);
      const int I = p + 1 + a, J = p + 1 + bcol;
      const double* TA = L + (((I * (I + 1)) >> 1) + p) * 64;
      const double* TBt = L + (((J * (J + 1)) >> 1) + p) * 64;
      double* C = L + (((I * (I + 1)) >> 1) + J) * 64 + (2 * fk) * 8 + fr;
      double c0 = 0.0, c1 = 0.0;
      tiled_dmma(c0, c1, TA[fk * 8 + fr], TBt[fk * 8 + fr]);
      tiled_dmma(c0, c1, TA[(4 + fk) * 8 + fr], TBt[(4 + fk) * 8 + fr]);
      C[0] -= c0;
      C[8] -= c1;
    }
    __syncthreads();
  }
  for (int j = tid; j < n; j += G) w[j] = L[tiled_idx(npad, j)];
  __syncthreads();
}

// a = L^-T v (back substitution with the transposed factor); v (n, shared) is destroyed.
static __device__ __noinline__ void tiled_back_solve(const double* L, const double* dinv, double* v, double* a, int n) {
  const int nt = tiled_nt(n), tid = threadIdx.x, G = blockDim.x, lane = tid & 31, warp = tid >> 5;
  for (int p = nt - 1; p >= 0; --p) {
    const double* D = L + (((p * (p + 1)) >> 1) + p) * 64;
    if (warp == 0) {
      // x_p = L(p,p)^-T v_p: 8 dependent steps, lane c keeps v[8p + c]
      double vc = (lane < 8 && p * 8 + lane < n) ? v[p * 8 + lane] : 0.0;
      for (int i = 7; i >= 0; --i) {
        const double xi = __shfl_sync(0xffffffffu, vc, i) * dinv[p * 8 + i];
        if (lane < i) vc = fma(-D[lane * 8 + i], xi, vc);          // L[i][lane]
        if (lane == i) vc = xi;
      }
      if (lane < 8 && p * 8 + lane < n) { v[p * 8 + lane] = vc; a[p * 8 + lane] = vc; }
    }
    __syncthreads();
    // v_c -= sum_{r in tile row p} L[r][c] x[r]  for the columns c < 8p
    for (int c = tid; c < p * 8; c += G) {
      const double* R = L + tiled_idx(p * 8, c);                   // element (8p, c); next row is +1
      double s = v[c];
#pragma unroll
      for (int r = 0; r < 8; ++r) s = fma(-R[r], (p * 8 + r < n) ? v[p * 8 + r] : 0.0, s);
      v[c] = s;
    }
    __syncthreads();
  }
}
#endif

}  // namespace bvg
