// bvg_tiled.cuh -- CTA-wide FP64 Cholesky / triangular solves on an 8 x 8 tile-packed lower triangle in shared memory,
// trailing updates on the FP64 tensor pipe (mma.sync m8n8k4 f64 = DMMA).  Used by the CTA-per-chain variant of the
// constant-parameter sweep (config c3: n = 150) in place of the row-cyclic Crout factorisation, whose n^2/2 dependent
// dot products per sweep left the SM idle.  Device only.
//
// Layout: the matrix is padded to nt = ceil(n / 8) tile rows.  Tile (I, J), J <= I, lives at (I (I + 1) / 2 + J) * 64,
// column-major inside the tile (element (i, j) at j * 8 + i): a warp reads a DMMA A/B fragment (32 doubles) from one
// contiguous, conflict-free half tile.  The padding rows are an identity block.  The right-hand side b (npad doubles,
// zero padded) is carried along: on exit it holds w = L^-1 b.
//
// Algorithm (right-looking with look-ahead), per tile column p:
//   panel     L(I,p) = A(I,p) T_p'  with T_p = L(p,p)^-1  -- two DMMAs per tile, no dependent substitution chain
//   trailing  A(I,J) -= L(I,p) L(J,p)'                     -- two DMMAs per tile, accumulating into the tile
//   look-ahead: warp 0 updates tile (p+1,p+1) first and factorises + inverts it in registers (row per lane, shuffle
//   broadcasts) while the other warps finish the trailing update.  T_p overwrites the diagonal tile (L(p,p) itself is
//   not needed again: the back substitution uses T_p' as well).
// Reference call sites: arma::chol + arma::solve in bvaralg.cpp:306-307.
#pragma once
#include "bvg_group.cuh"

namespace bvg {

BVG_HD int tiled_nt(int n) { return (n + 7) >> 3; }
BVG_HD int tiled_len(int n) { const int r = tiled_nt(n); return ((r * (r + 1)) >> 1) * 64; }
BVG_HD int tiled_idx(int i, int j) {      // element (i, j), j <= i (tile-wise: J <= I)
  const int I = i >> 3, J = j >> 3;
  return ((((I * (I + 1)) >> 1) + J) << 6) + ((j & 7) << 3) + (i & 7);
}

#if defined(__CUDACC__)
__device__ __forceinline__ void tiled_dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// identity padding rows and zero padding of the right-hand side b (length npad); call after the n x n lower triangle
// and b[0..n) have been written, before tiled_chol
__device__ __forceinline__ void tiled_finish_build(double* L, double* b, int n) {
  const int nt = tiled_nt(n), npad = nt * 8, tid = threadIdx.x, G = blockDim.x;
  for (int e = tid; e < (npad - n) * npad; e += G) {
    const int i = n + e / npad, j = e % npad;
    if (j <= i) L[tiled_idx(i, j)] = (i == j) ? 1.0 : 0.0;
  }
  for (int e = n + tid; e < npad; e += G) b[e] = 0.0;
  __syncthreads();
}

// One warp: Cholesky of the 8 x 8 diagonal tile D (lower part valid) and T = L^-1, written over D with an exactly
// zero upper triangle.  Lane i (mod 8) keeps row i in registers; pivots and multipliers travel by shuffle.
__device__ __forceinline__ void tiled_factor_invert(double* D, int& fail) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, i = lane & 7;
  double a[8], rsv[8], x[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) a[c] = D[c * 8 + i];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double d = __shfl_sync(FULL, a[j], j);
    if (!(d > 0.0)) fail = 1;
    const double rs = rsqrt(d);
    rsv[j] = rs;
    const double l = a[j] * rs;          // lane i >= j: L_ij   (lane j: sqrt(d))
    a[j] = l;
#pragma unroll
    for (int c = j + 1; c < 8; ++c) a[c] = fma(-l, __shfl_sync(FULL, l, c), a[c]);
  }
  // column c = i of T: x_c = 1 / L_cc, x_r = -(sum_{c <= m < r} L_rm x_m) / L_rr
#pragma unroll
  for (int r = 0; r < 8; ++r) x[r] = (r == i) ? rsv[r] : 0.0;
#pragma unroll
  for (int r = 1; r < 8; ++r) {
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < r; ++m) s = fma(__shfl_sync(FULL, a[m], r), x[m], s);
    if (r > i) x[r] = -rsv[r] * s;
  }
  __syncwarp();
  if (lane < 8) {
#pragma unroll
    for (int r = 0; r < 8; ++r) D[i * 8 + r] = x[r];
  }
}

// In-place Cholesky of the tile-packed matrix.  On exit the strictly-lower tiles hold L, the diagonal tiles hold
// L(p,p)^-1, and b holds w = L^-1 b.  fail is set on a non-positive pivot.
static __device__ __noinline__ void tiled_chol(double* L, double* b, int n, int& fail) {
  const int nt = tiled_nt(n), npad = nt * 8, tid = threadIdx.x, G = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = G >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int fo0 = fk * 8 + fr, fo1 = (4 + fk) * 8 + fr, co = (2 * fk) * 8 + fr;
  if (warp == 0) tiled_factor_invert(L, fail);
  __syncthreads();
  for (int p = 0; p < nt; ++p) {
    const double* Ti = L + (((p * (p + 1)) >> 1) + p) * 64;      // T_p = L(p,p)^-1
    // ---- panel: L(I,p) = A(I,p) T_p', one tile per warp and pass; the last warp also maps the rhs block
    {
      const double t0 = Ti[fo0], t1 = Ti[fo1];                    // B fragment: B[k][n] = T[n][k]
      for (int I = p + 1 + warp; I < nt; I += nwarp) {
        double* TA = L + (((I * (I + 1)) >> 1) + p) * 64;
        double c0 = 0.0, c1 = 0.0;
        tiled_dmma(c0, c1, TA[fo0], t0);
        tiled_dmma(c0, c1, TA[fo1], t1);
        TA[co] = c0;
        TA[co + 8] = c1;
      }
      if (warp == nwarp - 1) {
        double s = 0.0;
        const int i = lane & 7;
#pragma unroll
        for (int c = 0; c < 8; ++c) s = fma(Ti[c * 8 + i], b[p * 8 + c], s);
        __syncwarp();
        if (lane < 8) b[p * 8 + i] = s;                           // w_p = T_p b_p
      }
    }
    __syncthreads();
    if (p + 1 == nt) break;
    // ---- trailing update of the rows below: rhs first, then the tiles
    for (int r = (p + 1) * 8 + tid; r < npad; r += G) {
      const double* R = L + tiled_idx(r, p * 8);                  // element (r, 8p); next column is +8
      double s = b[r];
#pragma unroll
      for (int c = 0; c < 8; ++c) s = fma(-R[c * 8], b[p * 8 + c], s);
      b[r] = s;
    }
    const int m = nt - 1 - p;                                     // tile rows p+1 .. nt-1
    const int ntile = (m * (m + 1)) >> 1;
    // tile t = a (a + 1) / 2 + bc  <->  (I, J) = (p + 1 + a, p + 1 + bc); t = 0 is the next diagonal tile
    int t, stride;
    if (nwarp == 1) { t = 0; stride = 1; }
    else if (warp == 0) { t = 0; stride = ntile; }                // warp 0: tile (p+1,p+1) only, then the look-ahead
    else { t = warp; stride = nwarp - 1; }
    int a = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while (((a + 1) * (a + 2)) >> 1 <= t) ++a;
    while (((a * (a + 1)) >> 1) > t) --a;
    int bc = t - ((a * (a + 1)) >> 1);
    for (; t < ntile; t += stride) {
      const int I = p + 1 + a, J = p + 1 + bc;
      const double* TA = L + (((I * (I + 1)) >> 1) + p) * 64;
      const double* TB = L + (((J * (J + 1)) >> 1) + p) * 64;
      double* C = L + (((I * (I + 1)) >> 1) + J) * 64 + co;
      double c0 = C[0], c1 = C[8];
      tiled_dmma(c0, c1, -TA[fo0], TB[fo0]);
      tiled_dmma(c0, c1, -TA[fo1], TB[fo1]);
      C[0] = c0;
      C[8] = c1;
      bc += stride;
      while (bc > a) { bc -= a + 1; ++a; }
    }
    if (warp == 0) {
      __syncwarp();
      tiled_factor_invert(L + ((((p + 1) * (p + 2)) >> 1) + p + 1) * 64, fail);
    }
    __syncthreads();
  }
}

// a = L^-T v (back substitution with the transposed factor; diagonal tiles hold L(p,p)^-1); v (npad, shared, zero
// padded) is destroyed.
static __device__ __noinline__ void tiled_back_solve(const double* L, double* v, double* a, int n) {
  const int nt = tiled_nt(n), tid = threadIdx.x, G = blockDim.x, lane = tid & 31, warp = tid >> 5;
  for (int p = nt - 1; p >= 0; --p) {
    const double* Ti = L + (((p * (p + 1)) >> 1) + p) * 64;
    if (warp == 0) {
      // x_p = T_p' v_p
      const int i = lane & 7;
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < 8; ++r) s = fma(Ti[i * 8 + r], v[p * 8 + r], s);
      __syncwarp();
      if (lane < 8) { v[p * 8 + i] = s; if (p * 8 + i < n) a[p * 8 + i] = s; }
    }
    __syncthreads();
    // v_c -= sum_{r in tile row p} L[r][c] x[r]  for the columns c < 8p
    for (int c = tid; c < p * 8; c += G) {
      const double* R = L + tiled_idx(p * 8, c);                   // element (8p, c); next row is +1
      double s = v[c];
#pragma unroll
      for (int r = 0; r < 8; ++r) s = fma(-R[r], v[p * 8 + r], s);
      v[c] = s;
    }
    __syncthreads();
  }
}
#endif

}  // namespace bvg
