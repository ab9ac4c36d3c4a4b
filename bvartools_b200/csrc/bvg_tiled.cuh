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
#include "bvg_linalg.cuh"

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
// zero upper triangle.  EVERY lane factorises the whole tile redundantly in registers (36 broadcast loads, 112 FP64
// operations, no shuffles: the chain from one pivot to the next is rsqrt -> multiply -> FMA); lane c (mod 8) then forms
// column c of T by forward substitution.  Measured against a row-per-lane variant with shuffle broadcasts (132 cycles
// of dependent latency per pivot plus 128 SHFL): 2.6 k -> see tools/micro/chol_prof, tvp_prof.
__device__ __forceinline__ void tiled_factor_invert(double* D, int& fail) {
  const int lane = threadIdx.x & 31, c = lane & 7;
  double L[8][8], rsv[8], x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = j; i < 8; ++i) L[i][j] = D[j * 8 + i];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double d = L[j][j];
    bad = bad || !(d > 0.0);
    const double rs = bvg_rsqrt(d);
    rsv[j] = rs;
#pragma unroll
    for (int i = j + 1; i < 8; ++i) L[i][j] *= rs;
#pragma unroll
    for (int cc = j + 1; cc < 8; ++cc)
#pragma unroll
      for (int i = cc; i < 8; ++i) L[i][cc] = fma(-L[i][j], L[cc][j], L[i][cc]);
  }
  if (bad) fail = 1;
  // column c of T: x_c = 1 / L_cc, x_r = -(sum_{c <= m < r} L_rm x_m) / L_rr; entries above the diagonal stay zero
#pragma unroll
  for (int r = 0; r < 8; ++r) x[r] = (r == c) ? rsv[r] : 0.0;
#pragma unroll
  for (int r = 1; r < 8; ++r) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int m = 0; m < r; ++m) {
      if (m & 1) s1 = fma(L[r][m], x[m], s1); else s0 = fma(L[r][m], x[m], s0);
    }
    if (r > c) x[r] = -rsv[r] * (s0 + s1);
  }
  __syncwarp();
  if (lane < 8) {
#pragma unroll
    for (int r = 0; r < 8; ++r) D[c * 8 + r] = x[r];
  }
}

#ifdef BVG_TILED_PROF   // tools/micro/chol_prof.cu: cycles per phase seen by thread 0 (slots 0-3) and by warp 1 (4-5)
#define BVG_TP(slot) do { if (prof != nullptr && (tid == 0 || tid == 32)) { const long long now_ = clock64(); if ((slot) >= 0 && (tid == 0) == ((slot) < 4 || (slot) == 6)) prof[slot] += now_ - tp_; tp_ = now_; } } while (0)
#define BVG_TP_ARG , long long* prof
#define BVG_TP_INIT long long tp_ = clock64();
#else
#define BVG_TP(slot) do { } while (0)
#define BVG_TP_ARG
#define BVG_TP_INIT
#endif

// In-place Cholesky of the tile-packed matrix, left-looking by tile column with the accumulators in registers (no
// read-modify-write of the trailing matrix in shared memory), except the diagonal tiles, which are kept up to date
// right-looking so that warp 0 can start on the next one without an accumulation pass.  On exit the strictly-lower tiles hold L, the diagonal
// tiles hold L(J,J)^-1, and b holds w = L^-1 b.  fail is set on a non-positive pivot.
static __device__ __noinline__ void tiled_chol(double* L, double* b, int n, int& fail BVG_TP_ARG) {
  const int nt = tiled_nt(n), tid = threadIdx.x, G = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = G >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int fo0 = fk * 8 + fr, fo1 = (4 + fk) * 8 + fr, co = (2 * fk) * 8 + fr;
  const bool solo = (nwarp == 1);
  const int nwork = solo ? 1 : nwarp - 1;                         // warps 1.. (warp 0 owns the diagonal tile)
  const int wid = solo ? 0 : warp - 1;
  BVG_TP_INIT
  for (int J = 0; J < nt; ++J) {
    double* rowJ = L + ((J * (J + 1)) >> 1) * 64;
    double* D = rowJ + J * 64;
    // ---- phase 1: A(I,J) -= sum_{q<J} L(I,q) L(J,q)' for I > J; warp 0 factorises the diagonal tile, which is
    //      already up to date (the panel phase applies every column to the diagonal tiles below it)
    if (warp == 0) {
      tiled_factor_invert(D, fail);
      BVG_TP(2);
    }
    if (warp != 0 || solo) {
      for (int I = J + 1 + wid; I < nt; I += 2 * nwork) {
        const int I2 = I + nwork;
        double* r1 = L + ((I * (I + 1)) >> 1) * 64;
        if (I2 < nt) {                                            // two row tiles share the L(J,q) fragments
          double* r2 = L + ((I2 * (I2 + 1)) >> 1) * 64;
          double c0 = r1[J * 64 + co], c1 = r1[J * 64 + co + 8], e0 = 0.0, e1 = 0.0;
          double g0 = r2[J * 64 + co], g1 = r2[J * 64 + co + 8], h0 = 0.0, h1 = 0.0;
#pragma unroll 2
          for (int q = 0; q < J; ++q) {
            const double v0 = rowJ[q * 64 + fo0], v1 = rowJ[q * 64 + fo1];
            const double u0 = r1[q * 64 + fo0], u1 = r1[q * 64 + fo1];
            const double y0 = r2[q * 64 + fo0], y1 = r2[q * 64 + fo1];
            tiled_dmma(c0, c1, -u0, v0);
            tiled_dmma(e0, e1, -u1, v1);
            tiled_dmma(g0, g1, -y0, v0);
            tiled_dmma(h0, h1, -y1, v1);
          }
          r1[J * 64 + co] = c0 + e0;  r1[J * 64 + co + 8] = c1 + e1;
          r2[J * 64 + co] = g0 + h0;  r2[J * 64 + co + 8] = g1 + h1;
        } else {
          double c0 = r1[J * 64 + co], c1 = r1[J * 64 + co + 8], e0 = 0.0, e1 = 0.0;
#pragma unroll 4
          for (int q = 0; q < J; ++q) {
            const double v0 = rowJ[q * 64 + fo0], v1 = rowJ[q * 64 + fo1];
            tiled_dmma(c0, c1, -r1[q * 64 + fo0], v0);
            tiled_dmma(e0, e1, -r1[q * 64 + fo1], v1);
          }
          r1[J * 64 + co] = c0 + e0;  r1[J * 64 + co + 8] = c1 + e1;
        }
      }
      if (wid == nwork - 1) {
        // rhs: b_J -= sum_{q<J} L(J,q) w_q ; lane = (row i, quarter of the q range)
        const int i = lane & 7, part = lane >> 3;
        double s = 0.0;
        for (int q = part; q < J; q += 4) {
#pragma unroll
          for (int c = 0; c < 8; ++c) s = fma(rowJ[q * 64 + c * 8 + i], b[q * 8 + c], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        if (lane < 8) b[J * 8 + i] -= s;
      }
      BVG_TP(4);
    }
    __syncthreads();
    BVG_TP(tid == 0 ? 3 : 5);
    // ---- phase 2: panel L(I,J) = A(I,J) T_J' (two DMMAs per tile), w_J = T_J b_J
    {
      const double t0 = D[fo0], t1 = D[fo1];                      // B fragment: B[k][n] = T[n][k]
      for (int I = J + 1 + warp; I < nt; I += nwarp) {
        double* TA = L + (((I * (I + 1)) >> 1) + J) * 64;
        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
        tiled_dmma(c0, c1, TA[fo0], t0);
        tiled_dmma(e0, e1, TA[fo1], t1);
        TA[co] = c0 + e0;
        TA[co + 8] = c1 + e1;
        __syncwarp();
        // right-looking for the diagonal tile of this row only: A(I,I) -= L(I,J) L(I,J)'
        double* DI = L + (((I * (I + 1)) >> 1) + I) * 64;
        const double u0 = TA[fo0], u1 = TA[fo1];
        double d0 = DI[co], d1 = DI[co + 8], f0 = 0.0, f1 = 0.0;
        tiled_dmma(d0, d1, -u0, u0);
        tiled_dmma(f0, f1, -u1, u1);
        DI[co] = d0 + f0;
        DI[co + 8] = d1 + f1;
      }
      if (warp == nwarp - 1) {
        double s = 0.0;
        const int i = lane & 7;
#pragma unroll
        for (int c = 0; c < 8; ++c) s = fma(D[c * 8 + i], b[J * 8 + c], s);
        __syncwarp();
        if (lane < 8) b[J * 8 + i] = s;                           // w_J = T_J b_J
      }
    }
    __syncthreads();
    BVG_TP(0);
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
    // v_c -= sum_{r in tile row p} L[r][c] x[r]  for the columns c < 8p; the row order is rotated per lane so that the
    // eight lanes of a tile hit different banks
    for (int c = tid; c < p * 8; c += G) {
      const double* R = L + tiled_idx(p * 8, c);                   // element (8p, c); next row is +1
      const double* xp = v + p * 8;
      double s0 = v[c], s1 = 0.0;
#pragma unroll
      for (int t = 0; t < 8; t += 2) {
        const int ra = (c + t) & 7, rb = (c + t + 1) & 7;
        s0 = fma(-R[ra], xp[ra], s0);
        s1 = fma(-R[rb], xp[rb], s1);
      }
      v[c] = s0 + s1;
    }
    __syncthreads();
  }
}
#endif

}  // namespace bvg
