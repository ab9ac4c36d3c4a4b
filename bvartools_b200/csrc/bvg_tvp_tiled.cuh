// bvg_tvp_tiled.cuh -- warp-per-chain state draw of the TVP sampler on 8 x 8 tile-packed blocks (device only).
//
// Same recursion as the generic path in bvg_tvp_sweep.cuh (block-bidiagonal Cholesky factor of the state precision,
// bvartvpalg.cpp:290-297), with every n x n operation of a period mapped to FP64 tensor-core tile products
// (mma.sync m8n8k4 = DMMA) issued by the one warp that owns the chain:
//     D_t = B_t - Cq_{t-1}                   built half tile by half tile (32 consecutive doubles per store)
//     L_tt                                    right-looking over nt <= 4 tile columns; diagonal tiles factorised and
//                                             inverted in registers (tiled_factor_invert), panels = A T'
//     M_t = L_tt^-1                           in place, tile row by tile row:  M_IJ = -T_I sum_q L_Iq M_qJ
//     w_t = M_t rhs_t,  Cq_t = Q M_t'M_t Q,  cvec_t = Q M_t' w_t
// M_t is streamed to HBM as nt (nt + 1) / 2 full tiles (coalesced 512-byte rows), read back in the backward pass with a
// one-period register prefetch:  a_t = M_t' (w_t + eps_t + M_t Q a_{t+1}).
// Applies when the chain is owned by a full warp and n <= 32; everything else uses the generic path.
#pragma once
#include "bvg_tiled.cuh"

namespace bvg {

#if defined(__CUDACC__)
#ifdef BVG_TVP_PROF   // cycles per phase of the state draw, accumulated by lane 0 of chain 0 (tools/micro, never in the product build)
static __device__ long long g_tvp_prof[16];
#define BVG_TVP_TICK(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long now_ = clock64(); g_tvp_prof[slot] += now_ - tick_; tick_ = now_; } } while (0)
#define BVG_TVP_TICK0 long long tick_ = clock64();
#else
#define BVG_TVP_TICK(slot) do { } while (0)
#define BVG_TVP_TICK0
#endif
#define BVG_TT(P, I, J) ((P) + (((((I) * ((I) + 1)) >> 1) + (J)) << 6))

// y_I = sum_{J <= I} M_IJ x_J  for all tile rows (lower block-triangular mat-vec); result for row 8 I + (lane & 7) is
// returned in out[I] (valid in every lane).
template <int NTMAX>
__device__ __forceinline__ void tvp_tile_matvec(const double* Mt, const double* x, int nt, double (&out)[NTMAX]) {
  const int lane = threadIdx.x & 31, li = lane & 7, c = (lane >> 3) * 2;
#pragma unroll
  for (int I = 0; I < NTMAX; ++I) {
    double s = 0.0;
    if (I < nt) {
      for (int J = 0; J <= I; ++J) {
        const double* Tl = BVG_TT(Mt, I, J);
        s = fma(Tl[c * 8 + li], x[8 * J + c], s);
        s = fma(Tl[(c + 1) * 8 + li], x[8 * J + c + 1], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
    }
    out[I] = s;
  }
}

// y_J = sum_{I >= J} M_IJ' x_I ; result for column 8 J + (lane & 7) in out[J].
template <int NTMAX>
__device__ __forceinline__ void tvp_tile_matvec_t(const double* Mt, const double* x, int nt, double (&out)[NTMAX]) {
  const int lane = threadIdx.x & 31, li = lane & 7, r = (lane >> 3) * 2;
#pragma unroll
  for (int J = 0; J < NTMAX; ++J) {
    double s = 0.0;
    if (J < nt) {
      for (int I = J; I < nt; ++I) {
        const double* Tl = BVG_TT(Mt, I, J);
        s = fma(Tl[li * 8 + r], x[8 * I + r], s);
        s = fma(Tl[li * 8 + r + 1], x[8 * I + r + 1], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
    }
    out[J] = s;
  }
}

// One sweep's state path.  Shared arrays: A, Cq (tiled_len(n) each); wv, cvec, xl, anext, tmp (npad each); q, a_init
// (npad, zero padded); lam (n) or nullptr; kv (k).  Global: Mpath [T][tiled_len], wpath [T][n], apath [T][n].
static __device__ __noinline__ void tvp_state_draw_tiled(const BvarModel& m, const ChainRng& rng, int sweep, bool sv,
                                                         double* A, double* Cq, double* wv, double* cvec, double* xl,
                                                         double* anext, double* tmp, const double* q,
                                                         const double* a_init, const double* lam, const double* sig,
                                                         const double* ew, double* kv, double* Mpath, double* wpath,
                                                         double* apath, int& fail) {
  constexpr int NTMAX = 4;
  const int k = m.k, T = m.T, M = m.M, n = m.n;
  const int nt = tiled_nt(n), npad = nt * 8, TP = tiled_len(n);
  const int lane = threadIdx.x & 31, fr = lane >> 2, fk = lane & 3, li = lane & 7, lj = lane >> 3;
  const int fo0 = fk * 8 + fr, fo1 = fo0 + 32, co = (2 * fk) * 8 + fr;   // A fragment halves, C fragment
  const int to0 = fr * 8 + fk, to1 = to0 + 4;                            // transposed / B fragment halves
  const float kinv = 1.0f / (float)k;
  const unsigned FULL = 0xffffffffu;

  BVG_TVP_TICK0
  // ================= forward recursion =================
  for (int t = 0; t < T; ++t) {
    const double* xt = m.X + (size_t)M * t;
    const double* yt = m.Y + (size_t)k * t;
    if (lane < k) {
      double s;
      if (sv) s = ew[t + T * lane] * yt[lane];
      else { s = 0.0; for (int i2 = 0; i2 < k; ++i2) s = fma(sig[lane + k * i2], yt[i2], s); }
      kv[lane] = s;
    }
    if (lane < npad) {
      const int jr = (int)(((float)lane + 0.5f) * kinv);
      xl[lane] = (lane < n) ? xt[jr] * (lam != nullptr ? lam[lane] : 1.0) : 0.0;
    }
    __syncwarp();
    const double ct = (t < T - 1) ? 2.0 : 1.0;
    // ---- D_t and rhs_t
    for (int I = 0; I < nt; ++I) {
      const int r = 8 * I + li;
      const int jr = (int)(((float)r + 0.5f) * kinv), ir = r - jr * k;
      const double xr = xl[r];
      for (int J = 0; J <= I; ++J) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = 8 * J + lj + 4 * h;
          const int idx = (((((I * (I + 1)) >> 1) + J)) << 6) + lane + 32 * h;
          double gv = 0.0;
          if (r >= n) gv = (c == r) ? 1.0 : 0.0;
          else if (c <= r) {
            const int jc = (int)(((float)c + 0.5f) * kinv), ic = c - jc * k;
            const double sii = sv ? ((ir == ic) ? ew[t + T * ir] : 0.0) : sig[ir + k * ic];
            gv = xr * xl[c] * sii;
            if (c == r) gv += ct * q[r];
            if (t > 0) gv -= Cq[idx];
          }
          A[idx] = gv;
        }
      }
    }
    if (lane < npad) {
      double b = 0.0;
      if (lane < n) {
        const int jr = (int)(((float)lane + 0.5f) * kinv), ir = lane - jr * k;
        b = xl[lane] * kv[ir] + ((t == 0) ? q[lane] * a_init[lane] : cvec[lane]);        // :296 first block
      }
      wv[lane] = b;
    }
    __syncwarp();
    BVG_TVP_TICK(0);
    // ---- L_tt, diagonal tiles replaced by their inverses
    for (int J = 0; J < nt; ++J) {
      double* D = BVG_TT(A, J, J);
      tiled_factor_invert(D, fail);
      __syncwarp();
      if (J + 1 < nt) {
        const double t0 = D[fo0], t1 = D[fo1];
        for (int I = J + 1; I < nt; ++I) {
          double* TA = BVG_TT(A, I, J);
          double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
          tiled_dmma(c0, c1, TA[fo0], t0);
          tiled_dmma(e0, e1, TA[fo1], t1);
          TA[co] = c0 + e0;
          TA[co + 8] = c1 + e1;
        }
        __syncwarp();
        for (int I = J + 1; I < nt; ++I) {
          const double* TA = BVG_TT(A, I, J);
          const double u0 = -TA[fo0], u1 = -TA[fo1];
          for (int I2 = J + 1; I2 <= I; ++I2) {
            const double* TB = BVG_TT(A, I2, J);
            double* C = BVG_TT(A, I, I2);
            double c0 = C[co], c1 = C[co + 8], e0 = 0.0, e1 = 0.0;
            tiled_dmma(c0, c1, u0, TB[fo0]);
            tiled_dmma(e0, e1, u1, TB[fo1]);
            C[co] = c0 + e0;
            C[co + 8] = c1 + e1;
          }
        }
        __syncwarp();
      }
    }
    BVG_TVP_TICK(1);
    // ---- M_t = L_tt^-1 in place: row I from the finished rows above it
    for (int I = 1; I < nt; ++I) {
      double x0[NTMAX - 1], x1[NTMAX - 1];
#pragma unroll
      for (int J = 0; J < NTMAX - 1; ++J) {
        x0[J] = 0.0; x1[J] = 0.0;
        if (J < I) {
          double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
          for (int qq = J; qq < I; ++qq) {
            const double* LA = BVG_TT(A, I, qq);
            const double* MB = BVG_TT(A, qq, J);
            tiled_dmma(c0, c1, LA[fo0], MB[to0]);
            tiled_dmma(e0, e1, LA[fo1], MB[to1]);
          }
          x0[J] = c0 + e0; x1[J] = c1 + e1;
        }
      }
      __syncwarp();
#pragma unroll
      for (int J = 0; J < NTMAX - 1; ++J)
        if (J < I) { double* X = BVG_TT(A, I, J); X[co] = x0[J]; X[co + 8] = x1[J]; }
      __syncwarp();
      const double* TI = BVG_TT(A, I, I);
      const double a0 = -TI[fo0], a1 = -TI[fo1];
#pragma unroll
      for (int J = 0; J < NTMAX - 1; ++J)
        if (J < I) {
          const double* X = BVG_TT(A, I, J);
          double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
          tiled_dmma(c0, c1, a0, X[to0]);
          tiled_dmma(e0, e1, a1, X[to1]);
          x0[J] = c0 + e0; x1[J] = c1 + e1;
        }
      __syncwarp();
#pragma unroll
      for (int J = 0; J < NTMAX - 1; ++J)
        if (J < I) { double* X = BVG_TT(A, I, J); X[co] = x0[J]; X[co + 8] = x1[J]; }
      __syncwarp();
    }
    BVG_TVP_TICK(2);
    // ---- w_t = M_t rhs_t
    {
      double wI[NTMAX];
      tvp_tile_matvec<NTMAX>(A, wv, nt, wI);
      __syncwarp();
#pragma unroll
      for (int I = 0; I < NTMAX; ++I)
        if (I < nt && lj == 0) wv[8 * I + li] = wI[I];
      __syncwarp();
    }
    BVG_TVP_TICK(3);
    // ---- stream M_t and w_t + eps_t to HBM
    {
      double* Mg = Mpath + (size_t)t * TP;
      for (int e = lane; e < TP; e += 32) Mg[e] = A[e];
      if (lane < n) wpath[(size_t)t * n + lane] = wv[lane] + rng.normal(VB_STATE_N, sweep, t * n + lane);
    }
    BVG_TVP_TICK(4);
    if (t < T - 1) {
      // ---- carry: Cq = Q M'M Q,  cvec = Q M' w
      for (int I = 0; I < nt; ++I) {
        const double qr = q[8 * I + fr];
        for (int J = 0; J <= I; ++J) {
          double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
          for (int qq = I; qq < nt; ++qq) {
            const double* MA = BVG_TT(A, qq, I);
            const double* MB = BVG_TT(A, qq, J);
            tiled_dmma(c0, c1, MA[to0], MB[to0]);
            tiled_dmma(e0, e1, MA[to1], MB[to1]);
          }
          double* C = BVG_TT(Cq, I, J);
          C[co] = (c0 + e0) * qr * q[8 * J + 2 * fk];
          C[co + 8] = (c1 + e1) * qr * q[8 * J + 2 * fk + 1];
        }
      }
      double cJ[NTMAX];
      tvp_tile_matvec_t<NTMAX>(A, wv, nt, cJ);
#pragma unroll
      for (int J = 0; J < NTMAX; ++J)
        if (J < nt && lj == 0) cvec[8 * J + li] = q[8 * J + li] * cJ[J];
    }
    __syncwarp();
    BVG_TVP_TICK(5);
  }

  // ================= backward recursion =================
  constexpr int PF = (NTMAX * (NTMAX + 1) / 2) * 64 / 32;               // doubles per lane of one M_t
  double pre[PF], prez = 0.0;
  {
    const double* Mg = Mpath + (size_t)(T - 1) * TP;
#pragma unroll
    for (int u = 0; u < PF; ++u) pre[u] = (lane + 32 * u < TP) ? Mg[lane + 32 * u] : 0.0;
    prez = (lane < n) ? wpath[(size_t)(T - 1) * n + lane] : 0.0;
  }
  for (int t = T - 1; t >= 0; --t) {
#pragma unroll
    for (int u = 0; u < PF; ++u) if (lane + 32 * u < TP) A[lane + 32 * u] = pre[u];
    if (lane < npad) tmp[lane] = prez;                                 // z_t = w_t + eps_t
    if (t > 0) {                                                       // prefetch period t-1 behind the mat-vecs
      const double* Mg = Mpath + (size_t)(t - 1) * TP;
#pragma unroll
      for (int u = 0; u < PF; ++u) pre[u] = (lane + 32 * u < TP) ? Mg[lane + 32 * u] : 0.0;
      prez = (lane < n) ? wpath[(size_t)(t - 1) * n + lane] : 0.0;
    }
    if (t < T - 1 && lane < npad) xl[lane] = q[lane] * anext[lane];
    __syncwarp();
    if (t < T - 1) {
      double vI[NTMAX];
      tvp_tile_matvec<NTMAX>(A, xl, nt, vI);
#pragma unroll
      for (int I = 0; I < NTMAX; ++I)
        if (I < nt && lj == 0) tmp[8 * I + li] += vI[I];
      __syncwarp();
    }
    double aJ[NTMAX];
    tvp_tile_matvec_t<NTMAX>(A, tmp, nt, aJ);
#pragma unroll
    for (int J = 0; J < NTMAX; ++J)
      if (J < nt && lj == 0) {
        const int r = 8 * J + li;
        anext[r] = aJ[J];
        if (r < n) apath[(size_t)t * n + r] = aJ[J];
      }
    __syncwarp();
  }
  BVG_TVP_TICK(6);
}
#undef BVG_TT
#endif

}  // namespace bvg
