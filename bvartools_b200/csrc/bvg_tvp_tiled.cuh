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
// accumulated in shared memory (an LDS / STS pair per tick; a global read-modify-write would stall the in-order warp for an
// L2 round trip at every tick) and flushed to g_tvp_prof at the end of the kernel (bvg_kern_tvp.cu)
static __device__ long long g_tvp_prof[16];
__shared__ long long s_tvp_prof[16];
#define BVG_TVP_TICK(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long now_ = clock64(); s_tvp_prof[slot] += now_ - tick_; tick_ = now_; } } while (0)
#define BVG_TVP_TICK0 long long tick_ = clock64();
#else
#define BVG_TVP_TICK(slot) do { } while (0)
#define BVG_TVP_TICK0
#endif
#define BVG_TT(P, I, J) ((P) + (((((I) * ((I) + 1)) >> 1) + (J)) << 6))

// y_I = sum_{J <= I} M_IJ x_J  for all tile rows (lower block-triangular mat-vec); result for row 8 I + (lane & 7) is
// returned in out[I] (valid in every lane).
template <int NT>
__device__ __forceinline__ void tvp_tile_matvec(const double* Mt, const double* x, double (&out)[NT]) {
  const int lane = threadIdx.x & 31, li = lane & 7, c = (lane >> 3) * 2;
#pragma unroll
  for (int I = 0; I < NT; ++I) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int J = 0; J <= I; ++J) {
      const double* Tl = BVG_TT(Mt, I, J);
      s0 = fma(Tl[c * 8 + li], x[8 * J + c], s0);
      s1 = fma(Tl[(c + 1) * 8 + li], x[8 * J + c + 1], s1);
    }
    out[I] = s0 + s1;
  }
#pragma unroll
  for (int I = 0; I < NT; ++I) out[I] += __shfl_xor_sync(0xffffffffu, out[I], 8);
#pragma unroll
  for (int I = 0; I < NT; ++I) out[I] += __shfl_xor_sync(0xffffffffu, out[I], 16);
}

// y_J = sum_{I >= J} M_IJ' x_I ; result for column 8 J + (lane & 7) in out[J].
template <int NT>
__device__ __forceinline__ void tvp_tile_matvec_t(const double* Mt, const double* x, double (&out)[NT]) {
  const int lane = threadIdx.x & 31, li = lane & 7, r = (lane >> 3) * 2;
#pragma unroll
  for (int J = 0; J < NT; ++J) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int I = J; I < NT; ++I) {
      const double* Tl = BVG_TT(Mt, I, J);
      s0 = fma(Tl[li * 8 + r], x[8 * I + r], s0);
      s1 = fma(Tl[li * 8 + r + 1], x[8 * I + r + 1], s1);
    }
    out[J] = s0 + s1;
  }
#pragma unroll
  for (int J = 0; J < NT; ++J) out[J] += __shfl_xor_sync(0xffffffffu, out[J], 8);
#pragma unroll
  for (int J = 0; J < NT; ++J) out[J] += __shfl_xor_sync(0xffffffffu, out[J], 16);
}

// One sweep's state path, NT = ceil(n / 8) tile rows at compile time (every tile loop unrolls: the loads of
// independent tile products are issued together).  Shared arrays: A, Cq (tiled_len(n) each); wv, cvec, xl, anext, tmp
// (npad each); q, a_init (npad, zero padded); lam (n) or nullptr; kv (k).  Global: Mpath [T][tiled_len], wpath [T][n],
// apath [T][n].
template <int NT>
static __device__ __noinline__ void tvp_state_draw_tiled_nt(const BvarModel& m, const ChainRng& rng, int sweep, bool sv,
                                                            double* A, double* Cq, double* wv, double* cvec, double* xl,
                                                            double* anext, double* tmp, const double* q,
                                                            const double* a_init, const double* lam, const double* sig,
                                                            const double* ew, double* kv, double* Mpath, double* wpath,
                                                            double* apath, int& fail) {
  constexpr int NPAD = NT * 8, TP = (NT * (NT + 1) / 2) * 64, PF = TP / 32;
  const int k = m.k, T = m.T, M = m.M, n = m.n;
  const int lane = threadIdx.x & 31, fr = lane >> 2, fk = lane & 3, li = lane & 7, lj = lane >> 3;
  const int fo0 = fk * 8 + fr, fo1 = fo0 + 32, co = (2 * fk) * 8 + fr;   // A fragment halves, C fragment
  const int to0 = fr * 8 + fk, to1 = to0 + 4;                            // transposed / B fragment halves
  const float kinv = 1.0f / (float)k;
  BVG_TVP_TICK0

  // ---- all state normals of the sweep in one throughput-bound pass (site = t n + r, Box-Muller pairs): eps -> wpath
  {
    const int nsite = n * T, npair = (nsite + 1) >> 1;
    for (int p = lane; p < npair; p += 32) {
      const Normal2 nn = rng.normal_pair(VB_STATE_N, sweep, p, nsite);
      wpath[2 * p] = nn.a;
      if (2 * p + 1 < nsite) wpath[2 * p + 1] = nn.b;
    }
    __syncwarp();
  }
  BVG_TVP_TICK(7);

  // ================= forward recursion =================
  // x_t, y_t and eps_t of the NEXT period are fetched one period ahead (their L2 latency would otherwise sit in front
  // of every period's build)
  const int jr_lane = (int)(((float)lane + 0.5f) * kinv);
  double x_next = (lane < n) ? m.X[jr_lane] : 0.0;
  double y_next = (lane < k) ? m.Y[lane] : 0.0;
  double eps_next = (lane < n) ? wpath[lane] : 0.0;
  for (int t = 0; t < T; ++t) {
    const double x_cur = x_next, y_cur = y_next, eps_t = eps_next;
    if (t + 1 < T) {
      x_next = (lane < n) ? m.X[(size_t)M * (t + 1) + jr_lane] : 0.0;
      y_next = (lane < k) ? m.Y[(size_t)k * (t + 1) + lane] : 0.0;
      eps_next = (lane < n) ? wpath[(size_t)(t + 1) * n + lane] : 0.0;
    }
    {
      double s;
      if (sv) s = (lane < k) ? ew[t + T * lane] * y_cur : 0.0;
      else {
        s = 0.0;
        for (int i2 = 0; i2 < k; ++i2) {
          const double yv = __shfl_sync(0xffffffffu, y_cur, i2);
          if (lane < k) s = fma(sig[lane + k * i2], yv, s);
        }
      }
      if (lane < k) kv[lane] = s;
    }
    if (lane < NPAD) xl[lane] = (lane < n) ? x_cur * (lam != nullptr ? lam[lane] : 1.0) : 0.0;
    __syncwarp();
    const double ct = (t < T - 1) ? 2.0 : 1.0;
    // ---- D_t and rhs_t (branch-free: every load of the period is issued up front)
#pragma unroll
    for (int I = 0; I < NT; ++I) {
      const int r = 8 * I + li;
      const int jr = (int)(((float)r + 0.5f) * kinv), ir = r - jr * k;
      const bool rin = r < n;
      const double xr = xl[r], qr = q[r];
      const double ewr = sv ? ew[t + T * (rin ? ir : 0)] : 0.0;
#pragma unroll
      for (int J = 0; J <= I; ++J) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = 8 * J + lj + 4 * h;
          const int idx = ((((I * (I + 1)) >> 1) + J) << 6) + lane + 32 * h;
          const int jc = (int)(((float)c + 0.5f) * kinv), ic = c - jc * k;
          const double sii = sv ? ((ir == ic) ? ewr : 0.0) : sig[(rin ? ir : 0) + k * ((c < n) ? ic : 0)];
          const double cq = (t > 0) ? Cq[idx] : 0.0;
          double gv = xr * xl[c] * sii;
          if (I == J) gv += (c == r) ? ct * qr : 0.0;
          gv -= cq;
          const bool valid = rin && (I != J || c <= r);
          const double pad = (I == J && c == r) ? 1.0 : 0.0;
          A[idx] = valid ? gv : pad;
        }
      }
    }
    if (lane < NPAD) {
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
#pragma unroll
    for (int J = 0; J < NT; ++J) {
      double* D = BVG_TT(A, J, J);
      BVG_TVP_TICK(1);
      tiled_factor_invert(D, fail);
      __syncwarp();
      BVG_TVP_TICK(14);
      if (J + 1 < NT) {
        const double t0 = D[fo0], t1 = D[fo1];
        double p0[NT], p1[NT];
#pragma unroll
        for (int I = J + 1; I < NT; ++I) {
          const double* TA = BVG_TT(A, I, J);
          double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
          tiled_dmma(c0, c1, TA[fo0], t0);
          tiled_dmma(e0, e1, TA[fo1], t1);
          p0[I] = c0 + e0; p1[I] = c1 + e1;
        }
#pragma unroll
        for (int I = J + 1; I < NT; ++I) { double* TA = BVG_TT(A, I, J); TA[co] = p0[I]; TA[co + 8] = p1[I]; }
        __syncwarp();
#pragma unroll
        for (int I = J + 1; I < NT; ++I) {
          const double* TA = BVG_TT(A, I, J);
          const double u0 = -TA[fo0], u1 = -TA[fo1];
#pragma unroll
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
#pragma unroll
    for (int I = 1; I < NT; ++I) {
      double x0[NT], x1[NT];
#pragma unroll
      for (int J = 0; J < I; ++J) {
        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
        for (int qq = J; qq < I; ++qq) {
          const double* LA = BVG_TT(A, I, qq);
          const double* MB = BVG_TT(A, qq, J);
          tiled_dmma(c0, c1, LA[fo0], MB[to0]);
          tiled_dmma(e0, e1, LA[fo1], MB[to1]);
        }
        x0[J] = c0 + e0; x1[J] = c1 + e1;
      }
      __syncwarp();
#pragma unroll
      for (int J = 0; J < I; ++J) { double* X = BVG_TT(A, I, J); X[co] = x0[J]; X[co + 8] = x1[J]; }
      __syncwarp();
      const double* TI = BVG_TT(A, I, I);
      const double a0 = -TI[fo0], a1 = -TI[fo1];
#pragma unroll
      for (int J = 0; J < I; ++J) {
        const double* X = BVG_TT(A, I, J);
        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
        tiled_dmma(c0, c1, a0, X[to0]);
        tiled_dmma(e0, e1, a1, X[to1]);
        x0[J] = c0 + e0; x1[J] = c1 + e1;
      }
      __syncwarp();
#pragma unroll
      for (int J = 0; J < I; ++J) { double* X = BVG_TT(A, I, J); X[co] = x0[J]; X[co + 8] = x1[J]; }
      __syncwarp();
    }
    BVG_TVP_TICK(2);
    // ---- w_t = M_t rhs_t ; stream M_t and w_t + eps_t to HBM
    {
      double wI[NT];
      tvp_tile_matvec<NT>(A, wv, wI);
      double* Mg = Mpath + (size_t)t * TP;
#pragma unroll
      for (int u = 0; u < PF; ++u) Mg[lane + 32 * u] = A[lane + 32 * u];
      __syncwarp();
#pragma unroll
      for (int I = 0; I < NT; ++I)
        if (lj == 0) wv[8 * I + li] = wI[I];
      __syncwarp();
      if (lane < n) wpath[(size_t)t * n + lane] = wv[lane] + eps_t;
    }
    BVG_TVP_TICK(3);
    if (t < T - 1) {
      // ---- carry: Cq = Q M'M Q,  cvec = Q M' w
#pragma unroll
      for (int I = 0; I < NT; ++I) {
        const double qr = q[8 * I + fr];
#pragma unroll
        for (int J = 0; J <= I; ++J) {
          double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
          for (int qq = I; qq < NT; ++qq) {
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
      double cJ[NT];
      tvp_tile_matvec_t<NT>(A, wv, cJ);
#pragma unroll
      for (int J = 0; J < NT; ++J)
        if (lj == 0) cvec[8 * J + li] = q[8 * J + li] * cJ[J];
    }
    __syncwarp();
    BVG_TVP_TICK(4);
  }

  // ================= backward recursion =================
  double pre[PF], prez = 0.0;
  {
    const double* Mg = Mpath + (size_t)(T - 1) * TP;
#pragma unroll
    for (int u = 0; u < PF; ++u) pre[u] = Mg[lane + 32 * u];
    prez = (lane < n) ? wpath[(size_t)(T - 1) * n + lane] : 0.0;
  }
  for (int t = T - 1; t >= 0; --t) {
#pragma unroll
    for (int u = 0; u < PF; ++u) A[lane + 32 * u] = pre[u];
    if (lane < NPAD) tmp[lane] = prez;                                 // z_t = w_t + eps_t
    if (t > 0) {                                                       // prefetch period t-1 behind the mat-vecs
      const double* Mg = Mpath + (size_t)(t - 1) * TP;
#pragma unroll
      for (int u = 0; u < PF; ++u) pre[u] = Mg[lane + 32 * u];
      prez = (lane < n) ? wpath[(size_t)(t - 1) * n + lane] : 0.0;
    }
    if (t < T - 1 && lane < NPAD) xl[lane] = q[lane] * anext[lane];
    __syncwarp();
    if (t < T - 1) {
      double vI[NT];
      tvp_tile_matvec<NT>(A, xl, vI);
#pragma unroll
      for (int I = 0; I < NT; ++I)
        if (lj == 0) tmp[8 * I + li] += vI[I];
      __syncwarp();
    }
    double aJ[NT];
    tvp_tile_matvec_t<NT>(A, tmp, aJ);
#pragma unroll
    for (int J = 0; J < NT; ++J)
      if (lj == 0) {
        const int r = 8 * J + li;
        anext[r] = aJ[J];
        if (r < n) apath[(size_t)t * n + r] = aJ[J];
      }
    __syncwarp();
  }
  BVG_TVP_TICK(6);
}

// dispatch on the number of tile rows
static __device__ __forceinline__ void tvp_state_draw_tiled(const BvarModel& m, const ChainRng& rng, int sweep, bool sv,
                                                            double* A, double* Cq, double* wv, double* cvec, double* xl,
                                                            double* anext, double* tmp, const double* q,
                                                            const double* a_init, const double* lam, const double* sig,
                                                            const double* ew, double* kv, double* Mpath, double* wpath,
                                                            double* apath, int& fail) {
#define BVG_TVP_ARGS m, rng, sweep, sv, A, Cq, wv, cvec, xl, anext, tmp, q, a_init, lam, sig, ew, kv, Mpath, wpath, apath, fail
  switch (tiled_nt(m.n)) {
    case 1: tvp_state_draw_tiled_nt<1>(BVG_TVP_ARGS); break;
    case 2: tvp_state_draw_tiled_nt<2>(BVG_TVP_ARGS); break;
    case 3: tvp_state_draw_tiled_nt<3>(BVG_TVP_ARGS); break;
    default: tvp_state_draw_tiled_nt<4>(BVG_TVP_ARGS); break;
  }
#undef BVG_TVP_ARGS
}
#undef BVG_TT
#endif

}  // namespace bvg
