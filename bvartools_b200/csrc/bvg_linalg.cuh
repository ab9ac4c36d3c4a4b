// bvg_linalg.cuh -- group-parallel FP64 dense kernels on shared-memory operands.
//
// Storage: symmetric / triangular matrices are packed lower, row-major: element (i,j), j <= i, at tri(i)+j.
// Row ownership is cyclic and fixed: row i belongs to thread i % G, so a thread re-reads only what it wrote
// between barriers, and the only cross-thread traffic per column is row j (the pivot row).
#pragma once
#include "bvg_group.cuh"

namespace bvg {

// Cholesky-Crout, P = L L', with the right-hand side carried as an extra row (augmented system), so that the
// forward substitution w = L^-1 b costs no extra barriers.
//   in : L   packed lower with P;  b (n) or nullptr
//   out: strict lower part of L = factor, dinv[j] = 1 / L_jj  (the diagonal slots of L are left holding the
//        pivots d_j = L_jj^2 and must not be used);  b <- L^-1 b
//   fail is set when a pivot is not positive (or NaN).
// Reference call sites: arma::chol + arma::solve in bvaralg.cpp:306-307, bvartvpalg.cpp:296-297, :429, :482.
template <class Grp>
BVG_HD void chol_crout_generic(const Grp& g, double* L, double* dinv, double* b, int n, int& fail) {
  const int G = g.size(), tid = g.tid();
  const int bowner = n % G;
  for (int j = 0; j < n; ++j) {
    const double* Lj = L + tri(j);
    int i0 = j - (j % G) + tid;       // first row >= j owned by this thread
    if (i0 < j) i0 += G;
    for (int i = i0; i < n; i += G) {
      double* Li = L + tri(i);
      double s0 = Li[j], s1 = 0.0;
      int c = 0;
      for (; c + 1 < j; c += 2) {
        s0 = fma(-Li[c], Lj[c], s0);
        s1 = fma(-Li[c + 1], Lj[c + 1], s1);
      }
      if (c < j) s0 = fma(-Li[c], Lj[c], s0);
      Li[j] = s0 + s1;
    }
    if (b != nullptr && tid == bowner) {
      double s0 = b[j], s1 = 0.0;
      int c = 0;
      for (; c + 1 < j; c += 2) {
        s0 = fma(-b[c], Lj[c], s0);
        s1 = fma(-b[c + 1], Lj[c + 1], s1);
      }
      if (c < j) s0 = fma(-b[c], Lj[c], s0);
      b[j] = s0 + s1;
    }
    g.sync();
    const double d = Lj[j];
    if (!(d > 0.0)) fail = 1;
    const double rs = 1.0 / sqrt(d);
    for (int i = (i0 == j ? i0 + G : i0); i < n; i += G) L[tri(i) + j] *= rs;
    if (b != nullptr && tid == bowner) b[j] *= rs;
    if (tid == 0) dinv[j] = rs;
    g.sync();
  }
}

// 1 / sqrt(d) for the pivots of the Cholesky factorisations, branch-free: the hardware seed (MUFU.RSQ64H, ~20 bits) and
// one third-order step, y = y0 + y0 e (1/2 + 3/8 e), e = 1 - d y0^2 (error ~ e^3: a few ulp, far inside the 1e-10 parity
// bar).  The library rsqrt() carries a range check and a slow-path call behind a reconvergence barrier: ~74 cycles of
// dependent latency against ~50 here, on every pivot of every chain, and independent calls do not overlap (on the
// latency-bound state recursion of the Kronecker sweep it was where the recursion warp spent its time,
// profiles/r02_c2a_512_*).  Pivots of an SPD matrix are positive normal numbers; zero / negative / NaN arguments give
// inf / NaN, and every caller flags non-positive pivots itself.
BVG_HD double bvg_rsqrt(double d) {
#if defined(__CUDA_ARCH__)
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  const double e = fma(-d, y0 * y0, 1.0);
  const double p = fma(e, 0.375, 0.5);
  return fma(p, y0 * e, y0);
#else
  return 1.0 / sqrt(d);
#endif
}
BVG_HD double bvg_rsqrt_fast(double d) { return bvg_rsqrt(d); }

// Same factorisation for n + 1 <= G <= 32: ONE row per lane (lane r owns row r, lane n owns the rhs row).  The
// pivot is broadcast by shuffle instead of a shared-memory round trip, so a column costs one barrier.  The dot
// product over the j already-final columns enters a 31-deep unrolled chain at depth j (switch with fall-through,
// one uniform indirect branch per column): its body is two LDS.64 with immediate offsets and one DFMA per
// element, without loop control, and -- unlike a fully unrolled factorisation -- the same ~100 instructions are
// reused by every column, so the kernel stays inside the instruction cache.
#define BVG_ROWLANE_MAX 32
#define BVG_DOT_CASE(c) case (c) + 1: acc[(c) & 3] = fma(-Lr[c], Lj[c], acc[(c) & 3]);
template <class Grp>
BVG_HD void chol_crout_rowlane(const Grp& g, double* L, double* dinv, double* b, int n, int& fail) {
  const int r = g.tid();
  const bool act = (r < n) || (b != nullptr && r == n);
  double* Lr = (r < n) ? L + tri(r) : (b != nullptr ? b : L);
  const double* Lj = L;
  for (int j = 0; j < n; ++j) {
    double s = 0.0;
    if (act && r >= j) {
      double acc[4] = {Lr[j], 0.0, 0.0, 0.0};
      switch (j) {
        BVG_DOT_CASE(30) BVG_DOT_CASE(29) BVG_DOT_CASE(28) BVG_DOT_CASE(27) BVG_DOT_CASE(26) BVG_DOT_CASE(25)
        BVG_DOT_CASE(24) BVG_DOT_CASE(23) BVG_DOT_CASE(22) BVG_DOT_CASE(21) BVG_DOT_CASE(20) BVG_DOT_CASE(19)
        BVG_DOT_CASE(18) BVG_DOT_CASE(17) BVG_DOT_CASE(16) BVG_DOT_CASE(15) BVG_DOT_CASE(14) BVG_DOT_CASE(13)
        BVG_DOT_CASE(12) BVG_DOT_CASE(11) BVG_DOT_CASE(10) BVG_DOT_CASE(9) BVG_DOT_CASE(8) BVG_DOT_CASE(7)
        BVG_DOT_CASE(6) BVG_DOT_CASE(5) BVG_DOT_CASE(4) BVG_DOT_CASE(3) BVG_DOT_CASE(2) BVG_DOT_CASE(1)
        BVG_DOT_CASE(0)
        default: break;
      }
      s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    }
    const double d = g.bcast(s, j);
    if (!(d > 0.0)) fail = 1;
    const double rs = bvg_rsqrt(d);
    if (act && r > j) Lr[j] = s * rs;
    if (r == j) dinv[j] = rs;
    Lj += j + 1;
    g.sync();
  }
}
#undef BVG_DOT_CASE

// Fully unrolled variant (column index compile-time, early exit on n dynamic): ~15 % fewer executed instructions
// than the fall-through chain but ~2000 instructions of straight-line code per call site -- used once, for the
// n x n factor of the constant-parameter sweep kernel (profiles/r01_sweep_kernel_notes.md).
template <class Grp>
BVG_HD void chol_crout_rowlane_unrolled(const Grp& g, double* L, double* dinv, double* b, int n, int& fail) {
  const int r = g.tid();
  const bool act = (r < n) || (b != nullptr && r == n);
  double* Lr = (r < n) ? L + tri(r) : (b != nullptr ? b : L);
  BVG_UNROLL
  for (int j = 0; j < BVG_ROWLANE_MAX; ++j) {
    if (j >= n) break;
    const double* Lj = L + tri(j);
    double s = 0.0;
    if (act && r >= j) {
      double acc[4] = {Lr[j], 0.0, 0.0, 0.0};
      BVG_UNROLL
      for (int c = 0; c < j; ++c) acc[c & 3] = fma(-Lr[c], Lj[c], acc[c & 3]);
      s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    }
    const double d = g.bcast(s, j);
    if (!(d > 0.0)) fail = 1;
    const double rs = bvg_rsqrt(d);
    if (act && r > j) Lr[j] = s * rs;
    if (r == j) dinv[j] = rs;
    g.sync();
  }
}

// a = L^-T w for n <= G <= 32: lane r keeps w_r in a register, a_i is broadcast by shuffle; no barriers in the loop.
template <class Grp>
BVG_HD void back_solve_lt_rowlane(const Grp& g, const double* L, const double* dinv, const double* w, double* a,
                                  int n) {
  const int r = g.tid();
  double wr = (r < n) ? w[r] : 0.0;
  const double dr = (r < n) ? dinv[r] : 0.0;
  const double* Li = L + tri(n - 1) + ((r < n) ? r : 0);      // &L[n-1][r]
  for (int i = n - 1; i >= 0; --i) {
    const double ai = g.bcast(wr * dr, i);
    if (r < i) wr = fma(-Li[0], ai, wr);
    if (r == i) a[i] = ai;
    Li -= i;
  }
  g.sync();
}

// a = L^-T w  (back substitution with the transposed factor); w is destroyed.  a may alias nothing in w.
template <class Grp>
BVG_HD void back_solve_lt_generic(const Grp& g, const double* L, const double* dinv, double* w, double* a, int n) {
  const int G = g.size(), tid = g.tid();
  for (int i = n - 1; i >= 0; --i) {
    const double ai = w[i] * dinv[i];
    const double* Li = L + tri(i);
    for (int r = tid; r < i; r += G) w[r] = fma(-Li[r], ai, w[r]);
    if (tid == 0) a[i] = ai;
    g.sync();
  }
}

template <class Grp>
BVG_HD void chol_crout(const Grp& g, double* L, double* dinv, double* b, int n, int& fail) {
  if (Grp::kCheapBcast && n + 1 <= g.size() && n <= BVG_ROWLANE_MAX) chol_crout_rowlane(g, L, dinv, b, n, fail);
  else chol_crout_generic(g, L, dinv, b, n, fail);
}

// call-site choice: UNROLLED = true only where the straight-line code is affordable
template <bool UNROLLED, class Grp>
BVG_HD void chol_crout_sel(const Grp& g, double* L, double* dinv, double* b, int n, int& fail) {
  if (Grp::kCheapBcast && n + 1 <= g.size() && n <= BVG_ROWLANE_MAX) {
    if (UNROLLED) chol_crout_rowlane_unrolled(g, L, dinv, b, n, fail);
    else chol_crout_rowlane(g, L, dinv, b, n, fail);
  } else {
    chol_crout_generic(g, L, dinv, b, n, fail);
  }
}

template <class Grp>
BVG_HD void back_solve_lt(const Grp& g, const double* L, const double* dinv, double* w, double* a, int n) {
  if (Grp::kCheapBcast && n <= g.size() && n <= BVG_ROWLANE_MAX) back_solve_lt_rowlane(g, L, dinv, w, a, n);
  else back_solve_lt_generic(g, L, dinv, w, a, n);
}

// w = L^-1 b in place (forward substitution), for factors whose rhs was not carried through chol_crout.
template <class Grp>
BVG_HD void fwd_solve_l(const Grp& g, const double* L, const double* dinv, double* b, int n) {
  const int G = g.size(), tid = g.tid();
  for (int j = 0; j < n; ++j) {
    const double wj = b[j] * dinv[j];
    g.sync();                                  // everyone has read b[j] before the owner overwrites it
    if (tid == 0) b[j] = wj;
    for (int i = j + 1 + tid; i < n; i += G) b[i] = fma(-L[tri(i) + j], wj, b[i]);
    g.sync();
  }
}

// M = L^-1 (lower triangular inverse, packed), column per thread.  L's diagonal comes from dinv.
template <class Grp>
BVG_HD void tri_inverse(const Grp& g, const double* L, const double* dinv, double* Minv, int n) {
  const int G = g.size(), tid = g.tid();
  for (int c = tid; c < n; c += G) {
    Minv[tri(c) + c] = dinv[c];
    for (int i = c + 1; i < n; ++i) {
      const double* Li = L + tri(i);
      double s = 0.0;
      for (int m = c; m < i; ++m) s = fma(Li[m], Minv[tri(m) + c], s);
      Minv[tri(i) + c] = -s * dinv[i];
    }
  }
  g.sync();
}

// C (packed lower) = M' M for lower-triangular packed M:  C_ij = sum_{m >= i} M_mi M_mj  (i >= j)
template <class Grp>
BVG_HD void mtm_lower(const Grp& g, const double* Minv, double* C, int n) {
  const int G = g.size(), tid = g.tid();
  const int ne = tri(n);
  int i = 0;
  for (int e = tid; e < ne; e += G) {
    while (tri(i + 1) <= e) ++i;
    const int j = e - tri(i);
    double s = 0.0;
    for (int m = i; m < n; ++m) s = fma(Minv[tri(m) + i], Minv[tri(m) + j], s);
    C[e] = s;
  }
  g.sync();
}

// C (packed lower) = M M' for lower-triangular packed M:  C_ij = sum_{m <= j} M_im M_jm  (i >= j)
template <class Grp>
BVG_HD void mmt_lower(const Grp& g, const double* Mlow, double* C, int n) {
  const int G = g.size(), tid = g.tid();
  const int ne = tri(n);
  int i = 0;
  for (int e = tid; e < ne; e += G) {
    while (tri(i + 1) <= e) ++i;
    const int j = e - tri(i);
    double s = 0.0;
    for (int m = 0; m <= j; ++m) s = fma(Mlow[tri(i) + m], Mlow[tri(j) + m], s);
    C[e] = s;
  }
  g.sync();
}

// write the true diagonal (1/dinv) into the packed factor so it can be used as a plain triangular matrix
template <class Grp>
BVG_HD void fix_diag(const Grp& g, double* L, const double* dinv, int n) {
  for (int i = g.tid(); i < n; i += g.size()) L[tri(i) + i] = 1.0 / dinv[i];
  g.sync();
}

}  // namespace bvg
