// bvg_tvp_eq8.cuh -- equation-decoupled state draw of the TVP sampler, warp per chain, 8 lanes per equation (device only).
//
// With a DIAGONAL measurement precision (stochastic volatility or inverse-gamma variances without the covariance block:
// Sigma_t^-1 = diag(omega_t)) the blocks of the state precision of bvartvpalg.cpp:290-297 are
//     B_t = c_t Q + (x_t x_t') (x) diag(omega_t),          Q = diag(Sigma_v^-1) diagonal,
// i.e. coefficient (regressor j, equation i) only couples with the coefficients of the SAME equation: the (nT) x (nT)
// precision is -- up to a symmetric permutation that keeps the order inside every equation -- block diagonal over the k
// equations.  The Cholesky factor of a reducible SPD matrix is the union of the factors of its components (same argument
// as for the psi block, bvg_tvp_sweep.cuh), so the reference's draw  a = P^-1 rhs + chol(P)^-1 eps  EQUALS k independent
// draws of dimension M T (M regressors per equation) fed with the matching entries of eps: M x M blocks instead of
// n x n (n = k M), k^2 fewer flops, a pivot chain of M instead of n per period, and k tri(M) instead of tri(n) streamed
// doubles per period.
//
// Mapping: lanes 8 i .. 8 i + 7 of the chain's warp own equation i (k <= 4, M <= 8).  Per period every lane of a group
// factorises the group's M x M block D_t REDUNDANTLY in registers (no shuffles on the pivot chain: rsqrt -> multiply ->
// FMA), lane c then forms column c of M_t = L_tt^-1 (forward substitution) and column c of D_t^-1 = L_tt^-T L_tt^-1
// (back substitution) locally.  With  z_t = L_tt^-1 rhs_t + eps_t  the recursions of the file header of bvg_tvp_sweep.cuh
// become
//     forward :  D_t = B_t - Cq_{t-1},   Cq_t = Q D_t^-1 Q,   cvec_t = Q M_t' w_t,   g_t = M_t' z_t
//     backward:  a_t = g_t + D_t^-1 Q a_{t+1} = g_t + Q^-1 Cq_t a_{t+1}
// so only Cq_t (packed lower, 36 slots) and g_t (8 slots) per equation are streamed through HBM: written forward, read
// backward -- 44 k doubles per period instead of tiled_len(n) + n.  The lanes of a group exchange Cq / cvec / x_t / eps_t
// through a few hundred bytes of shared memory, two __syncwarp per period.
#pragma once
#include "bvg_tiled.cuh"

namespace bvg {

// 1 = this path applies (host and device agree on it: scratch / shared-memory sizes depend on it)
BVG_HD bool tvp_eq8_ok(int k, int M, int n_psi, int diag_sigma) { return diag_sigma != 0 && n_psi == 0 && k <= 4 && M >= 1 && M <= 8; }
BVG_HD int tvp_eq8_smem(int k) { return (44 * k + 12) + 4 * 8 * k + 4; }             // block (CQ|g, x, y), cvec, x, eps, a_next, y
BVG_HD int tvp_eq8_stream(int k) { return 44 * k + 12; }                             // doubles per period in HBM

#if defined(__CUDACC__)
#define BVG_EQ8_OFF(c) (8 * (c) - (((c) * ((c) - 1)) >> 1))      // first slot of packed column c (rows c..7)

// 1 / sqrt(d) for the pivots: the hardware seed (MUFU.RSQ64H, ~20 bits) and one third-order step,
// y = y0 + y0 e (1/2 + 3/8 e), e = 1 - d y0^2 -- the fast path of the library's rsqrt() without its range check, call and
// reconvergence barrier (7 of them per period sit on the pivot chain).  Pivots of an SPD block are normal numbers; zero /
// negative / NaN pivots give inf / NaN, which the caller flags.
__device__ __forceinline__ double eq8_rsqrt(double d) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  const double e = fma(-d, y0 * y0, 1.0);
  const double p = fma(e, 0.375, 0.5);
  return fma(p, y0 * e, y0);
}

// X, Y, k, T, n by value: nothing in the period loops goes through the model struct (a by-reference kernel parameter
// lives in local memory, and with ~200 KB of shared memory per SM the L1 that would cache it is a few KB: a miss is an L2
// round trip on the critical path).  Forward and backward passes are separate functions so that each gets its own register
// allocation (the backward ring is 96 registers that the forward loop needs for the factor).
// Values that must live in registers across the period loops.  Kernel parameters and function arguments are otherwise
// re-materialised by ptxas from the constant bank / the stack at the top of every period (an L2 round trip when the small
// caches miss); the result of a shuffle is opaque to it.  (The lane reads its own value: identity.)
// Pointers keep their address space: global ones through __builtin_assume(__isGlobal), shared ones as an (opaque) offset
// into the dynamic shared-memory array -- a pointer rebuilt from two shuffled halves would be generic, and every access
// through it a generic load with 64-bit address arithmetic.
extern __shared__ double eq8_dyn_smem[];
__device__ __forceinline__ int eq8_pin(int v) { return __shfl_sync(0xffffffffu, v, threadIdx.x & 31); }
template <class P>
__device__ __forceinline__ P* eq8_pin(P* p) {            // global memory
  const unsigned long long a = (unsigned long long)p;
  const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)a, threadIdx.x & 31);
  const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(a >> 32), threadIdx.x & 31);
  P* r = (P*)(((unsigned long long)hi << 32) | lo);
  __builtin_assume(__isGlobal(r));
  return r;
}
__device__ __forceinline__ double* eq8_pin_smem(const double* p) {   // dynamic shared memory
  const int off = (int)(((unsigned)__cvta_generic_to_shared(p) - (unsigned)__cvta_generic_to_shared(eq8_dyn_smem)) >> 3);
  return eq8_dyn_smem + eq8_pin(off);
}

struct Eq8Smem {
  // CQ | BX | BY are one contiguous block per period -> Gpath
  double* CQ;   // [k][44]: packed lower columns of Cq (36) | g (8)
  double* BX;   // [8]      x_t (unmasked)
  double* BY;   // [4]      y_t
  double* CV;   // [k][8]   cvec
  double* XE;   // [k][8]   x_t, BVS-masked per equation
  double* EP;   // [k][8]   eps_t
  double* AN;   // [k][8]   a_{t+1} (backward pass)
  double* YS;   // [4]      y_t of the equations
  __device__ __forceinline__ Eq8Smem(double* xs, int k)
      : CQ(xs), BX(xs + 44 * k), BY(xs + 44 * k + 8), CV(xs + 44 * k + 12), XE(xs + 52 * k + 12), EP(xs + 60 * k + 12),
        AN(xs + 68 * k + 12), YS(xs + 76 * k + 12) {}
};

// all state normals of the sweep in one throughput-bound pass (site = t n + r, Box-Muller pairs): eps -> wpath
static __device__ __noinline__ void tvp_eq8_normals(const ChainRng& rng, int sweep, int nsite, double* wpath) {
  const int lane = threadIdx.x & 31, npair = (nsite + 1) >> 1;
  if (rng.injected(VB_STATE_N)) {
    for (int p = lane; p < nsite; p += 32) wpath[p] = rng.inj_at(VB_STATE_N, sweep, p);
  } else {
    int p = lane;
    for (; p + 32 < npair; p += 64) {   // two independent Philox / log / sincos chains per lane in flight
      const Normal4 nn = philox_normal_pair_x2(rng.seed, rng.chain, VB_STATE_N, sweep, p, rng.chain, VB_STATE_N, sweep, p + 32);
      wpath[2 * p] = nn.p.a; wpath[2 * p + 1] = nn.p.b;
      wpath[2 * (p + 32)] = nn.q.a;
      if (2 * (p + 32) + 1 < nsite) wpath[2 * (p + 32) + 1] = nn.q.b;
    }
    if (p < npair) {
      const Normal2 nn = rng.normal_pair(VB_STATE_N, sweep, p, nsite);
      wpath[2 * p] = nn.a;
      if (2 * p + 1 < nsite) wpath[2 * p + 1] = nn.b;
    }
  }
  __syncwarp();
}

// Forward recursion.  om: the lane's measurement precisions omega_{i,t} at om[oms * t] (shared memory).
template <int MB>
static __device__ __noinline__ bool tvp_eq8_forward(const double* __restrict__ X, const double* __restrict__ Y,
                                                    const int k_, const int T_, const int n_, double* xs, const double* q,
                                                    const double* a_init, const double* lam, const double* om_base,
                                                    const int oms_, double* Gpath, const double* wpath) {
  const int k = eq8_pin(k_), T = eq8_pin(T_), n = eq8_pin(n_), oms = eq8_pin(oms_);
  xs = eq8_pin_smem(xs);
  Gpath = eq8_pin(Gpath);
  const int lane = threadIdx.x & 31, gi = lane >> 3, c = lane & 7;
  const bool act = gi < k, own = act && c < MB;
  const int ie = act ? gi : 0;
  const Eq8Smem S(xs, k);
  double* cq = S.CQ + 44 * ie;
  double* cv = S.CV + 8 * ie;
  const double* xe = S.XE + 8 * ie;
  const double* ep = S.EP + 8 * ie;
  const int offc = BVG_EQ8_OFF(c);
  const int ST = 44 * k + 12;            // stream doubles per period
  const int rown = c * k + ie;           // the lane's own coefficient (regressor c, equation ie)
  BVG_TVP_TICK0

  // ---- per-lane constants: state precisions of the lane's equation, its own BVS mask
  double qv[MB];
#pragma unroll
  for (int j = 0; j < MB; ++j) qv[j] = q[j * k + ie];
  const double qc = (c < MB) ? q[rown] : 0.0;
  const double lamc = own ? ((lam != nullptr) ? lam[rown] : 1.0) : 0.0;
  // initial carry: Cq_{-1} = 0, cvec_{-1} = Q a_init  (bvartvpalg.cpp:296, first block of the right-hand side)
  for (int e = lane; e < 44 * k; e += 32) S.CQ[e] = 0.0;
  for (int e = lane; e < 8 * k; e += 32) { S.CV[e] = 0.0; S.AN[e] = 0.0; }
  __syncwarp();
  if (own) cv[c] = qc * a_init[rown];
  // x_0, eps_0, y_0 staged; the values of period 1 in flight.  The loads are consumed (stored to shared memory) a whole
  // period after they are issued; the pointers are laundered so that they stay in registers (as kernel parameters they
  // would be re-read from the constant bank in every period, a miss there is an L2 round trip).
  const double* Xp = X + c;
  const double* Yp = Y + ie;
  const double* Wp = wpath + rown;
  Xp = eq8_pin(Xp); Yp = eq8_pin(Yp); Wp = eq8_pin(Wp);
  double* YS = S.YS;
  const bool ylane = act && c == 0;
  if (act) {
    S.XE[lane] = own ? Xp[0] * lamc : 0.0;
    S.EP[lane] = own ? Wp[0] : 0.0;
  }
  if (ylane) YS[gi] = Yp[0];
  double pfx = 0.0, pfe = 0.0, pfy = 0.0;
  if (T > 1) {
    if (own) { pfx = Xp[MB]; pfe = Wp[n]; }
    if (ylane) pfy = Yp[k];
  }
  Xp += 2 * MB; Wp += 2 * (size_t)n; Yp += 2 * k;       // period t + 2 relative to the loop counter
  const double* om_p = eq8_pin_smem(om_base);
  for (int e = lane; e < 12; e += 32) S.BX[e] = 0.0;
  __syncwarp();

  bool bad = false;
  for (int t = 0; t < T; ++t) {
    // ---- phase A: read the carry and the data of the period, everything else in registers
    double x[MB], L[MB][MB], b[MB], rs[MB], xc[MB];
    const double om = *om_p;
    om_p += oms;
    const double yv = YS[ie];
    const double xown = S.XE[c];                    // x_{c,t} (equation 0's copy; unmasked whenever the block's copy is used)
    const double ct = (t < T - 1) ? 2.0 : 1.0;
#pragma unroll
    for (int j = 0; j < MB; ++j) x[j] = xe[j];
#pragma unroll
    for (int r = 0; r < MB; ++r) {
      const double xo = x[r] * om;
#pragma unroll
      for (int cc = 0; cc <= r; ++cc) L[r][cc] = fma(xo, x[cc], -cq[BVG_EQ8_OFF(cc) + r - cc]);
      L[r][r] = fma(ct, qv[r], L[r][r]);
      b[r] = fma(xo, yv, cv[r]);
      xc[r] = (r == c) ? 1.0 : 0.0;                 // right-hand side e_c of the column solve
    }
    BVG_TVP_TICK(0);
    // ---- Cholesky of D_t; carried along in axpy form (one multiply + one FMA per pivot on their own chains, off the
    //      pivot chain rsqrt -> multiply -> FMA):  w = L^-1 rhs (in b)  and  column c of M_t = L^-1 (in xc; zero above row c)
#pragma unroll
    for (int j = 0; j < MB; ++j) {
      const double d = L[j][j];
      bad = bad || !(d > 0.0);
      const double r0 = eq8_rsqrt(d);
      rs[j] = r0;
      b[j] *= r0;
      xc[j] *= r0;
#pragma unroll
      for (int i = j + 1; i < MB; ++i) L[i][j] *= r0;
#pragma unroll
      for (int i = j + 1; i < MB; ++i) {
        b[i] = fma(-L[i][j], b[j], b[i]);
        xc[i] = fma(-L[i][j], xc[j], xc[i]);
      }
#pragma unroll
      for (int cc = j + 1; cc < MB; ++cc)
#pragma unroll
        for (int i = cc; i < MB; ++i) L[i][cc] = fma(-L[i][j], L[cc][j], L[i][cc]);
    }
    BVG_TVP_TICK(1);
    // ---- g_c = (M' z)_c and (M' w)_c with z = w + eps ...
    double g0 = 0.0, g1 = 0.0, u0 = 0.0, u1 = 0.0;
#pragma unroll
    for (int r = 0; r < MB; ++r) {
      const double zr = b[r] + ep[r];
      if (r & 1) { g1 = fma(xc[r], zr, g1); u1 = fma(xc[r], b[r], u1); }
      else { g0 = fma(xc[r], zr, g0); u0 = fma(xc[r], b[r], u0); }
    }
    // ... and column c of D_t^-1 = L^-T (column c of L^-1): column-oriented back substitution, in place
#pragma unroll
    for (int r = MB - 1; r >= 0; --r) {
      xc[r] *= rs[r];
#pragma unroll
      for (int mm = 0; mm < r; ++mm) xc[mm] = fma(-L[r][mm], xc[r], xc[mm]);
    }
    BVG_TVP_TICK(2);
    __syncwarp();                                   // every lane has read the carry of this period
    // ---- phase B: publish Cq_t (rows >= c of column c), cvec_t, g_t and the inputs of the next period
    if (own) {
#pragma unroll
      for (int r = 0; r < MB; ++r)
        if (r >= c) cq[offc + r - c] = (qc * qv[r]) * xc[r];
      cv[c] = qc * (u0 + u1);
      cq[36 + c] = g0 + g1;
    }
    if (lane < 8) S.BX[lane] = xown;                // x_t, y_t travel with the block: the backward pass forms the residuals
    if (ylane) S.BY[gi] = yv;
    if (act) { S.XE[lane] = pfx * lamc; S.EP[lane] = pfe; }
    if (ylane) YS[gi] = pfy;
    __syncwarp();
    // ---- phase C: stream [Cq_t | g_t] to HBM (coalesced 16-byte stores); fetch the inputs of period t + 2
    {
      double2* dst = reinterpret_cast<double2*>(Gpath + (size_t)t * ST);
      const double2* src = reinterpret_cast<const double2*>(S.CQ);
      for (int u = lane; u < 22 * k + 6; u += 32) dst[u] = src[u];
    }
    if (t + 2 < T) {
      if (own) { pfx = *Xp; pfe = *Wp; }
      if (ylane) pfy = *Yp;
    }
    Xp += MB; Wp += n; Yp += k;
    BVG_TVP_TICK(3);
  }
  return bad;
}

// Backward recursion:  a_{T-1} = g_{T-1};  a_t = g_t + (1 / q) (Cq_t a_{t+1}):  lane c gathers row c of the symmetric Cq_t
// from the packed columns.  A period is ~100 cycles of work, an L2 / HBM round trip ~1000: the blocks come back through a
// ring of D periods in shared memory filled by cp.async (LDGSTS, 16 bytes per lane and word) -- the loop stays rolled (an
// unrolled register ring of the same depth is 30 KB of code per instantiation, and the instruction cache, not the data
// path, became the limit).  `ring`: 16-byte aligned shared memory free during the pass (the SV work arrays), `ring_len`
// doubles; D = 8 or 4 periods as it fits, otherwise the blocks are loaded synchronously.  ss_out[r] (shared, n) receives
// sum_t (a_t - a_{t-1})^2 with a_{-1} = a_init, the input of the state-variance draw (bvartvpalg.cpp:471-477).
// FUSED (models without BVS): a_t goes straight into the retained-draw buffer when coef_out != nullptr ([T][n],
// bvartvpalg.cpp:521-524) and the residuals u_t = y_t - Z_t a_t (:332-336) are formed from the x_t / y_t that travel
// with the block; the scratch copy of the path (apath) is not written then.
__device__ __forceinline__ void eq8_cp16(double* dst_shared, const double* src_global) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_shared)), "l"(src_global) : "memory");
}
__device__ __forceinline__ void eq8_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void eq8_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int MB, bool FUSED>
static __device__ __noinline__ void tvp_eq8_backward(const int k_, const int T_, const int n_, double* xs, const double* q,
                                                     const double* a_init, const double* Gpath, double* apath,
                                                     double* coef_out, double* U, double* ss_out, double* ring,
                                                     int ring_len) {
  const int k = eq8_pin(k_), T = eq8_pin(T_), n = eq8_pin(n_);
  xs = eq8_pin_smem(xs);
  Gpath = eq8_pin(Gpath);
  if (ring != nullptr) ring = eq8_pin_smem(ring);
  if (FUSED) { U = eq8_pin_smem(U); if (coef_out != nullptr) coef_out = eq8_pin(coef_out); }
  else apath = eq8_pin(apath);
  const int lane = threadIdx.x & 31, gi = lane >> 3, c = lane & 7;
  const bool act = gi < k, own = act && c < MB;
  const int ie = act ? gi : 0;
  const Eq8Smem S(xs, k);
  double* an = S.AN + 8 * ie;
  const int offc = BVG_EQ8_OFF(c);
  const int ST = 44 * k + 12;
  const int rown = c * k + ie;
  BVG_TVP_TICK0
  int gidx[MB];
#pragma unroll
  for (int j = 0; j < MB; ++j) gidx[j] = 44 * ie + ((j <= c) ? (BVG_EQ8_OFF(j) + c - j) : (offc + j - c));
  const int gslot = 44 * ie + 36 + (own ? c : 0);   // g_c
  const int xslot = 44 * k + (own ? c : 0), yslot = 44 * k + 8 + ie;
  const double invq = own ? 1.0 / q[rown] : 0.0;
  const double a0c = own ? a_init[rown] : 0.0;
  const bool ylane = act && c == 0;
  const int nw = 22 * k + 6;                        // 16-byte words per period (<= 94: three per lane)
  double aprev = 0.0, ss = 0.0;
  // one period from the block at blk_ (shared memory)
#define BVG_EQ8_BACK_STEP(t_, blk_)                                                                   \
  do {                                                                                                \
    double a_ = (blk_)[gslot];                                                                        \
    if ((t_) < T - 1) {                                                                               \
      double s0_ = 0.0, s1_ = 0.0;                                                                    \
      _Pragma("unroll") for (int j = 0; j < MB; ++j) {                                                \
        if (j & 1) s1_ = fma((blk_)[gidx[j]], an[j], s1_); else s0_ = fma((blk_)[gidx[j]], an[j], s0_); \
      }                                                                                               \
      a_ = fma(s0_ + s1_, invq, a_);                                                                  \
      const double dd_ = aprev - a_;                                                                  \
      ss = fma(dd_, dd_, ss);                                                                         \
    }                                                                                                 \
    aprev = a_;                                                                                       \
    if (FUSED) {                                /* residual of the group's equation */               \
      double pr_ = own ? (blk_)[xslot] * a_ : 0.0;                                                    \
      const double yr_ = (blk_)[yslot];                                                               \
      pr_ += __shfl_xor_sync(0xffffffffu, pr_, 1);                                                    \
      pr_ += __shfl_xor_sync(0xffffffffu, pr_, 2);                                                    \
      pr_ += __shfl_xor_sync(0xffffffffu, pr_, 4);                                                    \
      if (ylane) U[gi + k * (t_)] = yr_ - pr_;                                                        \
    }                                                                                                 \
    __syncwarp();                               /* all reads of a_{t+1} and of the block done */      \
    if (own) {                                                                                        \
      an[c] = a_;                                                                                     \
      if (!FUSED) apath[(size_t)(t_) * n + rown] = a_;                                                \
      else if (coef_out != nullptr) coef_out[(size_t)(t_) * n + rown] = a_;                           \
    }                                                                                                 \
  } while (0)
  // the lane's words of period p into slot s_ (an empty group is committed for p < 0 so that the group count stays in step)
#define BVG_EQ8_FETCH(p_, s_)                                                                         \
  do {                                                                                                \
    if ((p_) >= 0) {                                                                                  \
      const double* src_ = Gpath + (size_t)(p_) * ST;                                                 \
      double* dst_ = ring + (s_) * ST;                                                                \
      for (int w_ = lane; w_ < nw; w_ += 32) eq8_cp16(dst_ + 2 * w_, src_ + 2 * w_);                  \
    }                                                                                                 \
    eq8_cp_commit();                                                                                  \
  } while (0)
#define BVG_EQ8_BACK_LOOP(D_)                                                                         \
  do {                                                                                                \
    for (int u_ = 0; u_ < (D_); ++u_) BVG_EQ8_FETCH(T - 1 - u_, u_);                                  \
    int s_ = 0;                                                                                       \
    for (int t = T - 1; t >= 0; --t) {                                                                \
      eq8_cp_wait<(D_) - 1>();                  /* the oldest group in flight (period t) has landed */ \
      __syncwarp();                             /* ... for every lane; a_{t+1} visible */             \
      const double* blk = ring + s_ * ST;                                                             \
      BVG_EQ8_BACK_STEP(t, blk);                                                                      \
      BVG_EQ8_FETCH(t - (D_), s_);              /* the slot is free again (barrier inside the step) */ \
      s_ = (s_ + 1 == (D_)) ? 0 : s_ + 1;                                                             \
    }                                                                                                 \
  } while (0)
  if (ring != nullptr && ring_len >= 8 * ST) BVG_EQ8_BACK_LOOP(8);
  else if (ring != nullptr && ring_len >= 4 * ST) BVG_EQ8_BACK_LOOP(4);
  else {                                            // synchronous: stage the block in the (idle) carry area
    double2* stage = reinterpret_cast<double2*>(S.CQ);
    for (int t = T - 1; t >= 0; --t) {
      const double2* src = reinterpret_cast<const double2*>(Gpath + (size_t)t * ST);
      for (int w = lane; w < nw; w += 32) stage[w] = src[w];
      __syncwarp();
      const double* blk = S.CQ;
      BVG_EQ8_BACK_STEP(t, blk);
    }
  }
#undef BVG_EQ8_BACK_LOOP
#undef BVG_EQ8_FETCH
#undef BVG_EQ8_BACK_STEP
  if (own) {
    const double dd = aprev - a0c;                  // a_0 - a_init
    ss_out[rown] = fma(dd, dd, ss);
  }
  __syncwarp();
  BVG_TVP_TICK(6);
}

template <int MB>
static __device__ __forceinline__ void tvp_eq8_mb(const BvarModel& m, bool sv, double* xs, const double* q,
                                                  const double* a_init, const double* lam, const double* sig,
                                                  const double* ew, double* Gpath, double* wpath, double* apath,
                                                  double* coef_out, double* U, double* ss_out, double* ring, int ring_len,
                                                  int& fail) {
  const int k = m.k, T = m.T, lane = threadIdx.x & 31;
  const int ie = ((lane >> 3) < k) ? (lane >> 3) : 0;
  const bool bad = tvp_eq8_forward<MB>(m.X, m.Y, k, T, m.n, xs, q, a_init, lam, sv ? ew + (size_t)T * ie : sig + ie + k * ie,
                                       sv ? 1 : 0, Gpath, wpath);
  if (__any_sync(0xffffffffu, bad)) fail = 1;
  __syncwarp();                                     // the last periods' stores are ordered before the loads of the backward pass
  if (U != nullptr) tvp_eq8_backward<MB, true>(k, T, m.n, xs, q, a_init, Gpath, apath, coef_out, U, ss_out, ring, ring_len);
  else tvp_eq8_backward<MB, false>(k, T, m.n, xs, q, a_init, Gpath, apath, coef_out, U, ss_out, ring, ring_len);
}

// state normals, then dispatch on the regressors per equation
static __device__ __forceinline__ void tvp_state_draw_eq8(const BvarModel& m, const ChainRng& rng, int sweep, bool sv,
                                                          double* xs, const double* q, const double* a_init,
                                                          const double* lam, const double* sig, const double* ew,
                                                          double* Gpath, double* wpath, double* apath, double* coef_out,
                                                          double* U, double* ss_out, double* ring, int ring_len, int& fail) {
  BVG_TVP_TICK0
  tvp_eq8_normals(rng, sweep, m.n * m.T, wpath);
  BVG_TVP_TICK(7);
#define BVG_EQ8_ARGS m, sv, xs, q, a_init, lam, sig, ew, Gpath, wpath, apath, coef_out, U, ss_out, ring, ring_len, fail
  switch (m.M) {
    case 1: tvp_eq8_mb<1>(BVG_EQ8_ARGS); break;
    case 2: tvp_eq8_mb<2>(BVG_EQ8_ARGS); break;
    case 3: tvp_eq8_mb<3>(BVG_EQ8_ARGS); break;
    case 4: tvp_eq8_mb<4>(BVG_EQ8_ARGS); break;
    case 5: tvp_eq8_mb<5>(BVG_EQ8_ARGS); break;
    case 6: tvp_eq8_mb<6>(BVG_EQ8_ARGS); break;
    case 7: tvp_eq8_mb<7>(BVG_EQ8_ARGS); break;
    default: tvp_eq8_mb<8>(BVG_EQ8_ARGS); break;
  }
#undef BVG_EQ8_ARGS
}
#endif

}  // namespace bvg
