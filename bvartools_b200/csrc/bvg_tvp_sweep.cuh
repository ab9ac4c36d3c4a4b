// bvg_tvp_sweep.cuh -- one chain of the TVP-BVAR Gibbs sampler (src/bvartvpalg.cpp:282-553).
//
// The reference assembles the (nT) x (nT) precision  P = H'(I_T (x) Sv^-1)H + zz' Su^-1 zz  (bvartvpalg.cpp:293-295),
// converts it to a DENSE matrix and runs LU solve + Cholesky on it (:296-297).  P is block tridiagonal with
// n x n blocks:  diag  B_t = c_t Q + Z_t' S_t^-1 Z_t  (c_t = 2, last 1; Q = diag(Sv^-1)),  off-diagonal -Q.
// Its (unique) Cholesky factor is block bidiagonal, so the reference's  a = P^-1 rhs + chol(P)^-1 eps  is
// reproduced exactly by a block forward recursion (per t: one n x n Cholesky, one triangular inverse M_t = L_tt^-1,
// one M'M product) and a backward recursion made only of triangular mat-vecs:
//     D_t   = B_t - Q M_{t-1}' M_{t-1} Q                      L_tt L_tt' = D_t
//     w_t   = L_tt^-1 (rhs_t + Q M_{t-1}' w_{t-1})
//     a_t   = M_t' ( w_t + eps_t + M_t Q a_{t+1} )
// M_t (packed lower) is streamed through HBM: T tri(n) doubles written forward, read backward.
#pragma once
#include "bvg_bvar_sweep.cuh"
#include "bvg_tvp_tiled.cuh"
#include "bvg_tvp_eq8.cuh"

namespace bvg {

#if defined(BVG_TVP_PROF) && defined(__CUDA_ARCH__)
#define BVG_TVPC_TICK0 long long tick_ = clock64();
#define BVG_TVPC_TICK(slot) BVG_TVP_TICK(slot)
#else
#define BVG_TVPC_TICK0
#define BVG_TVPC_TICK(slot) do { } while (0)
#endif

#define BVG_TVP_PSI_MAXK 8      // the psi block of TVP models keeps k x k matrices in registers / local memory

struct TvpModel {
  BvarModel b;                   // data, error-variance prior, BVS prior; mu/vdiag0/vi_off = prior of the initial state
  const double* v_shape_post;    // n  shape + T/2                                   (bvartvpalg.cpp:104)
  const double* v_rate;          // n                                                (bvartvpalg.cpp:105)
  // covariance (psi) block (bvartvpalg.cpp:143-176): b.n_psi, b.psi_mu, b.psi_vi (full matrix) + state variances
  const double* psi_shape_post;  // n_psi  shape + T/2                               (bvartvpalg.cpp:149)
  const double* psi_rate;        // n_psi                                            (bvartvpalg.cpp:150)
};

struct TvpOut {
  double* coef;      // [chain][iter][T][n]
  double* sigma;     // [chain][iter][k*k] or [chain][iter][T][k*k]
  double* lambda;    // [chain][iter][n] or nullptr
  double* sigma_h;   // [chain][iter][k] or nullptr
  double* sigma_v;   // [chain][iter][n]
  int* status;
  long long iters;
};

// state: [0,kk) Sigma_u^-1 | n q = diag(Sigma_v^-1) | n a_init | n lambda | SV: T*k h, k sigma_h, k h_init
//        | covar: n_psi psi state precisions, n_psi psi_init, n_psi psi lambda, T*kk Sigma_t^-1 = Psi_t' Omega_t^-1 Psi_t
BVG_HD int tvp_state_len(int k, int T, int n, int sigma_type, int n_psi = 0) {
  return k * k + 3 * n + (sigma_type == SIG_SV ? T * k + 2 * k : 0) + (n_psi > 0 ? 3 * n_psi + T * k * k : 0);
}
// global scratch per chain: T*tri(n) (M_t) + T*n (w_t + eps_t) + T*n (a path)
// (the warp-per-chain tile path streams full 8 x 8 tiles: tiled_len(n) >= tri(n) doubles per period)
// (models with the psi block run the generic recursion: their Sigma_t^-1 is a full matrix per period)
// (diagonal Sigma_t^-1, warp per chain: the equation-decoupled path streams 44 k doubles per period, bvg_tvp_eq8.cuh)
enum TvpMode : int { TVP_GENERIC = 0, TVP_TILED = 1, TVP_EQ8 = 2 };
BVG_HD bool tvp_tiled_ok(int n, int n_psi = 0) { return n <= 32 && n_psi == 0; }
// the path a chain owned by `lanes` lanes takes (lanes = 32: one warp per chain)
BVG_HD int tvp_mode(int lanes, int k, int M, int n, int n_psi, int diag_sigma) {
  if (lanes != 32) return TVP_GENERIC;
  if (tvp_eq8_ok(k, M, n_psi, diag_sigma)) return TVP_EQ8;
  return tvp_tiled_ok(n, n_psi) ? TVP_TILED : TVP_GENERIC;
}
BVG_HD int tvp_stream_len(int mode, int k, int n) {
  return mode == TVP_EQ8 ? tvp_eq8_stream(k) : (mode == TVP_TILED ? tiled_len(n) : tri(n));
}
// large enough for whichever path the launch plan picks (even, so that chains stay 16-byte aligned)
BVG_HD long long tvp_scratch_len(int T, int k, int M, int n, int n_psi, int diag_sigma) {
  int st = tri(n);
  if (tvp_tiled_ok(n, n_psi) && tiled_len(n) > st) st = tiled_len(n);
  if (tvp_eq8_ok(k, M, n_psi, diag_sigma)) st = (tvp_eq8_stream(k) > tri(n)) ? tvp_eq8_stream(k) : tri(n);
  const long long len = (long long)T * (st + 2 * n);
  return len + (len & 1);
}

BVG_HD int tvp_smem_len(int k, int T, int n, int sigma_type, int mode = TVP_GENERIC, int n_psi = 0) {
  int len = 6 * k * k + 2 * k + k * T;
  if (mode == TVP_TILED) len += 2 * tiled_len(n) + 10 * (n + 8);
  else if (mode == TVP_EQ8) len += tvp_eq8_smem(k) + tri(n) + 9 * n;
  else len += 3 * tri(n) + 9 * n;
  if (sigma_type == SIG_SV) len += 3 * k * T + 2 * k + 20;                   // h, ew, work, sigma_h, h_init, mixture look-up
  if (n_psi > 0) len += T * k * k + T * n_psi + 4 * n_psi;
  return len + (len & 1);
}

template <class Grp>
BVG_HD void tvp_residuals(const Grp& g, const BvarModel& m, const double* apath, const double* lam, bool mask,
                          double* U) {
  const int k = m.k, T = m.T, M = m.M, n = m.n;
  for (int e = g.tid(); e < k * T; e += g.size()) {
    const int i = e % k, t = e / k;
    const double* at = apath + (size_t)t * n;
    double s = m.Y[e];
    for (int j = 0; j < M; ++j) {
      const double av = mask ? lam[j * k + i] * at[j * k + i] : at[j * k + i];
      s = fma(-av, m.X[j + M * t], s);
    }
    U[e] = s;
  }
  g.sync();
}

// One equation with time-varying coefficients:  y_t = x_t' b_t + e_t,  e_t ~ N(0, 1 / w_t),  b_t = b_{t-1} + v_t,
// v_t ~ N(0, diag(1 / q)).  Row i of the covariance draw of bvartvpalg.cpp:339-363 is exactly this with
// x_t = -u_{0..i-1,t}, y_t = u_{i,t}, w_t = omega_{i,t}^-1: psi_z only couples the i elements of row i and the state
// variances are diagonal, so the (n_psi T)^2 precision of :361 decouples into k - 1 independent block-tridiagonal
// systems whose Cholesky factors are the matching sub-blocks of the joint factor -- the per-row draws with the matching
// entries of the normal vector ARE the joint draw.  Same recursion as the coefficient draw (see the file header).
// X[xs * t + c] (times xsign, times lam[c] when BVS masks the regressors), Y[ys * t], w[ws * t]; normals from site VB_PSI_N, index t * site_stride + site_off + r;
// draws to out[out_stride * t + r].
template <class Grp>
BVG_HD void tvp_single_eq_draw(const Grp& g, const ChainRng& rng, int sweep, int nn, int T, const double* X, int xs,
                               double xsign, const double* Y, int ys, const double* w, int ws, const double* lam,
                               const double* q, const double* b_init, int site_stride, int site_off, double* Dt, double* Mi, double* Cq,
                               double* wv, double* cvec, double* dinv, double* anext, double* tmp, double* tmp2,
                               double* Mpath, double* wpath, double* out, int out_stride, int& fail) {
  const int G = g.size(), tid = g.tid(), tn = tri(nn);
  for (int t = 0; t < T; ++t) {
    const double* xt = X + (size_t)xs * t;
    const double wt = w[(size_t)ws * t], yt = Y[(size_t)ys * t];
    const double ct = (t < T - 1) ? 2.0 : 1.0;
    for (int r = tid; r < nn; r += G) {
      double* Dr = Dt + tri(r);
      const double xr = xsign * xt[r] * (lam ? lam[r] : 1.0);                  // psi_z = psi_z_bvs Lambda (:354-357)
      for (int c = 0; c <= r; ++c) {
        double gv = xr * (xsign * xt[c] * (lam ? lam[c] : 1.0)) * wt;
        if (c == r) gv += ct * q[r];
        if (t > 0) gv -= Cq[tri(r) + c];
        Dr[c] = gv;
      }
      double b = xr * (wt * yt);
      b += (t == 0) ? q[r] * b_init[r] : cvec[r];
      wv[r] = b;
    }
    g.sync();
    chol_crout(g, Dt, dinv, wv, nn, fail);
    tri_inverse(g, Dt, dinv, Mi, nn);
    double* Mg = Mpath + (size_t)t * tn;
    for (int e = tid; e < tn; e += G) Mg[e] = Mi[e];
    for (int r = tid; r < nn; r += G)
      wpath[(size_t)t * nn + r] = wv[r] + rng.normal(VB_PSI_N, sweep, t * site_stride + site_off + r);
    if (t < T - 1) {
      mtm_lower(g, Mi, Cq, nn);
      int i = 0;
      for (int e = tid; e < tn; e += G) {
        while (tri(i + 1) <= e) ++i;
        Cq[e] *= q[i] * q[e - tri(i)];
      }
      for (int r = tid; r < nn; r += G) {
        double s = 0.0;
        for (int mm = r; mm < nn; ++mm) s = fma(Mi[tri(mm) + r], wv[mm], s);
        cvec[r] = q[r] * s;
      }
    }
    g.sync();
  }
  for (int t = T - 1; t >= 0; --t) {
    const double* Mg = Mpath + (size_t)t * tn;
    for (int e = tid; e < tn; e += G) Mi[e] = Mg[e];
    for (int r = tid; r < nn; r += G) tmp[r] = wpath[(size_t)t * nn + r];
    g.sync();
    for (int r = tid; r < nn; r += G) {
      double s = 0.0;
      if (t < T - 1) {
        const double* Mr = Mi + tri(r);
        for (int c = 0; c <= r; ++c) s = fma(Mr[c], q[c] * anext[c], s);
      }
      tmp2[r] = tmp[r] + s;
    }
    g.sync();
    for (int r = tid; r < nn; r += G) {
      double s = 0.0;
      for (int mm = r; mm < nn; ++mm) s = fma(Mi[tri(mm) + r], tmp2[mm], s);
      out[(size_t)out_stride * t + r] = s;
      tmp[r] = s;
    }
    g.sync();
    for (int r = tid; r < nn; r += G) anext[r] = tmp[r];
    g.sync();
  }
}

template <class Grp>
BVG_HD void tvp_chain(const Grp& g, const TvpModel& tm, const ChainRng& rng, double* state, double* scratch,
                      double* sm, const TvpOut& out, long long chain_local, int sweep_begin, int sweep_end,
                      int burnin) {
  const BvarModel& m = tm.b;
  const int G = g.size(), tid = g.tid();
  const int k = m.k, T = m.T, M = m.M, n = m.n, kk = k * k, tn = tri(n);
  const bool sv = m.sigma_type == SIG_SV;
  const bool bvs = m.varsel == VS_BVS;
  const int n_psi = m.n_psi;
  const bool covar = n_psi > 0;
#if defined(__CUDA_ARCH__)
  const int mode = tvp_mode(Grp::kWarp32 ? 32 : 0, k, M, n, n_psi, m.diag_sigma);
#else
  const int mode = TVP_GENERIC;
#endif
  const bool tiled = mode == TVP_TILED;                 // warp per chain, n <= 32: DMMA tile path (bvg_tvp_tiled.cuh)
  const bool eq8 = mode == TVP_EQ8;                     // warp per chain, diagonal Sigma_t^-1: bvg_tvp_eq8.cuh
  const bool fused = eq8 && !bvs && m.a0_from < 0;      // its backward pass also stores the path and forms the residuals
  const int nv = tiled ? n + 8 : n;                     // vectors are zero padded to whole tiles
  double* xs = sm;      sm += eq8 ? tvp_eq8_smem(k) : 0;   // first: 16-byte aligned exchange area of the eq8 path
  double* Dt = sm;      sm += tiled ? tiled_len(n) : tn;
  double* Mi = sm;      sm += (tiled || eq8) ? 0 : tn;
  double* Cq = sm;      sm += tiled ? tiled_len(n) : (eq8 ? 0 : tn);
  double* wv = sm;      sm += nv;     // rhs_t -> w_t
  double* cvec = sm;    sm += nv;     // Q M_t' w_t (carry into t+1)
  double* dinv = sm;    sm += nv;
  double* q = sm;       sm += nv;
  double* a_init = sm;  sm += nv;
  double* lam = sm;     sm += nv;
  double* anext = sm;   sm += nv;
  double* tmp = sm;     sm += nv;
  double* tmp2 = sm;    sm += nv;
  double* xl = sm;      sm += tiled ? nv : 0;
  double* sig = sm;     sm += kk;
  double* S = sm;       sm += kk;
  double* W1 = sm;      sm += kk;
  double* W2 = sm;      sm += kk;
  double* W3 = sm;      sm += kk;
  double* W4 = sm;      sm += kk;
  double* kd = sm;      sm += k;
  double* kv = sm;      sm += k;
  double* U = sm;       sm += k * T;
  double* h = nullptr; double* ew = nullptr; double* svw = nullptr; double* sigma_h = nullptr; double* h_init = nullptr;
  double* svtab = nullptr;
  if (sv) { h = sm; sm += k * T; ew = sm; sm += k * T; svw = sm; sm += k * T; sigma_h = sm; sm += k; h_init = sm; sm += k; svtab = sm; sm += 20; }
  double* sigt = nullptr; double* psip = nullptr; double* pq = nullptr; double* pinit = nullptr;
  double* plam = nullptr; double* pnew = nullptr;
  const bool psi_bvs = covar && m.psi_varsel == VS_BVS;
  if (covar) {
    sigt = sm; sm += T * kk; psip = sm; sm += T * n_psi; pq = sm; sm += n_psi; pinit = sm; sm += n_psi;
    plam = sm; sm += n_psi; pnew = sm; sm += n_psi;
  }
  const int lam_rows = (bvs ? n : 0) + (psi_bvs ? n_psi : 0);
  double* Mpath = scratch;                         // [T][tn]  (tiled: [T][tiled_len(n)], eq8: [T][44 k])
  double* wpath = scratch + (size_t)T * tvp_stream_len(mode, k, n);   // [T][n]
  double* apath = wpath + (size_t)T * n;           // [T][n]
  int fail = 0;

  double* st_q = state + kk;
  double* st_ai = st_q + n;
  double* st_lam = st_ai + n;
  double* st_h = st_lam + n;
  double* st_psi = st_h + (sv ? T * k + 2 * k : 0);     // n_psi pq | n_psi psi_init | T*kk Sigma_t^-1
  if (covar) {
    for (int e = tid; e < n_psi; e += G) { pq[e] = st_psi[e]; pinit[e] = st_psi[n_psi + e]; plam[e] = st_psi[2 * n_psi + e]; }
    for (int e = tid; e < T * kk; e += G) sigt[e] = st_psi[3 * n_psi + e];
  }
  for (int e = tid; e < kk; e += G) sig[e] = state[e];
  for (int e = tid; e < n; e += G) { q[e] = st_q[e]; a_init[e] = st_ai[e]; lam[e] = st_lam[e]; }
  for (int e = n + tid; e < nv; e += G) { q[e] = 0.0; a_init[e] = 0.0; anext[e] = 0.0; }
  if (sv) {
    for (int e = tid; e < T * k; e += G) h[e] = st_h[e];
    for (int e = tid; e < k; e += G) { sigma_h[e] = st_h[T * k + e]; h_init[e] = st_h[T * k + k + e]; }
  }
  g.sync();

  for (int sweep = sweep_begin; sweep < sweep_end; ++sweep) {
    const bool keep = sweep >= burnin;
    const long long pos = (long long)sweep - burnin;
    if (sv && (sweep == sweep_begin || !Grp::kWarp32 || covar)) {
      // bvartvpalg.cpp:446; the very first sweep uses h.row(0) for every period (:224-225,230)
      // (warp per chain without the psi block: later sweeps take exp(-h) from the end of the previous sweep's SV block,
      //  one exp pass per sweep; with the psi block that pass does not run and ew is a work array of the SV step)
      for (int e = tid; e < k * T; e += G) ew[e] = 1.0 / exp(sweep == 0 ? h[T * (e / T)] : h[e]);
      g.sync();
    }
    if (covar && sweep == 0) {
      // before the first psi draw Sigma_t^-1 is the diagonal initial value (bvartvpalg.cpp:224-233)
      for (int e = tid; e < T * kk; e += G) {
        const int t = e / kk, f = e % kk, i = f % k, j = f / k;
        sigt[e] = (i == j) ? (sv ? ew[t + T * i] : sig[i + k * i]) : 0.0;
      }
      g.sync();
    }
    const bool svd = sv && !covar;                      // diagonal Sigma_t^-1 read from ew; covar: full matrix per period
#if defined(__CUDA_ARCH__)
    if (eq8) {
      // without BVS (and without the row selection of structural models) the backward pass stores the retained path and
      // forms the residuals itself; the scratch copy of the path is not needed then
      double* oc = (fused && keep) ? out.coef + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)n * (size_t)T : nullptr;
      if constexpr (Grp::kWarp32)
      {
        // backward-stream ring: the SV work arrays ew | svw are idle during the backward pass (ew is consumed by the forward
        // pass and refreshed at the end of the SV block), 16-byte aligned inside them
        double* ring = nullptr;
        int ring_len = 0;
        if (sv) {                                       // (the BVS step still reads ew: only svw is free then)
          double* r0 = bvs ? svw : ew;
          ring = r0 + (((unsigned long long)r0 >> 3) & 1);
          ring_len = (bvs ? 1 : 2) * k * T - (int)(ring - r0);
        }
        tvp_state_draw_eq8(m, rng, sweep, sv, xs, q, a_init, bvs ? lam : nullptr, sig, ew, Mpath, wpath,
                           fused ? nullptr : apath, oc, fused ? U : nullptr, tmp2, ring, ring_len, fail);
      }
    } else if (tiled) {
      if constexpr (Grp::kWarp32)
        tvp_state_draw_tiled(m, rng, sweep, sv, Dt, Cq, wv, cvec, xl, anext, tmp, q, a_init, bvs ? lam : nullptr, sig,
                             ew, kv, Mpath, wpath, apath, fail);
    } else
#endif
    {
    // ================= forward recursion =================
    for (int t = 0; t < T; ++t) {
      const double* xt = m.X + (size_t)M * t;
      const double* yt = m.Y + (size_t)k * t;
      const double* sgt = covar ? sigt + (size_t)kk * t : sig;
      // kv = Sigma_t^-1 y_t
      for (int i = tid; i < k; i += G) {
        double s;
        if (svd) s = ew[t + T * i] * yt[i];
        else { s = 0.0; for (int i2 = 0; i2 < k; ++i2) s = fma(sgt[i + k * i2], yt[i2], s); }
        kv[i] = s;
      }
      g.sync();
      const double ct = (t < T - 1) ? 2.0 : 1.0;
      for (int r = tid; r < n; r += G) {
        const int j = r / k, i = r % k;
        double* Dr = Dt + tri(r);
        const double lr = bvs ? lam[r] : 1.0;
        const double xr = xt[j] * lr;
        int j2 = 0, i2 = 0;
        for (int c = 0; c <= r; ++c) {
          double sv_ii;
          if (svd) sv_ii = (i == i2) ? ew[t + T * i] : 0.0;
          else sv_ii = sgt[i + k * i2];
          double gv = xr * (xt[j2] * (bvs ? lam[c] : 1.0)) * sv_ii;
          if (c == r) gv += ct * q[r];
          if (t > 0) gv -= Cq[tri(r) + c];
          Dr[c] = gv;
          if (++i2 == k) { i2 = 0; ++j2; }
        }
        double b = xr * kv[i];
        if (t == 0) b += q[r] * a_init[r];                                              // :296 first block
        else b += cvec[r];
        wv[r] = b;
      }
      g.sync();
      chol_crout(g, Dt, dinv, wv, n, fail);
      tri_inverse(g, Dt, dinv, Mi, n);
      // stream M_t and w_t + eps_t to HBM
      double* Mg = Mpath + (size_t)t * tn;
      for (int e = tid; e < tn; e += G) Mg[e] = Mi[e];
      for (int r = tid; r < n; r += G) wpath[(size_t)t * n + r] = wv[r] + rng.normal(VB_STATE_N, sweep, t * n + r);
      if (t < T - 1) {
        // carry: Cq = Q M'M Q,  cvec = Q M' w
        mtm_lower(g, Mi, Cq, n);
        int i = 0;
        for (int e = tid; e < tn; e += G) {
          while (tri(i + 1) <= e) ++i;
          Cq[e] *= q[i] * q[e - tri(i)];
        }
        for (int r = tid; r < n; r += G) {
          double s = 0.0;
          for (int mm = r; mm < n; ++mm) s = fma(Mi[tri(mm) + r], wv[mm], s);
          cvec[r] = q[r] * s;
        }
      }
      g.sync();
    }
    // ================= backward recursion =================
    for (int t = T - 1; t >= 0; --t) {
      const double* Mg = Mpath + (size_t)t * tn;
      for (int e = tid; e < tn; e += G) Mi[e] = Mg[e];
      for (int r = tid; r < n; r += G) tmp[r] = wpath[(size_t)t * n + r];
      g.sync();
      if (t < T - 1) {
        for (int r = tid; r < n; r += G) {
          double s = 0.0;
          const double* Mr = Mi + tri(r);
          for (int c = 0; c <= r; ++c) s = fma(Mr[c], q[c] * anext[c], s);
          tmp2[r] = tmp[r] + s;
        }
      } else {
        for (int r = tid; r < n; r += G) tmp2[r] = tmp[r];
      }
      g.sync();
      for (int r = tid; r < n; r += G) {
        double s = 0.0;
        for (int mm = r; mm < n; ++mm) s = fma(Mi[tri(mm) + r], tmp2[mm], s);
        apath[(size_t)t * n + r] = s;
        tmp[r] = s;
      }
      g.sync();
      for (int r = tid; r < n; r += G) anext[r] = tmp[r];
      g.sync();
    }
    }

    // ================= BVS (bvartvpalg.cpp:300-330) and residuals (:332-336) =================
    BVG_TVPC_TICK0
    if (!fused) tvp_residuals(g, m, apath, lam, bvs, U);
    BVG_TVPC_TICK(8);
    if (bvs) {
      for (int r = tid; r < n; r += G) {
        double newlam = lam[r];
        if (m.include[r]) {
          const int j = r / k, i = r % k;
          const double lp0 = log(1.0 - m.inprior[r]), lp1 = log(m.inprior[r]);
          const double g1 = log(rng.uniform(VB_BVS_U1, sweep, r));
          const bool gate = (lam[r] == 1.0) ? (g1 < lp1) : (g1 < lp0);
          if (gate) {
            double dq = 0.0;
            for (int t = 0; t < T; ++t) {
              const double x = m.X[j + M * t];
              double v, sii;
              if (svd) { sii = ew[t + T * i]; v = sii * U[i + k * t]; }
              else {
                const double* sgt = covar ? sigt + (size_t)kk * t : sig;
                sii = sgt[i + k * i];
                v = 0.0;
                for (int i2 = 0; i2 < k; ++i2) v = fma(sgt[i + k * i2], U[i2 + k * t], v);
              }
              const double ar = apath[(size_t)t * n + r];
              const double d0 = lam[r] * ar, d1 = (lam[r] - 1.0) * ar;
              dq += 2.0 * (d1 - d0) * (x * v) + (d1 * d1 - d0 * d0) * (x * x * sii);
            }
            const double bayes = -0.5 * dq + (lp1 - lp0);
            const double g2 = log(rng.uniform(VB_BVS_U2, sweep, r));
            newlam = (bayes >= g2) ? 1.0 : 0.0;
          }
        }
        tmp[r] = newlam;
      }
      g.sync();
      for (int r = tid; r < n; r += G) lam[r] = tmp[r];
      g.sync();
      for (int e = tid; e < n * T; e += G) apath[e] *= lam[e % n];                     // :328
      g.sync();
      tvp_residuals(g, m, apath, lam, false, U);
    }

    // ================= covariance (psi) block (bvartvpalg.cpp:339-407) =================
    if (covar) {
      for (int i = 1; i < k; ++i) {
        const int o = (i * (i - 1)) / 2;
        tvp_single_eq_draw(g, rng, sweep, i, T, U, k, -1.0, U + i, k, sv ? ew + (size_t)T * i : sig + i + k * i, sv ? 1 : 0,
                           psi_bvs ? plam + o : nullptr, pq + o, pinit + o, n_psi, o, Dt, Mi, Cq, wv, cvec, dinv, anext,
                           tmp, tmp2, Mpath, wpath, psip + o, n_psi, fail);
      }
      if (psi_bvs) {
        // BVS on the time-varying covariances (:365-398).  psi_AG = Lambda psi is computed once (:372), so the
        // decisions of different positions are independent given the pre-sweep Lambda; theta0 / theta1 differ only in
        // row r = (i, j), i.e. only the residuals of equation i:  e0_t = u_it + sum_{j' != j} u_j't lam_j' psi_j't,
        // e1_t = e0_t + u_jt psi_rt, and l1 - l0 = -1/2 sum_t w_it (e1_t^2 - e0_t^2) + log prior odds.
        for (int r = tid; r < n_psi; r += G) {
          double newlam = plam[r];
          if (m.psi_include[r]) {
            int i = 1;
            while ((i * (i + 1)) / 2 <= r) ++i;
            const int o = (i * (i - 1)) / 2, j = r - o;
            const double lp0 = log(1.0 - m.psi_inprior[r]), lp1 = log(m.psi_inprior[r]);
            const double g1 = log(rng.uniform(VB_PSI_U1, sweep, r));
            const bool gate = (plam[r] == 1.0) ? (g1 < lp1) : (g1 < lp0);
            if (gate) {
              double dq = 0.0;
              for (int t = 0; t < T; ++t) {
                const double* ut = U + (size_t)k * t;
                const double* ps = psip + (size_t)n_psi * t + o;
                double e0 = ut[i];
                for (int j2 = 0; j2 < i; ++j2)
                  if (j2 != j) e0 = fma(ut[j2], plam[o + j2] * ps[j2], e0);
                const double d = ut[j] * ps[j];
                const double wt = sv ? ew[t + T * i] : sig[i + k * i];
                dq = fma(wt, d * (2.0 * e0 + d), dq);
              }
              const double bayes = -0.5 * dq + (lp1 - lp0);
              const double g2 = log(rng.uniform(VB_PSI_U2, sweep, r));
              newlam = (bayes >= g2) ? 1.0 : 0.0;
            }
          }
          pnew[r] = newlam;
        }
        g.sync();
        for (int r = tid; r < n_psi; r += G) plam[r] = pnew[r];
        g.sync();
        for (int e = tid; e < n_psi * T; e += G) psip[e] *= plam[e % n_psi];                 // :396
        g.sync();
      }
      // u_t <- Psi_t u_t (:406): rows from the last to the first, in place
      for (int t = tid; t < T; t += G) {
        const double* ps = psip + (size_t)n_psi * t;
        double* ut = U + (size_t)k * t;
        for (int i = k - 1; i >= 1; --i) {
          double s = ut[i];
          for (int j = 0; j < i; ++j) s = fma(ps[(i * (i - 1)) / 2 + j], ut[j], s);
          ut[i] = s;
        }
      }
      g.sync();
    }

    // ================= error variances =================
    double* out_sig = nullptr;
    if (keep) out_sig = out.sigma + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)((sv || covar) ? kk * T : kk);
    if (sv) {
#if defined(__CUDA_ARCH__)
      if constexpr (Grp::kWarp32) sv_step_w32(rng, sweep, m, U, h, sigma_h, h_init, ew, svw, svtab);
      else
#endif
      sv_step(g, rng, sweep, m, U, h, sigma_h, h_init, ew, svw);                       // :414
      BVG_TVPC_TICK(9);
      for (int i = tid; i < k; i += G) {
        double ss = 0.0, prev = h_init[i];
        for (int t = 0; t < T; ++t) { const double d = h[t + T * i] - prev; ss = fma(d, d, ss); prev = h[t + T * i]; }
        const double scale = 1.0 / (m.g_rate[i] + ss * 0.5);
        kv[i] = 1.0 / (scale * rng.gamma(VB_SIGH_G, sweep, i, m.g_shape_post[i]));     // :422
      }
      g.sync();
      for (int i = tid; i < k; i += G) sigma_h[i] = kv[i];
      for (int e = tid; e < kk; e += G) {
        const int i = e % k, j = e / k;
        S[e] = m.sv_vi[e] + ((i == j) ? 1.0 / kv[i] : 0.0);
      }
      for (int i = tid; i < k; i += G) {
        double s = (1.0 / kv[i]) * h[0 + T * i];
        for (int j = 0; j < k; ++j) s = fma(m.sv_vi[i + k * j], m.sv_mu[j], s);
        W4[i] = s;
      }
      g.sync();
      for (int e = tid; e < tri(k); e += G) {
        int i = 0; while (tri(i + 1) <= e) ++i;
        W1[e] = S[i + k * (e - tri(i))];
      }
      g.sync();
      chol_crout_generic(g, W1, kd, W4, k, fail);
      for (int i = tid; i < k; i += G) W4[i] += rng.normal(VB_HINIT_N, sweep, i);
      g.sync();
      back_solve_lt_generic(g, W1, kd, W4, h_init, k);                                         // :429
      if (Grp::kWarp32 && !covar) {
        // exp(-h) once per sweep: the measurement precisions of the next sweep (:446) and, inverted, the stored Sigma_t
        for (int e = tid; e < k * T; e += G) { const double w = 1.0 / exp(h[e]); ew[e] = w; svw[e] = 1.0 / w; }
        g.sync();
        if (keep)
          for (int e = tid; e < kk * T; e += G) {
            const int t = e / kk, f = e - t * kk, j = f / k, i = f - j * k;
            out_sig[e] = (i == j) ? svw[t + T * i] : 0.0;                              // :507-509
          }
      } else if (keep && !covar) {
        for (int e = tid; e < kk * T; e += G) {
          const int t = e / kk, f = e % kk, i = f % k, j = f / k;
          out_sig[e] = (i == j) ? 1.0 / (1.0 / exp(h[t + T * i])) : 0.0;               // :507-509
        }
      }
      if (keep) {
        double* osh = out.sigma_h + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)k;
        for (int i = tid; i < k; i += G) osh[i] = kv[i];
      }
      g.sync();
    } else {
      for (int e = tid; e < kk; e += G) {
        const int i = e % k, j = e / k;
        double s = 0.0;
        for (int t = 0; t < T; ++t) s = fma(U[i + k * t], U[j + k * t], s);
        S[e] = s;
      }
      g.sync();
      if (m.sigma_type == SIG_GAMMA && covar) {
        // bvartvpalg.cpp:433-437,454-455: the drawn omega_i never reaches diag_omega_i (only the SV branch refreshes
        // it, :446), so Sigma_t^-1 = Psi_t' Omega_0^-1 Psi_t with the INITIAL measurement precisions -- restated as is
      } else if (m.sigma_type == SIG_GAMMA) {
        for (int i = tid; i < k; i += G) {
          const double gs = rng.gamma(VB_OMEGA_G, sweep, i, m.g_shape_post[i]);
          kv[i] = gs * (1.0 / (m.g_rate[i] + S[i + k * i] * 0.5));                     // :436
        }
        g.sync();
        for (int e = tid; e < kk; e += G) {
          const int i = e % k, j = e / k;
          sig[e] = (i == j) ? kv[i] : m.omega_off[e];
        }
        g.sync();
        if (keep) {
          small_spd_inverse(g, sig, W4, W1, W2, W3, kd, k, fail);
          for (int e = tid; e < kk; e += G) out_sig[e] = W4[e];
        }
      } else {
        for (int e = tid; e < kk; e += G) S[e] += m.S0[e];
        g.sync();
        wishart_block(g, rng, sweep, k, m.wish_df_post, S, sig, W4, keep, W1, W2, W3, W4, kd, fail);   // :461
        if (keep) for (int e = tid; e < kk; e += G) out_sig[e] = W4[e];
      }
      g.sync();
    }

    if (covar) {
      // Sigma_t^-1 = Psi_t' Omega_t^-1 Psi_t (:448,455) and the stored Sigma_t = Psi_t^-1 Omega_t Psi_t^-T (:507-509)
      for (int t = tid; t < T; t += G) {
        const double* ps = psip + (size_t)n_psi * t;
        double P[BVG_TVP_PSI_MAXK][BVG_TVP_PSI_MAXK], om[BVG_TVP_PSI_MAXK];
        for (int i = 0; i < k; ++i) {
          om[i] = sv ? 1.0 / exp(h[t + T * i]) : sig[i + k * i];
          for (int j = 0; j < k; ++j) P[i][j] = (i == j) ? 1.0 : ((j < i) ? ps[(i * (i - 1)) / 2 + j] : 0.0);
        }
        double* sg = sigt + (size_t)kk * t;
        for (int a = 0; a < k; ++a)
          for (int c = 0; c <= a; ++c) {
            double s = 0.0;
            for (int b = a; b < k; ++b) s = fma(P[b][a] * om[b], P[b][c], s);
            sg[a + k * c] = s; sg[c + k * a] = s;
          }
        if (keep) {
          // unit-lower inverse by forward substitution, in place of P
          for (int i = 1; i < k; ++i)
            for (int c = 0; c < i; ++c) {
              double s = -P[i][c];
              for (int j = c + 1; j < i; ++j) s = fma(-P[i][j], (j > c) ? P[j][c] : 0.0, s);
              P[i][c] = s;
            }
          // (rows above i already hold the inverse when row i is formed: P[i][j] on the right is the ORIGINAL row i)
          double* os = out_sig + (size_t)kk * t;
          for (int a = 0; a < k; ++a)
            for (int c = 0; c <= a; ++c) {
              double s = 0.0;
              for (int b = 0; b <= c; ++b) s = fma(P[a][b] / om[b], P[c][b], s);
              os[a + k * c] = s; os[c + k * a] = s;
            }
        }
      }
      g.sync();
      // psi state variances (:485-493) and initial state (:495-498)
      for (int r = tid; r < n_psi; r += G) {
        double ss = 0.0, prev = pinit[r];
        for (int t = 0; t < T; ++t) { const double v = psip[(size_t)n_psi * t + r]; const double d = v - prev; ss = fma(d, d, ss); prev = v; }
        tmp[r] = rng.gamma(VB_PSI_G, sweep, r, tm.psi_shape_post[r]) * (1.0 / (tm.psi_rate[r] + ss * 0.5));
      }
      g.sync();
      for (int r = tid; r < n_psi; r += G) pq[r] = tmp[r];
      g.sync();
      for (int r = tid; r < n_psi; r += G) {
        double* Dr = Dt + tri(r);
        double b = pq[r] * psip[r];
        for (int c = 0; c < n_psi; ++c) {
          const double v = m.psi_vi[r + n_psi * c];
          if (c <= r) Dr[c] = v + ((c == r) ? pq[r] : 0.0);
          b = fma(v, m.psi_mu[c], b);
        }
        wv[r] = b;
      }
      g.sync();
      chol_crout(g, Dt, dinv, wv, n_psi, fail);
      for (int r = tid; r < n_psi; r += G) wv[r] += rng.normal(VB_PSI_IN, sweep, r);
      g.sync();
      back_solve_lt(g, Dt, dinv, wv, pinit, n_psi);
    }

    BVG_TVPC_TICK(10);
    // ================= state variances (:469-477) and initial state (:480-482) =================
    for (int r = tid; r < n; r += G) {
      double ss = 0.0, prev = a_init[r];
      if (eq8 && !bvs) ss = tmp2[r];                  // accumulated by the backward pass of the state draw
      else
        for (int t = 0; t < T; ++t) { const double v = apath[(size_t)t * n + r]; const double d = v - prev; ss = fma(d, d, ss); prev = v; }
      const double scale = 1.0 / (tm.v_rate[r] + ss * 0.5);
      tmp[r] = rng.gamma(VB_SIGV_G, sweep, r, tm.v_shape_post[r]) * scale;             // :476
    }
    g.sync();
    BVG_TVPC_TICK(11);
    for (int r = tid; r < n; r += G) q[r] = tmp[r];
    g.sync();
    if (eq8 && m.vi_off == nullptr) {
      // diagonal prior precision: the n x n system of :480-482 is diagonal, L = sqrt(d), a_init = (b / L + eps) / L
      for (int r = tid; r < n; r += G) {
        const double a0 = fused ? xs[68 * k + 12 + 8 * (r % k) + r / k] : apath[r];          // a_0 (Eq8Smem::AN after the backward pass)
        const double v0 = m.vdiag0[r], d = v0 + q[r];
        if (!(d > 0.0)) fail = 1;
        const double rs = bvg_rsqrt(d);
        a_init[r] = (fma(v0, m.mu[r], q[r] * a0) * rs + rng.normal(VB_AINIT_N, sweep, r)) * rs;
      }
      g.sync();
    } else {
    for (int r = tid; r < n; r += G) {
      double* Dr = Dt + tri(r);
      for (int c = 0; c <= r; ++c) {
        double v = (m.vi_off != nullptr) ? m.vi_off[tri(r) + c] : 0.0;
        if (c == r) v += m.vdiag0[r] + q[r];
        Dr[c] = v;
      }
      double b = m.vdiag0[r] * m.mu[r] + q[r] * (fused ? xs[68 * k + 12 + 8 * (r % k) + r / k] : apath[r]);
      if (m.vmu_off != nullptr) b += m.vmu_off[r];
      wv[r] = b;
    }
    g.sync();
    chol_crout(g, Dt, dinv, wv, n, fail);
    for (int r = tid; r < n; r += G) wv[r] += rng.normal(VB_AINIT_N, sweep, r);
    g.sync();
    back_solve_lt(g, Dt, dinv, wv, a_init, n);
    }

    BVG_TVPC_TICK(12);
    // ================= store (:502-551) =================
    if (keep && m.a0_from >= 0) {
      // structural model: only the kept rows [a | b | c | a0 (variable j, equation i > j)] of the masked augmented system
      const int a0 = m.a0_from, n_out = a0 + (k * (k - 1)) / 2;
      double* oc = out.coef + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)n_out * (size_t)T;
      double* ov = out.sigma_v + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)n_out;
      double* ol = (out.lambda != nullptr) ? out.lambda + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)n_out
                                           : nullptr;
      for (int r = tid; r < n; r += G) {
        int o = r;
        if (r >= a0) {
          const int j = (r - a0) / k, i = (r - a0) % k;
          if (i <= j) continue;
          o = a0 + j * (k - 1) - (j * (j - 1)) / 2 + (i - j - 1);
        }
        for (int t = 0; t < T; ++t) oc[(size_t)t * n_out + o] = apath[(size_t)t * n + r];
        ov[o] = 1.0 / q[r];
        if (ol != nullptr) ol[o] = lam[r];
      }
    } else if (keep) {
      double* oc = out.coef + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)n * (size_t)T;
      if (!fused)
        for (int e = tid; e < n * T; e += G) oc[e] = apath[e];
      double* ov = out.sigma_v + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)n;
      for (int r = tid; r < n; r += G) ov[r] = 1.0 / q[r];                             // :525
      if (out.lambda != nullptr && lam_rows > 0) {
        double* ol = out.lambda + ((size_t)chain_local * (size_t)out.iters + (size_t)pos) * (size_t)lam_rows;
        if (bvs) for (int r = tid; r < n; r += G) ol[r] = lam[r];
        if (psi_bvs) for (int r = tid; r < n_psi; r += G) ol[(bvs ? n : 0) + r] = plam[r];   // draws_lambda_a0, :517-519
      }
    }
    g.sync();
    BVG_TVPC_TICK(13);
  }

  for (int e = tid; e < kk; e += G) state[e] = sig[e];
  for (int e = tid; e < n; e += G) { st_q[e] = q[e]; st_ai[e] = a_init[e]; st_lam[e] = lam[e]; }
  if (sv) {
    for (int e = tid; e < T * k; e += G) st_h[e] = h[e];
    for (int e = tid; e < k; e += G) { st_h[T * k + e] = sigma_h[e]; st_h[T * k + k + e] = h_init[e]; }
  }
  if (covar) {
    for (int e = tid; e < n_psi; e += G) { st_psi[e] = pq[e]; st_psi[n_psi + e] = pinit[e]; st_psi[2 * n_psi + e] = plam[e]; }
    for (int e = tid; e < T * kk; e += G) st_psi[3 * n_psi + e] = sigt[e];
  }
  if (fail && tid == 0) out.status[chain_local] = 1;
  g.sync();
}

}  // namespace bvg
