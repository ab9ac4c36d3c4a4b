// bvg_blocks.cuh -- group-generic logic of the batched stand-alone building blocks:
//   post_normal      src/post_normal.cpp:61-93        post_normal_sur  src/post_normal_sur.cpp:60-113
//   ssvs             src/svss.cpp:76-111              bvs              src/bvs.cpp:73-159 (constant parameters)
//   stochvol_*       src/stochvol_ksc1998.cpp:60-108, src/stochvol_ocsn2007.cpp:60-114   (sv_step, bvg_bvar_sweep.cuh)
//   kalman_dk        src/kalman_dk.cpp:91-184
// One problem of the batch is owned by one group (a CTA in the CUDA kernels).  Matrices are full column-major
// unless stated otherwise.
#pragma once
#include "bvg_bvar_sweep.cuh"

namespace bvg {

// ---- cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (the arma::eig_sym behind the reference's
//      symmetric square roots, post_normal.cpp:91-93, post_normal_sur.cpp:109-113, kalman_dk.cpp:130-147).
//      A is destroyed (diagonal = eigenvalues on exit), V = eigenvectors (columns).  The symmetric square root
//      U diag(f(s)) U' is independent of eigenvector order / sign, so any convergent eigen solver reproduces it.
template <class Grp>
BVG_HD void jacobi_eig_sym(const Grp& g, double* A, double* V, int n) {
  const int G = g.size(), tid = g.tid();
  for (int e = tid; e < n * n; e += G) V[e] = (e % n == e / n) ? 1.0 : 0.0;
  g.sync();
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int e = tid; e < n * n; e += G) {
      const double v = A[e] * A[e];
      if (e % n == e / n) dg += v; else off += v;
    }
    off = g.sum(off);
    dg = g.sum(dg);
    if (off <= 1e-30 * dg || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p) {
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p + n * q];
        const double app = A[p + n * p], aqq = A[q + n * q];
        g.sync();                                    // everyone has read the pivot block
        if (fabs(apq) > 1e-300 && fabs(apq) > 1e-19 * sqrt(fabs(app * aqq)) ) {
          const double theta = (aqq - app) / (2.0 * apq);
          const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
          // rows/cols p, q of A and columns p, q of V; thread i handles index i
          for (int i = tid; i < n; i += G) {
            if (i != p && i != q) {
              const double aip = A[i + n * p], aiq = A[i + n * q];
              const double nip = c * aip - s * aiq, niq = s * aip + c * aiq;
              A[i + n * p] = nip; A[p + n * i] = nip;
              A[i + n * q] = niq; A[q + n * i] = niq;
            }
            const double vip = V[i + n * p], viq = V[i + n * q];
            V[i + n * p] = c * vip - s * viq;
            V[i + n * q] = s * vip + c * viq;
          }
          if (tid == 0) {
            A[p + n * p] = app - t * apq;
            A[q + n * q] = aqq + t * apq;
            A[p + n * q] = 0.0;
            A[q + n * p] = 0.0;
          }
        }
        g.sync();
      }
    }
  }
  g.sync();
}

// out = U diag(f) U'  with f = sqrt(s) (power = +1/2), 1/sqrt(s) (-1/2) or 1/s (-1)
template <class Grp>
BVG_HD void sym_func(const Grp& g, const double* A_eig, const double* V, double* out, int n, int mode) {
  for (int e = g.tid(); e < n * n; e += g.size()) {
    const int i = e % n, j = e / n;
    double s = 0.0;
    for (int m = 0; m < n; ++m) {
      const double lam = A_eig[m + n * m];
      const double f = (mode == 0) ? sqrt(lam) : (mode == 1 ? 1.0 / sqrt(lam) : 1.0 / lam);
      s = fma(V[i + n * m] * f, V[j + n * m], s);
    }
    out[e] = s;
  }
  g.sync();
}

// ---- posterior normal draw from a built precision P (full n x n) and rhs b
//      method 0: a = P^-1 b + sqrtm(P^-1) eps  (symmetric eigen square root: post_normal.cpp:86-93)
//      method 1: a = L^-T (L^-1 b + eps)       (Cholesky mapping of the built-in samplers: bvaralg.cpp:306-307)
// work: W1, W2 (n*n each), v1 (n).  P is destroyed.  eps read through rng (block VB_COEF_N, sweep 0).
template <class Grp>
BVG_HD void normal_draw_from_precision(const Grp& g, const ChainRng& rng, double* P, double* b, double* W1,
                                       double* W2, double* v1, double* out, int n, int method, int& fail) {
  const int G = g.size(), tid = g.tid();
  if (method == 1) {
    for (int e = tid; e < tri(n); e += G) {
      int i = 0; while (tri(i + 1) <= e) ++i;
      W1[e] = P[i + n * (e - tri(i))];
    }
    g.sync();
    chol_crout_generic(g, W1, v1, b, n, fail);
    for (int r = tid; r < n; r += G) b[r] += rng.normal(VB_COEF_N, 0, r);
    g.sync();
    back_solve_lt_generic(g, W1, v1, b, W2, n);
    for (int r = tid; r < n; r += G) out[r] = W2[r];
    g.sync();
    return;
  }
  jacobi_eig_sym(g, P, W1, n);                    // P = U diag(l) U'
  for (int m = tid; m < n; m += G) if (!(P[m + n * m] > 0.0)) fail = 1;
  // c = U' b, d = U' eps ;  a = U (c / l + d / sqrt(l))
  for (int r = tid; r < n; r += G) W2[r] = rng.normal(VB_COEF_N, 0, r);
  g.sync();
  for (int m = tid; m < n; m += G) {
    double c = 0.0, d = 0.0;
    for (int i = 0; i < n; ++i) {
      c = fma(W1[i + n * m], b[i], c);
      d = fma(W1[i + n * m], W2[i], d);
    }
    const double lam = P[m + n * m];
    v1[m] = c / lam + d / sqrt(lam);
  }
  g.sync();
  for (int i = tid; i < n; i += G) {
    double s = 0.0;
    for (int m = 0; m < n; ++m) s = fma(W1[i + n * m], v1[m], s);
    out[i] = s;
  }
  g.sync();
}

// post_normal: P = V^-1 + (x x') (x) Sigma^-1, b = V^-1 mu + vec(Sigma^-1 y x')        (post_normal.cpp:86-87)
template <class Grp>
BVG_HD void post_normal_problem(const Grp& g, const ChainRng& rng, int k, int M, const double* XX, const double* YX,
                                const double* sig, const double* mu, const double* Vi, double* sm, double* out,
                                int method, int& fail) {
  const int n = k * M;
  double* P = sm; double* W1 = P + n * n; double* W2 = W1 + n * n; double* b = W2 + n * n; double* v1 = b + n;
  for (int e = g.tid(); e < n * n; e += g.size()) {
    const int r = e % n, c = e / n;
    P[e] = Vi[e] + XX[(r / k) + M * (c / k)] * sig[(r % k) + k * (c % k)];
  }
  for (int r = g.tid(); r < n; r += g.size()) {
    const int j = r / k, i = r % k;
    double s = 0.0;
    for (int c = 0; c < n; ++c) s = fma(Vi[r + n * c], mu[c], s);
    for (int i2 = 0; i2 < k; ++i2) s = fma(sig[i + k * i2], YX[i2 + k * j], s);
    b[r] = s;
  }
  g.sync();
  normal_draw_from_precision(g, rng, P, b, W1, W2, v1, out, n, method, fail);
}
BVG_HD int post_normal_smem_len(int n) { return 3 * n * n + 2 * n; }

// post_normal_sur: general SUR regressor matrix z (kT x n); sigma_i k x k, or kT x k when time varying
// (post_normal_sur.cpp:84-104)
template <class Grp>
BVG_HD void post_normal_sur_problem(const Grp& g, const ChainRng& rng, int k, int T, int n, const double* y,
                                    const double* z, const double* sig, int sigma_tv, const double* mu,
                                    const double* Vi, double* sm, double* out, int method, int& fail) {
  const int kT = k * T;
  double* P = sm; double* W1 = P + n * n; double* W2 = W1 + n * n; double* b = W2 + n * n; double* v1 = b + n;
  for (int e = g.tid(); e < n * n; e += g.size()) {
    const int r = e % n, c = e / n;
    double s = Vi[e];
    for (int t = 0; t < T; ++t) {
      for (int i = 0; i < k; ++i) {
        const double zr = z[(t * k + i) + kT * r];
        if (zr == 0.0) continue;
        double v = 0.0;
        for (int i2 = 0; i2 < k; ++i2) {
          const double sv = sigma_tv ? sig[(t * k + i) + kT * i2] : sig[i + k * i2];
          v = fma(sv, z[(t * k + i2) + kT * c], v);
        }
        s = fma(zr, v, s);
      }
    }
    P[e] = s;
  }
  for (int r = g.tid(); r < n; r += g.size()) {
    double s = 0.0;
    for (int c = 0; c < n; ++c) s = fma(Vi[r + n * c], mu[c], s);
    for (int t = 0; t < T; ++t)
      for (int i = 0; i < k; ++i) {
        const double zr = z[(t * k + i) + kT * r];
        if (zr == 0.0) continue;
        double v = 0.0;
        for (int i2 = 0; i2 < k; ++i2) {
          const double sv = sigma_tv ? sig[(t * k + i) + kT * i2] : sig[i + k * i2];
          v = fma(sv, y[i2 + k * t], v);
        }
        s = fma(zr, v, s);
      }
    b[r] = s;
  }
  g.sync();
  normal_draw_from_precision(g, rng, P, b, W1, W2, v1, out, n, method, fail);
}

// ssvs (svss.cpp:78-108) for one coefficient: posterior inclusion probability, Bernoulli draw, prior precision.
// Non-included positions keep lambda = 1 and 1 / tau1^2 (:84-86, :94-97).
BVG_HD void ssvs_element(const ChainRng& rng, int r, double ar, double t0, double t1, double pi, bool included,
                         double* lam, double* vdiag) {
  const double t0sq = t0 * t0, t1sq = t1 * t1;
  double ld = 1.0, v = 1.0 / t1sq;
  if (included) {
    const double u0 = 1.0 / t0 * exp(-((ar * ar) / (2.0 * t0sq))) * (1.0 - pi);  // :81
    const double u1 = 1.0 / t1 * exp(-((ar * ar) / (2.0 * t1sq))) * pi;          // :82
    const double p = u1 / (u0 + u1);
    ld = bernoulli_rbinom(p, rng.uniform(VB_SSVS_U, 0, r));                      // :104
    if (ld == 0.0) v = 1.0 / t0sq;                                               // :106-108
  }
  *lam = ld;
  *vdiag = v;
}

// bvs (bvs.cpp:73-159) in its full generality: SUR regressor matrix z (kT x n), a constant (n) or time varying
// (n x T, col-major), Sigma^-1 constant (k x k) or time varying (kT x k stack).  Rank-1 form of the likelihood
// ratio, coefficient-parallel (every decision uses the stale A_G = Lambda a, bvs.cpp:105):
//   r_theta,t = r_base,t + z_t[,j] delta_t,   delta0_t = lambda_j a_jt,  delta1_t = (lambda_j - 1) a_jt
//   l1 - l0 = -1/2 sum_t [ 2 (d1_t - d0_t) c_t + (d1_t^2 - d0_t^2) g_t ] + (lp1 - lp0),
//   c_t = z_t[,j]' S_t r_base,t,   g_t = z_t[,j]' S_t z_t[,j].        sm: U (k*T) + W (k*T).
template <class Grp>
BVG_HD void bvs_problem(const Grp& g, const ChainRng& rng, int k, int T, int n, const double* y, const double* z,
                        const double* a, int a_tv, const double* lam, const double* sig, int sigma_tv,
                        const double* prob, const unsigned char* include, double* sm, double* lam_out) {
  const int kT = k * T;
  double* U = sm;
  double* W = sm + kT;
  for (int e = g.tid(); e < kT; e += g.size()) {
    const int t = e / k;
    double s = y[e];
    for (int r = 0; r < n; ++r) {
      const double zr = z[e + (size_t)kT * r];
      if (zr == 0.0) continue;
      s = fma(-(lam[r] * (a_tv ? a[r + (size_t)n * t] : a[r])), zr, s);
    }
    U[e] = s;
  }
  g.sync();
  for (int e = g.tid(); e < kT; e += g.size()) {
    const int i = e % k, t = e / k;
    double v = 0.0;
    for (int i2 = 0; i2 < k; ++i2)
      v = fma(sigma_tv ? sig[(t * k + i) + (size_t)kT * i2] : sig[i + k * i2], U[i2 + k * t], v);
    W[e] = v;
  }
  g.sync();
  for (int r = g.tid(); r < n; r += g.size()) {
    double nl = lam[r];
    if (include[r]) {
      const double lp0 = log(1.0 - prob[r]), lp1 = log(prob[r]);
      const double g1 = log(rng.uniform(VB_BVS_U1, 0, r));
      const bool gate = (lam[r] == 1.0) ? (g1 < lp1) : (g1 < lp0);
      if (gate) {
        const double* zr = z + (size_t)kT * r;
        double dq = 0.0;
        for (int t = 0; t < T; ++t) {
          double ct = 0.0, gt = 0.0;
          for (int i = 0; i < k; ++i) {
            const double zi = zr[t * k + i];
            if (zi == 0.0) continue;
            ct = fma(zi, W[i + k * t], ct);
            double v = 0.0;
            for (int i2 = 0; i2 < k; ++i2)
              v = fma(sigma_tv ? sig[(t * k + i) + (size_t)kT * i2] : sig[i + k * i2], zr[t * k + i2], v);
            gt = fma(zi, v, gt);
          }
          const double ar = a_tv ? a[r + (size_t)n * t] : a[r];
          const double d0 = lam[r] * ar, d1 = (lam[r] - 1.0) * ar;
          dq += 2.0 * (d1 - d0) * ct + (d1 * d1 - d0 * d0) * gt;
        }
        const double bayes = -0.5 * dq + (lp1 - lp0);
        nl = (bayes >= log(rng.uniform(VB_BVS_U2, 0, r))) ? 1.0 : 0.0;
      }
    }
    lam_out[r] = nl;
  }
  g.sync();
}

// ---- kalman_dk (kalman_dk.cpp:91-184), constant sigma_u / sigma_v / B per problem.
// Shared memory (doubles): 7 n*n + 4 k*k + n*k*2 + 6 n + 3 k.  Global scratch per problem:
//   aplus n(T+1) | yplus kT | v kT | Fi k*k*T | Kt n*k*T | r n*T
BVG_HD int kalman_smem_len(int k, int n) { return 7 * n * n + 4 * k * k + 2 * n * k + 6 * n + 3 * k; }
BVG_HD long long kalman_scratch_len(int k, int T, int n) {
  return (long long)n * (T + 1) + 2LL * k * T + (long long)k * k * T + (long long)n * k * T + (long long)n * T;
}

// C (r x c) = A (r x m) * B (m x c) [optionally B transposed], all col-major with leading dims = rows
template <class Grp>
BVG_HD void mm(const Grp& g, const double* A, const double* B, double* C, int r, int m, int c, bool bt) {
  for (int e = g.tid(); e < r * c; e += g.size()) {
    const int i = e % r, j = e / r;
    double s = 0.0;
    for (int l = 0; l < m; ++l) s = fma(A[i + r * l], bt ? B[j + c * l] : B[l + m * j], s);
    C[e] = s;
  }
  g.sync();
}

// inverse of a symmetric positive definite k x k matrix (the arma::inv at kalman_dk.cpp:165)
template <class Grp>
BVG_HD void spd_inverse_full(const Grp& g, const double* A, double* out, double* W1, double* W2, double* W3,
                             double* kd, int k, int& fail) {
  small_spd_inverse(g, A, out, W1, W2, W3, kd, k, fail);
}

// `tv` bit mask: 1 = sigma_u is a kT x k stack, 2 = sigma_v is an nT x n stack, 4 = B is an nT x n stack
// (kalman_dk.cpp:99-121 expands the constant forms to these stacks; here the constant forms are indexed directly).
template <class Grp>
BVG_HD void kalman_dk_problem(const Grp& g, const ChainRng& rng, int k, int T, int n, const double* y,
                              const double* z, const double* su, const double* sv, const double* Bm, int tv,
                              const double* a_init, const double* P_init, double* sm, double* scratch,
                              double* out, int& fail) {
  const int G = g.size(), tid = g.tid(), nn = n * n, kk = k * k, kT = k * T;
  const bool su_tv = tv & 1, sv_tv = tv & 2, b_tv = tv & 4;
  const size_t ldu = su_tv ? (size_t)kT : (size_t)k, ldv = sv_tv ? (size_t)n * T : (size_t)n,
               ldb = b_tv ? (size_t)n * T : (size_t)n;
  double* A1 = sm;            // n*n work / eigen input
  double* A2 = A1 + nn;       // eigenvectors
  double* RP = A2 + nn;       // sqrtm(P_init)
  double* RV = RP + nn;       // sqrtm(sigma_v)
  double* Pm = RV + nn;       // P_t
  double* Lm = Pm + nn;       // L_t
  double* T1 = Lm + nn;       // n*n temp
  double* RU = T1 + nn;       // sqrtm(sigma_u) k*k
  double* F = RU + kk; double* Fi = F + kk; double* K1 = Fi + kk;          // k*k each
  double* PZ = K1 + kk;       // n*k  (P Z')
  double* Km = PZ + n * k;    // n*k  K_t
  double* av = Km + n * k; double* av2 = av + n; double* rv = av2 + n; double* rv2 = rv + n;
  double* tn = rv2 + n; double* tn2 = tn + n;                               // n each
  double* vk = tn2 + n; double* tk = vk + k; double* kd = tk + k;           // k each
  (void)tn; (void)tn2;
  double* aplus = scratch;                   // n x (T+1)
  double* yplus = aplus + (size_t)n * (T + 1);
  double* vst = yplus + kT;                  // k x T
  double* Fis = vst + kT;                    // k*k x T
  double* Ks = Fis + (size_t)kk * T;         // n*k x T
  double* rs = Ks + (size_t)n * k * T;       // n x T

  // symmetric square roots (kalman_dk.cpp:130-131, :141-142, :145-146)
  for (int e = tid; e < nn; e += G) A1[e] = P_init[e];
  g.sync();
  jacobi_eig_sym(g, A1, A2, n);
  sym_func(g, A1, A2, RP, n, 0);

  // Step 1: simulate a+, y+ ; variates: eps0 (n), then per t eps_y (k), eps_a (n)   (:132, :143, :147)
  for (int i = tid; i < n; i += G) {
    double s = 0.0;
    for (int m = 0; m < n; ++m) s = fma(RP[i + n * m], rng.normal(VB_STATE_N, 0, m), s);
    aplus[i] = s;
  }
  g.sync();
  for (int t = 0; t < T; ++t) {
    const double* sut = su + (su_tv ? (size_t)t * k : 0);
    const double* svt = sv + (sv_tv ? (size_t)t * n : 0);
    const double* Bt = Bm + (b_tv ? (size_t)t * n : 0);
    if (t == 0 || sv_tv) {
      for (int e = tid; e < nn; e += G) A1[e] = svt[(e % n) + ldv * (e / n)];
      g.sync();
      jacobi_eig_sym(g, A1, A2, n);
      sym_func(g, A1, A2, RV, n, 0);
    }
    if (t == 0 || su_tv) {
      for (int e = tid; e < kk; e += G) F[e] = sut[(e % k) + ldu * (e / k)];
      g.sync();
      jacobi_eig_sym(g, F, Fi, k);
      sym_func(g, F, Fi, RU, k, 0);
    }
    const int base = n + t * (k + n);
    const double* ap = aplus + (size_t)n * t;
    for (int i = tid; i < k; i += G) {
      double s = 0.0;
      for (int m = 0; m < n; ++m) s = fma(z[(t * k + i) + (size_t)kT * m], ap[m], s);
      for (int m = 0; m < k; ++m) s = fma(RU[i + k * m], rng.normal(VB_STATE_N, 0, base + m), s);
      yplus[i + k * t] = s;
    }
    for (int i = tid; i < n; i += G) {
      double s = 0.0;
      for (int m = 0; m < n; ++m) s = fma(Bt[i + ldb * m], ap[m], s);
      for (int m = 0; m < n; ++m) s = fma(RV[i + n * m], rng.normal(VB_STATE_N, 0, base + k + m), s);
      aplus[(size_t)n * (t + 1) + i] = s;
    }
    g.sync();
  }

  // Step 2: filter on y* = y - y+   (:151-170)
  for (int i = tid; i < n; i += G) av[i] = a_init[i];
  for (int e = tid; e < nn; e += G) Pm[e] = P_init[e];
  g.sync();
  for (int t = 0; t < T; ++t) {
    const double* sut = su + (su_tv ? (size_t)t * k : 0);
    const double* svt = sv + (sv_tv ? (size_t)t * n : 0);
    const double* Bt = Bm + (b_tv ? (size_t)t * n : 0);
    // v = y* - Z a ;  PZ = P Z'
    for (int i = tid; i < k; i += G) {
      double s = y[i + k * t] - yplus[i + k * t];
      for (int m = 0; m < n; ++m) s = fma(-z[(t * k + i) + (size_t)kT * m], av[m], s);
      vk[i] = s;
      vst[i + k * t] = s;
    }
    for (int e = tid; e < n * k; e += G) {
      const int i = e % n, j = e / n;
      double s = 0.0;
      for (int m = 0; m < n; ++m) s = fma(Pm[i + n * m], z[(t * k + j) + (size_t)kT * m], s);
      PZ[e] = s;
    }
    g.sync();
    // F = Z P Z' + sigma_u ; Fi = inv(F)
    for (int e = tid; e < kk; e += G) {
      const int i = e % k, j = e / k;
      double s = sut[i + ldu * j];
      for (int m = 0; m < n; ++m) s = fma(z[(t * k + i) + (size_t)kT * m], PZ[m + n * j], s);
      F[e] = s;
    }
    g.sync();
    small_spd_inverse(g, F, Fi, K1, RU /*free after step 1*/, T1, kd, k, fail);
    for (int e = tid; e < kk; e += G) Fis[(size_t)kk * t + e] = Fi[e];
    // K = B P Z' Fi : T1(n x k) = PZ Fi ; Km = B T1
    mm(g, PZ, Fi, T1, n, k, k, false);
    for (int e = tid; e < n * k; e += G) {
      const int i = e % n, j = e / n;
      double s = 0.0;
      for (int m = 0; m < n; ++m) s = fma(Bt[i + ldb * m], T1[m + n * j], s);
      Km[e] = s;
      Ks[(size_t)n * k * t + e] = s;
    }
    g.sync();
    // L = B - K Z_t
    for (int e = tid; e < nn; e += G) {
      const int i = e % n, j = e / n;
      double s = Bt[i + ldb * j];
      for (int m = 0; m < k; ++m) s = fma(-Km[i + n * m], z[(t * k + m) + (size_t)kT * j], s);
      Lm[e] = s;
    }
    // a = B a + K v
    for (int i = tid; i < n; i += G) {
      double s = 0.0;
      for (int m = 0; m < n; ++m) s = fma(Bt[i + ldb * m], av[m], s);
      for (int m = 0; m < k; ++m) s = fma(Km[i + n * m], vk[m], s);
      av2[i] = s;
    }
    // T1 = B P
    for (int e = tid; e < nn; e += G) {
      const int i = e % n, j = e / n;
      double s = 0.0;
      for (int m = 0; m < n; ++m) s = fma(Bt[i + ldb * m], Pm[m + n * j], s);
      T1[e] = s;
    }
    g.sync();
    // P = B P L' + sigma_v
    for (int e = tid; e < nn; e += G) {
      const int i = e % n, j = e / n;
      double s = svt[i + ldv * j];
      for (int m = 0; m < n; ++m) s = fma(T1[i + n * m], Lm[j + n * m], s);
      A1[e] = s;
    }
    g.sync();
    for (int e = tid; e < nn; e += G) Pm[e] = A1[e];
    for (int i = tid; i < n; i += G) av[i] = av2[i];
    g.sync();
  }

  // backward recursion r_{t-1} = Z_t' Fi_t v_t + L_t' r_t   (:172-176), r_{T-1} = 0
  for (int i = tid; i < n; i += G) { rv[i] = 0.0; rs[(size_t)n * (T - 1) + i] = 0.0; }
  g.sync();
  for (int t = T - 1; t >= 0; --t) {
    const double* Bt = Bm + (b_tv ? (size_t)t * n : 0);
    // tk = Fi_t v_t ; L_t = B - K_t Z_t (recomputed rather than stored: n*n doubles per period saved)
    for (int i = tid; i < k; i += G) {
      double s = 0.0;
      for (int m = 0; m < k; ++m) s = fma(Fis[(size_t)kk * t + i + k * m], vst[m + k * t], s);
      tk[i] = s;
    }
    for (int e = tid; e < nn; e += G) {
      const int i = e % n, j = e / n;
      double s = Bt[i + ldb * j];
      for (int m = 0; m < k; ++m) s = fma(-Ks[(size_t)n * k * t + i + n * m], z[(t * k + m) + (size_t)kT * j], s);
      Lm[e] = s;
    }
    g.sync();
    for (int i = tid; i < n; i += G) {
      double s = 0.0;
      for (int m = 0; m < k; ++m) s = fma(z[(t * k + m) + (size_t)kT * i], tk[m], s);
      for (int m = 0; m < n; ++m) s = fma(Lm[m + n * i], rv[m], s);
      rv2[i] = s;
    }
    g.sync();
    for (int i = tid; i < n; i += G) {
      rv[i] = rv2[i];
      if (t > 0) rs[(size_t)n * (t - 1) + i] = rv2[i];
    }
    g.sync();
  }
  // rv now holds r0 (:177);  a_0 = a_init + P_init r0 ; a_{t+1} = B a_t + sigma_v r_t   (:178-181)
  for (int i = tid; i < n; i += G) {
    double s = a_init[i];
    for (int m = 0; m < n; ++m) s = fma(P_init[i + n * m], rv[m], s);
    av[i] = s;
    out[i] = s + aplus[i];
  }
  g.sync();
  for (int t = 0; t < T; ++t) {
    const double* svt = sv + (sv_tv ? (size_t)t * n : 0);
    const double* Bt = Bm + (b_tv ? (size_t)t * n : 0);
    for (int i = tid; i < n; i += G) {
      double s = 0.0;
      for (int m = 0; m < n; ++m) s = fma(Bt[i + ldb * m], av[m], s);
      for (int m = 0; m < n; ++m) s = fma(svt[i + ldv * m], rs[(size_t)n * t + m], s);
      av2[i] = s;
      out[(size_t)n * (t + 1) + i] = s + aplus[(size_t)n * (t + 1) + i];     // Step 3 (:184)
    }
    g.sync();
    for (int i = tid; i < n; i += G) av[i] = av2[i];
    g.sync();
  }
}

}  // namespace bvg
