/* blocks_ref.c -- TEST INFRASTRUCTURE ONLY: a second, independent restatement (plain C, dense matrices, plain loops) of
 * three blocks that the NumPy oracle (oracle/blocks.py, oracle/samplers.py) also restates, so that every block of the
 * hot path is read twice from the reference's source:
 *
 *   stochvol_ref   src/stochvol_ksc1998.cpp:60-108 / src/stochvol_ocsn2007.cpp:60-114  (mixture indicators, dense T x T
 *                  precision hh / sigma + diag(1 / sigma2_s), LU solve for the mean, upper Cholesky for the draw)
 *   bvs_ref        src/bvs.cpp:73-159   (theta0 / theta1 residuals, dense kT x kT quadratic forms, constant or TVP a,
 *                  constant or per-period Sigma^-1)
 *   tvp_state_ref  src/bvartvpalg.cpp:290-297  (dense (nT) x (nT) precision  H' (I (x) Sv^-1) H + zz' Su^-1 zz, LU solve,
 *                  upper Cholesky, a = mean + R^-1 eps)
 *
 * tests/test_oracle.py requires NumPy oracle == this file to rounding on the same injected variates.  Never linked into
 * the product.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static const double KSC_P[7] = {0.00730, 0.10556, 0.00002, 0.04395, 0.34001, 0.24566, 0.25750};
static const double KSC_MU[7] = {-10.12999, -3.97281, -8.56686, 2.77786, 0.61942, 1.79518, -1.08819};
static const double KSC_S2[7] = {5.79596, 2.61369, 5.17950, 0.16735, 0.64009, 0.34023, 1.26261};
static const double OCSN_P[10] = {0.00609, 0.04775, 0.13057, 0.20674, 0.22715, 0.18842, 0.12047, 0.05591, 0.01575, 0.00115};
static const double OCSN_MU[10] = {1.92677, 1.34744, 0.73504, 0.02266, -0.85173, -1.97278, -3.46788, -5.55246, -8.68384, -14.65000};
static const double OCSN_S2[10] = {0.11265, 0.17788, 0.26768, 0.40611, 0.62699, 0.98583, 1.57469, 2.54498, 4.16591, 7.33342};

/* A x = b by LU with partial pivoting (A n x n column-major, destroyed; b overwritten); 1 on a zero pivot */
static int lu_solve(int n, double* A, double* b) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double mx = fabs(A[c + (size_t)n * c]);
    for (int r = c + 1; r < n; ++r)
      if (fabs(A[r + (size_t)n * c]) > mx) { mx = fabs(A[r + (size_t)n * c]); piv = r; }
    if (mx == 0.0) return 1;
    if (piv != c) {
      for (int j = 0; j < n; ++j) { double t = A[c + (size_t)n * j]; A[c + (size_t)n * j] = A[piv + (size_t)n * j]; A[piv + (size_t)n * j] = t; }
      double t = b[c]; b[c] = b[piv]; b[piv] = t;
    }
    for (int r = c + 1; r < n; ++r) {
      const double f = A[r + (size_t)n * c] / A[c + (size_t)n * c];
      if (f == 0.0) continue;
      for (int j = c + 1; j < n; ++j) A[r + (size_t)n * j] -= f * A[c + (size_t)n * j];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double v = b[r];
    for (int j = r + 1; j < n; ++j) v -= A[r + (size_t)n * j] * b[j];
    b[r] = v / A[r + (size_t)n * r];
  }
  return 0;
}

/* upper Cholesky factor R (A = R'R) in the upper triangle of A (column-major); 1 when A is not positive definite */
static int chol_upper(int n, double* A) {
  for (int j = 0; j < n; ++j) {
    double d = A[j + (size_t)n * j];
    for (int q = 0; q < j; ++q) d -= A[q + (size_t)n * j] * A[q + (size_t)n * j];
    if (!(d > 0.0)) return 1;
    d = sqrt(d);
    A[j + (size_t)n * j] = d;
    for (int c = j + 1; c < n; ++c) {
      double v = A[j + (size_t)n * c];
      for (int q = 0; q < j; ++q) v -= A[q + (size_t)n * j] * A[q + (size_t)n * c];
      A[j + (size_t)n * c] = v / d;
    }
  }
  return 0;
}
/* R x = b, R upper (column-major) */
static void back_solve_upper(int n, const double* R, double* b) {
  for (int r = n - 1; r >= 0; --r) {
    double v = b[r];
    for (int j = r + 1; j < n; ++j) v -= R[r + (size_t)n * j] * b[j];
    b[r] = v / R[r + (size_t)n * r];
  }
}

/* y, h: T x k column-major (h updated in place); sigma, h_init, constant: k; u, eps: k x T ([series][t]);
 * J = 7 (KSC) or 10 (OCSN); s_out: T x k mixture indicators (may be NULL).  Returns 0, 1 (LU) or 2 (Cholesky). */
int stochvol_ref(int T, int k, const double* y, double* h, const double* sigma, const double* h_init,
                 const double* constant, const double* u, const double* eps, int J, int* s_out) {
  const double* P = (J == 7) ? KSC_P : OCSN_P;
  const double* S2 = (J == 7) ? KSC_S2 : OCSN_S2;
  double mu[10];
  for (int j = 0; j < J; ++j) mu[j] = (J == 7) ? KSC_MU[j] - 1.2704 : OCSN_MU[j];          /* stochvol_ksc1998.cpp:77 */
  double* hh = (double*)calloc((size_t)T * T, sizeof(double));
  double* V = (double*)malloc((size_t)T * T * sizeof(double));
  double* R = (double*)malloc((size_t)T * T * sizeof(double));
  double* ys = (double*)malloc((size_t)T * sizeof(double));
  double* rhs = (double*)malloc((size_t)T * sizeof(double));
  double* w = (double*)malloc((size_t)T * sizeof(double));
  int* s = (int*)malloc((size_t)T * sizeof(int));
  int rc = 0;
  /* hh = H'H with H = I - subdiagonal ones  (:85-87): diag 2,...,2,1; off-diagonals -1 */
  for (int t = 0; t < T; ++t) {
    hh[t + (size_t)T * t] = (t < T - 1) ? 2.0 : 1.0;
    if (t + 1 < T) { hh[t + 1 + (size_t)T * t] = -1.0; hh[t + (size_t)T * (t + 1)] = -1.0; }
  }
  for (int i = 0; i < k && rc == 0; ++i) {
    for (int t = 0; t < T; ++t) {
      const double yv = y[t + (size_t)T * i];
      ys[t] = log(yv * yv + constant[i]);                                                   /* :91 */
      double q[10], qs = 0.0;
      for (int j = 0; j < J; ++j) {
        const double sd = sqrt(S2[j]), z = (ys[t] - (h[t + (size_t)T * i] + mu[j])) / sd;
        q[j] = P[j] * (exp(-0.5 * z * z) / (sd * sqrt(2.0 * 3.14159265358979323846)));      /* :94 */
        qs += q[j];
      }
      double cum = 0.0;
      int cnt = 0;
      for (int j = 0; j < J; ++j) { cum += q[j] / qs; if (u[(size_t)i * T + t] < cum) ++cnt; }   /* :95-96 */
      s[t] = J - cnt;
      if (s[t] > J - 1) s[t] = J - 1;
      if (s_out) s_out[t + (size_t)T * i] = s[t];
    }
    double hsum;
    for (int c = 0; c < T; ++c)
      for (int r = 0; r < T; ++r) V[r + (size_t)T * c] = hh[r + (size_t)T * c] / sigma[i] + ((r == c) ? 1.0 / S2[s[r]] : 0.0);
    for (int r = 0; r < T; ++r) {                                                           /* :103 right-hand side */
      hsum = 0.0;
      for (int c = 0; c < T; ++c) hsum += hh[r + (size_t)T * c] / sigma[i];
      rhs[r] = hsum * h_init[i] + (ys[r] - mu[s[r]]) / S2[s[r]];
    }
    memcpy(R, V, (size_t)T * T * sizeof(double));
    if (lu_solve(T, V, rhs)) { rc = 1; break; }
    if (chol_upper(T, R)) { rc = 2; break; }
    for (int t = 0; t < T; ++t) w[t] = eps[(size_t)i * T + t];
    back_solve_upper(T, R, w);                                                              /* :104 */
    for (int t = 0; t < T; ++t) h[t + (size_t)T * i] = rhs[t] + w[t];
  }
  free(hh); free(V); free(R); free(ys); free(rhs); free(w); free(s);
  return rc;
}

/* y: k x T; z: kT x n; a: n x na (na = 1 or T); lam: n diagonal indicators (updated in place); sigma_i: k x k
 * (sig_rows = k) or kT x k stack (sig_rows = kT); prior: n inclusion probabilities; u1, u2: n uniforms by coefficient id;
 * include: 0-based positions in the order they are visited. */
int bvs_ref(int k, int T, int n, const double* y, const double* z, const double* a, int na, double* lam,
            const double* sigma_i, int sig_rows, const double* prior, const double* u1, const double* u2,
            const int* include, int n_inc) {
  const int kT = k * T;
  double* AG = (double*)malloc((size_t)n * na * sizeof(double));
  double* th = (double*)malloc((size_t)n * na * sizeof(double));
  double* res = (double*)malloc((size_t)kT * sizeof(double));
  for (int c = 0; c < na; ++c)
    for (int r = 0; r < n; ++r) AG[r + (size_t)n * c] = lam[r] * a[r + (size_t)n * c];       /* :111 (computed once) */
  for (int e = 0; e < n_inc; ++e) {
    const int var = include[e];
    const double g = log(u1[var]), lp0 = log(1.0 - prior[var]), lp1 = log(prior[var]);
    if (lam[var] == 1.0 && g >= lp1) continue;                                               /* :127-133 */
    if (lam[var] == 0.0 && g >= lp0) continue;
    double l[2];
    for (int which = 0; which < 2; ++which) {
      memcpy(th, AG, (size_t)n * na * sizeof(double));
      for (int c = 0; c < na; ++c) th[var + (size_t)n * c] = which ? a[var + (size_t)n * c] : 0.0;
      for (int t = 0; t < T; ++t)
        for (int i = 0; i < k; ++i) {
          double v = y[i + (size_t)k * t];
          const double* thc = th + (size_t)n * ((na == 1) ? 0 : t);
          for (int r = 0; r < n; ++r) v -= z[(t * k + i) + (size_t)kT * r] * thc[r];
          res[t * k + i] = v;
        }
      double quad = 0.0;                                    /* res' S_i res with S_i block diagonal (dense in the reference) */
      for (int t = 0; t < T; ++t)
        for (int i = 0; i < k; ++i) {
          double v = 0.0;
          for (int j = 0; j < k; ++j)
            v += ((sig_rows == k) ? sigma_i[i + (size_t)k * j] : sigma_i[(t * k + i) + (size_t)kT * j]) * res[t * k + j];
          quad += res[t * k + i] * v;
        }
      l[which] = -0.5 * quad + (which ? lp1 : lp0);
    }
    lam[var] = ((l[1] - l[0]) >= log(u2[var])) ? 1.0 : 0.0;                                  /* :150-156 */
  }
  free(AG); free(th); free(res);
  return 0;
}

/* y: k x T; z: kT x n (row t k + i = regressors of equation i at t, already multiplied by the BVS mask if any);
 * sig: kT x k stack of Sigma_t^-1; sv_i: n diagonal of Sigma_v^-1; a_init: n; eps: nT; a_out: nT (time-major).
 * Returns 0, 1 (LU) or 2 (Cholesky). */
int tvp_state_ref(int k, int T, int n, const double* y, const double* z, const double* sig, const double* sv_i,
                  const double* a_init, const double* eps, double* a_out) {
  const int N = n * T, kT = k * T;
  double* P = (double*)calloc((size_t)N * N, sizeof(double));
  double* R = (double*)malloc((size_t)N * N * sizeof(double));
  double* rhs = (double*)calloc((size_t)N, sizeof(double));
  double* w = (double*)malloc((size_t)N * sizeof(double));
  double* zs = (double*)malloc((size_t)n * k * sizeof(double));
  int rc = 0;
  for (int t = 0; t < T; ++t) {
    /* H' (I (x) Sv^-1) H: diagonal blocks 2 Sv^-1 (last: Sv^-1), off-diagonal blocks -Sv^-1 */
    for (int r = 0; r < n; ++r) {
      P[(t * n + r) + (size_t)N * (t * n + r)] += ((t < T - 1) ? 2.0 : 1.0) * sv_i[r];
      if (t + 1 < T) {
        P[((t + 1) * n + r) + (size_t)N * (t * n + r)] -= sv_i[r];
        P[(t * n + r) + (size_t)N * ((t + 1) * n + r)] -= sv_i[r];
      }
    }
    /* zz' Su^-1 zz and zz' Su^-1 y of period t */
    for (int r = 0; r < n; ++r)
      for (int j = 0; j < k; ++j) {
        double v = 0.0;
        for (int i = 0; i < k; ++i) v += z[(t * k + i) + (size_t)kT * r] * sig[(t * k + i) + (size_t)kT * j];
        zs[r + (size_t)n * j] = v;
      }
    for (int r = 0; r < n; ++r) {
      for (int c = 0; c < n; ++c) {
        double v = 0.0;
        for (int j = 0; j < k; ++j) v += zs[r + (size_t)n * j] * z[(t * k + j) + (size_t)kT * c];
        P[(t * n + r) + (size_t)N * (t * n + c)] += v;
      }
      double v = 0.0;
      for (int j = 0; j < k; ++j) v += zs[r + (size_t)n * j] * y[j + (size_t)k * t];
      rhs[t * n + r] += v;
    }
  }
  /* a_h_sigmav_i_h * kron(vec_tt, a_init): rows of H'(I (x) Sv^-1)H sum to Sv^-1 in the first block and to 0 elsewhere */
  for (int r = 0; r < n; ++r) rhs[r] += sv_i[r] * a_init[r];
  memcpy(R, P, (size_t)N * N * sizeof(double));
  if (lu_solve(N, P, rhs)) rc = 1;
  if (rc == 0 && chol_upper(N, R)) rc = 2;
  if (rc == 0) {
    memcpy(w, eps, (size_t)N * sizeof(double));
    back_solve_upper(N, R, w);
    for (int e = 0; e < N; ++e) a_out[e] = rhs[e] + w[e];
  }
  free(P); free(R); free(rhs); free(w); free(zs);
  return rc;
}
