/* bvar_dense_ref.c -- TEST INFRASTRUCTURE ONLY: plain-C restatement of the reference's constant-parameter BVAR
 * sweep with the SAME dense data structures the reference materialises, used (a) as a second, independent
 * checker next to the NumPy oracle and (b) as the CPU baseline ("port") timed by bench.py.
 *
 * Follows src/bvaralg.cpp:292-560 literally for the Wishart prior with optional SSVS:
 *   diag_sigma_i = kron(I_T, Sigma^-1)   dense kT x kT                       (:219, :514)
 *   ZtD  = trans(z) * diag_sigma_i       dense GEMM  n x kT x kT             (:305)
 *   P    = V^-1 + ZtD * z,  b = V^-1 mu + ZtD * yvec                          (:305-306)
 *   mean = solve(P, b)       LU with partial pivoting (dgesv)                (:306)
 *   a    = mean + solve(chol(P), eps)   upper Cholesky + back substitution   (:307)
 *   SSVS inclusion update                                                    (:314-331)
 *   u    = yvec - z a                                                        (:364)
 *   Sigma^-1 ~ wishrnd(solve(S0 + u u', I), df)                              (:511)
 *   store a and Sigma = solve(Sigma^-1, I)                                   (:518-559)
 * Plain triple loops stand in for the reference BLAS/LAPACK R links by default (src/Makevars:14).
 * Chains are spread over OpenMP threads the way parallel::mclapply spreads models over cores
 * (R/draw_posterior.bvarmodel.R:57-58).  Variates are injected (parity) or drawn from a private generator.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { uint64_t s[4]; int has; double spare; } rng_t;

static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t next64(rng_t* r) {                      /* xoshiro256** */
  const uint64_t res = rotl(r->s[1] * 5, 7) * 9, t = r->s[1] << 17;
  r->s[2] ^= r->s[0]; r->s[3] ^= r->s[1]; r->s[1] ^= r->s[2]; r->s[0] ^= r->s[3]; r->s[2] ^= t; r->s[3] = rotl(r->s[3], 45);
  return res;
}
static void seed_rng(rng_t* r, uint64_t seed) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < 4; ++i) { z += 0x9E3779B97F4A7C15ull; uint64_t x = z; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; r->s[i] = x ^ (x >> 31); }
  r->has = 0;
}
static double unif(rng_t* r) { return ((double)(next64(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static double normal(rng_t* r) {
  if (r->has) { r->has = 0; return r->spare; }
  const double u1 = unif(r), u2 = unif(r), m = sqrt(-2.0 * log(u1));
  r->spare = m * sin(6.283185307179586 * u2); r->has = 1;
  return m * cos(6.283185307179586 * u2);
}
static double gamma_mt(rng_t* r, double a) {            /* Marsaglia-Tsang, a >= 1 here */
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (;;) {
    double x = normal(r), v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    if (log(unif(r)) < 0.5 * x * x + d - d * v + d * log(v)) return d * v;
  }
}

/* C (m x n) = A' (k x m)' * B (k x n); all column-major */
static void gemm_tn(int m, int n, int k, const double* A, const double* B, double* C) {
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) {
      double s = 0.0;
      const double* a = A + (size_t)k * i; const double* b = B + (size_t)k * j;
      for (int l = 0; l < k; ++l) s += a[l] * b[l];
      C[i + (size_t)m * j] = s;
    }
}
/* C (m x n) = A (m x k) * B (k x n) */
static void gemm_nn(int m, int n, int k, const double* A, const double* B, double* C) {
  for (int j = 0; j < n; ++j) {
    double* c = C + (size_t)m * j;
    for (int i = 0; i < m; ++i) c[i] = 0.0;
    for (int l = 0; l < k; ++l) {
      const double b = B[l + (size_t)k * j];
      if (b == 0.0) continue;            /* the kron factor is block diagonal; dgemm would not skip, see note */
      const double* a = A + (size_t)m * l;
      for (int i = 0; i < m; ++i) c[i] += a[i] * b;
    }
  }
}
/* same without the zero skip: what a dense dgemm executes */
static void gemm_nn_dense(int m, int n, int k, const double* A, const double* B, double* C) {
  for (int j = 0; j < n; ++j) {
    double* c = C + (size_t)m * j;
    for (int i = 0; i < m; ++i) c[i] = 0.0;
    for (int l = 0; l < k; ++l) {
      const double b = B[l + (size_t)k * j];
      const double* a = A + (size_t)m * l;
      for (int i = 0; i < m; ++i) c[i] += a[i] * b;
    }
  }
}
/* solve A x = b by LU with partial pivoting (A n x n col-major destroyed, b overwritten) */
static int lu_solve(int n, double* A, double* b, int nrhs) {
  for (int j = 0; j < n; ++j) {
    int p = j; double mx = fabs(A[j + (size_t)n * j]);
    for (int i = j + 1; i < n; ++i) if (fabs(A[i + (size_t)n * j]) > mx) { mx = fabs(A[i + (size_t)n * j]); p = i; }
    if (mx == 0.0) return 1;
    if (p != j) {
      for (int c = 0; c < n; ++c) { double t = A[j + (size_t)n * c]; A[j + (size_t)n * c] = A[p + (size_t)n * c]; A[p + (size_t)n * c] = t; }
      for (int c = 0; c < nrhs; ++c) { double t = b[j + (size_t)n * c]; b[j + (size_t)n * c] = b[p + (size_t)n * c]; b[p + (size_t)n * c] = t; }
    }
    const double piv = A[j + (size_t)n * j];
    for (int i = j + 1; i < n; ++i) A[i + (size_t)n * j] /= piv;
    for (int c = j + 1; c < n; ++c) {
      const double f = A[j + (size_t)n * c];
      for (int i = j + 1; i < n; ++i) A[i + (size_t)n * c] -= A[i + (size_t)n * j] * f;
    }
    for (int c = 0; c < nrhs; ++c) {
      const double f = b[j + (size_t)n * c];
      for (int i = j + 1; i < n; ++i) b[i + (size_t)n * c] -= A[i + (size_t)n * j] * f;
    }
  }
  for (int c = 0; c < nrhs; ++c)
    for (int i = n - 1; i >= 0; --i) {
      double s = b[i + (size_t)n * c];
      for (int l = i + 1; l < n; ++l) s -= A[i + (size_t)n * l] * b[l + (size_t)n * c];
      b[i + (size_t)n * c] = s / A[i + (size_t)n * i];
    }
  return 0;
}
/* upper Cholesky R'R = A in place (upper triangle of A col-major); returns non-zero if not SPD */
static int chol_upper(int n, double* A) {
  for (int j = 0; j < n; ++j) {
    double d = A[j + (size_t)n * j];
    for (int l = 0; l < j; ++l) d -= A[l + (size_t)n * j] * A[l + (size_t)n * j];
    if (!(d > 0.0)) return 1;
    d = sqrt(d);
    A[j + (size_t)n * j] = d;
    for (int c = j + 1; c < n; ++c) {
      double s = A[j + (size_t)n * c];
      for (int l = 0; l < j; ++l) s -= A[l + (size_t)n * j] * A[l + (size_t)n * c];
      A[j + (size_t)n * c] = s / d;
    }
    for (int i = j + 1; i < n; ++i) A[i + (size_t)n * j] = 0.0;
  }
  return 0;
}
static void backsolve_upper(int n, const double* R, double* x) {   /* R x = b in place */
  for (int i = n - 1; i >= 0; --i) {
    double s = x[i];
    for (int l = i + 1; l < n; ++l) s -= R[i + (size_t)n * l] * x[l];
    x[i] = s / R[i + (size_t)n * i];
  }
}

/* returns 0 ok, 1 allocation failure, 2 numerical failure in some chain */
int bvar_dense_ref(int k, int T, int M, const double* Y, const double* Z, const double* mu, const double* Vi,
                   double df_post, const double* S0, const double* sigma_i0, int ssvs, const double* tau0,
                   const double* tau1, const double* inprior, const int* include, int n_include, int iterations,
                   int burnin, int n_chains, uint64_t seed, int n_threads, int dense_gemm, const double* inj_coef,
                   const double* inj_chi2, const double* inj_wn, const double* inj_ssvs_u, double* out_coef,
                   double* out_sigma, double* out_lambda) {
  const int n = k * M, kT = k * T, sweeps = iterations + burnin, nw = k * (k - 1) / 2;
  int status = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int ch = 0; ch < n_chains; ++ch) {
    rng_t rng; seed_rng(&rng, seed + 0x1234567ull * (uint64_t)ch);
    double* D = calloc((size_t)kT * kT, 8);         /* kron(I_T, Sigma^-1) */
    double* ZtD = malloc((size_t)n * kT * 8);
    double* P = malloc((size_t)n * n * 8);
    double* Pc = malloc((size_t)n * n * 8);
    double* b = malloc((size_t)n * 8), *a = malloc((size_t)n * 8), *eps = malloc((size_t)n * 8);
    double* vi = malloc((size_t)n * n * 8);
    double* lam = malloc((size_t)n * 8);
    double* u = malloc((size_t)kT * 8);
    double* sig = malloc((size_t)k * k * 8), *S = malloc((size_t)k * k * 8), *I = malloc((size_t)k * k * 8);
    double* A = malloc((size_t)k * k * 8), *AD = malloc((size_t)k * k * 8), *tmp = malloc((size_t)k * k * 8);
    double* Zt = malloc((size_t)n * kT * 8);        /* trans(z) materialised once (n x kT) */
    if (!D || !ZtD || !P || !Pc || !b || !a || !eps || !vi || !lam || !u || !sig || !S || !I || !A || !AD || !tmp || !Zt) {
#pragma omp critical
      status = 1;
    } else {
      memcpy(vi, Vi, (size_t)n * n * 8);
      for (int r = 0; r < n; ++r) lam[r] = 1.0;
      for (int r = 0; r < kT; ++r) for (int c = 0; c < n; ++c) Zt[c + (size_t)n * r] = Z[r + (size_t)kT * c];
      /* initial diag_sigma_i: only the diagonal of sigma_i (:246) */
      for (int t = 0; t < T; ++t) for (int i = 0; i < k; ++i) D[(t * k + i) + (size_t)kT * (t * k + i)] = sigma_i0[i + k * i];
      for (int draw = 0; draw < sweeps; ++draw) {
        /* ZtD = trans(z) * D  (n x kT) */
        if (dense_gemm) gemm_nn_dense(n, kT, kT, Zt, D, ZtD); else gemm_nn(n, kT, kT, Zt, D, ZtD);
        /* P = vi + ZtD * z ;  b = vi mu + ZtD yvec */
        for (int c = 0; c < n; ++c)
          for (int r = 0; r < n; ++r) {
            double s = 0.0;
            for (int l = 0; l < kT; ++l) s += ZtD[r + (size_t)n * l] * Z[l + (size_t)kT * c];
            P[r + (size_t)n * c] = vi[r + (size_t)n * c] + s;
          }
        for (int r = 0; r < n; ++r) {
          double s = 0.0;
          for (int c = 0; c < n; ++c) s += vi[r + (size_t)n * c] * mu[c];
          for (int l = 0; l < kT; ++l) s += ZtD[r + (size_t)n * l] * Y[l];
          b[r] = s;
        }
        memcpy(Pc, P, (size_t)n * n * 8);
        if (lu_solve(n, Pc, b, 1)) status = 2;                       /* post_mu */
        memcpy(Pc, P, (size_t)n * n * 8);
        if (chol_upper(n, Pc)) status = 2;
        for (int r = 0; r < n; ++r)
          eps[r] = inj_coef ? inj_coef[((size_t)ch * sweeps + draw) * n + r] : normal(&rng);
        backsolve_upper(n, Pc, eps);
        for (int r = 0; r < n; ++r) a[r] = b[r] + eps[r];
        if (ssvs) {
          for (int q = 0; q < n_include; ++q) {
            const int r = include[q];
            const double t0 = tau0[r], t1 = tau1[r], t0sq = t0 * t0, t1sq = t1 * t1;
            const double u0 = 1.0 / t0 * exp(-((a[r] * a[r]) / (2.0 * t0sq))) * (1.0 - inprior[r]);
            const double u1 = 1.0 / t1 * exp(-((a[r] * a[r]) / (2.0 * t1sq))) * inprior[r];
            const double p = u1 / (u0 + u1);
            const double uu = inj_ssvs_u ? inj_ssvs_u[((size_t)ch * sweeps + draw) * n + r] : unif(&rng);
            const double ld = (p > 0.5) ? (uu < p ? 1.0 : 0.0) : (uu >= 1.0 - p ? 1.0 : 0.0);
            lam[r] = ld;
            vi[r + (size_t)n * r] = (ld == 0.0) ? 1.0 / t0sq : 1.0 / t1sq;
          }
        }
        /* u = yvec - z a */
        for (int l = 0; l < kT; ++l) {
          double s = Y[l];
          for (int c = 0; c < n; ++c) s -= Z[l + (size_t)kT * c] * a[c];
          u[l] = s;
        }
        /* S = S0 + u u' ; Sbar = solve(S, I) */
        for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) {
          double s = S0[i + k * j];
          for (int t = 0; t < T; ++t) s += u[i + k * t] * u[j + k * t];
          S[i + k * j] = s; I[i + k * j] = (i == j) ? 1.0 : 0.0;
        }
        if (lu_solve(k, S, I, k)) status = 2;                        /* I <- S^-1 */
        /* wishrnd: D_w = chol(Sbar) upper; A Bartlett; Sigma^-1 = (A D_w)'(A D_w) */
        if (chol_upper(k, I)) status = 2;
        memset(A, 0, (size_t)k * k * 8);
        for (int i = 0; i < k; ++i) {
          const double c2 = inj_chi2 ? inj_chi2[((size_t)ch * sweeps + draw) * k + i] : 2.0 * gamma_mt(&rng, 0.5 * (df_post - i));
          A[i + k * i] = sqrt(c2);
        }
        { int c = 0; for (int i = 1; i < k; ++i) for (int j = 0; j < i; ++j, ++c)
            A[j + k * i] = inj_wn ? inj_wn[((size_t)ch * sweeps + draw) * nw + c] : normal(&rng); }
        gemm_nn_dense(k, k, k, A, I, AD);
        gemm_tn(k, k, k, AD, AD, sig);
        /* diag_sigma_i = kron(I_T, sigma_i) */
        for (int t = 0; t < T; ++t) for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j)
          D[(t * k + i) + (size_t)kT * (t * k + j)] = sig[i + k * j];
        if (draw >= burnin) {
          const size_t pos = (size_t)ch * iterations + (draw - burnin);
          memcpy(tmp, sig, (size_t)k * k * 8);
          for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) I[i + k * j] = (i == j) ? 1.0 : 0.0;
          if (lu_solve(k, tmp, I, k)) status = 2;
          memcpy(out_sigma + pos * k * k, I, (size_t)k * k * 8);
          memcpy(out_coef + pos * n, a, (size_t)n * 8);
          if (out_lambda) memcpy(out_lambda + pos * n, lam, (size_t)n * 8);
        }
      }
    }
    free(D); free(ZtD); free(P); free(Pc); free(b); free(a); free(eps); free(vi); free(lam); free(u);
    free(sig); free(S); free(I); free(A); free(AD); free(tmp); free(Zt);
  }
  return status;
}
