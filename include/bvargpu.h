/* bvargpu.h -- C ABI of libbvargpu.so, the B200-native Gibbs-sampler hot path behind bvartools'
 * draw_posterior() / bvarpost().
 *
 * Every entry point replaces one FFI entry of the reference (file:line into franzmohr/bvartools):
 *
 *   bvg_bvaralg          <-  RcppExport SEXP _bvartools_bvaralg(SEXP)      src/RcppExports.cpp:19-27   (R: .bvaralg,    R/RcppExports.R:4-6)
 *   bvg_bvartvpalg       <-  RcppExport SEXP _bvartools_bvartvpalg(SEXP)   src/RcppExports.cpp:30-38   (R: .bvartvpalg, R/RcppExports.R:8-10)
 *   bvg_post_normal      <-  _bvartools_post_normal       (5 args)          src/RcppExports.cpp:318-332
 *   bvg_post_normal_sur  <-  _bvartools_post_normal_sur   (6 args)          src/RcppExports.cpp:333-348
 *   bvg_ssvs             <-  _bvartools_ssvs              (5 args)          src/RcppExports.cpp:477-490
 *   bvg_bvs              <-  _bvartools_bvs               (7 args)          src/RcppExports.cpp:62-77
 *   bvg_stochvol         <-  _bvartools_stochvol_ksc1998 / _ocsn2007 / _stoch_vol (5 args)   src/RcppExports.cpp:349-463
 *   bvg_kalman_dk        <-  _bvartools_kalman_dk         (7 args)          src/RcppExports.cpp:196-234
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary: every function returns 0 on success, < 0 for an invalid
 *     argument / unsupported model, > 0 for a CUDA error (the cudaError_t value); bvg_last_error() returns the
 *     message of the last failure on the calling thread (the glue maps it to Rcpp::stop()).
 *   - all matrices are FP64, column-major (Armadillo / R layout); the caller owns every buffer it passes and
 *     allocates the outputs; device memory is owned by the library.
 *   - many independent chains per call ("batch"); random variates are Philox4x32-10 keyed by
 *     (seed, GLOBAL chain id, sweep, site), so results do not depend on thread mapping, launch chunking or on
 *     how chains are sharded over GPUs / ranks (chain_offset).
 *   - no CPU fallback: without a CUDA device every compute entry point fails with a message.
 */
#ifndef BVARGPU_H
#define BVARGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVG_ABI_VERSION 2

/* error-variance prior (priors$sigma, R/add_priors.bvarmodel.R:521-569) */
#define BVG_SIGMA_WISHART 0
#define BVG_SIGMA_GAMMA 1
#define BVG_SIGMA_SV 2
/* variable selection (priors$coefficients$ssvs / $bvs) */
#define BVG_VARSEL_NONE 0
#define BVG_VARSEL_SSVS 1
#define BVG_VARSEL_BVS 2
/* residual evaluation: 0 = library decides */
#define BVG_RESID_AUTO 0
#define BVG_RESID_DIRECT 1 /* u_t = y_t - A x_t, O(T) per sweep                    (bvaralg.cpp:364)      */
#define BVG_RESID_MOMENT 2 /* u u' = SSE_ols + (A-A_ols) XX' (A-A_ols)', T-free     (needs X of full rank) */
/* mixture table of the SV step */
#define BVG_MIX_KSC1998 7   /* src/stochvol_ksc1998.cpp:74-78 */
#define BVG_MIX_OCSN2007 10 /* src/stochvol_ocsn2007.cpp:74-84 */

/* Injected random variates (tier-1 parity tests).  NULL = generate with Philox on the device.  Each array is
 * [chain][sweep][count] with the counts of oracle/variates.py; sweeps = iterations + burnin. */
typedef struct bvg_inject {
  const double* coef_normal;    /* n            */
  const double* ssvs_u;         /* n            */
  const double* bvs_u1;         /* n            */
  const double* bvs_u2;         /* n            */
  const double* sv_u;           /* k*T  ([series][t]) */
  const double* sv_normal;      /* k*T          */
  const double* sigma_h_gamma;  /* k            */
  const double* h_init_normal;  /* k            */
  const double* omega_gamma;    /* k            */
  const double* wishart_chi2;   /* k            */
  const double* wishart_normal; /* k(k-1)/2     */
  const double* state_normal;   /* n*T          */
  const double* sigma_v_gamma;  /* n            */
  const double* a_init_normal;  /* n            */
  const double* psi_normal;     /* k(k-1)/2  covariance (psi) block, bvaralg.cpp:394; TVP: k(k-1)/2 * T, period-major, bvartvpalg.cpp:363 */
  const double* psi_u1;         /* k(k-1)/2  SSVS Bernoulli / BVS prior gate on psi, bvaralg.cpp:409,424 */
  const double* psi_u2;         /* k(k-1)/2  BVS acceptance on psi, bvaralg.cpp:437 */
  const double* psi_sigma_v_gamma; /* k(k-1)/2  Gamma(shape + T/2, 1), TVP psi state precisions, bvartvpalg.cpp:492 */
  const double* psi_init_normal;   /* k(k-1)/2  TVP psi initial state, bvartvpalg.cpp:498 */
} bvg_inject;

/* The flat POD the Rcpp glue fills from the `bvarmodel` list (what src/bvaralg.cpp:9-249 and
 * src/bvartvpalg.cpp:11-279 unpack by name).  Pure VAR regressors: SUR = kron(Z, I_k) (R/gen_var.R:281). */
typedef struct bvg_spec {
  int32_t abi_version; /* BVG_ABI_VERSION */
  int32_t device;      /* CUDA device ordinal, -1 = current */
  /* dimensions */
  int32_t k; /* endogenous variables                                       */
  int32_t T; /* observations                                               */
  int32_t M; /* regressors per equation (k p + exogenous + deterministic)  */
  int32_t tvp;
  int32_t iterations, burnin;
  /* batch */
  int64_t n_chains;     /* chains computed by this call                    */
  int64_t chain_offset; /* global id of the first one (RNG key)            */
  uint64_t seed;
  /* data: data$Y (T x k) and data$Z (T x M) as the samplers use them, transposed */
  const double* Y; /* k x T  (y = t(Y), bvaralg.cpp:10)               */
  const double* X; /* M x T  (x = t(Z)); may be NULL when M == 0      */
  /* priors$coefficients */
  const double* mu;      /* n = k M                                        */
  const double* Vi;      /* V^-1: n (diagonal) or n x n, see vi_dense     */
  int32_t vi_dense;
  int32_t varsel;        /* BVG_VARSEL_*                                   */
  const double* tau0;    /* n, SSVS                                        */
  const double* tau1;    /* n, SSVS                                        */
  const double* inprior; /* n, SSVS / BVS prior inclusion probability      */
  const int32_t* include; /* 0-based positions taking part (R's include - 1) */
  int32_t n_include;
  /* priors$sigma + initial$sigma */
  int32_t sigma_type;        /* BVG_SIGMA_*                                */
  double wishart_df;         /* prior df (posterior df = df + T)           */
  const double* wishart_scale; /* k x k                                    */
  const double* sigma_shape; /* k   gamma / SV prior shape                 */
  const double* sigma_rate;  /* k                                          */
  const double* sv_mu;       /* k                                          */
  const double* sv_vi;       /* k x k                                      */
  const double* sv_sigma_h;  /* k   initial$sigma$sigma_h                  */
  const double* sv_constant; /* k   initial$sigma$constant                 */
  const double* sv_h;        /* T x k initial$sigma$h                      */
  int32_t sv_mixture;        /* BVG_MIX_*                                  */
  const double* sigma_i;     /* k x k initial$sigma$sigma_i (non-SV)       */
  /* TVP (priors$coefficients$shape/rate, initial$coefficients$draw) */
  const double* tvp_shape; /* n */
  const double* tvp_rate;  /* n */
  const double* a_init;    /* n */
  /* priors$psi: covariance block (bvaralg.cpp:157-162,373-455; TVP models: bvartvpalg.cpp:143-176,339-407); NULL = none.
   * Needs a gamma or SV error prior.  sigma output of a TVP model with this block: [T][k*k] per draw. */
  const double* psi_mu;    /* k(k-1)/2 */
  const double* psi_vi;    /* k(k-1)/2 x k(k-1)/2, column-major */
  /* priors$psi$ssvs / $bvs: variable selection on the covariances (bvaralg.cpp:164-202,396-449); the inclusion draws
   * are appended to the `lambda` output: rows [n (if coefficient selection is on) | k(k-1)/2] */
  int32_t psi_varsel;          /* BVG_VARSEL_* */
  const double* psi_tau0;      /* k(k-1)/2, SSVS */
  const double* psi_tau1;      /* k(k-1)/2, SSVS */
  const double* psi_inprior;   /* k(k-1)/2 */
  const int32_t* psi_include;  /* 0-based positions taking part */
  int32_t psi_n_include;
  /* TVP models (bvartvpalg.cpp:143-176,339-407,485-499): the free elements of Psi_t follow random walks; inverse-gamma
   * prior of their state variances and the initial draw (initial$psi$draw).  NULL for constant-parameter models.
   * BVS on the time-varying covariances (psi_varsel = BVG_VARSEL_BVS, bvartvpalg.cpp:365-398) is supported; SSVS is not
   * defined for TVP models (-7). */
  const double* psi_shape;     /* k(k-1)/2 */
  const double* psi_rate;      /* k(k-1)/2 */
  const double* psi_init;      /* k(k-1)/2 */
  /* model$structural (R/gen_var.R:243-251): equation i also contains -y_jt, j < i (the A0 columns of data$SUR).  With
   * structural = 1 every per-coefficient array (mu, Vi, inprior, include; the coef / lambda outputs; injected coef_normal
   * excepted, see below) has k M + k(k-1)/2 entries in the reference's order [a | b | c | a0], a0 by variable j then
   * equation i > j.  X stays the M x T regressor matrix WITHOUT the -y block (the library appends it).  Needs a gamma or SV
   * error prior (R/add_priors.bvarmodel.R); the covariance block is not available with it (-7); SSVS works on the kept
   * positions (tau0 / tau1 / inprior / include in the same [a | b | c | a0] order).  TVP models:
   * tvp_shape / tvp_rate / a_init and the sigma_v output follow the same order and length.
   * Injected coefficient-level variates (tests) are indexed by the INTERNAL position i + k m over the M + k regressors. */
  int32_t structural;
  /* tuning / testing */
  int32_t resid_mode;       /* BVG_RESID_*                                 */
  int32_t threads_per_chain; /* 0 = auto; 1,2,4,8,16,32 (sub-warp tiles) or a multiple of 32 (CTA per chain) */
  int32_t sweeps_per_launch; /* 0 = auto: kernel launches are chunked so that D2H copies overlap compute */
  /* bvg_bvaralg / bvg_bvartvpalg only: shard the n_chains chains over this many GPUs of the box (0, 1 = one device;
   * -1 = all devices).  Contiguous blocks of GLOBAL chain ids, one host thread per shard, shard i on device
   * (device + i) mod device count, results straight into the caller's buffers: bit-identical to n_gpus = 1 (the RNG is
   * keyed by the global chain id).  The interrupt callback is polled from the CALLING thread only.  What
   * parallel::mclapply(mc.cores) is to the reference (R/draw_posterior.bvarmodel.R:57-58), without forking a CUDA
   * context.  The handle API drives one device (-54 otherwise). */
  int32_t n_gpus;
  const bvg_inject* inject;  /* NULL = Philox                              */
} bvg_spec;

/* Output buffers (caller-allocated; host for bvg_bvaralg / bvg_sampler_fetch).  Per chain a column-major
 * rows x iterations matrix -- exactly the reference's draws_* matrices (bvaralg.cpp:265-275,
 * bvartvpalg.cpp:249-265); the glue slices rows into a / b / c.  NULL = not wanted. */
typedef struct bvg_out {
  double* coef;    /* [chain][iter][n]     (TVP: [chain][iter][T][n], time-major like vectorise(a_mat)) */
  double* sigma;   /* [chain][iter][k*k]   (SV: [chain][iter][T][k*k])                                   */
  double* lambda;  /* [chain][iter][n]     inclusion indicators (SSVS / BVS)                             */
  double* sigma_h; /* [chain][iter][k]     SV: sigma_h (the reference stores diag(sigma_h) as k*k)       */
  double* sigma_v; /* [chain][iter][n]     TVP: state variances 1/diag(Sigma_v^-1)                       */
  int32_t* status; /* [chain]              0 ok, 1 = a Cholesky pivot was not positive (chain continues) */
} bvg_out;

typedef struct bvg_sampler bvg_sampler; /* opaque: validated spec + device buffers */

/* library */
const char* bvg_version(void);
const char* bvg_last_error(void);
int bvg_device_count(void);
void* bvg_host_alloc(size_t bytes); /* pinned host memory for outputs (NULL on failure) */
void bvg_host_free(void* p);

/* sizes of the output rows per retained draw, so the caller can allocate */
int bvg_out_rows(const bvg_spec* spec, int64_t* coef_rows, int64_t* sigma_rows, int64_t* lambda_rows,
                 int64_t* sigma_h_rows, int64_t* sigma_v_rows);

/* one-shot drop-ins for .bvaralg / .bvartvpalg: host in, host out */
int bvg_bvaralg(const bvg_spec* spec, const bvg_out* out);
int bvg_bvartvpalg(const bvg_spec* spec, const bvg_out* out);

/* handle API: inputs uploaded once, draws stay resident in HBM until fetched */
int bvg_sampler_create(const bvg_spec* spec, bvg_sampler** sampler);
int bvg_sampler_reset(bvg_sampler* s);                    /* back to the initial state (re-run the same job) */
int bvg_sampler_run(bvg_sampler* s, void* cuda_stream);   /* all sweeps, asynchronous on the given stream    */
int bvg_sampler_run_to_host(bvg_sampler* s, const bvg_out* host_out, void* cuda_stream); /* chunked, D2H overlapped; returns after completion */
int bvg_sampler_fetch(bvg_sampler* s, const bvg_out* host_out); /* synchronous D2H of everything            */
int bvg_sampler_device_out(bvg_sampler* s, bvg_out* dev_out);   /* device pointers of the draw buffers       */
int bvg_sampler_sync(bvg_sampler* s);
double bvg_sampler_last_kernel_ms(bvg_sampler* s);        /* CUDA-event time of the sweep kernels of the last run */
int64_t bvg_sampler_launch_count(bvg_sampler* s);         /* kernels launched by the last run                 */
const char* bvg_sampler_kernel_name(bvg_sampler* s);      /* which sweep kernel the plan selected, e.g. "kron_k3_g8" */
void bvg_sampler_destroy(bvg_sampler* s);

/* set a callback polled between kernel batches (Rcpp::checkUserInterrupt analogue, bvaralg.cpp:294-296);
 * a non-zero return aborts the run with error code -100 */
void bvg_set_interrupt_poll(int (*poll)(void));

/* on-device posterior moments of a draw buffer [chains][iters][rows]: sums[r] = sum x, sumsq[r] = sum x^2
 * (inputs to an NCCL all-reduce over ranks); device pointers */
int bvg_moments(const double* dev_draws, int64_t chains, int64_t iters, int64_t rows, double* dev_sums,
                double* dev_sumsq, void* cuda_stream);

/* ---- model comparison: log-likelihood of posterior draws ---------------------------------------
 * Replaces the per-draw R loop of summary.bvarlist (R/summary.bvarlist.R:160-186), which calls
 * loglik_normal (src/loglik_normal.cpp:37-58, R/RcppExports.R:313) once per draw:
 *   LL[t, j] = -k/2 ln(2 pi) - ln|Sigma_t^(j)|/2 - u_t' (Sigma_t^(j))^-1 u_t / 2,   u_t = y_t - A_t^(j) x_t .
 * dev_ll_t[t] (T doubles) receives the SUM over the n_draws draws (the caller divides by the number of draws; ranks
 * all-reduce the sums); ll = sum_t dev_ll_t[t] / n_draws, AIC = 2 M - 2 ll, BIC = ln(T) M - 2 ll, HQ = 2 ln ln(T) M - 2 ll
 * (tot_pars = NCOL(x) = M, R/summary.bvarlist.R:62,181-184).
 * Layouts as produced by the samplers: dev_coef [draw][n] (tvp: [draw][T][n]), dev_sigma [draw][k*k] (tv_sigma:
 * [draw][T][k*k]), dev_Y k x T, dev_X M x T column-major; all device pointers.  1 <= k <= 32. */
int bvg_loglik_draws(const double* dev_coef, const double* dev_sigma, int64_t n_draws, int32_t k, int32_t M, int32_t T,
                     int32_t tvp, int32_t tv_sigma, const double* dev_Y, const double* dev_X, double* dev_ll_t,
                     void* cuda_stream);
/* the same with every buffer on the HOST (draws of a finished `bvar` object, as summary.bvarlist receives them): the
 * draws pass through the device in chunks; ll_t: T doubles on the host */
int bvg_loglik_draws_host(const double* coef, const double* sigma, int64_t n_draws, int32_t k, int32_t M, int32_t T,
                          int32_t tvp, int32_t tv_sigma, const double* Y, const double* X, double* ll_t);
/* the same over all retained draws resident in a sampler (chains x iterations); host_ll_t: T doubles on the host */
int bvg_sampler_loglik(bvg_sampler* s, double* host_ll_t, void* cuda_stream);
/* TOTAL log-likelihood sum_draws sum_t LL[t, draw] -- all the information criteria need.  For constant-parameter,
 * constant-variance models this does not loop over the periods: sum_t u_t' Sigma^-1 u_t = tr(Sigma^-1 S) with
 * S = SSE_ols + (A - A_ols) XX' (A - A_ols)', an HBM-bound pass over the draws; other models (and a rank-deficient X)
 * fall back to the per-period kernel.  *total on the host. */
int bvg_sampler_loglik_total(bvg_sampler* s, double* total, void* cuda_stream);
/* Same sum written to DEVICE memory (dev_total: one double), asynchronously on cuda_stream: no host synchronisation, so
 * the sums of a whole model grid can be all-reduced in one collective (constant-parameter models with T > M only: -6
 * otherwise). */
int bvg_sampler_loglik_total_device(bvg_sampler* s, double* dev_total, void* cuda_stream);
int bvg_loglik_total_host(const double* coef, const double* sigma, int64_t n_draws, int32_t k, int32_t M, int32_t T,
                          const double* Y, const double* X, double* total);
/* stand-alone block, host buffers: loglik_normal(u, sigma) of src/loglik_normal.cpp:37-58 for n_batch problems.
 * u: k x T per problem; sigma: k x k (sigma_rows = k) or the (k T) x k stack of per-period matrices (sigma_rows = k T),
 * column-major as in R; out: T per problem.  NaN in u -> -20 (the R wrapper's stop()). */
int bvg_loglik_normal(const double* u, const double* sigma, int32_t k, int32_t T, int32_t sigma_rows, int64_t n_batch,
                      double* out);

/* ---- posterior summaries of resident draws -------------------------------------------------------
 * What summary.bvar takes from coda::summary.mcmc for every coefficient (R/summary.bvar.R:121-140): Mean, SD (naive SE =
 * SD / sqrt(N) follows) and the quantiles, R's default type 7 -- exact order statistics by an 8-pass radix select on the
 * device -- pooled over all n_draws draws of the buffer [n_draws][rows] (device pointer).  Inclusion frequencies are
 * the means of the lambda buffer.  mean, sd: rows doubles; quantiles: [n_probs][rows]; host pointers (mean / sd may be
 * NULL).  The time-series SE is a separate call: bvg_tsse_draws. */
int bvg_summary_draws(const double* dev_draws, int64_t n_draws, int64_t rows, const double* probs, int32_t n_probs,
                      double* mean, double* sd, double* quantiles, void* cuda_stream);
/* the same for an output of a sampler: which = 0 coef, 1 sigma, 2 lambda, 3 sigma_h, 4 sigma_v */
int bvg_sampler_summary(bvg_sampler* s, int32_t which, const double* probs, int32_t n_probs, double* mean, double* sd,
                        double* quantiles, void* cuda_stream);

/* Time-series SE of coda::summary.mcmc (the column summary.bvar copies at R/summary.bvar.R:133,141,195,203): coda's
 * spectrum0.ar -- Yule-Walker AR fit with AIC order selection (stats::ar.yw, order.max = min(n - 1, floor(10 log10 n)))
 * per chain and row, spec = var.pred / (1 - sum ar)^2 (0 when the residuals of lm(x ~ time) vanish) -- and, as in
 * summary.mcmc.list, tsse[row] = sqrt(mean_chains spec / (iters n_chains)).  dev_draws: [n_chains][iters][rows] on the
 * device; tsse: rows doubles on the host.  coda is not vendored by the reference: see oracle/blocks.py spectrum0_ar. */
int bvg_tsse_draws(const double* dev_draws, int64_t n_chains, int64_t iters, int64_t rows, double* tsse,
                   void* cuda_stream);
/* the same for an output of a sampler (which as in bvg_sampler_summary) */
int bvg_sampler_tsse(bvg_sampler* s, int32_t which, double* tsse, void* cuda_stream);

/* ---- impulse responses per draw --------------------------------------------------------------------
 * Replaces the per-draw loop of irf.bvar (R/irf.bvar.R:140-216) around .ir (src/ir.cpp:4-72): for every draw
 * theta_i = (Phi_i P)[response, impulse], i = 0..n_ahead, Phi_i = sum_j Phi_{i-j} A_j.
 * type: 0 feir, 1 oir, 2 gir (sir / sgir: bvg_irf_draws_a0); impulse / response 0-based;
 * shock_kind: 0 = the number `shock`, 1 = "sd", 2 = "nsd" (R/irf.bvar.R:176-188); cumulative as in :204-206.
 * dev_coef: draw d starts at dev_coef + d * coef_stride, its first k*k*p doubles are A = [A_1 .. A_p] column-major;
 * dev_sigma likewise (k x k).  dev_out: [n_draws][n_ahead + 1] on the device. */
int bvg_irf_draws(const double* dev_coef, int64_t coef_stride, const double* dev_sigma, int64_t sigma_stride,
                  int64_t n_draws, int32_t k, int32_t p, int32_t n_ahead, int32_t type, int32_t impulse, int32_t response,
                  int32_t shock_kind, double shock, int32_t cumulative, double* dev_out, void* cuda_stream);
/* the same with the structural types (src/ir.cpp:35-37,42-46): type 3 sir P = A0^-1 shock, 4 sgir P = A0^-1 Sigma / Sigma_jj
 * shock.  dev_a0: draw d starts at dev_a0 + d * a0_stride; a0_compact = 0: the full k x k matrix, column-major (k <= 16);
 * a0_compact = 1: the k(k-1)/2 free elements of the unit lower-triangular A0 by column -- the a0 rows the structural
 * sampler stores at the end of every coefficient draw.  Types 0-2 ignore dev_a0. */
int bvg_irf_draws_a0(const double* dev_coef, int64_t coef_stride, const double* dev_sigma, int64_t sigma_stride,
                     const double* dev_a0, int64_t a0_stride, int32_t a0_compact, int64_t n_draws, int32_t k, int32_t p,
                     int32_t n_ahead, int32_t type, int32_t impulse, int32_t response, int32_t shock_kind, double shock,
                     int32_t cumulative, double* dev_out, void* cuda_stream);
/* on the retained draws of a sampler + the bands of irf.bvar (type-7 quantiles over all draws, :208-213): quantiles
 * [n_probs][n_ahead + 1], mean [n_ahead + 1] (may be NULL), host_draws [draws][n_ahead + 1] (keep_draws; may be NULL) on
 * the host.  period: 0-based period for TVP / SV draws, < 0 = the last one (:127-133). */
int bvg_sampler_irf(bvg_sampler* s, int32_t p, int32_t n_ahead, int32_t type, int32_t impulse, int32_t response,
                    int32_t shock_kind, double shock, int32_t cumulative, int32_t period, const double* probs,
                    int32_t n_probs, double* quantiles, double* mean, double* host_draws, void* cuda_stream);

/* ---- forecast error variance decomposition per draw -----------------------------------------------
 * Replaces the per-draw loop of fevd.bvar (R/fevd.bvar.R:128-165) around .vardecomp (src/vardecomp.cpp:4-87) and its
 * rowMeans over the draws: host_out is the (n_ahead + 1) x k table (row i = horizon, column c = shock variable; row-major),
 * the MEAN over all draws.  type: 1 oir, 2 gir (sir / sgir: bvg_fevd_draws_a0); response 0-based; normalise_gir as :166-170.
 * Buffer layouts as in bvg_irf_draws. */
int bvg_fevd_draws(const double* dev_coef, int64_t coef_stride, const double* dev_sigma, int64_t sigma_stride,
                   int64_t n_draws, int32_t k, int32_t p, int32_t n_ahead, int32_t type, int32_t response,
                   int32_t normalise_gir, double* host_out, void* cuda_stream);
/* the same with the structural types (src/vardecomp.cpp:32-36,41-45): type 3 sir, 4 sgir; dev_a0 as in bvg_irf_draws_a0 */
int bvg_fevd_draws_a0(const double* dev_coef, int64_t coef_stride, const double* dev_sigma, int64_t sigma_stride,
                      const double* dev_a0, int64_t a0_stride, int32_t a0_compact, int64_t n_draws, int32_t k, int32_t p,
                      int32_t n_ahead, int32_t type, int32_t response, int32_t normalise_gir, double* host_out,
                      void* cuda_stream);
int bvg_sampler_fevd(bvg_sampler* s, int32_t p, int32_t n_ahead, int32_t type, int32_t response, int32_t normalise_gir,
                     int32_t period, double* host_out, void* cuda_stream);

/* ---- forecasts per draw -----------------------------------------------------------------------------
 * Replaces the per-draw loop of predict.bvar (R/predict.bvar.R:203-217) around .draw_forecast
 * (src/draw_forecast.cpp:6-51): y_{T+j+1} = A pred_j + Sigma^{1/2} eps_j with the symmetric square root, lags shifted.
 * pred: the reference's prediction matrix, tot x (n_ahead + 1) column-major on the HOST (rows: k*p lags of y | exogenous
 * | deterministic; column 0 = data at T, later columns = future exogenous / deterministic values, :176-201).  The first
 * k * (tot, or tot - k when p = 0) doubles of a coefficient draw are A = cbind(A, B, C), column-major.  dev_eps: injected
 * N(0,1) [n_draws][n_ahead][k] on the device, or NULL = Philox4x32-10 keyed (seed; draw, period, variable).
 * dev_out: [n_draws][n_ahead][k] on the device. */
int bvg_forecast_draws(const double* dev_coef, int64_t coef_stride, const double* dev_sigma, int64_t sigma_stride,
                       int64_t n_draws, int32_t k, int32_t p, int32_t tot, int32_t n_ahead, const double* pred,
                       const double* dev_eps, uint64_t seed, double* dev_out, void* cuda_stream);
/* the same for structural models: y_{T+j+1} = A0^-1 (A pred_j + Sigma^{1/2} eps_j) (R/predict.bvar.R:200-209,
 * draw_forecast.cpp:33-41); dev_a0 as in bvg_irf_draws_a0 (NULL = identity) */
int bvg_forecast_draws_a0(const double* dev_coef, int64_t coef_stride, const double* dev_sigma, int64_t sigma_stride,
                          const double* dev_a0, int64_t a0_stride, int32_t a0_compact, int64_t n_draws, int32_t k, int32_t p,
                          int32_t tot, int32_t n_ahead, const double* pred, const double* dev_eps, uint64_t seed,
                          double* dev_out, void* cuda_stream);
/* on the retained draws of a sampler (TVP / SV: the last period) + the bands of predict.bvar (:219-221): quantiles
 * [n_probs][n_ahead * k] (entry j * k + i = horizon j + 1, variable i), mean, host_draws [draws][n_ahead][k] (may be NULL),
 * host_eps injected normals on the HOST (may be NULL) */
int bvg_sampler_forecast(bvg_sampler* s, int32_t p, int32_t tot, int32_t n_ahead, const double* pred, const double* host_eps,
                         uint64_t seed, const double* probs, int32_t n_probs, double* quantiles, double* mean,
                         double* host_draws, void* cuda_stream);

/* measured FP64 FMA throughput of the device (TFLOP/s) -- roofline denominator for the sweep kernels */
int bvg_fp64_peak(double* tflops_fma, void* cuda_stream);
/* measured FP64 tensor-core (DMMA m8n8k4) throughput (TFLOP/s) -- denominator for the large-VAR Cholesky */
int bvg_dmma_peak(double* tflops, void* cuda_stream);

/* ---- batched stand-alone building blocks (B independent problems per call; host pointers) ---------------- */
/* post_normal (src/post_normal.cpp:61-93): out[b] = a_post + sqrtm(V_post) eps.  method 0 = symmetric
 * eigen square root (reference mapping), 1 = Cholesky L^-T (bvaralg mapping). eps NULL = Philox(seed). */
int bvg_post_normal(int64_t B, int k, int T, int M, const double* y, const double* x, const double* sigma_i,
                    const double* a_prior, const double* v_i_prior, const double* eps, uint64_t seed,
                    int method, double* out);
/* post_normal_sur (src/post_normal_sur.cpp:60-113): z kT x n shared, sigma_i [B][k x k] or [B][kT x k] */
int bvg_post_normal_sur(int64_t B, int k, int T, int n, const double* y, const double* z, const double* sigma_i,
                        int sigma_tv, const double* a_prior, const double* v_i_prior, const double* eps,
                        uint64_t seed, int method, double* out);
/* ssvs (src/svss.cpp:76-111): a [B][n]; lambda_out [B][n], vdiag_out [B][n] (diagonal of v_i) */
int bvg_ssvs(int64_t B, int n, const double* a, const double* tau0, const double* tau1, const double* prob_prior,
             const int32_t* include, int n_include, const double* u, uint64_t seed, double* lambda_out,
             double* vdiag_out);
/* bvs (src/bvs.cpp:73-159): y k x T and z kT x n shared; per problem a [B][n] (a_tv = 0) or [B][n x T]
 * (a_tv = 1, time-varying coefficients, bvs.cpp:83-86,141-146), lambda [B][n] (the diagonal of the reference's
 * n x n indicator matrix), sigma_i [B][k x k] (sigma_tv = 0) or [B][kT x k] (sigma_tv = 1, bvs.cpp:88-100) */
int bvg_bvs(int64_t B, int k, int T, int n, const double* y, const double* z, const double* a, int a_tv,
            const double* lambda, const double* sigma_i, int sigma_tv, const double* prob_prior,
            const int32_t* include, int n_include, const double* u1, const double* u2, uint64_t seed,
            double* lambda_out);
/* stochvol_ksc1998 / stochvol_ocsn2007 / stoch_vol (src/stochvol_*.cpp:60-114): y,h [B][T x k] */
int bvg_stochvol(int64_t B, int T, int k, const double* y, const double* h, const double* sigma,
                 const double* h_init, const double* constant, int mixture, const double* u, const double* eps,
                 uint64_t seed, double* h_out, int32_t* s_out);
/* kalman_dk (src/kalman_dk.cpp:91-184): y k x T and z kT x n shared; per problem sigma_u (k x k, or kT x k
 * when tv_mask & 1), sigma_v (n x n, or nT x n when tv_mask & 2), B (n x n, or nT x n when tv_mask & 4) -- the
 * stacked forms of kalman_dk.cpp:99-121 --, a_init (n), P_init (n x n); out [B][n x (T+1)].
 * eps: [B][n + T (k + n)] N(0,1) in the reference's consumption order, or NULL = Philox(seed) */
int bvg_kalman_dk(int64_t B, int k, int T, int n, const double* y, const double* z, const double* sigma_u,
                  const double* sigma_v, const double* Bmat, int tv_mask, const double* a_init,
                  const double* P_init, const double* eps, uint64_t seed, double* out);

#ifdef __cplusplus
}
#endif
#endif /* BVARGPU_H */
