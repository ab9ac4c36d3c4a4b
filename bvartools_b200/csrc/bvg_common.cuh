// bvg_common.cuh -- shared definitions: host/device macros, Philox4x32-10, variate transforms, site addressing.
//
// All kernel logic in this directory is written as group-generic __host__ __device__ templates (see
// bvg_group.cuh).  The product path instantiates them inside CUDA kernels for sm_100a; tests/emu compiles the
// very same headers with g++ (single-thread group) purely to check the arithmetic against the oracle on
// machines without a GPU.  There is no CPU execution path in the product library.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define BVG_HD __host__ __device__ __forceinline__
#define BVG_D __device__ __forceinline__
// out-of-line on the device: the variate generators are called from many sites and would otherwise be inlined
// (Philox unrolled + log/cospi expansions) at each of them, blowing the kernel past the instruction cache
#define BVG_HD_NOINLINE static __host__ __device__ __noinline__
#define BVG_UNROLL _Pragma("unroll")
#else
#define BVG_UNROLL
#define BVG_HD inline
#define BVG_D inline
#define BVG_HD_NOINLINE inline
#endif

namespace bvg {

// ---- variate blocks (site addressing, SURVEY.md Appendix B / oracle/variates.py) --------------------------
enum VarBlock : int {
  VB_COEF_N = 0,   // N(0,1) x n           coefficient draw                 bvaralg.cpp:307
  VB_SSVS_U = 1,   // U(0,1) x n           SSVS Bernoulli                   bvaralg.cpp:322
  VB_BVS_U1 = 2,   // U(0,1) x n           BVS prior gate                   bvaralg.cpp:338
  VB_BVS_U2 = 3,   // U(0,1) x n           BVS acceptance                   bvaralg.cpp:351
  VB_SV_U = 4,     // U(0,1) x k*T         mixture indicators               stochvol_ksc1998.cpp:96
  VB_SV_N = 5,     // N(0,1) x k*T         log-volatility path              stochvol_ksc1998.cpp:104
  VB_SIGH_G = 6,   // Gamma(shape,1) x k   sigma_h                          bvaralg.cpp:470
  VB_HINIT_N = 7,  // N(0,1) x k           h_init                           bvaralg.cpp:477
  VB_OMEGA_G = 8,  // Gamma(shape,1) x k   inverse-gamma error variances    bvaralg.cpp:484
  VB_WISH_C = 9,   // chi2(df-i) x k       Bartlett diagonal                bvaralg.cpp:511
  VB_WISH_N = 10,  // N(0,1) x k(k-1)/2    Bartlett strict upper triangle
  VB_STATE_N = 11, // N(0,1) x n*T         TVP state path                   bvartvpalg.cpp:297
  VB_SIGV_G = 12,  // Gamma(shape,1) x n   TVP state precisions             bvartvpalg.cpp:476
  VB_AINIT_N = 13, // N(0,1) x n           TVP initial state                bvartvpalg.cpp:482
  VB_PSI_N = 14,   // N(0,1) x k(k-1)/2 (TVP: x T)  covariance (psi) block   bvaralg.cpp:394, bvartvpalg.cpp:363
  VB_PSI_U1 = 15,  // U(0,1) x k(k-1)/2    SSVS Bernoulli / BVS gate on psi bvaralg.cpp:409,424
  VB_PSI_U2 = 16,  // U(0,1) x k(k-1)/2    BVS acceptance on psi            bvaralg.cpp:437
  VB_PSI_G = 17,   // Gamma(shape,1) x k(k-1)/2  TVP psi state precisions     bvartvpalg.cpp:492
  VB_PSI_IN = 18,  // N(0,1) x k(k-1)/2    TVP psi initial state             bvartvpalg.cpp:498
  VB_COUNT = 19
};

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: no state, any site can be drawn by any thread ------
struct Philox4 { uint32_t x, y, z, w; };
struct Normal2 { double a, b; };

BVG_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

BVG_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  BVG_UNROLL
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

// Uniform strictly inside (0, 1) from 52 random bits: (v + 1/2) 2^-52, v = the 52 bits.  Built without an integer ->
// double conversion: the bits become the mantissa of a double in [1, 2), and one exact subtraction of 1 - 2^-53 gives
// (v + 1/2) 2^-52 (the result has at most 53 significant bits).  Three instructions instead of seven (shift, shift-merge,
// I2F.F64.U64, DADD, DMUL) -- the generators are issue bound.
BVG_HD double u53(uint32_t hi, uint32_t lo) {
  const uint32_t h = (hi >> 12) | 0x3ff00000u;
#if defined(__CUDA_ARCH__)
  const double d = __hiloint2double((int)h, (int)lo);
#else
  const uint64_t bits = ((uint64_t)h << 32) | (uint64_t)lo;
  double d;
  memcpy(&d, &bits, 8);
#endif
  return d - 0.99999999999999988897769753748434595763683319091796875;
}

BVG_HD double bvg_cospi(double x) {
#if defined(__CUDA_ARCH__)
  return cospi(x);
#else
  return cos(3.14159265358979323846 * x);
#endif
}
BVG_HD double bvg_sinpi(double x) {
#if defined(__CUDA_ARCH__)
  return sinpi(x);
#else
  return sin(3.14159265358979323846 * x);
#endif
}
BVG_HD void bvg_sincospi(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincospi(x, s, c);
#else
  *s = sin(3.14159265358979323846 * x);
  *c = cos(3.14159265358979323846 * x);
#endif
}

// ---- FP64 transforms of the variate generators --------------------------------------------------------------------
// The generators are instruction-issue bound (profiles/r02_c2a_*): what counts is the number of issued instructions per
// variate, not FP64 flops.  Three things are done about that here:
//  * polynomial coefficients live in __constant__ tables: sm_100a DFMA takes no immediate / constant-bank operand, a
//    literal double costs two UMOV (or IMAD.MOV) per use (17 % of the executed instructions of the round-1 kernel); from
//    a table one LDCU.128 brings two coefficients into uniform registers;
//  * log / sqrt / sincos are written for the ranges the generators produce (positive normal arguments, the angle in
//    (0, 1) turns): no special-case handling, no slow-path calls, MUFU seed + Newton steps spelled out;
//  * the same source is the host build of tests/emu (plain division / sqrt there: results agree to ~1e-16 relative).
#if defined(__CUDACC__)
#define BVG_TABLE(name, n, ...)                                             \
  static __constant__ double name##_d[n] __attribute__((unused)) = {__VA_ARGS__}; \
  static const double name##_h[n] __attribute__((unused)) = {__VA_ARGS__};
#else
#define BVG_TABLE(name, n, ...) static const double name##_h[n] __attribute__((unused)) = {__VA_ARGS__};
#endif
#if defined(__CUDA_ARCH__)
#define BVG_TAB(name) name##_d
#else
#define BVG_TAB(name) name##_h
#endif

// sin(x) / x and cos(x) in z = x^2 on |x| <= pi/4, and P(z) of 2 atanh(f) = 2 f + f z P(z), z = f^2 <= 0.02944:
// interpolation at the Chebyshev nodes of the interval (near-minimax; 50-digit fit, tools/fit_variate_polys.py):
// max errors 3.2e-18 (sin), 3.0e-20 (cos), 4.6e-18 relative to the logarithm.  Two coefficients shorter each than the
// Taylor polynomials of the same accuracy.
BVG_TABLE(kSinTab, 7, 1.0, -0.16666666666666616, 0.008333333333320363, -0.000198412698286503, 2.755731337640013e-06,
          -2.5050716974102745e-08, 1.5894736651849094e-10)
BVG_TABLE(kCosTab, 8, 1.0, -0.5, 0.04166666666666645, -0.0013888888888861095, 2.4801587283881152e-05,
          -2.7557313097790086e-07, 2.0875582380663952e-09, -1.1353379638297575e-11)
BVG_TABLE(kLogTab, 8, 0.666666666666667, 0.39999999999899505, 0.28571428625975487, 0.2222221113479508,
          0.18182889125261723, 0.15331721600556042, 0.14616449685043406, 0.6931471805599453)

// log(x) for a positive NORMAL double (no zero / denormal / inf / nan handling): x = m 2^e with m in [2^-1/2, 2^1/2),
// log m = 2 atanh((m - 1) / (m + 1)).  ~2 ulp.
BVG_HD double bvg_logpos(double x) {
#if defined(__CUDA_ARCH__)
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
#else
  uint64_t bits;
  memcpy(&bits, &x, 8);
  int hi = (int)(bits >> 32);
  const int lo = (int)(uint32_t)bits;
#endif
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  if (hi >= 0x3ff6a09f) { hi -= 0x00100000; e += 1; }
#if defined(__CUDA_ARCH__)
  const double m = __hiloint2double(hi, lo);
#else
  double m;
  bits = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  memcpy(&m, &bits, 8);
#endif
  const double num = m - 1.0, den = m + 1.0;
  double f;
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));      // MUFU.RCP64H, then two Newton steps and a residual fix
  double t = fma(-den, r, 1.0);
  r = fma(r, t, r);
  t = fma(-den, r, 1.0);
  r = fma(r, t, r);
  f = num * r;
  f = fma(fma(-den, f, num), r, f);
#else
  f = num / den;
#endif
  const double z = f * f;
  const double* c = BVG_TAB(kLogTab);
  double p = c[6];
  p = fma(p, z, c[5]);
  p = fma(p, z, c[4]);
  p = fma(p, z, c[3]);
  p = fma(p, z, c[2]);
  p = fma(p, z, c[1]);
  p = fma(p, z, c[0]);
  const double lm = fma(f * z, p, f + f);
  return fma((double)e, c[7], lm);
}

// sqrt(x) for a positive NORMAL double: MUFU.RSQ64H seed, two coupled (Goldschmidt) steps, residual fix.  ~1 ulp.
BVG_HD double bvg_sqrtpos(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double g = x * r, h = 0.5 * r;
  double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  return fma(fma(-g, g, x), h, g);
#else
  return sqrt(x);
#endif
}

// cos(2 pi u), sin(2 pi u) for u in (0, 1): the Box-Muller angle.  t = 4 u, q = rint(t) in {0..4}, phi = (pi / 2)(t - q)
// in [-pi/4, pi/4]; sin(phi) and cos(phi) by polynomials in phi^2, then a rotation by q quarter turns.
BVG_HD void bvg_sincos2pi(double u, double* sn, double* cs) {
  const double t = 4.0 * u;
  const double qd = rint(t);
  const int q = (int)qd;
  const double x = (t - qd) * 1.5707963267948966;
  const double z = x * x;
  const double* ts = BVG_TAB(kSinTab);
  const double* tc = BVG_TAB(kCosTab);
  double ps = ts[6];
  double pc = tc[7];
  pc = fma(pc, z, tc[6]);
  ps = fma(ps, z, ts[5]);
  pc = fma(pc, z, tc[5]);
  ps = fma(ps, z, ts[4]);
  pc = fma(pc, z, tc[4]);
  ps = fma(ps, z, ts[3]);
  pc = fma(pc, z, tc[3]);
  ps = fma(ps, z, ts[2]);
  pc = fma(pc, z, tc[2]);
  ps = fma(ps, z, ts[1]);
  pc = fma(pc, z, tc[1]);
  ps = fma(ps, z, ts[0]);
  pc = fma(pc, z, tc[0]);
  const double s = x * ps, c = pc;
  const double a = (q & 1) ? s : c, b = (q & 1) ? c : s;           // |cos|, |sin| before the quadrant signs
  *cs = (((q + 1) >> 1) & 1) ? -a : a;                              // q = 0: c  1: -s  2: -c  3: s  4: c
  *sn = ((q >> 1) & 1) ? -b : b;                                    // q = 0: s  1: c   2: -s  3: -c 4: s
}

// Box-Muller: (u1, u2) in (0, 1)^2 -> two independent N(0, 1)
BVG_HD void bvg_box_muller(double u1, double u2, double* a, double* b) {
  const double r = bvg_sqrtpos(-2.0 * bvg_logpos(u1));
  double sn, cs;
  bvg_sincos2pi(u2, &sn, &cs);
  *a = r * cs;
  *b = r * sn;
}

// ---- random source of one chain ----------------------------------------------------------------------------
// Injected variates (tier-1 parity tests): device arrays [chain][sweep][count] or nullptr per block.  Lives in
// kernel parameter (constant) space; the chain only keeps a pointer to it so that the Philox path carries no
// per-block pointers in registers.
struct RngParams {
  uint64_t seed;
  uint64_t chain_offset;
  const double* inj[VB_COUNT];
  int inj_stride[VB_COUNT];               // count per sweep
  long long inj_chain_stride[VB_COUNT];   // sweeps * count
};

// Site = (block, sweep, index[, attempt]).  Key = seed, counter = (index | attempt<<24, sweep, chain_lo,
// chain_hi ^ block<<24): results are independent of thread mapping, launch chunking and GPU count.
//
// The generators proper are out-of-line FREE functions over by-value scalars (one copy of the Philox / log / sincospi
// code per kernel, and no `this` pointer that would force the caller's ChainRng onto the stack); the ChainRng members
// are inlined wrappers that test for injected variates first.
BVG_HD Philox4 philox_site(uint64_t seed, uint64_t chain, int block, int sweep, int idx, int attempt) {
  return philox4x32_10((uint32_t)idx | ((uint32_t)attempt << 24), (uint32_t)sweep, (uint32_t)chain,
                       (uint32_t)(chain >> 32) ^ ((uint32_t)block << 24), (uint32_t)seed, (uint32_t)(seed >> 32));
}
BVG_HD_NOINLINE double philox_uniform(uint64_t seed, uint64_t chain, int block, int sweep, int idx, int attempt) {
  const Philox4 p = philox_site(seed, chain, block, sweep, idx, attempt);
  return u53(p.x, p.y);
}
// Normals come in Box-Muller pairs: sites 2p and 2p+1 of a block are the cosine and sine branch of ONE Philox call
// (counter index p).
BVG_HD_NOINLINE Normal2 philox_normal_pair(uint64_t seed, uint64_t chain, int block, int sweep, int pair, int attempt) {
  const Philox4 p = philox_site(seed, chain, block, sweep, pair, attempt);
  Normal2 o;
  bvg_box_muller(u53(p.x, p.y), u53(p.z, p.w), &o.a, &o.b);
  return o;
}
// Two independent pairs at once (two chains / two sites): the two Philox + log + sincospi dependency chains are
// interleaved by the instruction scheduler, which roughly doubles the issue rate of a latency-bound warp.
struct Normal4 { Normal2 p, q; };
BVG_HD_NOINLINE Normal4 philox_normal_pair_x2(uint64_t seed, uint64_t chain_a, int block_a, int sweep_a, int pair_a,
                                              uint64_t chain_b, int block_b, int sweep_b, int pair_b) {
  const Philox4 pa = philox_site(seed, chain_a, block_a, sweep_a, pair_a, 0);
  const Philox4 pb = philox_site(seed, chain_b, block_b, sweep_b, pair_b, 0);
  Normal4 o;
  bvg_box_muller(u53(pa.x, pa.y), u53(pa.z, pa.w), &o.p.a, &o.p.b);
  bvg_box_muller(u53(pb.x, pb.y), u53(pb.z, pb.w), &o.q.a, &o.q.b);
  return o;
}
BVG_HD_NOINLINE double philox_normal(uint64_t seed, uint64_t chain, int block, int sweep, int idx, int attempt) {
  const Philox4 p = philox_site(seed, chain, block, sweep, idx >> 1, attempt);
  double a, b;
  bvg_box_muller(u53(p.x, p.y), u53(p.z, p.w), &a, &b);
  return (idx & 1) ? b : a;
}
// Gamma(shape, 1), Marsaglia-Tsang 2000 (shape < 1 boosted); attempts are separate sites.  `x0` is the N(0,1) of
// attempt 0, i.e. philox_normal(.., idx, 0): callers may draw it earlier together with other normals so that all
// lanes of a tile run one generator call site instead of diverging.
BVG_HD_NOINLINE double philox_gamma_from_x0(uint64_t seed, uint64_t chain, int block, int sweep, int idx, double shape,
                                            double x0) {
  double boost = 1.0, a = shape;
  if (a < 1.0) {
    const Philox4 p = philox_site(seed, chain, block, sweep, idx, 255);
    boost = pow(u53(p.x, p.y), 1.0 / a);
    a += 1.0;
  }
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  double x = x0;
  for (int attempt = 0; attempt < 100; ++attempt) {
    if (attempt > 0) {
      const Philox4 p = philox_site(seed, chain, block, sweep, idx, attempt);
      double xb;
      bvg_box_muller(u53(p.x, p.y), u53(p.z, p.w), &x, &xb);
    }
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    const Philox4 q = philox_site(seed, chain, block, sweep, idx, attempt + 100);
    const double u = u53(q.x, q.y);
    const double x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;       // squeeze: sufficient for the test below, no logs
    if (bvg_logpos(u) < 0.5 * x2 + d - d * v + d * bvg_logpos(v)) return boost * d * v;
  }
  return boost * d;   // unreachable in practice (acceptance > 95 %)
}

struct ChainRng {
  uint64_t seed;
  uint64_t chain;                 // GLOBAL chain id
  const RngParams* rp;            // injected variates (or all-null)
  long long chain_local;

  BVG_HD ChainRng(const RngParams* p, long long local)
      : seed(p->seed), chain(p->chain_offset + (uint64_t)local), rp(p), chain_local(local) {}

  BVG_HD Philox4 raw(int block, int sweep, int idx, int attempt) const {
    return philox_site(seed, chain, block, sweep, idx, attempt);
  }
  BVG_HD bool injected(int block) const { return rp->inj[block] != nullptr; }
  BVG_HD double inj_at(int block, int sweep, int idx) const {
    return rp->inj[block][(size_t)chain_local * (size_t)rp->inj_chain_stride[block] +
                          (size_t)sweep * (size_t)rp->inj_stride[block] + (size_t)idx];
  }
  BVG_HD double uniform(int block, int sweep, int idx, int attempt = 0) const {
    if (rp->inj[block]) return inj_at(block, sweep, idx);
    return philox_uniform(seed, chain, block, sweep, idx, attempt);
  }
  // one member of a Box-Muller pair
  BVG_HD double normal(int block, int sweep, int idx, int attempt = 0) const {
    if (rp->inj[block]) return inj_at(block, sweep, idx);
    return philox_normal(seed, chain, block, sweep, idx, attempt);
  }
  // normals of sites 2*pair and 2*pair+1 (`count` = sites in the block: an odd block's last pair has one member)
  BVG_HD Normal2 normal_pair(int block, int sweep, int pair, int count) const {
    if (rp->inj[block]) {
      Normal2 o;
      o.a = inj_at(block, sweep, 2 * pair);
      o.b = (2 * pair + 1 < count) ? inj_at(block, sweep, 2 * pair + 1) : 0.0;
      return o;
    }
    return philox_normal_pair(seed, chain, block, sweep, pair, 0);
  }
  BVG_HD double gamma_from_x0(int block, int sweep, int idx, double shape, double x0) const {
    if (rp->inj[block]) return inj_at(block, sweep, idx);
    return philox_gamma_from_x0(seed, chain, block, sweep, idx, shape, x0);
  }
  BVG_HD double gamma(int block, int sweep, int idx, double shape) const {
    if (rp->inj[block]) return inj_at(block, sweep, idx);
    return philox_gamma_from_x0(seed, chain, block, sweep, idx, shape, philox_normal(seed, chain, block, sweep, idx, 0));
  }
  BVG_HD double chi2_from_x0(int block, int sweep, int idx, double df, double x0) const {
    if (rp->inj[block]) return inj_at(block, sweep, idx);
    return 2.0 * philox_gamma_from_x0(seed, chain, block, sweep, idx, 0.5 * df, x0);
  }
  BVG_HD double chi2(int block, int sweep, int idx, double df) const {
    if (rp->inj[block]) return inj_at(block, sweep, idx);
    return 2.0 * gamma(block, sweep, idx, 0.5 * df);
  }
};

BVG_HD int tri(int i) { return (i * (i + 1)) >> 1; }   // offset of row i in packed lower row-major storage

// R's rbinom(1, 1, p) driven by one uniform (convention: oracle/variates.py)
BVG_HD double bernoulli_rbinom(double p, double u) { return (p > 0.5) ? ((u < p) ? 1.0 : 0.0) : ((u >= 1.0 - p) ? 1.0 : 0.0); }

}  // namespace bvg
