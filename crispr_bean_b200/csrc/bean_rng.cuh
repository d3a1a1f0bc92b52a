// bean_rng.cuh -- counter-based noise for the guide program and the Dirichlet pathwise derivative.
//
// Replaces the reparameterised draws of the pyro guides (bean/model/model.py:810-811 Normal/LogNormal,
// :839-847 Dirichlet -> torch._sample_dirichlet) and torch._dirichlet_grad, the backward of
// Dirichlet.rsample (44 % of the reference's CPU step time at scale).
//   * Philox4x32-10 (Salmon et al., SC'11) keyed by the run seed and indexed by (entity, replicate,
//     step, stream): reproducible and independent of launch geometry.
//   * Gamma(alpha) by Marsaglia & Tsang (2000) with the alpha < 1 boost -- the algorithm behind
//     torch's sample_gamma (ATen/native/Distributions.h); pi = g / sum g, clamped to
//     [tiny, 1 - eps/2] exactly like _sample_dirichlet, so log(pi) stays finite.
//   * dirichlet_grad_one: the four-regime approximation of -(dCDF/dalpha)/pdf of torch
//     (ATen/native/Distributions.h: _beta_grad_alpha_small / _beta_grad_beta_small /
//     _beta_grad_alpha_mid / rational correction).  It is restated here, coefficient for coefficient,
//     because parity means reproducing the reference's gradient, approximation included.
#pragma once
#include "bean_math.cuh"

namespace bean {

static __device__ __noinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// The same ten rounds with the round keys k + i (0x9E3779B9, 0xBB67AE85) handed in precomputed (by the host, in the kernel's
// parameter block: each becomes a constant-bank operand of the round's LOP3), inlined: 4 instructions per round instead of 6
// plus the call -- for the one call site that runs once per (guide, replicate).
struct PhiloxKeys { uint32_t k[20]; };
__device__ __forceinline__ uint4 philox4x32_10_keys(uint4 c, const PhiloxKeys& rk) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ rk.k[2 * i], lo1, hi0 ^ c.w ^ rk.k[2 * i + 1], lo0);
  }
  return c;
}
static inline void philox_round_keys(uint64_t seed, PhiloxKeys& rk) {
  for (int i = 0; i < 10; ++i) {
    rk.k[2 * i] = (uint32_t)seed + (uint32_t)i * 0x9E3779B9u;
    rk.k[2 * i + 1] = (uint32_t)(seed >> 32) + (uint32_t)i * 0xBB67AE85u;
  }
}

__device__ __forceinline__ uint2 seed_key(uint64_t seed) { return make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)); }

// uniform in (0, 1), 24 bits
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * 5.9604644775390625e-8f; }

// two standard normals from two 32-bit words (Box-Muller)
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float r = sqrtf(-2.0f * log_ftz(u01(a)));
  float s, c;
  __sincosf(6.283185307179586f * (u01(b) - 0.5f), &s, &c);  // angle in (-pi, pi): MUFU.SIN/COS range
  n0 = r * c;
  n1 = r * s;
}

enum : uint32_t { STREAM_VARIANT = 0, STREAM_NOISE = 4, STREAM_BOOST = 8, STREAM_PI = 16 };

// (eps_mu, eps_sd) of variant v at `step`
__device__ __forceinline__ void variant_noise(uint64_t seed, uint32_t v, uint32_t step, float& e0, float& e1) {
  const uint4 w = philox4x32_10(make_uint4(v, 0u, step, STREAM_VARIANT), seed_key(seed));
  box_muller(w.x, w.y, e0, e1);
}

// standard normal behind logit_pi_noise of guide g at `step`
__device__ __forceinline__ float guide_noise(uint64_t seed, uint32_t g, uint32_t step) {
  const uint4 w = philox4x32_10(make_uint4(g, 0u, step, STREAM_NOISE), seed_key(seed));
  float e0, e1;
  box_muller(w.x, w.y, e0, e1);
  return e0;
}

// Marsaglia-Tsang constants of one shape parameter (boosted to >= 1); they depend on the guide only, so
// they are computed once per guide, not once per replicate.
template <typename real>
struct GammaMT {
  real d, c, inv_alpha;  // d = a - 1/3, c = 1/sqrt(9 d); inv_alpha = 1/alpha if alpha < 1 (boost) else 0
  __device__ __forceinline__ void init(real alpha) {
    inv_alpha = real(0);
    if (alpha < real(1)) {
      inv_alpha = real(1) / alpha;
      alpha += real(1);
    }
    d = alpha - real(1.0 / 3.0);
    c = real(1) / Num<real>::sqrt(real(9) * d);
  }
  // one proposal; returns true and the draw when accepted
  __device__ __forceinline__ bool attempt(float nrm, float uni, real& out) const {
    const real x = real(nrm);
    const real y = real(1) + c * x;
    if (y <= real(0)) return false;
    const real v = y * y * y;
    const real xx = x * x;
    if (real(uni) < real(1) - real(0.0331) * xx * xx ||
        Num<real>::flog(real(uni)) < real(0.5) * xx + d * (real(1) - v + Num<real>::flog(v))) {
      out = d * v;
      return true;
    }
    return false;
  }
};

template <typename real> struct Lim;
template <> struct Lim<float> {
  static __device__ __forceinline__ float tiny() { return 1.17549435e-38f; }
  static __device__ __forceinline__ float one_minus() { return 0.99999994f; }  // nexttoward(1, 0)
  static __device__ __forceinline__ float eps() { return 1.1920929e-7f; }
};
template <> struct Lim<double> {
  static __device__ __forceinline__ double tiny() { return 2.2250738585072014e-308; }
  static __device__ __forceinline__ double one_minus() { return 0.99999999999999989; }
  static __device__ __forceinline__ double eps() { return 2.220446049250313e-16; }
};

// pi ~ Dirichlet(c0, c1) for guide g, replicate r at `step` (a Beta draw as two gammas).
template <typename real>
__device__ __forceinline__ void sample_pi2(uint64_t seed, uint32_t g, uint32_t r, uint32_t step, const GammaMT<real>& m0,
                                           const GammaMT<real>& m1, real& pi0, real& pi1, const PhiloxKeys* rk = nullptr) {
  const uint2 key = seed_key(seed);
  real s0 = real(1), s1 = real(1);
  if (m0.inv_alpha != real(0) || m1.inv_alpha != real(0)) {  // boost: Gamma(a) = Gamma(a + 1) * U^(1/a)
    const uint4 w = philox4x32_10(make_uint4(g, r, step, STREAM_BOOST), key);
    if (m0.inv_alpha != real(0)) s0 = Num<real>::pow(real(1) - real(u01(w.x)), m0.inv_alpha);
    if (m1.inv_alpha != real(0)) s1 = Num<real>::pow(real(1) - real(u01(w.y)), m1.inv_alpha);
  }
  real g0 = real(0), g1 = real(0);
  bool ok0 = false, ok1 = false;
  for (uint32_t k = 0; k < 32u && !(ok0 && ok1); ++k) {
    const uint4 w = (rk && k == 0u) ? philox4x32_10_keys(make_uint4(g, r, step, STREAM_PI), *rk)
                                    : philox4x32_10(make_uint4(g, r, step, STREAM_PI + k), key);  // retries (rare): out of line
    float n0, n1;
    box_muller(w.x, w.y, n0, n1);
    if (!ok0) ok0 = m0.attempt(n0, 1.0f - u01(w.z), g0);
    if (!ok1) ok1 = m1.attempt(n1, 1.0f - u01(w.w), g1);
  }
  g0 = Num<real>::fmax(g0 * s0, Lim<real>::tiny());
  g1 = Num<real>::fmax(g1 * s1, Lim<real>::tiny());
  const real inv = real(1) / (g0 + g1);
  pi0 = Num<real>::fmin(Num<real>::fmax(g0 * inv, Lim<real>::tiny()), Lim<real>::one_minus());
  pi1 = Num<real>::fmin(Num<real>::fmax(g1 * inv, Lim<real>::tiny()), Lim<real>::one_minus());
}

// ---- reparameterised gradient of a Dirichlet/Beta draw ---------------------------------------------
template <typename real>
__device__ __forceinline__ real digamma_full(real z) {
  real lg, dg;
  lgamma_digamma(z, lg, dg);
  return dg;
}
template <>
__device__ __forceinline__ double digamma_full<double>(double z) { return digamma_f64(z); }  // no lgamma needed

// x near 0: Taylor series in x (torch: _beta_grad_alpha_small).  torch's `factor + 1/alpha` with
// factor = psi(alpha) - psi(alpha + beta) - ln x cancels from O(1/alpha) to O(1) for small alpha (4 digits lost in
// float at alpha = 3e-4); psi(alpha) + 1/alpha = psi(alpha + 1) removes the cancellation, the series is unchanged:
//   factor + 1/(alpha + i) = f1 - i / (alpha (alpha + i)),   f1 = psi(alpha + 1) - psi(alpha + beta) - ln x.
template <typename real>
__device__ __forceinline__ real beta_grad_alpha_small(real x, real alpha, real beta) {
  const real f1 = digamma_full(alpha + real(1)) - digamma_full(alpha + beta) - Num<real>::log(x);
  const real ialpha = real(1) / alpha;
  real numer = real(1);
  real series = numer * ialpha * f1;
#pragma unroll 1
  for (int i = 1; i <= 10; ++i) {
    const real ci = real(i);
    numer *= (ci - beta) * x / ci;
    const real idenom = real(1) / (alpha + ci);
    series += numer * idenom * (f1 - ci * ialpha * idenom);
  }
  const real result = x * Num<real>::pow(real(1) - x, -beta) * series;
  return isnan(result) ? real(0) : result;
}

// double: torch's expression AS WRITTEN (factor + 1 / (alpha + i) with factor = psi(alpha) - psi(alpha + beta) - ln x).  For
// tiny concentrations (non-existent tiling alleles: ~1e-7) its cancellation costs the reference ~1e-9 of relative accuracy,
// and fp64 parity means reproducing that number, not the better one.
// `psi_diff` = psi(alpha) - psi(alpha + beta): a caller that evaluates many draws of one concentration vector passes it in
// (NaN = compute it here); it is O(1) accurate either way -- the cancellation is between psi(alpha) and 1 / alpha below.
__device__ __forceinline__ double beta_grad_alpha_small_f64(double x, double alpha, double beta, double psi_diff) {
  if (isnan(psi_diff)) psi_diff = digamma_f64(alpha) - digamma_f64(alpha + beta);
  const double factor = psi_diff - ::log(x);
  double numer = 1.0;
  double series = numer / alpha * (factor + 1.0 / alpha);
#pragma unroll 1
  for (int i = 1; i <= 10; ++i) {
    const double ci = (double)i;
    numer *= (ci - beta) * x / ci;
    // draws of absent alleles sit at ~1e-300: numer underflows, and every division by a denormal takes the FP64 divide's slow
    // path (30 % of the fused tiling kernel's instructions before this line).  Terms below 1e-280 cannot change `series`,
    // whose first term is >= (-ln x) / alpha ~ 600 / alpha there: the result is bit-identical.
    if (::fabs(numer) < 1e-280) break;
    const double denom = alpha + ci;
    series += numer / denom * (factor + 1.0 / denom);
  }
  const double result = x * ::pow(1.0 - x, -beta) * series;
  return isnan(result) ? 0.0 : result;
}
template <>
__device__ __forceinline__ double beta_grad_alpha_small<double>(double x, double alpha, double beta) {
  return beta_grad_alpha_small_f64(x, alpha, beta, nan(""));
}

// float: the same series with every division a MUFU reciprocal (or a constant), ln x and (1 - x)^-beta through MUFU lg2 / ex2
// -- in this regime beta |ln(1 - x)| <~ 5, so the exponent's rounding costs < 1e-6 -- and the ten terms unrolled: ~110
// instructions instead of ~450.  These regimes run with a handful of active lanes (the draws the per-warp queue collects), and
// in the survival step, whose concentrations stay below 6, they are ALL of the alpha kernel's work.
__device__ __forceinline__ float exp_mufu(float t) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t * 1.4426950408889634f));
  return r;
}
template <>
__device__ __forceinline__ float beta_grad_alpha_small<float>(float x, float alpha, float beta) {
  const float f1 = digamma_full(alpha + 1.0f) - digamma_full(alpha + beta) - log_ftz(x);
  const float ialpha = rcp_ftz(alpha);
  float numer = 1.0f;
  float series = ialpha * f1;
#pragma unroll
  for (int i = 1; i <= 10; ++i) {
    const float ci = (float)i;
    numer *= (ci - beta) * x * (1.0f / ci);
    const float idenom = rcp_ftz(alpha + ci);
    series = fmaf(numer * idenom, fmaf(-ci * ialpha, idenom, f1), series);
  }
  const float result = x * exp_mufu(-beta * log1p_ratio_series(-x)) * series;  // x <= 0.5: inside the series' range
  return isnan(result) ? 0.0f : result;
}

// x near 0, derivative w.r.t. beta (torch: _beta_grad_beta_small)
template <typename real>
__device__ __forceinline__ real beta_grad_beta_small(real x, real alpha, real beta) {
  const real factor = digamma_full(alpha + beta) - digamma_full(beta);
  real numer = real(1), betas = real(1), dbetas = real(0), series = factor / alpha;
#pragma unroll 1
  for (int i = 1; i <= 8; ++i) {
    const real ci = real(i);
    numer *= -x / ci;
    dbetas = dbetas * (beta - ci) + betas;
    betas = betas * (beta - ci);
    series += numer / (alpha + ci) * (dbetas + factor * betas);
  }
  const real result = -Num<real>::pow(real(1) - x, real(1) - beta) * series;
  return isnan(result) ? real(0) : result;
}
template <>
__device__ __forceinline__ float beta_grad_beta_small<float>(float x, float alpha, float beta) {
  const float factor = digamma_full(alpha + beta) - digamma_full(beta);
  float numer = 1.0f, betas = 1.0f, dbetas = 0.0f, series = factor * rcp_ftz(alpha);
#pragma unroll
  for (int i = 1; i <= 8; ++i) {
    const float ci = (float)i;
    numer *= -x * (1.0f / ci);
    dbetas = fmaf(dbetas, beta - ci, betas);
    betas = betas * (beta - ci);
    series = fmaf(numer * rcp_ftz(alpha + ci), fmaf(factor, betas, dbetas), series);
  }
  const float result = -exp_mufu((1.0f - beta) * log1p_ratio_series(-x)) * series;
  return isnan(result) ? 0.0f : result;
}

// |x - mean| <= 0.1 std inside the saddle-point regime: torch's polynomial in (x, alpha, beta)
template <typename real>
__device__ __forceinline__ real beta_grad_mid_near_mean(real x, real alpha, real beta, real iT) {
  const real total = alpha + beta;
  const real b2 = beta * beta;
  const real poly = real(47) * x * b2 * b2 +
                    alpha * ((real(43) + real(20) * (real(16) + real(27) * beta) * x) * b2 * beta +
                             alpha * (real(3) * (real(59) + real(180) * beta - real(90) * x) * b2 +
                                      alpha * ((real(453) + real(1620) * beta * (real(1) - x) - real(455) * x) * beta +
                                               alpha * (real(8) * (real(1) - x) * (real(135) * beta - real(11))))));
  const real pre_num = (real(1) + real(12) * alpha) * (real(1) + real(12) * beta) * iT * iT;
  const real pre_den = real(12960) * alpha * alpha * alpha * beta * beta * (real(1) + real(12) * total);
  return pre_num * poly / ((real(1) - x) * pre_den);
}

// alpha, beta both large: Rice saddle-point expansion (torch: _beta_grad_alpha_mid).  Same expression as
// torch, algebraically regrouped so that it costs 2 logs, 2 rsqrt and a handful of divisions instead of
// 3 logs, 2 pows, 4 sqrts and ~15 divisions:
//   q = 2ab/T, s = sqrt(q);  prefactor = -x/s;  term3 = 2s/axbx;  term1_den = s T^2 axbx^2 / b;
//   term4 = base^-1.5 = rsqrt(base)^3;  the three Stirling factors share one division.
// It must run in double: term1 and term2*term4 are both O((x-mean)^-2) and cancel to O(1).
template <typename real>
__device__ __forceinline__ real beta_grad_alpha_mid(real x, real alpha, real beta) {
  const real total = alpha + beta;
  const real iT = real(1) / total;
  const real mean = alpha * iT;
  const real dx = x - mean;
  // |x - mean| <= 0.1 std  <=>  dx^2 (T+1) T^2 <= 0.01 a b
  if (dx * dx * (total + real(1)) * total * total <= real(0.01) * alpha * beta)
    return beta_grad_mid_near_mean(x, alpha, beta, iT);
  const real q = real(2) * alpha * beta * iT;
  const real rs = real(1) / Num<real>::sqrt(q);
  const real s = q * rs;
  const real a2 = alpha * alpha, be2 = beta * beta, t2 = total * total;
  // (1 + 1/12a + 1/288a^2)(1 + 1/12b + 1/288b^2) / (1 + 1/12T + 1/288T^2) with one division
  const real stirling = (real(288) * a2 + real(24) * alpha + real(1)) * (real(288) * be2 + real(24) * beta + real(1)) * t2 /
                        (real(288) * a2 * be2 * (real(288) * t2 + real(24) * total + real(1)));
  const real xm1 = x - real(1);
  const real axbx = alpha * xm1 + beta * x;
  const real iax = real(1) / axbx;
  const real term1 = (real(2) * a2 * xm1 + alpha * beta * xm1 - x * be2) * beta * rs * iT * iT * iax * iax;
  const real L1 = Num<real>::log(mean / x);
  const real L2 = Num<real>::log(beta * iT / (real(1) - x));
  const real term3 = real(2) * s * iax;
  const real base = beta * L2 + alpha * L1;
  const real rb = real(1) / Num<real>::sqrt(base);
  const real term4 = rb * rb * rb;
  const real term1234 = term1 + real(0.5) * L1 * (term3 + (x < mean ? term4 : -term4));
  return stirling * (-x * rs) * term1234;
}

// double: torch's expression AS WRITTEN (ATen/native/Distributions.h: _beta_grad_alpha_mid).  Its two leading terms cancel
// from O((x - mean)^-2) to O(1), so a regrouped evaluation -- though algebraically identical -- differs from torch's number at
// the 1e-9 level for large total concentrations (many-allele tiling guides); fp64 parity means the same rounding.
template <>
__device__ __forceinline__ double beta_grad_alpha_mid<double>(double x, double alpha, double beta) {
  const double total = alpha + beta;
  const double mean = alpha / total;
  const double sd = ::sqrt(alpha * beta / (total + 1.0)) / total;
  if (mean - 0.1 * sd <= x && x <= mean + 0.1 * sd) {
    const double poly = 47.0 * x * (beta * beta) * (beta * beta) +
                        alpha * ((43.0 + 20.0 * (16.0 + 27.0 * beta) * x) * (beta * beta) * beta +
                                 alpha * (3.0 * (59.0 + 180.0 * beta - 90.0 * x) * (beta * beta) +
                                          alpha * ((453.0 + 1620.0 * beta * (1.0 - x) - 455.0 * x) * beta +
                                                   alpha * (8.0 * (1.0 - x) * (135.0 * beta - 11.0)))));
    const double prefactor_num = (1.0 + 12.0 * alpha) * (1.0 + 12.0 * beta) / (total * total);
    const double prefactor_den = 12960.0 * alpha * alpha * alpha * beta * beta * (1.0 + 12.0 * total);
    return prefactor_num / (1.0 - x) * poly / prefactor_den;
  }
  const double prefactor = -x / ::sqrt(2.0 * alpha * beta / total);
  const double stirling = (1.0 + 1.0 / (12.0 * alpha) + 1.0 / (288.0 * alpha * alpha)) *
                          (1.0 + 1.0 / (12.0 * beta) + 1.0 / (288.0 * beta * beta)) /
                          (1.0 + 1.0 / (12.0 * total) + 1.0 / (288.0 * total * total));
  const double term1_num = 2.0 * (alpha * alpha) * (x - 1.0) + alpha * beta * (x - 1.0) - x * (beta * beta);
  const double axbx = alpha * (x - 1.0) + beta * x;
  const double term1_den = ::sqrt(2.0 * alpha / beta) * ::pow(total, 1.5) * axbx * axbx;
  const double term1 = term1_num / term1_den;
  const double term2 = 0.5 * ::log(alpha / (total * x));
  const double term3_num = ::sqrt(8.0 * alpha * beta / total);
  const double term3_den = beta * x + alpha * (x - 1.0);
  const double term3 = term3_num / term3_den;
  const double term4_base = beta * ::log(beta / (total * (1.0 - x))) + alpha * ::log(alpha / (total * x));
  const double term4 = ::pow(term4_base, -1.5);
  const double term1234 = term1 + term2 * (term3 + (x < mean ? term4 : -term4));
  return prefactor * stirling * term1234;
}

// torch's rational correction to the asymptotic approximation x (psi(total) - psi(alpha)) / beta (dirichlet_grad_one, table
// c[2][3][3][4]); `dpsi` = psi(total) - psi(alpha)
template <typename real>
__device__ __forceinline__ real dirichlet_grad_rational(real x, real alpha, real total, real dpsi) {
  const real beta = total - alpha;
  // rational-correction coefficients (torch: dirichlet_grad_one, table c[2][3][3][4])
  constexpr double kDirGradC[2][3][3][4] = {
    {{{1.003668233, -0.01061107488, -0.0657888334, 0.01201642863},
      {0.6336835991, -0.3557432599, 0.05486251648, -0.001465281033},
      {-0.03276231906, 0.004474107445, 0.002429354597, -0.0001557569013}},
     {{0.221950385, -0.3187676331, 0.01799915743, 0.01074823814},
      {-0.2951249643, 0.06219954479, 0.01535556598, 0.001550077057},
      {0.02155310298, 0.004170831599, 0.001292462449, 6.976601077e-05}},
     {{-0.05980841433, 0.008441916499, 0.01085618172, 0.002319392565},
      {0.02911413504, 0.01400243777, -0.002721828457, 0.000751041181},
      {0.005900514878, -0.001936558688, -9.495446725e-06, 5.385558597e-05}}},
    {{{1, -0.02924021934, -0.04438342661, 0.007285809825},
      {0.6357567472, -0.3473456711, 0.05454656494, -0.002407477521},
      {-0.03301322327, 0.004845219414, 0.00231480583, -0.0002307248149}},
     {{0.5925320577, -0.1757678135, 0.01505928619, 0.000564515273},
      {0.1014815858, -0.06589186703, 0.01272886114, -0.0007316646956},
      {-0.007258481865, 0.001096195486, 0.0003934994223, -4.12701925e-05}},
     {{0.06469649321, -0.0236701437, 0.002902096474, -5.896963079e-05},
      {0.001925008108, -0.002869809258, 0.0008000589141, -6.063713228e-05},
      {-0.0003477407336, 6.959756487e-05, 1.097287507e-05, -1.650964693e-06}}},
};
  const real u = Num<real>::flog(x);  // float: MUFU lg2 (the logs enter cubic polynomials with O(1) coefficients)
  const real a = Num<real>::flog(alpha) - u;
  const real b = Num<real>::flog(total) - a;
  const real pow_u[3] = {real(1), u, u * u};
  const real pow_a[3] = {real(1), a, a * a};
  real p = real(0), q = real(0);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const real ua = pow_u[i] * pow_a[j];
      p += ua * (real(kDirGradC[0][i][j][0]) + b * (real(kDirGradC[0][i][j][1]) + b * (real(kDirGradC[0][i][j][2]) + b * real(kDirGradC[0][i][j][3]))));
      q += ua * (real(kDirGradC[1][i][j][0]) + b * (real(kDirGradC[1][i][j][1]) + b * (real(kDirGradC[1][i][j][2]) + b * real(kDirGradC[1][i][j][3]))));
    }
  }
  const real approx = x * dpsi / beta;
  return p / q * approx;
}

// -(d/dalpha cdf(x; alpha, total - alpha)) / pdf / (1 - x): what torch._dirichlet_grad evaluates per element.
template <typename real>
__device__ __forceinline__ real dirichlet_grad_one(real x, real alpha, real total) {
  const real beta = total - alpha;
  const real boundary = total * x * (real(1) - x);
  if (x <= real(0.5) && boundary < real(2.5)) return beta_grad_alpha_small(x, alpha, beta);
  if (x >= real(0.5) && boundary < real(0.75)) return -beta_grad_beta_small(real(1) - x, beta, alpha);
  if (alpha > real(6) && beta > real(6)) return beta_grad_alpha_mid(x, alpha, beta);
  return dirichlet_grad_rational<real>(x, alpha, total, digamma_full(total) - digamma_full(alpha));
}

// Single out-of-line copies (code size: the SVI kernel lives or dies by its instruction-cache footprint -- with
// these inlined per call site it reached 235 KB of SASS and stalled on instruction fetch once warps diverged):
//   the full four-regime evaluation in double (double kernels),
//   the three cancellation-free regimes in float (float kernels; the caller has excluded the saddle-point regime),
//   the saddle-point regime alone in double.
static __device__ __noinline__ double dirichlet_grad_one_f64(double x, double alpha, double total) {
  return dirichlet_grad_one<double>(x, alpha, total);
}
// The same with psi(total) - psi(alpha) supplied (replicate-independent: the tiling kernel has it from the Dirichlet sites): the
// boundary series and the rational regime need exactly this difference; the other two regimes do not use digamma at all.
static __device__ __noinline__ double dirichlet_grad_one_f64_psi(double x, double alpha, double total, double psi_total_minus_alpha) {
  const double beta = total - alpha;
  const double boundary = total * x * (1.0 - x);
  if (x <= 0.5 && boundary < 2.5) return beta_grad_alpha_small_f64(x, alpha, beta, -psi_total_minus_alpha);
  if (x >= 0.5 && boundary < 0.75) return -beta_grad_beta_small(1.0 - x, beta, alpha);
  if (alpha > 6.0 && beta > 6.0) return beta_grad_alpha_mid<double>(x, alpha, beta);
  return dirichlet_grad_rational<double>(x, alpha, total, psi_total_minus_alpha);
}
static __device__ __noinline__ float dirichlet_grad_tail_f32(float x, float alpha, float total) {
  return dirichlet_grad_one<float>(x, alpha, total);
}
static __device__ __noinline__ double beta_grad_mid_f64(double x, double alpha, double beta) {
  return beta_grad_alpha_mid<double>(x, alpha, beta);
}

// ---- the two components of a two-allele (Beta) draw --------------------------------------------------------------
// g0 = dirichlet_grad_one(x0, a, a + b), g1 = dirichlet_grad_one(x1, b, a + b).
//
// Regimes.  Once both concentrations exceed 6 nearly every draw lands in torch's saddle-point regime for BOTH
// components; there the two evaluations share everything expensive (`dirichlet_pair_saddle_f64`).  The other three
// regimes (two boundary series, the rational correction) are the minority of draws -- but a 3 % minority already puts
// one in 60 % of the warps, and evaluated in place they cost 40 % of the step: 1-2 active lanes walk through ~1,000
// instructions of cold code per replicate (measured: 1.84 vs 1.11 ms/step with the tails artificially skipped).  So
// the SVI kernel does not evaluate them in place: `dirichlet_pair_is_saddle` routes a draw either to the shared
// saddle-point evaluation or into a per-warp queue in shared memory (`TailQueue`), which is flushed with all lanes busy
// when 32 requests have gathered and once after the replicate loop.

// true  <=>  both components are in the saddle-point regime (torch's dirichlet_grad_one branch order)
__device__ __forceinline__ bool dirichlet_pair_is_saddle(double x0, double x1, double a, double b) {
  const double T = a + b;
  const double bnd0 = T * x0 * (1.0 - x0), bnd1 = T * x1 * (1.0 - x1);
  const bool big = a > 6.0 && b > 6.0;
  const bool mid0 = big && !(x0 <= 0.5 && bnd0 < 2.5) && !(x0 >= 0.5 && bnd0 < 0.75);
  const bool mid1 = big && !(x1 <= 0.5 && bnd1 < 2.5) && !(x1 >= 0.5 && bnd1 < 0.75);
  return mid0 && mid1;
}

// Both components in the saddle-point regime: the roles of (alpha, L1) and (beta, L2) just swap between them, and
// q = 2ab/T, the Stirling factor and term4 = (b L2 + a L1)^-1.5 are symmetric.  Component 1 is evaluated at 1 - x0
// throughout (x1 differs from it by one rounding of the sample: a smooth perturbation of the argument, not amplified by
// the formula's cancellation).
static __device__ __noinline__ void dirichlet_pair_saddle_f64(double x0, double x1, double a, double b, double& g0, double& g1) {
  const double T = a + b;
  const double iT = 1.0 / T;
  const double m0 = a * iT, m1 = b * iT;
  const double d0 = x0 - m0, d1 = x1 - m1;
  // |x - mean| <= 0.1 std: torch switches to a polynomial there (8 % of draws, so nearly every warp has such a
  // lane: it stays inside this path, per component)
  const double lim = 0.01 * a * b, w = (T + 1.0) * T * T;
  const bool near0 = d0 * d0 * w <= lim, near1 = d1 * d1 * w <= lim;
  const double q = 2.0 * a * b * iT;
  const double rs = rsqrt(q);
  const double s = q * rs;
  const double a2 = a * a, b2 = b * b, t2 = T * T;
  const double stirling = (288.0 * a2 + 24.0 * a + 1.0) * (288.0 * b2 + 24.0 * b + 1.0) * t2 /
                          (288.0 * a2 * b2 * (288.0 * t2 + 24.0 * T + 1.0));
  const double La = ool_log(m0 / x0);          // component 0: L1, component 1: L2
  const double Lb = ool_log(m1 / (1.0 - x0));  // component 0: L2, component 1: L1
  const double base = b * Lb + a * La;
  const double rb = rsqrt(base);
  const double term4 = rb * rb * rb;
  const double k = rs * iT * iT;
  {
    const double xm1 = x0 - 1.0;
    const double iax = 1.0 / (a * xm1 + b * x0);
    const double term1 = (2.0 * a2 * xm1 + a * b * xm1 - x0 * b2) * b * k * iax * iax;
    const double t = term1 + 0.5 * La * (2.0 * s * iax + (x0 < m0 ? term4 : -term4));
    g0 = stirling * (-x0 * rs) * t;
  }
  {
    // every x of component 1 must be the SAME number the shared logs saw (1 - x0): term1 and term2 * term4 cancel
    // to O(1) from O((x - mean)^-2), and a sample rounded differently in the two would break that cancellation
    const double y1 = 1.0 - x0, xm1 = -x0;
    const double iax = 1.0 / (b * xm1 + a * y1);
    const double term1 = (2.0 * b2 * xm1 + a * b * xm1 - y1 * a2) * a * k * iax * iax;
    const double t = term1 + 0.5 * Lb * (2.0 * s * iax + (y1 < m1 ? term4 : -term4));
    g1 = stirling * (-y1 * rs) * t;
  }
  if (near0) g0 = beta_grad_mid_near_mean(x0, a, b, iT);
  if (near1) g1 = beta_grad_mid_near_mean(x1, b, a, iT);
}

// The same, split into the part that depends on the guide's concentrations only (hoisted out of the replicate loop by
// `svi_alpha_kernel`: one division, one rsqrt and ~25 multiplies per guide instead of per draw) and the per-draw part.
struct SaddlePair {
  double a, b, iT, m0, m1, lim_w, rs, s, stirling, k, a2, b2;
  __device__ __forceinline__ void prepare_near_mean(float*, int) {}
  __device__ __forceinline__ void init(double a_, double b_) {
    a = a_; b = b_;
    const double T = a + b;
    iT = 1.0 / T;
    m0 = a * iT; m1 = b * iT;
    lim_w = 0.01 * a * b / ((T + 1.0) * T * T);  // near the mean <=> d^2 <= lim_w
    const double q = 2.0 * a * b * iT;
    rs = rsqrt(q);
    s = q * rs;
    a2 = a * a; b2 = b * b;
    const double t2 = T * T;
    stirling = (288.0 * a2 + 24.0 * a + 1.0) * (288.0 * b2 + 24.0 * b + 1.0) * t2 /
               (288.0 * a2 * b2 * (288.0 * t2 + 24.0 * T + 1.0));
    k = rs * iT * iT;
  }
  __device__ __forceinline__ void eval(double x0, double x1, double& g0, double& g1) const {
    const double d0 = x0 - m0, d1 = x1 - m1;
    const double La = ool_log(m0 / x0), Lb = ool_log(m1 / (1.0 - x0));
    const double base = b * Lb + a * La;
    const double rb = rsqrt(base);
    const double term4 = rb * rb * rb;
    {
      const double xm1 = x0 - 1.0;
      const double iax = 1.0 / (a * xm1 + b * x0);
      const double term1 = (2.0 * a2 * xm1 + a * b * xm1 - x0 * b2) * b * k * iax * iax;
      g0 = stirling * (-x0 * rs) * (term1 + 0.5 * La * (2.0 * s * iax + (x0 < m0 ? term4 : -term4)));
    }
    {
      const double y1 = 1.0 - x0, xm1 = -x0;  // component 1 at 1 - x0 throughout (see dirichlet_pair_saddle_f64)
      const double iax = 1.0 / (b * xm1 + a * y1);
      const double term1 = (2.0 * b2 * xm1 + a * b * xm1 - y1 * a2) * a * k * iax * iax;
      g1 = stirling * (-y1 * rs) * (term1 + 0.5 * Lb * (2.0 * s * iax + (y1 < m1 ? term4 : -term4)));
    }
    if (d0 * d0 <= lim_w) g0 = beta_grad_mid_near_mean(x0, a, b, iT);
    if (d1 * d1 <= lim_w) g1 = beta_grad_mid_near_mean(x1, b, a, iT);
  }
};

// ---- the saddle-point regime in FLOAT ------------------------------------------------------------------------------
// torch's expression cancels twice: term1 against term2 * term4 at O(delta^-2), and what is left against the rest of
// term1 at O(delta^-1) (delta = x - mean).  Both cancellations are removed analytically here, so single precision is
// enough (worst 9e-7 relative against torch's double evaluation over alpha, beta in [6.2, 5000], |z| in [0.1, 8],
// tests/test_saddle_float_form.py -- given delta itself to 1e-7, see eval()).  With m = alpha / T, u = delta / m,
// v = -delta / (1 - m):
//   log(m / x) = -u l(u),                      l(y) = log1p(y) / y           = 1 + l1(y)
//   T KL(m || x) = T delta^2 / (2 m (1 - m)) G,  G = (1 - m) g(u) + m g(v),  g(y) = 2 (y - log1p(y)) / y^2 = 1 + g1(y)
//   term1 + term2 (+-term4) = C2 D / delta^2 + beta (2 alpha - beta) / (s T^3 delta),   D = l(u) G^-1.5 - 1,
//   C2 = 2 alpha beta^2 / (s T^4), s = sqrt(2 alpha beta / T); the 1 / delta term equals -C2 c1 / delta with
//   c1 = T (2 alpha - beta) / (2 alpha beta) and D = -c1 delta + O(delta^2), so
//   term1234 = C2 (D + c1 delta) / delta^2 - s l(u) / alpha,      D = (1 + l1(u)) (1 + [(1 + G1)^-1.5 - 1]) - 1.
// l1 and g1 are series for small |y| (no cancellation), the defining expressions otherwise.  The second component is the
// same with alpha <-> beta, u <-> v, delta -> -delta, and shares g1(u), g1(v), G.
// (l1, g1)(y) = (log1p(y) / y - 1,  2 (y - log1p(y)) / y^2 - 1): series for small |y|, one shared log1p otherwise
__device__ __forceinline__ void saddle_l1_g1(float y, float& l1, float& g1) {
  if (fabsf(y) < 0.3f) {
    // ONE series: g1 = y q(y), q = -2/3 + 2y/4 - 2y^2/5 ...; and log1p(y)/y = 1 - y/2 (1 + g1) identically, so l1 = -y/2 (1 + g1)
    float q = 2.0f / 14.0f;
#pragma unroll
    for (int k = 13; k >= 3; --k) q = fmaf(q, y, (k & 1) ? -2.0f / k : 2.0f / k);
    g1 = q * y;
    l1 = -0.5f * y * (1.0f + g1);
  } else {
    const float lp = log1pf(y), iy = rcp_ftz(y);
    l1 = lp * iy - 1.0f;
    g1 = 2.0f * (y - lp) * iy * iy - 1.0f;
  }
}

// (1 + z)^-1.5 - 1 without cancellation and without libm: with q = sqrt(1 + z), e = 1/q - 1 = -z / (q (1 + q)) and
// (1 + e)^3 - 1 = e (3 + e (3 + e)).  ~12 instructions against ~90 for expm1f(-1.5f * log1pf(z)); same worst error of the
// assembled gradient (tests/test_saddle_float_form.py).
__device__ __forceinline__ float pow_m15_minus1(float z) {
  const float q = sqrtf(1.0f + z);
  const float e = -z * rcp_ftz(q * (1.0f + q));
  return e * (3.0f + e * (3.0f + e));
}

// torch's near-mean polynomial (beta_grad_mid_near_mean) is LINEAR in x: with the prefactor folded in it is
// (k0 + kx x + k1 (1 - x)) / (1 - x), k0, kx, k1 functions of the concentrations only (all three groups free of cancellation for
// alpha, beta > 6: 5e-7 against torch's double evaluation, tests/test_saddle_float_form.py).
__device__ __forceinline__ void near_mean_coeffs(float al, float be, float T, float& k0, float& kx, float& k1) {
  const float b2 = be * be, b3 = b2 * be, a3 = al * al * al;
  const float c0 = al * (43.0f * b3 + al * (3.0f * (59.0f + 180.0f * be) * b2 + al * 453.0f * be));
  const float cx = 47.0f * b2 * b2 + al * (20.0f * (16.0f + 27.0f * be) * b3 - al * (270.0f * b2 + al * 455.0f * be));
  const float c1 = a3 * (1620.0f * b2 + al * 8.0f * (135.0f * be - 11.0f));
  const float K = (1.0f + 12.0f * al) * (1.0f + 12.0f * be) / (T * T) / (12960.0f * a3 * b2 * (1.0f + 12.0f * T));
  k0 = c0 * K; kx = cx * K; k1 = c1 * K;
}

struct SaddlePairF {
  float a, b, iT, m, om, im, iom, lim_w, s, is, stirling, C2a, C2b, c1a, c1b, ia, ib;
  const float* nm;  // this thread's six near-mean coefficients in shared memory, stride nm_stride (prepare_near_mean)
  int nm_stride;
  // The near-mean regime holds 8 % of the draws, so nearly every warp meets it in every replicate with 2-3 lanes active: the
  // polynomial's ~70 instructions were 13 % of the alpha kernel.  Its coefficients are made once per guide with all lanes busy
  // and parked in shared memory; a draw then costs one reciprocal and three FMAs per component.
  __device__ __forceinline__ void prepare_near_mean(float* smem, int stride) {
    nm = smem; nm_stride = stride;
    if (a > 6.0f && b > 6.0f) {
      float k0, kx, k1;
      near_mean_coeffs(a, b, a + b, k0, kx, k1);
      smem[0] = k0; smem[stride] = kx; smem[2 * stride] = k1;
      near_mean_coeffs(b, a, a + b, k0, kx, k1);
      smem[3 * stride] = k0; smem[4 * stride] = kx; smem[5 * stride] = k1;
    }
  }
  __device__ __forceinline__ void init(float a_, float b_) {
    a = a_; b = b_;
    const float T = a + b;
    iT = 1.0f / T;
    m = a * iT; om = b * iT;
    im = T / a; iom = T / b;
    ia = 1.0f / a; ib = 1.0f / b;
    lim_w = 0.01f * a * b / ((T + 1.0f) * T * T);
    const float q = 2.0f * a * b * iT;
    is = rsqrtf(q);
    s = q * is;
    const float a2 = a * a, b2 = b * b, t2 = T * T;
    stirling = (288.0f * a2 + 24.0f * a + 1.0f) * (288.0f * b2 + 24.0f * b + 1.0f) * t2 /
               (288.0f * a2 * b2 * (288.0f * t2 + 24.0f * T + 1.0f));
    const float k = 2.0f * is * iT * iT * iT * iT;  // 2 / (s T^4)
    C2a = k * a * b2;
    C2b = k * b * a2;
    const float h = 0.5f * T * ia * ib;             // T / (2 alpha beta)
    c1a = h * (2.0f * a - b);
    c1b = h * (2.0f * b - a);
  }
  __device__ __forceinline__ void eval(float x0, float x1, float& g0, float& g1) const {
    // delta = x0 - a / T.  Formed as x0 - m it inherits the rounding of m: 6e-8 m, which for a lopsided pair (m ~ 1, std
    // ~ 1e-3) is 1e-4 of delta -- and the result is linear in delta (this was the whole 3e-5 worst case of round 1).
    // Instead delta = (x0 b - (1 - x0) a) / T with error-free products and the exact split 1 - x0 = hi + lo: ~1e-7 relative.
    const float hi = 1.0f - x0, lo = (1.0f - hi) - x0;
    const float pa = hi * a, ea = fmaf(hi, a, -pa);
    const float d = fmaf(-lo, a, fmaf(x0, b, -pa) - ea) * iT;
    if (d * d <= lim_w) {  // |x - mean| <= 0.1 std: torch's polynomial, from the per-guide coefficients
      const float o0 = 1.0f - x0, o1 = 1.0f - x1;
      g0 = fmaf(nm[2 * nm_stride], o0, fmaf(nm[nm_stride], x0, nm[0])) * rcp_ftz(o0);
      g1 = fmaf(nm[5 * nm_stride], o1, fmaf(nm[4 * nm_stride], x1, nm[3 * nm_stride])) * rcp_ftz(o1);
      return;
    }
    const float u = d * im, v = -d * iom;
    float l1u, l1v, g1u, g1v;
    saddle_l1_g1(u, l1u, g1u);
    saddle_l1_g1(v, l1v, g1v);
    const float wm1 = pow_m15_minus1(om * g1u + m * g1v);                   // G^-1.5 - 1
    const float D0 = l1u + wm1 + l1u * wm1, D1 = l1v + wm1 + l1v * wm1;    // (1 + l1)(1 + wm1) - 1
    const float id2 = rcp_ftz(d * d);
    const float t0 = C2a * (D0 + c1a * d) * id2 - s * (1.0f + l1u) * ia;
    const float t1 = C2b * (D1 - c1b * d) * id2 - s * (1.0f + l1v) * ib;
    g0 = stirling * (-x0 * is) * t0;
    g1 = stirling * (-(1.0f - x0) * is) * t1;  // component 1 at 1 - x0, like the double path
  }
};

// One component, any regime.  FLOAT_TAILS (the float kernels): the three regimes without cancellation are evaluated in
// float (1e-6 relative, against the fp32 path's 2e-4 budget for this gradient); the saddle-point regime stays double.
template <bool FLOAT_TAILS>
__device__ __forceinline__ double dirichlet_grad_any(double x, double alpha, double beta) {
  if (!FLOAT_TAILS) return dirichlet_grad_one_f64(x, alpha, alpha + beta);
  const double T = alpha + beta, bnd = T * x * (1.0 - x);
  const bool mid = alpha > 6.0 && beta > 6.0 && !(x <= 0.5 && bnd < 2.5) && !(x >= 0.5 && bnd < 0.75);
  return mid ? beta_grad_mid_f64(x, alpha, beta) : (double)dirichlet_grad_tail_f32((float)x, (float)alpha, (float)T);
}

#ifndef BEAN_FLOAT_TAILS
#define BEAN_FLOAT_TAILS 1
#endif

// Per-warp queue of deferred (non-saddle) draws.  Entry: the draw (x0, x1), the guide's concentrations (a, b), the
// upstream weights (w0, w1) = (go - gbar) of the two components and the owning lane.  `flush` evaluates the requests
// with every lane of the warp busy, then each owner adds its own results in queue order (deterministic).
template <typename real>
struct TailQueue {
  static constexpr int CAP = 64;  // flushed as soon as it holds >= 32, a push adds <= 32
  real x0[CAP], x1[CAP], a[CAP], b[CAP], w0[CAP], w1[CAP];
  int owner[CAP];
};

template <typename real>
__device__ __forceinline__ int tail_queue_push(TailQueue<real>& q, int count, unsigned wmask, int lane, bool need, real x0, real x1,
                                               real a, real b, real w0, real w1) {
  const unsigned m = __ballot_sync(wmask, need);
  if (need) {
    const int j = count + __popc(m & ((1u << lane) - 1u));
    q.x0[j] = x0; q.x1[j] = x1; q.a[j] = a; q.b[j] = b; q.w0[j] = w0; q.w1[j] = w1;
    q.owner[j] = lane;
  }
  return count + __popc(m);
}

template <typename real>
static __device__ __noinline__ void tail_queue_flush(TailQueue<real>& q, int n, unsigned wmask, int lane, real& acc0, real& acc1) {
  __syncwarp(wmask);
  const int rank = __popc(wmask & ((1u << lane) - 1u)), width = __popc(wmask);
  for (int j = rank; j < n; j += width) {
    const double x0 = (double)q.x0[j], x1 = (double)q.x1[j], a = (double)q.a[j], b = (double)q.b[j];
    // float kernels: the three cancellation-free regimes in float.  BEAN_FLOAT_TAILS = 0 (double) was measured on B200: the
    // alpha kernel goes from 0.185 to 0.291 ms per step at c5 while the worst alpha_pi error of the golden cases only moves
    // from 3.7e-5 to 3.0e-5 (it is not in these regimes), so float stays
    const double g0 = dirichlet_grad_any<(sizeof(real) == 4) && BEAN_FLOAT_TAILS>(x0, a, b);
    const double g1 = dirichlet_grad_any<(sizeof(real) == 4) && BEAN_FLOAT_TAILS>(x1, b, a);
    q.w0[j] = real(g0 * (double)q.w0[j]);
    q.w1[j] = real(g1 * (double)q.w1[j]);
  }
  __syncwarp(wmask);
  for (int j = 0; j < n; ++j)
    if (q.owner[j] == lane) {
      acc0 += q.w0[j];
      acc1 += q.w1[j];
    }
  __syncwarp(wmask);
}

}  // namespace bean
