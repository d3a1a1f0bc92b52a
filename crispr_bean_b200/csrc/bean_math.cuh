// bean_math.cuh -- in-register special functions for the bean SVI kernels (sm_100a).
//
// These replace the torch.lgamma / torch.digamma / erf sweeps behind pyro's
// DirichletMultinomial.log_prob, its backward and get_std_normal_prob
// (reference call sites: bean/model/model.py:138-164, :531-547; bean/model/utils.py:34-76).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace bean {

// MUFU reciprocal / log2 in their flush-to-zero form.  Without -ftz nvcc wraps every `__fdividef` / `__logf` in denormal
// pre-scaling (FSETP + FMUL 2^24 + FSEL ... ~4 extra instructions each); their arguments here are never denormal
// (concentrations >= 1e-5, counts, sums of those), and an issue-bound kernel pays for every instruction: this removed
// a fifth of the per-bin instruction count.
__device__ __forceinline__ float rcp_ftz(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float log_ftz(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * 0.69314718055994530942f;
}

// Library functions with long bodies are called through ONE out-of-line copy each: the SVI kernel is bound by
// instruction issue AND by its instruction-cache footprint, and a call costs far less than a duplicated body.
static __device__ __noinline__ float ool_logf(float x) { return logf(x); }
static __device__ __noinline__ float ool_log1pf(float x) { return log1pf(x); }
static __device__ __noinline__ float ool_powf(float x, float y) { return powf(x, y); }
static __device__ __noinline__ float ool_erfcf(float x) { return erfcf(x); }
static __device__ __noinline__ double ool_log(double x) { return ::log(x); }

template <typename real> struct Num;
template <> struct Num<float> {
  static __device__ __forceinline__ float log(float x) { return ool_logf(x); }
  static __device__ __forceinline__ float exp(float x) { return expf(x); }
  static __device__ __forceinline__ float erf(float x) { return erff(x); }
  static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float lgamma(float x) { return lgammaf(x); }
  static __device__ __forceinline__ float pow(float x, float y) { return ool_powf(x, y); }
  static __device__ __forceinline__ float log1p(float x) { return ool_log1pf(x); }
  static __device__ __forceinline__ float log1p_inl(float x) { return log1pf(x); }
  static __device__ __forceinline__ float fmax(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ float fabs(float a) { return fabsf(a); }
  static __device__ __forceinline__ float fmin(float a, float b) { return fminf(a, b); }
  static __device__ __forceinline__ float rcp(float x) { return rcp_ftz(x); }  // MUFU.RCP, ~1 ulp
  // 2-ulp division (MUFU.RCP + FMUL) for quantities whose own rounding already dominates
  static __device__ __forceinline__ float div(float a, float b) { return a * rcp_ftz(b); }
  // 3-ulp log (MUFU.LG2 + FMUL): only where the value is multiplied by O(1) factors
  static __device__ __forceinline__ float flog(float x) { return log_ftz(x); }
};
template <> struct Num<double> {
  static __device__ __forceinline__ double log(double x) { return ool_log(x); }
  static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
  static __device__ __forceinline__ double erf(double x) { return ::erf(x); }
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ double lgamma(double x) { return ::lgamma(x); }
  static __device__ __forceinline__ double pow(double x, double y) { return ::pow(x, y); }
  static __device__ __forceinline__ double log1p(double x) { return ::log1p(x); }
  static __device__ __forceinline__ double log1p_inl(double x) { return ::log1p(x); }
  static __device__ __forceinline__ double fmax(double a, double b) { return ::fmax(a, b); }
  static __device__ __forceinline__ double fabs(double a) { return ::fabs(a); }
  static __device__ __forceinline__ double fmin(double a, double b) { return ::fmin(a, b); }
  static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
  static __device__ __forceinline__ double div(double a, double b) { return a / b; }
  static __device__ __forceinline__ double flog(double x) { return ::log(x); }
};

// ---- Gamma-function corrections to the Stirling main part, z > 0 ---------------------------------
// gamma_corr(z) returns
//   cv = lgamma(z)  - [(z - 1/2) ln z - z + ln(2 pi)/2]
//   dl = digamma(z) - ln z
// The Dirichlet-Multinomial row is assembled from these in "KL form" (bean_row.cuh): every large
// z ln z product cancels ANALYTICALLY and only logs of near-1 ratios remain, so float keeps ~1e-6
// absolute accuracy on a row whose individual lgamma terms are O(1e3..1e4) (a float lgamma(300) alone
// is already off by 1e-4, which is the noise floor of the reference's own fp32 torch.lgamma path).
// float, z >= 4: pure series in 1/z (no log at all); z < 4: 4-step upward recurrence first.
__device__ __forceinline__ void gamma_corr(float z, float& cv, float& dl) {
  float zs = z;
  if (z < 4.0f) zs = z + 4.0f;
  const float rz = rcp_ftz(zs);
  const float r2 = rz * rz;
  // lgamma tail: 1/(12 z) - 1/(360 z^3) + 1/(1260 z^5) - 1/(1680 z^7)
  cv = rz * (8.3333333333e-2f + r2 * (-2.7777777778e-3f + r2 * (7.9365079365e-4f + r2 * -5.9523809524e-4f)));
  // digamma tail: -1/(2z) - 1/(12 z^2) + 1/(120 z^4) - 1/(252 z^6) + 1/(240 z^8)
  dl = -0.5f * rz - r2 * (8.3333333333e-2f + r2 * (-8.3333333333e-3f + r2 * (3.9682539683e-3f + r2 * -4.1666666667e-3f)));
  if (z < 4.0f) {
    // Gamma(z) = Gamma(z+4) / P(z), P = z (z+1) (z+2) (z+3);  psi(z) = psi(z+4) - P'(z)/P(z).
    //   cv(z) = cv(z+4) + (z + 1/2) ln(1 + 4/z) + ln((z+4)^3 / ((z+1)(z+2)(z+3))) - 4
    //   dl(z) = dl(z+4) + ln(1 + 4/z) - P'/P
    // Both logs are multiplied by factors <= 4.5, so the 3-ulp MUFU log is accurate enough here.
    const float z1 = z + 1.0f, z2 = z + 2.0f, z3 = z + 3.0f;
    const float p123 = z1 * z2 * z3;
    const float iz = rcp_ftz(z), ip = rcp_ftz(p123);
    const float ls = log_ftz(zs * iz);
    const float lr = log_ftz(zs * zs * zs * ip);
    cv += (z + 0.5f) * ls + lr - 4.0f;
    // P'/P = 1/z + ((z+1)(z+2) + (z+1)(z+3) + (z+2)(z+3)) / ((z+1)(z+2)(z+3))
    dl += ls - (iz + (z1 * z2 + z1 * z3 + z2 * z3) * ip);
  }
}

__device__ __forceinline__ double digamma_f64(double z) {
  // recurrence up to z >= 10, then the asymptotic series (7 Bernoulli terms)
  double sub = 0.0, zs = z;
  while (zs < 10.0) {
    sub += 1.0 / zs;
    zs += 1.0;
  }
  const double rz = 1.0 / zs, r2 = rz * rz;
  const double s = r2 * (1.0 / 12 + r2 * (-1.0 / 120 + r2 * (1.0 / 252 + r2 * (-1.0 / 240 + r2 * (1.0 / 132 + r2 * (-691.0 / 32760 + r2 * (1.0 / 12)))))));
  return ::log(zs) - 0.5 * rz - s - sub;
}

__device__ __forceinline__ void gamma_corr(double z, double& cv, double& dl) {
  const double lz = ::log(z);
  cv = ::lgamma(z) - ((z - 0.5) * lz - z + 0.91893853320467274178);
  dl = digamma_f64(z) - lz;
}

// full lgamma / digamma of one argument (off the hot row loop: Dirichlet normalisers)
template <typename real>
__device__ __noinline__ void lgamma_digamma(real z, real& lg, real& dg) {
  real cv, dl;
  gamma_corr(z, cv, dl);
  const real lz = Num<real>::log(z);
  lg = (z - real(0.5)) * lz - z + real(0.91893853320467274178) + cv;
  dg = lz + dl;
}

// ---- Normal CDF pieces ----------------------------------------------------------------------------
template <typename real>
__device__ __forceinline__ real std_normal_pdf(real z) {
  return real(0.39894228040143267794) * Num<real>::exp(real(-0.5) * z * z);
}

// Probability mass of one sorting bin and its derivatives w.r.t. (mu, sd).
// thr_u = +inf / thr_l = -inf encode quantile 1 / 0 (model/utils.py:48-54, :60-72).
// double: exactly the reference expression, Normal.cdf = 0.5 (1 + erf(z / sqrt 2)).
__device__ __forceinline__ void bin_prob_sorting(double thr_u, double thr_l, double mu, double sd, double& P,
                                                 double& dP_dmu, double& dP_dsd) {
  const double rs = 1.0 / sd;
  double cu = 1.0, fu = 0.0, zfu = 0.0, cl = 0.0, fl = 0.0, zfl = 0.0;
  if (!isinf(thr_u)) {
    const double z = (thr_u - mu) * rs;
    cu = 0.5 * (1.0 + ::erf(z * 0.70710678118654752440));
    fu = std_normal_pdf(z);
    zfu = z * fu;
  }
  if (!isinf(thr_l)) {
    const double z = (thr_l - mu) * rs;
    cl = 0.5 * (1.0 + ::erf(z * 0.70710678118654752440));
    fl = std_normal_pdf(z);
    zfl = z * fl;
  }
  P = cu - cl;
  dP_dmu = -(fu - fl) * rs;
  dP_dsd = -(zfu - zfl) * rs;
}

// float: the same mass written with erfc on the tail side, Q(z) = erfc(z / sqrt 2) / 2, so that a bin far
// in a tail keeps its RELATIVE accuracy (0.5 (1 + erf) in float loses it below ~1e-3 and that error is
// amplified by a0 / sum(p) downstream).
static __device__ __noinline__ void bin_prob_sorting(float thr_u, float thr_l, float mu, float sd, float& P,
                                                 float& dP_dmu, float& dP_dsd) {
  const float rs = 1.0f / sd;
  const bool hu = !isinf(thr_u), hl = !isinf(thr_l);
  const float zu = hu ? (thr_u - mu) * rs : INFINITY;
  const float zl = hl ? (thr_l - mu) * rs : -INFINITY;
  const float k = 0.70710678118654752440f;
  if (zl > 0.0f)
    P = 0.5f * (ool_erfcf(zl * k) - ool_erfcf(zu * k));
  else
    P = 0.5f * (ool_erfcf(-zu * k) - ool_erfcf(-zl * k));
  const float fu = hu ? std_normal_pdf(zu) : 0.0f, fl = hl ? std_normal_pdf(zl) : 0.0f;
  const float zfu = hu ? zu * fu : 0.0f, zfl = hl ? zl * fl : 0.0f;
  dP_dmu = -(fu - fl) * rs;
  dP_dsd = -(zfu - zfl) * rs;
}

// ---- block reduction of a double (deterministic order) ------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// all threads must call; result valid in thread 0.  blockDim.x <= 1024.
__device__ __forceinline__ double block_sum(double v, double* smem32) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) smem32[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? smem32[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

}  // namespace bean
