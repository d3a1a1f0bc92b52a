// bean_math.cuh -- in-register special functions for the bean SVI kernels (sm_100a).
//
// lgamma/digamma are evaluated as a PAIR sharing log(z) and 1/z (Stirling series, with a fixed
// 4-step upward recurrence for z < 4 in float so the branch is two-way only).  These replace the
// torch.lgamma / torch.digamma sweeps behind pyro's DirichletMultinomial.log_prob and its backward
// (reference call sites: bean/model/model.py:138-164, :531-547).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace bean {

template <typename real> struct Num;
template <> struct Num<float> {
  static __device__ __forceinline__ float log(float x) { return logf(x); }
  static __device__ __forceinline__ float exp(float x) { return expf(x); }
  static __device__ __forceinline__ float erf(float x) { return erff(x); }
  static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float lgamma(float x) { return lgammaf(x); }
  static __device__ __forceinline__ float pow(float x, float y) { return powf(x, y); }
  static __device__ __forceinline__ float log1p(float x) { return log1pf(x); }
  static __device__ __forceinline__ float fmax(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ float fmin(float a, float b) { return fminf(a, b); }
  static __device__ __forceinline__ float rcp(float x) { return 1.0f / x; }
};
template <> struct Num<double> {
  static __device__ __forceinline__ double log(double x) { return ::log(x); }
  static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
  static __device__ __forceinline__ double erf(double x) { return ::erf(x); }
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ double lgamma(double x) { return ::lgamma(x); }
  static __device__ __forceinline__ double pow(double x, double y) { return ::pow(x, y); }
  static __device__ __forceinline__ double log1p(double x) { return ::log1p(x); }
  static __device__ __forceinline__ double fmax(double a, double b) { return ::fmax(a, b); }
  static __device__ __forceinline__ double fmin(double a, double b) { return ::fmin(a, b); }
  static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
};

// ---- lgamma + digamma pair, z > 0 --------------------------------------------------------------
// float: |err| ~ 1e-7 relative to max(1, |value|) for z >= 1e-6 (checked in tests/test_gpu_math.py).
__device__ __forceinline__ void lgamma_digamma(float z, float& lg, float& dg) {
  float zs = z, sub_lg = 0.0f, sub_dg = 0.0f;
  if (z < 4.0f) {
    // Gamma(z) = Gamma(z+4) / (z (z+1) (z+2) (z+3));  psi(z) = psi(z+4) - P'(z)/P(z)
    const float z1 = z + 1.0f, z2 = z + 2.0f, z3 = z + 3.0f;
    const float p01 = z * z1, p23 = z2 * z3;
    const float P = p01 * p23;
    const float dP = (z + z1) * p23 + p01 * (z2 + z3);
    sub_lg = logf(P);
    sub_dg = dP / P;
    zs = z + 4.0f;
  }
  const float lz = logf(zs);
  const float rz = 1.0f / zs;
  const float r2 = rz * rz;
  // Stirling: lgamma(z) = (z-1/2) ln z - z + ln(2 pi)/2 + 1/(12 z) - 1/(360 z^3) + 1/(1260 z^5) - 1/(1680 z^7)
  const float s_lg = rz * (8.3333333333e-2f + r2 * (-2.7777777778e-3f + r2 * (7.9365079365e-4f + r2 * -5.9523809524e-4f)));
  lg = (zs - 0.5f) * lz - zs + 0.91893853320467274f + s_lg - sub_lg;
  // psi(z) = ln z - 1/(2z) - 1/(12 z^2) + 1/(120 z^4) - 1/(252 z^6) + 1/(240 z^8)
  const float s_dg = r2 * (8.3333333333e-2f + r2 * (-8.3333333333e-3f + r2 * (3.9682539683e-3f + r2 * -4.1666666667e-3f)));
  dg = lz - 0.5f * rz - s_dg - sub_dg;
}

__device__ __forceinline__ void lgamma_digamma(double z, double& lg, double& dg) {
  lg = ::lgamma(z);
  // digamma: recurrence up to z >= 10, then the asymptotic series (7 Bernoulli terms)
  double sub = 0.0, zs = z;
  while (zs < 10.0) {
    sub += 1.0 / zs;
    zs += 1.0;
  }
  const double rz = 1.0 / zs, r2 = rz * rz;
  const double s = r2 * (1.0 / 12 + r2 * (-1.0 / 120 + r2 * (1.0 / 252 + r2 * (-1.0 / 240 + r2 * (1.0 / 132 + r2 * (-691.0 / 32760 + r2 * (1.0 / 12)))))));
  dg = ::log(zs) - 0.5 * rz - s - sub;
}

// lgamma(1 + x) for a non-negative integer-valued count (data-only term of the DM log-pmf)
template <typename real>
__device__ __forceinline__ real lgamma1p_count(real x) {
  return Num<real>::lgamma(x + real(1));
}

// ---- Normal CDF pieces (reference: torch Normal.cdf = 0.5 * (1 + erf((x - mu) / sd / sqrt(2)))) ---
template <typename real>
__device__ __forceinline__ real std_normal_cdf(real z) {
  return real(0.5) * (real(1) + Num<real>::erf(z * real(0.70710678118654752440)));
}
template <typename real>
__device__ __forceinline__ real std_normal_pdf(real z) {
  return real(0.39894228040143267794) * Num<real>::exp(real(-0.5) * z * z);
}

// Probability mass of one sorting bin and its derivatives w.r.t. (mu, sd).
// thr_u = +inf / thr_l = -inf encode quantile 1 / 0 (model/utils.py:48-54, :60-72).
template <typename real>
__device__ __forceinline__ void bin_prob_sorting(real thr_u, real thr_l, real mu, real sd, real& P,
                                                 real& dP_dmu, real& dP_dsd) {
  const real rs = real(1) / sd;
  real cu = real(1), fu = real(0), zfu = real(0);
  real cl = real(0), fl = real(0), zfl = real(0);
  if (!isinf(thr_u)) {
    const real z = (thr_u - mu) * rs;
    cu = std_normal_cdf(z);
    fu = std_normal_pdf(z);
    zfu = z * fu;
  }
  if (!isinf(thr_l)) {
    const real z = (thr_l - mu) * rs;
    cl = std_normal_cdf(z);
    fl = std_normal_pdf(z);
    zfl = z * fl;
  }
  P = cu - cl;
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
