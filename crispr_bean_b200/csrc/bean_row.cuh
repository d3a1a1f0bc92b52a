// bean_row.cuh -- one Dirichlet-Multinomial row: value and digamma differences in "KL form".
//
// Replaces, per (replicate, guide, layer) row, pyro's
//   DirichletMultinomial(a).log_prob(x) = lgG(A) + lgG(N+1) - lgG(N+A) - sum_b [lgG(x_b+1) + lgG(a_b) - lgG(x_b+a_b)]
// and the digamma differences autograd would produce, written so that no large
// term is ever formed.  With u_b = x_b + a_b, U = N + A, num_b = x_b A - a_b N (computed exactly via FMA):
//
//   L2_b = ln(u_b A / (a_b U)) = -log1p(-num_b / (u_b A))         L1_b = ln(u_b N / (U x_b)) = -log1p(num_b / (u_b N))
//   V    = sum_b [x_b L1_b + a_b L2_b] - 1/2 [sum_b L2_b + (B-1) ln(U/A)] + sum_b [cv(u_b) - cv(a_b)] - [cv(U) - cv(A)]
//   gb   = psi(A) - psi(U) + psi(u_b) - psi(a_b) = L2_b + [dl(A) - dl(U)] + [dl(u_b) - dl(a_b)]
//   log_prob = V + K,   K = lgG(N+1) - sum_b lgG(x_b+1) + sum_{x_b>0} x_b ln(x_b / N)      (data only, hoisted)
//
// (cv, dl) = gamma_corr: the parts of lgamma / digamma beyond the Stirling main term (bean_math.cuh).  Both ratios are
// written over u_b because 1/u_b, 1/A and 1/U fall out of gamma_corr for free and 1/N is one reciprocal per row: a bin
// costs 4 MUFU operations (1/a_b, 1/u_b and one per log1p) and ~75 FP32 instructions, everything inlined.
#pragma once
#include "bean_math.cuh"

namespace bean {

template <typename real> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<double> { typedef double4 type; };
template <typename real> struct Vec2;
template <> struct Vec2<float> { typedef float2 type; };
template <> struct Vec2<double> { typedef double2 type; };

// x, a: counts and concentrations of the row (a_b > 0, N = sum x > 0); nb = number of bins actually used (<= NB).
// Returns V; gb[b] receives the digamma difference of bin b.
template <typename real, int NB>
__device__ __forceinline__ real dm_row_kl_scalar(int nb, const real (&x)[NB], const real (&a)[NB], real N, real A,
                                                 real (&gb)[NB]) {
  const real U = N + A;
  real cvA, dlA, iA, cvU, dlU, iU;
  gamma_corr(A, cvA, dlA, iA);
  gamma_corr(U, cvU, dlU, iU);
  const real iN = Num<real>::rcp(N);
  const real lUA = Num<real>::flog(U * iA);  // enters V with weight (B-1)/2 only
  const real dlAU = dlA - dlU;
  real V = real(0), sumL2 = real(0), csum = real(0);
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    if (b < nb) {
      const real u = x[b] + a[b];
      real cva, dla, ia, cvu, dlu, iu;
      gamma_corr(a[b], cva, dla, ia);
      gamma_corr(u, cvu, dlu, iu);
      // num = x A - a N without cancellation error: p + e == a N exactly
      const real p = a[b] * N;
      const real e = fma(a[b], N, -p);
      const real num = fma(x[b], A, -p) - e;
      const real t = num * iu;
      const real Uu = U * iu;
      const real L2 = -log1p_ratio(-t * iA, a[b] * iA * Uu);  // -ln(a U / (u A))
      const real L1 = -log1p_ratio(t * iN, x[b] * iN * Uu);   // -ln(x U / (u N)); x = 0: +inf, discarded below
      V += x[b] > real(0) ? fma(x[b], L1, a[b] * L2) : a[b] * L2;
      sumL2 += L2;
      csum += cvu - cva;
      gb[b] = L2 + (dlAU + (dlu - dla));
    } else {
      gb[b] = real(0);
    }
  }
  V += real(-0.5) * (sumL2 + real(nb - 1) * lUA) + (csum - (cvU - cvA));
  return V;
}

// float: the same row through the packed (two-at-a-time) functions of bean_math.cuh -- gamma_corr of (a_b, u_b) and of
// (A, U), the bin's two log1p.  Bit-identical to the generic template above evaluated in float.
template <int NB>
__device__ __forceinline__ float dm_row_kl_packed(int nb, const float (&x)[NB], const float (&a)[NB], float N, float A,
                                                  float (&gb)[NB]) {
  const float U = N + A;
  float2 cvAU, dlAU2, iAU;
  gamma_corr2(make_float2(A, U), cvAU, dlAU2, iAU);
  const float iA = iAU.x;
  const float iN = rcp_ftz(N);
  const float lUA = log_ftz(U * iA);  // enters V with weight (B-1)/2 only
  const float dlAU = dlAU2.x - dlAU2.y;
  float V = 0.0f, sumL2 = 0.0f, csum = 0.0f;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    if (b < nb) {
      const float u = x[b] + a[b];
      float2 cv, dl, iz;
      gamma_corr2(make_float2(a[b], u), cv, dl, iz);
      // num = x A - a N without cancellation error: p + e == a N exactly
      const float p = a[b] * N;
      const float e = fmaf(a[b], N, -p);
      const float num = fmaf(x[b], A, -p) - e;
      const float t = num * iz.y;
      const float2 y = make_float2(-t * iA, t * iN);
      float2 L = log1p_ratio_series2(y);  // (-L2, -L1) = (ln(a U / (u A)), ln(x U / (u N)))
      if (!log1p_ratio_in_range2(y)) {  // outlier bin: rare, one branch for the pair
        const float Uu = U * iz.y;
        if (!log1p_ratio_in_range(y.x)) L.x = log_ftz(a[b] * iA * Uu);
        if (!log1p_ratio_in_range(y.y)) L.y = log_ftz(x[b] * iN * Uu);  // x = 0: -inf, discarded below
      }
      const float L2 = -L.x;
      V += x[b] > 0.0f ? fmaf(-x[b], L.y, a[b] * L2) : a[b] * L2;
      sumL2 += L2;
      csum += cv.y - cv.x;
      gb[b] = L2 + (dlAU + (dl.y - dl.x));
    } else {
      gb[b] = 0.0f;
    }
  }
  V += -0.5f * (sumL2 + float(nb - 1) * lUA) + (csum - (cvAU.y - cvAU.x));
  return V;
}

template <typename real, int NB>
__device__ __forceinline__ real dm_row_kl(int nb, const real (&x)[NB], const real (&a)[NB], real N, real A,
                                          real (&gb)[NB]) {
#ifndef BEAN_NO_PACKED_FP32
  if constexpr (sizeof(real) == 4) return dm_row_kl_packed<NB>(nb, x, a, N, A, gb);
#endif
  return dm_row_kl_scalar<real, NB>(nb, x, a, N, A, gb);
}

}  // namespace bean
