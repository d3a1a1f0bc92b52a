// bean_row.cuh -- one Dirichlet-Multinomial row: value and digamma differences in "KL form".
//
// Replaces, per (replicate, guide, layer) row, pyro's
//   DirichletMultinomial(a).log_prob(x) = lgG(A) + lgG(N+1) - lgG(N+A) - sum_b [lgG(x_b+1) + lgG(a_b) - lgG(x_b+a_b)]
// and the digamma differences autograd would produce, written so that no large
// term is ever formed.  With u_b = x_b + a_b, U = N + A, num_b = x_b A - a_b N (computed exactly via FMA):
//
//   L2_b = log1p( num_b / (U a_b)) = ln(u_b A / (a_b U))          L1_b = log1p(-num_b / (U x_b)) = ln(u_b N / (U x_b))
//   V    = sum_b [x_b L1_b + a_b L2_b] - 1/2 [sum_b L2_b + (B-1) ln(U/A)] + sum_b [cv(u_b) - cv(a_b)] - [cv(U) - cv(A)]
//   gb   = psi(A) - psi(U) + psi(u_b) - psi(a_b) = L2_b + [dl(A) - dl(U)] + [dl(u_b) - dl(a_b)]
//   log_prob = V + K,   K = lgG(N+1) - sum_b lgG(x_b+1) + sum_{x_b>0} x_b ln(x_b / N)      (data only, hoisted)
//
// (cv, dl) = gamma_corr: the parts of lgamma / digamma beyond the Stirling main term (bean_math.cuh).
#pragma once
#include "bean_math.cuh"

namespace bean {

template <typename real> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<double> { typedef double4 type; };

// Everything one bin contributes, as ONE out-of-line function: the SVI kernel calls it B times per row
// instead of inlining B copies (the inlined kernel was 230 KB of SASS and stalled on instruction fetch).
// Returns {x L1 + a L2,  L2,  cv(u) - cv(a),  dl(u) - dl(a)}.
template <typename real>
__device__ __noinline__ typename Vec4<real>::type dm_bin_terms(real x, real a, real N, real A, real rU) {
  const real u = x + a;
  real cva, dla, cvu, dlu;
  gamma_corr(a, cva, dla);
  gamma_corr(u, cvu, dlu);
  // num = x A - a N without cancellation error: p + e == a N exactly
  const real p = a * N;
  const real e = fma(a, N, -p);
  const real num = fma(x, A, -p) - e;
  const real t = num * rU;
  // log1p inlined here (its only hot call site): two fewer calls per bin than through the shared out-of-line copy
  const real L2 = Num<real>::log1p_inl(Num<real>::div(t, a));
  const real L1 = x > real(0) ? Num<real>::log1p_inl(Num<real>::div(-t, x)) : real(0);
  typename Vec4<real>::type out;
  out.x = x * L1 + a * L2;
  out.y = L2;
  out.z = cvu - cva;
  out.w = dlu - dla;
  return out;
}

// x, a: counts and concentrations of the row (a_b > 0); nb = number of bins actually used (<= NB).
// Returns V; gb[b] receives the digamma difference of bin b.
template <typename real, int NB>
__device__ __forceinline__ real dm_row_kl(int nb, const real (&x)[NB], const real (&a)[NB], real N, real A,
                                          real (&gb)[NB]) {
  const real U = N + A;
  const real rU = Num<real>::rcp(U);
  real cvA, dlA, cvU, dlU;
  gamma_corr(A, cvA, dlA);
  gamma_corr(U, cvU, dlU);
  const real lUA = Num<real>::flog(Num<real>::div(U, A));  // enters V with weight (B-1)/2 only
  const real dlAU = dlA - dlU;
  real V = real(0), sumL2 = real(0), csum = real(0);
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    if (b < nb) {
      const typename Vec4<real>::type t = dm_bin_terms<real>(x[b], a[b], N, A, rU);
      V += t.x;
      sumL2 += t.y;
      csum += t.z;
      gb[b] = t.y + (dlAU + t.w);
    } else {
      gb[b] = real(0);
    }
  }
  V += real(-0.5) * (sumL2 + real(nb - 1) * lUA) + (csum - (cvU - cvA));
  return V;
}

}  // namespace bean
