// bean_pi_sites.cu -- the editing-rate sites of the MultiMixtureNormal / survival MixtureNormal programs as ONE
// forward + local-gradient kernel (general allele count A, R replicates, C control conditions):
//
//   V =   sum_{r,g} m[r,g] log Dirichlet(pi[r,g,:]; conc_model[g,:])                       (model `pi` site)
//       + sum_{r,c,g} m[r,g] log Multinomial(counts[r,c,g,:]; q / sum_a q),  q_a = pi_a exp(growth[g,a] t_c)
//       - sum_{r,g} m'[r,g] log Dirichlet(pi[r,g,:]; conc_guide[g,:])                      (guide `pi` site)
//
// reference: model.py:632-670 + guide :938-950 (tiling), survival_model.py:313-346 / :496-548 and guides :699-712 / :790-833;
// torch.distributions.Dirichlet.log_prob / Multinomial.log_prob (probs normalised, clamped to [eps, 1 - eps], logged).
// In the torch-op version these sites and their autograd backward were ~150 of the ~290 kernel launches of a step.
// The data-only part of the Multinomial (lgamma(N + 1) - sum lgamma(x + 1)) is a constant the caller adds once.
//
// One warp (tiling) or one thread (A < 8) owns one guide; a warp's lanes stride over the alleles (coalesced rows).  Three sweeps over the
// alleles: (1) concentration sums and the Multinomial normalisers S[r,c]; (2) the Multinomial value and
// hbar[r,c] = sum_a h_a n_a; (3) the Dirichlet values and every gradient.
#include "bean_common.cuh"
#include "bean_math.cuh"

namespace bean {

constexpr int PI_WARPS = 4;
constexpr int PI_WIDE_MIN_ALLELES = 8;  // from this many alleles per guide on, a warp owns a guide

template <typename real>
struct PiSitesParams {
  int G, R, A, C, mask_guide_site;
  const real* conc_g;
  const real* conc_m;
  const real* pi;
  const real* counts;
  const uint8_t* mask;
  const real* growth;
  real tc[BEAN_MAX_BINS];
  real lo, hi;
  double* partial;
  real* d_conc_g;
  real* d_conc_m;
  real* d_pi;
  real* d_growth;
};

template <int LPG, typename real>
__device__ __forceinline__ real warp_all_sum(real v) {
  if (LPG == 1) return v;
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// LPG lanes per guide: 32 (tiling: tens to hundreds of alleles) or 1 (variant designs, A = 2: a warp per guide left 30 of
// 32 lanes idle -- 5.2 ms of a 7.5 ms survival step at 1M guides)
template <typename real, int LPG>
__global__ void __launch_bounds__(PI_WARPS * 32) pi_sites_kernel(const PiSitesParams<real> p) {
  const int lane = LPG == 1 ? 0 : (threadIdx.x & 31);
  const int g = LPG == 1 ? blockIdx.x * (PI_WARPS * 32) + threadIdx.x : blockIdx.x * PI_WARPS + (threadIdx.x >> 5);
  if (g >= p.G) return;  // LPG == 32: warp-uniform; LPG == 1: no warp collectives below
  const int R = p.R, A = p.A, C = p.C, G = p.G;
  const size_t ga = (size_t)g * A;
  // ---- sweep 1: sum of concentrations; S[r][c] = sum_a pi[r,g,a] w[c,a] ---------------------------------------
  real S[BEAN_MAX_RB], H[BEAN_MAX_RB];
  for (int i = 0; i < R * C; ++i) S[i] = H[i] = real(0);
  real sum_g = real(0), sum_m = real(0);
  for (int a = lane; a < A; a += LPG) {
    sum_g += p.conc_g[ga + a];
    sum_m += p.conc_m[ga + a];
    const real mu = p.growth ? p.growth[ga + a] : real(0);
    for (int c = 0; c < C; ++c) {
      const real w = p.growth ? Num<real>::exp(mu * p.tc[c]) : real(1);
      for (int r = 0; r < R; ++r) S[r * C + c] += p.pi[((size_t)r * G + g) * A + a] * w;
    }
  }
  sum_g = warp_all_sum<LPG>(sum_g);
  sum_m = warp_all_sum<LPG>(sum_m);
  for (int i = 0; i < R * C; ++i) S[i] = warp_all_sum<LPG>(S[i]);
  int n_in = 0;  // replicates of this guide inside the repguide mask
  for (int r = 0; r < R; ++r) n_in += p.mask[(size_t)r * G + g] != 0;
  const int n_guide = p.mask_guide_site ? n_in : R;
  // ---- sweep 2: Multinomial value, hbar[r][c] ---------------------------------------------------------------
  double val = 0.0;
  for (int a = lane; a < A; a += LPG) {
    const real mu = p.growth ? p.growth[ga + a] : real(0);
    for (int c = 0; c < C; ++c) {
      const real w = p.growth ? Num<real>::exp(mu * p.tc[c]) : real(1);
      for (int r = 0; r < R; ++r) {
        if (!p.mask[(size_t)r * G + g]) continue;
        const real x = p.counts[(((size_t)r * C + c) * G + g) * A + a];
        const real n = p.pi[((size_t)r * G + g) * A + a] * w / S[r * C + c];
        const bool inside = n >= p.lo && n <= p.hi;
        const real cl = n < p.lo ? p.lo : (n > p.hi ? p.hi : n);
        if (x != real(0)) val += (double)(x * Num<real>::log(cl));
        if (inside) H[r * C + c] += x;  // h_a n_a = x_a where the clamp passes the gradient
      }
    }
  }
  for (int i = 0; i < R * C; ++i) H[i] = warp_all_sum<LPG>(H[i]);
  // ---- sweep 3: Dirichlet values, all gradients -----------------------------------------------------------------
  real lg_sg, dg_sg, lg_sm, dg_sm;
  lgamma_digamma(sum_g, lg_sg, dg_sg);
  lgamma_digamma(sum_m, lg_sm, dg_sm);
  if (lane == 0) val += (double)n_in * (double)lg_sm - (double)n_guide * (double)lg_sg;
  for (int a = lane; a < A; a += LPG) {
    const real cg = p.conc_g[ga + a], cm = p.conc_m[ga + a];
    real lg_g, dg_g, lg_m, dg_m;
    lgamma_digamma(cg, lg_g, dg_g);
    lgamma_digamma(cm, lg_m, dg_m);
    val += -(double)n_in * (double)lg_m + (double)n_guide * (double)lg_g;
    real dcg = -real(n_guide) * (dg_sg - dg_g), dcm = real(n_in) * (dg_sm - dg_m);
    const real mu = p.growth ? p.growth[ga + a] : real(0);
    real dmu = real(0);
    for (int r = 0; r < R; ++r) {
      const bool in = p.mask[(size_t)r * G + g] != 0;
      const bool in_g = p.mask_guide_site ? in : true;
      const size_t ia = ((size_t)r * G + g) * A + a;
      const real pi = p.pi[ia];
      const real lp = Num<real>::log(pi), ip = real(1) / pi;
      real dpi = real(0);
      if (in) {
        if (cm != real(1)) val += (double)((cm - real(1)) * lp);  // xlogy(conc - 1, pi)
        dcm += lp;
        dpi += (cm - real(1)) * ip;
        for (int c = 0; c < C; ++c) {
          const real w = p.growth ? Num<real>::exp(mu * p.tc[c]) : real(1);
          const real s = S[r * C + c];
          const real n = pi * w / s;
          const real x = p.counts[(((size_t)r * C + c) * G + g) * A + a];
          const real h = (n >= p.lo && n <= p.hi) ? x / n : real(0);
          const real dq = (h - H[r * C + c]) / s;  // d / d q_a, q_a = pi_a w_a
          dpi += w * dq;
          dmu += pi * w * p.tc[c] * dq;
        }
      }
      if (in_g) {
        if (cg != real(1)) val -= (double)((cg - real(1)) * lp);
        dcg -= lp;
        dpi -= (cg - real(1)) * ip;
      }
      p.d_pi[ia] = dpi;
    }
    p.d_conc_g[ga + a] = dcg;
    p.d_conc_m[ga + a] = dcm;
    if (p.d_growth) p.d_growth[ga + a] = dmu;
  }
  val = warp_all_sum<LPG>(val);
  if (lane == 0) p.partial[g] = val;
}

template <typename real>
static int launch_pi_sites(const BeanPiSitesArgs* a, void* stream) {
  BEAN_REQUIRE(a != nullptr, BEAN_EINVAL, "args is NULL");
  BEAN_REQUIRE(a->n_guides > 0 && a->n_reps > 0 && a->n_alleles >= 2 && a->n_controls > 0, BEAN_EINVAL,
               "sizes must be positive (A >= 2)");
  BEAN_REQUIRE(a->n_alleles <= BEAN_MAX_ALLELES, BEAN_EINVAL, "n_alleles %d > %d", a->n_alleles, BEAN_MAX_ALLELES);
  BEAN_REQUIRE(a->n_controls <= BEAN_MAX_BINS && a->n_reps * a->n_controls <= BEAN_MAX_RB, BEAN_EINVAL,
               "n_controls %d / n_reps * n_controls %d out of range", a->n_controls, a->n_reps * a->n_controls);
  BEAN_REQUIRE(a->conc_guide && a->conc_model && a->pi && a->counts && a->rep_guide_mask, BEAN_EINVAL,
               "conc_guide / conc_model / pi / counts / rep_guide_mask must be non-NULL");
  BEAN_REQUIRE(a->partial && a->d_conc_guide && a->d_conc_model && a->d_pi, BEAN_EINVAL, "outputs must be non-NULL");
  BEAN_REQUIRE((a->growth == nullptr) == (a->d_growth == nullptr), BEAN_EINVAL, "growth and d_growth go together");
  BEAN_REQUIRE(a->growth == nullptr || a->control_time != nullptr, BEAN_EINVAL, "growth needs control_time");
  BEAN_REQUIRE(a->prob_eps > 0 && a->prob_eps < 0.5, BEAN_EINVAL, "prob_eps out of range");
  PiSitesParams<real> p{};
  p.G = a->n_guides; p.R = a->n_reps; p.A = a->n_alleles; p.C = a->n_controls; p.mask_guide_site = a->mask_guide_site;
  p.conc_g = static_cast<const real*>(a->conc_guide);
  p.conc_m = static_cast<const real*>(a->conc_model);
  p.pi = static_cast<const real*>(a->pi);
  p.counts = static_cast<const real*>(a->counts);
  p.mask = a->rep_guide_mask;
  p.growth = static_cast<const real*>(a->growth);
  for (int c = 0; c < BEAN_MAX_BINS; ++c) p.tc[c] = (a->control_time && c < a->n_controls) ? real(a->control_time[c]) : real(0);
  p.lo = real(a->prob_eps);
  p.hi = real(1) - real(a->prob_eps);
  p.partial = a->partial;
  p.d_conc_g = static_cast<real*>(a->d_conc_guide);
  p.d_conc_m = static_cast<real*>(a->d_conc_model);
  p.d_pi = static_cast<real*>(a->d_pi);
  p.d_growth = static_cast<real*>(a->d_growth);
  if (a->n_alleles < PI_WIDE_MIN_ALLELES) {
    const int grid = (a->n_guides + PI_WARPS * 32 - 1) / (PI_WARPS * 32);
    pi_sites_kernel<real, 1><<<grid, PI_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  } else {
    const int grid = (a->n_guides + PI_WARPS - 1) / PI_WARPS;
    pi_sites_kernel<real, 32><<<grid, PI_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  }
  BEAN_CUDA(cudaPeekAtLastError());
  return BEAN_OK;
}

}  // namespace bean

extern "C" {
int bean_pi_sites_f32(const BeanPiSitesArgs* a, void* stream) { return bean::launch_pi_sites<float>(a, stream); }
int bean_pi_sites_f64(const BeanPiSitesArgs* a, void* stream) { return bean::launch_pi_sites<double>(a, stream); }
}
