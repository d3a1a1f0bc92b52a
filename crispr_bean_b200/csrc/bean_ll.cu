// bean_ll.cu -- count log-likelihood forward + local gradients (generic in R, B, A, layers, mode).
//
// One thread (few alleles) or one warp (tiling: many alleles) owns one guide: it builds
// e[r][b] = sum_a pi[r][a] P[b][a] in per-thread scratch, scores
// every (replicate, layer) row with the Dirichlet-Multinomial and turns the digamma differences into
// d ll / d e in the same pass, then contracts d ll / d e back onto (mu, sd, pi) per allele.  Replaces
// bean/model/utils.py:10-76 + bean/model/model.py:495-547 (survival_model.py:352-424) and their
// autograd backward; closed-form gradients per SURVEY App. A.4.
#include <stdarg.h>
#include <string.h>

#include "bean_common.cuh"
#include "bean_math.cuh"
#include "bean_row.cuh"

namespace bean {

static thread_local char g_err[512] = "";
char* err_slot() { return g_err; }
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int validate_screen(const BeanScreen* s) {
  BEAN_REQUIRE(s != nullptr, BEAN_EINVAL, "screen is NULL");
  BEAN_REQUIRE(s->n_guides > 0, BEAN_EINVAL, "n_guides must be > 0 (got %d)", s->n_guides);
  BEAN_REQUIRE(s->n_reps > 0 && s->n_bins > 0, BEAN_EINVAL, "n_reps / n_bins must be > 0");
  BEAN_REQUIRE(s->n_bins <= BEAN_MAX_BINS, BEAN_EINVAL, "n_bins %d > BEAN_MAX_BINS %d", s->n_bins, BEAN_MAX_BINS);
  BEAN_REQUIRE(s->n_reps * s->n_bins <= BEAN_MAX_RB, BEAN_EINVAL, "n_reps*n_bins %d > BEAN_MAX_RB %d",
               s->n_reps * s->n_bins, BEAN_MAX_RB);
  BEAN_REQUIRE(s->n_layers >= 1 && s->n_layers <= BEAN_MAX_LAYERS, BEAN_EINVAL, "n_layers must be 1 or 2");
  BEAN_REQUIRE(s->mode == BEAN_MODE_SORTING || s->mode == BEAN_MODE_SURVIVAL, BEAN_EINVAL, "bad mode %d", s->mode);
  BEAN_REQUIRE(s->x && s->a0 && s->row_mask, BEAN_EINVAL, "x / a0 / row_mask must be non-NULL");
  BEAN_REQUIRE(s->size_factor && s->sample_mask, BEAN_EINVAL, "size_factor / sample_mask must be non-NULL");
  if (s->mode == BEAN_MODE_SORTING)
    BEAN_REQUIRE(s->upper_thres && s->lower_thres, BEAN_EINVAL, "sorting mode needs upper_thres / lower_thres");
  else
    BEAN_REQUIRE(s->timepoints, BEAN_EINVAL, "survival mode needs timepoints");
  BEAN_REQUIRE(aligned16(s->x), BEAN_EALIGN, "x is not 16-byte aligned");
  return BEAN_OK;
}

constexpr int LL_THREADS = 128;
constexpr int LL_WIDE_WARPS = 4;          // guides per CTA of the warp-per-guide kernel
constexpr int LL_WIDE_MIN_ALLELES = 8;    // from this many alleles per guide on, a warp owns a guide

template <typename real>
struct LLParams {
  int G, R, B, L, A, mode;
  real mask_thres;
  const real* x;
  const real* a0;
  const uint8_t* row_mask;
  const double* row_const;
  const real* mu;
  const real* sd;
  const real* pi;
  const uint8_t* allele_mask;
  real* ll_row;
  double* ll_partial;
  real* d_mu;
  real* d_sd;
  real* d_pi;
  SampleTables<real> t;
};

template <typename real>
__device__ __forceinline__ void allele_bin_probs(const LLParams<real>& p, real mu, real sd, bool exists, real* P,
                                                 real* dPm, real* dPs) {
  for (int b = 0; b < p.B; ++b) {
    if (!exists) {
      P[b] = dPm[b] = dPs[b] = real(0);  // model/utils.py:73-74: res[~mask] = 0
    } else if (p.mode == BEAN_MODE_SORTING) {
      bin_prob_sorting(p.t.thr_u[b], p.t.thr_l[b], mu, sd, P[b], dPm[b], dPs[b]);
    } else {
      P[b] = Num<real>::exp(mu * p.t.tp[b]);  // survival_model.py:358-361
      dPm[b] = p.t.tp[b] * P[b];
      dPs[b] = real(0);
    }
  }
}

// Dirichlet-Multinomial over every (replicate, layer) row of guide g.  In: e[r][b] expected bin fractions.
// Out: e[r][b] = d ll / d e[r][b]; returns the guide's masked log-likelihood.  `writer` stores ll_row.
template <typename real>
__device__ __forceinline__ double ll_rows(const LLParams<real>& p, int g, real* e, bool writer) {
  const int R = p.R, B = p.B;
  const real eps = real(1e-5);
  double ll_acc = 0.0;
  for (int r = 0; r < R; ++r) {
    const bool rmask = p.row_mask[(size_t)r * p.G + g] != 0;
    real de[BEAN_MAX_BINS];
    for (int b = 0; b < B; ++b) de[b] = real(0);
    for (int l = 0; l < p.L; ++l) {
      const real* xr = p.x + (((size_t)l * R + r) * p.G + g) * B;
      const real a0 = p.a0[(size_t)l * p.G + g];
      real xb[BEAN_MAX_BINS] = {}, pb[BEAN_MAX_BINS];
      real N = real(0), S = real(0);
      for (int b = 0; b < B; ++b) {
        xb[b] = xr[b];
        N += xb[b];
        pb[b] = e[r * B + b] * p.t.sf[l][r * B + b];
        S += pb[b];
      }
      const bool w = rmask && (N > p.mask_thres);
      const real inv = real(1) / (S + eps);
      real ab[BEAN_MAX_BINS] = {}, frac[BEAN_MAX_BINS];
      bool live[BEAN_MAX_BINS];
      real Asum = real(0);
      for (int b = 0; b < B; ++b) {
        frac[b] = (pb[b] + eps / real(B)) * inv;
        const real raw = frac[b] * a0 * p.t.smask[r * B + b];
        live[b] = raw >= eps;  // clamp(min=eps) passes gradient where input >= eps
        ab[b] = live[b] ? raw : eps;
        Asum += ab[b];
      }
      if (!w) {  // poutine.mask: the row contributes nothing
        if (writer && p.ll_row) p.ll_row[((size_t)l * R + r) * p.G + g] = real(0);
        continue;
      }
      real psi_diff[BEAN_MAX_BINS];
      const real V = dm_row_kl<real, BEAN_MAX_BINS>(B, xb, ab, N, Asum, psi_diff);
      // data-only part of the log-pmf, hoisted to tensorisation time (row_const is NULL -> 0)
      const double K = p.row_const ? p.row_const[((size_t)l * R + r) * p.G + g] : 0.0;
      const double ll = (double)V + K;
      real gb[BEAN_MAX_BINS];
      real dot = real(0);
      for (int b = 0; b < B; ++b) {
        gb[b] = live[b] ? psi_diff[b] * p.t.smask[r * B + b] : real(0);
        dot += gb[b] * frac[b];
      }
      if (writer && p.ll_row) p.ll_row[((size_t)l * R + r) * p.G + g] = real(ll);
      ll_acc += ll;
      const real c = a0 * inv;
      for (int b = 0; b < B; ++b) de[b] += p.t.sf[l][r * B + b] * c * (gb[b] - dot);
    }
    for (int b = 0; b < B; ++b) e[r * B + b] = de[b];
  }
  return ll_acc;
}

// One guide, alleles a0, a0 + stride, ...: pass 1 accumulates e[r][b] += pi[r][a] P[b][a].
template <typename real>
__device__ __forceinline__ void ll_mix_alleles(const LLParams<real>& p, int g, int a0, int stride, real* e) {
  const int R = p.R, B = p.B, A = p.A;
  real P[BEAN_MAX_BINS], dPm[BEAN_MAX_BINS], dPs[BEAN_MAX_BINS];
  for (int a = a0; a < A; a += stride) {
    const bool exists = p.allele_mask == nullptr || p.allele_mask[(size_t)g * A + a] != 0;
    allele_bin_probs(p, p.mu[(size_t)g * A + a], p.sd[(size_t)g * A + a], exists, P, dPm, dPs);
    for (int r = 0; r < R; ++r) {
      const real w = p.pi ? p.pi[((size_t)g * R + r) * A + a] : real(1);
      for (int b = 0; b < B; ++b) e[r * B + b] += w * P[b];
    }
  }
}

// Pass 3: contract d ll / d e (in e[]) onto (mu, sd, pi) of alleles a0, a0 + stride, ...
template <typename real>
__device__ __forceinline__ void ll_contract_alleles(const LLParams<real>& p, int g, int a0, int stride, const real* e) {
  const int R = p.R, B = p.B, A = p.A;
  real P[BEAN_MAX_BINS], dPm[BEAN_MAX_BINS], dPs[BEAN_MAX_BINS];
  for (int a = a0; a < A; a += stride) {
    const bool exists = p.allele_mask == nullptr || p.allele_mask[(size_t)g * A + a] != 0;
    allele_bin_probs(p, p.mu[(size_t)g * A + a], p.sd[(size_t)g * A + a], exists, P, dPm, dPs);
    real dmu = real(0), dsd = real(0);
    for (int r = 0; r < R; ++r) {
      const real w = p.pi ? p.pi[((size_t)g * R + r) * A + a] : real(1);
      real dpi = real(0);
      for (int b = 0; b < B; ++b) {
        const real d = e[r * B + b];
        dpi += d * P[b];
        dmu += d * w * dPm[b];
        dsd += d * w * dPs[b];
      }
      if (p.d_pi) p.d_pi[((size_t)g * R + r) * A + a] = dpi;
    }
    p.d_mu[(size_t)g * A + a] = dmu;
    p.d_sd[(size_t)g * A + a] = dsd;
  }
}

// Few alleles (variant designs, A <= LL_WIDE_MIN_ALLELES): one thread owns one guide.
template <typename real>
__global__ void __launch_bounds__(LL_THREADS) ll_generic_kernel(const LLParams<real> p) {
  __shared__ double red[32];
  const int g = blockIdx.x * LL_THREADS + threadIdx.x;
  double ll_acc = 0.0;
  if (g < p.G) {
    real e[BEAN_MAX_RB];
    for (int i = 0; i < p.R * p.B; ++i) e[i] = real(0);
    ll_mix_alleles(p, g, 0, 1, e);
    ll_acc = ll_rows(p, g, e, true);
    ll_contract_alleles(p, g, 0, 1, e);
  }
  const double tot = block_sum(ll_acc, red);
  if (threadIdx.x == 0) p.ll_partial[blockIdx.x] = tot;
}

// Many alleles per guide (tiling designs: tens to hundreds): one WARP owns one guide, lanes stride over the
// alleles so that mu/sd/pi/d_pi rows are read and written coalesced; e[r][b] is completed by a butterfly
// all-reduce (every lane ends with the same bits), the R x L rows are then scored redundantly by all lanes.
template <typename real>
__global__ void __launch_bounds__(LL_WIDE_WARPS * 32) ll_wide_kernel(const LLParams<real> p) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * LL_WIDE_WARPS + (threadIdx.x >> 5);
  double ll_acc = 0.0;
  if (g < p.G) {  // warp-uniform
    real e[BEAN_MAX_RB];
    const int RB = p.R * p.B;
    for (int i = 0; i < RB; ++i) e[i] = real(0);
    ll_mix_alleles(p, g, lane, 32, e);
    for (int i = 0; i < RB; ++i) {
      real v = e[i];
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
      e[i] = v;
    }
    const double ll = ll_rows(p, g, e, lane == 0);
    if (lane == 0) ll_acc = ll;
    ll_contract_alleles(p, g, lane, 32, e);
  }
  const double tot = block_sum(ll_acc, red);
  if (threadIdx.x == 0) p.ll_partial[blockIdx.x] = tot;
}

template <typename real>
static int launch_ll(const BeanScreen* s, const BeanLLArgs* a, void* stream) {
  int rc = validate_screen(s);
  if (rc != BEAN_OK) return rc;
  BEAN_REQUIRE(a != nullptr, BEAN_EINVAL, "args is NULL");
  BEAN_REQUIRE(a->n_alleles >= 1 && a->n_alleles <= BEAN_MAX_ALLELES, BEAN_EINVAL, "n_alleles %d out of range [1, %d]",
               a->n_alleles, BEAN_MAX_ALLELES);
  BEAN_REQUIRE(a->mu_allele && a->sd_allele, BEAN_EINVAL, "mu_allele / sd_allele must be non-NULL");
  BEAN_REQUIRE(a->pi != nullptr || a->n_alleles == 1, BEAN_EINVAL, "pi may only be NULL when n_alleles == 1");
  BEAN_REQUIRE(a->ll_partial && a->d_mu && a->d_sd, BEAN_EINVAL, "ll_partial / d_mu / d_sd must be non-NULL");
  BEAN_REQUIRE(a->pi == nullptr || a->d_pi != nullptr, BEAN_EINVAL, "d_pi must be non-NULL when pi is given");
  LLParams<real> p{};
  p.G = s->n_guides; p.R = s->n_reps; p.B = s->n_bins; p.L = s->n_layers; p.A = a->n_alleles; p.mode = s->mode;
  p.mask_thres = real(s->mask_thres);
  p.x = static_cast<const real*>(s->x);
  p.a0 = static_cast<const real*>(s->a0);
  p.row_mask = s->row_mask;
  p.row_const = s->row_const;
  p.mu = static_cast<const real*>(a->mu_allele);
  p.sd = static_cast<const real*>(a->sd_allele);
  p.pi = static_cast<const real*>(a->pi);
  p.allele_mask = a->allele_mask;
  p.ll_row = static_cast<real*>(a->ll_row);
  p.ll_partial = a->ll_partial;
  p.d_mu = static_cast<real*>(a->d_mu);
  p.d_sd = static_cast<real*>(a->d_sd);
  p.d_pi = static_cast<real*>(a->d_pi);
  fill_tables(s, p.t);
  const int grid = bean_ll_num_partials(s->n_guides, a->n_alleles);
  if (a->n_alleles >= LL_WIDE_MIN_ALLELES)
    ll_wide_kernel<real><<<grid, LL_WIDE_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else
    ll_generic_kernel<real><<<grid, LL_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  BEAN_CUDA(cudaPeekAtLastError());
  return BEAN_OK;
}

}  // namespace bean

extern "C" {

int bean_abi_version(void) { return BEAN_ABI_VERSION; }
const char* bean_last_error(void) { return bean::err_slot(); }

int bean_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return bean::fail(BEAN_ECUDA, "cudaGetDevice failed (no CUDA device?)");
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return bean::fail(BEAN_ECUDA, "cudaDeviceGetAttribute failed");
  return n;
}

int bean_ll_num_partials(int32_t n_guides, int32_t n_alleles) {
  const int per_cta = n_alleles >= bean::LL_WIDE_MIN_ALLELES ? bean::LL_WIDE_WARPS : bean::LL_THREADS;
  return (n_guides + per_cta - 1) / per_cta;
}
int bean_ll_f32(const BeanScreen* s, const BeanLLArgs* a, void* stream) { return bean::launch_ll<float>(s, a, stream); }
int bean_ll_f64(const BeanScreen* s, const BeanLLArgs* a, void* stream) { return bean::launch_ll<double>(s, a, stream); }

}  // extern "C"
