// bean_latent_sites.cu -- the per-variant / per-edit latent sites (`mu_targets`, `sd_targets`) of the programs that run
// on torch autograd around the likelihood kernel: reparameterised draws, prior and guide densities and their closed-form
// gradients in ONE launch (+ one for the backward combination).
//   guide  (model.py:893-921, survival_model.py:640-652): mu ~ Normal(mu_loc, mu_scale), sd ~ LogNormal(sd_loc, sd_scale)
//   model  (model.py:579-610, survival_model.py:246-274): mu ~ Laplace(0, 1) | Normal(prior), sd ~ LogNormal(prior)
//   mu = mu_loc + s e,  s = exp(mu_log_scale);        y = sd_loc + t e', t = exp(sd_log_scale), sd = exp(y)
//   V  = sum_i [log p(mu_i) - log q(mu_i)] + [log p(sd_i) - log q(sd_i)]
//      = sum_i  log p(mu_i) + log s + e^2/2 + log(2 pi)/2   +   -log b - ((y - a)/b)^2/2 + log t + e'^2/2
// (the -y and log(2 pi)/2 of the two LogNormals cancel).  In torch-op form these four sites and their backward were
// ~100 of a step's kernel launches.  The fused sorting step carries the same formulas in svi_variant_kernel (bean_svi.cu).
#include "bean_common.cuh"
#include "bean_math.cuh"

namespace bean {

constexpr int LAT_THREADS = 256;

template <typename real>
struct LatentParams {
  long long n;
  int has_sd, mu_prior_normal;
  const real *mu_loc, *mu_ls, *sd_loc, *sd_ls, *eps_mu, *eps_sd;
  real p_mu_loc, p_mu_scale, p_sd_loc, p_sd_scale;
  const real *p_mu_loc_v, *p_mu_scale_v, *p_sd_loc_v, *p_sd_scale_v;
  real *mu, *sd;
  double* partial;
  real* dv;
};

template <typename real>
__global__ void __launch_bounds__(LAT_THREADS) latent_sites_kernel(const LatentParams<real> p) {
  __shared__ double red[32];
  const long long i = (long long)blockIdx.x * LAT_THREADS + threadIdx.x;
  double v = 0.0;
  if (i < p.n) {
    const real HALF_LOG_2PI = real(0.91893853320467274178);
    const real ls = p.mu_ls[i], s = Num<real>::exp(ls), e = p.eps_mu[i];
    const real mu = p.mu_loc[i] + s * e;
    real lp, dlp;  // log prior density of mu and its derivative
    if (p.mu_prior_normal) {
      const real m = p.p_mu_loc_v ? p.p_mu_loc_v[i] : p.p_mu_loc, sc = p.p_mu_scale_v ? p.p_mu_scale_v[i] : p.p_mu_scale;
      const real z = (mu - m) / sc;
      lp = -Num<real>::log(sc) - real(0.5) * z * z - HALF_LOG_2PI;
      dlp = -z / sc;
    } else {
      lp = -real(0.69314718055994530942) - Num<real>::fabs(mu);  // Laplace(0, 1)
      dlp = mu > real(0) ? real(-1) : (mu < real(0) ? real(1) : real(0));
    }
    p.mu[i] = mu;
    v = (double)lp + (double)(ls + real(0.5) * e * e + HALF_LOG_2PI);
    p.dv[i] = dlp;
    p.dv[p.n + i] = dlp * s * e + real(1);
    if (p.has_sd) {
      const real lt = p.sd_ls[i], t = Num<real>::exp(lt), e2 = p.eps_sd[i];
      const real y = p.sd_loc[i] + t * e2;
      const real a = p.p_sd_loc_v ? p.p_sd_loc_v[i] : p.p_sd_loc, b = p.p_sd_scale_v ? p.p_sd_scale_v[i] : p.p_sd_scale;
      const real z = (y - a) / b;
      p.sd[i] = Num<real>::exp(y);
      v += (double)(-Num<real>::log(b) - real(0.5) * z * z + lt + real(0.5) * e2 * e2);
      const real dy = -z / b;
      p.dv[2 * p.n + i] = dy;
      p.dv[3 * p.n + i] = dy * t * e2 + real(1);
    }
  }
  const double tot = block_sum(v, red);
  if (threadIdx.x == 0) p.partial[blockIdx.x] = tot;
}

template <typename real>
struct LatentGradParams {
  long long n;
  int has_sd;
  const real *mu_ls, *sd_ls, *eps_mu, *eps_sd, *sd, *dv, *g_mu, *g_sd, *g_v;
  real* grad;
};

// d L / d(mu_loc, mu_log_scale, sd_loc, sd_log_scale) from the upstream gradients of (mu, sd, V)
template <typename real>
__global__ void __launch_bounds__(LAT_THREADS) latent_sites_grad_kernel(const LatentGradParams<real> p) {
  const long long i = (long long)blockIdx.x * LAT_THREADS + threadIdx.x;
  if (i >= p.n) return;
  const real gv = p.g_v ? p.g_v[0] : real(0);
  const real gm = p.g_mu ? p.g_mu[i] : real(0);
  p.grad[i] = gm + gv * p.dv[i];
  p.grad[p.n + i] = gm * Num<real>::exp(p.mu_ls[i]) * p.eps_mu[i] + gv * p.dv[p.n + i];
  if (p.has_sd) {
    const real gs = (p.g_sd ? p.g_sd[i] : real(0)) * p.sd[i];  // d sd / d y = sd
    p.grad[2 * p.n + i] = gs + gv * p.dv[2 * p.n + i];
    p.grad[3 * p.n + i] = gs * Num<real>::exp(p.sd_ls[i]) * p.eps_sd[i] + gv * p.dv[3 * p.n + i];
  }
}

template <typename real>
static int launch_latent(const BeanLatentSitesArgs* a, void* stream) {
  BEAN_REQUIRE(a != nullptr && a->n > 0, BEAN_EINVAL, "args is NULL or n <= 0");
  BEAN_REQUIRE(a->mu_loc && a->mu_log_scale && a->eps_mu && a->mu && a->partial && a->dv, BEAN_EINVAL, "mu site buffers must be non-NULL");
  BEAN_REQUIRE(!a->has_sd || (a->sd_loc && a->sd_log_scale && a->eps_sd && a->sd), BEAN_EINVAL, "sd site buffers must be non-NULL");
  BEAN_REQUIRE(!a->mu_prior_normal || a->mu_prior_scale_v || a->mu_prior_scale > 0, BEAN_EINVAL, "mu prior scale must be > 0");
  BEAN_REQUIRE(!a->has_sd || a->sd_prior_scale_v || a->sd_prior_scale > 0, BEAN_EINVAL, "sd prior scale must be > 0");
  LatentParams<real> p{};
  p.n = a->n; p.has_sd = a->has_sd; p.mu_prior_normal = a->mu_prior_normal;
  p.mu_loc = static_cast<const real*>(a->mu_loc); p.mu_ls = static_cast<const real*>(a->mu_log_scale);
  p.sd_loc = static_cast<const real*>(a->sd_loc); p.sd_ls = static_cast<const real*>(a->sd_log_scale);
  p.eps_mu = static_cast<const real*>(a->eps_mu); p.eps_sd = static_cast<const real*>(a->eps_sd);
  p.p_mu_loc = real(a->mu_prior_loc); p.p_mu_scale = real(a->mu_prior_scale);
  p.p_sd_loc = real(a->sd_prior_loc); p.p_sd_scale = real(a->sd_prior_scale);
  p.p_mu_loc_v = static_cast<const real*>(a->mu_prior_loc_v); p.p_mu_scale_v = static_cast<const real*>(a->mu_prior_scale_v);
  p.p_sd_loc_v = static_cast<const real*>(a->sd_prior_loc_v); p.p_sd_scale_v = static_cast<const real*>(a->sd_prior_scale_v);
  p.mu = static_cast<real*>(a->mu); p.sd = static_cast<real*>(a->sd);
  p.partial = a->partial; p.dv = static_cast<real*>(a->dv);
  latent_sites_kernel<real><<<bean_latent_sites_num_partials(a->n), LAT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  BEAN_CUDA(cudaPeekAtLastError());
  return BEAN_OK;
}

template <typename real>
static int launch_latent_grad(const BeanLatentSitesGradArgs* a, void* stream) {
  BEAN_REQUIRE(a != nullptr && a->n > 0, BEAN_EINVAL, "args is NULL or n <= 0");
  BEAN_REQUIRE(a->mu_log_scale && a->eps_mu && a->dv && a->grad, BEAN_EINVAL, "mu site buffers must be non-NULL");
  BEAN_REQUIRE(!a->has_sd || (a->sd_log_scale && a->eps_sd && a->sd), BEAN_EINVAL, "sd site buffers must be non-NULL");
  LatentGradParams<real> p{};
  p.n = a->n; p.has_sd = a->has_sd;
  p.mu_ls = static_cast<const real*>(a->mu_log_scale); p.sd_ls = static_cast<const real*>(a->sd_log_scale);
  p.eps_mu = static_cast<const real*>(a->eps_mu); p.eps_sd = static_cast<const real*>(a->eps_sd);
  p.sd = static_cast<const real*>(a->sd); p.dv = static_cast<const real*>(a->dv);
  p.g_mu = static_cast<const real*>(a->g_mu); p.g_sd = static_cast<const real*>(a->g_sd); p.g_v = static_cast<const real*>(a->g_v);
  p.grad = static_cast<real*>(a->grad);
  latent_sites_grad_kernel<real><<<bean_latent_sites_num_partials(a->n), LAT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  BEAN_CUDA(cudaPeekAtLastError());
  return BEAN_OK;
}

}  // namespace bean

extern "C" {
int bean_latent_sites_num_partials(int64_t n) { return (int)((n + bean::LAT_THREADS - 1) / bean::LAT_THREADS); }
int bean_latent_sites_f32(const BeanLatentSitesArgs* a, void* stream) { return bean::launch_latent<float>(a, stream); }
int bean_latent_sites_f64(const BeanLatentSitesArgs* a, void* stream) { return bean::launch_latent<double>(a, stream); }
int bean_latent_sites_grad_f32(const BeanLatentSitesGradArgs* a, void* stream) { return bean::launch_latent_grad<float>(a, stream); }
int bean_latent_sites_grad_f64(const BeanLatentSitesGradArgs* a, void* stream) { return bean::launch_latent_grad<double>(a, stream); }
}
