// bean_dirichlet.cu -- reparameterised Dirichlet draws of the editing-rate site `pi` and their pathwise derivative, as
// stand-alone C-ABI kernels for the programs that run on torch autograd around bean_ll_* (tiling, survival).
//
// Replaces  dist.Dirichlet(concentration).rsample()  of the guides (bean/model/model.py:942-950 MultiMixtureNormalGuide,
// bean/model/survival_model.py:699-712 / :822-833), i.e. torch._sample_dirichlet forward and torch._dirichlet_grad
// (`_Dirichlet_backward`) backward.  Same generator as the fused sorting step (bean_rng.cuh): Philox4x32-10 keyed by the
// run seed, counter = (GLOBAL guide id, replicate | allele << 8, step, stream + attempt), Marsaglia-Tsang gammas with the
// alpha < 1 boost, normalised and clamped to [tiny, 1 - eps/2] like _sample_dirichlet -- so the draws do not depend on
// launch geometry or on how the guides are sharded over GPUs.  The step index is read from DEVICE memory, so the launch can
// be captured in a CUDA graph and replayed.
//
// Mapping: a group of W = 2 .. 32 lanes owns one guide (W = the power of two >= A, capped at 32; lanes stride over the
// alleles), loops over its replicates, and reduces the row sums with shuffles inside the group.
#include "bean_common.cuh"
#include "bean_math.cuh"
#include "bean_rng.cuh"

namespace bean {

constexpr int DIR_THREADS = 128;
enum : uint32_t { STREAM_DIRICHLET = 64 };  // + 8 * attempt (+ 4 for the boost word); bean_rng.cuh uses 0 .. 47

template <typename real>
struct DirichletParams {
  int G, R, A;
  const real* conc;   // [G][A]
  real* x;            // [R][G][A]
  const real* grad_x; // [R][G][A] (backward)
  real* d_conc;       // [G][A]   (backward)
  uint64_t seed;
  uint32_t guide_offset, site;
  const int64_t* step_dev;
  int64_t step_host;
};

template <int W>
__device__ __forceinline__ double group_sum(double v, unsigned mask) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// Gamma(alpha) for (guide g, replicate r, allele a) at `step`
template <typename real>
__device__ __forceinline__ real gamma_draw(uint64_t seed, uint32_t g, uint32_t r, uint32_t a, uint32_t step, uint32_t site, real alpha) {
  GammaMT<real> mt;
  mt.init(alpha);
  const uint2 key = seed_key(seed);
  const uint32_t cy = r | (a << 8);
  real out = real(0);
  bool ok = false;
  for (uint32_t k = 0; k < 16u && !ok; ++k) {
    const uint4 w = philox4x32_10(make_uint4(g, cy, step, STREAM_DIRICHLET + site * 256u + 8u * k), key);
    float n0, n1;
    box_muller(w.x, w.y, n0, n1);
    ok = mt.attempt(n0, 1.0f - u01(w.z), out);
    if (!ok) ok = mt.attempt(n1, 1.0f - u01(w.w), out);
  }
  if (mt.inv_alpha != real(0)) {  // boost: Gamma(a) = Gamma(a + 1) * U^(1/a)
    const uint4 w = philox4x32_10(make_uint4(g, cy, step, STREAM_DIRICHLET + site * 256u + 4u), key);
    out *= Num<real>::pow(real(1) - real(u01(w.x)), mt.inv_alpha);
  }
  return Num<real>::fmax(out, Lim<real>::tiny());
}

template <typename real, int W>
__global__ void __launch_bounds__(DIR_THREADS) dirichlet_rsample_kernel(const DirichletParams<real> p) {
  const int tid = blockIdx.x * DIR_THREADS + threadIdx.x;
  const int g = tid / W, sub = tid % W;
  const unsigned mask = __ballot_sync(0xffffffffu, g < p.G);
  if (g >= p.G) return;  // whole groups leave together (G groups of W lanes; W divides 32)
  const uint32_t step = (uint32_t)(p.step_dev ? *p.step_dev : p.step_host);
  for (int r = 0; r < p.R; ++r) {
    real* xr = p.x + ((size_t)r * p.G + g) * p.A;
    double sum = 0.0;
    for (int a = sub; a < p.A; a += W) {
      const real gam = gamma_draw<real>(p.seed, (uint32_t)g + p.guide_offset, (uint32_t)r, (uint32_t)a, step, p.site, p.conc[(size_t)g * p.A + a]);
      xr[a] = gam;
      sum += (double)gam;
    }
    sum = group_sum<W>(sum, mask);
    const real inv = real(1.0 / sum);
    for (int a = sub; a < p.A; a += W)
      xr[a] = Num<real>::fmin(Num<real>::fmax(xr[a] * inv, Lim<real>::tiny()), Lim<real>::one_minus());
  }
}

// d L / d conc[g][a] = sum_r D(x_ra; c_a, C) (gout_ra - sum_b x_rb gout_rb),  D = torch._dirichlet_grad  (torch:
// _Dirichlet_backward).  Evaluated in double also on the float path, like torch's CPU kernel (accscalar_t = double).
template <typename real, int W>
__global__ void __launch_bounds__(DIR_THREADS) dirichlet_rsample_grad_kernel(const DirichletParams<real> p) {
  const int tid = blockIdx.x * DIR_THREADS + threadIdx.x;
  const int g = tid / W, sub = tid % W;
  const unsigned mask = __ballot_sync(0xffffffffu, g < p.G);
  if (g >= p.G) return;
  double total = 0.0;
  for (int a = sub; a < p.A; a += W) total += (double)p.conc[(size_t)g * p.A + a];
  total = group_sum<W>(total, mask);
  for (int a0 = 0; a0 < p.A; a0 += W) {
    const int a = a0 + sub;
    if (a < p.A) p.d_conc[(size_t)g * p.A + a] = real(0);
  }
  for (int r = 0; r < p.R; ++r) {
    const size_t row = ((size_t)r * p.G + g) * p.A;
    double dot = 0.0;
    for (int a = sub; a < p.A; a += W) dot += (double)p.x[row + a] * (double)p.grad_x[row + a];
    dot = group_sum<W>(dot, mask);
    for (int a = sub; a < p.A; a += W) {
      const double c = (double)p.conc[(size_t)g * p.A + a];
      const double D = dirichlet_grad_one_f64((double)p.x[row + a], c, total);  // double also on the float path (not a hot kernel)
      p.d_conc[(size_t)g * p.A + a] += real(D * ((double)p.grad_x[row + a] - dot));
    }
  }
}

template <typename real, bool GRAD>
static int launch_dirichlet(const BeanDirichletArgs* a, void* stream) {
  BEAN_REQUIRE(a != nullptr, BEAN_EINVAL, "args is NULL");
  BEAN_REQUIRE(a->n_guides > 0 && a->n_reps > 0 && a->n_reps < 256, BEAN_EINVAL, "n_guides must be > 0, n_reps in [1, 255]");
  BEAN_REQUIRE(a->n_alleles >= 2 && a->n_alleles <= BEAN_MAX_ALLELES, BEAN_EINVAL, "n_alleles %d out of range [2, %d]", a->n_alleles,
               BEAN_MAX_ALLELES);
  BEAN_REQUIRE(a->conc && a->x, BEAN_EINVAL, "conc / x must be non-NULL");
  if (GRAD) BEAN_REQUIRE(a->grad_x && a->d_conc, BEAN_EINVAL, "grad_x / d_conc must be non-NULL");
  DirichletParams<real> p{};
  p.G = a->n_guides; p.R = a->n_reps; p.A = a->n_alleles;
  p.conc = static_cast<const real*>(a->conc);
  p.x = static_cast<real*>(a->x);
  p.grad_x = static_cast<const real*>(a->grad_x);
  p.d_conc = static_cast<real*>(a->d_conc);
  p.seed = a->seed; p.guide_offset = a->guide_offset; p.site = a->site;
  p.step_dev = a->step; p.step_host = a->step_value;
  int W = 2;
  while (W < p.A && W < 32) W *= 2;
  const long long threads = (long long)p.G * W;
  const int grid = (int)((threads + DIR_THREADS - 1) / DIR_THREADS);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define BEAN_DIR_LAUNCH(WW)                                                                   \
  if (GRAD) dirichlet_rsample_grad_kernel<real, WW><<<grid, DIR_THREADS, 0, st>>>(p);         \
  else dirichlet_rsample_kernel<real, WW><<<grid, DIR_THREADS, 0, st>>>(p)
  switch (W) {
    case 2: BEAN_DIR_LAUNCH(2); break;
    case 4: BEAN_DIR_LAUNCH(4); break;
    case 8: BEAN_DIR_LAUNCH(8); break;
    case 16: BEAN_DIR_LAUNCH(16); break;
    default: BEAN_DIR_LAUNCH(32); break;
  }
#undef BEAN_DIR_LAUNCH
  BEAN_CUDA(cudaPeekAtLastError());
  return BEAN_OK;
}

}  // namespace bean

extern "C" {
int bean_dirichlet_rsample_f32(const BeanDirichletArgs* a, void* s) { return bean::launch_dirichlet<float, false>(a, s); }
int bean_dirichlet_rsample_f64(const BeanDirichletArgs* a, void* s) { return bean::launch_dirichlet<double, false>(a, s); }
int bean_dirichlet_rsample_grad_f32(const BeanDirichletArgs* a, void* s) { return bean::launch_dirichlet<float, true>(a, s); }
int bean_dirichlet_rsample_grad_f64(const BeanDirichletArgs* a, void* s) { return bean::launch_dirichlet<double, true>(a, s); }
}
