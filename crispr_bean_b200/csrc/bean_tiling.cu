// bean_tiling.cu -- allele <- edit contraction of the tiling (MultiMixtureNormal) models as CSR gather /
// CSC segmented scatter.
//
// Replaces `mu_alleles = allele_to_edit @ mu_edits`, `sd_alleles = ||allele_to_edit * sd_edits||_2` and the
// WT column (0, 1) (bean/model/model.py:618-625, :907-918; survival_model.py:484-493) and their autograd
// backward.  The reference's `allele_to_edit` is a dense 0/1 tensor (G, A-1, E) -- 429 MB and > 99.9 % zeros
// at the published tiling example -- so the "GEMM" is a sparse gather; here it is stored as
//   CSR  allele_ptr [G*(A-1)+1], allele_edit [nnz]     (slot = g*(A-1) + (a-1) -> its edits)
//   CSC  edit_ptr   [E+1],       edit_slot   [nnz]     (edit -> the slots that contain it)
// and the backward is a per-edit segmented sum over the CSC view (deterministic, no atomics).
#include "bean_common.cuh"
#include "bean_math.cuh"

namespace bean {

constexpr int TILE_THREADS = 256;

template <typename real>
__global__ void __launch_bounds__(TILE_THREADS) allele_gather_kernel(int G, int A, const int32_t* __restrict__ allele_ptr,
                                                                      const int32_t* __restrict__ allele_edit,
                                                                      const real* __restrict__ mu_edit,
                                                                      const real* __restrict__ sd_edit, real* mu_allele,
                                                                      real* sd_allele) {
  const long long i = (long long)blockIdx.x * TILE_THREADS + threadIdx.x;  // index into [G][A]
  if (i >= (long long)G * A) return;
  const int g = (int)(i / A), a = (int)(i % A);
  if (a == 0) {  // wild type: N(0, 1)
    mu_allele[i] = real(0);
    sd_allele[i] = real(1);
    return;
  }
  const long long slot = (long long)g * (A - 1) + (a - 1);
  real m = real(0), v = real(0);
  for (int k = allele_ptr[slot]; k < allele_ptr[slot + 1]; ++k) {
    const int e = allele_edit[k];
    m += mu_edit[e];
    const real s = sd_edit[e];
    v += s * s;
  }
  mu_allele[i] = m;
  sd_allele[i] = Num<real>::sqrt(v);
}

template <typename real>
__global__ void __launch_bounds__(TILE_THREADS) allele_scatter_kernel(int E, int A, const int32_t* __restrict__ edit_ptr,
                                                                       const int32_t* __restrict__ edit_slot,
                                                                       const real* __restrict__ sd_edit,
                                                                       const real* __restrict__ sd_allele,
                                                                       const real* __restrict__ d_mu_allele,
                                                                       const real* __restrict__ d_sd_allele, real* d_mu_edit,
                                                                       real* d_sd_edit) {
  const int e = blockIdx.x * TILE_THREADS + threadIdx.x;
  if (e >= E) return;
  real dm = real(0), ds = real(0);
  const real s = sd_edit[e];
  for (int k = edit_ptr[e]; k < edit_ptr[e + 1]; ++k) {
    const long long slot = edit_slot[k];
    const long long i = (slot / (A - 1)) * A + (slot % (A - 1)) + 1;  // [g][a]
    dm += d_mu_allele[i];
    const real n = sd_allele[i];
    if (n > real(0)) ds += d_sd_allele[i] * s / n;  // d||v||/dv = v/||v||, subgradient 0 at 0
  }
  d_mu_edit[e] = dm;
  d_sd_edit[e] = ds;
}

static int check_map(const BeanAlleleMap* m) {
  BEAN_REQUIRE(m != nullptr, BEAN_EINVAL, "allele map is NULL");
  BEAN_REQUIRE(m->n_guides > 0 && m->n_alleles >= 2 && m->n_edits > 0, BEAN_EINVAL, "bad allele map sizes");
  BEAN_REQUIRE(m->allele_ptr && m->allele_edit && m->edit_ptr && m->edit_slot, BEAN_EINVAL, "allele map arrays must be non-NULL");
  return BEAN_OK;
}

template <typename real>
static int gather(const BeanAlleleMap* m, const void* mu_edit, const void* sd_edit, void* mu_allele, void* sd_allele, void* stream) {
  int rc = check_map(m);
  if (rc != BEAN_OK) return rc;
  BEAN_REQUIRE(mu_edit && sd_edit && mu_allele && sd_allele, BEAN_EINVAL, "gather buffers must be non-NULL");
  const long long n = (long long)m->n_guides * m->n_alleles;
  allele_gather_kernel<real><<<(unsigned)((n + TILE_THREADS - 1) / TILE_THREADS), TILE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      m->n_guides, m->n_alleles, m->allele_ptr, m->allele_edit, static_cast<const real*>(mu_edit), static_cast<const real*>(sd_edit),
      static_cast<real*>(mu_allele), static_cast<real*>(sd_allele));
  BEAN_CUDA(cudaPeekAtLastError());
  return BEAN_OK;
}

template <typename real>
static int scatter(const BeanAlleleMap* m, const void* sd_edit, const void* sd_allele, const void* d_mu_allele,
                   const void* d_sd_allele, void* d_mu_edit, void* d_sd_edit, void* stream) {
  int rc = check_map(m);
  if (rc != BEAN_OK) return rc;
  BEAN_REQUIRE(sd_edit && sd_allele && d_mu_allele && d_sd_allele && d_mu_edit && d_sd_edit, BEAN_EINVAL, "scatter buffers must be non-NULL");
  allele_scatter_kernel<real><<<(m->n_edits + TILE_THREADS - 1) / TILE_THREADS, TILE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      m->n_edits, m->n_alleles, m->edit_ptr, m->edit_slot, static_cast<const real*>(sd_edit), static_cast<const real*>(sd_allele),
      static_cast<const real*>(d_mu_allele), static_cast<const real*>(d_sd_allele), static_cast<real*>(d_mu_edit),
      static_cast<real*>(d_sd_edit));
  BEAN_CUDA(cudaPeekAtLastError());
  return BEAN_OK;
}

}  // namespace bean

extern "C" {
int bean_allele_gather_f32(const BeanAlleleMap* m, const void* mu, const void* sd, void* mu_a, void* sd_a, void* st) {
  return bean::gather<float>(m, mu, sd, mu_a, sd_a, st);
}
int bean_allele_gather_f64(const BeanAlleleMap* m, const void* mu, const void* sd, void* mu_a, void* sd_a, void* st) {
  return bean::gather<double>(m, mu, sd, mu_a, sd_a, st);
}
int bean_allele_scatter_f32(const BeanAlleleMap* m, const void* sd, const void* sd_a, const void* dmu_a, const void* dsd_a,
                            void* dmu, void* dsd, void* st) {
  return bean::scatter<float>(m, sd, sd_a, dmu_a, dsd_a, dmu, dsd, st);
}
int bean_allele_scatter_f64(const BeanAlleleMap* m, const void* sd, const void* sd_a, const void* dmu_a, const void* dsd_a,
                            void* dmu, void* dsd, void* st) {
  return bean::scatter<double>(m, sd, sd_a, dmu_a, dsd_a, dmu, dsd, st);
}
}
