// bean_ceiling.cu -- register-only micro-kernel: the empirical FP32 / XU (SFU) ceiling of the likelihood rows.
//
// lgamma / digamma have no spec-sheet peak, so the compute roofline of the SVI step is
// MEASURED: this kernel evaluates, per guide, exactly the special-function work the Dirichlet-Multinomial rows
// of `svi_guide_kernel` need -- R * L rows of B bins through the same `dm_row_kl` (2B + 2 gamma corrections,
// 2B log1p, 1 log per row) -- on operands made up in registers, with no global loads and one store per CTA.
// Its duration is the floor for the row maths of one SVI step at the same launch geometry; bench.py reports
// the step against it next to the HBM roofline.  Nothing of the reference corresponds to it.
#include "bean_common.cuh"
#include "bean_math.cuh"
#include "bean_row.cuh"

namespace bean {

template <int NB>
__global__ void __launch_bounds__(128, 4) row_ceiling_kernel(int G, int rows, int nb, float seed, float* out) {
  __shared__ double red[32];
  const int g = blockIdx.x * 128 + threadIdx.x;
  float acc = 0.0f;
  if (g < G) {
    // counts ~ O(100), concentrations ~ O(1..10): the magnitudes the c5 screen has
    float u = seed + 1e-7f * (float)g;
    for (int r = 0; r < rows; ++r) {
      float x[NB], a[NB], gb[NB];
      float N = 0.0f, A = 0.0f;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        u = fmaf(u, 1.000173f, 0.377f);
        u = u > 4.0f ? u - 3.9f : u;  // keeps u in (0.3, 4.4): both branches of gamma_corr are exercised
        x[b] = b < nb ? floorf(40.0f * u + (float)(r + b)) : 0.0f;
        a[b] = b < nb ? 0.6f + 1.9f * u : 0.0f;
        N += x[b];
        A += a[b];
      }
      acc += dm_row_kl<float, NB>(nb, x, a, N, A, gb);
#pragma unroll
      for (int b = 0; b < NB; ++b) acc += gb[b];
    }
  }
  const double tot = block_sum((double)acc, red);
  if (threadIdx.x == 0) out[blockIdx.x] = (float)tot;
}

}  // namespace bean

extern "C" int bean_row_ceiling_f32(int32_t n_guides, int32_t n_rows_per_guide, int32_t n_bins, void* out, void* stream) {
  using namespace bean;
  BEAN_REQUIRE(n_guides > 0 && n_rows_per_guide > 0, BEAN_EINVAL, "n_guides / n_rows_per_guide must be > 0");
  BEAN_REQUIRE(n_bins > 0 && n_bins <= BEAN_MAX_BINS, BEAN_EINVAL, "n_bins must be in [1, %d]", BEAN_MAX_BINS);
  BEAN_REQUIRE(out != nullptr, BEAN_EINVAL, "out is NULL");
  const int grid = (n_guides + 127) / 128;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_bins <= 4)
    row_ceiling_kernel<4><<<grid, 128, 0, st>>>(n_guides, n_rows_per_guide, n_bins, 0.71f, static_cast<float*>(out));
  else
    row_ceiling_kernel<BEAN_MAX_BINS><<<grid, 128, 0, st>>>(n_guides, n_rows_per_guide, n_bins, 0.71f, static_cast<float*>(out));
  BEAN_CUDA(cudaPeekAtLastError());
  return BEAN_OK;
}
