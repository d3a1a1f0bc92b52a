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

// The same row maths in the mapping BASELINE.json's north_star describes -- one LANE per (row, bin) cell, a warp = 8 rows x 4
// bins, bin-axis reductions (N, A, the row value) by 4-lane shuffles -- for the A/B against the thread-per-guide mapping
// above.  Every lane evaluates its bin (one gamma_corr pair + one log1p pair) and, redundantly with its 3 neighbours, the
// row-level pair (A, U): per row that is 4 bin evaluations + 4 row-level evaluations in issue slots against 4 + 1 in the
// thread-per-guide mapping, plus 6 shuffles.  Requires 4 bins; n_rows_per_guide rows per guide are walked 8 at a time.
__global__ void __launch_bounds__(128, 4) row_ceiling_lanes_kernel(int G, int rows, float seed, float* out) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, b = lane & 3, rsub = lane >> 2;
  const int warp = (blockIdx.x * 128 + threadIdx.x) >> 5;  // one guide per warp-iteration
  const int n_warps = (gridDim.x * 128) >> 5;
  float acc = 0.0f;
  for (int g = warp; g < G; g += n_warps) {
    for (int r0 = 0; r0 < rows; r0 += 8) {
      const int r = r0 + rsub;
      // operands of cell (g, r, b) in the same range as row_ceiling_kernel's (u in (0.3, 4.4)), made up in ~5 instructions
      const float h = fmaf(0.6180339f, (float)(r * 4 + b), seed + 9.765625e-4f * (float)(g & 1023));
      const float u = fmaf(4.1f, h - floorf(h), 0.3f);
      const float x = floorf(40.0f * u + (float)(r + b)), a = 0.6f + 1.9f * u;
      float N = x, A = a;
      N += __shfl_xor_sync(0xffffffffu, N, 1); N += __shfl_xor_sync(0xffffffffu, N, 2);
      A += __shfl_xor_sync(0xffffffffu, A, 1); A += __shfl_xor_sync(0xffffffffu, A, 2);
      const float U = N + A;
      float2 cvAU, dlAU2, iAU, cv, dl, iz;
      gamma_corr2(make_float2(A, U), cvAU, dlAU2, iAU);
      const float iN = rcp_ftz(N), lUA = log_ftz(U * iAU.x);
      const float uu = x + a;
      gamma_corr2(make_float2(a, uu), cv, dl, iz);
      const float p = a * N, e = fmaf(a, N, -p), num = fmaf(x, A, -p) - e, t = num * iz.y;
      const float2 y = make_float2(-t * iAU.x, t * iN);
      float2 L = log1p_ratio_series2(y);
      if (!log1p_ratio_in_range2(y)) {
        const float Uu = U * iz.y;
        if (!log1p_ratio_in_range(y.x)) L.x = log_ftz(a * iAU.x * Uu);
        if (!log1p_ratio_in_range(y.y)) L.y = log_ftz(x * iN * Uu);
      }
      const float L2 = -L.x;
      float V = (x > 0.0f ? fmaf(-x, L.y, a * L2) : a * L2) - 0.5f * L2 + (cv.y - cv.x);
      V += __shfl_xor_sync(0xffffffffu, V, 1); V += __shfl_xor_sync(0xffffffffu, V, 2);
      V += -0.5f * 3.0f * lUA - (cvAU.y - cvAU.x);
      const float gb = L2 + ((dlAU2.x - dlAU2.y) + (dl.y - dl.x));
      if (r < rows) acc += gb + (b == 0 ? V : 0.0f);
    }
  }
  const double tot = block_sum((double)acc, red);
  if (threadIdx.x == 0) out[blockIdx.x] = (float)tot;
}

}  // namespace bean

extern "C" int bean_row_ceiling_lanes_f32(int32_t n_guides, int32_t n_rows_per_guide, void* out, void* stream) {
  using namespace bean;
  BEAN_REQUIRE(n_guides > 0 && n_rows_per_guide > 0, BEAN_EINVAL, "n_guides / n_rows_per_guide must be > 0");
  BEAN_REQUIRE(out != nullptr, BEAN_EINVAL, "out is NULL");
  // persistent: 148 SMs x 4 CTAs of 4 warps, each warp strides over the guides
  int sms = bean_device_sm_count();
  if (sms <= 0) sms = 148;
  const int grid = sms * 4;
  row_ceiling_lanes_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(n_guides, n_rows_per_guide, 0.71f, static_cast<float*>(out));
  BEAN_CUDA(cudaPeekAtLastError());
  return BEAN_OK;
}

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
