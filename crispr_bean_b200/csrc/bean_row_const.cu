// bean_row_const.cu -- the data-only part of a row's Dirichlet-Multinomial / Multinomial log-pmf, once per screen.
//
//   row_const[i] = lgamma(N + 1) - sum_b lgamma(x_b + 1) [+ sum_{x_b > 0} x_b ln(x_b / N)],   N = sum_b x_b
//
// is what the reference recomputes inside every SVI step (L (B + 1) of the L (3B + 3) lgammas of a row, model.py:531-547 through
// pyro's DirichletMultinomial.log_prob; the Multinomial coefficient of the reporter counts, model.py:464-474) and what the
// fused steps hoist into `BeanSviConfig.ll_const` / `BeanScreen.row_const`.  It is on the END-TO-END path (screen upload ->
// first step): as torch ops on a 1M-guide screen it was 4 of the 12 ms between the host tensors and the first step -- a float64
// copy of the counts and ten passes over it.  One pass here: counts are integers, so lgamma(x + 1) = ln x! comes from a
// 4096-entry table (L1-resident) and only larger or non-integer entries take ::lgamma.
#include "bean_common.cuh"

namespace bean {

constexpr int ROWC_THREADS = 256;

__device__ __forceinline__ double log_factorial(double x, const double* table, int n_table) {
  if (x < (double)n_table && x == ::floor(x)) return table[(int)x];
  return ::lgamma(x + 1.0);
}

template <typename real>
__global__ void __launch_bounds__(ROWC_THREADS) row_const_kernel(const real* x, long long n_rows, int B, int with_xlogx, const double* table,
                                                                 int n_table, double* row_const, double* row_total) {
  const long long i = (long long)blockIdx.x * ROWC_THREADS + threadIdx.x;
  if (i >= n_rows) return;
  const real* xr = x + i * B;
  double N = 0.0;
  for (int b = 0; b < B; ++b) N += (double)xr[b];
  double rc = log_factorial(N, table, n_table);
  const double lN = with_xlogx ? ::log(N > 1.0 ? N : 1.0) : 0.0;  // x / max(N, 1), as the torch expression it replaces
  for (int b = 0; b < B; ++b) {
    const double xb = (double)xr[b];
    rc -= log_factorial(xb, table, n_table);
    if (with_xlogx && xb > 0.0) rc += xb * (::log(xb) - lN);
  }
  row_const[i] = rc;
  if (row_total) row_total[i] = N;
}

template <typename real>
static int row_const_run(const void* x, int64_t n_rows, int32_t n_bins, int32_t with_xlogx, const double* table, int32_t n_table,
                         double* row_const, double* row_total, void* stream) {
  BEAN_REQUIRE(x && row_const && table, BEAN_EINVAL, "x / row_const / log_factorial table is NULL");
  BEAN_REQUIRE(n_rows >= 0 && n_bins >= 1 && n_table >= 1, BEAN_EINVAL, "n_rows >= 0, n_bins >= 1, table_size >= 1");
  if (n_rows == 0) return BEAN_OK;
  const long long grid = (n_rows + ROWC_THREADS - 1) / ROWC_THREADS;
  BEAN_REQUIRE(grid < 2147483647LL, BEAN_EINVAL, "too many rows");
  row_const_kernel<real><<<(unsigned)grid, ROWC_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const real*>(x), (long long)n_rows, n_bins, with_xlogx, table, n_table, row_const, row_total);
  BEAN_CUDA(cudaPeekAtLastError());
  return BEAN_OK;
}

}  // namespace bean

extern "C" {
int bean_row_const_f32(const void* x, int64_t n_rows, int32_t n_bins, int32_t with_xlogx, const double* log_factorial, int32_t table_size,
                       double* row_const, double* row_total, void* stream) {
  return bean::row_const_run<float>(x, n_rows, n_bins, with_xlogx, log_factorial, table_size, row_const, row_total, stream);
}
int bean_row_const_f64(const void* x, int64_t n_rows, int32_t n_bins, int32_t with_xlogx, const double* log_factorial, int32_t table_size,
                       double* row_const, double* row_total, void* stream) {
  return bean::row_const_run<double>(x, n_rows, n_bins, with_xlogx, log_factorial, table_size, row_const, row_total, stream);
}
}
