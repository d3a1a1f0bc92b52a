// bean_common.cuh -- shared host-side plumbing of the C-ABI (error slot, argument checks, tables).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/bean_b200.h"

namespace bean {

// last error message of the calling thread (returned by bean_last_error)
char* err_slot();
int fail(int code, const char* fmt, ...);

#define BEAN_REQUIRE(cond, code, ...)            \
  do {                                           \
    if (!(cond)) return ::bean::fail(code, __VA_ARGS__); \
  } while (0)

#define BEAN_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess)                                                          \
      return ::bean::fail(BEAN_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Per-sample tables travel by value inside the kernel argument block (< 4 KB).
template <typename real>
struct SampleTables {
  real sf[BEAN_MAX_LAYERS][BEAN_MAX_RB];  // size_factor[l][r*B + b]
  real smask[BEAN_MAX_RB];                // sample_mask[r*B + b]
  real thr_u[BEAN_MAX_BINS];              // sorting: upper threshold (+inf = quantile 1)
  real thr_l[BEAN_MAX_BINS];              // sorting: lower threshold (-inf = quantile 0)
  real tp[BEAN_MAX_BINS];                 // survival: timepoints
};

int validate_screen(const BeanScreen* s);

template <typename real>
inline void fill_tables(const BeanScreen* s, SampleTables<real>& t) {
  const int RB = s->n_reps * s->n_bins;
  for (int l = 0; l < BEAN_MAX_LAYERS; ++l)
    for (int i = 0; i < BEAN_MAX_RB; ++i)
      t.sf[l][i] = (l < s->n_layers && i < RB) ? real(s->size_factor[l * RB + i]) : real(0);
  for (int i = 0; i < BEAN_MAX_RB; ++i) t.smask[i] = i < RB ? real(s->sample_mask[i]) : real(0);
  for (int b = 0; b < BEAN_MAX_BINS; ++b) {
    const bool ok = b < s->n_bins;
    t.thr_u[b] = (ok && s->mode == BEAN_MODE_SORTING) ? real(s->upper_thres[b]) : real(0);
    t.thr_l[b] = (ok && s->mode == BEAN_MODE_SORTING) ? real(s->lower_thres[b]) : real(0);
    t.tp[b] = (ok && s->mode == BEAN_MODE_SURVIVAL) ? real(s->timepoints[b]) : real(0);
  }
}

}  // namespace bean
