// rpsf_builder.cuh — per-cell averaging of star cutouts for ArrayPSFBuilder.
//
// Reference: regularizepsf/builder.py:45-125 (_find_matches, _average_patches_by_mean,
// _average_patches_by_percentile, _average_patches).  Every star cutout is divided by its centre
// pixel (builder.py:63,90) and joins the stack of each covering cell its centre falls in; per pixel
// the stack is reduced with np.nansum / count of finite values (mean), np.nanmedian (median, or
// percentile == 50) or np.nanpercentile (linear interpolation); NaN results become 0
// (builder.py:118-122).
//
// All arithmetic is float64 and mirrors numpy operation by operation (no FMA contraction), so the
// result is bit-identical to the reference's:
//   mean        acc = nan_to_zero(acc) + nan_to_zero(v), in stack order; count += isfinite(v)
//   median      (v[(m-1)/2] + v[m/2]) / 2 over the m non-NaN values
//   percentile  numpy's "linear" method: virtual index (m-1)*q, gamma = index - floor(index),
//               _lerp(a, b, gamma) with its t >= 0.5 branch
//
// Order statistics are found without sorting: doubles are mapped to order-preserving 64-bit keys
// and the k-th smallest key is built bit by bit (64 counting passes over the stack), then one more
// pass gives the next distinct key.  Every thread owns one pixel of one cell and walks the stack in
// lockstep with its neighbours, so the passes are coalesced and divergence-free whatever the data.
// Stacks of up to 440 cutouts are staged once into shared memory (64 pixels per CTA); larger ones
// into a global scratch with the same [item][pixel] layout.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rpsf {

constexpr int AVG_TPB = 64;                    // pixels per CTA
constexpr int AVG_MAX_STAGED = 440;            // 440 items x 64 px x 8 B = 220 KB of shared memory

enum AvgMethod : int { AVG_MEAN = 0, AVG_MEDIAN = 1, AVG_PERCENTILE = 2 };

__device__ __forceinline__ unsigned long long order_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}
constexpr unsigned long long NAN_KEY = ~0ull;  // above +inf (0xfff0...): never counted, never selected

// builder.py:52-75.  One thread per (cell, pixel); items in stack (= dict insertion) order.
__global__ void average_mean(const double* __restrict__ cutouts, const long long* __restrict__ offsets,
                             const int* __restrict__ items, int pp, int centre, double* __restrict__ out) {
  const int cell = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= pp) return;
  const long long b = offsets[cell], e = offsets[cell + 1];
  double acc = 0.0, cnt = 0.0;
  for (long long i = b; i < e; ++i) {
    const double* cut = cutouts + (long long)items[i] * pp;
    const double v = cut[pix] / cut[centre];
    const double a0 = isnan(acc) ? 0.0 : acc;           // np.nansum([acc, v], axis=0)
    const double v0 = isnan(v) ? 0.0 : v;
    acc = __dadd_rn(a0, v0);
    if (isfinite(v)) cnt += 1.0;                        // accumulator_counts += np.isfinite(patch)
  }
  const double r = acc / cnt;                           // 0/0 -> NaN -> 0 (builder.py:118-122)
  out[(long long)cell * pp + pix] = isnan(r) ? 0.0 : r;
}

// numpy.lib._function_base_impl._lerp
__device__ __forceinline__ double lerp_like_numpy(double a, double b, double t) {
  const double d = __dsub_rn(b, a);
  double r = __dadd_rn(a, __dmul_rn(d, t));
  if (t >= 0.5) r = __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, t)));
  return r;
}

// builder.py:77-104.  `cells` lists the cells of this launch (all with at most `cap` items when
// STAGED); `scratch_off[j]` is the element offset of cell j's key matrix in `scratch` otherwise.
template <bool STAGED>
__global__ void __launch_bounds__(AVG_TPB)
average_select(const double* __restrict__ cutouts, const long long* __restrict__ offsets,
               const int* __restrict__ items, const int* __restrict__ cells, int pp, int centre, int method,
               double quantile, unsigned long long* __restrict__ scratch,
               const long long* __restrict__ scratch_off, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned long long staged[];
  const int cell = cells[blockIdx.y];
  const int pix = blockIdx.x * AVG_TPB + threadIdx.x;
  const bool live = pix < pp;
  const int col = live ? pix : pp - 1;                  // idle lanes shadow a real pixel, store nothing
  const long long b = offsets[cell];
  const int n = (int)(offsets[cell + 1] - b);
  unsigned long long* keys;
  if constexpr (STAGED) keys = staged + threadIdx.x;
  else keys = scratch + scratch_off[blockIdx.y] + (long long)blockIdx.x * AVG_TPB * n + threadIdx.x;

  int m = 0;                                            // non-NaN values of this pixel's stack
  for (int i = 0; i < n; ++i) {
    const double* cut = cutouts + (long long)items[b + i] * pp;
    const double v = cut[col] / cut[centre];            // builder.py:90
    const bool nan = isnan(v);
    keys[(long long)i * AVG_TPB] = nan ? NAN_KEY : order_key(v);
    m += nan ? 0 : 1;
  }
  double r = __longlong_as_double(0x7ff8000000000000ll);   // all-NaN / empty stack -> NaN -> 0
  if (m > 0) {
    int k1, k2;
    double gamma = 0.0;
    if (method == AVG_MEDIAN) {
      k1 = (m - 1) / 2; k2 = m / 2;
    } else {
      const double vi = __dmul_rn((double)(m - 1), quantile);
      if (vi >= (double)(m - 1)) { k1 = k2 = m - 1; gamma = __dadd_rn(vi, 1.0); }   // previous = next = -1
      else { k1 = (int)floor(vi); k2 = k1 + 1; gamma = __dsub_rn(vi, (double)k1); }
    }
    // largest K with count(keys < K) <= k1 is the k1-th smallest key
    unsigned long long lo = 0;
    for (int bit = 63; bit >= 0; --bit) {
      const unsigned long long trial = lo | (1ull << bit);
      int c = 0;
      for (int i = 0; i < n; ++i) c += keys[(long long)i * AVG_TPB] < trial ? 1 : 0;
      if (c <= k1) lo = trial;
    }
    unsigned long long hi = lo;
    if (k2 != k1) {
      int le = 0;
      unsigned long long above = NAN_KEY;
      for (int i = 0; i < n; ++i) {
        const unsigned long long k = keys[(long long)i * AVG_TPB];
        le += k <= lo ? 1 : 0;
        if (k > lo && k < above) above = k;
      }
      if (le < k2 + 1) hi = above;
    }
    const double a = key_value(lo), bb = key_value(hi);
    if (method == AVG_MEDIAN) r = __dadd_rn(a, bb) / 2.0;
    else r = lerp_like_numpy(a, bb, gamma);
  }
  if (live) out[(long long)cell * pp + pix] = isnan(r) ? 0.0 : r;
}

}  // namespace rpsf
