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
// Stacks of up to 416 cutouts are staged once into shared memory (64 pixels per CTA); larger ones
// into a global scratch with the same [item][pixel] layout.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rpsf {

constexpr int AVG_TPB = 64;                    // pixels per CTA
constexpr int AVG_SUBS = 8;                    // threads per pixel for deep stacks
constexpr int AVG_SHALLOW = 32;                // up to here: one thread per pixel
constexpr int AVG_MAX_STAGED = 416;            // 416 items x 68 keys x 8 B = 221 KB of shared memory

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
#pragma unroll 4
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
//
// SUBS threads share one pixel: thread `sub` stages and counts the items i = sub (mod SUBS) and the
// partial counts meet in a shuffle butterfly, so a deep stack — which leaves room for only one CTA of
// 64 pixels in an SM's shared memory — still runs 16 warps.  The SUBS lanes of a pixel sit in one
// warp (lane = sub * 4 + pixel % 4) and the key matrix rows are padded to 68 keys, which spreads the
// 8 rows a warp reads at once over all banks.  Shallow stacks use SUBS = 1 (one thread per pixel).
template <int SUBS> struct AvgLayout {
  static constexpr int PX_PER_WARP = 32 / SUBS;
  static constexpr int THREADS = AVG_TPB * SUBS;
  static constexpr int STRIDE = SUBS == 1 ? AVG_TPB : AVG_TPB + 4;      // keys per item row
};

template <bool STAGED, int SUBS>
__global__ void __launch_bounds__(AvgLayout<SUBS>::THREADS)
average_select(const double* __restrict__ cutouts, const long long* __restrict__ offsets,
               const int* __restrict__ items, const int* __restrict__ cells, int pp, int centre, int method,
               double quantile, unsigned long long* __restrict__ scratch,
               const long long* __restrict__ scratch_off, double* __restrict__ out) {
  using L = AvgLayout<SUBS>;
  extern __shared__ __align__(16) unsigned long long staged[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = SUBS == 1 ? 0 : lane / L::PX_PER_WARP;
  const int px = warp * L::PX_PER_WARP + (SUBS == 1 ? lane : lane % L::PX_PER_WARP);   // pixel inside the CTA
  const int cell = cells[blockIdx.y];
  const int pix = blockIdx.x * AVG_TPB + px;
  const bool live = pix < pp;
  const int col = live ? pix : pp - 1;                  // idle lanes shadow a real pixel, store nothing
  const long long b = offsets[cell];
  const int n = (int)(offsets[cell + 1] - b);
  unsigned long long* keys;
  if constexpr (STAGED) keys = staged + px;
  else keys = scratch + scratch_off[blockIdx.y] + (long long)blockIdx.x * L::STRIDE * n + px;
  auto all_subs = [](auto v, auto op) {                 // butterfly over the SUBS lanes of a pixel
#pragma unroll
    for (int d = L::PX_PER_WARP; d < 32; d <<= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
  };
  auto add = [](int x, int y) { return x + y; };

  int m = 0;                                            // non-NaN values of this pixel's stack
  for (int i = sub; i < n; i += SUBS) {
    const double* cut = cutouts + (long long)items[b + i] * pp;
    const double v = cut[col] / cut[centre];            // builder.py:90
    const bool nan = isnan(v);
    keys[(long long)i * L::STRIDE] = nan ? NAN_KEY : order_key(v);
    m += nan ? 0 : 1;
  }
  if constexpr (SUBS > 1) { m = all_subs(m, add); __syncwarp(); }
  // Every lane runs every pass — the lanes of a warp belong to different pixels, and the shuffles
  // below need all of them — an empty stack (m == 0) just has its result discarded.
  int k1 = 0, k2 = 0;
  double gamma = 0.0;
  if (m > 0) {
    if (method == AVG_MEDIAN) {
      k1 = (m - 1) / 2; k2 = m / 2;
    } else {
      const double vi = __dmul_rn((double)(m - 1), quantile);
      if (vi >= (double)(m - 1)) { k1 = k2 = m - 1; gamma = __dadd_rn(vi, 1.0); }   // previous = next = -1
      else { k1 = (int)floor(vi); k2 = k1 + 1; gamma = __dsub_rn(vi, (double)k1); }
    }
  }
  // largest K with count(keys < K) <= k1 is the k1-th smallest key
  unsigned long long lo = 0;
  for (int bit = 63; bit >= 0; --bit) {
    const unsigned long long trial = lo | (1ull << bit);
    int c = 0;
#pragma unroll 4
    for (int i = sub; i < n; i += SUBS) c += keys[(long long)i * L::STRIDE] < trial ? 1 : 0;
    if constexpr (SUBS > 1) c = all_subs(c, add);
    if (c <= k1) lo = trial;
  }
  // the next order statistic: lo again if it is repeated often enough, else the smallest key above it
  int le = 0;
  unsigned long long above = NAN_KEY;
  for (int i = sub; i < n; i += SUBS) {
    const unsigned long long k = keys[(long long)i * L::STRIDE];
    le += k <= lo ? 1 : 0;
    if (k > lo && k < above) above = k;
  }
  if constexpr (SUBS > 1) {
    le = all_subs(le, add);
    above = all_subs(above, [](unsigned long long x, unsigned long long y) { return x < y ? x : y; });
  }
  const unsigned long long hi = (k2 != k1 && le < k2 + 1) ? above : lo;
  const double a = key_value(lo), bb = key_value(hi);
  double r = method == AVG_MEDIAN ? __dadd_rn(a, bb) / 2.0 : lerp_like_numpy(a, bb, gamma);
  if (m == 0) r = __longlong_as_double(0x7ff8000000000000ll);      // all-NaN / empty stack -> NaN -> 0
  if (live && sub == 0) out[(long long)cell * pp + pix] = isnan(r) ? 0.0 : r;
}

}  // namespace rpsf
