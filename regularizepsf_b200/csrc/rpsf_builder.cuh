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

// ============================================================================ core isolation (round 2)
// Reference: regularizepsf/image_processing.py:13-46 (calculate_background) and regularizepsf/builder.py:236-258
// (the per-patch tail of ArrayPSFBuilder.build).  One CTA per averaged patch, the patch and two byte masks in global
// memory (any patch size), every step a strided pass over the pixels with a CTA barrier in between:
//   ring   = dilate(core) & ~core of the eroded non-zero interior, fainter than the centre       (:29-39)
//   plane  = least squares c0 * col + c1 * row + c2 through the ring, NaN where the patch is 0       (:41-44)
//   patch -= plane; zeros -> NaN; pixels below 0.5 % of the centre whose four neighbours are too -> NaN;
//   non-finite -> 0; keep the 4-connected island of non-zero pixels that holds the centre, dilated by one pixel;
//   divide by the sum.                                                                      (builder.py:239-258)
// The masks are integer decisions and follow the reference exactly; the plane comes from the 3 x 3 normal equations
// (eigen-decomposition, pseudo-inverse when the ring is degenerate = LAPACK gelsd's minimum-norm answer) instead of
// an SVD of the n x 3 design matrix, and the final sum is a tree instead of numpy's pairwise order: the result agrees
// with the reference to a few 1e-13 of the patch maximum, not bit for bit (tolerance stated in the tests).
constexpr int ISO_TPB = 256;

// sum of K doubles per thread over the CTA, fixed order (lane tree, then warps in order): run-to-run stable
template <int K>
__device__ __forceinline__ void iso_block_sum(double (&v)[K], double* red /* shared [K * (ISO_TPB / 32)] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    if (lane == 0) red[k * (ISO_TPB / 32) + warp] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < ISO_TPB / 32; ++w) acc += red[k * (ISO_TPB / 32) + w];
    v[k] = acc;
  }
  __syncthreads();
}

// eigen-decomposition of a symmetric 3 x 3 matrix by cyclic Jacobi rotations: a -> diagonal, q -> eigenvectors (columns)
__device__ inline void iso_jacobi3(double a[3][3], double q[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) q[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 24; ++sweep) {
    const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (off == 0.0) break;
    for (int pi = 0; pi < 2; ++pi)
      for (int qi = pi + 1; qi < 3; ++qi) {
        if (a[pi][qi] == 0.0) continue;
        const double theta = (a[qi][qi] - a[pi][pi]) / (2.0 * a[pi][qi]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < 3; ++k) {            // columns pi, qi of a
          const double akp = a[k][pi], akq = a[k][qi];
          a[k][pi] = c * akp - sn * akq;
          a[k][qi] = sn * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {            // rows pi, qi of a
          const double apk = a[pi][k], aqk = a[qi][k];
          a[pi][k] = c * apk - sn * aqk;
          a[qi][k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double qkp = q[k][pi], qkq = q[k][qi];
          q[k][pi] = c * qkp - sn * qkq;
          q[k][qi] = sn * qkp + c * qkq;
        }
      }
  }
}
// minimum-norm least-squares solution of (m) x = b for a symmetric positive semi-definite m; returns the rank
__device__ inline int iso_solve3(const double m[3][3], const double b[3], double x[3]) {
  double a[3][3], q[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = m[i][j];
  iso_jacobi3(a, q);
  const double top = fmax(fmax(fabs(a[0][0]), fabs(a[1][1])), fabs(a[2][2]));
  x[0] = x[1] = x[2] = 0.0;
  int rank = 0;
  for (int k = 0; k < 3; ++k) {
    if (!(fabs(a[k][k]) > 1e-11 * top)) continue;         // eigenvalue = (singular value)^2 of the design matrix
    ++rank;
    const double proj = (q[0][k] * b[0] + q[1][k] * b[1] + q[2][k] * b[2]) / a[k][k];
    for (int i = 0; i < 3; ++i) x[i] += proj * q[i][k];
  }
  return rank;
}

// image_processing.py:13-46 for the patch at `v` (P x P, row-major).  `core` is a P*P byte scratch.  Returns the plane
// coefficients to every thread: background(row, col) = coef[0] * col + coef[1] * row + coef[2].
__device__ inline void iso_plane_fit(const double* __restrict__ v, int P, unsigned char* __restrict__ core, double* red,
                                     double* coef_s /* shared [3] */, double (&coef)[3]) {
  const int pp = P * P;
  auto nz = [&](int r, int c) { return r >= 0 && r < P && c >= 0 && c < P && v[r * P + c] != 0.0; };   // NaN != 0
  for (int i = threadIdx.x; i < pp; i += ISO_TPB) {
    const int r = i / P, c = i % P;
    const bool inner = r > 0 && r < P - 1 && c > 0 && c < P - 1;
    core[i] = inner && nz(r, c) && nz(r - 1, c) && nz(r + 1, c) && nz(r, c - 1) && nz(r, c + 1);
  }
  __syncthreads();
  const double centre = v[(P / 2) * P + P / 2];
  const double mid = 0.5 * (P - 1);
  // [0..8] centred sums (n, x, y, xx, xy, yy, v, xv, yv), [9..15] the same moments of the raw coordinates
  double sums[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) sums[k] = 0.0;
  auto in_core = [&](int r, int c) { return r >= 0 && r < P && c >= 0 && c < P && core[r * P + c] != 0; };
  for (int i = threadIdx.x; i < pp; i += ISO_TPB) {
    const int r = i / P, c = i % P;
    const bool grown = in_core(r, c) || in_core(r - 1, c) || in_core(r + 1, c) || in_core(r, c - 1) || in_core(r, c + 1);
    const double val = v[i];
    if (grown && !core[i] && val < centre) {
      const double x = c - mid, y = r - mid, ux = c, uy = r;
      sums[0] += 1.0; sums[1] += x; sums[2] += y; sums[3] += x * x; sums[4] += x * y; sums[5] += y * y;
      sums[6] += val; sums[7] += x * val; sums[8] += y * val;
      sums[9] += ux; sums[10] += uy; sums[11] += ux * ux; sums[12] += ux * uy; sums[13] += uy * uy;
      sums[14] += ux * val; sums[15] += uy * val;
    }
  }
  iso_block_sum<16>(sums, red);
  if (threadIdx.x == 0) {
    double x[3] = {0.0, 0.0, 0.0};
    if (sums[0] > 0.0) {
      const double m[3][3] = {{sums[3], sums[4], sums[1]}, {sums[4], sums[5], sums[2]}, {sums[1], sums[2], sums[0]}};
      const double b[3] = {sums[7], sums[8], sums[6]};
      if (iso_solve3(m, b, x) == 3) {
        x[2] -= x[0] * mid + x[1] * mid;                  // centred -> raw coordinates (unique solution: any basis)
      } else {                                            // degenerate ring: minimum norm in the reference's own basis
        const double mu[3][3] = {{sums[11], sums[12], sums[9]}, {sums[12], sums[13], sums[10]}, {sums[9], sums[10], sums[0]}};
        const double bu[3] = {sums[14], sums[15], sums[6]};
        iso_solve3(mu, bu, x);
      }
    }
    coef_s[0] = x[0]; coef_s[1] = x[1]; coef_s[2] = x[2];
  }
  __syncthreads();
  coef[0] = coef_s[0]; coef[1] = coef_s[1]; coef[2] = coef_s[2];
}
__device__ __forceinline__ double iso_plane_at(const double (&coef)[3], int r, int c) {
  // coefficients[0] * patch_x + coefficients[1] * patch_y + coefficients[2], rounded term by term like numpy
  return __dadd_rn(__dadd_rn(__dmul_rn(coef[0], (double)c), __dmul_rn(coef[1], (double)r)), coef[2]);
}

// calculate_background alone (used by tests and by callers that only want the plane): out = plane, NaN where patch == 0
__global__ void __launch_bounds__(ISO_TPB)
plane_background(const double* __restrict__ patches, double* __restrict__ out, unsigned char* __restrict__ scratch, int P) {
  __shared__ double red[16 * (ISO_TPB / 32)];
  __shared__ double coef_s[3];
  const int pp = P * P;
  const double* v = patches + (size_t)blockIdx.x * pp;
  double coef[3];
  iso_plane_fit(v, P, scratch + (size_t)blockIdx.x * pp, red, coef_s, coef);
  for (int i = threadIdx.x; i < pp; i += ISO_TPB)
    out[(size_t)blockIdx.x * pp + i] = v[i] == 0.0 ? __longlong_as_double(0x7ff8000000000000ll) : iso_plane_at(coef, i / P, i % P);
}

// builder.py:236-258 in place.  scratch: 2 * P * P bytes per patch.
__global__ void __launch_bounds__(ISO_TPB)
isolate_cores(double* __restrict__ patches, unsigned char* __restrict__ scratch, int P) {
  __shared__ double red[16 * (ISO_TPB / 32)];
  __shared__ double coef_s[3];
  const int pp = P * P;
  double* v = patches + (size_t)blockIdx.x * pp;
  unsigned char* m1 = scratch + (size_t)blockIdx.x * 2 * pp;
  unsigned char* m2 = m1 + pp;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  double coef[3];
  iso_plane_fit(v, P, m1, red, coef_s, coef);
  // patch -= background (NaN where the patch is 0); patch[patch == 0] = nan                       builder.py:239-242
  for (int i = threadIdx.x; i < pp; i += ISO_TPB) {
    const double val = v[i];
    const double d = val == 0.0 ? nan : __dsub_rn(val, iso_plane_at(coef, i / P, i % P));
    v[i] = d == 0.0 ? nan : d;
  }
  __syncthreads();
  // faint pixels: below 0.5 % of the centre, eroded with the outside counted as faint             builder.py:244-248
  const double limit = __dmul_rn(0.005, v[(P / 2) * P + P / 2]);
  for (int i = threadIdx.x; i < pp; i += ISO_TPB) m1[i] = v[i] < limit;
  __syncthreads();
  auto faint = [&](int r, int c) { return r < 0 || r >= P || c < 0 || c >= P || m1[r * P + c] != 0; };
  for (int i = threadIdx.x; i < pp; i += ISO_TPB) {
    const int r = i / P, c = i % P;
    m2[i] = faint(r, c) && faint(r - 1, c) && faint(r + 1, c) && faint(r, c - 1) && faint(r, c + 1);
  }
  __syncthreads();
  // non-finite and faint -> 0                                                                     builder.py:248-251
  for (int i = threadIdx.x; i < pp; i += ISO_TPB) {
    const double d = v[i];
    v[i] = (m2[i] || !isfinite(d)) ? 0.0 : d;
  }
  __syncthreads();
  // the island of the centre pixel: 4-connected non-zero pixels (scipy.ndimage.label's default structure); when the
  // centre itself is 0 its label is the background's, and the reference keeps every zero pixel    builder.py:253-254
  const int ci = (P / 2) * P + P / 2;
  const bool centre_set = v[ci] != 0.0;
  for (int i = threadIdx.x; i < pp; i += ISO_TPB) m1[i] = centre_set ? (i == ci) : (v[i] == 0.0);
  __syncthreads();
  if (centre_set) {
    auto in = [&](int r, int c) { return r >= 0 && r < P && c >= 0 && c < P && m1[r * P + c] != 0; };
    for (;;) {
      int changed = 0;
      for (int i = threadIdx.x; i < pp; i += ISO_TPB) {
        if (m1[i] || v[i] == 0.0) continue;
        const int r = i / P, c = i % P;
        if (in(r - 1, c) || in(r + 1, c) || in(r, c - 1) || in(r, c + 1)) { m1[i] = 1; changed = 1; }
      }
      if (!__syncthreads_or(changed)) break;
    }
  }
  // dilate by one pixel, cut, normalise                                                           builder.py:256-259
  auto in = [&](int r, int c) { return r >= 0 && r < P && c >= 0 && c < P && m1[r * P + c] != 0; };
  double total[1] = {0.0};
  for (int i = threadIdx.x; i < pp; i += ISO_TPB) {
    const int r = i / P, c = i % P;
    const bool keep = in(r, c) || in(r - 1, c) || in(r + 1, c) || in(r, c - 1) || in(r, c + 1);
    const double kept = __dmul_rn(v[i], keep ? 1.0 : 0.0);
    m2[i] = keep;
    total[0] += kept;
  }
  iso_block_sum<1>(total, red);
  for (int i = threadIdx.x; i < pp; i += ISO_TPB) v[i] = __ddiv_rn(__dmul_rn(v[i], m2[i] ? 1.0 : 0.0), total[0]);
}

// ============================================================================ star cutouts (round 2)
// Reference: regularizepsf/image_processing.py:78-121, the per-star body of _find_patches.  One CTA per detected star:
//   window   width x width pixels of the frame at the rounded corner, through np.pad(mode="reflect")     (:79-99)
//   shift    scipy.ndimage.shift(order=3, mode="mirror") by (-corner + round(corner) - 0.5): cubic B-spline
//            prefilter along both axes (pole sqrt(3) - 2, mirror initialisation) and 4 x 4 interpolation      (:100-101)
//   plane    calculate_background of the shifted patch, subtracted; NaN where the shifted patch is 0   (:108-110)
//   accept   every pixel below the saturation threshold (a NaN fails, as `nan < t` does) and the centre inside
//            (star_minimum, star_maximum)                                                              (:114-118)
// The arithmetic is the same exact-interpolation spline scipy computes, in float64, but not scipy's instruction
// order: values agree to ~1e-15 of the patch maximum.  (The pixel-mask patch of :102-106 is NOT computed here: the
// reference casts its spline-shifted values to bool by truncation, which only scipy itself reproduces; the host
// applies it.)
constexpr double SPLINE_POLE = -0.267949192431122706472553658494127633;   // sqrt(3) - 2

__device__ __forceinline__ int reflect_index(long long i, int n) {        // np.pad "reflect" / ndimage "mirror"
  if (n <= 1) return 0;
  const long long period = 2LL * (n - 1);
  long long m = i % period;
  if (m < 0) m += period;
  return (int)(m < n ? m : period - m);
}
__device__ __forceinline__ double mirror_coordinate(double x, int n) {
  if (n <= 1) return 0.0;
  const double period = 2.0 * (n - 1);
  double m = fmod(x, period);
  if (m < 0.0) m += period;
  return m <= (double)(n - 1) ? m : period - m;
}
// cubic B-spline coefficients of one line in place (gain 6, causal and anti-causal recursion, mirror boundaries)
__device__ inline void spline_prefilter_line(double* c, int n, int stride) {
  if (n < 2) return;
  const double z = SPLINE_POLE;
  for (int i = 0; i < n; ++i) c[(size_t)i * stride] *= 6.0;
  const double zn1 = pow(z, (double)(n - 1));
  double c0 = c[0] + zn1 * c[(size_t)(n - 1) * stride], zi = z;
  for (int i = 1; i < n - 1; ++i) {
    c0 += zi * (c[(size_t)i * stride] + zn1 * c[(size_t)(n - 1 - i) * stride]);
    zi *= z;
  }
  c[0] = c0 / (1.0 - zn1 * zn1);
  for (int i = 1; i < n; ++i) c[(size_t)i * stride] += z * c[(size_t)(i - 1) * stride];
  c[(size_t)(n - 1) * stride] = (z * c[(size_t)(n - 2) * stride] + c[(size_t)(n - 1) * stride]) * z / (z * z - 1.0);
  for (int i = n - 2; i >= 0; --i) c[(size_t)i * stride] = z * (c[(size_t)(i + 1) * stride] - c[(size_t)i * stride]);
}
struct SplineTap { int idx[4]; double w[4]; };
__device__ __forceinline__ SplineTap spline_taps(int k, double shift, int n) {
  SplineTap t;
  const double cc = mirror_coordinate((double)k - shift, n);
  const double fl = floor(cc);
  const double x = cc - fl, y = 1.0 - x;
  t.w[0] = y * y * y / 6.0;
  t.w[1] = (x * x * (x - 2.0) * 3.0 + 4.0) / 6.0;
  t.w[2] = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
  t.w[3] = 1.0 - t.w[0] - t.w[1] - t.w[2];
  const long long start = (long long)fl - 1;
#pragma unroll
  for (int l = 0; l < 4; ++l) t.idx[l] = reflect_index(start + l, n);
  return t;
}

template <typename TIn>
__global__ void __launch_bounds__(ISO_TPB)
star_cutouts(const TIn* __restrict__ frame, int H, int W, const double* __restrict__ corners, int P, double sat, double smin,
             double smax, double* __restrict__ out, unsigned char* __restrict__ accepted, double* __restrict__ coeffs,
             unsigned char* __restrict__ scratch) {
  __shared__ double red[16 * (ISO_TPB / 32)];
  __shared__ double coef_s[3];
  const int pp = P * P;
  const size_t star = blockIdx.x;
  double* cf = coeffs + star * pp;
  double* v = out + star * pp;
  const double cr = corners[2 * star], cc = corners[2 * star + 1];
  const double rr = rint(cr), rc = rint(cc);                       // Python round(): half to even
  const long long r0 = (long long)rr, c0 = (long long)rc;
  const double shift_r = -cr + rr - 0.5, shift_c = -cc + rc - 0.5;
  for (int i = threadIdx.x; i < pp; i += ISO_TPB) {
    const int r = reflect_index(r0 + i / P, H), c = reflect_index(c0 + i % P, W);
    cf[i] = (double)frame[(size_t)r * W + c];
  }
  __syncthreads();
  for (int line = threadIdx.x; line < P; line += ISO_TPB) spline_prefilter_line(cf + line, P, P);        // axis 0
  __syncthreads();
  for (int line = threadIdx.x; line < P; line += ISO_TPB) spline_prefilter_line(cf + (size_t)line * P, P, 1);   // axis 1
  __syncthreads();
  for (int i = threadIdx.x; i < pp; i += ISO_TPB) {
    const SplineTap tr = spline_taps(i / P, shift_r, P), tc = spline_taps(i % P, shift_c, P);
    double t = 0.0;
#pragma unroll
    for (int l = 0; l < 4; ++l)
#pragma unroll
      for (int m = 0; m < 4; ++m) t += cf[(size_t)tr.idx[l] * P + tc.idx[m]] * (tr.w[l] * tc.w[m]);
    v[i] = (double)(TIn)t;                                         // scipy writes the frame's dtype
  }
  __syncthreads();
  double coef[3];
  iso_plane_fit(v, P, scratch + star * pp, red, coef_s, coef);
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  int bad = 0;
  for (int i = threadIdx.x; i < pp; i += ISO_TPB) {
    const double val = v[i];
    const double d = val == 0.0 ? nan : __dsub_rn(val, iso_plane_at(coef, i / P, i % P));
    v[i] = d;
    bad |= !(d < sat);
  }
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) {
    const double centre = v[(P / 2) * P + P / 2];
    accepted[star] = !bad && centre > smin && centre < smax;
  }
}

}  // namespace rpsf
