// rpsf_small.cuh — patches whose packed half-spectrum fits one SM's shared memory (P <= 128 in float32: 64 KB):
// the whole of transform.py:141-165 for one patch of one frame inside ONE CTA, nothing of the spectrum ever in HBM.
//
//   phase 1  gather + apodize + row FFT (same arithmetic as k1_gather_window_rowfft): teams of N1 threads take a
//            pair of patch rows each and leave their two packed half-spectra in the shared-memory plane S[P][P/2]
//   phase 2  column FFT x transfer kernel x column IFFT (same arithmetic as k2_colfft_mul_colifft), in place on S,
//            the transfer-kernel tile straight from the private layout in HBM / L2
//   phase 3  row IFFT + both windows (transform.py:164-165): the corrected patch, P x P reals, goes to a per-patch
//            plane W[frame][patch][P][P]
//
// and a second, elementwise kernel adds the planes of the patches that cover an output tile in colour order
// (transform.py:167-177: the reference's own `+=` order), writing every output pixel once.  The three-kernel path
// moves a P = 128 patch through HBM four times as a spectrum; this one writes it once as a real plane and reads it
// back once.  Per-pixel sums have a fixed order, so results are bit-stable and row slabs stitch bit-identically.
#pragma once
#include "rpsf_stream.cuh"

namespace rpsf {

template <int P, typename T> struct Small {
  using TL = Tile<P>;
  static constexpr int THREADS = 256;
  static constexpr int N1 = TL::N1, N2 = TL::N2, HALF = TL::HALF;
  static constexpr int SROW = HALF + 4;                              // padded row of the spectrum plane (bank spread)
  static constexpr int TEAMS = THREADS / N1;
  static constexpr int ROW_SCRATCH = TEAMS * TL::SCR;                // complex elements: per-team exchange / staging
  static constexpr int COL_SCRATCH = TL::SLOTS * P * TL::C;          // complex elements: column exchange
  static constexpr int SCRATCH = ROW_SCRATCH > COL_SCRATCH ? ROW_SCRATCH : COL_SCRATCH;
  static constexpr size_t SMEM = sizeof(cplx<T>) * (size_t)(P + P * SROW + SCRATCH) + sizeof(T) * P + 16;
  static constexpr bool OK = SMEM <= 110 * 1024 && TL::K2_THREADS <= THREADS && (HALF / TL::C) >= 1;
};

// One CTA = one (active patch, frame).  blockIdx.x = patch * batch + frame: the frames of a patch are neighbours,
// so its transfer kernel reaches HBM once per launch.
template <int P, typename T>
__global__ void __launch_bounds__(256, 2)
small_patch(const T* __restrict__ image, T* __restrict__ planes, const int2* __restrict__ corners,
            const int* __restrict__ active, const cplx<T>* __restrict__ kmain, const cplx<T>* __restrict__ knyq,
            const cplx<T>* __restrict__ tw_g, const T* __restrict__ win_g, ApplyGeom g, int batch, int bulk_ok) {
  using SM = Small<P, T>;
  using TL = Tile<P>;
  constexpr int N1 = TL::N1, N2 = TL::N2, HALF = TL::HALF, C = TL::C, NTILE = TL::NTILE, SROW = SM::SROW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);
  cplx<T>* S = tw + P;                                         // [P][SROW]
  cplx<T>* scratch = S + P * SROW;
  T* win = reinterpret_cast<T*>(scratch + SM::SCRATCH);
  uint64_t* bar = reinterpret_cast<uint64_t*>(win + P);
  for (int i = threadIdx.x; i < P; i += blockDim.x) { tw[i] = tw_g[i]; win[i] = win_g[i]; }
  const int a = blockIdx.x / batch, f = blockIdx.x % batch;
  const int2 corner = corners[a];
  const T* img = image + (long long)f * g.img_frame_stride;
  const bool direct = g.pad_mode == PAD_NONE;
  // ---------------------------------------------------------------- the patch arrives: one bulk copy (1-D TMA) per row
  // Row r of the patch lands at the start of row r of the spectrum plane (a real row is exactly as long as its packed
  // half-spectrum), so phase 1 transforms in place.  Rows whose in-frame span is 16-byte aligned are bulk copies of
  // that span; columns that hang over the frame edge, unaligned patches and `constant` rows are filled by the threads.
  constexpr unsigned RS = sizeof(T);
  const int x_lo = direct ? corner.y : max(corner.y, 0);
  const int x_hi = direct ? corner.y + P : min(corner.y + P, g.W);
  const bool span_ok = bulk_ok && x_hi > x_lo && ((x_lo * (int)RS) & 15) == 0 && ((x_hi * (int)RS) & 15) == 0 &&
                       (((x_lo - corner.y) * (int)RS) & 15) == 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  {
    // rows that take a bulk copy: in-frame source row (or a materialised pad); thread r looks at row r (P <= THREADS)
    static_assert(P <= SM::THREADS, "one thread per patch row");
    const bool my_row_bulk = span_ok && (int)threadIdx.x < P &&
                             (direct || pad_index(corner.x + (int)threadIdx.x, g.H, g.pad_mode) >= 0);
    const int rows = __syncthreads_count(my_row_bulk ? 1 : 0);
    if (threadIdx.x == 0) mbar_expect_tx(bar, (unsigned)rows * (unsigned)(x_hi - x_lo) * RS);
    __syncthreads();
    for (int r = threadIdx.x; r < P; r += blockDim.x) {
      const int y = pad_index(corner.x + r, g.H, g.pad_mode);
      T* dst = reinterpret_cast<T*>(S + r * SROW);
      if (span_ok && (direct || y >= 0)) {
        bulk_load(dst + (x_lo - corner.y), img + (long long)(y - g.img_row0) * g.img_pitch + x_lo, (unsigned)(x_hi - x_lo) * RS, bar);
      }
    }
    // everything a bulk copy does not bring: whole rows (unaligned / constant rows) or the overhanging columns
    const int lo_c = span_ok ? x_lo - corner.y : 0, hi_c = span_ok ? x_hi - corner.y : 0;
    const bool partial = !span_ok || lo_c > 0 || hi_c < P || (!direct && g.pad_mode == PAD_CONSTANT);
    mbar_wait(bar, 0);                                         // (an expect of 0 bytes completes at once)
    if (partial) {
      // only what the copies did not bring: the columns outside [lo_c, hi_c) of bulk rows (mirrored columns are read
      // back from the row itself when their source lies in the copied span), and whole rows that had no copy
      __syncthreads();                                         // every thread sees the landed rows
      const int n_fill = lo_c + (P - hi_c);
#pragma unroll 4
      for (int i = threadIdx.x; i < P * n_fill; i += blockDim.x) {
        const int r = i / n_fill, ci = i % n_fill;
        const int cidx = ci < lo_c ? ci : ci - lo_c + hi_c;
        const int y = pad_index(corner.x + r, g.H, g.pad_mode);
        if (!(direct || y >= 0)) continue;                     // a row without a copy: filled whole below
        const int x = pad_index(corner.y + cidx, g.W, g.pad_mode);
        const int cs = x - corner.y;
        T* row = reinterpret_cast<T*>(S + r * SROW);
        T val = T(0);
        if (!direct && x >= 0 && cs >= lo_c && cs < hi_c) val = row[cs];
        else if (direct || x >= 0) val = img[(long long)(y - g.img_row0) * g.img_pitch + x];
        row[cidx] = val;
      }
      if (!span_ok || (!direct && g.pad_mode == PAD_CONSTANT)) {
#pragma unroll 4
        for (int i = threadIdx.x; i < P * P; i += blockDim.x) {
          const int r = i / P, cidx = i % P;
          const int y = pad_index(corner.x + r, g.H, g.pad_mode);
          if (span_ok && (direct || y >= 0)) continue;         // this row came by bulk copy
          const int x = pad_index(corner.y + cidx, g.W, g.pad_mode);
          T val = T(0);
          if (direct || (x >= 0 && y >= 0)) val = img[(long long)(y - g.img_row0) * g.img_pitch + x];
          reinterpret_cast<T*>(S + r * SROW)[cidx] = val;
        }
      }
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 1: rows forward, in place
  {
    const int team = threadIdx.x / N1, t = threadIdx.x % N1;
    cplx<T>* scr = scratch + team * TL::SCR;
    const unsigned mask = team_mask(N1);
    auto ex = [](int k2, int n1) { return k2 * TL::EX_STRIDE + n1; };
    auto sync = [mask]() { __syncwarp(mask); };
    for (int pair = team; pair < HALF; pair += SM::TEAMS) {
      const int ra = 2 * pair, rb = ra + 1;
      const T* rowa = reinterpret_cast<const T*>(S + ra * SROW);
      const T* rowb = reinterpret_cast<const T*>(S + rb * SROW);
      cplx<T> v[N2];
      static_for<0, N2>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const int n = t + N1 * j;
        v[j] = cscale(mk<T>(rowa[n], rowb[n]), win[n]);
      });
      sync();                                                // the team has its two rows in registers: they may be overwritten
      coop_fft_forward<P, T>(v, t, scr, tw, ex, sync);
      static_for<0, N2>([&](auto ee) {                       // natural-order staging: Z[k], k = (t + N1*m) + N2*k1
        constexpr int e = decltype(ee)::value;
        scr[(t + N1 * (e / N1)) + N2 * (e % N1)] = v[e];
      });
      sync();
      // A[k] = (Z[k] + conj Z[P-k]) / 2,  B[k] = (Z[k] - conj Z[P-k]) / (2i), scaled by the row windows; bin 0 packs (DC, Nyquist)
      const T wa = T(0.5) * win[ra], wb = T(0.5) * win[rb];
      cplx<T>* outa = S + ra * SROW;
      cplx<T>* outb = S + rb * SROW;
#pragma unroll
      for (int i = 0; i < HALF / N1; ++i) {
        const int k = t + N1 * i;
        const cplx<T> z1 = scr[k];
        const cplx<T> z2 = scr[(P - k) & (P - 1)];
        const cplx<T> D = padd(z1, mk<T>(-z2.x, z2.y));
        cplx<T> A = cscale(padd(z1, mk<T>(z2.x, -z2.y)), wa);
        cplx<T> B = cscale(mk<T>(D.y, -D.x), wb);
        if (k == 0) {
          const cplx<T> zn = scr[HALF];
          A = mk<T>(T(2) * wa * z1.x, T(2) * wa * zn.x);
          B = mk<T>(T(2) * wb * z1.y, T(2) * wb * zn.y);
        }
        outa[k] = A;
        outb[k] = B;
      }
      sync();                                                // the team's scratch goes to its next row pair
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 2: columns, in place on S
  {
    const int c = threadIdx.x % C;
    const int n1 = (threadIdx.x / C) % N1;
    const int slot = threadIdx.x / (C * N1);
    const bool lane_ok = threadIdx.x < TL::K2_THREADS;       // fewer column threads than the CTA has (small P)
    const int gp = active[a];
    auto ex = [=](int k2, int nn) { return ((slot * N2 + k2) * N1 + nn) * C + c; };
    auto sync = []() { __syncthreads(); };
    for (int t0 = 0; t0 < NTILE; t0 += TL::SLOTS) {
      const int tile = t0 + slot;
      const bool valid = lane_ok && tile < NTILE;
      const bool special = valid && tile == 0 && c == 0;
      const cplx<T>* kp = kmain + (((long long)gp * NTILE + (valid ? tile : 0)) * N2) * (N1 * C) + n1 * C + c;
      cplx<T> kv[N2];
      static_for<0, N2>([&](auto ee) {
        constexpr int e = decltype(ee)::value;
        kv[e] = valid ? kp[(long long)e * (N1 * C)] : mk<T>(T(0), T(0));
      });
      cplx<T>* col = S + tile * C + c;
      cplx<T> v[N2];
      static_for<0, N2>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        v[j] = valid ? col[(n1 + N1 * j) * SROW] : mk<T>(T(0), T(0));
      });
      coop_fft_forward<P, T>(v, lane_ok ? n1 : 0, scratch, tw, ex, sync);
      if (t0 == 0) {                                         // CTA-uniform: this pass holds tile 0 with the packed column
        cplx<T>* zs = scratch + slot * P;
        if (special) {
          static_for<0, N2>([&](auto ee) {
            constexpr int e = decltype(ee)::value;
            zs[(n1 + N1 * (e / N1)) + N2 * (e % N1)] = v[e];
          });
        }
        __syncthreads();
        if (special) {
          const cplx<T>* kn = knyq + (long long)gp * P + n1;
          static_for<0, N2>([&](auto ee) {
            constexpr int e = decltype(ee)::value;
            const int k = (n1 + N1 * (e / N1)) + N2 * (e % N1);
            const cplx<T> zr = zs[(P - k) & (P - 1)];
            const cplx<T> zm = mk<T>(zr.x, -zr.y);
            const cplx<T> sum = mk<T>(T(0.5) * (v[e].x + zm.x), T(0.5) * (v[e].y + zm.y));
            const cplx<T> dif = mk<T>(T(0.5) * (v[e].x - zm.x), T(0.5) * (v[e].y - zm.y));
            v[e] = cadd(cmul(sum, kv[e]), cmul(dif, kn[e * N1]));
          });
        }
        __syncthreads();
      }
      if (!special) {
        static_for<0, N2>([&](auto ee) { v[decltype(ee)::value] = cmul(v[decltype(ee)::value], kv[decltype(ee)::value]); });
      }
      coop_fft_inverse<P, T>(v, lane_ok ? n1 : 0, scratch, tw, ex, sync);
      if (valid) {
        static_for<0, N2>([&](auto jj) {
          constexpr int j = decltype(jj)::value;
          col[(n1 + N1 * j) * SROW] = v[j];
        });
      }
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 3: rows inverse, windows, the patch plane
  {
    const int team = threadIdx.x / N1, t = threadIdx.x % N1;
    cplx<T>* scr = scratch + team * TL::SCR;
    const unsigned mask = team_mask(N1);
    auto ex = [](int k2, int n1) { return k2 * TL::EX_STRIDE + n1; };
    auto sync = [mask]() { __syncwarp(mask); };
    T* plane = planes + ((long long)f * g.n_active + a) * P * P;
    for (int pair = team; pair < HALF; pair += SM::TEAMS) {
      const int ra = 2 * pair, rb = ra + 1;
      const cplx<T>* ua = S + ra * SROW;
      const cplx<T>* ub = S + rb * SROW;
      cplx<T> v[N2];
      static_for<0, N2>([&](auto ee) {                       // Z[k] = Ua[k] + i*Ub[k] for k <= P/2, Hermitian mirror above
        constexpr int e = decltype(ee)::value;
        const int k = (t + N1 * (e / N1)) + N2 * (e % N1);
        const int src = k <= HALF ? k : P - k;
        const cplx<T> pa = ua[src == HALF ? 0 : src];
        const cplx<T> pb = ub[src == HALF ? 0 : src];
        cplx<T> z;
        if (k == 0)           z = mk<T>(pa.x, pb.x);
        else if (k == HALF)   z = mk<T>(pa.y, pb.y);
        else if (k < HALF)    z = mk<T>(pa.x - pb.y, pa.y + pb.x);
        else                  z = mk<T>(pa.x + pb.y, pb.x - pa.y);
        v[e] = z;
      });
      coop_fft_inverse<P, T>(v, t, scr, tw, ex, sync);
      const T wa = win[ra], wb = win[rb];
      T* oa = plane + (long long)ra * P;
      T* ob = oa + P;
      static_for<0, N2>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const int n = t + N1 * j;
        const T w = win[n];
        oa[n] = v[j].x * w * wa;
        ob[n] = v[j].y * w * wb;
      });
    }
  }
}

// Overlap-add of the patch planes: one CTA per output tile of TS x TS pixels; `tile_patches[tile][k]` are the active
// patches that cover the tile, in colour (= list) order, -1 padded.  Every output pixel is written once.
struct SmallTile { int y0, x0; };

template <int P, typename T>
__global__ void __launch_bounds__(256)
small_overlap_add(const T* __restrict__ planes, T* __restrict__ out, const SmallTile* __restrict__ tiles,
                  const int* __restrict__ tile_patches, int max_cover, const int2* __restrict__ corners, int tile_size,
                  ApplyGeom g) {
  const SmallTile tl = tiles[blockIdx.x];
  const int f = blockIdx.y;
  const int* cover = tile_patches + (long long)blockIdx.x * max_cover;
  const T* fplanes = planes + (long long)f * g.n_active * P * P;
  T* frame = out + (long long)f * g.out_frame_stride;
  const int y_end = min(min(tl.y0 + tile_size, g.row_end), g.H), x_end = min(tl.x0 + tile_size, g.W);
  const int y_begin = max(tl.y0, g.row_begin);
  const int width = x_end - tl.x0, rows = y_end - y_begin;
  if (width <= 0 || rows <= 0) return;
  // the (at most 4 for a covering) planes of this tile, as pointers to their pixel (y_begin, x0)
  constexpr int MAXC = 8;
  const T* src[MAXC];
  int n_src = 0;
  for (int k = 0; k < max_cover && k < MAXC; ++k) {
    const int a = cover[k];
    if (a < 0) break;
    const int2 c = corners[a];
    src[n_src++] = fplanes + ((long long)a * P + (y_begin - c.x)) * P + (tl.x0 - c.y);
  }
  T* dst = frame + (long long)(y_begin - g.out_row0) * g.out_pitch + tl.x0;
  constexpr int V = 16 / (int)sizeof(T);
  const bool vec = n_src > 0 && max_cover <= MAXC && width % V == 0 && (g.out_pitch % V) == 0 &&
                   (reinterpret_cast<uintptr_t>(dst) & 15) == 0;     // plane rows are P reals: 16-byte aligned at x0 - c.y (multiples of P/2)
  if (vec) {
    const int per_row = width / V;
    for (int i = threadIdx.x; i < rows * per_row; i += blockDim.x) {
      const int r = i / per_row, q = (i % per_row) * V;
      T acc[V];
#pragma unroll
      for (int k = 0; k < MAXC; ++k) {
        if (k < n_src) {
          T v[V];
          if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(src[k] + (long long)r * P + q);
          else *reinterpret_cast<double2*>(v) = *reinterpret_cast<const double2*>(src[k] + (long long)r * P + q);
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] = k == 0 ? v[e] : acc[e] + v[e];      // list order; the first term is stored
        }
      }
      if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(dst + (long long)r * g.out_pitch + q) = *reinterpret_cast<float4*>(acc);
      else *reinterpret_cast<double2*>(dst + (long long)r * g.out_pitch + q) = *reinterpret_cast<double2*>(acc);
    }
    return;
  }
  for (int i = threadIdx.x; i < rows * width; i += blockDim.x) {
    const int r = i / width, q = i % width;
    T acc = T(0);
    for (int k = 0; k < max_cover; ++k) {
      const int a = cover[k];
      if (a < 0) break;
      const int2 c = corners[a];
      const T v = fplanes[((long long)a * P + (y_begin + r - c.x)) * P + (tl.x0 + q - c.y)];
      acc = k == 0 ? v : acc + v;
    }
    dst[(long long)r * g.out_pitch + q] = acc;
  }
}

}  // namespace rpsf
