// rpsf_stream.cuh — persistent, bulk-async (TMA 1-D) versions of the two row kernels.
//
// K1 (transform.py:141-163) and K3 (transform.py:164-177) each touch every row of every patch
// exactly once, so the kernels are streaming passes whose only enemies are memory latency and
// issue slots spent on address arithmetic.  Here a warp is a self-contained pipeline:
//
//   * one elected lane per row issues `cp.async.bulk.shared::cluster.global` copies (UBLKCP, the
//     1-D TMA path) that land whole patch rows in a ring of per-warp shared-memory stages and
//     signal an mbarrier with the byte count — no register staging, no per-element addresses;
//   * the warp's teams (N1 lanes each, see rpsf_fft.cuh) wait on the mbarrier of the oldest
//     stage, pull their samples out of shared memory with constant offsets, and reuse the very
//     same stage as the FFT exchange buffer;
//   * the grid is persistent (one CTA per SM), warps walk the item list with a fixed stride, and
//     the copy for item i+STAGES-1 is in flight while item i is transformed.
//
// Items the bulk path cannot serve (patches that hang over the frame edge in x, unaligned
// corners, `constant` padding rows) are gathered through pad_index() by the warp itself into
// the same stage and then take the same compute path.
#pragma once
#include <cstdint>
#include "rpsf_kernels.cuh"

namespace rpsf {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  unsigned done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// orders this thread's earlier generic-proxy shared-memory accesses before later async-proxy ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef RPSF_STREAM_SMEM_KB
#define RPSF_STREAM_SMEM_KB 224
#endif
constexpr int STREAM_SMEM_BUDGET = RPSF_STREAM_SMEM_KB * 1024;
#ifndef RPSF_STREAM_MAX_WARPS
#define RPSF_STREAM_MAX_WARPS 16
#endif
#ifndef RPSF_STREAM_STAGES
#define RPSF_STREAM_STAGES 3
#endif

template <int P, typename T, int STAGES_REQ = RPSF_STREAM_STAGES> struct Stream {
  static constexpr int N1 = Split<P>::N1, N2 = Split<P>::N2, HALF = P / 2;
  static constexpr int TPW = 32 / N1;                                      // teams per warp
  static constexpr int ROWS = 2 * TPW;                                     // patch rows per warp item
  static constexpr int IPP = HALF / TPW;                                   // warp items per patch
  // A team's slot holds its row pair and, later, its padded N2 x (N1+1) exchange matrix (constant offsets for
  // every access, conflict-free).  A skew of N1 reals per team puts the sample reads of the teams of a warp on
  // disjoint banks — the row kernels are bound by the shared-memory data pipe (128 B/cycle/SM), so every
  // avoidable wavefront counts.  (This is also why rows arrive as 1-D bulk copies and not as one tiled
  // tensor-map box per item: a box lands densely, 128-byte aligned, and the 2-way conflict that costs was
  // measured to outweigh the cheaper issue.)
  static constexpr int EX_STRIDE = N1 + 1;
  static constexpr int SLOT_ELEMS = N2 * EX_STRIDE > P ? N2 * EX_STRIDE : P;            // complex elements
  static constexpr int TEAM_BYTES = SLOT_ELEMS * (int)sizeof(cplx<T>) + N1 * (int)sizeof(T);
  static constexpr int STAGE_BYTES = TPW * TEAM_BYTES;
  static constexpr int TABLE_BYTES = P * (int)sizeof(cplx<T>) + P * (int)sizeof(T);
  static constexpr int RING_OFFSET = (TABLE_BYTES + 1024 + 127) / 128 * 128;   // tables, mbarriers, then the ring (128-byte aligned)
  static constexpr int AVAIL = STREAM_SMEM_BUDGET - RING_OFFSET;
  static constexpr int WS = AVAIL / (STAGES_REQ * STAGE_BYTES);
  static constexpr int STAGES = WS >= 8 ? STAGES_REQ : 2;
  static constexpr int WFIT = AVAIL / (STAGES * STAGE_BYTES);
  static constexpr int WARPS = WFIT > RPSF_STREAM_MAX_WARPS ? RPSF_STREAM_MAX_WARPS : WFIT;
  static constexpr int THREADS = WARPS * 32;
  static constexpr size_t SMEM = RING_OFFSET + (size_t)WARPS * STAGES * STAGE_BYTES;
  static_assert(WARPS >= 2, "stage ring does not fit shared memory");
  static_assert(WARPS * STAGES * 8 <= 768, "mbarriers overlap the words behind them (rpsf_fused.cuh keeps a watermark at +768)");
  static_assert(TEAM_BYTES % 16 == 0, "team slots must keep 16-byte alignment for bulk copies");

  __device__ static __forceinline__ int ex(int k2, int n1) { return k2 * EX_STRIDE + n1; }
};

// The gather kernel's own ring.  At 128 px three stages leave room for 15 warps (4 + 4 + 4 + 3 on the schedulers); its
// items are short enough that two stages and 16 warps are faster (single 1024^2 frame: 17.0 -> 15.1 us, 8 frames: 7.0
// -> 6.8 us per frame), while the overlap-add kernel keeps three (6.2 vs 6.5 us).
template <int P, typename T>
using StreamK1 = Stream<P, T, (P == 128 && sizeof(T) == 4) ? 2 : RPSF_STREAM_STAGES>;

#ifndef RPSF_K1_EVICT_LAST   // 1: K1's spectrum stores carry an L2 evict_last hint (K2 reads them next)
#define RPSF_K1_EVICT_LAST 1
#endif
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_spec(float2* p, float2 v, unsigned long long pol) {
#if RPSF_K1_EVICT_LAST
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
#else
  (void)pol; *p = v;
#endif
}
__device__ __forceinline__ void st_spec(double2* p, double2 v, unsigned long long) { *p = v; }

// ============================================================================ K1, streaming
// gather + apodize + row FFT (same arithmetic as k1_gather_window_rowfft).  Warp item = ROWS
// consecutive rows of one patch of one frame; team tm of the warp owns rows (2*tm, 2*tm+1) of it.
// Where a patch's spectrum lives and who is told when it is complete.  The stand-alone kernels use the plain
// policies: patch `a` of frame `f` sits at slot f * n_active + a of the workspace and nobody is signalled.  The
// fused pipeline (rpsf_fused.cuh) substitutes a ring of L2-resident slots with producer / consumer counters.
struct PlainK1 {
  // workspace slot (in patches) of patch `a` of frame `f`; may block until the slot may be overwritten
  __device__ __forceinline__ unsigned slot(int f, int a, int n_active) { return (unsigned)(f * n_active + a); }
  // Warp item number `item` = (f * n_active + a) * IPP + q, which this warp stored one iteration ago, is complete
  // (kNone: there is none).  Items are dealt round robin to the warps of the grid, so the WARPS items
  // [WARPS * k, WARPS * (k + 1)) always belong to one CTA, and (WARPS dividing IPP) to one patch.  Called between the transform and the stores of the NEXT item, so that a policy that
  // publishes with a memory fence finds those older stores already performed instead of stalling on fresh ones.
  static constexpr unsigned kNone = 0xffffffffu;
  __device__ __forceinline__ void publish(unsigned /*item*/, int /*lane*/) {}
  // the head of the pipeline moved to patch `a`: its slot() follows one iteration later
  __device__ __forceinline__ void prefetch(int /*a*/) {}
};

template <int P, typename T, typename ST = Stream<P, T>, typename Pol>
__device__ __forceinline__ void
k1_stream_body(const T* __restrict__ image, cplx<T>* __restrict__ spec, const int2* __restrict__ corners,
               const cplx<T>* __restrict__ tw_g, const T* __restrict__ win_g, const ApplyGeom& g, int batch, int bulk_ok,
               unsigned cta, unsigned n_cta, Pol pol, unsigned char* smem_raw) {
  constexpr int N1 = ST::N1, N2 = ST::N2, HALF = ST::HALF, ROWS = ST::ROWS, IPP = ST::IPP;
  constexpr int STAGES = ST::STAGES, WARPS = ST::WARPS;
  constexpr unsigned ROW_BYTES = P * sizeof(T);
  cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);
  T* win = reinterpret_cast<T*>(tw + P);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + ST::TABLE_BYTES);            // [WARPS][STAGES]
  unsigned char* ring = smem_raw + ST::RING_OFFSET;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tm = lane / N1, t = lane % N1;
  for (int i = threadIdx.x; i < P; i += blockDim.x) { tw[i] = tw_g[i]; win[i] = win_g[i]; }
  uint64_t* bar = bars + warp * STAGES;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(bar + s, 1);
    mbar_fence_init();
  }
  __syncthreads();
  grid_dependency_wait();         // tables and barriers are ready; frames / spectra may come from the kernel before
  grid_launch_dependents();

  unsigned char* my_ring = ring + (size_t)warp * STAGES * ST::STAGE_BYTES;
  const unsigned long long l2_keep = RPSF_K1_EVICT_LAST ? l2_policy_evict_last() : 0ull;
  const unsigned n_items = (unsigned)batch * (unsigned)g.n_active * IPP;       // host guarantees < 2^30
  const unsigned stride = n_cta * WARPS;
  const unsigned first = cta * WARPS + warp;
  const bool direct = g.pad_mode == PAD_NONE;

  // How an item's rows reach its stage:
  //   BULK    every column lies inside the frame: one bulk copy per row
  //   PARTIAL the patch hangs over the left/right frame edge: bulk copy of the in-frame span, then the
  //           overhanging columns are filled from their pad_index() sources (mirrored columns are
  //           read back from the stage itself when the source lies in the copied span)
  //   MANUAL  unaligned corners, `constant` rows above/below the frame: the warp gathers every sample
  enum : unsigned { BULK = 0, PARTIAL = 1, MANUAL = 2 };

  // ---- the head of the pipeline: the next item to be issued, decoded incrementally ----------
  // item = (f * n_active + a) * IPP + q; one step adds `stride` = (df, da, dq) with carries
  unsigned head = first;
  int hq = int(first % IPP), ha = int((first / IPP) % (unsigned)g.n_active), hf = int((first / IPP) / (unsigned)g.n_active);
  const int dq = int(stride % IPP), da = int((stride / IPP) % (unsigned)g.n_active),
            df = int((stride / IPP) / (unsigned)g.n_active);
  int2 hcorner = make_int2(0, 0);
  if (head < n_items) hcorner = __ldg(corners + ha);
  auto advance_head = [&]() {
    head += stride;
    hq += dq; ha += da; hf += df;
    if (hq >= IPP) { hq -= IPP; ++ha; }
    if (ha >= g.n_active) { ha -= g.n_active; ++hf; }
    if (head < n_items) { hcorner = __ldg(corners + ha); pol.prefetch(ha); }   // consumed one iteration later
  };
  // first row of the head item in the workspace's (frame, patch, row) order, and its patch row
  auto src_row = [&](int corner_row, int patch_row) { return pad_index(corner_row + patch_row, g.H, g.pad_mode); };

  unsigned phase = 0;      // bit s: parity the next wait on stage s must see

  // issue the head item into stage `st`; returns (workspace row index << 2) | kind
  auto issue_head = [&](int st) -> unsigned {
    const unsigned rowidx = pol.slot(hf, ha, g.n_active) * P + hq * ROWS;
    unsigned char* stage = my_ring + st * ST::STAGE_BYTES;
    const int x_lo = direct ? hcorner.y : max(hcorner.y, 0);
    const int x_hi = direct ? hcorner.y + P : min(hcorner.y + P, g.W);
    // source span AND its landing offset inside the stage row must keep the 16-byte granularity of bulk copies
    // (a patch overhanging both frame edges can have aligned ends but a misaligned offset x_lo - corner.y)
    const bool aligned = bulk_ok && x_hi > x_lo && ((x_lo * (int)sizeof(T)) & 15) == 0 && ((x_hi * (int)sizeof(T)) & 15) == 0 &&
                         (((x_lo - hcorner.y) * (int)sizeof(T)) & 15) == 0;
    int y = 0;
    if (lane < ROWS) y = src_row(hcorner.x, hq * ROWS + lane);
    const bool rows_ok = __all_sync(0xffffffffu, direct || y >= 0);
    const unsigned kind = !(aligned && rows_ok) ? MANUAL : (x_hi - x_lo == P ? BULK : PARTIAL);
    if (kind != MANUAL) {
      const unsigned bytes = (unsigned)(x_hi - x_lo) * (unsigned)sizeof(T);
      // The stage was last written through the generic proxy (exchange matrix, PARTIAL / MANUAL fills); the
      // closing __syncwarp of that iteration ordered those stores before lane 0, whose proxy fence now orders
      // them before the async-proxy writes of the refill.
      if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(bar + st, ROWS * bytes);
      }
      __syncwarp();
      if (lane < ROWS) {
        const T* src = image + (long long)hf * g.img_frame_stride + (long long)(y - g.img_row0) * g.img_pitch + x_lo;
        bulk_load(stage + (lane >> 1) * ST::TEAM_BYTES + (lane & 1) * ROW_BYTES + (x_lo - hcorner.y) * (int)sizeof(T), src, bytes,
                  bar + st);
      }
    }
    return (rowidx << 2) | kind;
  };

  unsigned pend[STAGES - 1], pend_item[STAGES - 1];    // (workspace row << 2) | kind, and the item's number
#pragma unroll
  for (int k = 0; k < STAGES - 1; ++k) {
    pend[k] = 0; pend_item[k] = 0;
    if (head < n_items) { pend_item[k] = head; pend[k] = issue_head(k); advance_head(); }
  }

  // per-thread constants: offset inside a stage, column window of this thread's samples
  const unsigned team_off = tm * ST::TEAM_BYTES;
  T wcol[N2];
  static_for<0, N2>([&](auto jj) { wcol[decltype(jj)::value] = win[t + N1 * decltype(jj)::value]; });

  int s = 0;
  unsigned prev_item = PlainK1::kNone;
  for (unsigned it = first; it < n_items; it += stride) {
    const unsigned cur = pend[0], item = pend_item[0], pf = item / IPP;
#pragma unroll
    for (int k = 0; k + 1 < STAGES - 1; ++k) { pend[k] = pend[k + 1]; pend_item[k] = pend_item[k + 1]; }
    // the stage the previous iteration released receives the head item
    if (head < n_items) {
      pend_item[STAGES - 2] = head;
      pend[STAGES - 2] = issue_head(s == 0 ? STAGES - 1 : s - 1);
      advance_head();
    }
    const unsigned kind = cur & 3u, rowidx = cur >> 2;
    unsigned char* stage = my_ring + s * ST::STAGE_BYTES;
    if (kind != MANUAL) {
      mbar_wait(bar + s, (phase >> s) & 1u);
      phase ^= 1u << s;
    }
    if (kind >= PARTIAL) {
      // rare path: recover the item's coordinates from its patch-frame and workspace row index
      const int q_rows = int(rowidx & (P - 1));
      const int a = int(pf % (unsigned)g.n_active), f = int(pf / (unsigned)g.n_active);
      const int2 corner = __ldg(corners + a);
      const T* img = image + (long long)f * g.img_frame_stride;
      const int lo_c = kind == PARTIAL ? max(-corner.y, 0) : 0;                   // bulk-copied columns [lo_c, hi_c)
      const int hi_c = kind == PARTIAL ? min(g.W - corner.y, P) : 0;
      const int n_fill = lo_c + (P - hi_c);                                       // columns outside [lo_c, hi_c)
#pragma unroll 1
      for (int ci = lane; ci < n_fill; ci += 32) {
        const int c = ci < lo_c ? ci : ci - lo_c + hi_c;
        const int x = pad_index(corner.y + c, g.W, g.pad_mode);
        const int cs = x - corner.y;                                              // source column inside the stage?
        const bool in_stage = !direct && x >= 0 && cs >= lo_c && cs < hi_c;
#pragma unroll 1
        for (int r = 0; r < ROWS; ++r) {
          T* dst = reinterpret_cast<T*>(stage + (r >> 1) * ST::TEAM_BYTES + (r & 1) * ROW_BYTES);
          T val = T(0);
          if (in_stage) {
            val = dst[cs];
          } else {
            const int y = src_row(corner.x, q_rows + r);
            if (direct || (x >= 0 && y >= 0)) val = img[(long long)(y - g.img_row0) * g.img_pitch + x];
          }
          dst[c] = val;
        }
      }
      __syncwarp();
    }

    const T* ra_p = reinterpret_cast<const T*>(stage + team_off);
    const T* rb_p = ra_p + P;
    cplx<T>* scr = reinterpret_cast<cplx<T>*>(stage + team_off);
    cplx<T> v[N2];
    static_for<0, N2>([&](auto jj) {
      constexpr int j = decltype(jj)::value;
      const int n = t + N1 * j;
      v[j] = cscale(mk<T>(ra_p[n], rb_p[n]), wcol[j]);
    });
    __syncwarp();                                            // samples are in registers: the slot becomes the exchange buffer
    auto sync = []() { __syncwarp(); };
    coop_fft_forward<P, T>(v, t, scr, tw, [](int k2, int n1) { return ST::ex(k2, n1); }, sync);
    pol.publish(prev_item, lane);                            // the previous item's rows (team barriers above order them)
    prev_item = item;

    // This thread holds Z[k] for k = (t + N1*m) + N2*k1 in v[m*N1 + k1].  Split into the two rows' Hermitian
    // half-spectra: A[k] = (Z[k] + conj Z[P-k]) / 2, B[k] = (Z[k] - conj Z[P-k]) / (2i), scaled by the row
    // window.  Z[P-k] for k = t + N1*i sits in lane N1-t of the team at a compile-time register, so it comes by
    // shuffle — no second pass through shared memory (the data pipe is this kernel's bound).
    const int ra = int(rowidx & (P - 1)) + 2 * tm;
    const T wa = T(0.5) * win[ra], wb = T(0.5) * win[ra + 1];
    cplx<T>* outa = spec + ((size_t)rowidx + 2 * tm) * HALF + t;
    constexpr int R = N2 / N1;
    auto reg_of = [](int i) constexpr { return (i % R) * N1 + i / R; };   // register holding bin (lane) + N1*i
    const int partner = (N1 - t) & (N1 - 1);
    static_for<0, HALF / N1>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      const cplx<T> z1 = v[reg_of(i)];
      const cplx<T> mine = v[reg_of((N2 - i) % N2)];                     // lane 0: bin P - N1*i is its own
      const cplx<T> theirs = v[reg_of(N2 - 1 - i)];                      // lane t > 0: bin P - t - N1*i = (N1-t) + N1*(N2-1-i)
      cplx<T> z2;
      z2.x = __shfl_sync(0xffffffffu, theirs.x, partner, N1);
      z2.y = __shfl_sync(0xffffffffu, theirs.y, partner, N1);
      if (t == 0) z2 = mine;
      const cplx<T> D = padd(z1, mk<T>(-z2.x, z2.y));
      cplx<T> A = cscale(padd(z1, mk<T>(z2.x, -z2.y)), wa);
      cplx<T> B = cscale(mk<T>(D.y, -D.x), wb);
      if constexpr (i == 0) {
        if (t == 0) {                                        // pack (DC, Nyquist): both real
          const cplx<T> zn = v[reg_of(N2 / 2)];
          A = mk<T>(T(2) * wa * z1.x, T(2) * wa * zn.x);
          B = mk<T>(T(2) * wb * z1.y, T(2) * wb * zn.y);
        }
      }
      st_spec(outa + N1 * i, A, l2_keep);
      st_spec(outa + HALF + N1 * i, B, l2_keep);
    });
    // coop_fft_forward ended with a team barrier after its last exchange read: the stage may be refilled
    s = s + 1 == STAGES ? 0 : s + 1;
  }
  __syncwarp();
  pol.publish(prev_item, lane);                              // the last item
}

template <int P, typename T>
__global__ void __launch_bounds__(StreamK1<P, T>::THREADS, 1)
k1_stream(const T* __restrict__ image, cplx<T>* __restrict__ spec, const int2* __restrict__ corners,
          const cplx<T>* __restrict__ tw_g, const T* __restrict__ win_g, ApplyGeom g, int batch, int bulk_ok) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  k1_stream_body<P, T, StreamK1<P, T>>(image, spec, corners, tw_g, win_g, g, batch, bulk_ok, blockIdx.x, gridDim.x, PlainK1{},
                                       smem_raw);
}

// ============================================================================ K3, streaming
// row IFFT + window + overlap-add (transform.py:164-177; same arithmetic as k3_rowpair_gather).
//
// For a covering the (patch, row pair) items that land on one pair of output rows form a chain of
// groups: group g holds the items whose patch corner column is cx0 + g*P/2 (one per overlapping
// patch row, summed in the frequency domain in colour order, one inverse FFT per group), and
// consecutive groups overlap by half a patch.  A team walks a chain left to right keeping the right
// half of the previous group's rows in registers:
//
//     out[cx_g .. cx_g + P/2) = right half of group g-1  +  left half of group g
//
// so every output pixel is produced exactly once, straight from registers — no shared-memory
// plane, no zero fill, no read-modify-write, and a fixed two-term sum per pixel (bit-stable; the
// same sum on every slab and every chunking).  Chains are cut into chunks for parallelism; a
// chunk that does not start its chain recomputes the group before it (flag SEAM) only for its
// right half.  Items arrive by bulk copy into the warp's stage ring exactly like K1's rows; the
// stage of a group's last item doubles as the exchange buffer of the group's inverse FFT.
struct StreamTask {
  int y;            // output rows y, y+1
  int cx0;          // corner column of the first group this task computes
  int item_begin;   // into the item-code list
  int n_steps;      // item codes to walk; 0 = idle team
  int flags;        // TASK_SEAM | TASK_LAST
  int pad0, pad1, pad2;
};
enum : int { TASK_SEAM = 1, TASK_LAST = 2 };
constexpr unsigned ITEM_LAST_OF_GROUP = 1u << 30;            // item code = (active*P/2 + pair) | flag

// Fused output gather for patch-row slabs across GPUs: besides `out`, every pixel is stored to the same
// position of up to 7 peer buffers (byte offsets from `out`; the peers' frames mapped over NVLink with
// CUDA IPC), so the band lands in every rank's full frame while the kernel runs and no all-gather follows.
struct OutMirrors {
  int n;
  long long delta[7];
};

// Plain policy of the stand-alone kernel: item code = active * P/2 + pair, patch `a` of frame `f` at workspace slot
// f * n_active + a, no waiting, no signalling.  (The fused pipeline's policy is in rpsf_fused.cuh.)
template <int P> struct PlainK3 {
  // row pair inside its patch (for the row windows)
  __device__ __forceinline__ int pair(unsigned code) const { return int((code & (ITEM_LAST_OF_GROUP - 1)) % (P / 2)); }
  // element offset of the item's two half-spectrum rows in the workspace
  __device__ __forceinline__ size_t offset(int f, int n_active, unsigned code, const StreamTask&) const {
    return ((size_t)f * n_active * (P / 2) + (code & (ITEM_LAST_OF_GROUP - 1))) * P;
  }
  // row windows of the item's two rows (transform.py:165)
  __device__ __forceinline__ void row_windows(const float* win, unsigned code, float& wa, float& wb) const { wa = win[2 * pair(code)]; wb = win[2 * pair(code) + 1]; }
  __device__ __forceinline__ void row_windows(const double* win, unsigned code, double& wa, double& wb) const { wa = win[2 * pair(code)]; wb = win[2 * pair(code) + 1]; }
  __device__ __forceinline__ void wait(int /*f*/, const StreamTask&, bool /*live*/) const {}
  __device__ __forceinline__ void done(int /*f*/, const StreamTask&, bool /*leader*/) const {}
};

// Policy for the paired workspace of k2_chain (rpsf_kernels.cuh): a group is ONE item — the two patches that overlap
// on a pair of output rows were summed, row windows included, by the column pass — so item code = band * P/4 + row
// pair inside the band, every item is the last of its group, and the row windows here are 1.
template <int P> struct PairedK3 {
  long long items_per_frame;          // bands * P / 4
  __device__ __forceinline__ int pair(unsigned code) const { return int((code & (ITEM_LAST_OF_GROUP - 1)) % (P / 4)); }
  __device__ __forceinline__ size_t offset(int f, int /*n_active*/, unsigned code, const StreamTask&) const {
    return ((size_t)f * items_per_frame + (code & (ITEM_LAST_OF_GROUP - 1))) * P;
  }
  template <typename T> __device__ __forceinline__ void row_windows(const T*, unsigned, T& wa, T& wb) const { wa = T(1); wb = T(1); }
  __device__ __forceinline__ void wait(int, const StreamTask&, bool) const {}
  __device__ __forceinline__ void done(int, const StreamTask&, bool) const {}
};

template <int P, typename T, bool MIRROR, typename Pol>
__device__ __forceinline__ void
k3_stream_body(const cplx<T>* __restrict__ spec, T* __restrict__ out, const StreamTask* __restrict__ tasks,
               const unsigned* __restrict__ codes, int n_warp_items, const cplx<T>* __restrict__ tw_g,
               const T* __restrict__ win_g, const ApplyGeom& g, int batch, const OutMirrors& mir, unsigned cta,
               unsigned n_cta, Pol pol, unsigned char* smem_raw) {
  auto put = [&](T* p, T val) {
    *p = val;
    if constexpr (MIRROR) {
      for (int d = 0; d < mir.n; ++d) *reinterpret_cast<T*>(reinterpret_cast<char*>(p) + mir.delta[d]) = val;
    }
  };
  using ST = Stream<P, T>;
  constexpr int N1 = ST::N1, N2 = ST::N2, HALF = ST::HALF, TPW = ST::TPW;
  constexpr int STAGES = ST::STAGES, WARPS = ST::WARPS;
  constexpr int HN = N2 / 2;                                 // registers per half row
  constexpr unsigned ITEM_BYTES = P * sizeof(cplx<T>);      // one row pair of half-spectra
  cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);
  T* win = reinterpret_cast<T*>(tw + P);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + ST::TABLE_BYTES);
  unsigned char* ring = smem_raw + ST::RING_OFFSET;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tm = lane / N1, t = lane % N1;
  for (int i = threadIdx.x; i < P; i += blockDim.x) { tw[i] = tw_g[i]; win[i] = win_g[i]; }
  uint64_t* bar = bars + warp * STAGES;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(bar + s, 1);
    mbar_fence_init();
  }
  __syncthreads();
  grid_dependency_wait();         // tables and barriers are ready; frames / spectra may come from the kernel before
  grid_launch_dependents();

  unsigned char* my_ring = ring + (size_t)warp * STAGES * ST::STAGE_BYTES;
  const unsigned team_off = tm * ST::TEAM_BYTES;
  T wcol[N2];                                                // column window of this thread's samples
  static_for<0, N2>([&](auto jj) { wcol[decltype(jj)::value] = win[t + N1 * decltype(jj)::value]; });

  const unsigned total = (unsigned)n_warp_items * (unsigned)batch;
  const unsigned stride = n_cta * WARPS;
  unsigned phase = 0;
  int s = 0;                                                 // stage of the next step to consume

  for (unsigned wi = cta * WARPS + warp; wi < total; wi += stride) {
    const int f = int(wi / (unsigned)n_warp_items);
    const unsigned local = wi - (unsigned)f * (unsigned)n_warp_items;
    const StreamTask task = tasks[(size_t)local * TPW + tm];
    const int K = __reduce_max_sync(0xffffffffu, task.n_steps);
    const bool live = task.n_steps > 0;
    const unsigned live_teams = __popc(__ballot_sync(0xffffffffu, live && t == 0));
    const unsigned* my_codes = codes + task.item_begin;
    pol.wait(f, task, live);                                 // fused pipeline: the patches this task reads are complete

    // issue step k into stage st: every live team's row pair of half-spectra, one bulk copy each
    auto issue = [&](unsigned code, int st) {
      if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(bar + st, live_teams * ITEM_BYTES);
      }
      __syncwarp();
      if (live && t == 0) {
        const cplx<T>* src = spec + pol.offset(f, g.n_active, code, task);
        bulk_load(my_ring + st * ST::STAGE_BYTES + team_off, src, ITEM_BYTES, bar + st);
      }
    };

    // prologue: STAGES-1 steps in flight, the code of the next one to issue in a register
    unsigned inflight[STAGES - 1];
#pragma unroll
    for (int k = 0; k < STAGES - 1; ++k) {
      inflight[k] = 0;
      if (k < K) {
        inflight[k] = live ? __ldg(my_codes + k) : 0u;
        int st = s + k; if (st >= STAGES) st -= STAGES;
        issue(inflight[k], st);
      }
    }
    unsigned next_code = (live && STAGES - 1 < K) ? __ldg(my_codes + STAGES - 1) : 0u;

    cplx<T> v[N2];
    static_for<0, N2>([&](auto ee) { v[decltype(ee)::value] = mk<T>(T(0), T(0)); });
    cplx<T> prev[HN];
    static_for<0, HN>([&](auto jj) { prev[decltype(jj)::value] = mk<T>(T(0), T(0)); });
    int cx = task.cx0;
    bool emit_left = !(task.flags & TASK_SEAM);
    const bool oka = live && task.y >= g.row_begin && task.y < g.row_end;
    const bool okb = live && task.y + 1 >= g.row_begin && task.y + 1 < g.row_end;
    T* oa = out + (size_t)f * g.out_frame_stride + (long long)(task.y - g.out_row0) * g.out_pitch + t;
    T* ob = oa + g.out_pitch;
    // 16-byte stores: row starts aligned in `out` (and hence in the mirrors, whose offsets are checked on the host)
    const bool vec_ok = MIRROR && ((reinterpret_cast<uintptr_t>(oa - t) | (uintptr_t)(g.out_pitch * sizeof(T))) & 15) == 0;

    for (int k = 0; k < K; ++k) {
      const unsigned code = inflight[0];
#pragma unroll
      for (int q = 0; q + 1 < STAGES - 1; ++q) inflight[q] = inflight[q + 1];
      if (k + STAGES - 1 < K) {
        inflight[STAGES - 2] = next_code;
        issue(next_code, s == 0 ? STAGES - 1 : s - 1);
        next_code = (live && k + STAGES < K) ? __ldg(my_codes + k + STAGES) : 0u;
      }
      mbar_wait(bar + s, (phase >> s) & 1u);
      phase ^= 1u << s;
      cplx<T>* slot = reinterpret_cast<cplx<T>*>(my_ring + s * ST::STAGE_BYTES + team_off);

      if (live) {
        // Z[k] += wa*Ua[k] + i*wb*Ub[k] for k <= P/2, Hermitian mirror above; bin 0 packs (DC, Nyquist)
        T wa_s, wb_s;
        pol.row_windows(win, code, wa_s, wb_s);
        const cplx<T> wa = mk<T>(wa_s, wa_s), wb = mk<T>(wb_s, wb_s);
        const cplx<T>* lo = slot + t;                      // Ua[bin] = lo[bin - t], Ub[bin] = lo[P/2 + bin - t]
        const cplx<T>* neg = slot - t;                     // Ua[P - bin] = neg[P - (bin - t)]
        const cplx<T> p0a = slot[0], p0b = slot[HALF];     // packed (DC, Nyquist) of the two rows
        static_for<0, N2>([&](auto ee) {
          constexpr int e = decltype(ee)::value;
          constexpr int m = e / N1, k1 = e % N1;
          constexpr int base = N1 * m + N2 * k1;            // bin = base + t; [base, base + N1) never straddles P/2
          cplx<T> a2, b2;
          if constexpr (base < HALF) {
            const cplx<T> pa = lo[base], pb = lo[HALF + base];
            a2 = pa; b2 = mk<T>(-pb.y, pb.x);
            if constexpr (base == 0) {
              if (t == 0) { a2 = mk<T>(p0a.x, T(0)); b2 = mk<T>(T(0), p0b.x); }
            }
          } else {
            // bin > P/2: conjugate mirror of bin P - (base + t).  (base == P/2, t == 0 is the Nyquist bin:
            // the generic read lands on a harmless in-slot word and is replaced below.)
            const cplx<T> pa = neg[P - base], pb = neg[HALF + P - base];
            a2 = mk<T>(pa.x, -pa.y); b2 = mk<T>(pb.y, pb.x);
            if constexpr (base == HALF) {
              if (t == 0) { a2 = mk<T>(p0a.y, T(0)); b2 = mk<T>(T(0), p0b.y); }
            }
          }
          v[e] = pfma(b2, wb, pfma(a2, wa, v[e]));
        });
      }
      __syncwarp();                                        // every lane has read the item

      if (__any_sync(0xffffffffu, (code & ITEM_LAST_OF_GROUP) != 0)) {
        auto sync = []() { __syncwarp(); };
        coop_fft_inverse<P, T>(v, t, slot, tw, [](int k2, int n1) { return ST::ex(k2, n1); }, sync, sync);
        // Column window: the right half is scaled and kept for the next group; the left half meets the
        // previous group's right half in ONE fused multiply-add, o = v * w + prev, written out as such so
        // that every instantiation of this kernel rounds the same way (left to the compiler, the packed
        // multiply and add were contracted in one instantiation and not in another).
        static_for<HN, N2>([&](auto jj) { v[decltype(jj)::value] = cscale(v[decltype(jj)::value], wcol[decltype(jj)::value]); });
        auto left = [&](auto jj) -> cplx<T> {
          constexpr int j = decltype(jj)::value;
          return pfma(v[j], mk<T>(wcol[j], wcol[j]), prev[j]);
        };
        bool wide = false;
        if constexpr (MIRROR && sizeof(T) == 4 && HN % 4 == 0) {
          // Mirrored stores travel over NVLink, where 64-byte pieces are expensive: transpose the two half
          // rows through the (now idle) exchange slot so that every lane holds 4 consecutive pixels, and
          // store 16 bytes per lane — 256 contiguous bytes per team and instruction — to `out` and every mirror.
          wide = emit_left && live && cx >= 0 && cx + HALF <= g.W && vec_ok;
          T* plane = reinterpret_cast<T*>(slot);
          if (wide) {
            static_for<0, HN>([&](auto jj) {
              constexpr int j = decltype(jj)::value;
              const cplx<T> o = left(jj);
              plane[t + N1 * j] = o.x;
              plane[HALF + t + N1 * j] = o.y;
            });
          }
          __syncwarp();
          if (wide) {
            static_for<0, HN / 4>([&](auto qq) {
              constexpr int q = decltype(qq)::value;
              const int i0 = q * 4 * N1 + 4 * t;
              const float4 pa = *reinterpret_cast<const float4*>(plane + i0);
              const float4 pb = *reinterpret_cast<const float4*>(plane + HALF + i0);
              float* da = reinterpret_cast<float*>(oa - t + cx + i0);       // oa / ob carry + t
              float* db = reinterpret_cast<float*>(ob - t + cx + i0);
              if (oka) {
                *reinterpret_cast<float4*>(da) = pa;
                for (int d = 0; d < mir.n; ++d) *reinterpret_cast<float4*>(reinterpret_cast<char*>(da) + mir.delta[d]) = pa;
              }
              if (okb) {
                *reinterpret_cast<float4*>(db) = pb;
                for (int d = 0; d < mir.n; ++d) *reinterpret_cast<float4*>(reinterpret_cast<char*>(db) + mir.delta[d]) = pb;
              }
            });
          }
          __syncwarp();                                     // the slot goes back to the ring after this step
        }
        if (emit_left && !wide) {
          const bool inside = cx >= 0 && cx + HALF <= g.W;
          static_for<0, HN>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            const cplx<T> o = left(jj);
            const int x = cx + N1 * j;                      // + t folded into oa / ob
            if (inside || (x + t >= 0 && x + t < g.W)) {
              if (oka) put(oa + x, o.x);
              if (okb) put(ob + x, o.y);
            }
          });
        }
        static_for<0, HN>([&](auto jj) { prev[decltype(jj)::value] = v[HN + decltype(jj)::value]; });
        if (k == K - 1 && (task.flags & TASK_LAST)) {
          const int cr = cx + HALF;                         // the chain ends: its right half stands alone
          static_for<0, HN>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            const int x = cr + N1 * j;
            if (x + t >= 0 && x + t < g.W) {
              if (oka) put(oa + x, prev[j].x);
              if (okb) put(ob + x, prev[j].y);
            }
          });
        }
        static_for<0, N2>([&](auto ee) { v[decltype(ee)::value] = mk<T>(T(0), T(0)); });
        cx += HALF;
        emit_left = true;
      }
      s = s + 1 == STAGES ? 0 : s + 1;
    }
    pol.done(f, task, live && t == 0);                       // every step's copy has landed and been consumed
  }
}

template <int P, typename T, bool MIRROR>
__global__ void __launch_bounds__(Stream<P, T>::THREADS, 1)
k3_stream(const cplx<T>* __restrict__ spec, T* __restrict__ out, const StreamTask* __restrict__ tasks,
          const unsigned* __restrict__ codes, int n_warp_items, const cplx<T>* __restrict__ tw_g,
          const T* __restrict__ win_g, ApplyGeom g, int batch, OutMirrors mir) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  k3_stream_body<P, T, MIRROR>(spec, out, tasks, codes, n_warp_items, tw_g, win_g, g, batch, mir, blockIdx.x, gridDim.x,
                               PlainK3<P>{}, smem_raw);
}
// the same over the paired workspace of k2_chain (one item per group, row windows already applied)
template <int P, typename T, bool MIRROR>
__global__ void __launch_bounds__(Stream<P, T>::THREADS, 1)
k3_stream_paired(const cplx<T>* __restrict__ paired, T* __restrict__ out, const StreamTask* __restrict__ tasks,
                 const unsigned* __restrict__ codes, int n_warp_items, const cplx<T>* __restrict__ tw_g,
                 const T* __restrict__ win_g, ApplyGeom g, int batch, OutMirrors mir, long long items_per_frame) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  k3_stream_body<P, T, MIRROR>(paired, out, tasks, codes, n_warp_items, tw_g, win_g, g, batch, mir, blockIdx.x, gridDim.x,
                               PairedK3<P>{items_per_frame}, smem_raw);
}

}  // namespace rpsf
