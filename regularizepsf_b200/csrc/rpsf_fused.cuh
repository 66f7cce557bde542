// rpsf_fused.cuh — the whole of apply() as ONE persistent launch whose intermediates never leave L2.
//
// The three stand-alone kernels (rpsf_stream.cuh, rpsf_kernels.cuh) hand the spectrum workspace over through HBM
// twice: K1 writes it, K2 reads and rewrites it, K3 reads it — 4 x 75.8 MB per 2048^2 frame against 33.5 MB of
// frame bytes in and out.  One patch (P = 256: 256 KB as a packed half-spectrum) does not fit one SM's shared
// memory, so the hand-over cannot stay on chip; but it can stay in the 126 MB L2 if consumer follows producer
// closely.  Here the SMs of one cooperative launch are split into three roles that run the SAME arithmetic as
// the stand-alone kernels (k1_stream_body / the column pass below / k3_stream_body) and are chained by counters
// in global memory:
//
//   role K1 (gather + window + row FFT)   produces patch spectra into a RING of `ring` band slots
//   role K2 (column FFT x kernel x IFFT)  takes (patch, tile) units from a ticket as their patch completes
//   role K3 (row IFFT + window + overlap-add) runs a row-pair task when the two patch rows it sums are complete,
//                                         and frees a band slot for K1 when every task that reads it is done
//
// A band = the active patches that share one corner row (a "patch row": 17 patches of a 2048^2 / 256-px covering,
// 4.45 MB of spectrum); bands are numbered in corner-row order, frame-major: seq = frame * n_bands + band, and
// band seq lives in ring slot seq mod ring.  Every wait points at a strictly earlier position of that one
// sequence, each role walks its work in sequence order and all CTAs are co-resident (cooperative launch), so the
// pipeline cannot deadlock.  HBM then sees the frame in, the frame out and the transfer kernel; the spectrum
// lives in L2 (the ring is rewritten in place, so its lines are never written back either).
//
// Counters (zeroed by the host before the launch):
//   ready1[frame * n_active + a]  += 1 per group of WARPS K1 warp items of the patch     complete at IPP / WARPS
//   ready2[seq]                   += 1 per K2 unit of the band              complete at band_units[band]
//   done3[seq]                    += 1 per K3 task that read the band       slot free at band_tasks[band]
//   ticket                        K2's unit counter
#pragma once
#include "rpsf_stream.cuh"

namespace rpsf {

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// `waited` (may be null): cycles this wait cost, for the pipeline statistics of rpsf_plan_fused_stats
__device__ __forceinline__ void spin_until(const int* p, int target, unsigned long long* waited = nullptr) {
  if (ld_acquire(p) >= target) return;
  const long long t0 = clock64();
  do { __nanosleep(64); } while (ld_acquire(p) < target);
  if (waited) atomicAdd(waited, (unsigned long long)(clock64() - t0));
}
// publish this thread's (and, through the preceding barrier, its team's / CTA's) global writes, then count
// (red.release.gpu = MEMBAR.ALL.GPU + REDG: no L1 invalidation and no sequentially consistent fence, unlike
// __threadfence() + atomicAdd)
__device__ __forceinline__ void release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int atomic_release_add(int* p, int v) {
  int old;
  asm volatile("atom.release.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct FusedGeom {
  int n_bands;              // patch-row bands per frame
  int ring_mask;            // ring slots (bands) - 1; power of two
  int band_cap;             // patches per ring slot (the widest band)
  int n1, n2, n3;           // CTAs per role
  int units_per_patch;      // K2 units per patch = NTILE / SLOTS
  const int2* slot;         // per active patch (sorted by band): (band, index inside the band)
  const int* band_units;    // [n_bands] K2 units of the band
  const int* band_tasks;    // [n_bands] K3 tasks that read the band
  int* ready1;              // [batch * n_active]
  int* ready2;              // [batch * n_bands]
  int* done3;               // [batch * n_bands]
  unsigned* ticket;         // [1]
  // optional statistics (null = off): [0..2] cycles roles K1 / K2 / K3 spent waiting on a counter (one sample per
  // warp / group), [3..5] cycles from role start to role end summed over the role's CTAs, [6] K2 units whose tile
  // could not be prefetched because their patch was not complete yet, [7] K2 units
  unsigned long long* stats;
  // optional timeline (null = off): [3][batch * n_bands] globaltimer ns at which band seq was completed by K1
  // (all its patches published), by K2 (ready2 full) and by K3 (done3 full = slot free); [3 * batch * n_bands ..]
  // scratch: patches of the band K1 has completed
  unsigned long long* trace;
  int trace_stride;         // batch * n_bands
  int n_active;
  // diagnostics: 0 = the pipeline; 1 / 2 / 3 = only that role runs and never waits (its throughput in isolation on
  // its share of the SMs; results are garbage); 11 / 12 = role 1 / 2 alone without its publishes, 22 = role 2 alone
  // without publishes and without its transfer-kernel reads
  int solo;
};
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// item code of the fused overlap-add tables: pair | index-in-band << 9 | (band - task.band_lo) << 21 | LAST flag
constexpr int FUSED_IDX_SHIFT = 9, FUSED_BAND_SHIFT = 21;

// One poller per CTA.  Hundreds of warps spinning on the same few counters in L2 would hot-spot one L2 slice (and
// queue behind each other the very atomics they wait for), so the counters of a sequence are folded into a per-CTA
// watermark in shared memory: "every position <= wm is complete".  A warp that needs more takes the CTA's lock and
// polls global memory (one lane, with back-off) position by position — the producers complete them in order —
// while the CTA's other warps watch the shared word.
struct CtaWatermark { int wm; int lock; };

// The watermark word is read and written with shared-memory ATOMICS carrying acquire / release semantics at CTA
// scope: warps poll it while another one advances it, by design without a barrier in between (plain loads and
// stores with the same semantics work too, but compute-sanitizer's racecheck cannot tell them from a data race).
__device__ __forceinline__ int lds_acquire(const int* p) {
  int v;
  asm volatile("atom.acquire.cta.shared.or.b32 %0, [%1], 0;" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void sts_release(int* p, int v) {
  int old;
  asm volatile("atom.release.cta.shared.exch.b32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
}
// lane 0 of a warp: return once every position <= need of `counter` has reached target[position % n_bands]
__device__ __forceinline__ void watermark_wait(CtaWatermark* w, int need, const int* counter, const int* target, int n_bands,
                                               unsigned long long* waited) {
  if (lds_acquire(&w->wm) >= need) return;
  const long long t0 = clock64();
  for (;;) {
    if (lds_acquire(&w->wm) >= need) break;
    if (atomicCAS(&w->lock, 0, 1) == 0) {
      int have = lds_acquire(&w->wm);
      while (have < need) {
        const int pos = have + 1;
        const int tgt = __ldg(target + pos % n_bands);
        while (ld_acquire(counter + pos) < tgt) __nanosleep(200);
        have = pos;
        sts_release(&w->wm, have);
      }
      atomicExch(&w->lock, 0);
      break;
    }
    __nanosleep(100);
  }
  if (waited) atomicAdd(waited, (unsigned long long)(clock64() - t0));
}

struct FusedK1 {
  FusedGeom fg;
  CtaWatermark* wm;         // shared: ring slots of every band seq <= wm + ring are free
  int cleared;              // this thread's copy of what it has already checked
  int2 next;                // (band, index) of the patch after the head item, fetched one iteration early
  int next_a;
  __device__ __forceinline__ void prefetch(int a) { next = __ldg(fg.slot + a); next_a = a; }
  __device__ __forceinline__ unsigned slot(int f, int a, int /*n_active*/) {
    const int2 s = a == next_a ? next : __ldg(fg.slot + a);
    const int seq = f * fg.n_bands + s.x;
    const int prev = seq - (fg.ring_mask + 1);               // the band that used this slot before
    if (prev > cleared && !fg.solo) {                         // warp-uniform: every lane decodes the same head item
      if ((threadIdx.x & 31) == 0) watermark_wait(wm, prev, fg.done3, fg.band_tasks, fg.n_bands, fg.stats ? fg.stats + 0 : nullptr);
      __syncwarp();
      cleared = prev;
    }
    return (unsigned)((seq & fg.ring_mask) * fg.band_cap + s.y);
  }
  // A finished item is first counted in shared memory (a CTA-scope release: cheap).  The `warps` items
  // [warps * k, warps * (k + 1)) are processed by the warps of ONE CTA at about the same time and lie in one
  // patch; the item that completes such a group publishes it to the other SMs — one gpu-scope release per group
  // instead of one per warp item.  (A gpu-scope release drains the SM's outstanding stores, about a microsecond:
  // per item it cost the role a third of its throughput.)  Groups are counted by their turn in this CTA, modulo
  // 32 — far more turns than the ring lets warps of a CTA drift apart.
  int* counts;              // shared: [32]
  unsigned n_cta, warps;
  __device__ __forceinline__ void publish(unsigned item, int lane) {
    if (lane == 0 && item != PlainK1::kNone) {
      if (fg.solo == 11) return;                              // diagnostics: K1 alone, publishing nothing
      const unsigned group = item / warps, pf = item / (unsigned)ipp;
      int* c = counts + ((group / n_cta) & 31u);
      int old;
      asm volatile("atom.acq_rel.cta.shared.add.s32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(c)) : "memory");
      if (old + 1 < (int)warps) return;
      *c = 0;                                                 // the group is complete: nobody adds to this counter before its next turn
      const int before = atomic_release_add(fg.ready1 + pf, 1);
      if (fg.trace && before + 1 == ipp / (int)warps) {       // that group completed its patch
        const int a = int(pf % (unsigned)fg.n_active), f = int(pf / (unsigned)fg.n_active);
        const int band = __ldg(fg.slot + a).x, seq = f * fg.n_bands + band;
        const int units = __ldg(fg.band_units + band) / fg.units_per_patch;     // patches of the band
        if (atomicAdd(fg.trace + 3 * (size_t)fg.trace_stride + seq, 1ull) + 1 == (unsigned long long)units)
          fg.trace[seq] = globaltimer_ns();
      }
    }
  }
  int ipp;
};

template <int P> struct FusedK3 {
  FusedGeom fg;
  CtaWatermark* wm;         // shared: K2 has finished every band seq <= wm
  __device__ __forceinline__ int pair(unsigned code) const { return int(code & ((1u << FUSED_IDX_SHIFT) - 1)); }
  __device__ __forceinline__ size_t offset(int f, int /*n_active*/, unsigned code, const StreamTask& task) const {
    const int idx = int((code >> FUSED_IDX_SHIFT) & ((1u << (FUSED_BAND_SHIFT - FUSED_IDX_SHIFT)) - 1));
    const int band = task.pad0 + int((code >> FUSED_BAND_SHIFT) & 7u);
    const int slot = ((f * fg.n_bands + band) & fg.ring_mask) * fg.band_cap + idx;
    return ((size_t)slot * (P / 2) + pair(code)) * P;
  }
  template <typename T> __device__ __forceinline__ void row_windows(const T* win, unsigned code, T& wa, T& wb) const {
    wa = win[2 * pair(code)]; wb = win[2 * pair(code) + 1];
  }
  // task.pad0 / pad1 = first / last band the task reads; the warp waits for the last band any of its teams reads
  __device__ __forceinline__ void wait(int f, const StreamTask& task, bool live) const {
    const int need = __reduce_max_sync(0xffffffffu, live ? f * fg.n_bands + task.pad1 : -1);
    if ((threadIdx.x & 31) == 0 && !fg.solo) watermark_wait(wm, need, fg.ready2, fg.band_units, fg.n_bands, fg.stats ? fg.stats + 2 : nullptr);
    __syncwarp();
    fence_proxy_async_all();        // the bulk (async-proxy) copies below read what other SMs wrote through the generic proxy
  }
  // no fence: what must precede the slot's reuse are this task's READS of the ring, and those completed when
  // their mbarriers did (the data is in shared memory); the task's own output stores are nobody's input
  __device__ __forceinline__ void done(int f, const StreamTask& task, bool leader) const {
    if (leader)
      for (int b = task.pad0; b <= task.pad1; ++b) {
        const int seq = f * fg.n_bands + b;
        const int old = atomicAdd(fg.done3 + seq, 1);
        if (fg.trace && old + 1 == __ldg(fg.band_tasks + b)) fg.trace[2 * (size_t)fg.trace_stride + seq] = globaltimer_ns();
      }
  }
};

// ---------------------------------------------------------------------------- role K2
// One group = Tile<P>::K2_THREADS threads (a named barrier each) = SLOTS column tiles of one patch of one frame.
// Same arithmetic as k2_frames (rpsf_kernels.cuh); the unit stream replaces the frame loop: while unit u is
// transformed, the tile(s) of unit u+1 stream into the other stage if their patch is already complete.
template <int P, typename T> struct FusedK2 {
  using TL = Tile<P>;
  static constexpr int GROUPS = 512 / TL::K2_THREADS;
  static constexpr int STAGE = TL::SLOTS * P * TL::C;                       // complex elements per stage
  static constexpr size_t SMEM = sizeof(cplx<T>) * (P + (size_t)GROUPS * 2 * STAGE) + 64;
  static_assert(512 % TL::K2_THREADS == 0, "column groups must tile the CTA");
};

// `mid()` runs right after the barrier of the forward exchange: every thread of the group has then left the
// previous unit, whose stage may be refilled (the prefetch of the next unit's tile).
// `kv` holds this unit's transfer-kernel tile (loaded by the caller, possibly still in flight); `after_mul()` runs
// once the tile has been consumed and may refill kv with the next unit's (the loads then have the whole inverse
// transform and the next forward transform to land).
template <int P, typename T, bool TILE0, typename Mid, typename AfterMul>
__device__ __forceinline__ void fused_k2_unit(cplx<T>* __restrict__ base, cplx<T> (&kv)[Tile<P>::N2],
                                              const cplx<T>* __restrict__ kn, const cplx<T>* tw, cplx<T>* xbuf,
                                              bool special, int c, int n1, int slot, int bar_id, Mid mid, AfterMul after_mul) {
  using TL = Tile<P>;
  constexpr int N1 = TL::N1, N2 = TL::N2, HALF = TL::HALF, C = TL::C;
  auto ex = [=](int k2, int nn) { return ((slot * N2 + k2) * N1 + nn) * C + c; };
  auto sync = [=]() { bar_sync(bar_id, TL::K2_THREADS); };
  auto nosync = []() {};
  cplx<T> v[N2];
  static_for<0, N2>([&](auto jj) { v[decltype(jj)::value] = xbuf[ex(decltype(jj)::value, n1)]; });
  auto sync_mid = [=]() { bar_sync(bar_id, TL::K2_THREADS); mid(); };
  if constexpr (TILE0) coop_fft_forward<P, T>(v, n1, xbuf, tw, ex, sync_mid, sync);
  else coop_fft_forward<P, T>(v, n1, xbuf, tw, ex, sync_mid, nosync);
  if constexpr (TILE0) {
    cplx<T>* zs = xbuf + slot * P;                          // natural order, one column per slot
    if (special) {
      static_for<0, N2>([&](auto ee) {
        constexpr int e = decltype(ee)::value;
        zs[(n1 + N1 * (e / N1)) + N2 * (e % N1)] = v[e];
      });
    }
    sync();
    static_for<0, N2>([&](auto ee) {
      constexpr int e = decltype(ee)::value;
      const int k = (n1 + N1 * (e / N1)) + N2 * (e % N1);
      cplx<T> zm = v[e], kny = mk<T>(T(0), T(0));
      if (special) {
        const cplx<T> zr = zs[(P - k) & (P - 1)];
        zm = mk<T>(zr.x, -zr.y);
        kny = kn[e * N1];
      }
      const cplx<T> sum = mk<T>(T(0.5) * (v[e].x + zm.x), T(0.5) * (v[e].y + zm.y));
      const cplx<T> dif = mk<T>(T(0.5) * (v[e].x - zm.x), T(0.5) * (v[e].y - zm.y));
      v[e] = cadd(cmul(sum, kv[e]), cmul(dif, kny));
    });
    after_mul();
    sync();
  } else {
    static_for<0, N2>([&](auto ee) { v[decltype(ee)::value] = cmul(v[decltype(ee)::value], kv[decltype(ee)::value]); });
    after_mul();
  }
  coop_fft_inverse<P, T>(v, n1, xbuf, tw, ex, sync, nosync);
  static_for<0, N2>([&](auto jj) {
    constexpr int j = decltype(jj)::value;
    base[(long long)(n1 + N1 * j) * HALF] = v[j];
  });
}

template <int P, typename T>
__device__ __forceinline__ void fused_k2_role(cplx<T>* __restrict__ spec, const cplx<T>* __restrict__ kmain,
                                              const cplx<T>* __restrict__ knyq, const int* __restrict__ active,
                                              const cplx<T>* __restrict__ tw_g, int batch, int n_active,
                                              const FusedGeom& fg, unsigned char* smem_raw) {
  using TL = Tile<P>;
  using F2 = FusedK2<P, T>;
  constexpr int N1 = TL::N1, N2 = TL::N2, HALF = TL::HALF, C = TL::C, NTILE = TL::NTILE, SLOTS = TL::SLOTS;
  constexpr int GT = TL::K2_THREADS, STAGE = F2::STAGE;
  constexpr int PATCH_GROUPS = Stream<P, T>::IPP / Stream<P, T>::WARPS;     // K1 publishes a patch in this many pieces
  constexpr int CH = 16 / (int)sizeof(cplx<T>);
  constexpr int ROW_CHUNKS = C / CH;
  constexpr int PER_THREAD = (P * ROW_CHUNKS) / TL::SLOT_THREADS;
  constexpr int ROWS_PER_PASS = TL::SLOT_THREADS / ROW_CHUNKS;
  cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);
  cplx<T>* stages = reinterpret_cast<cplx<T>*>(smem_raw + sizeof(cplx<T>) * P + 64);
  for (int i = threadIdx.x; i < P; i += blockDim.x) tw[i] = tw_g[i];
  __syncthreads();

  const int grp = threadIdx.x / GT, gt = threadIdx.x % GT;
  const int bar_id = 1 + grp;
  const int c = gt % C, n1 = (gt / C) % N1, slot = gt / (C * N1), lt = gt % TL::SLOT_THREADS;
  cplx<T>* stage0 = stages + (size_t)grp * 2 * STAGE;
  const int upp = fg.units_per_patch;
  const unsigned upf = (unsigned)n_active * (unsigned)upp;     // units per frame
  const unsigned total = (unsigned)batch * upf;
  const int row0 = lt / ROW_CHUNKS, part = lt % ROW_CHUNKS;

  struct Unit { int f, a, tg; };
  auto decode = [&](unsigned u) { Unit r; r.f = int(u / upf); const unsigned w = u % upf; r.a = int(w / upp); r.tg = int(w % upp); return r; };
  auto patch_base = [&](const Unit& u) -> cplx<T>* {
    const int2 s = __ldg(fg.slot + u.a);
    const int ps = ((u.f * fg.n_bands + s.x) & fg.ring_mask) * fg.band_cap + s.y;
    return spec + (size_t)ps * P * HALF;
  };
  auto issue = [&](const Unit& u, cplx<T>* stage) {
    const cplx<T>* src = patch_base(u) + (u.tg * SLOTS + slot) * C + (long long)row0 * HALF + part * CH;
    cplx<T>* dst = stage + slot * (P * C) + row0 * C + part * CH;
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) cp_async16(dst + i * (ROWS_PER_PASS * C), src + (long long)i * (ROWS_PER_PASS * HALF));
  };
  // Units go to the groups round robin (they all cost the same).  Thread 0 of a group is its only poller: its
  // verdict on the next unit travels through shared memory under the barrier of the forward exchange.
  const unsigned n_groups = (unsigned)fg.n2 * F2::GROUPS;
  const unsigned my = (unsigned)(blockIdx.x - fg.n1) * F2::GROUPS + grp;
  int* verdict = reinterpret_cast<int*>(smem_raw + sizeof(cplx<T>) * P) + grp;
  auto is_ready = [&](const Unit& u) { return fg.solo || ld_acquire(fg.ready1 + u.f * n_active + u.a) >= PATCH_GROUPS; };
  auto wait_ready = [&](const Unit& u) {      // blocking; ends with a group barrier
    if (gt == 0 && !fg.solo) {
      const int* flag = fg.ready1 + u.f * n_active + u.a;
      if (ld_acquire(flag) < PATCH_GROUPS) {
        const long long t0 = clock64();
        do { __nanosleep(200); } while (ld_acquire(flag) < PATCH_GROUPS);
        if (fg.stats) atomicAdd(fg.stats + 1, (unsigned long long)(clock64() - t0));
      }
    }
    bar_sync(bar_id, GT);
  };

  auto publish_band = [&](int seq) {
    if (fg.solo == 12 || fg.solo == 22) return;                // diagnostics: the role without its publishes
    const int old = atomic_release_add(fg.ready2 + seq, 1);
    if (fg.trace && old + 1 == __ldg(fg.band_units + seq % fg.n_bands)) fg.trace[(size_t)fg.trace_stride + seq] = globaltimer_ns();
  };
  auto kernel_ptr = [&](const Unit& u) {
    const int tile = u.tg * SLOTS + slot;
    return kmain + (((long long)__ldg(active + u.a) * NTILE + tile) * N2) * (N1 * C) + n1 * C + c;
  };
  unsigned cur = my;
  if (cur >= total) return;
  Unit cu = decode(cur);
  cplx<T> kv[N2];                                              // transfer-kernel tile of the current unit
  static_for<0, N2>([&](auto ee) { kv[decltype(ee)::value] = mk<T>(T(1), T(0)); });
  wait_ready(cu);
  int st = 0;
  issue(cu, stage0);
  cp_async_commit();
  int pending_band = -1;                                       // band seq of the unit stored last
  for (unsigned it = 0;; ++it) {
    cp_async_wait_all();
    bar_sync(bar_id, GT);                                      // unit `cur` has landed for the whole group; the previous unit's stores are issued
    // one thread publishes the previous unit (the barrier makes the group's stores its own); a different warp each
    // time, so the fence's wait for that thread's stores is not always the same warp's
    if (pending_band >= 0 && gt == (int)(it % (GT / 32)) * 32) publish_band(pending_band);
    const unsigned nxt = cur + n_groups;
    const bool has_next = nxt < total;
    const Unit nu = decode(has_next ? nxt : cur);
    if (gt == 0) *verdict = has_next && is_ready(nu);          // read by the group after the forward exchange barrier
    bool pre = false;
    auto mid = [&]() {
      pre = *verdict != 0;
      if (pre) issue(nu, stage0 + (st ^ 1) * STAGE);
      cp_async_commit();
    };
    {
      const int tile = cu.tg * SLOTS + slot;
      const cplx<T>* kn = knyq + (long long)__ldg(active + cu.a) * P + n1;
      cplx<T>* base = patch_base(cu) + tile * C + c;
      const bool special = tile == 0 && c == 0;
      if (fg.solo != 22) {                                     // (22: diagnostics, the role without its kernel reads)
        const cplx<T>* kp = kernel_ptr(cu);
        static_for<0, N2>([&](auto ee) { kv[decltype(ee)::value] = kp[(long long)decltype(ee)::value * (N1 * C)]; });
      }
      auto after_mul = [&]() {};
      if (cu.tg == 0) fused_k2_unit<P, T, true>(base, kv, kn, tw, stage0 + st * STAGE, special, c, n1, slot, bar_id, mid, after_mul);
      else fused_k2_unit<P, T, false>(base, kv, kn, tw, stage0 + st * STAGE, special, c, n1, slot, bar_id, mid, after_mul);
      pending_band = cu.f * fg.n_bands + __ldg(fg.slot + cu.a).x;
    }
    if (fg.stats && gt == 0) { atomicAdd(fg.stats + 7, 1ull); if (has_next && !pre) atomicAdd(fg.stats + 6, 1ull); }
    if (!has_next) break;
    if (!pre) {                                                // the patch was not complete in time: wait for it now
      wait_ready(nu);                                          // (its barrier also separates this read of the verdict from the next write)
      issue(nu, stage0 + (st ^ 1) * STAGE);
      cp_async_commit();
    }
    cur = nxt; cu = nu; st ^= 1;
  }
  bar_sync(bar_id, GT);                                        // the last unit's stores
  if (gt == 0) publish_band(pending_band);
}

// ---------------------------------------------------------------------------- the launch
template <int P, typename T> struct Fused {
  static constexpr int THREADS = 512;
  static constexpr size_t SMEM = Stream<P, T>::SMEM > FusedK2<P, T>::SMEM ? Stream<P, T>::SMEM : FusedK2<P, T>::SMEM;
  static constexpr bool OK = Stream<P, T>::THREADS == 512 && Stream<P, T>::WARPS == 16 && FusedK2<P, T>::SMEM <= (size_t)STREAM_SMEM_BUDGET &&
                             Tile<P>::NTILE % Tile<P>::SLOTS == 0 && P / 2 <= (1 << FUSED_IDX_SHIFT) &&
                             Stream<P, T>::IPP % Stream<P, T>::WARPS == 0;      // a CTA's warps share out whole patches
};

template <int P, typename T>
__global__ void __launch_bounds__(512, 1)
fused_apply(const T* __restrict__ image, cplx<T>* __restrict__ spec, T* __restrict__ out,
            const int2* __restrict__ corners, const int* __restrict__ active, const cplx<T>* __restrict__ kmain,
            const cplx<T>* __restrict__ knyq, const StreamTask* __restrict__ tasks, const unsigned* __restrict__ codes,
            int n_warp_items, const cplx<T>* __restrict__ tw_g, const T* __restrict__ win_g, ApplyGeom g_in, ApplyGeom g,
            int batch, int bulk_ok, FusedGeom fg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const long long t_start = fg.stats ? clock64() : 0;
  const int role = b < fg.n1 ? 0 : b < fg.n1 + fg.n2 ? 1 : 2;
  if (fg.solo && fg.solo % 10 != role + 1) return;
  CtaWatermark* wm = reinterpret_cast<CtaWatermark*>(smem_raw + Stream<P, T>::TABLE_BYTES + 768);   // behind the mbarriers
  if (role != 1 && threadIdx.x == 0) { wm->wm = -1; wm->lock = 0; }      // the role bodies start with a CTA barrier
  if (b < fg.n1) {
    int* counts = reinterpret_cast<int*>(wm + 1);
    if (threadIdx.x < 32) counts[threadIdx.x] = 0;
    FusedK1 pol{fg, wm, -1, make_int2(0, 0), -1, counts, (unsigned)fg.n1, (unsigned)Stream<P, T>::WARPS, Stream<P, T>::IPP};
    k1_stream_body<P, T>(image, spec, corners, tw_g, win_g, g_in, batch, bulk_ok, (unsigned)b, (unsigned)fg.n1, pol, smem_raw);
  } else if (b < fg.n1 + fg.n2) {
    fused_k2_role<P, T>(spec, kmain, knyq, active, tw_g, batch, g.n_active, fg, smem_raw);
  } else {
    FusedK3<P> pol{fg, wm};
    k3_stream_body<P, T, false>(spec, out, tasks, codes, n_warp_items, tw_g, win_g, g, batch, OutMirrors{},
                                (unsigned)(b - fg.n1 - fg.n2), (unsigned)fg.n3, pol, smem_raw);
  }
  if (fg.stats && (threadIdx.x & 31) == 0)                    // per warp: warps of a role finish at different times
    atomicAdd(fg.stats + 3 + role, (unsigned long long)(clock64() - t_start));
}

}  // namespace rpsf
