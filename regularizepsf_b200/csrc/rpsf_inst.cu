// rpsf_inst.cu — instantiates every kernel for one patch size (compile with -DRPSF_P=<P>).
#include <cstdlib>
#include "rpsf_ops.h"

#ifndef RPSF_P
#error "compile with -DRPSF_P=<patch size>"
#endif

namespace rpsf {
namespace {

constexpr int P = RPSF_P;
using TL = Tile<P>;

template <typename T> constexpr size_t row_smem() {
  return sizeof(cplx<T>) * P + sizeof(T) * P + sizeof(cplx<T>) * size_t(TL::TEAMS) * TL::SCR;
}
template <typename T> constexpr size_t col_smem() {
  return sizeof(cplx<T>) * P + sizeof(cplx<T>) * size_t(TL::SLOTS) * P * TL::C;
}
template <typename T> constexpr size_t col_pipe_smem() {
  return sizeof(cplx<T>) * P + RPSF_K2_STAGES * sizeof(cplx<T>) * size_t(TL::SLOTS) * P * TL::C;
}
// the pipelined column kernel needs RPSF_K2_STAGES tile stages per CTA; use it while a stage is at most
// 32 KB (two or three CTAs then fit an SM)
template <typename T> constexpr bool use_col_pipe() {
  return sizeof(cplx<T>) * size_t(TL::SLOTS) * P * TL::C <= 32 * 1024 + 4096;
}
template <typename T> constexpr size_t fft2_row_smem() {
  return sizeof(cplx<T>) * P + sizeof(cplx<T>) * size_t(TL::TEAMS) * TL::SCR;
}

template <typename T> constexpr size_t chain_smem() {
  return sizeof(cplx<T>) * P + sizeof(T) * P + 2 * sizeof(cplx<T>) * size_t(TL::SLOTS) * P * TL::C +
         sizeof(cplx<T>) * size_t(TL::K2_THREADS) * (TL::N2 / 2);
}
// two CTAs per SM must fit, as for k2_pipelined
template <typename T> constexpr bool use_chain() {
  return use_col_pipe<T>() && 2 * chain_smem<T>() <= 220 * 1024 && TL::NTILE % TL::SLOTS == 0 && TL::NTILE >= TL::SLOTS;
}

// Launch one of the chained kernels (k1_stream -> k2_pipelined -> k3_stream -> the next call's k1_stream) so that its
// CTAs may start while the kernel before it in the stream drains (programmatic dependent launch; the kernels wait in
// grid_dependency_wait() before they touch anything a predecessor wrote).  RPSF_PDL=0 launches them the plain way.
// RPSF_PDL is a mask: 1 = k1_stream, 2 = k2_pipelined, 4 = k3_stream may start early.  Measured (scripts/chain_ab.py):
// all three at P <= 128 (one 1024^2 frame: 41.1 -> 34.1 us at 128 px, 39.4 -> 31.8 at 64 px); at P >= 256 an early
// column pass costs 1 % (its first wave of CTAs is placed on the SMs the row pass leaves first), so only the row
// kernels start early there (2048^2 / 256 px: 97.0 -> 95.2 us for one frame, 64.4 -> 63.5 per frame at 8).
inline int pdl_mask() {
  static const int mask = [] { const char* v = getenv("RPSF_PDL"); return v ? atoi(v) : (P <= 128 ? 7 : 5); }();
  return mask;
}
template <typename... KArgs, typename... Args>
int launch_chain(int which, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_mask() & which) ? 1 : 0;
  return (int)cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename F> int set_smem(F* fn, size_t bytes) {
  return (int)cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int init() {
  int e = 0;
  if ((e = set_smem(k1_gather_window_rowfft<P, float>, row_smem<float>()))) return e;
  if ((e = set_smem(k1_gather_window_rowfft<P, double>, row_smem<double>()))) return e;
  if ((e = set_smem(k1_stream<P, float>, StreamK1<P, float>::SMEM))) return e;
  if ((e = set_smem(k1_stream<P, double>, StreamK1<P, double>::SMEM))) return e;
  if ((e = set_smem(k3_stream<P, float, false>, Stream<P, float>::SMEM))) return e;
  if ((e = set_smem(k3_stream<P, double, false>, Stream<P, double>::SMEM))) return e;
  if ((e = set_smem(k3_stream<P, float, true>, Stream<P, float>::SMEM))) return e;
  if ((e = set_smem(k3_stream<P, double, true>, Stream<P, double>::SMEM))) return e;
  if constexpr (Fused<P, float>::OK)
    if ((e = set_smem(fused_apply<P, float>, Fused<P, float>::SMEM))) return e;
  if ((e = set_smem(k2_colfft_mul_colifft<P, float>, col_smem<float>()))) return e;
  if ((e = set_smem(k2_colfft_mul_colifft<P, double>, col_smem<double>()))) return e;
  if ((e = set_smem(k2_colfft_mul_colifft<P, float, true>, col_smem<float>()))) return e;
  if ((e = set_smem(k2_colfft_mul_colifft<P, double, true>, col_smem<double>()))) return e;
  if constexpr (use_col_pipe<float>())
    if ((e = set_smem(k2_pipelined<P, float, 1, true>, col_pipe_smem<float>()))) return e;
  if constexpr (use_col_pipe<double>())
    if ((e = set_smem(k2_pipelined<P, double, 1, true>, col_pipe_smem<double>()))) return e;
  if constexpr (use_col_pipe<float>())
    if ((e = set_smem(k2_pipelined<P, float>, col_pipe_smem<float>()))) return e;
  if constexpr (use_col_pipe<double>())
    if ((e = set_smem(k2_pipelined<P, double>, col_pipe_smem<double>()))) return e;
  if constexpr (use_col_pipe<float>() && TL::SLOTS == 1 && TL::NTILE % 2 == 0)
    if ((e = set_smem(k2_pipelined<P, float, 2>, col_pipe_smem<float>()))) return e;
  if constexpr (use_col_pipe<double>() && TL::SLOTS == 1 && TL::NTILE % 2 == 0)
    if ((e = set_smem(k2_pipelined<P, double, 2>, col_pipe_smem<double>()))) return e;
  if constexpr (use_col_pipe<float>() && TL::SLOTS == 1 && TL::NTILE % 4 == 0)
    if ((e = set_smem(k2_pipelined<P, float, 4>, col_pipe_smem<float>()))) return e;
  if constexpr (use_col_pipe<double>() && TL::SLOTS == 1 && TL::NTILE % 4 == 0)
    if ((e = set_smem(k2_pipelined<P, double, 4>, col_pipe_smem<double>()))) return e;
  if constexpr (Small<P, float>::OK)
    if ((e = set_smem(small_patch<P, float>, Small<P, float>::SMEM))) return e;
  if constexpr (Small<P, double>::OK)
    if ((e = set_smem(small_patch<P, double>, Small<P, double>::SMEM))) return e;
  if constexpr (use_chain<float>()) {
    if ((e = set_smem(k2_chain<P, float>, chain_smem<float>()))) return e;
    if ((e = set_smem(k3_stream_paired<P, float, false>, Stream<P, float>::SMEM))) return e;
    if ((e = set_smem(k3_stream_paired<P, float, true>, Stream<P, float>::SMEM))) return e;
  }
  if constexpr (use_chain<double>()) {
    if ((e = set_smem(k2_chain<P, double>, chain_smem<double>()))) return e;
    if ((e = set_smem(k3_stream_paired<P, double, false>, Stream<P, double>::SMEM))) return e;
    if ((e = set_smem(k3_stream_paired<P, double, true>, Stream<P, double>::SMEM))) return e;
  }
  if ((e = set_smem(k3_rowifft_window_overlap_add<P, float>, row_smem<float>()))) return e;
  if ((e = set_smem(k3_rowifft_window_overlap_add<P, double>, row_smem<double>()))) return e;
  if ((e = set_smem(k3_rowpair_gather<P, float>, K3G_SMEM_MAX))) return e;
  if ((e = set_smem(k3_rowpair_gather<P, double>, K3G_SMEM_MAX))) return e;
  if ((e = set_smem(fft2_rows<P, float, float>, fft2_row_smem<float>()))) return e;
  if ((e = set_smem(fft2_rows<P, double, double>, fft2_row_smem<double>()))) return e;
  if ((e = set_smem(fft2_cols<P, float>, col_smem<float>()))) return e;
  if ((e = set_smem(fft2_cols<P, double>, col_smem<double>()))) return e;
  return 0;
}

inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

template <typename T>
int k1_t(const void* image, void* spec, const int2* corners, const void* tw, const void* win,
         const ApplyGeom& g, int batch, cudaStream_t s) {
  dim3 grid(cdiv((long long)g.n_active * TL::HALF, TL::TEAMS), batch);
  k1_gather_window_rowfft<P, T><<<grid, TL::ROW_THREADS, row_smem<T>(), s>>>(
      (const T*)image, (cplx<T>*)spec, corners, (const cplx<T>*)tw, (const T*)win, g);
  return (int)cudaGetLastError();
}
int k1(int dt, const void* image, void* spec, const int2* corners, const void* tw, const void* win,
       const ApplyGeom& g, int batch, cudaStream_t s) {
  return dt == DT_F32 ? k1_t<float>(image, spec, corners, tw, win, g, batch, s)
                      : k1_t<double>(image, spec, corners, tw, win, g, batch, s);
}

template <typename T>
int k1s_t(const void* image, void* spec, const int2* corners, const void* tw, const void* win,
          const ApplyGeom& g, int batch, int bulk_ok, int sm_count, cudaStream_t s) {
  using ST = StreamK1<P, T>;
  const long long items = (long long)batch * g.n_active * ST::IPP;
  if (items == 0) return 0;
  const long long ctas = (items + ST::WARPS - 1) / ST::WARPS;
  const unsigned grid = (unsigned)(ctas < sm_count ? ctas : sm_count);
  return launch_chain(1, k1_stream<P, T>, dim3(grid), dim3(ST::THREADS), ST::SMEM, s, (const T*)image, (cplx<T>*)spec, corners,
                      (const cplx<T>*)tw, (const T*)win, g, batch, bulk_ok);
}
int k1s(int dt, const void* image, void* spec, const int2* corners, const void* tw, const void* win,
        const ApplyGeom& g, int batch, int bulk_ok, int sm_count, cudaStream_t s) {
  return dt == DT_F32 ? k1s_t<float>(image, spec, corners, tw, win, g, batch, bulk_ok, sm_count, s)
                      : k1s_t<double>(image, spec, corners, tw, win, g, batch, bulk_ok, sm_count, s);
}

template <typename T>
int k2_t(void* spec, const void* kmain, const void* knyq, const int* active, const void* tw,
         const ApplyGeom& g, int batch, int sm_count, cudaStream_t s) {
  // Walk several frames per CTA so the transfer-kernel tile is read once per batch, but keep
  // at least ~4 waves of CTAs in flight (SMs x 3 resident CTAs).
  const long long ctas = cdiv((long long)g.n_active * TL::NTILE, TL::SLOTS);
  int fpc = batch;
  while (fpc > 1 && ctas * cdiv(batch, fpc) < (long long)sm_count * 3 * 4) fpc = (fpc + 1) / 2;
  dim3 grid((unsigned)ctas, cdiv(batch, fpc));
  if (g.win_len < P) {
    // embedded patch size: only the first win_len rows of a patch's spectrum carry data (rpsf_kernels.cuh: PRUNED)
    if constexpr (use_col_pipe<T>())
      return launch_chain(2, k2_pipelined<P, T, 1, true>, grid, dim3(TL::K2_THREADS), col_pipe_smem<T>(), s, (cplx<T>*)spec,
                          (const cplx<T>*)kmain, (const cplx<T>*)knyq, active, (const cplx<T>*)tw, batch, fpc, g);
    else
      k2_colfft_mul_colifft<P, T, true><<<grid, TL::K2_THREADS, col_smem<T>(), s>>>(
          (cplx<T>*)spec, (const cplx<T>*)kmain, (const cplx<T>*)knyq, active, (const cplx<T>*)tw, batch, fpc, g);
    return (int)cudaGetLastError();
  }
  if constexpr (use_col_pipe<T>() && TL::SLOTS == 1 && TL::NTILE % 2 == 0) {
    // few frames per CTA: walk two adjacent tiles per CTA so that the stage ring and the transfer-kernel loads have
    // something to overlap with (RPSF_K2_TPC=1 switches it off).  Only for one-slot CTAs (P >= 256): with two slots
    // per CTA (P = 128) the same change was measured 2 % (1024^2) to 8 % (2048^2) slower, at P = 64 1 % faster.
    static const int tpc_env = [] { const char* v = getenv("RPSF_K2_TPC"); return v ? atoi(v) : 0; }();
    const int tpc = tpc_env ? tpc_env : (fpc <= 1 ? 2 : 1);
    if (tpc == 2) {
      grid.x = (unsigned)cdiv(ctas, 2);
      return launch_chain(2, k2_pipelined<P, T, 2>, grid, dim3(TL::K2_THREADS), col_pipe_smem<T>(), s, (cplx<T>*)spec,
                          (const cplx<T>*)kmain, (const cplx<T>*)knyq, active, (const cplx<T>*)tw, batch, fpc, g);
    }
    if constexpr (TL::NTILE % 4 == 0) {
      if (tpc == 4) {
        grid.x = (unsigned)cdiv(ctas, 4);
        return launch_chain(2, k2_pipelined<P, T, 4>, grid, dim3(TL::K2_THREADS), col_pipe_smem<T>(), s, (cplx<T>*)spec,
                            (const cplx<T>*)kmain, (const cplx<T>*)knyq, active, (const cplx<T>*)tw, batch, fpc, g);
      }
    }
  }
  if constexpr (use_col_pipe<T>())
    return launch_chain(2, k2_pipelined<P, T>, grid, dim3(TL::K2_THREADS), col_pipe_smem<T>(), s, (cplx<T>*)spec,
                        (const cplx<T>*)kmain, (const cplx<T>*)knyq, active, (const cplx<T>*)tw, batch, fpc, g);
  else
    k2_colfft_mul_colifft<P, T><<<grid, TL::K2_THREADS, col_smem<T>(), s>>>(
        (cplx<T>*)spec, (const cplx<T>*)kmain, (const cplx<T>*)knyq, active, (const cplx<T>*)tw, batch, fpc, g);
  return (int)cudaGetLastError();
}
int k2(int dt, void* spec, const void* kmain, const void* knyq, const int* active, const void* tw,
       const ApplyGeom& g, int batch, int sm_count, cudaStream_t s) {
  return dt == DT_F32 ? k2_t<float>(spec, kmain, knyq, active, tw, g, batch, sm_count, s)
                      : k2_t<double>(spec, kmain, knyq, active, tw, g, batch, sm_count, s);
}

template <typename T>
int k3_t(const void* spec, void* out, const int2* corners, const int* items, int n_items, const void* tw,
         const void* win, int store_only, const ApplyGeom& g, int batch, cudaStream_t s) {
  if (n_items == 0) return 0;
  dim3 grid(cdiv(n_items, TL::TEAMS), batch);
  k3_rowifft_window_overlap_add<P, T><<<grid, TL::ROW_THREADS, row_smem<T>(), s>>>(
      (const cplx<T>*)spec, (T*)out, corners, items, n_items, (const cplx<T>*)tw, (const T*)win, store_only, g);
  return (int)cudaGetLastError();
}
int k3(int dt, const void* spec, void* out, const int2* corners, const int* items, int n_items, const void* tw,
       const void* win, int store_only, const ApplyGeom& g, int batch, cudaStream_t s) {
  return dt == DT_F32 ? k3_t<float>(spec, out, corners, items, n_items, tw, win, store_only, g, batch, s)
                      : k3_t<double>(spec, out, corners, items, n_items, tw, win, store_only, g, batch, s);
}

template <typename T> size_t gather_smem(int teams, int seg_w) {
  constexpr size_t BUF = TL::SCR > P ? TL::SCR : P;
  return sizeof(cplx<T>) * P + sizeof(T) * P + sizeof(T) * 2 * (size_t)seg_w + sizeof(cplx<T>) * (size_t)teams * BUF;
}
size_t k3g_smem(int dt, int teams, int seg_w) {
  return dt == DT_F32 ? gather_smem<float>(teams, seg_w) : gather_smem<double>(teams, seg_w);
}
template <typename T>
int k3g_t(const void* spec, void* out, const RowTile* tiles, int n_tiles, const RowGroup* groups, const int* items,
          const void* tw, const void* win, int teams, int seg_w, const ApplyGeom& g, int batch, cudaStream_t s) {
  if (n_tiles == 0) return 0;
  dim3 grid(n_tiles, batch);
  k3_rowpair_gather<P, T><<<grid, teams * TL::N1, gather_smem<T>(teams, seg_w), s>>>(
      (const cplx<T>*)spec, (T*)out, tiles, groups, items, (const cplx<T>*)tw, (const T*)win, seg_w, g);
  return (int)cudaGetLastError();
}
int k3g(int dt, const void* spec, void* out, const RowTile* tiles, int n_tiles, const RowGroup* groups,
        const int* items, const void* tw, const void* win, int teams, int seg_w, const ApplyGeom& g, int batch,
        cudaStream_t s) {
  return dt == DT_F32 ? k3g_t<float>(spec, out, tiles, n_tiles, groups, items, tw, win, teams, seg_w, g, batch, s)
                      : k3g_t<double>(spec, out, tiles, n_tiles, groups, items, tw, win, teams, seg_w, g, batch, s);
}

template <typename T>
int k3s_t(const void* spec, void* out, const StreamTask* tasks, const unsigned* codes, int n_warp_items, const void* tw,
          const void* win, const ApplyGeom& g, int batch, int sm_count, const OutMirrors* mirrors, cudaStream_t s) {
  using ST = Stream<P, T>;
  const long long items = (long long)n_warp_items * batch;
  if (items == 0) return 0;
  const long long ctas = (items + ST::WARPS - 1) / ST::WARPS;
  const unsigned grid = (unsigned)(ctas < sm_count ? ctas : sm_count);
  if (mirrors && mirrors->n > 0)
    return launch_chain(4, k3_stream<P, T, true>, dim3(grid), dim3(ST::THREADS), ST::SMEM, s, (const cplx<T>*)spec, (T*)out,
                        tasks, codes, n_warp_items, (const cplx<T>*)tw, (const T*)win, g, batch, *mirrors);
  return launch_chain(4, k3_stream<P, T, false>, dim3(grid), dim3(ST::THREADS), ST::SMEM, s, (const cplx<T>*)spec, (T*)out,
                      tasks, codes, n_warp_items, (const cplx<T>*)tw, (const T*)win, g, batch, OutMirrors{});
}
int k3s(int dt, const void* spec, void* out, const StreamTask* tasks, const unsigned* codes, int n_warp_items,
        const void* tw, const void* win, const ApplyGeom& g, int batch, int sm_count, const OutMirrors* mirrors,
        cudaStream_t s) {
  return dt == DT_F32 ? k3s_t<float>(spec, out, tasks, codes, n_warp_items, tw, win, g, batch, sm_count, mirrors, s)
                      : k3s_t<double>(spec, out, tasks, codes, n_warp_items, tw, win, g, batch, sm_count, mirrors, s);
}
template <typename T>
int k2c_t(const void* spec, void* paired, const void* kmain, const void* knyq, const int* active, const ChainDesc* chains,
          int n_segments, const int* patches, const void* tw, const void* win, int batch, int n_active, long long bands_total,
          cudaStream_t s) {
  if constexpr (use_chain<T>()) {
    const long long ctas = (long long)n_segments * (TL::NTILE / TL::SLOTS) * batch;
    if (ctas == 0) return 0;
    k2_chain<P, T><<<(unsigned)ctas, TL::K2_THREADS, chain_smem<T>(), s>>>(
        (const cplx<T>*)spec, (cplx<T>*)paired, (const cplx<T>*)kmain, (const cplx<T>*)knyq, active, chains, patches,
        (const cplx<T>*)tw, (const T*)win, batch, n_active, bands_total);
    return (int)cudaGetLastError();
  } else {
    return (int)cudaErrorInvalidValue;
  }
}
int k2c(int dt, const void* spec, void* paired, const void* kmain, const void* knyq, const int* active, const ChainDesc* chains,
        int n_segments, const int* patches, const void* tw, const void* win, int batch, int n_active, long long bands_total,
        cudaStream_t s) {
  return dt == DT_F32 ? k2c_t<float>(spec, paired, kmain, knyq, active, chains, n_segments, patches, tw, win, batch, n_active, bands_total, s)
                      : k2c_t<double>(spec, paired, kmain, knyq, active, chains, n_segments, patches, tw, win, batch, n_active, bands_total, s);
}
template <typename T>
int k3p_t(const void* paired, void* out, const StreamTask* tasks, const unsigned* codes, int n_warp_items, const void* tw,
          const void* win, const ApplyGeom& g, int batch, int sm_count, const OutMirrors* mirrors, long long bands_total,
          cudaStream_t s) {
  if constexpr (use_chain<T>()) {
    using ST = Stream<P, T>;
    const long long items = (long long)n_warp_items * batch;
    if (items == 0) return 0;
    const long long ctas = (items + ST::WARPS - 1) / ST::WARPS;
    const unsigned grid = (unsigned)(ctas < sm_count ? ctas : sm_count);
    const long long ipf = bands_total * (P / 4);
    if (mirrors && mirrors->n > 0)
      k3_stream_paired<P, T, true><<<grid, ST::THREADS, ST::SMEM, s>>>((const cplx<T>*)paired, (T*)out, tasks, codes, n_warp_items,
                                                                       (const cplx<T>*)tw, (const T*)win, g, batch, *mirrors, ipf);
    else
      k3_stream_paired<P, T, false><<<grid, ST::THREADS, ST::SMEM, s>>>((const cplx<T>*)paired, (T*)out, tasks, codes, n_warp_items,
                                                                        (const cplx<T>*)tw, (const T*)win, g, batch, OutMirrors{}, ipf);
    return (int)cudaGetLastError();
  } else {
    return (int)cudaErrorInvalidValue;
  }
}
int k3p(int dt, const void* paired, void* out, const StreamTask* tasks, const unsigned* codes, int n_warp_items, const void* tw,
        const void* win, const ApplyGeom& g, int batch, int sm_count, const OutMirrors* mirrors, long long bands_total,
        cudaStream_t s) {
  return dt == DT_F32 ? k3p_t<float>(paired, out, tasks, codes, n_warp_items, tw, win, g, batch, sm_count, mirrors, bands_total, s)
                      : k3p_t<double>(paired, out, tasks, codes, n_warp_items, tw, win, g, batch, sm_count, mirrors, bands_total, s);
}
template <typename T>
int small_t(const void* image, void* planes, void* out, const int2* corners, const int* active, const void* kmain,
            const void* knyq, const void* tw, const void* win, const SmallTile* tiles, int n_tiles, const int* tile_patches,
            int max_cover, int tile_size, const ApplyGeom& g_in, const ApplyGeom& g_out, int batch, int bulk_ok, cudaStream_t s) {
  if constexpr (Small<P, T>::OK) {
    if (g_in.n_active > 0) {
      small_patch<P, T><<<(unsigned)(g_in.n_active * batch), Small<P, T>::THREADS, Small<P, T>::SMEM, s>>>(
          (const T*)image, (T*)planes, corners, active, (const cplx<T>*)kmain, (const cplx<T>*)knyq, (const cplx<T>*)tw,
          (const T*)win, g_in, batch, bulk_ok);
      int e = (int)cudaGetLastError();
      if (e) return e;
    }
    if (n_tiles > 0) {
      small_overlap_add<P, T><<<dim3((unsigned)n_tiles, (unsigned)batch), 256, 0, s>>>(
          (const T*)planes, (T*)out, tiles, tile_patches, max_cover, corners, tile_size, g_out);
    }
    return (int)cudaGetLastError();
  } else {
    return (int)cudaErrorInvalidValue;
  }
}
int small(int dt, const void* image, void* planes, void* out, const int2* corners, const int* active, const void* kmain,
          const void* knyq, const void* tw, const void* win, const SmallTile* tiles, int n_tiles, const int* tile_patches,
          int max_cover, int tile_size, const ApplyGeom& g_in, const ApplyGeom& g_out, int batch, int bulk_ok, cudaStream_t s) {
  return dt == DT_F32 ? small_t<float>(image, planes, out, corners, active, kmain, knyq, tw, win, tiles, n_tiles, tile_patches,
                                       max_cover, tile_size, g_in, g_out, batch, bulk_ok, s)
                      : small_t<double>(image, planes, out, corners, active, kmain, knyq, tw, win, tiles, n_tiles, tile_patches,
                                        max_cover, tile_size, g_in, g_out, batch, bulk_ok, s);
}
int small_ok(int dt) { return dt == DT_F32 ? (Small<P, float>::OK ? 1 : 0) : (Small<P, double>::OK ? 1 : 0); }
int chain_ok(int dt) { return dt == DT_F32 ? (use_chain<float>() ? 1 : 0) : (use_chain<double>() ? 1 : 0); }
int stream_tpw() { return Stream<P, float>::TPW; }

void fused_info(int dt, int info[4]) {
  info[0] = (dt == DT_F32 && Fused<P, float>::OK) ? 1 : 0;
  info[1] = Stream<P, float>::IPP;
  info[2] = TL::NTILE / TL::SLOTS;
  info[3] = Stream<P, float>::WARPS * Stream<P, float>::STAGES;     // issued ahead + the item published one iteration late
}
int fused(int dt, const void* image, void* ring, void* out, const int2* corners, const int* active, const void* kmain,
          const void* knyq, const StreamTask* tasks, const unsigned* codes, int n_warp_items, const void* tw,
          const void* win, const ApplyGeom& g_in, const ApplyGeom& g_out, int batch, int bulk_ok, const FusedGeom& fg,
          cudaStream_t s) {
  if constexpr (Fused<P, float>::OK) {
    if (dt != DT_F32) return (int)cudaErrorInvalidValue;
    using T = float;
    const T* image_t = (const T*)image; cplx<T>* ring_t = (cplx<T>*)ring; T* out_t = (T*)out;
    const cplx<T>* kmain_t = (const cplx<T>*)kmain; const cplx<T>* knyq_t = (const cplx<T>*)knyq;
    const cplx<T>* tw_t = (const cplx<T>*)tw; const T* win_t = (const T*)win;
    ApplyGeom gi = g_in, go = g_out; FusedGeom f = fg;
    void* args[] = {&image_t, &ring_t, &out_t, &corners, &active, &kmain_t, &knyq_t, &tasks, &codes, &n_warp_items,
                    &tw_t, &win_t, &gi, &go, &batch, &bulk_ok, &f};
    // cooperative: every CTA of the three roles must be resident at once (they wait on each other's counters)
    return (int)cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(fused_apply<P, T>), dim3(fg.n1 + fg.n2 + fg.n3),
                                            dim3(Fused<P, T>::THREADS), args, Fused<P, T>::SMEM, s);
  } else {
    return (int)cudaErrorInvalidValue;
  }
}

template <typename T, typename TK>
int prep_t(const void* full, void* kmain, void* knyq, int n, cudaStream_t s) {
  const long long total = (long long)n * ((long long)P * TL::HALF + P);
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long cap = (long long)sms * 32;
  const unsigned blocks = (unsigned)((total + 255) / 256 < cap ? (total + 255) / 256 : cap);
  prep_transfer_kernel<P, T, TK><<<blocks ? blocks : 1, 256, 0, s>>>((const TK*)full, (cplx<T>*)kmain, (cplx<T>*)knyq, n);
  return (int)cudaGetLastError();
}
int prep(int dt, int kdt, const void* full, void* kmain, void* knyq, int n, cudaStream_t s) {
  if (dt == DT_F32) return kdt == DT_F32 ? prep_t<float, float2>(full, kmain, knyq, n, s)
                                         : prep_t<float, double2>(full, kmain, knyq, n, s);
  return kdt == DT_F32 ? prep_t<double, float2>(full, kmain, knyq, n, s)
                       : prep_t<double, double2>(full, kmain, knyq, n, s);
}

template <typename T>
int fft2_t(const void* values, void* out, const void* tw, long long n, cudaStream_t s) {
  if (n == 0) return 0;
  fft2_rows<P, T, T><<<cdiv(n * P, TL::TEAMS), TL::ROW_THREADS, fft2_row_smem<T>(), s>>>(
      (const T*)values, (cplx<T>*)out, (const cplx<T>*)tw, n * P);
  int e = (int)cudaGetLastError();
  if (e) return e;
  fft2_cols<P, T><<<cdiv(n * (P / TL::C), TL::SLOTS), TL::K2_THREADS, col_smem<T>(), s>>>(
      (cplx<T>*)out, (const cplx<T>*)tw, n);
  return (int)cudaGetLastError();
}
int fft2(int dt, int in_dt, const void* values, void* out, const void* tw, long long n, cudaStream_t s) {
  if (dt != in_dt) return (int)cudaErrorInvalidValue;
  return dt == DT_F32 ? fft2_t<float>(values, out, tw, n, s) : fft2_t<double>(values, out, tw, n, s);
}

const Ops kOps = {P, init, k1, k1s, k2, k3, k3g, k3s, k2c, k3p, chain_ok, small_ok, small, stream_tpw, fused_info, fused, k3g_smem, prep, fft2};

}  // namespace

#define RPSF_CAT2(a, b) a##b
#define RPSF_CAT(a, b) RPSF_CAT2(a, b)
const Ops* RPSF_CAT(ops_p, RPSF_P)() { return &kOps; }

}  // namespace rpsf
