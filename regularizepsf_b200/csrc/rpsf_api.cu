// rpsf_api.cu — C ABI (include/rpsf_b200.h) over the sm_100a kernels.
// Host-side planning: coordinate validation, colour classes of the patch overlap graph,
// per-colour work lists, resident-row range, workspace ownership.
#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/rpsf_b200.h"
#include "rpsf_ops.h"
#include "rpsf_saturation.cuh"
#include "rpsf_builder.cuh"

namespace rpsf {
const Ops* ops_for(int P) {
  switch (P) {
    case 16: return ops_p16();
    case 32: return ops_p32();
    case 64: return ops_p64();
    case 128: return ops_p128();
    case 256: return ops_p256();
    case 512: return ops_p512();
    default: return nullptr;
  }
}
}  // namespace rpsf

using namespace rpsf;

namespace {

thread_local std::string g_error;
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}

#define CU(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(RPSF_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define LAUNCH(expr)                                                                               \
  do {                                                                                             \
    int e_ = (expr);                                                                               \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                            \
    if (e_ != 0)                                                                                   \
      return fail(RPSF_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString((cudaError_t)e_));       \
  } while (0)

size_t real_size(int dt) { return dt == DT_F32 ? 4 : 8; }
size_t elem_size(int dt) {
  switch (dt) {
    case RPSF_F32: case RPSF_I32: case RPSF_U32: return 4;
    case RPSF_F64: case RPSF_I64: return 8;
    case RPSF_U8: return 1;
    case RPSF_I16: case RPSF_U16: return 2;
    default: return 0;
  }
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// FFT length for a patch size without a native one: the next power of two >= 2 P - 1 (0 if none fits 16..512)
int embedded_length(int P) {
  if (P < 2) return 0;
  int m = 16;
  while (m < 2 * P - 1) m *= 2;
  return m <= 512 ? m : 0;
}

int sm_count_of(int device) {
  static std::mutex mu;
  static std::map<int, int> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(device);
  if (it != cache.end()) return it->second;
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n < 1) n = 148;
  cache[device] = n;
  return n;
}
int current_sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  return sm_count_of(dev);
}

template <typename T>
int upload_tables(int P, void** tw_out, void** win_out) {
  const int N1 = P == 16 ? 4 : P == 32 ? 4 : P == 64 ? 8 : P == 128 ? 8 : 16;
  const int N2 = P / N1;
  std::vector<T> tw(2 * (size_t)P), win(P);
  const double two_pi = 6.283185307179586476925286766559;
  for (int k2 = 0; k2 < N2; ++k2)
    for (int n1 = 0; n1 < N1; ++n1) {
      // exp(-2*pi*i*n1*k2/P); reduce the product mod P first so the angle stays exact
      const int r = (n1 * k2) % P;
      const double ang = two_pi * double(r) / double(P);
      tw[2 * (k2 * N1 + n1)] = T(std::cos(ang));
      tw[2 * (k2 * N1 + n1) + 1] = T(-std::sin(ang));
    }
  // apodization window, transform.py:151-154: sin((n + 0.5) * (pi / P)) per axis
  const double pi = 3.14159265358979323846264338327950288;
  for (int n = 0; n < P; ++n) win[n] = T(std::sin((n + 0.5) * (pi / P)));
  CU(cudaMalloc(tw_out, tw.size() * sizeof(T)));
  CU(cudaMalloc(win_out, win.size() * sizeof(T)));
  CU(cudaMemcpy(*tw_out, tw.data(), tw.size() * sizeof(T), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(*win_out, win.data(), win.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// Twiddle / window tables are a pure function of (P, dtype): one copy per device for the life of the
// library (3 KB at P = 256), shared by every transform and by rpsf_psf_fft2 — nothing to free, no
// synchronisation when a call returns.
int shared_tables(int P, int dtype, int device, void** tw, void** win) {
  struct Entry { void* tw; void* win; };
  static std::mutex mu;
  static std::map<std::tuple<int, int, int>, Entry> cache;
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_tuple(P, dtype, device);
  auto it = cache.find(key);
  if (it == cache.end()) {
    Entry e{nullptr, nullptr};
    int rc = dtype == RPSF_F32 ? upload_tables<float>(P, &e.tw, &e.win) : upload_tables<double>(P, &e.tw, &e.win);
    if (rc) { cudaFree(e.tw); cudaFree(e.win); return rc; }
    it = cache.emplace(key, e).first;
  }
  *tw = it->second.tw; *win = it->second.win;
  return 0;
}

}  // namespace

namespace {
template <typename TI, typename TO>
void launch_convert(const void* src, long long sp, void* dst, long long dp, int rows, int cols, cudaStream_t s) {
  const long long total = (long long)rows * cols;
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)current_sm_count() * 16);
  convert_2d<TI, TO><<<blocks ? blocks : 1, 256, 0, s>>>((const TI*)src, sp, (TO*)dst, dp, rows, cols);
}
template <typename TO>
int convert_to(const void* src, int sdt, long long sp, void* dst, long long dp, int rows, int cols, cudaStream_t s) {
  switch (sdt) {
    case RPSF_F32: launch_convert<float, TO>(src, sp, dst, dp, rows, cols, s); break;
    case RPSF_F64: launch_convert<double, TO>(src, sp, dst, dp, rows, cols, s); break;
    case RPSF_U8: launch_convert<uint8_t, TO>(src, sp, dst, dp, rows, cols, s); break;
    case RPSF_I16: launch_convert<int16_t, TO>(src, sp, dst, dp, rows, cols, s); break;
    case RPSF_U16: launch_convert<uint16_t, TO>(src, sp, dst, dp, rows, cols, s); break;
    case RPSF_I32: launch_convert<int32_t, TO>(src, sp, dst, dp, rows, cols, s); break;
    case RPSF_U32: launch_convert<uint32_t, TO>(src, sp, dst, dp, rows, cols, s); break;
    case RPSF_I64: launch_convert<long long, TO>(src, sp, dst, dp, rows, cols, s); break;
    default: return -1;
  }
  return 0;
}
}  // namespace

namespace {
// Tables of the streaming overlap-add (k3_stream): per owned row pair the (patch, row pair) items are
// grouped by corner column (colour order inside a group); the kernel needs the groups of a row pair
// to form one chain cx, cx + P/2, cx + P, ... that covers [0, W).  Chains are cut into `chunks`
// pieces for parallelism; teams of one warp must walk identically shaped tasks, so tasks are packed
// `tpw` at a time and padded with idle tasks where the shape changes.
struct StreamTables {
  std::vector<StreamTask> tasks;
  std::vector<unsigned> codes;
  int n_warp_items = 0;
  bool ok = false;
};

// `slots` (fused pipeline only): per active patch (band, index in band).  Then tasks are ordered row-pair major
// (a warp item = the same chunk of `tpw` consecutive row pairs), chains are cut into `fused_chunks` pieces, item
// codes name ring slots (FUSED_*_SHIFT) and every task carries the first / last band it reads in pad0 / pad1.
StreamTables build_stream_tables(const std::vector<int2>& corners, const std::vector<int>& colours, int P, int W,
                                 int row_begin, int row_end, int tpw, long long team_slots, int max_batch,
                                 const std::vector<int2>* slots = nullptr, int fused_chunks = 0,
                                 const std::vector<int>* paired_band = nullptr) {
  // `paired_band` (paired column pass, k2_chain): per active patch the band of the paired workspace that holds its
  // UPPER half (the lower half is in the next band).  A group then is one item, code = band * P/4 + row pair in band.
  StreamTables st;
  const int half = P / 2;
  if (corners.empty() || row_end <= row_begin) return st;
  const int parity = ((corners[0].x % 2) + 2) % 2;
  for (const int2& c : corners)
    if (((c.x % 2) + 2) % 2 != parity) return st;
  struct Entry { int item, colour, cx; };
  const int y_start = row_begin - ((((row_begin - parity) % 2) + 2) % 2);
  const int n_rp = (row_end - y_start + 1) / 2;
  std::vector<std::vector<Entry>> bucket(n_rp);
  for (size_t a = 0; a < corners.size(); ++a) {
    const int2 c = corners[a];
    if (c.y + P <= 0 || c.y >= W) continue;
    for (int pair = 0; pair < half; ++pair) {
      const int y = c.x + 2 * pair;
      if (y + 1 < row_begin || y >= row_end) continue;
      bucket[(y - y_start) / 2].push_back({(int)a * half + pair, colours[a], c.y});
    }
  }
  struct Group { int cx; std::vector<int> items; };
  std::vector<std::vector<Group>> chains(n_rp);
  size_t min_groups = SIZE_MAX;
  for (int rp = 0; rp < n_rp; ++rp) {
    auto& b = bucket[rp];
    if (b.empty()) return st;
    std::stable_sort(b.begin(), b.end(), [](const Entry& u, const Entry& v) {
      return u.cx != v.cx ? u.cx < v.cx : u.colour < v.colour;
    });
    auto& ch = chains[rp];
    for (const Entry& e : b) {
      if (ch.empty() || ch.back().cx != e.cx) ch.push_back({e.cx, {}});
      ch.back().items.push_back(e.item);
    }
    for (size_t i = 1; i < ch.size(); ++i)
      if (ch[i].cx != ch[i - 1].cx + half) return st;
    if (ch.front().cx > 0 || ch.back().cx + P < W) return st;
    min_groups = std::min(min_groups, ch.size());
  }
  // chunks per chain.  A chunk that does not start its chain recomputes one seam group, so c chunks cost a team
  // ceil(groups / c) + 1 transforms per task, and the launch takes ceil(tasks / teams) rounds of them: pick the c
  // that minimises the product (small launches want many short tasks, large ones whole chains).
  long long chunks = 1;
  const long long whole = (long long)n_rp * std::max(max_batch, 1);
  if (const char* v = getenv("RPSF_K3_CHUNKS")) {
    chunks = std::max(1, std::min(atoi(v), (int)min_groups));
  } else if (team_slots > 0) {
    long long best = LLONG_MAX;
    for (long long c = 1; c <= (long long)min_groups; ++c) {
      const long long per_task = ((long long)min_groups + c - 1) / c + (c > 1 ? 1 : 0);
      const long long rounds = (whole * c + team_slots - 1) / team_slots;
      if (per_task * rounds < best) { best = per_task * rounds; chunks = c; }
    }
  }
  std::vector<std::vector<int>> shape;        // per task: items per computed group
  if (slots) chunks = std::max<long long>(1, std::min<long long>(fused_chunks, (long long)min_groups));
  // visiting order of (chunk, row pair): chunk major for the stand-alone kernel; for the fused pipeline blocks of
  // `tpw` row pairs, chunk by chunk, so that work becomes ready top to bottom
  std::vector<std::pair<long long, int>> visit;
  if (!slots) {
    for (long long chn = 0; chn < chunks; ++chn)
      for (int rp = 0; rp < n_rp; ++rp) visit.push_back({chn, rp});
  } else {
    for (int rp0 = 0; rp0 < n_rp; rp0 += tpw)
      for (long long chn = 0; chn < chunks; ++chn)
        for (int rp = rp0; rp < std::min(n_rp, rp0 + tpw); ++rp) visit.push_back({chn, rp});
  }
  for (const auto& vis : visit) {
    {
      const long long chn = vis.first;
      const int rp = vis.second;
      const auto& ch = chains[rp];
      const size_t n_g = ch.size();
      const size_t gb = n_g * chn / chunks, ge = n_g * (chn + 1) / chunks;
      if (gb >= ge) continue;
      const size_t g0 = gb > 0 ? gb - 1 : 0;
      StreamTask t{};
      t.y = y_start + 2 * rp;
      t.cx0 = ch[g0].cx;
      t.item_begin = (int)st.codes.size();
      t.flags = (gb > 0 ? TASK_SEAM : 0) | (ge == n_g ? TASK_LAST : 0);
      std::vector<int> sh;
      if (slots) {
        int lo = INT_MAX, hi = INT_MIN;
        for (size_t gi = g0; gi < ge; ++gi)
          for (int item : ch[gi].items) {
            const int band = (*slots)[item / half].x;
            lo = std::min(lo, band); hi = std::max(hi, band);
          }
        if (hi - lo > 7) return StreamTables{};                // the item code has 3 bits for the band offset
        t.pad0 = lo; t.pad1 = hi;
      }
      for (size_t gi = g0; gi < ge; ++gi) {
        if (paired_band) {
          const int item = ch[gi].items[0], a = item / half, pair = item % half;
          const bool upper = 2 * pair < half;
          const unsigned code = (unsigned)(((*paired_band)[a] + (upper ? 0 : 1)) * (half / 2) + (upper ? pair : pair - half / 2));
          st.codes.push_back(code | ITEM_LAST_OF_GROUP);
          sh.push_back(1);
          continue;
        }
        for (size_t k = 0; k < ch[gi].items.size(); ++k) {
          unsigned code = (unsigned)ch[gi].items[k];
          if (slots) {
            const int2 sl = (*slots)[ch[gi].items[k] / half];
            code = (unsigned)(ch[gi].items[k] % half) | ((unsigned)sl.y << FUSED_IDX_SHIFT) |
                   ((unsigned)(sl.x - t.pad0) << FUSED_BAND_SHIFT);
          }
          st.codes.push_back(code | (k + 1 == ch[gi].items.size() ? ITEM_LAST_OF_GROUP : 0u));
        }
        sh.push_back((int)ch[gi].items.size());
      }
      t.n_steps = (int)st.codes.size() - t.item_begin;
      // pack: a warp item holds `tpw` tasks of one shape
      const size_t fill = st.tasks.size() % tpw;
      if (fill != 0 && shape.back() != sh)
        for (size_t k = fill; k < (size_t)tpw; ++k) { st.tasks.push_back(StreamTask{}); shape.push_back(sh); }
      st.tasks.push_back(t);
      shape.push_back(sh);
    }
  }
  while (st.tasks.size() % tpw) st.tasks.push_back(StreamTask{});
  st.n_warp_items = (int)(st.tasks.size() / tpw);
  st.ok = st.n_warp_items > 0;
  return st;
}
}  // namespace

struct rpsf_transform {
  int device = 0, P = 0, n = 0, dtype = 0;
  const Ops* ops = nullptr;
  std::vector<int2> corners;       // (row, col)
  std::vector<int> colour;         // greedy colouring in list order
  int n_colours = 0;
  void* tw = nullptr;
  void* win = nullptr;
  void* kmain = nullptr;           // private kernel layout of the KEPT patches, in list order
  void* knyq = nullptr;
  std::vector<int> kslot;          // patch -> index into kmain / knyq, or -1 (no kernel on this device: row-slab shards)
  int n_kept = 0;
  // patch sizes without a native FFT length run embedded in P = the next power of two >= 2 * win_len - 1
  // (embed_transfer_kernel); win_len == P otherwise
  int win_len = 0;
  bool own_win = false;            // the window table belongs to this transform (embedded sizes), not to the shared cache
  bool has_kernel = false;
  int sm_count = 148;
};

struct rpsf_plan {
  rpsf_transform* tr = nullptr;
  int H = 0, W = 0, pad_mode = 0, row_begin = 0, row_end = 0, max_batch = 1;
  int n_active = 0;
  int* active_dev = nullptr;       // active index -> kernel slot of the patch (rpsf_transform::kslot)
  int2* corners_dev = nullptr;     // per active patch
  std::vector<int*> items_dev;     // per colour: active*P/2 + pair
  std::vector<int> n_items;
  bool colour0_covers = false;
  // single-launch overlap-add (row-pair gather); falls back to colour phases when patch corner
  // rows do not share one parity or there are more than 15 colours
  bool gather = false;
  RowTile* tiles_dev = nullptr;
  RowGroup* groups_dev = nullptr;
  int* gitems_dev = nullptr;
  int n_tiles = 0, teams = 0, seg_w = 0;
  bool force_phases = false;   // test hook: run the colour-phase kernel even when gather is possible
  // kernel variants (test / A-B hooks; default = the persistent bulk-async kernels of rpsf_stream.cuh)
  bool k1_stream = true;
  bool k3_stream = true;       // allowed (RPSF_K3=old disables); used when stream_ok
  bool force_gather = false;   // test hook: the shared-memory row-pair gather even when chains exist
  bool stream_ok = false;      // every owned row pair is a chain of half-overlapping groups covering [0, W)
  StreamTask* stasks_dev = nullptr;
  unsigned* scodes_dev = nullptr;
  int n_warp_items = 0;
  void* workspace = nullptr;       // allocated at the first unfused apply (the fused pipeline only needs its ring)
  size_t workspace_bytes = 0;
  // small patches (rpsf_small.cuh): one CTA per patch-frame + overlap-add of the patch planes
  bool small_ok = false;
  int small_mode = 0;              // 0 = automatic (= the three kernels today), 1 = never, 2 = the single-CTA path
  SmallTile* small_tiles_dev = nullptr; int* small_cover_dev = nullptr; int small_n_tiles = 0, small_max_cover = 0;
  void* planes = nullptr; size_t planes_bytes = 0;
  // paired column pass (k2_chain + k3_stream_paired): chains of patches sharing a corner column, the workspace of
  // their band sums, and overlap-add tables with one item per group
  bool paired_ok = false;
  int column_mode = 0;             // 0 = automatic (= classic today), 1 = classic in-place column pass, 2 = paired
  ChainDesc* chains_dev = nullptr; int n_segments = 0;
  int* chain_patches_dev = nullptr;
  long long bands_total = 0;
  void* paired = nullptr; size_t paired_bytes = 0;
  StreamTask* ptasks_dev = nullptr; unsigned* pcodes_dev = nullptr; int p_warp_items = 0;
  // fused persistent pipeline (rpsf_fused.cuh): ring of band slots in place of the workspace, counters, its own
  // overlap-add tables (row-pair major, item codes that name ring slots)
  bool fused_ok = false;
  int fused_mode = 0;              // 0 = automatic (the three kernels: they are faster), 1 = never, 2 = the fused pipeline
  FusedGeom fg{};
  void* ring = nullptr; size_t ring_bytes = 0;
  uint64_t* fstats = nullptr; uint64_t* ftrace = nullptr;
  int* fcounters = nullptr; size_t fcounter_ints = 0;
  int2* fslot_dev = nullptr; int* fbands_dev = nullptr;   // fbands: [band_units | band_tasks]
  StreamTask* ftasks_dev = nullptr; unsigned* fcodes_dev = nullptr; int f_warp_items = 0;
  int img_lo = 0, img_hi = 0;      // resident frame rows needed: [img_lo, img_hi)
  // host path: a ring of HOST_SLOTS chunk buffers so the upload of chunk i+1, the kernels of
  // chunk i and the download of chunk i-1 run concurrently on three streams (two copy engines)
  static constexpr int HOST_SLOTS = 3;
  cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[HOST_SLOTS] = {}, ev_comp[HOST_SLOTS] = {}, ev_out[HOST_SLOTS] = {};
  void* d_in_raw[HOST_SLOTS] = {}; size_t d_in_raw_bytes = 0;
  void* h_stage[HOST_SLOTS] = {}; size_t h_stage_bytes = 0;   // pinned staging of pageable input chunks
  void* d_in[HOST_SLOTS] = {};
  void* d_out[HOST_SLOTS] = {};
  void* d_out_conv[HOST_SLOTS] = {}; size_t d_out_conv_bytes = 0;
  // saturation branch (transform.py:125-138,171-172); threshold = +inf disables it
  double sat_threshold = INFINITY;
  int sat_dilation = 1, sat_width = 7;
  void* sat_pf = nullptr;                 // [max_batch][Hp][Wp] padded frames, compute dtype
  unsigned char* sat_mask[2] = {nullptr, nullptr};
  int* sat_list = nullptr;                // [max_batch][Hp*Wp] raster-ordered masked pixels
  int* sat_rows = nullptr;                // [max_batch][Hp+1] row offsets, total last
  int* sat_flags = nullptr;               // [2*max_batch]: any-masked flags, fill tickets
  // fused slab gather: peer buffers K3 stores to as well (same layout as `out`)
  std::vector<void*> mirrors;
  // per-stage timing (bench only)
  bool timing = false;
  std::vector<cudaEvent_t> events;   // 4 per recorded apply call
  size_t events_used = 0;
};

extern "C" {

int rpsf_abi_version(void) { return RPSF_ABI_VERSION; }
const char* rpsf_last_error(void) { return g_error.c_str(); }
int rpsf_patch_size_supported(int P) { return ops_for(P) ? 1 : (embedded_length(P) && ops_for(embedded_length(P))) ? 2 : 0; }
int64_t rpsf_launch_count(void) { return g_launches.load(); }
int rpsf_pad_index(int i, int n, int pad_mode) { return n > 0 ? pad_index(i, n, pad_mode) : -1; }

int rpsf_transform_create_subset(rpsf_transform** out, const int32_t* coords, int n, int P, int dtype, int device,
                                 const uint8_t* keep) {
  if (!out || (!coords && n > 0) || n < 0) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (dtype != RPSF_F32 && dtype != RPSF_F64) return fail(RPSF_E_UNSUPPORTED, "compute dtype must be f32 or f64");
  const int win_len = P;
  const Ops* ops = ops_for(P);
  if (!ops) {
    P = embedded_length(win_len);                  // run embedded in the next power of two >= 2 P - 1
    ops = P ? ops_for(P) : nullptr;
  }
  if (!ops)
    return fail(RPSF_E_UNSUPPORTED, "patch size %d has no device path (supported: powers of two 16..512 natively, any size "
                "from 2 to 256 embedded in the next power of two >= 2 P - 1)", win_len);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(RPSF_E_CUDA, "cannot select CUDA device %d", device);
  int e = ops->init();
  if (e) return fail(RPSF_E_CUDA, "kernel attribute setup failed: %s", cudaGetErrorString((cudaError_t)e));
  auto* t = new rpsf_transform;
  t->device = device; t->P = P; t->n = n; t->dtype = dtype; t->ops = ops; t->win_len = win_len;
  t->sm_count = sm_count_of(device);
  t->corners.resize(n);
  for (int i = 0; i < n; ++i) t->corners[i] = make_int2(coords[2 * i], coords[2 * i + 1]);
  // Greedy colouring in list order: same-colour patches are pairwise disjoint.  For
  // calculate_covering (util.py:27-53) this recovers its four grids, in its order.
  t->colour.assign(n, 0);
  {
    std::vector<std::vector<int>> members;
    for (int i = 0; i < n; ++i) {
      int c = 0;
      for (;; ++c) {
        if (c == (int)members.size()) { members.emplace_back(); break; }
        bool clash = false;
        for (int j : members[c]) {
          // (the footprint of a patch is its true size: an embedded transform's FFT length P is larger)
          if (std::abs(t->corners[i].x - t->corners[j].x) < win_len && std::abs(t->corners[i].y - t->corners[j].y) < win_len) {
            clash = true; break;
          }
        }
        if (!clash) break;
      }
      members[c].push_back(i);
      t->colour[i] = c;
    }
    t->n_colours = (int)members.size();
  }
  int rc = shared_tables(P, dtype, device, &t->tw, &t->win);     // owned by the library, not by the transform
  if (rc) { delete t; return rc; }
  if (win_len != P) {
    // the window of the true patch size, zero beyond it (transform.py:151-154 with P = win_len)
    const double pi = 3.14159265358979323846264338327950288;
    std::vector<double> w64(P, 0.0);
    for (int i = 0; i < win_len; ++i) w64[i] = std::sin((i + 0.5) * (pi / win_len));
    const size_t rsz = real_size(dtype);
    void* wdev = nullptr;
    if (cudaMalloc(&wdev, P * rsz) != cudaSuccess) { delete t; return fail(RPSF_E_CUDA, "out of device memory for the window table"); }
    cudaError_t ce;
    if (dtype == RPSF_F32) {
      std::vector<float> w32(w64.begin(), w64.end());
      ce = cudaMemcpy(wdev, w32.data(), P * rsz, cudaMemcpyHostToDevice);
    } else {
      ce = cudaMemcpy(wdev, w64.data(), P * rsz, cudaMemcpyHostToDevice);
    }
    if (ce != cudaSuccess) { cudaFree(wdev); delete t; return fail(RPSF_E_CUDA, "window upload failed: %s", cudaGetErrorString(ce)); }
    t->win = wdev; t->own_win = true;
  }
  const size_t cs = 2 * real_size(dtype);
  t->kslot.assign(n, -1);
  for (int i = 0; i < n; ++i)
    if (!keep || keep[i]) t->kslot[i] = t->n_kept++;
  if (t->n_kept > 0) {
    if (cudaMalloc(&t->kmain, (size_t)t->n_kept * P * (P / 2) * cs) != cudaSuccess ||
        cudaMalloc(&t->knyq, (size_t)t->n_kept * P * cs) != cudaSuccess) {
      rpsf_transform_destroy(t);
      return fail(RPSF_E_CUDA, "out of device memory for the transfer kernel (%d patches of %d)", t->n_kept, P);
    }
  }
  *out = t;
  return RPSF_OK;
}

int rpsf_transform_create(rpsf_transform** out, const int32_t* coords, int n, int P, int dtype, int device) {
  return rpsf_transform_create_subset(out, coords, n, P, dtype, device, nullptr);
}

int rpsf_transform_destroy(rpsf_transform* t) {
  if (!t) return RPSF_OK;
  DeviceGuard guard(t->device);
  cudaFree(t->kmain); cudaFree(t->knyq);
  if (t->own_win) cudaFree(t->win);
  delete t;
  return RPSF_OK;
}

int rpsf_transform_num_colours(const rpsf_transform* t) { return t ? t->n_colours : 0; }

int rpsf_transform_set_kernel(rpsf_transform* t, const void* kernel_full, int kernel_dtype, void* stream) {
  if (!t || (!kernel_full && t->n_kept > 0)) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (kernel_dtype != RPSF_F32 && kernel_dtype != RPSF_F64)
    return fail(RPSF_E_UNSUPPORTED, "kernel dtype must be complex64 (RPSF_F32) or complex128 (RPSF_F64)");
  DeviceGuard guard(t->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (t->n_kept > 0 && t->win_len != t->P) {
    // embedded patch size: K (n, P, P) -> K' (n, M, M) = A K A^T in double (rpsf_kernels.cuh), then the usual layout
    const int P = t->win_len, M = t->P;
    std::vector<double> A(2 * (size_t)M * P);
    const double two_pi = 6.283185307179586476925286766559;
    for (int u = 0; u < M; ++u)
      for (int p = 0; p < P; ++p) {
        // (1/P) sum_{i=-(P-1)}^{P-1} exp(2 pi i (p/P - u/M) i), the phase reduced exactly: (p M - u P) i mod (P M)
        // = sin((2P-1) theta/2) / sin(theta/2) / P with theta/2 = pi r / (P M), r = (p M - u P) mod (P M); real, because the
        // terms i and -i are conjugates
        const long long den = (long long)P * M;
        long long r = ((long long)p * M - (long long)u * P) % den; if (r < 0) r += den;
        double re = double(2 * P - 1);
        if (r != 0) {
          const long long rn = ((2LL * P - 1) * r) % (2 * den);         // numerator angle, reduced mod 2 pi
          re = std::sin(0.5 * two_pi * double(rn) / double(den)) / std::sin(0.5 * two_pi * double(r) / double(den));
        }
        A[2 * ((size_t)u * P + p)] = re / P; A[2 * ((size_t)u * P + p) + 1] = 0.0;
      }
    double2 *dA = nullptr, *dT = nullptr, *dK = nullptr;
    auto release = [&]() { for (void* q : {(void*)dA, (void*)dT, (void*)dK}) if (q) cudaFreeAsync(q, s); };
    const size_t n = (size_t)t->n_kept;
    if (cudaMallocAsync((void**)&dA, sizeof(double2) * M * P, s) != cudaSuccess ||
        cudaMallocAsync((void**)&dT, sizeof(double2) * n * M * P, s) != cudaSuccess ||
        cudaMallocAsync((void**)&dK, sizeof(double2) * n * M * M, s) != cudaSuccess) {
      release(); cudaGetLastError();
      return fail(RPSF_E_CUDA, "out of device memory embedding %zu kernels of %d in %d", n, P, M);
    }
    // the host matrix is read by the copy before this function returns (pageable source: staged synchronously)
    CU(cudaMemcpyAsync(dA, A.data(), sizeof(double2) * M * P, cudaMemcpyHostToDevice, s));
    const unsigned blocks = (unsigned)t->sm_count * 16;
    if (kernel_dtype == RPSF_F32) embed_pass1<float2><<<blocks, 256, 0, s>>>((const float2*)kernel_full, dA, dT, (int)n, P, M);
    else embed_pass1<double2><<<blocks, 256, 0, s>>>((const double2*)kernel_full, dA, dT, (int)n, P, M);
    LAUNCH((int)cudaGetLastError());
    embed_pass2<double2><<<blocks, 256, 0, s>>>(dT, dA, dK, (int)n, P, M);
    LAUNCH((int)cudaGetLastError());
    int e = t->ops->prep(t->dtype, RPSF_F64, dK, t->kmain, t->knyq, t->n_kept, s);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    release();
    if (e) return fail(RPSF_E_CUDA, "kernel layout launch failed: %s", cudaGetErrorString((cudaError_t)e));
    t->has_kernel = true;
    return RPSF_OK;
  }
  if (t->n_kept > 0) LAUNCH(t->ops->prep(t->dtype, kernel_dtype, kernel_full, t->kmain, t->knyq, t->n_kept, s));
  t->has_kernel = true;
  return RPSF_OK;
}

int rpsf_construct_kernel(const void* S, const void* Tg, void* K, int64_t count, int dtype, double alpha,
                          double epsilon, int device, void* stream) {
  if (count < 0 || (count > 0 && (!S || !Tg || !K))) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (dtype != RPSF_F32 && dtype != RPSF_F64) return fail(RPSF_E_UNSUPPORTED, "dtype must be f32 or f64");
  if (count == 0) return RPSF_OK;
  DeviceGuard guard(device);
  const unsigned blocks = (unsigned)std::min<long long>((count + 255) / 256, (long long)sm_count_of(device) * 16);
  if (dtype == RPSF_F32)
    construct_transfer_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (const float2*)S, (const float2*)Tg, (float2*)K, count, (float)alpha, (float)epsilon);
  else
    construct_transfer_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (const double2*)S, (const double2*)Tg, (double2*)K, count, alpha, epsilon);
  LAUNCH((int)cudaGetLastError());
  return RPSF_OK;
}

int rpsf_psf_fft2(const void* values, void* out, int64_t n, int P, int dtype, int device, void* stream) {
  if (n < 0 || (n > 0 && (!values || !out))) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (dtype != RPSF_F32 && dtype != RPSF_F64) return fail(RPSF_E_UNSUPPORTED, "dtype must be f32 or f64");
  const Ops* ops = ops_for(P);
  if (!ops && (P < 1 || P > 512)) return fail(RPSF_E_UNSUPPORTED, "patch size %d has no device path", P);
  if (n == 0) return RPSF_OK;
  DeviceGuard guard(device);
  if (!ops) {
    // no native FFT length: separable direct DFT in double (setup cost only, O(P^3) per patch)
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<double> tw(2 * (size_t)P);
    const double two_pi = 6.283185307179586476925286766559;
    for (int m = 0; m < P; ++m) { tw[2 * m] = std::cos(two_pi * m / P); tw[2 * m + 1] = -std::sin(two_pi * m / P); }
    double2 *d_tw = nullptr, *d_tmp = nullptr;
    if (cudaMallocAsync((void**)&d_tw, sizeof(double2) * P, s) != cudaSuccess ||
        cudaMallocAsync((void**)&d_tmp, sizeof(double2) * (size_t)n * P * P, s) != cudaSuccess) {
      if (d_tw) cudaFreeAsync(d_tw, s);
      cudaGetLastError();
      return fail(RPSF_E_CUDA, "out of device memory for the direct DFT of %lld patches of %d", (long long)n, P);
    }
    CU(cudaMemcpyAsync(d_tw, tw.data(), sizeof(double2) * P, cudaMemcpyHostToDevice, s));
    const unsigned blocks = (unsigned)sm_count_of(device) * 16;
    if (dtype == RPSF_F32) dft_rows_direct<float><<<blocks, 256, 0, s>>>((const float*)values, d_tw, d_tmp, n * P, P);
    else dft_rows_direct<double><<<blocks, 256, 0, s>>>((const double*)values, d_tw, d_tmp, n * P, P);
    LAUNCH((int)cudaGetLastError());
    if (dtype == RPSF_F32) dft_cols_direct<float2><<<blocks, 256, 0, s>>>(d_tmp, d_tw, (float2*)out, n, P);
    else dft_cols_direct<double2><<<blocks, 256, 0, s>>>(d_tmp, d_tw, (double2*)out, n, P);
    LAUNCH((int)cudaGetLastError());
    cudaFreeAsync(d_tw, s); cudaFreeAsync(d_tmp, s);
    return RPSF_OK;
  }
  int e = ops->init();
  if (e) return fail(RPSF_E_CUDA, "kernel attribute setup failed: %s", cudaGetErrorString((cudaError_t)e));
  void* tw = nullptr; void* win = nullptr;
  int rc = shared_tables(P, dtype, device, &tw, &win);           // cached per (P, dtype, device): no sync, no free
  if (rc) return rc;
  e = ops->fft2(dtype, dtype, values, out, tw, n, (cudaStream_t)stream);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  if (e) return fail(RPSF_E_CUDA, "fft2 launch failed: %s", cudaGetErrorString((cudaError_t)e));
  return RPSF_OK;
}

int rpsf_average_patches(const double* cutouts, int64_t n_cutouts, int P, const int64_t* cell_offsets,
                         const int32_t* cell_items, int64_t n_cells, int method, double percentile, double* out,
                         int device, void* stream) {
  if (P <= 0 || n_cutouts < 0 || n_cells < 0) return fail(RPSF_E_INVALID_ARGUMENT, "negative size");
  if (n_cells == 0) return RPSF_OK;
  if (!cell_offsets || !out) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (method != RPSF_AVG_MEAN && method != RPSF_AVG_MEDIAN && method != RPSF_AVG_PERCENTILE)
    return fail(RPSF_E_UNSUPPORTED, "unknown averaging method %d", method);
  if (method == RPSF_AVG_PERCENTILE && !(percentile >= 0.0 && percentile <= 100.0))
    return fail(RPSF_E_INVALID_ARGUMENT, "percentile must be in [0, 100]");
  if (method == RPSF_AVG_PERCENTILE && percentile == 50.0) method = RPSF_AVG_MEDIAN;      // builder.py:79-82
  const long long total = cell_offsets[n_cells];
  if (cell_offsets[0] != 0 || total < 0 || (total > 0 && (!cell_items || !cutouts)))
    return fail(RPSF_E_INVALID_ARGUMENT, "bad cell index arrays");
  for (int64_t c = 0; c < n_cells; ++c)
    if (cell_offsets[c + 1] < cell_offsets[c] || cell_offsets[c + 1] - cell_offsets[c] > INT_MAX)
      return fail(RPSF_E_INVALID_ARGUMENT, "cell offsets must be non-decreasing");
  for (long long i = 0; i < total; ++i)
    if (cell_items[i] < 0 || cell_items[i] >= n_cutouts)
      return fail(RPSF_E_INVALID_ARGUMENT, "cell item %lld refers to cutout %d of %lld", i, cell_items[i], (long long)n_cutouts);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(RPSF_E_CUDA, "cannot select CUDA device %d", device);
  cudaStream_t s = (cudaStream_t)stream;
  const int pp = P * P, centre = (P / 2) * P + P / 2;
  const unsigned gx = (unsigned)((pp + AVG_TPB - 1) / AVG_TPB);

  // launch lists: staged classes by stack depth, then the stacks too deep for shared memory
  const int caps[] = {16, AVG_SHALLOW, 64, 128, 256, AVG_MAX_STAGED};
  constexpr int NCLASS = 7;
  std::vector<int> lists[NCLASS];
  std::vector<long long> big_off;
  long long scratch_elems = 0;
  if (method != RPSF_AVG_MEAN)
    for (int64_t c = 0; c < n_cells; ++c) {
      const long long n = cell_offsets[c + 1] - cell_offsets[c];
      int k = 0;
      while (k < NCLASS - 1 && n > caps[k]) ++k;
      lists[k].push_back((int)c);
      if (k == NCLASS - 1) { big_off.push_back(scratch_elems); scratch_elems += (long long)gx * AvgLayout<AVG_SUBS>::STRIDE * n; }
    }
  std::vector<int> cells_flat;
  for (auto& l : lists) cells_flat.insert(cells_flat.end(), l.begin(), l.end());

  long long* d_off = nullptr; int* d_items = nullptr; int* d_cells = nullptr;
  long long* d_big = nullptr; unsigned long long* d_scratch = nullptr;
  // stream-ordered temporaries: freed behind the kernels, no host synchronisation
  auto release = [&]() {
    for (void* ptr : {(void*)d_off, (void*)d_items, (void*)d_cells, (void*)d_big, (void*)d_scratch})
      if (ptr) cudaFreeAsync(ptr, s);
  };
#define AVG_CU(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { release(); \
    return fail(RPSF_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); } } while (0)
  static_assert(sizeof(long long) == sizeof(int64_t), "offset width");
  AVG_CU(cudaMallocAsync(&d_off, sizeof(long long) * (size_t)(n_cells + 1), s));
  AVG_CU(cudaMemcpyAsync(d_off, cell_offsets, sizeof(long long) * (size_t)(n_cells + 1), cudaMemcpyHostToDevice, s));
  if (total > 0) {
    AVG_CU(cudaMallocAsync(&d_items, sizeof(int) * (size_t)total, s));
    AVG_CU(cudaMemcpyAsync(d_items, cell_items, sizeof(int) * (size_t)total, cudaMemcpyHostToDevice, s));
  }
  if (!cells_flat.empty()) {
    AVG_CU(cudaMallocAsync(&d_cells, sizeof(int) * cells_flat.size(), s));
    AVG_CU(cudaMemcpyAsync(d_cells, cells_flat.data(), sizeof(int) * cells_flat.size(), cudaMemcpyHostToDevice, s));
  }
  if (!big_off.empty()) {
    AVG_CU(cudaMallocAsync(&d_big, sizeof(long long) * big_off.size(), s));
    AVG_CU(cudaMemcpyAsync(d_big, big_off.data(), sizeof(long long) * big_off.size(), cudaMemcpyHostToDevice, s));
    AVG_CU(cudaMallocAsync(&d_scratch, sizeof(unsigned long long) * (size_t)scratch_elems, s));
  }
  const double quantile = percentile / 100.0;             // np.true_divide(q, 100)
  if (method == RPSF_AVG_MEAN) {
    const unsigned mx = (unsigned)((pp + 255) / 256);
    for (int64_t c0 = 0; c0 < n_cells; c0 += 65535) {
      const unsigned gy = (unsigned)std::min<int64_t>(65535, n_cells - c0);
      average_mean<<<dim3(mx, gy), 256, 0, s>>>(cutouts, d_off + c0, d_items, pp, centre, out + c0 * pp);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      AVG_CU(cudaGetLastError());
    }
  } else {
    using Deep = AvgLayout<AVG_SUBS>;
    AVG_CU(cudaFuncSetAttribute(reinterpret_cast<const void*>(average_select<true, AVG_SUBS>),
                                cudaFuncAttributeMaxDynamicSharedMemorySize, AVG_MAX_STAGED * Deep::STRIDE * 8));
    size_t first = 0;
    for (int k = 0; k < NCLASS; ++k) {
      const size_t count = lists[k].size();
      for (size_t c0 = 0; c0 < count; c0 += 65535) {
        const unsigned gy = (unsigned)std::min<size_t>(65535, count - c0);
        const int* cl = d_cells + first + c0;
        if (k < NCLASS - 1 && caps[k] <= AVG_SHALLOW)
          average_select<true, 1><<<dim3(gx, gy), AvgLayout<1>::THREADS, (size_t)caps[k] * AvgLayout<1>::STRIDE * 8, s>>>(
              cutouts, d_off, d_items, cl, pp, centre, method, quantile, nullptr, nullptr, out);
        else if (k < NCLASS - 1)
          average_select<true, AVG_SUBS><<<dim3(gx, gy), Deep::THREADS, (size_t)caps[k] * Deep::STRIDE * 8, s>>>(
              cutouts, d_off, d_items, cl, pp, centre, method, quantile, nullptr, nullptr, out);
        else
          average_select<false, AVG_SUBS><<<dim3(gx, gy), Deep::THREADS, 0, s>>>(
              cutouts, d_off, d_items, cl, pp, centre, method, quantile, d_scratch, d_big + c0, out);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        AVG_CU(cudaGetLastError());
      }
      first += count;
    }
  }
#undef AVG_CU
  release();
  return RPSF_OK;
}

namespace {
// the two per-patch builder stages share their argument checks and the stream-ordered byte scratch
int patch_stage(const double* in, double* out, int64_t n, int P, int device, void* stream, bool isolate) {
  if (P <= 0 || n < 0) return fail(RPSF_E_INVALID_ARGUMENT, "negative size");
  if (n == 0) return RPSF_OK;
  if (!out || (!isolate && !in)) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (n > INT_MAX || (long long)P * P > INT_MAX / 2) return fail(RPSF_E_UNSUPPORTED, "too many patches or patch too large");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(RPSF_E_CUDA, "cannot select CUDA device %d", device);
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* scratch = nullptr;
  const size_t bytes = (size_t)n * P * P * (isolate ? 2 : 1);
  cudaError_t e = cudaMallocAsync(&scratch, bytes, s);
  if (e != cudaSuccess) return fail(RPSF_E_CUDA, "cudaMallocAsync of %zu scratch bytes failed: %s", bytes, cudaGetErrorString(e));
  if (isolate) isolate_cores<<<(unsigned)n, ISO_TPB, 0, s>>>(out, scratch, P);
  else plane_background<<<(unsigned)n, ISO_TPB, 0, s>>>(in, out, scratch, P);
  e = cudaGetLastError();
  cudaFreeAsync(scratch, s);
  if (e != cudaSuccess) return fail(RPSF_E_CUDA, "launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return RPSF_OK;
}
}  // namespace

int rpsf_star_cutouts(const void* frame, int frame_dtype, int H, int W, const double* corners, int64_t n_stars, int width,
                      double saturation_threshold, double star_minimum, double star_maximum, double* out,
                      unsigned char* accepted, int device, void* stream) {
  if (H <= 0 || W <= 0 || width <= 0 || n_stars < 0) return fail(RPSF_E_INVALID_ARGUMENT, "negative size");
  if (n_stars == 0) return RPSF_OK;
  if (!frame || !corners || !out || !accepted) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (frame_dtype != RPSF_F32 && frame_dtype != RPSF_F64) return fail(RPSF_E_UNSUPPORTED, "frames must be float32 or float64");
  if (n_stars > INT_MAX || (long long)width * width > INT_MAX / 2) return fail(RPSF_E_UNSUPPORTED, "too many stars or cutout too large");
  for (int64_t i = 0; i < 2 * n_stars; ++i)
    if (!(std::fabs(corners[i]) < 1e9)) return fail(RPSF_E_INVALID_COORDINATE, "star corner %lld is not a finite pixel position", (long long)(i / 2));
  DeviceGuard guard(device);
  if (!guard.ok) return fail(RPSF_E_CUDA, "cannot select CUDA device %d", device);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t pp = (size_t)width * width;
  double* d_corners = nullptr; double* d_coeffs = nullptr; unsigned char* d_scratch = nullptr;
  auto release = [&]() {
    for (void* ptr : {(void*)d_corners, (void*)d_coeffs, (void*)d_scratch})
      if (ptr) cudaFreeAsync(ptr, s);
  };
#define CUT_CU(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { release(); \
    return fail(RPSF_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); } } while (0)
  CUT_CU(cudaMallocAsync(&d_corners, sizeof(double) * 2 * (size_t)n_stars, s));
  CUT_CU(cudaMemcpyAsync(d_corners, corners, sizeof(double) * 2 * (size_t)n_stars, cudaMemcpyHostToDevice, s));
  CUT_CU(cudaMallocAsync(&d_coeffs, sizeof(double) * pp * (size_t)n_stars, s));
  CUT_CU(cudaMallocAsync(&d_scratch, pp * (size_t)n_stars, s));
  if (frame_dtype == RPSF_F32)
    star_cutouts<float><<<(unsigned)n_stars, ISO_TPB, 0, s>>>((const float*)frame, H, W, d_corners, width, saturation_threshold,
                                                               star_minimum, star_maximum, out, accepted, d_coeffs, d_scratch);
  else
    star_cutouts<double><<<(unsigned)n_stars, ISO_TPB, 0, s>>>((const double*)frame, H, W, d_corners, width, saturation_threshold,
                                                                star_minimum, star_maximum, out, accepted, d_coeffs, d_scratch);
  CUT_CU(cudaGetLastError());
#undef CUT_CU
  g_launches.fetch_add(1, std::memory_order_relaxed);
  // the corner list is pageable host memory: its copy must have left the caller's buffer before we return
  cudaError_t e = cudaStreamSynchronize(s);
  release();
  if (e != cudaSuccess) return fail(RPSF_E_CUDA, "star cutouts failed: %s", cudaGetErrorString(e));
  return RPSF_OK;
}

int rpsf_plane_background(const double* patches, int64_t n_patches, int P, double* out, int device, void* stream) {
  return patch_stage(patches, out, n_patches, P, device, stream, false);
}

int rpsf_isolate_cores(double* patches, int64_t n_patches, int P, int device, void* stream) {
  return patch_stage(nullptr, patches, n_patches, P, device, stream, true);
}

int rpsf_plan_set_output_mirrors(rpsf_plan* p, int n, void* const* ptrs) {
  if (!p || n < 0 || n > 7 || (n > 0 && !ptrs)) return fail(RPSF_E_INVALID_ARGUMENT, "0..7 mirror buffers");
  p->mirrors.assign(ptrs, ptrs + n);
  return RPSF_OK;
}

int rpsf_ipc_alloc(void** ptr, int64_t bytes, int device, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!ptr || !handle || bytes <= 0) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(RPSF_E_CUDA, "cannot select CUDA device %d", device);
  CU(cudaMalloc(ptr, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) { cudaFree(*ptr); *ptr = nullptr; return fail(RPSF_E_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); }
  memcpy(handle, &h, sizeof h);
  return RPSF_OK;
}

int rpsf_ipc_open(void** ptr, const unsigned char handle[64], int device) {
  if (!ptr || !handle) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(RPSF_E_CUDA, "cannot select CUDA device %d", device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  CU(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return RPSF_OK;
}

int rpsf_ipc_close(void* ptr, int device) {
  if (!ptr) return RPSF_OK;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(RPSF_E_CUDA, "cannot select CUDA device %d", device);
  CU(cudaDeviceSynchronize());
  CU(cudaIpcCloseMemHandle(ptr));
  return RPSF_OK;
}

int rpsf_plan_create(rpsf_plan** out, rpsf_transform* t, int H, int W, int pad_mode, int row_begin, int row_end,
                     int max_batch) {
  // RPSF_PLAN_TIMING=1 prints where the planning time goes (host-side work lists and their uploads)
  const bool plan_timing = getenv("RPSF_PLAN_TIMING") != nullptr;
  auto plan_t0 = std::chrono::steady_clock::now();
  auto tick = [&](const char* what) {
    if (!plan_timing) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[rpsf plan] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - plan_t0).count());
    plan_t0 = now;
  };
  if (!out || !t) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (H <= 0 || W <= 0) return fail(RPSF_E_INCORRECT_SHAPE, "frame shape must be positive, got (%d, %d)", H, W);
  if (pad_mode < 0 || pad_mode > RPSF_PAD_MATERIALIZED) return fail(RPSF_E_UNSUPPORTED, "unknown pad mode %d", pad_mode);
  static_assert(RPSF_PAD_MATERIALIZED == PAD_NONE, "the materialised pad is the kernels' PAD_NONE");
  if (row_begin < 0 || row_end > H || row_begin > row_end)
    return fail(RPSF_E_INVALID_ARGUMENT, "row band [%d,%d) outside frame of %d rows", row_begin, row_end, H);
  if (max_batch < 1) return fail(RPSF_E_INVALID_ARGUMENT, "max_batch must be >= 1");
  const int P = t->P;
  // The reference pads 2P per side and slices [c+2P, c+3P) (transform.py:119-123,141-149); a
  // corner outside [-2P, dim+P] makes that slice short and numpy raises.  Reject it up front.
  // (P here is the true patch size; an embedded transform's FFT length t->P is larger.)
  for (int i = 0; i < t->n; ++i) {
    const int2 c = t->corners[i];
    const int P = t->win_len;
    if (c.x < -2 * P || c.x + P > H + 2 * P || c.y < -2 * P || c.y + P > W + 2 * P)
      return fail(RPSF_E_INVALID_COORDINATE,
                  "patch corner (%d, %d) lies outside the 2*P padded frame of shape (%d, %d)", c.x, c.y, H, W);
  }
  DeviceGuard guard(t->device);
  auto* p = new rpsf_plan;
  p->tr = t; p->H = H; p->W = W; p->pad_mode = pad_mode; p->row_begin = row_begin; p->row_end = row_end;
  p->max_batch = max_batch;
  if (const char* v = getenv("RPSF_K1")) p->k1_stream = strcmp(v, "old") != 0;
  if (const char* v = getenv("RPSF_K3")) p->k3_stream = strcmp(v, "old") != 0;
  const bool embedded = t->win_len != t->P;
  if (embedded) { p->k1_stream = false; p->force_phases = true; }      // the two kernels that honour ApplyGeom::win_len
  std::vector<int> active;
  std::vector<int2> corners;
  std::vector<std::vector<int>> items(std::max(t->n_colours, 1));
  long long colour0_area = 0;
  int lo = H, hi = 0;
  // active patches in (corner row, corner column) order: patches that share a corner row are consecutive (the
  // bands of the fused pipeline), and the row kernels walk the frame top to bottom
  std::vector<int> order;
  for (int i = 0; i < t->n; ++i) {
    const int2 c = t->corners[i];
    if (std::max(c.x, row_begin) >= std::min(c.x + t->win_len, row_end) || std::max(c.y, 0) >= std::min(c.y + t->win_len, W)) continue;
    order.push_back(i);                                       // contributes to the owned band
  }
  std::stable_sort(order.begin(), order.end(), [&](int u, int v) {
    const int2 cu = t->corners[u], cv = t->corners[v];
    return cu.x != cv.x ? cu.x < cv.x : cu.y < cv.y;
  });
  for (int i : order) {
    const int2 c = t->corners[i];
    const int r0 = std::max(c.x, row_begin), r1 = std::min(c.x + t->win_len, row_end);
    const int c0 = std::max(c.y, 0), c1 = std::min(c.y + t->win_len, W);
    const int a = (int)active.size();
    if (t->kslot[i] < 0) {
      rpsf_plan_destroy(p);
      return fail(RPSF_E_INVALID_ARGUMENT,
                  "patch %d at (%d, %d) contributes to rows [%d, %d) but this transform was created without its kernel", i, c.x,
                  c.y, row_begin, row_end);
    }
    active.push_back(i);
    corners.push_back(c);
    for (int r = 0; r < t->win_len; ++r) {
      const int y = pad_index(c.x + r, H, pad_mode);
      if (y >= 0) { lo = std::min(lo, y); hi = std::max(hi, y + 1); }
    }
    for (int pair = 0; pair < (t->win_len + 1) / 2; ++pair) {         // row pairs past the window contribute nothing
      const int ya = c.x + 2 * pair, yb = ya + 1;
      if ((ya >= row_begin && ya < row_end) || (yb >= row_begin && yb < row_end))
        items[t->colour[i]].push_back(a * (P / 2) + pair);
    }
    if (t->colour[i] == 0) colour0_area += (long long)(r1 - r0) * (c1 - c0);
  }
  if (lo >= hi) { lo = 0; hi = 0; }
  p->img_lo = lo; p->img_hi = hi;
  p->n_active = (int)active.size();
  p->colour0_covers = !embedded && colour0_area == (long long)(row_end - row_begin) * W;
  auto destroy_fail = [&](const char* what) {
    rpsf_plan_destroy(p);
    return fail(RPSF_E_CUDA, "out of device memory for %s", what);
  };
  auto upload_fail = [&]() {
    const cudaError_t e = cudaGetLastError();
    rpsf_plan_destroy(p);
    return fail(RPSF_E_CUDA, "plan upload failed: %s", cudaGetErrorString(e));
  };
  if (p->n_active > 0) {
    if (cudaMalloc(&p->active_dev, sizeof(int) * active.size()) != cudaSuccess) return destroy_fail("patch list");
    if (cudaMalloc(&p->corners_dev, sizeof(int2) * corners.size()) != cudaSuccess) return destroy_fail("corner list");
    std::vector<int> slots_of_active(active.size());          // what the column kernel needs of a patch: where its kernel lives
    for (size_t a = 0; a < active.size(); ++a) slots_of_active[a] = t->kslot[active[a]];
    if (cudaMemcpy(p->active_dev, slots_of_active.data(), sizeof(int) * active.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
    if (cudaMemcpy(p->corners_dev, corners.data(), sizeof(int2) * corners.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
  }
  p->items_dev.assign(items.size(), nullptr);
  p->n_items.assign(items.size(), 0);
  for (size_t c = 0; c < items.size(); ++c) {
    p->n_items[c] = (int)items[c].size();
    if (items[c].empty()) continue;
    if (cudaMalloc(&p->items_dev[c], sizeof(int) * items[c].size()) != cudaSuccess) return destroy_fail("work list");
    if (cudaMemcpy(p->items_dev[c], items[c].data(), sizeof(int) * items[c].size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
  }
  tick("active list and colour items");
  // ---- row-pair gather tables: tiles -> groups (same corner column, summed in registers in colour
  // order) -> at most two layers of disjoint groups (one shared-memory plane each)
  {
    bool aligned = p->n_active > 0;
    const int parity = p->n_active ? ((corners[0].x % 2) + 2) % 2 : 0;
    for (const int2& c : corners) aligned = aligned && (((c.x % 2) + 2) % 2 == parity);
    if (aligned) {
      struct Entry { int item, colour, cx; };
      const int seg = std::min(W, 2048);
      const int n_seg = (W + seg - 1) / seg;
      const int y_start = row_begin - ((((row_begin - parity) % 2) + 2) % 2);
      const int n_rp = row_end > y_start ? (row_end - y_start + 1) / 2 : 0;
      std::vector<std::vector<Entry>> bucket((size_t)n_rp * n_seg);
      for (int a = 0; a < p->n_active; ++a) {
        const int2 c = corners[a];
        const int col = t->colour[active[a]];
        for (int pair = 0; pair < P / 2; ++pair) {
          const int y = c.x + 2 * pair;
          if (y + 1 < row_begin || y >= row_end) continue;
          const int rp = (y - y_start) / 2;
          for (int sgm = 0; sgm < n_seg; ++sgm) {
            const int x0 = sgm * seg, x1 = std::min(W, x0 + seg);
            if (c.y + P <= x0 || c.y >= x1) continue;
            bucket[(size_t)rp * n_seg + sgm].push_back({a * (P / 2) + pair, col, c.y});
          }
        }
      }
      std::vector<RowTile> tiles;
      std::vector<RowGroup> groups;
      std::vector<int> gitems;
      size_t most = 0;
      bool two_layers = true;
      for (int rp = 0; rp < n_rp && two_layers; ++rp)
        for (int sgm = 0; sgm < n_seg && two_layers; ++sgm) {
          auto& b = bucket[(size_t)rp * n_seg + sgm];
          std::stable_sort(b.begin(), b.end(), [](const Entry& u, const Entry& v) {
            return u.cx != v.cx ? u.cx < v.cx : u.colour < v.colour;
          });
          const int x0 = sgm * seg, x1 = std::min(W, x0 + seg);
          RowTile rt;
          rt.group_begin = (int)groups.size(); rt.group_count = 0;
          rt.y = y_start + 2 * rp; rt.x0 = x0;
          int layer_end[2] = {INT_MIN, INT_MIN};           // one past the last column a layer holds so far
          for (size_t i = 0; i < b.size();) {
            size_t j = i;
            while (j < b.size() && b[j].cx == b[i].cx) ++j;
            const int cx = b[i].cx;
            int layer = -1;
            for (int l = 0; l < 2; ++l)
              if (cx >= layer_end[l]) { layer = l; break; }
            if (layer < 0) { two_layers = false; break; }
            layer_end[layer] = cx + P;
            RowGroup gr;
            gr.item_begin = (int)gitems.size(); gr.item_count = (int)(j - i); gr.cx = cx; gr.layer = layer;
            gr.clipped = (cx < x0 || cx + P > x1) ? 1 : 0;
            for (size_t k = i; k < j; ++k) gitems.push_back(b[k].item);
            groups.push_back(gr);
            ++rt.group_count;
            i = j;
          }
          tiles.push_back(rt);                                  // a tile with no group still zero-fills its rows
          most = std::max(most, (size_t)rt.group_count);
        }
      if (two_layers && !tiles.empty()) {
        const int n1 = P == 16 || P == 32 ? 4 : P == 64 || P == 128 ? 8 : 16;
        int teams_max = std::min(288 / n1, 32);
        while (teams_max > 1 && t->ops->k3g_smem(t->dtype, teams_max, seg) > (size_t)K3G_SMEM_MAX) --teams_max;
        if (t->ops->k3g_smem(t->dtype, teams_max, seg) <= (size_t)K3G_SMEM_MAX) {
          const int rounds = most ? (int)((most + teams_max - 1) / teams_max) : 1;
          p->teams = most ? (int)((most + rounds - 1) / rounds) : 1;
          p->seg_w = seg;
          p->n_tiles = (int)tiles.size();
          if (cudaMalloc(&p->tiles_dev, sizeof(RowTile) * tiles.size()) != cudaSuccess) return destroy_fail("row tiles");
          if (cudaMemcpy(p->tiles_dev, tiles.data(), sizeof(RowTile) * tiles.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
          if (!groups.empty()) {
            if (cudaMalloc(&p->groups_dev, sizeof(RowGroup) * groups.size()) != cudaSuccess) return destroy_fail("row groups");
            if (cudaMemcpy(p->groups_dev, groups.data(), sizeof(RowGroup) * groups.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
            if (cudaMalloc(&p->gitems_dev, sizeof(int) * gitems.size()) != cudaSuccess) return destroy_fail("gather items");
            if (cudaMemcpy(p->gitems_dev, gitems.data(), sizeof(int) * gitems.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
          }
          p->gather = true;
        }
      }
    }
  }
  tick("row-pair gather tables");
  // ---- streaming overlap-add tables
  if (p->n_active > 0 && (long long)p->n_active * (P / 2) < (1LL << 30)) {
    std::vector<int> colours(active.size());
    for (size_t a = 0; a < active.size(); ++a) colours[a] = t->colour[active[a]];
    const int tpw = t->ops->stream_tpw();
    StreamTables stt = build_stream_tables(corners, colours, P, W, row_begin, row_end, tpw,
                                           (long long)t->sm_count * 16 * tpw, max_batch);
    if (stt.ok) {
      if (cudaMalloc(&p->stasks_dev, sizeof(StreamTask) * stt.tasks.size()) != cudaSuccess) return destroy_fail("stream tasks");
      if (cudaMemcpy(p->stasks_dev, stt.tasks.data(), sizeof(StreamTask) * stt.tasks.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
      if (cudaMalloc(&p->scodes_dev, sizeof(unsigned) * stt.codes.size()) != cudaSuccess) return destroy_fail("stream item codes");
      if (cudaMemcpy(p->scodes_dev, stt.codes.data(), sizeof(unsigned) * stt.codes.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
      p->n_warp_items = stt.n_warp_items;
      p->stream_ok = true;
    }
  }
  tick("streaming overlap-add tables");
  // ---- small patches: output tiles of P/2 x P/2 and the patches that cover each, in list order (the reference's `+=` order)
  if (const char* v = getenv("RPSF_SMALL")) p->small_mode = strcmp(v, "0") == 0 ? 1 : strcmp(v, "1") == 0 ? 2 : 0;
  if (!embedded && p->n_active > 0 && P % 2 == 0 && t->ops->small_ok(t->dtype) && row_end > row_begin) {
    const int ts = P / 2;
    bool aligned = true;
    for (const int2& c : corners) aligned = aligned && (((c.x % ts) + ts) % ts == 0) && (((c.y % ts) + ts) % ts == 0);
    if (aligned) {
      // every patch covers the 2 x 2 tiles under it (corners are tile-aligned): one pass over the patches
      const int ty0 = row_begin / ts, ty1 = (row_end + ts - 1) / ts, ntx = (W + ts - 1) / ts;
      std::vector<std::vector<std::pair<int, int>>> hits((size_t)(ty1 - ty0) * ntx);      // (list index, active index)
      for (int a = 0; a < p->n_active; ++a) {
        const int cy = (int)std::floor((double)corners[a].x / ts), cx = (int)std::floor((double)corners[a].y / ts);
        for (int dy = 0; dy < 2; ++dy)
          for (int dx = 0; dx < 2; ++dx) {
            const int ty = cy + dy, tx = cx + dx;
            if (ty >= ty0 && ty < ty1 && tx >= 0 && tx < ntx) hits[(size_t)(ty - ty0) * ntx + tx].push_back({active[a], a});
          }
      }
      std::vector<SmallTile> tiles;
      std::vector<std::vector<int>> cover;
      size_t most = 0;
      for (int ty = ty0; ty < ty1; ++ty)
        for (int tx = 0; tx < ntx; ++tx) {
          auto& h = hits[(size_t)(ty - ty0) * ntx + tx];
          std::sort(h.begin(), h.end());
          tiles.push_back(SmallTile{ty * ts, tx * ts});
          cover.emplace_back();
          for (auto& e : h) cover.back().push_back(e.second);
          most = std::max(most, h.size());
        }
      const int mc = (int)std::max<size_t>(most, 1);
      std::vector<int> flat(tiles.size() * (size_t)mc, -1);
      for (size_t i = 0; i < tiles.size(); ++i) std::copy(cover[i].begin(), cover[i].end(), flat.begin() + i * mc);
      if (cudaMalloc(&p->small_tiles_dev, sizeof(SmallTile) * tiles.size()) != cudaSuccess) return destroy_fail("output tiles");
      if (cudaMalloc(&p->small_cover_dev, sizeof(int) * flat.size()) != cudaSuccess) return destroy_fail("tile cover lists");
      if (cudaMemcpy(p->small_tiles_dev, tiles.data(), sizeof(SmallTile) * tiles.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
      if (cudaMemcpy(p->small_cover_dev, flat.data(), sizeof(int) * flat.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
      p->small_n_tiles = (int)tiles.size(); p->small_max_cover = mc;
      p->planes_bytes = (size_t)max_batch * p->n_active * P * P * real_size(t->dtype);
      p->small_ok = true;
    }
  }
  tick("single-CTA tile lists");
  // ---- paired column pass: chains of active patches that share a corner column, corner rows P/2 apart
  if (const char* v = getenv("RPSF_PAIRED")) p->column_mode = strcmp(v, "1") == 0 ? 2 : 0;
  if (p->stream_ok && !embedded && t->ops->chain_ok(t->dtype) && (P / 2) % 2 == 0) {
    std::map<int, std::vector<int>> by_col;                       // corner column -> active patches, by corner row
    for (int a = 0; a < p->n_active; ++a) by_col[corners[a].y].push_back(a);
    bool ok = true;
    std::vector<int> chain_patches, band_upper(p->n_active, 0);
    struct Chain { int first, length, band0; };
    std::vector<Chain> chains;
    long long bands = 0;
    size_t min_len = SIZE_MAX;
    for (auto& kv : by_col) {
      const auto& list = kv.second;
      for (size_t i = 1; i < list.size() && ok; ++i) ok = corners[list[i]].x == corners[list[i - 1]].x + P / 2;
      if (!ok) break;
      chains.push_back({(int)chain_patches.size(), (int)list.size(), (int)bands});
      for (size_t i = 0; i < list.size(); ++i) { band_upper[list[i]] = (int)(bands + (long long)i); chain_patches.push_back(list[i]); }
      bands += (long long)list.size() + 1;
      min_len = std::min(min_len, list.size());
    }
    ok = ok && !chains.empty() && bands * (P / 4) < (1LL << 30);
    StreamTables pt;
    if (ok) {
      std::vector<int> colours(active.size());
      for (size_t a = 0; a < active.size(); ++a) colours[a] = t->colour[active[a]];
      const int tpw = t->ops->stream_tpw();
      pt = build_stream_tables(corners, colours, P, W, row_begin, row_end, tpw, (long long)t->sm_count * 16 * tpw, max_batch,
                               nullptr, 0, &band_upper);
      ok = pt.ok;
    }
    if (ok) {
      // segments per chain: a segment that does not start its chain recomputes one patch for its carry; pick the count
      // that minimises (steps per segment) x (rounds of CTAs over the 2-per-SM slots)
      const int C = std::min(16, P / 2), n1 = P == 16 || P == 32 ? 4 : P == 64 || P == 128 ? 8 : 16;
      const int slots_per_cta = std::max(1, 256 / (C * n1)), tg = std::max(1, (P / 2) / C / slots_per_cta);
      const long long cta_slots = 2LL * t->sm_count;
      long long segs = 1, best = LLONG_MAX;
      for (long long sgm = 1; sgm <= (long long)min_len; ++sgm) {
        const long long steps = ((long long)min_len + sgm - 1) / sgm + (sgm > 1 ? 1 : 0);
        const long long rounds = ((long long)chains.size() * sgm * tg * max_batch + cta_slots - 1) / cta_slots;
        if (steps * rounds < best) { best = steps * rounds; segs = sgm; }
      }
      if (const char* v = getenv("RPSF_CHAIN_SEGMENTS")) segs = std::max(1, std::min(atoi(v), (int)min_len));
      std::vector<ChainDesc> descs;
      for (const Chain& c : chains)
        for (long long sgm = 0; sgm < segs; ++sgm) {
          const int b = (int)((long long)c.length * sgm / segs), e = (int)((long long)c.length * (sgm + 1) / segs);
          if (e > b) descs.push_back(ChainDesc{c.first, c.length, c.band0, b, e - b});
        }
      if (cudaMalloc(&p->chains_dev, sizeof(ChainDesc) * descs.size()) != cudaSuccess) return destroy_fail("chain table");
      if (cudaMalloc(&p->chain_patches_dev, sizeof(int) * chain_patches.size()) != cudaSuccess) return destroy_fail("chain patches");
      if (cudaMalloc(&p->ptasks_dev, sizeof(StreamTask) * pt.tasks.size()) != cudaSuccess) return destroy_fail("paired tasks");
      if (cudaMalloc(&p->pcodes_dev, sizeof(unsigned) * pt.codes.size()) != cudaSuccess) return destroy_fail("paired item codes");
      if (cudaMemcpy(p->chains_dev, descs.data(), sizeof(ChainDesc) * descs.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
      if (cudaMemcpy(p->chain_patches_dev, chain_patches.data(), sizeof(int) * chain_patches.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
      if (cudaMemcpy(p->ptasks_dev, pt.tasks.data(), sizeof(StreamTask) * pt.tasks.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
      if (cudaMemcpy(p->pcodes_dev, pt.codes.data(), sizeof(unsigned) * pt.codes.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
      p->n_segments = (int)descs.size();
      p->bands_total = bands;
      p->p_warp_items = pt.n_warp_items;
      p->paired_bytes = (size_t)max_batch * bands * (P / 2) * (P / 2) * 2 * real_size(t->dtype);
      p->paired_ok = true;
    }
  }
  p->workspace_bytes = (size_t)max_batch * p->n_active * P * (P / 2) * 2 * real_size(t->dtype);
  tick("paired column tables");
  // ---- fused persistent pipeline: bands = runs of equal corner row in the (sorted) active list
  int finfo[4] = {0, 0, 0, 0};
  t->ops->fused_info(t->dtype, finfo);
  // Measured (profiles/r02_fused_*.txt): the fused pipeline cuts DRAM traffic from 346 to 121 MB per 2048^2 frame but
  // is slower than the three stand-alone kernels, so it is opt-in: RPSF_FUSED=1 or rpsf_plan_set_fused(plan, 2).
  const char* fenv = getenv("RPSF_FUSED");
  if (fenv && strcmp(fenv, "1") == 0) p->fused_mode = 2;
  if (finfo[0] && p->stream_ok && !(fenv && strcmp(fenv, "0") == 0)) {
    std::vector<int2> slots(p->n_active);
    std::vector<int> band_size;
    for (int a = 0; a < p->n_active; ++a) {
      if (a == 0 || corners[a].x != corners[a - 1].x) band_size.push_back(0);
      slots[a] = make_int2((int)band_size.size() - 1, band_size.back()++);
    }
    const int n_bands = (int)band_size.size();
    const int band_cap = *std::max_element(band_size.begin(), band_size.end());
    const int min_band = *std::min_element(band_size.begin(), band_size.end());
    const size_t patch_bytes = (size_t)P * (P / 2) * 2 * real_size(t->dtype);
    int ring = 8, chunks = 4, split[3] = {0, 0, 0};
    if (const char* v = getenv("RPSF_FUSED_RING")) ring = atoi(v);
    if (const char* v = getenv("RPSF_FUSED_CHUNKS")) chunks = atoi(v);
    if (const char* v = getenv("RPSF_FUSED_SPLIT")) sscanf(v, "%d,%d,%d", &split[0], &split[1], &split[2]);
    if (split[0] <= 0 || split[1] <= 0 || split[2] <= 0 || split[0] + split[1] + split[2] > t->sm_count) {
      // shares of the three roles measured on the stand-alone kernels (about 0.30 : 0.42 : 0.28)
      split[0] = std::max(1, (int)std::lround(t->sm_count * 0.31));
      split[2] = std::max(1, (int)std::lround(t->sm_count * 0.27));
      split[1] = t->sm_count - split[0] - split[2];
    }
    // A K1 warp holds up to `stages` + 1 items it has not published yet while it asks for the slot of the next
    // one, and a group of items is published when its slowest warp is through: what the role holds that way must
    // stay inside the ring's slack, or a warp could wait for a slot that only its own CTA's unfinished items can free
    const long long ahead_items = (long long)(finfo[3] + finfo[3] / 2) * split[0];
    const long long ahead_bands = (ahead_items + (long long)min_band * finfo[1] - 1) / ((long long)min_band * finfo[1]) + 1;
    while (ring < ahead_bands + 2 && ring < 1024) ring *= 2;
    bool ok = (ring & (ring - 1)) == 0 && ring >= 4 && (size_t)ring * band_cap * patch_bytes <= (size_t)96 << 20 &&
              band_cap < (1 << (FUSED_BAND_SHIFT - FUSED_IDX_SHIFT)) && t->sm_count >= 12 && ahead_bands <= ring - 2;
    StreamTables ft;
    if (ok) {
      std::vector<int> colours(active.size());
      for (size_t a = 0; a < active.size(); ++a) colours[a] = t->colour[active[a]];
      ft = build_stream_tables(corners, colours, P, W, row_begin, row_end, t->ops->stream_tpw(), 0, max_batch, &slots,
                               std::max(chunks, 1));
      ok = ft.ok;
    }
    if (ok) {
      std::vector<int> bands(2 * (size_t)n_bands, 0);            // [band_units | band_tasks]
      for (int b = 0; b < n_bands; ++b) bands[b] = band_size[b] * finfo[2];
      int span = 0;
      for (const StreamTask& tk : ft.tasks)
        if (tk.n_steps > 0) {
          for (int b = tk.pad0; b <= tk.pad1; ++b) ++bands[n_bands + b];
          span = std::max(span, tk.pad1 - tk.pad0 + 1);
        }
      ok = span + 2 <= ring;                                      // a task's bands, the band in K2 and the band K1 fills
      for (int b = 0; b < n_bands && ok; ++b) ok = bands[n_bands + b] > 0;   // every band is read (else its slot never frees)
      if (ok) {
        p->ring_bytes = (size_t)ring * band_cap * patch_bytes;
        p->fcounter_ints = (size_t)max_batch * (p->n_active + 2 * (size_t)n_bands) + 4;
        if (cudaMalloc(&p->ring, p->ring_bytes) != cudaSuccess) return destroy_fail("spectrum ring");
        if (cudaMalloc(&p->fcounters, p->fcounter_ints * sizeof(int)) != cudaSuccess) return destroy_fail("pipeline counters");
        if (cudaMalloc(&p->fslot_dev, sizeof(int2) * slots.size()) != cudaSuccess) return destroy_fail("band slots");
        if (cudaMalloc(&p->fbands_dev, sizeof(int) * bands.size()) != cudaSuccess) return destroy_fail("band targets");
        if (cudaMalloc(&p->ftasks_dev, sizeof(StreamTask) * ft.tasks.size()) != cudaSuccess) return destroy_fail("fused tasks");
        if (cudaMalloc(&p->fcodes_dev, sizeof(unsigned) * ft.codes.size()) != cudaSuccess) return destroy_fail("fused item codes");
        if (cudaMemcpy(p->fslot_dev, slots.data(), sizeof(int2) * slots.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
        if (cudaMemcpy(p->fbands_dev, bands.data(), sizeof(int) * bands.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
        if (cudaMemcpy(p->ftasks_dev, ft.tasks.data(), sizeof(StreamTask) * ft.tasks.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
        if (cudaMemcpy(p->fcodes_dev, ft.codes.data(), sizeof(unsigned) * ft.codes.size(), cudaMemcpyHostToDevice) != cudaSuccess) return upload_fail();
        p->f_warp_items = ft.n_warp_items;
        FusedGeom& fg = p->fg;
        fg.n_bands = n_bands; fg.ring_mask = ring - 1; fg.band_cap = band_cap;
        fg.n1 = split[0]; fg.n2 = split[1]; fg.n3 = split[2];
        fg.units_per_patch = finfo[2]; fg.n_active = p->n_active; fg.stats = nullptr; fg.trace = nullptr; fg.trace_stride = 0;
        fg.solo = 0;
        if (const char* v = getenv("RPSF_FUSED_SOLO")) fg.solo = atoi(v);
        fg.slot = p->fslot_dev; fg.band_units = p->fbands_dev; fg.band_tasks = p->fbands_dev + n_bands;
        fg.ready1 = p->fcounters;
        fg.ready2 = fg.ready1 + (size_t)max_batch * p->n_active;
        fg.done3 = fg.ready2 + (size_t)max_batch * n_bands;
        fg.ticket = reinterpret_cast<unsigned*>(fg.done3 + (size_t)max_batch * n_bands);
        p->fused_ok = true;
      }
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { rpsf_plan_destroy(p); return fail(RPSF_E_CUDA, "plan upload failed: %s", cudaGetErrorString(e)); }
  tick("fused pipeline tables");
  *out = p;
  return RPSF_OK;
}

int rpsf_plan_destroy(rpsf_plan* p) {
  if (!p) return RPSF_OK;
  DeviceGuard guard(p->tr->device);
  for (cudaStream_t s : {p->s_in, p->s_comp, p->s_out})
    if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
  for (int i = 0; i < rpsf_plan::HOST_SLOTS; ++i) {
    if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
    if (p->ev_comp[i]) cudaEventDestroy(p->ev_comp[i]);
    if (p->ev_out[i]) cudaEventDestroy(p->ev_out[i]);
    cudaFree(p->d_in_raw[i]); cudaFree(p->d_in[i]); cudaFree(p->d_out[i]); cudaFree(p->d_out_conv[i]);
    if (p->h_stage[i]) cudaFreeHost(p->h_stage[i]);
  }
  cudaFree(p->active_dev); cudaFree(p->corners_dev); cudaFree(p->workspace);
  cudaFree(p->tiles_dev); cudaFree(p->groups_dev); cudaFree(p->gitems_dev);
  cudaFree(p->stasks_dev); cudaFree(p->scodes_dev);
  cudaFree(p->small_tiles_dev); cudaFree(p->small_cover_dev); cudaFree(p->planes);
  cudaFree(p->chains_dev); cudaFree(p->chain_patches_dev); cudaFree(p->paired); cudaFree(p->ptasks_dev); cudaFree(p->pcodes_dev);
  cudaFree(p->ring); cudaFree(p->fstats); cudaFree(p->ftrace); cudaFree(p->fcounters); cudaFree(p->fslot_dev); cudaFree(p->fbands_dev);
  cudaFree(p->ftasks_dev); cudaFree(p->fcodes_dev);
  cudaFree(p->sat_pf); cudaFree(p->sat_mask[0]); cudaFree(p->sat_mask[1]); cudaFree(p->sat_list);
  cudaFree(p->sat_rows); cudaFree(p->sat_flags);
  for (int* d : p->items_dev) cudaFree(d);
  for (cudaEvent_t e : p->events) cudaEventDestroy(e);
  delete p;
  return RPSF_OK;
}

int rpsf_plan_set_overlap_mode(rpsf_plan* p, int mode) {
  if (!p) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (p->tr->win_len != p->tr->P) return RPSF_OK;             // embedded patch sizes only have the colour-phase kernel
  p->force_phases = mode == 1;
  p->force_gather = mode == 2;
  return RPSF_OK;
}

int rpsf_plan_set_gather_mode(rpsf_plan* p, int mode) {
  if (!p) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (p->tr->win_len != p->tr->P) return RPSF_OK;             // embedded patch sizes only have the plain gather kernel
  p->k1_stream = mode != 1;
  return RPSF_OK;
}

int rpsf_plan_info(const rpsf_plan* p, int64_t info[8]) {
  if (!p || !info) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  const bool stream = p->stream_ok && p->k3_stream && !p->force_phases && !p->force_gather;
  info[6] = stream ? 2 : (p->gather && !p->force_phases) ? 1 : 0; info[7] = p->teams;
  if (stream && p->fused_ok && p->fused_mode == 2 && p->k1_stream) info[6] += 8;
  info[0] = p->n_active; info[1] = p->tr->n_colours; info[2] = (int64_t)p->workspace_bytes;
  info[3] = p->img_lo; info[4] = p->img_hi; info[5] = p->colour0_covers ? 1 : 0;
  return RPSF_OK;
}

static int ensure_workspace(rpsf_plan* p) {
  if (p->workspace || p->workspace_bytes == 0) return RPSF_OK;
  DeviceGuard guard(p->tr->device);
  if (cudaMalloc(&p->workspace, p->workspace_bytes) != cudaSuccess) {
    cudaGetLastError();
    return fail(RPSF_E_CUDA, "out of device memory for the spectrum workspace (%zu bytes)", p->workspace_bytes);
  }
  return RPSF_OK;
}

int rpsf_plan_workspace(const rpsf_plan* p, void** ptr, int64_t* bytes) {
  if (!p || !ptr || !bytes) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (int rc = ensure_workspace(const_cast<rpsf_plan*>(p))) return rc;
  *ptr = p->workspace; *bytes = (int64_t)p->workspace_bytes;
  return RPSF_OK;
}

int rpsf_plan_fused_stats(rpsf_plan* p, int enable, uint64_t out[8]) {
  if (!p) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (!p->fused_ok) return fail(RPSF_E_UNSUPPORTED, "this plan has no fused pipeline");
  DeviceGuard guard(p->tr->device);
  if (out) {
    for (int i = 0; i < 8; ++i) out[i] = 0;
    if (p->fstats) {
      CU(cudaDeviceSynchronize());
      CU(cudaMemcpy(out, p->fstats, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
      CU(cudaMemset(p->fstats, 0, 8 * sizeof(uint64_t)));
    }
  }
  if (enable && !p->fstats) {
    CU(cudaMalloc((void**)&p->fstats, 8 * sizeof(uint64_t)));
    CU(cudaMemset(p->fstats, 0, 8 * sizeof(uint64_t)));
  }
  p->fg.stats = enable ? reinterpret_cast<unsigned long long*>(p->fstats) : nullptr;
  return RPSF_OK;
}

int rpsf_plan_fused_trace(rpsf_plan* p, int enable, uint64_t* out, int64_t out_len) {
  if (!p) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (!p->fused_ok) return fail(RPSF_E_UNSUPPORTED, "this plan has no fused pipeline");
  DeviceGuard guard(p->tr->device);
  const size_t stride = (size_t)p->max_batch * p->fg.n_bands, n = 4 * stride;
  if (out) {
    if (!p->ftrace || out_len < (int64_t)(3 * stride)) return fail(RPSF_E_INVALID_ARGUMENT, "trace is off or the buffer is too small (%zu needed)", 3 * stride);
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out, p->ftrace, 3 * stride * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  }
  if (enable && !p->ftrace) CU(cudaMalloc((void**)&p->ftrace, n * sizeof(uint64_t)));
  if (p->ftrace) CU(cudaMemset(p->ftrace, 0, n * sizeof(uint64_t)));
  p->fg.trace = enable ? reinterpret_cast<unsigned long long*>(p->ftrace) : nullptr;
  p->fg.trace_stride = (int)stride;
  return RPSF_OK;
}

int rpsf_plan_set_small_mode(rpsf_plan* p, int mode) {
  if (!p) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (mode < 0 || mode > 2) return fail(RPSF_E_INVALID_ARGUMENT, "small-patch mode must be 0, 1 or 2");
  if (mode == 2 && !p->small_ok)
    return fail(RPSF_E_UNSUPPORTED, "this plan has no single-CTA path (needs patches of at most 128 px with corners on multiples of P/2)");
  p->small_mode = mode;
  return RPSF_OK;
}

int rpsf_plan_column_info(const rpsf_plan* p, int64_t info[2]) {
  if (!p || !info) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  const bool stream = p->stream_ok && p->k3_stream && !p->force_phases && !p->force_gather;
  const bool want = p->column_mode == 2;
  info[0] = (p->paired_ok && want && stream) ? 1 : 0;
  info[1] = p->paired_ok ? (int64_t)(p->paired_bytes / (size_t)std::max(p->max_batch, 1)) : 0;
  return RPSF_OK;
}

int rpsf_plan_set_column_mode(rpsf_plan* p, int mode) {
  if (!p) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (mode < 0 || mode > 2) return fail(RPSF_E_INVALID_ARGUMENT, "column mode must be 0, 1 or 2");
  if (mode == 2 && !p->paired_ok) return fail(RPSF_E_UNSUPPORTED, "this plan has no paired column pass (needs a covering)");
  p->column_mode = mode;
  return RPSF_OK;
}

int rpsf_plan_set_fused(rpsf_plan* p, int mode) {
  if (!p) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (mode < 0 || mode > 2) return fail(RPSF_E_INVALID_ARGUMENT, "pipeline mode must be 0, 1 or 2");
  if (mode == 2 && !p->fused_ok) return fail(RPSF_E_UNSUPPORTED, "this plan has no fused pipeline (needs a covering, 256-px patches, float32)");
  p->fused_mode = mode;
  return RPSF_OK;
}

int rpsf_upload(void** device_ptr, const void* src_host, int64_t bytes, int device) {
  if (!device_ptr || bytes < 0 || (bytes > 0 && !src_host)) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  DeviceGuard guard(device);
  *device_ptr = nullptr;
  if (bytes == 0) return RPSF_OK;
  CU(cudaMalloc(device_ptr, (size_t)bytes));
  cudaError_t e = cudaMemcpy(*device_ptr, src_host, (size_t)bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(*device_ptr); *device_ptr = nullptr;
    return fail(RPSF_E_CUDA, "upload failed: %s", cudaGetErrorString(e));
  }
  return RPSF_OK;
}

int rpsf_device_free(void* device_ptr, int device) {
  if (!device_ptr) return RPSF_OK;
  DeviceGuard guard(device);
  CU(cudaDeviceSynchronize());
  CU(cudaFree(device_ptr));
  return RPSF_OK;
}

int rpsf_copy_to_host(void* dst, const void* src, int64_t bytes, int device) {
  if (bytes < 0 || (bytes > 0 && (!dst || !src))) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  DeviceGuard guard(device);
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost));
  return RPSF_OK;
}

int rpsf_plan_set_saturation(rpsf_plan* p, double threshold, int dilation, int neighborhood_width) {
  if (!p) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (threshold != threshold) return fail(RPSF_E_INVALID_ARGUMENT, "saturation threshold is NaN");
  if (neighborhood_width < 0) return fail(RPSF_E_INVALID_ARGUMENT, "neighborhood_width must be >= 0");
  p->sat_threshold = threshold;
  p->sat_dilation = dilation;
  p->sat_width = neighborhood_width;
  return RPSF_OK;
}

}  // extern "C"

namespace {
// output mirrors are only offered without the saturation branch (checked again on the classic path)
inline bool sat_blocks_mirrors(const rpsf_plan* p, bool sat) { return sat && !p->mirrors.empty(); }

// Saturation pre-fill on the stream: pad + mask, dilate, compact in raster order, ordered fill.
// On return *filled points at padded pixel (2P, 2P) of frame 0, i.e. at unpadded (0, 0).
template <typename T>
int sat_prefill(rpsf_plan* p, const void* image, int64_t img_pitch, int64_t img_frame_stride, int batch,
                cudaStream_t s, SatGeom* sg_out, const void** filled) {
  rpsf_transform* t = p->tr;
  const int P = t->P, pad = 2 * P;
  SatGeom sg;
  sg.H = p->H; sg.W = p->W; sg.pad = pad; sg.Hp = p->H + 2 * pad; sg.Wp = p->W + 2 * pad;
  sg.half = p->sat_width / 2; sg.row_begin = p->row_begin; sg.row_end = p->row_end;
  sg.img_pitch = img_pitch; sg.img_frame_stride = img_frame_stride;
  sg.out_pitch = 0; sg.out_frame_stride = 0; sg.out_row0 = 0; sg.pad_mode = p->pad_mode;
  const size_t px = (size_t)sg.Hp * sg.Wp, mb = (size_t)p->max_batch;
  if ((long long)px > 0x7fffffffLL) return fail(RPSF_E_UNSUPPORTED, "padded frame too large for the saturation branch");
  if (!p->sat_pf) {
    CU(cudaMalloc(&p->sat_pf, px * sizeof(T) * mb));
    CU(cudaMalloc((void**)&p->sat_mask[0], px * mb));
    CU(cudaMalloc((void**)&p->sat_mask[1], px * mb));
    CU(cudaMalloc((void**)&p->sat_list, px * sizeof(int) * mb));
    CU(cudaMalloc((void**)&p->sat_rows, (size_t)(sg.Hp + 1) * sizeof(int) * mb));
    CU(cudaMalloc((void**)&p->sat_flags, 2 * sizeof(int) * mb));
  }
  CU(cudaMemsetAsync(p->sat_flags, 0, 2 * sizeof(int) * mb, s));
  int* any_flag = p->sat_flags;
  int* tickets = p->sat_flags + mb;
  const dim3 rows_grid((unsigned)std::min((sg.Wp + 255) / 256, 8), (unsigned)sg.Hp, (unsigned)batch);
  sat_pad_mask<T><<<rows_grid, 256, 0, s>>>((const T*)image, (T*)p->sat_pf, p->sat_mask[0], any_flag,
                                            p->sat_threshold, sg);
  LAUNCH((int)cudaGetLastError());
  int cur = 0;
  if (p->sat_dilation < 1) {
    sat_dilate<<<rows_grid, 256, 0, s>>>(p->sat_mask[0], p->sat_mask[1], any_flag, sg.Hp, sg.Wp, 1);
    LAUNCH((int)cudaGetLastError());
    cur = 1;
  } else {
    for (int it = 0; it < p->sat_dilation; ++it) {
      sat_dilate<<<rows_grid, 256, 0, s>>>(p->sat_mask[cur], p->sat_mask[cur ^ 1], any_flag, sg.Hp, sg.Wp, 0);
      LAUNCH((int)cudaGetLastError());
      cur ^= 1;
    }
  }
  unsigned char* state = p->sat_mask[cur];
  const dim3 warp_rows((unsigned)((sg.Hp + 7) / 8), (unsigned)batch);
  sat_row_count<<<warp_rows, 256, 0, s>>>(state, p->sat_rows, sg.Hp, sg.Wp);
  LAUNCH((int)cudaGetLastError());
  sat_row_scan<<<batch, 1024, 0, s>>>(p->sat_rows, sg.Hp);
  LAUNCH((int)cudaGetLastError());
  sat_row_scatter<<<warp_rows, 256, 0, s>>>(state, p->sat_rows, p->sat_list, sg.Hp, sg.Wp);
  LAUNCH((int)cudaGetLastError());
  // every CTA of the fill must be resident at once only for speed, not for correctness (tickets)
  sat_fill<T><<<dim3((unsigned)t->sm_count * 2, (unsigned)batch), 256, 0, s>>>((T*)p->sat_pf, state, p->sat_list, p->sat_rows, tickets, sg);
  LAUNCH((int)cudaGetLastError());
  *sg_out = sg;
  *filled = (const T*)p->sat_pf + (size_t)pad * sg.Wp + pad;
  return RPSF_OK;
}

template <typename T>
int sat_restore_launch(rpsf_plan* p, const void* image, void* out, int64_t out_pitch, int64_t out_frame_stride,
                       int out_row0, int batch, cudaStream_t s, SatGeom sg) {
  sg.out_pitch = out_pitch; sg.out_frame_stride = out_frame_stride; sg.out_row0 = out_row0;
  sat_restore<T><<<dim3((unsigned)p->tr->sm_count, (unsigned)batch), 256, 0, s>>>((const T*)image, (T*)out, p->sat_list, p->sat_rows, sg);
  LAUNCH((int)cudaGetLastError());
  return RPSF_OK;
}
}  // namespace

extern "C" {

int rpsf_apply_stages(rpsf_plan* p, const void* image, int64_t img_pitch, int64_t img_frame_stride, int img_row0,
                      int img_rows, void* out, int64_t out_pitch, int64_t out_frame_stride, int out_row0,
                      int batch, int stages, void* stream_v) {
  if (!p) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (p->row_end == p->row_begin && stages >= 3) return RPSF_OK;      // an empty band (more ranks than half-patch rows): nothing to write
  if (!image || !out) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  rpsf_transform* t = p->tr;
  if (!t->has_kernel) return fail(RPSF_E_NO_KERNEL, "transfer kernel not loaded (call rpsf_transform_set_kernel)");
  if (batch < 1 || batch > p->max_batch)
    return fail(RPSF_E_INVALID_ARGUMENT, "batch %d outside [1, max_batch=%d]", batch, p->max_batch);
  if (p->pad_mode == PAD_NONE && (img_row0 != 0 || img_rows != p->H))
    return fail(RPSF_E_INVALID_ARGUMENT, "a materialised pad needs the whole frame: img_row0 = 0, img_rows = height");
  if (p->pad_mode != PAD_NONE && p->n_active > 0 && (img_row0 > p->img_lo || img_row0 + img_rows < p->img_hi))
    return fail(RPSF_E_INVALID_ARGUMENT, "resident rows [%d,%d) do not cover the rows this plan reads [%d,%d)",
                img_row0, img_row0 + img_rows, p->img_lo, p->img_hi);
  if (out_row0 > p->row_begin) return fail(RPSF_E_INVALID_ARGUMENT, "out_row0 %d is past row_begin %d", out_row0, p->row_begin);
  if (img_pitch < p->W || out_pitch < p->W) return fail(RPSF_E_INCORRECT_SHAPE, "row pitch smaller than frame width");
  DeviceGuard guard(t->device);
  cudaStream_t s = (cudaStream_t)stream_v;
  ApplyGeom g;
  g.H = p->H; g.W = p->W; g.img_row0 = img_row0; g.img_rows = img_rows; g.img_pitch = img_pitch;
  g.img_frame_stride = img_frame_stride; g.out_row0 = out_row0; g.row_begin = p->row_begin; g.row_end = p->row_end;
  g.out_pitch = out_pitch; g.out_frame_stride = out_frame_stride; g.n_active = p->n_active; g.pad_mode = p->pad_mode;
  g.win_len = t->win_len;
  const size_t rs = real_size(t->dtype);
  const int band = p->row_end - p->row_begin;
  const bool use_stream = p->stream_ok && p->k3_stream && !p->force_phases && !p->force_gather;
  const bool use_gather = p->gather && !p->force_phases && !use_stream;
  const bool need_zero = !((p->colour0_covers || use_gather || use_stream) && stages >= 3);
  cudaEvent_t* ev = nullptr;
  if (p->timing && stages >= 3 && p->n_active > 0) {
    if (p->events_used + 4 > p->events.size()) {
      for (int i = 0; i < 4; ++i) { cudaEvent_t e; CU(cudaEventCreate(&e)); p->events.push_back(e); }
    }
    ev = &p->events[p->events_used];
    p->events_used += 4;
    CU(cudaEventRecord(ev[0], s));
  }
  if (need_zero && band > 0 && stages >= 3) {
    for (int b = 0; b < batch; ++b) {
      char* dst = (char*)out + ((size_t)b * out_frame_stride + (size_t)(p->row_begin - out_row0) * out_pitch) * rs;
      CU(cudaMemset2DAsync(dst, (size_t)out_pitch * rs, 0, (size_t)p->W * rs, band, s));
    }
  }
  const bool sat = p->sat_threshold < INFINITY && stages >= 3;
  if (p->n_active == 0 && !sat) return RPSF_OK;
  SatGeom sg;
  const void* k1_image = image;
  ApplyGeom g1 = g;
  if (sat) {
    if (img_row0 != 0 || img_rows < p->H)
      return fail(RPSF_E_UNSUPPORTED, "the saturation branch needs the whole frame resident (rows [0,%d))", p->H);
    int rc = t->dtype == RPSF_F32 ? sat_prefill<float>(p, image, img_pitch, img_frame_stride, batch, s, &sg, &k1_image)
                                  : sat_prefill<double>(p, image, img_pitch, img_frame_stride, batch, s, &sg, &k1_image);
    if (rc) return rc;
    g1.img_row0 = 0; g1.img_rows = sg.Hp; g1.img_pitch = sg.Wp; g1.img_frame_stride = (long long)sg.Hp * sg.Wp;
    g1.pad_mode = PAD_NONE;
  }
  auto restore = [&]() -> int {
    if (!sat) return RPSF_OK;
    return t->dtype == RPSF_F32
               ? sat_restore_launch<float>(p, image, out, out_pitch, out_frame_stride, out_row0, batch, s, sg)
               : sat_restore_launch<double>(p, image, out, out_pitch, out_frame_stride, out_row0, batch, s, sg);
  };
  if (p->n_active == 0) return restore();
  const size_t rsz0 = real_size(t->dtype);
  const int bulk_ok0 = ((reinterpret_cast<uintptr_t>(k1_image) & 15) == 0 && ((size_t)g1.img_pitch * rsz0) % 16 == 0 &&
                        ((size_t)g1.img_frame_stride * rsz0) % 16 == 0) ? 1 : 0;
  if (p->fused_ok && p->fused_mode == 2 && stages >= 3 && use_stream && p->k1_stream && p->mirrors.empty() &&
      (long long)batch * p->n_active * (t->P / 2) < (1LL << 30)) {
    // one persistent launch: the spectrum goes from role to role through a ring that stays in L2
    CU(cudaMemsetAsync(p->fcounters, 0, p->fcounter_ints * sizeof(int), s));
    LAUNCH(t->ops->fused(t->dtype, k1_image, p->ring, out, p->corners_dev, p->active_dev, t->kmain, t->knyq, p->ftasks_dev,
                         p->fcodes_dev, p->f_warp_items, t->tw, t->win, g1, g, batch, bulk_ok0, p->fg, s));
    if (ev) { CU(cudaEventRecord(ev[1], s)); CU(cudaEventRecord(ev[2], s)); CU(cudaEventRecord(ev[3], s)); }
    return restore();
  }
  if (p->small_ok && p->small_mode == 2 && stages >= 3 && p->mirrors.empty() && !p->force_phases && !p->force_gather) {
    // Opt-in (rpsf_plan_set_small_mode(plan, 2) / RPSF_SMALL=1): the whole per-patch transform inside one CTA (the
    // half-spectrum fits shared memory), then the planes are added.  Measured at config 1: 29.0 us per frame at 8 frames
    // against 22.2 for the three kernels, 39.8 against 41.0 for a single frame — its phases are each slower than the
    // tuned stand-alone kernels, which outweighs the HBM passes it saves.
    if (!p->planes) {
      if (cudaMalloc(&p->planes, p->planes_bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(RPSF_E_CUDA, "out of device memory for the patch planes (%zu bytes)", p->planes_bytes);
      }
    }
    LAUNCH(t->ops->small(t->dtype, k1_image, p->planes, out, p->corners_dev, p->active_dev, t->kmain, t->knyq, t->tw, t->win,
                         p->small_tiles_dev, p->small_n_tiles, p->small_cover_dev, p->small_max_cover, t->P / 2, g1, g, batch, bulk_ok0, s));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (ev) { CU(cudaEventRecord(ev[1], s)); CU(cudaEventRecord(ev[2], s)); CU(cudaEventRecord(ev[3], s)); }
    return restore();
  }
  if (int rc = ensure_workspace(p)) return rc;
  if (p->k1_stream && (long long)batch * p->n_active * (t->P / 2) < (1LL << 30)) {
    // bulk (1-D TMA) row copies need 16-byte aligned rows; the kernel still checks each patch's column offset
    const size_t rsz = real_size(t->dtype);
    const int bulk_ok = ((reinterpret_cast<uintptr_t>(k1_image) & 15) == 0 && ((size_t)g1.img_pitch * rsz) % 16 == 0 &&
                         ((size_t)g1.img_frame_stride * rsz) % 16 == 0) ? 1 : 0;
    LAUNCH(t->ops->k1s(t->dtype, k1_image, p->workspace, p->corners_dev, t->tw, t->win, g1, batch, bulk_ok,
                       t->sm_count, s));
  } else {
    LAUNCH(t->ops->k1(t->dtype, k1_image, p->workspace, p->corners_dev, t->tw, t->win, g1, batch, s));
  }
  if (ev) CU(cudaEventRecord(ev[1], s));
  if (stages < 2) return RPSF_OK;
  // Opt-in (rpsf_plan_set_column_mode(plan, 2) / RPSF_PAIRED=1).  Measured at config 2: 21 % less DRAM traffic and 2 % (8
  // frames) to 5 % (32 frames) more throughput, slower below 8 frames; but the chain kernel is bound by shared-memory
  // instruction issue and latency at 16 warps per SM, not by HBM any more (0.54 of the peak against 0.80 for the
  // classic pass), so the classic pass stays the automatic choice.
  const bool want_paired = p->column_mode == 2;
  if (p->paired_ok && want_paired && use_stream && stages >= 3 && !sat_blocks_mirrors(p, sat)) {
    // paired column pass: the two patches that overlap on a band of rows are summed right after the column IFFT, so
    // the column pass writes half as much and the overlap-add reads half as much
    if (!p->paired) {
      if (cudaMalloc(&p->paired, p->paired_bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(RPSF_E_CUDA, "out of device memory for the paired workspace (%zu bytes)", p->paired_bytes);
      }
    }
    LAUNCH(t->ops->k2c(t->dtype, p->workspace, p->paired, t->kmain, t->knyq, p->active_dev, p->chains_dev, p->n_segments,
                       p->chain_patches_dev, t->tw, t->win, batch, p->n_active, p->bands_total, s));
    if (ev) CU(cudaEventRecord(ev[2], s));
    OutMirrors mir{};
    mir.n = (int)p->mirrors.size();
    for (int d = 0; d < mir.n; ++d) {
      mir.delta[d] = (long long)((char*)p->mirrors[d] - (char*)out);
      if (mir.delta[d] & 15) return fail(RPSF_E_INVALID_ARGUMENT, "mirror %d is not 16-byte congruent with out", d);
    }
    LAUNCH(t->ops->k3p(t->dtype, p->paired, out, p->ptasks_dev, p->pcodes_dev, p->p_warp_items, t->tw, t->win, g, batch,
                       t->sm_count, mir.n ? &mir : nullptr, p->bands_total, s));
    if (ev) CU(cudaEventRecord(ev[3], s));
    return restore();
  }
  LAUNCH(t->ops->k2(t->dtype, p->workspace, t->kmain, t->knyq, p->active_dev, t->tw, g, batch, t->sm_count, s));
  if (ev) CU(cudaEventRecord(ev[2], s));
  if (stages < 3) return RPSF_OK;
  if (!p->mirrors.empty() && (!use_stream || sat))
    return fail(RPSF_E_UNSUPPORTED, "output mirrors need the streaming overlap-add of a covering, without the saturation branch");
  if (use_stream) {
    OutMirrors mir{};
    mir.n = (int)p->mirrors.size();
    for (int d = 0; d < mir.n; ++d) {
      mir.delta[d] = (long long)((char*)p->mirrors[d] - (char*)out);
      if (mir.delta[d] & 15) return fail(RPSF_E_INVALID_ARGUMENT, "mirror %d is not 16-byte congruent with out", d);
    }
    LAUNCH(t->ops->k3s(t->dtype, p->workspace, out, p->stasks_dev, p->scodes_dev, p->n_warp_items, t->tw, t->win, g, batch,
                       t->sm_count, mir.n ? &mir : nullptr, s));
    if (ev) CU(cudaEventRecord(ev[3], s));
    return restore();
  }
  if (use_gather) {
    LAUNCH(t->ops->k3g(t->dtype, p->workspace, out, p->tiles_dev, p->n_tiles, p->groups_dev, p->gitems_dev, t->tw,
                       t->win, p->teams, p->seg_w, g, batch, s));
    if (ev) CU(cudaEventRecord(ev[3], s));
    return restore();
  }
  for (size_t c = 0; c < p->items_dev.size(); ++c) {
    if (p->n_items[c] == 0) continue;
    const int store_only = (c == 0 && p->colour0_covers) ? 1 : 0;
    LAUNCH(t->ops->k3(t->dtype, p->workspace, out, p->corners_dev, p->items_dev[c], p->n_items[c], t->tw, t->win,
                      store_only, g, batch, s));
  }
  if (ev) CU(cudaEventRecord(ev[3], s));
  return restore();
}

int rpsf_plan_enable_timing(rpsf_plan* p, int enabled) {
  if (!p) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  p->timing = enabled != 0;
  return RPSF_OK;
}

int rpsf_plan_read_timing(rpsf_plan* p, double ms[3], int* calls) {
  if (!p || !ms || !calls) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  DeviceGuard guard(p->tr->device);
  ms[0] = ms[1] = ms[2] = 0.0;
  *calls = 0;
  for (size_t i = 0; i + 4 <= p->events_used; i += 4) {
    CU(cudaEventSynchronize(p->events[i + 3]));
    for (int k = 0; k < 3; ++k) {
      float t = 0.f;
      CU(cudaEventElapsedTime(&t, p->events[i + k], p->events[i + k + 1]));
      ms[k] += t;
    }
    ++*calls;
  }
  p->events_used = 0;
  return RPSF_OK;
}

int rpsf_apply(rpsf_plan* p, const void* image, int64_t img_pitch, int64_t img_frame_stride, int img_row0,
               int img_rows, void* out, int64_t out_pitch, int64_t out_frame_stride, int out_row0, int batch,
               void* stream) {
  return rpsf_apply_stages(p, image, img_pitch, img_frame_stride, img_row0, img_rows, out, out_pitch,
                           out_frame_stride, out_row0, batch, 3, stream);
}


int rpsf_convert(const void* src, int sdt, int64_t sp, void* dst, int ddt, int64_t dp, int rows, int cols,
                 int device, void* stream) {
  if (!src || !dst) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (rows <= 0 || cols <= 0) return RPSF_OK;
  DeviceGuard guard(device);
  int rc = ddt == RPSF_F32   ? convert_to<float>(src, sdt, sp, dst, dp, rows, cols, (cudaStream_t)stream)
           : ddt == RPSF_F64 ? convert_to<double>(src, sdt, sp, dst, dp, rows, cols, (cudaStream_t)stream)
                             : -1;
  if (rc) return fail(RPSF_E_UNSUPPORTED, "unsupported conversion %d -> %d", sdt, ddt);
  LAUNCH((int)cudaGetLastError());
  return RPSF_OK;
}

}  // extern "C"

namespace {
// A numpy array is pageable memory.  cudaMemcpyAsync from pageable memory is staged by the driver through its own
// bounce buffer on ONE thread (about 12 GB/s on this pool's hosts against 55 GB/s from pinned memory), and it blocks
// the calling thread meanwhile.  Here the staging is done by a few host threads into the plan's own pinned buffers,
// chunk by chunk, while the DMA and the kernels of the previous chunks run.
bool is_pageable_host(const void* ptr) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) { cudaGetLastError(); return true; }
  return attr.type == cudaMemoryTypeUnregistered;
}
int host_copy_threads() {
  if (const char* v = getenv("RPSF_HOST_COPY_THREADS")) { const int n = atoi(v); if (n >= 1) return std::min(n, 64); }
  // measured on a 16-core host (scripts/pageable_probe.py): 4 threads 8.5 ms per 8 frames, 8 threads 9.2, 16 threads 10.1 —
  // the copy competes with the DMA for the host memory system, more threads only add contention
  const unsigned hw = std::thread::hardware_concurrency();
  return (int)std::max(1u, std::min(4u, hw / 2));
}
// The staged bytes are read next by the DMA engine, never by this CPU: write them with non-temporal stores so the
// destination lines are not first read into the cache (a third of a plain copy's memory traffic, and the host memory
// system — shared with the DMA both ways — is what bounds the staging).
void stream_copy(char* dst, const char* src, size_t bytes) {
#if defined(__SSE2__)
  size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
  if (head > bytes) head = bytes;
  memcpy(dst, src, head);
  dst += head; src += head; bytes -= head;
  const size_t blocks = bytes / 64;
  for (size_t i = 0; i < blocks; ++i) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 0);
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 1);
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 2);
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 3);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 0, a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 1, b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 2, c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 3, d);
    src += 64; dst += 64;
  }
  _mm_sfence();
  memcpy(dst, src, bytes - blocks * 64);
#else
  memcpy(dst, src, bytes);
#endif
}
// A call that fits one chunk has no earlier chunk to hide its staging behind: it is staged in pieces, the DMA of a
// piece under the staging of the next.  Measured (scripts/pageable_probe.py, one 2048^2 frame): float32 1.58 ms through
// the driver's own pageable path, 1.37 staged whole, 1.26 in 4 MB pieces (2 MB: slower); uint16 1.20 / 1.11 / 1.05.
// Several chunks are staged whole (6.7 ms per 8 frames; in 4 MB pieces 7.3).  RPSF_STAGE_SINGLE=0 keeps the
// driver's path for single chunks, RPSF_STAGE_PIECE_MB overrides the piece size.
size_t stage_piece_bytes(bool single_chunk) {
  static const size_t env = [] { const char* e = getenv("RPSF_STAGE_PIECE_MB"); const int mb = e ? atoi(e) : 0; return (size_t)(mb > 0 ? mb : 0) << 20; }();
  if (env) return env;
  return single_chunk ? (size_t)4 << 20 : ~(size_t)0;
}
bool stage_single_chunk(size_t chunk_bytes) {
  static const bool off = [] { const char* e = getenv("RPSF_STAGE_SINGLE"); return e && e[0] == '0'; }();
  return !off && chunk_bytes >= ((size_t)4 << 20);
}
void parallel_copy(void* dst, const void* src, size_t bytes, int threads) {
  const size_t grain = (size_t)1 << 20;
  const int n = (int)std::min<size_t>((size_t)threads, std::max<size_t>(bytes / grain, 1));
  if (n <= 1) { stream_copy((char*)dst, (const char*)src, bytes); return; }
  std::vector<std::thread> pool;
  const size_t per = ((bytes + n - 1) / n + 63) / 64 * 64;
  for (int i = 1; i < n; ++i) {
    const size_t b = std::min(bytes, per * i), e = std::min(bytes, per * (i + 1));
    if (e > b) pool.emplace_back([=]() { stream_copy((char*)dst + b, (const char*)src + b, e - b); });
  }
  stream_copy((char*)dst, (const char*)src, std::min(bytes, per));
  for (auto& t : pool) t.join();
}
}  // namespace

extern "C" {

int rpsf_apply_host(rpsf_plan* p, const void* image, int image_dtype, void* out, int out_dtype, int batch) {
  if (!p) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (batch < 1) return fail(RPSF_E_INVALID_ARGUMENT, "batch must be >= 1");
  if (p->row_end == p->row_begin) return RPSF_OK;                      // empty band: no bytes to produce
  if (!image || !out) return fail(RPSF_E_INVALID_ARGUMENT, "null argument");
  if (out_dtype != RPSF_F32 && out_dtype != RPSF_F64) return fail(RPSF_E_UNSUPPORTED, "output dtype must be f32 or f64");
  const size_t isz = elem_size(image_dtype);
  if (!isz) return fail(RPSF_E_UNSUPPORTED, "unsupported image dtype code %d", image_dtype);
  rpsf_transform* t = p->tr;
  DeviceGuard guard(t->device);
  constexpr int R = rpsf_plan::HOST_SLOTS;
  const int H = p->H, W = p->W, band = p->row_end - p->row_begin;
  const size_t rs = real_size(t->dtype), os = real_size(out_dtype);
  const size_t frame_px = (size_t)H * W, band_px = (size_t)band * W;
  const size_t band_alloc = std::max<size_t>(band_px, 1);
  const int mb = p->max_batch;                       // frames per chunk
  const bool conv_in = image_dtype != t->dtype;
  const bool conv_out = out_dtype != t->dtype;
  if (!p->s_in) {
    CU(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&p->s_comp, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < R; ++i) {
      CU(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&p->ev_comp[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming));
    }
  }
  // a single chunk needs one slot; allocate the ring lazily as far as this call uses it
  const int n_chunks = (batch + mb - 1) / mb;
  const int slots = std::min(R, n_chunks);
  for (int i = 0; i < slots; ++i) {
    if (!p->d_in[i]) CU(cudaMalloc(&p->d_in[i], frame_px * rs * mb));
    if (!p->d_out[i]) CU(cudaMalloc(&p->d_out[i], band_alloc * rs * mb));
  }
  if (conv_in && p->d_in_raw_bytes < frame_px * isz * mb) {
    for (int i = 0; i < R; ++i) { cudaFree(p->d_in_raw[i]); p->d_in_raw[i] = nullptr; }
    p->d_in_raw_bytes = frame_px * isz * mb;
  }
  if (conv_out && p->d_out_conv_bytes < band_alloc * os * mb) {
    for (int i = 0; i < R; ++i) { cudaFree(p->d_out_conv[i]); p->d_out_conv[i] = nullptr; }
    p->d_out_conv_bytes = band_alloc * os * mb;
  }
  for (int i = 0; i < slots; ++i) {
    if (conv_in && !p->d_in_raw[i]) CU(cudaMalloc(&p->d_in_raw[i], p->d_in_raw_bytes));
    if (conv_out && !p->d_out_conv[i]) CU(cudaMalloc(&p->d_out_conv[i], p->d_out_conv_bytes));
  }
  const bool stage = (n_chunks >= 2 || stage_single_chunk(frame_px * isz * (size_t)batch)) && is_pageable_host(image);
  const int copy_threads = stage ? host_copy_threads() : 1;
  if (stage && p->h_stage_bytes < frame_px * isz * mb) {
    for (int i = 0; i < R; ++i) { if (p->h_stage[i]) cudaFreeHost(p->h_stage[i]); p->h_stage[i] = nullptr; }
    p->h_stage_bytes = frame_px * isz * mb;
  }
  for (int i = 0; i < slots && stage; ++i)
    if (!p->h_stage[i]) CU(cudaHostAlloc(&p->h_stage[i], p->h_stage_bytes, cudaHostAllocDefault));
  int rc = RPSF_OK;
  for (int ci = 0; ci < n_chunks && rc == RPSF_OK; ++ci) {
    const int b0 = ci * mb, nb = std::min(mb, batch - b0), k = ci % R;
    const char* src = (const char*)image + (size_t)b0 * frame_px * isz;
    char* dst = (char*)out + (size_t)b0 * band_px * os;
    // upload: the slot's previous occupant must have been consumed by its kernels
    if (ci >= R) CU(cudaStreamWaitEvent(p->s_in, p->ev_comp[k], 0));
    char* d_dst = (char*)(conv_in ? p->d_in_raw[k] : p->d_in[k]);
    const size_t chunk_bytes = frame_px * isz * nb;
    if (!stage) {
      CU(cudaMemcpyAsync(d_dst, src, chunk_bytes, cudaMemcpyHostToDevice, p->s_in));
    } else {
      if (ci >= R) CU(cudaEventSynchronize(p->ev_in[k]));        // the staging slot's previous DMA has read it
      // piece by piece, so that the DMA of a piece runs under the staging of the next one (and under the DMA and
      // kernels of earlier chunks)
      const size_t piece = stage_piece_bytes(n_chunks == 1);
      for (size_t off = 0; off < chunk_bytes; off += std::min(piece, chunk_bytes)) {
        const size_t nbytes = std::min(piece, chunk_bytes - off);
        parallel_copy((char*)p->h_stage[k] + off, src + off, nbytes, copy_threads);
        CU(cudaMemcpyAsync(d_dst + off, (char*)p->h_stage[k] + off, nbytes, cudaMemcpyHostToDevice, p->s_in));
      }
    }
    CU(cudaEventRecord(p->ev_in[k], p->s_in));
    // kernels: need this chunk's upload, and the slot's previous download to be done with d_out
    CU(cudaStreamWaitEvent(p->s_comp, p->ev_in[k], 0));
    if (ci >= R) CU(cudaStreamWaitEvent(p->s_comp, p->ev_out[k], 0));
    if (conv_in) {
      int e = t->dtype == RPSF_F32
                  ? convert_to<float>(p->d_in_raw[k], image_dtype, W, p->d_in[k], W, H * nb, W, p->s_comp)
                  : convert_to<double>(p->d_in_raw[k], image_dtype, W, p->d_in[k], W, H * nb, W, p->s_comp);
      if (e) return fail(RPSF_E_UNSUPPORTED, "unsupported image dtype code %d", image_dtype);
      LAUNCH((int)cudaGetLastError());
    }
    rc = rpsf_apply(p, p->d_in[k], W, (int64_t)frame_px, 0, H, p->d_out[k], W, (int64_t)band_px, p->row_begin, nb,
                    p->s_comp);
    if (rc) break;
    if (band_px == 0) { CU(cudaEventRecord(p->ev_comp[k], p->s_comp)); continue; }
    if (conv_out) {
      int e = out_dtype == RPSF_F32
                  ? convert_to<float>(p->d_out[k], t->dtype, W, p->d_out_conv[k], W, band * nb, W, p->s_comp)
                  : convert_to<double>(p->d_out[k], t->dtype, W, p->d_out_conv[k], W, band * nb, W, p->s_comp);
      if (e) return fail(RPSF_E_UNSUPPORTED, "unsupported output conversion");
      LAUNCH((int)cudaGetLastError());
    }
    CU(cudaEventRecord(p->ev_comp[k], p->s_comp));
    // download
    CU(cudaStreamWaitEvent(p->s_out, p->ev_comp[k], 0));
    CU(cudaMemcpyAsync(dst, conv_out ? p->d_out_conv[k] : p->d_out[k], band_px * os * nb, cudaMemcpyDeviceToHost,
                       p->s_out));
    CU(cudaEventRecord(p->ev_out[k], p->s_out));
  }
  // drain all three streams even on error so no copy is in flight when the caller's buffers go away
  cudaError_t e1 = cudaStreamSynchronize(p->s_in), e2 = cudaStreamSynchronize(p->s_comp),
              e3 = cudaStreamSynchronize(p->s_out);
  if (rc) return rc;
  for (cudaError_t e : {e1, e2, e3})
    if (e != cudaSuccess) return fail(RPSF_E_CUDA, "host apply failed: %s", cudaGetErrorString(e));
  return RPSF_OK;
}

}  // extern "C"
