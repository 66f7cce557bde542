// widen_bench.cu — can the host widen float32 -> float64 fast enough to make "float32 on the wire" pay?
// Pipeline per frame: H2D of a float32 frame (other direction, concurrently), D2H of the result either as
// float64 (baseline) or as float32 into a pinned ring followed by a multi-threaded widen into the caller's array.
#include <cuda_runtime.h>
#include <immintrin.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

static void widen(const float* src, double* dst, size_t n) {
  size_t i = 0;
#ifdef __AVX2__
  for (; i + 8 <= n; i += 8) {
    __m256 v = _mm256_loadu_ps(src + i);
    _mm256_stream_pd(dst + i, _mm256_cvtps_pd(_mm256_castps256_ps128(v)));
    _mm256_stream_pd(dst + i + 4, _mm256_cvtps_pd(_mm256_extractf128_ps(v, 1)));
  }
#endif
  for (; i < n; ++i) dst[i] = (double)src[i];
}

int main(int argc, char** argv) {
  const size_t px = 2048 * 2048;
  const int frames = 64, ring = 4;
  float* d32; double* d64; float* din;
  cudaMalloc(&d32, px * 4 * ring); cudaMalloc(&d64, px * 8 * ring); cudaMalloc(&din, px * 4 * ring);
  float* hin; cudaHostAlloc(&hin, px * 4 * ring, 0);
  double* hout; cudaHostAlloc(&hout, px * 8 * frames, 0);           // caller's array (pinned, as bench.py uses)
  float* stage; cudaHostAlloc(&stage, px * 4 * ring, 0);
  memset(hin, 1, px * 4 * ring); memset(hout, 0, px * 8 * frames); memset(stage, 0, px * 4 * ring);
  cudaStream_t sin_, sout; cudaStreamCreate(&sin_); cudaStreamCreate(&sout);
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };

  // baseline: float64 on the wire, duplex with the float32 upload
  for (int rep = 0; rep < 2; ++rep) {
    auto t0 = now();
    for (int f = 0; f < frames; ++f) {
      cudaMemcpyAsync(din + (f % ring) * px, hin + (f % ring) * px, px * 4, cudaMemcpyHostToDevice, sin_);
      cudaMemcpyAsync(hout + (size_t)f * px, d64 + (f % ring) * px, px * 8, cudaMemcpyDeviceToHost, sout);
    }
    cudaDeviceSynchronize();
    double t = ms(t0, now());
    if (rep) printf("f64 on the wire:            %.3f ms/frame  -> %.2f Gpix/s\n", t / frames, px / (t / frames) / 1e6);
  }
  // float32 on the wire + N-thread widen
  for (int nthreads : {2, 4, 8, 12, 16, 24}) {
    if (nthreads > (int)std::thread::hardware_concurrency()) break;
    std::vector<cudaEvent_t> ev(frames);
    for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    std::vector<std::atomic<int>> converted(frames);
    for (auto& c : converted) c = 0;
    auto t0 = now();
    std::vector<std::thread> pool;
    for (int w = 0; w < nthreads; ++w)
      pool.emplace_back([&, w] {
        for (int f = 0; f < frames; ++f) {
          cudaEventSynchronize(ev[f]);                               // frame f has landed in the ring
          const size_t lo = (px * w / nthreads) & ~size_t(15), hi = w + 1 == nthreads ? px : (px * (w + 1) / nthreads) & ~size_t(15);
          widen(stage + (f % ring) * px + lo, hout + (size_t)f * px + lo, hi - lo);
          converted[f].fetch_add(1);
        }
      });
    for (int f = 0; f < frames; ++f) {
      if (f >= ring) while (converted[f - ring].load() < nthreads) std::this_thread::yield();   // ring slot free
      cudaMemcpyAsync(din + (f % ring) * px, hin + (f % ring) * px, px * 4, cudaMemcpyHostToDevice, sin_);
      cudaMemcpyAsync(stage + (f % ring) * px, d32 + (f % ring) * px, px * 4, cudaMemcpyDeviceToHost, sout);
      cudaEventRecord(ev[f], sout);
    }
    for (auto& th : pool) th.join();
    cudaDeviceSynchronize();
    double t = ms(t0, now());
    printf("f32 on the wire, %2d threads: %.3f ms/frame  -> %.2f Gpix/s\n", nthreads, t / frames, px / (t / frames) / 1e6); fflush(stdout);
    for (auto& e : ev) cudaEventDestroy(e);
  }
  printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
  return 0;
}
