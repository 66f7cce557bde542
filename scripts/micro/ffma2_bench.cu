// Microbenchmark: FP32 scalar vs packed (f32x2) add/fma throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float2* out, int iters, float2 seed) {
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
  const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, -0.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }        // 2 FFMA
      if (MODE == 1) { a[i] = __ffma2_rn(a[i], m, c); }                                          // 1 FFMA2
      if (MODE == 2) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }                            // 2 FADD
      if (MODE == 3) { a[i] = __fadd2_rn(a[i], c); }                                             // 1 FADD2
      if (MODE == 4) { a[i] = __fmul2_rn(a[i], m); }                                             // 1 FMUL2
    }
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float2* out) {
  const int iters = 4096, blocks = 148 * 8, threads = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, iters, make_float2(1, 2));
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters, make_float2(1, 2));
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double lane_ops = (double)blocks * threads * iters * 8 * 2;   // fp32 lane-operations (x and y)
  printf("%-8s %8.3f ms  %8.2f T lane-ops/s\n", name, ms, lane_ops / ms / 1e9);
}
int main() {
  float2* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
  run<0>("FFMA", out); run<1>("FFMA2", out); run<2>("FADD", out); run<3>("FADD2", out); run<4>("FMUL2", out);
  return 0;
}
