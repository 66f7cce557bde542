// Microbenchmark: what a gpu-scope release costs a warp that has just stored 2 KB (sm_100a).
// MODE 0: stores only; 1: __threadfence + atomicAdd (lane 0); 2: red.release.gpu (lane 0); 3: fence.acq_rel.gpu + red.relaxed;
// 4: as 2 but the release is issued after `work` of independent FMAs (stores are older when the fence comes)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float2* out, int* flags, int iters, int work) {
  const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float2* dst = out + (size_t)warp * 256;
  float a = lane, b = 1.0001f;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 4) { for (int w = 0; w < work; ++w) a = fmaf(a, b, 0.5f); }
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i * 32 + lane] = make_float2(a, (float)it);
    if (MODE != 4) { for (int w = 0; w < work; ++w) a = fmaf(a, b, 0.5f); }
    __syncwarp();
    if (lane == 0) {
      if (MODE == 1) { __threadfence(); atomicAdd(flags + (warp & 1023), 1); }
      if (MODE == 2 || MODE == 4) asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(flags + (warp & 1023)), "r"(1) : "memory");
      if (MODE == 3) { asm volatile("fence.acq_rel.gpu;" ::: "memory"); asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(flags + (warp & 1023)), "r"(1) : "memory"); }
    }
  }
  if (a == 12345.f) out[0].x = a;
}
template <int MODE> void run(const char* name, float2* out, int* flags, int work) {
  const int iters = 2000, blocks = 148, threads = 512;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, flags, iters, work);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, flags, iters, work);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-34s work %5d: %7.3f us per iteration per warp\n", name, work, 1e3 * ms / iters);
}
int main() {
  float2* out; int* flags;
  cudaMalloc(&out, (size_t)148 * 16 * 256 * sizeof(float2)); cudaMalloc(&flags, 1024 * sizeof(int)); cudaMemset(flags, 0, 4096);
  for (int work : {0, 2000}) {
    run<0>("stores only", out, flags, work);
    run<1>("__threadfence + atomicAdd", out, flags, work);
    run<2>("red.release.gpu", out, flags, work);
    run<3>("fence.acq_rel.gpu + red.relaxed", out, flags, work);
    run<4>("work, stores, red.release", out, flags, work);
  }
  return 0;
}
