// Microbenchmark: scalar FFMA/FADD vs packed FFMA2/FADD2 throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, float a, float b, int iters) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, blockIdx.x * 0.002f - i);
  const float2 va = make_float2(a, a * 1.0001f), vb = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(acc[i].x, va.x, vb.x); acc[i].y = fmaf(acc[i].y, va.y, vb.y); }
      if (MODE == 1) acc[i] = __ffma2_rn(acc[i], va, vb);
      if (MODE == 2) { acc[i].x = acc[i].x + vb.x; acc[i].y = acc[i].y + vb.y; }
      if (MODE == 3) acc[i] = __fadd2_rn(acc[i], vb);
    }
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float2* out) {
  const int iters = 4096, blocks = 148 * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f, iters);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * 256 * iters * 16;   // scalar fp32 ops (fma counts 1)
  printf("%-8s %8.3f ms  %8.2f Tops/s (fp32 lanes-ops)\n", name, ms, ops / ms / 1e9);
}
int main() {
  float2* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
  run<0>("FFMA", out); run<1>("FFMA2", out); run<2>("FADD", out); run<3>("FADD2", out);
  return 0;
}
