// Microbenchmark: does packing the complex adds of an FFT butterfly into fp32x2 instructions free issue slots for the
// shared-memory instructions that run beside them?  (the resample kernel issues at 0.73/cycle with
// 30 % of its stall samples "not selected", FMA pipe 58 %, shared-memory pipe 68 %.)
//   each loop pass: 16 complex values, a radix-4-like network of 32 complex add/sub, then 4 LDS.64 + 4 STS.64 per thread
//   mode 0: scalar FADD (2 per complex add)     mode 1: __fadd2_rn / __ffma2_rn(b, -1, a) (1 per complex add)
//   mode 2/3: the same without the shared-memory instructions
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o packed_fft_ops packed_fft_ops.cu && ./packed_fft_ops
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 cadd(float2 a, float2 b, bool packed) {
  if (packed) return __fadd2_rn(a, b);
  return make_float2(a.x + b.x, a.y + b.y);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b, bool packed) {
  if (packed) return __ffma2_rn(b, make_float2(-1.f, -1.f), a);
  return make_float2(a.x - b.x, a.y - b.y);
}

template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float2* out, const float2* in, int iters) {
  constexpr bool PACKED = (MODE & 1) != 0;
  constexpr bool SMEM = MODE < 2;
  __shared__ float2 sh[256 * 17];
  float2 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = in[i * 256 + threadIdx.x];
  float2* mine = sh + threadIdx.x * 17;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int st = 0; st < 4; ++st) {          // four stages of 8 add/sub pairs over 16 values
      const int h = 1 << st;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (!(i & h)) {
          const float2 a = x[i], b = x[i | h];
          x[i] = cadd(a, b, PACKED);
          x[i | h] = csub(a, b, PACKED);
        }
      }
    }
    if (SMEM) {
#pragma unroll
      for (int i = 0; i < 4; ++i) mine[i] = x[4 * i];
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) x[4 * i + 1] = cadd(x[4 * i + 1], sh[((threadIdx.x ^ 1) * 17) + i], PACKED);
      __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i].x *= 0.25f; x[i].y *= 0.25f; }   // keep the values finite (scalar in both modes)
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 16; ++i) { s.x += x[i].x; s.y += x[i].y; }
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, float2* out, const float2* in) {
  const int iters = 2048, blocks = 148 * 2 * 4;
  k<MODE><<<blocks, 256>>>(out, in, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, in, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double cadds = (double)blocks * 256 * iters * (64 + (MODE < 2 ? 4 : 0));
  printf("{\"mode\": \"%s\", \"ms\": %.3f, \"T_complex_add_per_s\": %.2f, \"err\": \"%s\"}\n", name, ms, cadds / ms / 1e9,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float2 *in, *out;
  cudaMalloc(&in, 16 * 256 * sizeof(float2));
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
  cudaMemset(in, 0, 16 * 256 * sizeof(float2));
  run<0>("scalar FADD + LDS/STS", out, in);
  run<1>("packed FADD2/FFMA2 + LDS/STS", out, in);
  run<2>("scalar FADD only", out, in);
  run<3>("packed FADD2/FFMA2 only", out, in);
  return 0;
}
