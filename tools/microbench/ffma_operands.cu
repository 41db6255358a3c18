// Microbenchmark behind DESIGN.md section 5 ("register-file bandwidth bounds the FIR"): fp32 FMA issue rate as a
// function of how many FRESH vector-register operands each instruction reads.
//   mode 0: FFMA  acc_i += tap_j * w_k          three fresh registers per instruction (the polyphase FIR inner loop)
//   mode 1: FFMA  acc_i += tap   * w_k          tap shared by consecutive instructions (operand-reuse cache)
//   mode 2: FFMA2 acc_i.xy += tap_j * w_k.xy    scalar tap + two fresh register pairs
//   mode 3: FFMA2 acc_i.xy += tap * w_k.xy      tap shared by consecutive instructions
//   mode 4: FFMA2 acc_i.xy += c[tap] * w_k.xy   tap from the constant bank / uniform register
//   mode 5: FFMA  acc_i = acc_i * a + b         a, b shared by all (the classic peak-rate loop)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_operands ffma_operands.cu && ./ffma_operands
#include <cstdio>
#include <cuda_runtime.h>

struct Taps { float t[64]; };

template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float2* out, const float2* in, const float* tg, const __grid_constant__ Taps ct, int iters) {
  float2 w[32];
  float tap[32];
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 32; ++i) { w[i] = in[i * 256 + threadIdx.x]; tap[i] = tg[i * 256 + threadIdx.x]; }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(0.f, 0.f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 32; ++t) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 v = w[(t + 5 * i) & 31];
        if (MODE == 0) { acc[i].x = fmaf(tap[(t + 3 * i) & 31], v.x, acc[i].x); acc[i].y = fmaf(tap[(t + 3 * i + 7) & 31], v.y, acc[i].y); }
        if (MODE == 1) { acc[i].x = fmaf(tap[t], v.x, acc[i].x); acc[i].y = fmaf(tap[t], v.y, acc[i].y); }
        if (MODE == 2) acc[i] = __ffma2_rn(make_float2(tap[(t + 3 * i) & 31], tap[(t + 3 * i) & 31]), v, acc[i]);
        if (MODE == 3) acc[i] = __ffma2_rn(make_float2(tap[t], tap[t]), v, acc[i]);
        if (MODE == 4) acc[i] = __ffma2_rn(make_float2(ct.t[t], ct.t[t]), v, acc[i]);
        if (MODE == 5) { acc[i].x = fmaf(acc[i].x, tap[0], tap[1]); acc[i].y = fmaf(acc[i].y, tap[0], tap[1]); }
      }
    }
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, float2* out, const float2* in, const float* tg, const Taps& ct) {
  const int iters = 512, blocks = 148 * 2 * 4;
  k<MODE><<<blocks, 256>>>(out, in, tg, ct, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, in, tg, ct, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double laneops = (double)blocks * 256 * iters * 32 * 8 * 2;      // fp32 multiply-adds
  printf("{\"mode\": \"%s\", \"ms\": %.3f, \"T_lane_fma_per_s\": %.2f, \"err\": \"%s\"}\n", name, ms, laneops / ms / 1e9,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float2 *out, *in; float* tg;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
  cudaMalloc(&in, 32 * 256 * sizeof(float2)); cudaMemset(in, 0, 32 * 256 * sizeof(float2));
  cudaMalloc(&tg, 32 * 256 * sizeof(float)); cudaMemset(tg, 0, 32 * 256 * sizeof(float));
  Taps ct; for (int i = 0; i < 64; ++i) ct.t[i] = 1.0f / (i + 1);
  run<0>("FFMA  3 fresh regs", out, in, tg, ct);
  run<1>("FFMA  tap reused", out, in, tg, ct);
  run<2>("FFMA2 scalar tap + 2 fresh pairs", out, in, tg, ct);
  run<3>("FFMA2 tap reused + 2 fresh pairs", out, in, tg, ct);
  run<4>("FFMA2 constant-bank tap + 2 fresh pairs", out, in, tg, ct);
  run<5>("FFMA  2 shared operands (peak loop)", out, in, tg, ct);
  return 0;
}
