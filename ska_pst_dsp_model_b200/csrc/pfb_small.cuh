// Shape-specialised FIR for few-channel, long-filter analysis: N = 8 channels, OS 8/7 (step 7), 384 taps per arm
// (BASELINE configs[0]: config/PST_filtertaps.txt used as an 8-channel prototype, polyphase_analysis.m:83-122).
// 1759 flop per input sample against 17 bytes: the shape is bound by the fp32 FMA rate, not by HBM (hard
// part 2).
//
//  * gcd(7, 8) = 1, but blocks k and k + 8 read the same sub-stream u[n] = x[7 p + j + 8 n] (p = k mod 8, arm j)
//    shifted by 7 elements, so a thread that owns (arm j, phase p) slides along u: output i needs u[7 i + m], m < 384.
//  * a CTA owns ONE arm (blockIdx.y), so its 384 taps are CTA-uniform: they come from the kernel-parameter constant bank
//    (one indexed LDC per tap) and stay in the operand-reuse cache for the 5 packed FMAs that use them
//    (SASS: FFMA2 Rd, Rtap.reuse.F32, Rw.F32x2, Racc.F32x2): two fresh register pairs per FFMA2.  Measured: the kernel
//    runs at ~0.3 of the nominal fp32 FMA roof - a packed FMA with two fresh 64-bit operands issues every ~4 cycles
//    (register-file bandwidth), so ~0.5 is the ceiling of this instruction mix;
//  * every thread computes R = 5 outputs i0..i0+4 at once: their sample indices at tap m are u[7 i0 + m + {0,7,..,28}],
//    a 29-element window that slides by ONE element per tap, held in a 32-slot circular register file with static
//    indices (the tap loop is unrolled 32-fold, 12 iterations): one shared-memory load per 5 packed FMAs.
//    R is odd so that the two chunks sharing a shared-memory wavefront sit 8 (mod 16) eight-byte units apart, next to
//    the eight phases' 7 p (mod 16): conflict-free;
//  * the arm sums go to a [k][slot] scratch with the oversampling rotation already applied (slot = (7 k + j) mod 8);
//    pfb_fft8_store runs the 8-point FFTs (one per thread, in registers) and the scaled store.  (A fused form - 8 warps
//    = 8 arms per CTA, FFT stage from shared memory in the same kernel - was measured: 22.5 us against 18.4 us for the
//    two kernels, the serial FFT tail of every CTA costs more than the second launch.)
#pragma once
#include "pfb_generic.cuh"

namespace pfb {
namespace small8 {

typedef Cx<float> C;
constexpr int kTaps = 384;                                   // per arm
constexpr int kR = 5;                                        // outputs per thread
constexpr int kW = 32;                                       // circular window (>= 7 (kR - 1) + 1 = 29)
constexpr int kThreads = 256;                                // 8 phases x 32 chunks
constexpr int kSpanBlocks = kThreads * kR;                   // 1280 output blocks per CTA
constexpr int kTileSamples = 7 * kSpanBlocks + 8 * kTaps;    // 12 032 input samples

struct Fir8Taps {
  float t[8][kTaps];                                         // t[j][m] = f[j + 8 m]  (12 KB of kernel parameters)
};
struct Fir8Args {
  const C* in;              // [t] single polarisation
  C* y;                     // [k][8] rotated arm sums
  long long n_dat, nblocks;
  long long start0;         // window start of block 0 relative to in[0] (>= 0)
  int rot;                  // (phase0 + start0) mod 8: FFT slot of arm j in block k = (rot + 7 k + j) mod 8
};

__global__ void __launch_bounds__(kThreads, 2)
    pfb_fir8_slide(const Fir8Args a, const __grid_constant__ Fir8Taps taps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* tile = reinterpret_cast<C*>(smem_raw);
  const int j = blockIdx.y;
  const long long k0 = (long long)blockIdx.x * kSpanBlocks;
  const long long s0 = a.start0 + 7 * k0;
  if ((s0 & 1) == 0 && ((uintptr_t)a.in % 16) == 0) {
    // two samples per 16-byte load, 8 independent loads in flight per thread (the tile is 96 KB: a serial load loop
    // would cost more than the FIR itself)
    const float4* src = reinterpret_cast<const float4*>(a.in + s0);
    float4* dst = reinterpret_cast<float4*>(tile);
    const long long avail = a.n_dat - s0;                  // samples left in the stream from the tile start
#pragma unroll 8
    for (int t = threadIdx.x; t < kTileSamples / 2; t += kThreads) {
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (2ll * t + 1 < avail) r = __ldg(src + t);
      dst[t] = r;
    }
    if ((avail & 1) && avail > 0 && avail <= kTileSamples && threadIdx.x == 0) {   // odd stream end inside the tile
      const float2 q = __ldg(reinterpret_cast<const float2*>(a.in + s0 + avail - 1));
      __syncwarp(0x1);
      tile[avail - 1] = mk<float>(q.x, q.y);
    }
  } else {
#pragma unroll 8
    for (int t = threadIdx.x; t < kTileSamples; t += kThreads) {
      const long long s = s0 + t;
      float2 r = make_float2(0.f, 0.f);
      if (s < a.n_dat) r = __ldg(reinterpret_cast<const float2*>(a.in + s));
      tile[t] = mk<float>(r.x, r.y);
    }
  }
  __syncthreads();
  const int p = threadIdx.x & 7, c = threadIdx.x >> 3;
  const int i0 = kR * c;
  const C* base = tile + 7 * p + j + 56 * i0;                // u[7 i0 + n] = base[8 n]
  C w[kW];
#pragma unroll
  for (int n = 0; n < 7 * (kR - 1) + 1; ++n) w[n] = base[8 * n];
  float2 acc[kR];
#pragma unroll
  for (int r = 0; r < kR; ++r) acc[r] = make_float2(0.f, 0.f);
  const float* tj = taps.t[j];
  for (int it = 0; it < kTaps / kW; ++it) {
#pragma unroll
    for (int mm = 0; mm < kW; ++mm) {
      const int m = it * kW + mm;
      const float tap = tj[m];
#pragma unroll
      for (int r = 0; r < kR; ++r) {
        const C v = w[(mm + 7 * r) % kW];
        acc[r] = __ffma2_rn(make_float2(tap, tap), make_float2(v.x, v.y), acc[r]);
      }
      // element m is dead now; element m + 29 moves into the free slot (needed up to 7 (kR - 1) + kTaps - 1)
      if (m + 7 * (kR - 1) + 1 < 7 * (kR - 1) + kTaps) w[(mm + 7 * (kR - 1) + 1) % kW] = base[8 * (m + 7 * (kR - 1) + 1)];
    }
  }
#pragma unroll
  for (int r = 0; r < kR; ++r) {
    const long long k = k0 + p + 8 * (i0 + r);
    if (k < a.nblocks) a.y[k * 8 + ((a.rot + 7 * k + j) & 7)] = mk<float>(acc[r].x, acc[r].y);
  }
}

// FFT stage: one thread per block, 8 rotated arm sums in, 8 channels out (x scale), 64 bytes each way
__global__ void __launch_bounds__(256) pfb_fft8_store(const C* __restrict__ y, C* __restrict__ out, long long nblocks,
                                                     long long kout0, float scale) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nblocks) return;
  C v[8];
  const float4* p = reinterpret_cast<const float4*>(y + k * 8);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 q = __ldg(p + i);
    v[2 * i] = mk<float>(q.x, q.y);
    v[2 * i + 1] = mk<float>(q.z, q.w);
  }
  Dft<8, -1, float>::run(v);
  float4* o = reinterpret_cast<float4*>(out + (kout0 + k) * 8);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    o[i] = make_float4(v[2 * i].x * scale, v[2 * i].y * scale, v[2 * i + 1].x * scale, v[2 * i + 1].y * scale);
}

inline int fir8_launch(const Fir8Args& a, const Fir8Taps& taps, cudaStream_t st) {
  const size_t smem = (size_t)kTileSamples * sizeof(C);
  cudaError_t e = cudaFuncSetAttribute(pfb_fir8_slide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const long long spans = (a.nblocks + kSpanBlocks - 1) / kSpanBlocks;
  if (spans > 2147483647ll) return (int)cudaErrorInvalidValue;
  pfb_fir8_slide<<<dim3((unsigned)spans, 8), kThreads, smem, st>>>(a, taps);
  return (int)cudaGetLastError();
}

}  // namespace small8
}  // namespace pfb
