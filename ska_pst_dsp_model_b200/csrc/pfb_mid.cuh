// Shape-specialised kernels for 4096-channel filterbanks (SKA-Mid, fp32, dual-pol):
//   pfb_rowfft4096 : one 4096-point transform per row (both polarisations), 256 threads x 16 points,
//                    three radix-16 passes in registers with two exchanges through shared memory.
//                    Used as the FFT stage of the analysis (after the sliding-window FIR) and as the
//                    channel IDFT of the inversion (K2).
// Everything else about these shapes runs the generic kernels.
#pragma once
#include "pfb_low.cuh"

namespace pfb {
namespace mid {

typedef Cx<float> C;
using low::fft_reg;
using low::slot;

struct RowFftArgs {
  const void* in;        // [row][4096][2]
  void* out;
  const float2* tw;      // [4096] exp(-2 pi i k / 4096)
  const int* jmap;       // input re-order (combine) or null
  long long nrows;
  // analysis mode: output row (r + kout0) mod kout_mod (0: no wrap), channel window, gain
  long long kout0, kout_mod;
  int chan_first, nchan_out;
  float scale;
};

constexpr int kRowPitch = 257;                       // 16-byte units per o1 plane: odd -> conflict-free transposes
constexpr size_t kRowFftSmem = (size_t)16 * kRowPitch * sizeof(float4);

__device__ __forceinline__ C tw_mul(C a, float2 w, int dir) {
  // a * (dir < 0 ? w : conj(w))
  const float s = dir < 0 ? w.y : -w.y;
  return mk<float>(a.x * w.x - a.y * s, a.x * s + a.y * w.x);
}

// DIR = -1 forward, +1 inverse (un-normalised).  ANALYSIS selects the output addressing.
template <int DIR, bool ANALYSIS>
__global__ void __launch_bounds__(256, 2) pfb_rowfft4096(const RowFftArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* ex = reinterpret_cast<float4*>(smem_raw);
  const int t = threadIdx.x;
  const float4* __restrict__ in = reinterpret_cast<const float4*>(a.in);
  for (long long r = blockIdx.x; r < a.nrows; r += gridDim.x) {
    C x0[16], x1[16];
    {
      // pass A: n = t + 256 i  ->  o1 = k mod 16
      const float4* p = in + r * 4096;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int n = t + 256 * i;
        const float4 v = __ldg(p + (a.jmap ? __ldg(a.jmap + n) : n));
        x0[i] = mk<float>(v.x, v.y);
        x1[i] = mk<float>(v.z, v.w);
      }
      fft_reg<4, 4, DIR>(x0);
      fft_reg<4, 4, DIR>(x1);
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        C u0 = x0[slot<4, 4>(o)], u1 = x1[slot<4, 4>(o)];
        if (o) {
          const float2 w = __ldg(a.tw + ((t * o) & 4095));
          u0 = tw_mul(u0, w, DIR);
          u1 = tw_mul(u1, w, DIR);
        }
        ex[o * kRowPitch + t] = make_float4(u0.x, u0.y, u1.x, u1.y);
      }
    }
    __syncthreads();
    {
      // pass B: thread (o1, q): t' = q + 16 i'  ->  o2
      const int o1 = t >> 4, q = t & 15;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 v = ex[o1 * kRowPitch + q + 16 * i];
        x0[i] = mk<float>(v.x, v.y);
        x1[i] = mk<float>(v.z, v.w);
      }
      __syncthreads();
      fft_reg<4, 4, DIR>(x0);
      fft_reg<4, 4, DIR>(x1);
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        C u0 = x0[slot<4, 4>(o)], u1 = x1[slot<4, 4>(o)];
        if (o) {
          const float2 w = __ldg(a.tw + ((16 * q * o) & 4095));
          u0 = tw_mul(u0, w, DIR);
          u1 = tw_mul(u1, w, DIR);
        }
        ex[o1 * kRowPitch + o * 16 + q] = make_float4(u0.x, u0.y, u1.x, u1.y);
      }
    }
    __syncthreads();
    {
      // pass C: thread (o1, o2), o1 fastest: q = 0..15  ->  o3; bin k = t + 256 o3
      const int o1 = t & 15, o2 = t >> 4;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 v = ex[o1 * kRowPitch + o2 * 16 + i];
        x0[i] = mk<float>(v.x, v.y);
        x1[i] = mk<float>(v.z, v.w);
      }
      fft_reg<4, 4, DIR>(x0);
      fft_reg<4, 4, DIR>(x1);
      if (ANALYSIS) {
        long long ko = r + a.kout0;
        if (a.kout_mod > 0) ko %= a.kout_mod;
        float4* op = reinterpret_cast<float4*>(a.out) + ko * a.nchan_out;
#pragma unroll
        for (int o3 = 0; o3 < 16; ++o3) {
          const int ci = (t + 256 * o3 - a.chan_first) & 4095;
          const C u0 = x0[slot<4, 4>(o3)], u1 = x1[slot<4, 4>(o3)];
          if (ci < a.nchan_out) op[ci] = make_float4(u0.x * a.scale, u0.y * a.scale, u1.x * a.scale, u1.y * a.scale);
        }
      } else {
        float4* op = reinterpret_cast<float4*>(a.out) + r * 4096 + t;
#pragma unroll
        for (int o3 = 0; o3 < 16; ++o3) {
          const C u0 = x0[slot<4, 4>(o3)], u1 = x1[slot<4, 4>(o3)];
          op[256 * o3] = make_float4(u0.x, u0.y, u1.x, u1.y);
        }
      }
    }
    __syncthreads();
  }
}

template <int DIR, bool ANALYSIS>
inline int rowfft4096_launch(const RowFftArgs& a, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaFuncSetAttribute(pfb_rowfft4096<DIR, ANALYSIS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kRowFftSmem);
  if (e != cudaSuccess) return (int)e;
  long long grid = a.nrows < 2ll * sms ? a.nrows : 2ll * sms;
  pfb_rowfft4096<DIR, ANALYSIS><<<(unsigned)grid, 256, kRowFftSmem, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace mid
}  // namespace pfb
