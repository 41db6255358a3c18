// Integer DADA payloads on the device (matlab/DADARead.m:41-53, 66-70: NBIT 8 / 16 samples are read as
// int8 / int16 and cast to floating point; matlab/write_dada_header.m:15-22: integer classes are written as
// NBIT 8 / 16).  Files and host buffers keep the integers (2-4x fewer PCIe / disk bytes); the conversion to the
// arithmetic type happens here, next to the kernels.  int -> float is exact (|v| < 2^15), so results are
// bit-identical to converting on the host.
#pragma once
#include "pfb_common.cuh"

namespace pfb {

// 16 values per thread-iteration; nvalues = real values (2 per complex sample)
template <typename I, typename F>
__global__ void __launch_bounds__(256) pfb_unpack_kernel(const I* __restrict__ in, F* __restrict__ out, long long n) {
  constexpr int V = 16 / sizeof(I);                   // values per 16-byte load
  const long long nv = n / V;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const int4 raw = __ldg(reinterpret_cast<const int4*>(in) + i);
    const I* v = reinterpret_cast<const I*>(&raw);
    F o[V];
#pragma unroll
    for (int k = 0; k < V; ++k) o[k] = (F)v[k];
    if (sizeof(F) == 4) {
      float4* p = reinterpret_cast<float4*>(out + i * V);
#pragma unroll
      for (int k = 0; k < V / 4; ++k) p[k] = *reinterpret_cast<const float4*>(&o[4 * k]);
    } else {
      double2* p = reinterpret_cast<double2*>(out + i * V);
#pragma unroll
      for (int k = 0; k < V / 2; ++k) p[k] = *reinterpret_cast<const double2*>(&o[2 * k]);
    }
  }
  // tail (n not a multiple of V)
  for (long long i = nv * V + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (F)in[i];
}

// Matlab int8()/int16() cast of scale*x: round half away from zero, saturate
template <typename F, typename I>
__global__ void __launch_bounds__(256) pfb_pack_kernel(const F* __restrict__ in, I* __restrict__ out, long long n, F scale) {
  const F lo = sizeof(I) == 1 ? (F)-128 : (F)-32768, hi = sizeof(I) == 1 ? (F)127 : (F)32767;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    F v = in[i] * scale;
    v = v < 0 ? -floor(-v + (F)0.5) : floor(v + (F)0.5);
    v = v < lo ? lo : (v > hi ? hi : v);
    out[i] = (I)v;
  }
}

}  // namespace pfb
