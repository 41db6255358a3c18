// Common device helpers for the B200 PFB kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pfb {

// interleaved (re, im) as in the DADA payload (NDIM=2); aligned so that one complex moves as a single
// 64-bit (fp32) / 128-bit (fp64) load or store
template <typename T> struct alignas(2 * sizeof(T)) Cx { T x, y; };
template <typename T> struct VecOf;
template <> struct VecOf<float>  { using type = float2;  };
template <> struct VecOf<double> { using type = double2; };

template <typename T> __host__ __device__ __forceinline__ Cx<T> mk(T a, T b) { Cx<T> r; r.x = a; r.y = b; return r; }
template <typename T> __device__ __forceinline__ Cx<T> operator+(Cx<T> a, Cx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T> __device__ __forceinline__ Cx<T> operator-(Cx<T> a, Cx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <typename T> __device__ __forceinline__ Cx<T> operator*(Cx<T> a, Cx<T> b) {
  return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename T> __device__ __forceinline__ Cx<T> operator*(Cx<T> a, T s) { return mk<T>(a.x * s, a.y * s); }
template <typename T> __device__ __forceinline__ Cx<T> cconj(Cx<T> a) { return mk<T>(a.x, -a.y); }
// multiply by +i (DIR=+1) or -i (DIR=-1)
template <int DIR, typename T> __device__ __forceinline__ Cx<T> mul_i(Cx<T> a) {
  return DIR > 0 ? mk<T>(-a.y, a.x) : mk<T>(a.y, -a.x);
}
// a * (DIR<0 ? w : conj(w)); tables hold the forward twiddle w = exp(-2 pi i k/n)
template <int DIR, typename T> __device__ __forceinline__ Cx<T> mul_tw(Cx<T> a, Cx<T> w) {
  return DIR < 0 ? a * w : a * cconj(w);
}

// shared-memory index padding: one extra element every 128 bytes so that the
// power-of-two strides of the FFT passes do not pile onto one bank group.
template <typename T> struct Pad;
template <> struct Pad<float>  { static constexpr int SHIFT = 4; };   // 16 x 8 B
template <> struct Pad<double> { static constexpr int SHIFT = 3; };   //  8 x 16 B
template <typename T> __host__ __device__ __forceinline__ int padi(int i) { return i + (i >> Pad<T>::SHIFT); }
template <typename T> __host__ __device__ __forceinline__ int padded_len(int n) { return n + (n >> Pad<T>::SHIFT) + 1; }

__device__ __forceinline__ int pmod(long long v, int n) {
  int r = (int)(v % n);
  return r < 0 ? r + n : r;
}

}  // namespace pfb
