// Generic in-place mixed-radix FFT on shared memory (radices 2,3,4,5,7,8).
//
// Two transposed algorithms share one digit-reversal table so that no separate
// permutation pass is ever needed:
//   * DIF (decimation in frequency): natural-order input, output bin k is left
//     at storage position perm[k];
//   * DIT (decimation in time): input element n is placed at perm[n], output is
//     in natural order.
// The analysis kernels run DIF and un-permute in the (coalesced) global store;
// the inversion kernels place their gathered spectra at perm[] and run DIT.
// A batch of FFTs lives as rows of `ld` elements, each row index-padded with
// padi<T>() (one element per 128 B) against bank conflicts.
#pragma once
#include "pfb_common.cuh"

namespace pfb {

constexpr int kMaxPass = 12;

struct FftDev {
  int n;
  int npass;
  int radix[kMaxPass];   // DIF pass order
  const void* tw;        // Cx<T>[n], tw[k] = exp(-2 pi i k / n)
  const int* perm;       // int[n]
};

template <typename T> __device__ __forceinline__ Cx<T> ldg_cx(const Cx<T>* p);
template <> __device__ __forceinline__ Cx<float> ldg_cx<float>(const Cx<float>* p) {
  float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return mk<float>(v.x, v.y);
}
template <> __device__ __forceinline__ Cx<double> ldg_cx<double>(const Cx<double>* p) {
  double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return mk<double>(v.x, v.y);
}

// ---- small DFTs:  X_o = sum_i a_i exp(DIR * 2 pi i * i o / R) ----------------
template <int R, int DIR, typename T> struct Dft;

template <int DIR, typename T> struct Dft<2, DIR, T> {
  static __device__ __forceinline__ void run(Cx<T>* a) {
    Cx<T> t = a[0];
    a[0] = t + a[1];
    a[1] = t - a[1];
  }
};

template <int DIR, typename T> struct Dft<4, DIR, T> {
  static __device__ __forceinline__ void run(Cx<T>* a) {
    Cx<T> t0 = a[0] + a[2], t1 = a[0] - a[2], t2 = a[1] + a[3], t3 = mul_i<DIR>(a[1] - a[3]);
    a[0] = t0 + t2; a[2] = t0 - t2; a[1] = t1 + t3; a[3] = t1 - t3;
  }
};

template <int DIR, typename T> struct Dft<8, DIR, T> {
  static __device__ __forceinline__ void run(Cx<T>* a) {
    Cx<T> e[4] = {a[0], a[2], a[4], a[6]};
    Cx<T> o[4] = {a[1], a[3], a[5], a[7]};
    Dft<4, DIR, T>::run(e);
    Dft<4, DIR, T>::run(o);
    const T h = (T)0.70710678118654752440;
    // w^1 = (1 + DIR i)/sqrt2, w^2 = DIR i, w^3 = (-1 + DIR i)/sqrt2
    Cx<T> o1 = mk<T>((o[1].x - DIR * o[1].y) * h, (o[1].y + DIR * o[1].x) * h);
    Cx<T> o2 = mul_i<DIR>(o[2]);
    Cx<T> o3 = mk<T>((-o[3].x - DIR * o[3].y) * h, (-o[3].y + DIR * o[3].x) * h);
    a[0] = e[0] + o[0]; a[4] = e[0] - o[0];
    a[1] = e[1] + o1;   a[5] = e[1] - o1;
    a[2] = e[2] + o2;   a[6] = e[2] - o2;
    a[3] = e[3] + o3;   a[7] = e[3] - o3;
  }
};

// odd primes: X_k, X_{R-k} = a0 + sum_j c_jk (a_j + a_{R-j}) +/- i DIR sum_j s_jk (a_j - a_{R-j})
template <int R> struct OddTab;
template <> struct OddTab<3> {
  static __device__ __forceinline__ double c(int m) { const double t[3] = {1.0, -0.5, -0.5}; return t[m]; }
  static __device__ __forceinline__ double s(int m) {
    const double t[3] = {0.0, 0.86602540378443864676, -0.86602540378443864676}; return t[m];
  }
};
template <> struct OddTab<5> {
  static __device__ __forceinline__ double c(int m) {
    const double t[5] = {1.0, 0.30901699437494742410, -0.80901699437494742410, -0.80901699437494742410, 0.30901699437494742410};
    return t[m];
  }
  static __device__ __forceinline__ double s(int m) {
    const double t[5] = {0.0, 0.95105651629515357212, 0.58778525229247312917, -0.58778525229247312917, -0.95105651629515357212};
    return t[m];
  }
};
template <> struct OddTab<7> {
  static __device__ __forceinline__ double c(int m) {
    const double t[7] = {1.0, 0.62348980185873353053, -0.22252093395631440429, -0.90096886790241912624,
                         -0.90096886790241912624, -0.22252093395631440429, 0.62348980185873353053};
    return t[m];
  }
  static __device__ __forceinline__ double s(int m) {
    const double t[7] = {0.0, 0.78183148246802980871, 0.97492791218182360702, 0.43388373911755812048,
                         -0.43388373911755812048, -0.97492791218182360702, -0.78183148246802980871};
    return t[m];
  }
};

template <int R, int DIR, typename T> struct DftOdd {
  static __device__ __forceinline__ void run(Cx<T>* a) {
    constexpr int H = (R - 1) / 2;
    Cx<T> t[H], u[H];
#pragma unroll
    for (int j = 1; j <= H; ++j) { t[j - 1] = a[j] + a[R - j]; u[j - 1] = a[j] - a[R - j]; }
    Cx<T> a0 = a[0];
    Cx<T> sum = a0;
#pragma unroll
    for (int j = 0; j < H; ++j) sum = sum + t[j];
    a[0] = sum;
#pragma unroll
    for (int k = 1; k <= H; ++k) {
      Cx<T> re = a0, im = mk<T>((T)0, (T)0);
#pragma unroll
      for (int j = 1; j <= H; ++j) {
        const int m = (j * k) % R;
        const T c = (T)OddTab<R>::c(m), s = (T)OddTab<R>::s(m);
        re.x += c * t[j - 1].x; re.y += c * t[j - 1].y;
        im.x += s * u[j - 1].x; im.y += s * u[j - 1].y;
      }
      Cx<T> iim = mul_i<DIR>(im);
      a[k] = re + iim;
      a[R - k] = re - iim;
    }
  }
};
template <int DIR, typename T> struct Dft<3, DIR, T> : DftOdd<3, DIR, T> {};
template <int DIR, typename T> struct Dft<5, DIR, T> : DftOdd<5, DIR, T> {};
template <int DIR, typename T> struct Dft<7, DIR, T> : DftOdd<7, DIR, T> {};

// ---- one pass over a batch of rows ------------------------------------------
template <typename T, int R, int DIR, bool DIT>
__device__ __forceinline__ void fft_pass(Cx<T>* __restrict__ s, int ld, int batch, int n, int M,
                                         const Cx<T>* __restrict__ tw) {
  const int sub = M / R, per = n / R, total = batch * per, tws = n / M;
  for (int w = threadIdx.x; w < total; w += blockDim.x) {
    const int b = w / per, rem = w - b * per;
    const int seg = rem / sub, q = rem - seg * sub;
    Cx<T>* row = s + (size_t)b * ld;
    const int base = seg * M + q;
    Cx<T> a[R];
#pragma unroll
    for (int i = 0; i < R; ++i) a[i] = row[padi<T>(base + i * sub)];
    if (DIT && q != 0) {
#pragma unroll
      for (int i = 1; i < R; ++i) a[i] = mul_tw<DIR>(a[i], ldg_cx<T>(tw + q * i * tws));
    }
    Dft<R, DIR, T>::run(a);
    if (!DIT && q != 0) {
#pragma unroll
      for (int o = 1; o < R; ++o) a[o] = mul_tw<DIR>(a[o], ldg_cx<T>(tw + q * o * tws));
    }
#pragma unroll
    for (int o = 0; o < R; ++o) row[padi<T>(base + o * sub)] = a[o];
  }
}

template <typename T, int DIR, bool DIT>
__device__ __forceinline__ void fft_pass_dispatch(int R, Cx<T>* s, int ld, int batch, int n, int M, const Cx<T>* tw) {
  switch (R) {
    case 2: fft_pass<T, 2, DIR, DIT>(s, ld, batch, n, M, tw); break;
    case 3: fft_pass<T, 3, DIR, DIT>(s, ld, batch, n, M, tw); break;
    case 4: fft_pass<T, 4, DIR, DIT>(s, ld, batch, n, M, tw); break;
    case 5: fft_pass<T, 5, DIR, DIT>(s, ld, batch, n, M, tw); break;
    case 7: fft_pass<T, 7, DIR, DIT>(s, ld, batch, n, M, tw); break;
    case 8: fft_pass<T, 8, DIR, DIT>(s, ld, batch, n, M, tw); break;
    default: break;
  }
}

// In-place batched FFT.  Caller must __syncthreads() before (data visible);
// the routine ends with a __syncthreads().  DIR=-1 forward, +1 inverse
// (un-normalised).
template <typename T, int DIR, bool DIT>
__device__ __forceinline__ void fft_smem(Cx<T>* s, int ld, int batch, const FftDev& p) {
  const Cx<T>* tw = reinterpret_cast<const Cx<T>*>(p.tw);
  if (!DIT) {
    int M = p.n;
    for (int l = 0; l < p.npass; ++l) {
      fft_pass_dispatch<T, DIR, false>(p.radix[l], s, ld, batch, p.n, M, tw);
      M /= p.radix[l];
      __syncthreads();
    }
  } else {
    int M = p.radix[p.npass - 1];
    for (int l = p.npass - 1; l >= 0; --l) {
      fft_pass_dispatch<T, DIR, true>(p.radix[l], s, ld, batch, p.n, M, tw);
      if (l > 0) M *= p.radix[l - 1];
      __syncthreads();
    }
  }
}

}  // namespace pfb
