// Test-vector synthesis and purity reductions on the device:
//   K6 pfb_generate_kernel : time_domain_impulse / complex_sinusoid of
//      python/data_gen/generate_test_vector.py:24-71 written straight into the analysis input
//   K5 pfb_metrics_kernel  : the reductions of python/verify/util.py:15-43 and purity.py:154-160
//      (|x|^2, zero the maximum, max / total spurious; |b - a| sum) in one pass per vector
#pragma once
#include "pfb_common.cuh"

namespace pfb {

struct GenArgs {
  void* out;            // [n][stride] complex; the first npol entries of every row are written
  long long n;
  long long stride;
  int npol, kind;       // 0 impulse, 1 tone, 2 frequency comb
  double p0, p1, p2;    // impulse: offset, width;  tone: frequency (bins), phase (rad), bin offset;
                        // comb: frequency shift (bins of n), phase (rad), extra offset (cycles/sample)
};

// Frequency comb of matlab/FrequencyComb.m:22-46 with the parameters of sgcht.m:394-432: 32 tones, amplitudes
// linspace(1, sqrt 2, 32), f_k = fmin + k/32 cycles per sample, fmin = -0.5 + 1/128 (+ offsets).  The tones sit 1/32
// apart, so x[t] = exp(2 pi i f0 t) * P[t mod 32] with P[m] = sum_k a_k exp(2 pi i k m / 32): one sincos per sample.
__device__ __forceinline__ void comb_table(double* pr, double* pi) {      // called by threads 0..31, fills P[32]
  const int m = threadIdx.x;
  double sr = 0, si = 0;
  for (int k = 0; k < 32; ++k) {
    const double amp = 1.0 + (1.4142135623730951 - 1.0) * (double)k / 31.0;
    double sn, cs;
    sincospi(2.0 * (double)((k * m) & 31) / 32.0, &sn, &cs);
    sr += amp * cs;
    si += amp * sn;
  }
  pr[m] = sr;
  pi[m] = si;
}
template <typename T>
__device__ __forceinline__ Cx<T> comb_sample(long long t, long long n, double p0, double p1, double p2, const double* pr,
                                             const double* pi) {
  const double f0 = -0.5 + 1.0 / 128.0 + p2 + p0 / (double)n;
  const double cyc = f0 * (double)t;
  double sn, cs;
  sincos(2.0 * 3.14159265358979323846 * (cyc - floor(cyc)) + p1, &sn, &cs);
  const int m = (int)(t & 31);
  return mk<T>((T)(cs * pr[m] - sn * pi[m]), (T)(cs * pi[m] + sn * pr[m]));
}

template <typename T>
__global__ void __launch_bounds__(256) pfb_generate_kernel(const GenArgs a) {
  __shared__ double comb_r[32], comb_i[32];
  if (a.kind == 2) {
    if (threadIdx.x < 32) comb_table(comb_r, comb_i);
    __syncthreads();
  }
  Cx<T>* out = reinterpret_cast<Cx<T>*>(a.out);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < a.n; t += (long long)gridDim.x * blockDim.x) {
    Cx<T> v = mk<T>((T)0, (T)0);
    if (a.kind == 0) {
      const long long o = (long long)a.p0, w = (long long)a.p1;
      if (t >= o && t < o + w) v.x = (T)1;
    } else if (a.kind == 2) {
      v = comb_sample<T>(t, a.n, a.p0, a.p1, a.p2, comb_r, comb_i);
    } else {
      // exp(i (2 pi (f + bin_offset) t / n + phase)), evaluated in double like numpy
      const double cyc = (a.p0 + a.p2) * (double)t / (double)a.n;
      const double ang = 2.0 * 3.14159265358979323846 * (cyc - floor(cyc)) + a.p1;
      double s, c;
      sincos(ang, &s, &c);
      v.x = (T)c;
      v.y = (T)s;
    }
    for (int p = 0; p < a.npol; ++p) out[t * a.stride + p] = v;
  }
}

// many vectors in one launch: blockIdx.y = vector; parameters come from a device array so that a captured
// CUDA graph of a whole batch can be replayed with new parameters
struct GenBatchArgs {
  void* out;
  long long n;
  long long group_stride;   // elements between consecutive groups of `group` interleaved vectors
  int group;                // vectors interleaved sample by sample (2: the polarisation slots of one stream)
  int kind;
  const double* params;     // [nvec][3]
};

template <typename T>
__global__ void __launch_bounds__(256) pfb_generate_batch_kernel(const GenBatchArgs a) {
  __shared__ double comb_r[32], comb_i[32];
  if (a.kind == 2) {
    if (threadIdx.x < 32) comb_table(comb_r, comb_i);
    __syncthreads();
  }
  const int v = blockIdx.y;
  const double p0 = a.params[3 * v], p1 = a.params[3 * v + 1], p2 = a.params[3 * v + 2];
  Cx<T>* out = reinterpret_cast<Cx<T>*>(a.out) + (long long)(v / a.group) * a.group_stride + (v % a.group);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < a.n; t += (long long)gridDim.x * blockDim.x) {
    Cx<T> val = mk<T>((T)0, (T)0);
    if (a.kind == 0) {
      const long long o = (long long)p0, w = (long long)p1;
      if (t >= o && t < o + w) val.x = (T)1;
    } else if (a.kind == 2) {
      val = comb_sample<T>(t, a.n, p0, p1, p2, comb_r, comb_i);
    } else {
      const double cyc = (p0 + p2) * (double)t / (double)a.n;
      const double ang = 2.0 * 3.14159265358979323846 * (cyc - floor(cyc)) + p1;
      double sn, cs;
      sincos(ang, &sn, &cs);
      val.x = (T)cs;
      val.y = (T)sn;
    }
    out[t * a.group] = val;
  }
}

struct MetricsArgs {
  const void* a;        // reference series (may be null), complex, element stride `stride`
  const void* b;        // series under test
  long long n;
  long long stride_a, stride_b;
  double scale_b;       // b is multiplied by this before use (purity.py divides by fft_size)
  double* out;          // [8]: sum|b-a|, max|b-a|, sum p, max1 p, max2 p, argmax, sum|b|, n
};

struct Top2 {
  double m1, m2;
  long long i1;
};
__device__ __forceinline__ Top2 merge(Top2 x, Top2 y) {
  Top2 r;
  if (x.m1 > y.m1 || (x.m1 == y.m1 && x.i1 <= y.i1)) { r.m1 = x.m1; r.i1 = x.i1; r.m2 = fmax(x.m2, y.m1); }
  else { r.m1 = y.m1; r.i1 = y.i1; r.m2 = fmax(y.m2, x.m1); }
  return r;
}

// one CTA per vector (grid = number of vectors; args array indexed by blockIdx.x)
template <typename T>
__global__ void __launch_bounds__(1024) pfb_metrics_kernel(const MetricsArgs* args) {
  const MetricsArgs a = args[blockIdx.x];
  const Cx<T>* pa = reinterpret_cast<const Cx<T>*>(a.a);
  const Cx<T>* pb = reinterpret_cast<const Cx<T>*>(a.b);
  double sdiff = 0, mdiff = 0, spow = 0, sabs = 0;
  Top2 t2{-1.0, -1.0, 0};
  for (long long i = threadIdx.x; i < a.n; i += blockDim.x) {
    const Cx<T> vb = pb[i * a.stride_b];
    const double br = (double)vb.x * a.scale_b, bi = (double)vb.y * a.scale_b;
    const double p = br * br + bi * bi;
    spow += p;
    sabs += sqrt(p);
    Top2 one{p, -1.0, i};
    t2 = merge(t2, one);
    if (pa) {
      const Cx<T> va = pa[i * a.stride_a];
      const double dr = br - (double)va.x * a.scale_b, di = bi - (double)va.y * a.scale_b;
      const double d = sqrt(dr * dr + di * di);
      sdiff += d;
      mdiff = fmax(mdiff, d);
    }
  }
  __shared__ double sh[5][32];
  __shared__ long long shi[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    sdiff += __shfl_down_sync(0xffffffff, sdiff, o);
    mdiff = fmax(mdiff, __shfl_down_sync(0xffffffff, mdiff, o));
    spow += __shfl_down_sync(0xffffffff, spow, o);
    sabs += __shfl_down_sync(0xffffffff, sabs, o);
    Top2 y{__shfl_down_sync(0xffffffff, t2.m1, o), __shfl_down_sync(0xffffffff, t2.m2, o),
           __shfl_down_sync(0xffffffff, t2.i1, o)};
    t2 = merge(t2, y);
  }
  if (lane == 0) { sh[0][warp] = sdiff; sh[1][warp] = mdiff; sh[2][warp] = spow; sh[3][warp] = t2.m1; sh[4][warp] = t2.m2; shi[warp] = t2.i1; }
  __shared__ double shabs[32];
  if (lane == 0) shabs[warp] = sabs;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) / 32;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    Top2 t{-1.0, -1.0, 0};
    for (int w = 0; w < nw; ++w) {
      s0 += sh[0][w]; s1 = fmax(s1, sh[1][w]); s2 += sh[2][w]; s3 += shabs[w];
      Top2 y{sh[3][w], sh[4][w], shi[w]};
      t = merge(t, y);
    }
    a.out[0] = s0; a.out[1] = s1; a.out[2] = s2; a.out[3] = t.m1; a.out[4] = t.m2 < 0 ? 0.0 : t.m2;
    a.out[5] = (double)t.i1; a.out[6] = s3; a.out[7] = (double)a.n;
  }
}

}  // namespace pfb
