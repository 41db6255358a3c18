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

// Tones and combs are exp(2 pi i f t) with one frequency per vector: t = t0 + 256 h + l gives
// exp(2 pi i f t) = A[h] * B[l], A[h] = exp(2 pi i frac(f (t0 + 256 h)) + i phase), B[l] = exp(2 pi i frac(f l)) (x the
// comb's period-32 envelope), so a CTA evaluates 2 x 256 double-precision sincos for its span of up to 65 536 samples
// and every sample costs one fp32 complex multiply (each factor is correctly rounded: the product is within 2 ulp of
// the double-precision evaluation the reference's numpy code does, generate_test_vector.py:40-48).
constexpr int kGenSpan = 65536;

template <typename T>
__global__ void __launch_bounds__(256) pfb_generate_batch_kernel(const GenBatchArgs a) {
  __shared__ double comb_r[32], comb_i[32];
  __shared__ float2 tabA[256], tabB[256];
  const int v = blockIdx.y;
  const double p0 = a.params[3 * v], p1 = a.params[3 * v + 1], p2 = a.params[3 * v + 2];
  Cx<T>* out = reinterpret_cast<Cx<T>*>(a.out) + (long long)(v / a.group) * a.group_stride + (v % a.group);
  if (a.kind == 0) {
    const long long o = (long long)p0, w = (long long)p1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < a.n; t += (long long)gridDim.x * blockDim.x) {
      Cx<T> val = mk<T>((T)0, (T)0);
      if (t >= o && t < o + w) val.x = (T)1;
      out[t * a.group] = val;
    }
    return;
  }
  if (a.kind == 2) {
    if (threadIdx.x < 32) comb_table(comb_r, comb_i);
    __syncthreads();
  }
  // cycles per sample
  const double f = a.kind == 1 ? (p0 + p2) / (double)a.n : -0.5 + 1.0 / 128.0 + p2 + p0 / (double)a.n;
  for (long long t0 = (long long)blockIdx.x * kGenSpan; t0 < a.n; t0 += (long long)gridDim.x * kGenSpan) {
    __syncthreads();
    {
      const int i = threadIdx.x;
      double cyc = f * (double)(t0 + 256ll * i);
      double sn, cs;
      sincos(2.0 * 3.14159265358979323846 * (cyc - floor(cyc)) + p1, &sn, &cs);
      tabA[i] = make_float2((float)cs, (float)sn);
      cyc = f * (double)i;
      sincos(2.0 * 3.14159265358979323846 * (cyc - floor(cyc)), &sn, &cs);
      if (a.kind == 2) {
        const int m = (int)((t0 + i) & 31);              // t0 is a multiple of 256: (t0 + 256 h + l) mod 32 = l mod 32
        const double br = cs * comb_r[m] - sn * comb_i[m], bi = cs * comb_i[m] + sn * comb_r[m];
        cs = br;
        sn = bi;
      }
      tabB[i] = make_float2((float)cs, (float)sn);
    }
    __syncthreads();
    const long long span = min((long long)kGenSpan, a.n - t0);
    for (long long d = threadIdx.x; d < span; d += blockDim.x) {
      const float2 A = tabA[d >> 8], B = tabB[d & 255];
      out[(t0 + d) * a.group] = mk<T>((T)(A.x * B.x - A.y * B.y), (T)(A.x * B.y + A.y * B.x));
    }
  }
}

// Pair form (complex64, group == 2): one CTA column generates BOTH vectors that share a [n][2] series and writes whole
// 16-byte elements; the one-vector kernel above writes 8 of every 16 bytes, i.e. half of every 32-byte sector.
constexpr int kPairSpan = 16384;   // samples per CTA pass of the pair generator (64 + 256 table entries per vector)
__global__ void __launch_bounds__(256) pfb_generate_pairs_kernel(const GenBatchArgs a) {
  __shared__ double comb_r[32], comb_i[32];
  __shared__ float2 tabA[2][kPairSpan / 256], tabB[2][256];
  const int pr = blockIdx.y;
  float4* out = reinterpret_cast<float4*>(reinterpret_cast<Cx<float>*>(a.out) + (long long)pr * a.group_stride);
  double p0[2], p1[2], p2[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    p0[s] = a.params[3 * (2 * pr + s)];
    p1[s] = a.params[3 * (2 * pr + s) + 1];
    p2[s] = a.params[3 * (2 * pr + s) + 2];
  }
  if (a.kind == 0) {
    const long long o0 = (long long)p0[0], w0 = (long long)p1[0], o1 = (long long)p0[1], w1 = (long long)p1[1];
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < a.n; t += (long long)gridDim.x * blockDim.x)
      out[t] = make_float4((t >= o0 && t < o0 + w0) ? 1.f : 0.f, 0.f, (t >= o1 && t < o1 + w1) ? 1.f : 0.f, 0.f);
    return;
  }
  if (a.kind == 2) {
    if (threadIdx.x < 32) comb_table(comb_r, comb_i);
    __syncthreads();
  }
  double f[2];
#pragma unroll
  for (int s = 0; s < 2; ++s)
    f[s] = a.kind == 1 ? (p0[s] + p2[s]) / (double)a.n : -0.5 + 1.0 / 128.0 + p2[s] + p0[s] / (double)a.n;
  for (long long t0 = (long long)blockIdx.x * kPairSpan; t0 < a.n; t0 += (long long)gridDim.x * kPairSpan) {
    __syncthreads();
    {
      const int i = threadIdx.x;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        double cyc = f[s] * (double)(t0 + 256ll * i);
        double sn, cs;
        if (i < kPairSpan / 256) {
          sincos(2.0 * 3.14159265358979323846 * (cyc - floor(cyc)) + p1[s], &sn, &cs);
          tabA[s][i] = make_float2((float)cs, (float)sn);
        }
        cyc = f[s] * (double)i;
        sincos(2.0 * 3.14159265358979323846 * (cyc - floor(cyc)), &sn, &cs);
        if (a.kind == 2) {
          const int m = (int)((t0 + i) & 31);              // t0 is a multiple of 256: (t0 + 256 h + l) mod 32 = l mod 32
          const double br = cs * comb_r[m] - sn * comb_i[m], bi = cs * comb_i[m] + sn * comb_r[m];
          cs = br;
          sn = bi;
        }
        tabB[s][i] = make_float2((float)cs, (float)sn);
      }
    }
    __syncthreads();
    const long long span = min((long long)kPairSpan, a.n - t0);
    for (long long d = threadIdx.x; d < span; d += blockDim.x) {
      const float2 A0 = tabA[0][d >> 8], B0 = tabB[0][d & 255], A1 = tabA[1][d >> 8], B1 = tabB[1][d & 255];
      out[t0 + d] = make_float4(A0.x * B0.x - A0.y * B0.y, A0.x * B0.y + A0.y * B0.x, A1.x * B1.x - A1.y * B1.y,
                                A1.x * B1.y + A1.y * B1.x);
    }
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

// one CTA per record (grid = number of records; args array indexed by blockIdx.x).  The per-element arithmetic is
// fp32 (the series are fp32); the power sum is accumulated in double, the |.| sums as fp32 partial sums flushed into
// doubles every 32 elements, and the reduction across threads is in double.
struct Top2f {
  float m1, m2;
  long long i1;
};
template <typename T>
__global__ void __launch_bounds__(1024) pfb_metrics_kernel(const MetricsArgs* args) {
  const MetricsArgs a = args[blockIdx.x];
  const Cx<T>* pa = reinterpret_cast<const Cx<T>*>(a.a);
  const Cx<T>* pb = reinterpret_cast<const Cx<T>*>(a.b);
  double sdiff = 0, mdiff = 0, spow = 0, sabs = 0;
  Top2 t2{-1.0, -1.0, 0};
  const float sc = (float)a.scale_b;
  for (long long i0 = threadIdx.x; i0 < a.n; i0 += 32ll * blockDim.x) {
    float fdiff = 0.f, fmaxd = 0.f, fabs_ = 0.f;
    Top2f f2{-1.f, -1.f, 0};
#pragma unroll 4
    for (int u = 0; u < 32; ++u) {
      const long long i = i0 + (long long)u * blockDim.x;
      if (i >= a.n) break;
      const Cx<T> vb = pb[i * a.stride_b];
      const float br = (float)vb.x * sc, bi = (float)vb.y * sc;
      const float p = br * br + bi * bi;
      spow += (double)p;       // exact sum: the report subtracts the largest bin from it (total spurious ~ 1e-8 of the sum)
      fabs_ += sqrtf(p);
      if (p > f2.m1) { f2.m2 = f2.m1; f2.m1 = p; f2.i1 = i; }
      else if (p > f2.m2) f2.m2 = p;
      if (pa) {
        const Cx<T> va = pa[i * a.stride_a];
        const float dr = br - (float)va.x * sc, di = bi - (float)va.y * sc;
        const float d = sqrtf(dr * dr + di * di);
        fdiff += d;
        fmaxd = fmaxf(fmaxd, d);
      }
    }
    sdiff += (double)fdiff;
    mdiff = fmax(mdiff, (double)fmaxd);
    sabs += (double)fabs_;
    Top2 one{(double)f2.m1, (double)f2.m2, f2.i1};
    t2 = merge(t2, one);
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

// Pair form used by the batched sweep: the two vectors that ride in the polarisation slots of one [n][2] series are
// reduced together, so every 32-byte sector read is fully used (the single-series kernel above touches every sector
// for 8 of its 16 bytes).  out0 / out1 receive the same 8 numbers per slot as pfb_metrics_kernel.  gridDim.y CTAs share a
// record: CTA (x, y) reduces elements [n y / Y, n (y + 1) / Y) and writes its 8 numbers at out + 8 y; the caller merges
// the Y partial rows (sums add, maxima and the top two are merged; argmax is an index into the whole series).
struct MetricsPairArgs {
  const void* a;        // [n][2] complex64 reference series (may be null)
  const void* b;        // [n][2] complex64 series under test
  long long n;
  double scale_b;
  double* out0;
  double* out1;
};

__global__ void __launch_bounds__(1024) pfb_metrics_pair_kernel(const MetricsPairArgs* args) {
  const MetricsPairArgs a = args[blockIdx.x];
  const float4* pa = reinterpret_cast<const float4*>(a.a);
  const float4* pb = reinterpret_cast<const float4*>(a.b);
  double sdiff[2] = {0, 0}, mdiff[2] = {0, 0}, spow[2] = {0, 0}, sabs[2] = {0, 0};
  Top2 t2[2] = {{-1.0, -1.0, 0}, {-1.0, -1.0, 0}};
  const float sc = (float)a.scale_b;
  const long long e0 = a.n * blockIdx.y / gridDim.y, e1 = a.n * (blockIdx.y + 1) / gridDim.y;
  for (long long i0 = e0 + threadIdx.x; i0 < e1; i0 += 32ll * blockDim.x) {
    float fdiff[2] = {0.f, 0.f}, fmaxd[2] = {0.f, 0.f}, fabs_[2] = {0.f, 0.f};
    Top2f f2[2] = {{-1.f, -1.f, 0}, {-1.f, -1.f, 0}};
#pragma unroll 4
    for (int u = 0; u < 32; ++u) {
      const long long i = i0 + (long long)u * blockDim.x;
      if (i >= e1) break;
      const float4 vb = __ldg(pb + i);
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pa) va = __ldg(pa + i);
      const float br[2] = {vb.x * sc, vb.z * sc}, bi[2] = {vb.y * sc, vb.w * sc};
      const float ar[2] = {va.x * sc, va.z * sc}, ai[2] = {va.y * sc, va.w * sc};
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const float p = br[s] * br[s] + bi[s] * bi[s];
        spow[s] += (double)p;
        fabs_[s] += sqrtf(p);
        if (p > f2[s].m1) { f2[s].m2 = f2[s].m1; f2[s].m1 = p; f2[s].i1 = i; }
        else if (p > f2[s].m2) f2[s].m2 = p;
        if (pa) {
          const float dr = br[s] - ar[s], di = bi[s] - ai[s];
          const float d = sqrtf(dr * dr + di * di);
          fdiff[s] += d;
          fmaxd[s] = fmaxf(fmaxd[s], d);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      sdiff[s] += (double)fdiff[s];
      mdiff[s] = fmax(mdiff[s], (double)fmaxd[s]);
      sabs[s] += (double)fabs_[s];
      Top2 one{(double)f2[s].m1, (double)f2[s].m2, f2[s].i1};
      t2[s] = merge(t2[s], one);
    }
  }
  __shared__ double sh[2][6][32];
  __shared__ long long shi[2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    for (int o = 16; o > 0; o >>= 1) {
      sdiff[s] += __shfl_down_sync(0xffffffff, sdiff[s], o);
      mdiff[s] = fmax(mdiff[s], __shfl_down_sync(0xffffffff, mdiff[s], o));
      spow[s] += __shfl_down_sync(0xffffffff, spow[s], o);
      sabs[s] += __shfl_down_sync(0xffffffff, sabs[s], o);
      Top2 y{__shfl_down_sync(0xffffffff, t2[s].m1, o), __shfl_down_sync(0xffffffff, t2[s].m2, o),
             __shfl_down_sync(0xffffffff, t2[s].i1, o)};
      t2[s] = merge(t2[s], y);
    }
    if (lane == 0) {
      sh[s][0][warp] = sdiff[s]; sh[s][1][warp] = mdiff[s]; sh[s][2][warp] = spow[s]; sh[s][3][warp] = t2[s].m1;
      sh[s][4][warp] = t2[s].m2; sh[s][5][warp] = sabs[s]; shi[s][warp] = t2[s].i1;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    const int s = threadIdx.x;
    const int nw = (blockDim.x + 31) / 32;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    Top2 t{-1.0, -1.0, 0};
    for (int w = 0; w < nw; ++w) {
      s0 += sh[s][0][w]; s1 = fmax(s1, sh[s][1][w]); s2 += sh[s][2][w]; s3 += sh[s][5][w];
      Top2 y{sh[s][3][w], sh[s][4][w], shi[s][w]};
      t = merge(t, y);
    }
    double* out = (s ? a.out1 : a.out0) + 8 * blockIdx.y;
    out[0] = s0; out[1] = s1; out[2] = s2; out[3] = t.m1; out[4] = t.m2 < 0 ? 0.0 : t.m2;
    out[5] = (double)t.i1; out[6] = s3; out[7] = (double)(e1 - e0);
  }
}

}  // namespace pfb
