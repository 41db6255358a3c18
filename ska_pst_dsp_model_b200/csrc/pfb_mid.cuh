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
      if (r + gridDim.x < a.nrows) {
        // the CTA's next row into L2 while this one is transformed (the loads of pass A are the kernel's longest stall)
        const char* nx = reinterpret_cast<const char*>(in + (r + gridDim.x) * 4096) + t * 256;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 128));
      }
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

// ---------------------------------------------------------------------------
// K3 for the SKA-Mid inversion shape (polyphase_synthesis.m:163-316 at config "mid": Nf = 512, overlap 128, OS 8/7):
// per column (t1, pol): taper -> 512-point FFT -> keep the 448 centre bins x W[t1][r] -> 448-point inverse FFT ->
// keep the middle 224 outputs.  Same scheme as pfb_resample_low: 16 lanes per transform, lanes of a warp = 16
// different columns (one 128-byte row segment per load / store), in-register butterflies with compile-time twiddles,
// two exchanges through shared memory per column-block:
//   512 = 32 x 16: lane q holds tau = q + 16 i (32 points) -> radix-32 -> x W512^(q o) -> exchange ->
//                  lane l runs the radix-16 transforms of o = l and l + 16: bins b = o + 32 o2;
//   kept bins are o2 in {0..6} U {9..15} for EVERY o, and their inverse-transform index is i' = o + 32 k, k = 0..13,
//   so the 14-point inverse DFTs over k need no exchange (448 = 14 x 32);
//   x W448^(o m) -> exchange -> lane m < 14 runs one radix-32 inverse: y[m + 14 j], kept j in [8, 24).
// One CTA = 16 columns (8 t1 x 2 pol) x one overlap-save block; weights come straight from the 14.7 MB table
// (L2-resident, every CTA of a column group re-reads the same 28 KB).
#ifndef PFB_MID_COLS
#define PFB_MID_COLS 16
#endif
constexpr int kMidCols = PFB_MID_COLS;
constexpr int kMidStrideA = 32 * 17 + 1;     // [o (32)][q (16)] padded: odd pitch per column
constexpr int kMidStrideB = 14 * 33 + 1;     // [m (14)][o (32)]

struct MidResArgs {
  const C* v;           // [tau][nchan][2]
  C* out;               // [T2][nchan][2]
  const float* taper;   // [512]
  const C* wt;          // [nchan][448] in FN order
  const float2* tw512;  // [512] exp(-2 pi i k/512)
  const float2* tw448;  // [448] exp(-2 pi i k/448) (conjugated on load: the inverse transform)
  int nchan, nblk, nsplit;
};

constexpr size_t mid_res_smem_bytes() {
  return (size_t)kMidCols * kMidStrideA * sizeof(C) + (512 + 448) * sizeof(C) + 512 * sizeof(float);
}

// UNITY_MID: the temporal taper is exactly 1 on [ov, Nf - ov) (Tukey, top-hat interior, no window): only the 16 edge
// samples of a lane are scaled.
template <bool UNITY_MID>
__global__ void __launch_bounds__(16 * kMidCols, 512 / (16 * kMidCols)) pfb_resample_mid(const MidResArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* ex = reinterpret_cast<C*>(smem_raw);              // [COLS][kMidStrideA], reused for the second exchange
  C* t1s = ex + kMidCols * kMidStrideA;                // [32 o][16 q]  exp(-2 pi i q o / 512)
  C* t2s = t1s + 512;                                  // [14 m][32 o]  exp(+2 pi i o m / 448)
  float* stp = reinterpret_cast<float*>(t2s + 448);
  const int tid = threadIdx.x;
  const int c = tid % kMidCols, q = tid / kMidCols;    // column (t1 pair, pol), lane-in-transform
  const int groups = a.nchan / (kMidCols / 2);
  const int g = blockIdx.x % groups, sp = blockIdx.x / groups;
  const int per = (a.nblk + a.nsplit - 1) / a.nsplit;
  const int n_begin = sp * per, n_end = min(a.nblk, n_begin + per);
  const int t1_0 = g * (kMidCols / 2);
  const long long rowlen = (long long)a.nchan * 2;

  for (int i = tid; i < 512; i += 16 * kMidCols) {
    stp[i] = a.taper[i];
    const float2 w = __ldg(a.tw512 + (((i >> 4) * (i & 15)) & 511));
    t1s[i] = mk<float>(w.x, w.y);
  }
  for (int i = tid; i < 448; i += 16 * kMidCols) {
    const float2 w = __ldg(a.tw448 + (((i >> 5) * (i & 31)) % 448));
    t2s[i] = mk<float>(w.x, -w.y);
  }
  __syncthreads();

  C* e = ex + c * kMidStrideA;
  // a CTA walks a run of consecutive blocks: the 256 rows two blocks share are read again a few microseconds later
  // and come from L2 (the register file has no room for a 16-element carry next to the 32-point butterflies)
  for (long long n = n_begin; n < n_end; ++n) {
  const C* vcol = a.v + (long long)t1_0 * 2 + c + (n * 256 + q) * rowlen;
  if (n + 1 < n_end) {
    // the 256 rows the NEXT block adds (one 128-byte segment per thread) into L2 while this block is transformed
    const C* nx = a.v + (long long)t1_0 * 2 + ((n + 1) * 256 + 256 + tid % 256) * rowlen;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
  }
  C x[32];
  // phase 1: lane q holds tau = q + 16 i
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    x[i] = low::ld_nc(vcol + (long long)i * 16 * rowlen);
    if (!UNITY_MID || i < 8 || i >= 24) x[i] = x[i] * stp[q + 16 * i];       // ov = 128 = 8 * 16
  }
  fft_reg<4, 8, -1>(x);
#pragma unroll
  for (int o = 0; o < 32; ++o) {
    C u = x[slot<4, 8>(o)];
    if (o) u = u * t1s[o * 16 + q];
    e[o * 17 + q] = u;
  }
  __syncthreads();
  // phase 2: lane l: o = l and l + 16; radix-16 over q -> bins b = o + 32 o2; keep o2 in {0..6} U {9..15}
  C z[28];
  const C* wrow = a.wt + (size_t)(t1_0 + (c >> 1)) * 448;      // (staging the 28 KB in shared memory was measured: 5 % slower)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int o = q + 16 * h;
    C y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = e[o * 17 + i];
    fft_reg<4, 4, -1>(y);
#pragma unroll
    for (int k = 0; k < 14; ++k) {
      const int o2 = k < 7 ? k : k + 2;
      const int b = o + 32 * o2;                           // FFT bin
      const int rr = b < 224 ? b + 224 : b - 288;          // FN-order index of that bin (fftshift + discard)
      z[14 * h + k] = y[slot<4, 4>(o2)] * low::ld_nc(wrow + rr);
    }
  }
  __syncthreads();                                         // all lanes done reading exchange A
  // z[14 h + k] sits at inverse index i' = o + 32 k: 14-point inverse DFT over k, twiddle, second exchange
  C* eb = ex + c * kMidStrideA;                            // [m][o] pitch 33
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int o = q + 16 * h;
    fft_reg<2, 7, +1>(z + 14 * h);
#pragma unroll
    for (int m = 0; m < 14; ++m) {
      C u = z[14 * h + slot<2, 7>(m)];
      if (m) u = u * t2s[m * 32 + o];
      eb[m * 33 + o] = u;
    }
  }
  __syncthreads();
  if (q < 14) {
    // phase 3: lane m = q: y[m + 14 j] = sum_o eb[m][o] exp(+2 pi i o j / 32); keep j in [8, 24)
    const int m = q;
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = eb[m * 33 + i];
    fft_reg<4, 8, +1>(x);
    C* p = a.out + (long long)t1_0 * 2 + c + (n * 224 - 112 + m) * rowlen;
#pragma unroll
    for (int j = 8; j < 24; ++j) p[(long long)14 * j * rowlen] = x[slot<4, 8>(j)];
  }
  __syncthreads();                                         // exchange buffer free for the next block
  }
}

inline int resample_mid_launch(const MidResArgs& a, bool unity_mid, cudaStream_t st) {
  auto kern = unity_mid ? pfb_resample_mid<true> : pfb_resample_mid<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mid_res_smem_bytes());
  if (e != cudaSuccess) return (int)e;
  // time splits: enough CTAs for ~4 resident waves (2 CTAs per SM), each walking a contiguous run of blocks
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long groups = a.nchan / (kMidCols / 2);
  long long ns = (8ll * sms + groups - 1) / groups;
  if (ns < 1) ns = 1;
  if (ns > a.nblk) ns = a.nblk;
  MidResArgs b = a;
  b.nsplit = (int)ns;
  const long long grid = ns * groups;
  if (grid > 2147483647ll) return (int)cudaErrorInvalidValue;
  kern<<<(unsigned)grid, 16 * kMidCols, mid_res_smem_bytes(), st>>>(b);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Long power-of-two-times-24 spectra of dual-pol series (the 2 * block_size = 98 304 = 24 x 4096 point transforms of
// the purity report, python/verify/purity.py:150,236-263, DomainPerformance.m:58-100): four-step FFT,
//   X[k1 + 24 k2] = sum_n2 W4096^(n2 k2) * [ W_N^(n2 k1) * sum_n1 x[4096 n1 + n2] W24^(n1 k1) ],
// step 1 (this kernel): the 24-point DFTs across n1 (3 x 8 in registers) and the twiddle, rows written as
// [k1][n2]; step 2: pfb_rowfft4096 over the 24 rows.  The result is stored transposed, spec[k1][k2] = X[k1 + 24 k2]:
// the report's reductions (sums, maxima, |b - a|) do not depend on the bin order.
struct SpecArgs {
  const float4* in;       // [pair][pitch] (pol0.re, pol0.im, pol1.re, pol1.im)
  float4* s1;             // [pair][24][4096]
  long long pitch;        // float4 units between consecutive pairs
};

__global__ void __launch_bounds__(256) pfb_spectrum24_step1(const SpecArgs a) {
  constexpr int M = 24, N = M * 4096;
  const int n2 = blockIdx.x * 256 + threadIdx.x;
  const long long pair = blockIdx.y;
  const float4* p = a.in + pair * a.pitch + n2;
  C x0[M], x1[M];
#pragma unroll
  for (int n1 = 0; n1 < M; ++n1) {
    const float4 v = __ldg(p + (long long)n1 * 4096);
    x0[n1] = mk<float>(v.x, v.y);
    x1[n1] = mk<float>(v.z, v.w);
  }
  fft_reg<3, 8, -1>(x0);
  fft_reg<3, 8, -1>(x1);
  float4* q = a.s1 + pair * (long long)N + n2;
#pragma unroll
  for (int k1 = 0; k1 < M; ++k1) {
    C u0 = x0[slot<3, 8>(k1)], u1 = x1[slot<3, 8>(k1)];
    if (k1) {
      float sn, cs;
      sincospif(-2.0f * (float)((n2 * k1) % N) / (float)N, &sn, &cs);      // n2 k1 < N: exact integer in fp32
      const C w = mk<float>(cs, sn);
      u0 = u0 * w;
      u1 = u1 * w;
    }
    q[(long long)k1 * 4096] = make_float4(u0.x, u0.y, u1.x, u1.y);
  }
}

// ---------------------------------------------------------------------------
// Sliding-window polyphase FIR with the block phases sharing their input (fp32, dual-pol, long filters).
//
// nu*S = de*N, so block k = k0 + nu*i of output arm j reads the sub-stream of SAMPLE arm
// a = (j - k0*(N - S)) mod N at rows Q = s + de*i + m, s = k0 - [a + k0*(N - S) >= N]: all nu phases walk the
// same rows, s rows apart.  pfb_fir_slide gives every phase its own CTA, so the input crosses L2 -> SM nu
// times; here one CTA owns 16 sample arms x 2 pols x nu phases, the rows (16 arms x 2 pols = 256 B each, N
// samples apart) are fetched ONCE by the TMA engine (one 2-D tensor copy per group of `de` rows; rows before
// the stream are zero-filled by the tensor map) into a shared ring, and every thread takes its `de` new
// window elements per step from there.  Taps and the circular register window are as in pfb_fir_slide.
// Requires the window start of block 0 to fall on a multiple of N samples.
// ---------------------------------------------------------------------------
#ifndef PFB_PH_ARMS
#define PFB_PH_ARMS 16
#endif
constexpr int kPhArms = PFB_PH_ARMS;                   // sample arms per CTA (x 2 pols x nu phases compute threads)
constexpr int kPhGroups = 12;                          // ring depth in row groups
constexpr int kPhLag = 3;                              // a group is refilled kPhLag steps after thread 0 released it

struct FirPhasesArgs {
  void* y;               // [k][slot][2]
  const float* taps;     // [P*N]
  int N, S, nu;
  long long start0, phase0, nblocks;
  long long i0, i1;      // step range of this launch; a step = one block of every phase
  long long steps_per_cta;
  long long row0;        // start0 / N (exact): tensor row of sub-stream row Q = 0
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   low::smem_u32(dst)),
               "l"(tm), "r"(x), "r"(y), "r"(low::smem_u32(bar))
               : "memory");
}

template <int P, int DE, int NU>
__global__ void __launch_bounds__(2 * kPhArms * NU, kPhArms <= 16 ? 2 : 1)
    pfb_fir_phases(const FirPhasesArgs a, const __grid_constant__ CUtensorMap in_map) {
  static_assert(DE < P, "long filters only");
  constexpr int NEW0 = P - DE;
  constexpr int W = (P + DE - 1) / DE * DE;
  constexpr int PER = W / DE;
  constexpr int NWARP = 2 * kPhArms * NU / 32;                    // compute warps
  constexpr int kRowC = kPhArms * 2;                              // complex elements per row
  constexpr uint32_t kGroupBytes = DE * kRowC * sizeof(C);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* ring = reinterpret_cast<C*>(smem_raw);                       // [kPhGroups + 1][DE][64]: the last group mirrors group 0
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (kPhGroups + 1) * DE * kRowC);
  uint64_t* empty = full + kPhGroups;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int a0 = blockIdx.x * kPhArms;                            // first sample arm of the CTA
  const long long nsteps_max = (a.nblocks + a.nu - 1) / a.nu;
  const long long i_begin = a.i0 + (long long)blockIdx.y * a.steps_per_cta;
  const long long i_end = min(min(i_begin + a.steps_per_cta, a.i1), nsteps_max);
  if (i_begin >= i_end) return;
  if (tid == 0) {
    for (int s = 0; s < kPhGroups; ++s) { low::mbar_init(&full[s], 1); low::mbar_init(&empty[s], NWARP); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // virtual steps v = i_begin - (PER - 1) .. i_end - 1: the first PER - 1 only fill the register window
  const long long v0 = i_begin - (PER - 1);
  // row groups are numbered from the first one any thread reads: group gi holds rows DE*(g0 + gi) .. + DE - 1,
  // g0 = v0 + NEW0 / DE (may be negative: rows before the stream, zero-filled)
  const long long g0 = v0 + NEW0 / DE;
  const long long ngroups = ((NU - 1) + (long long)DE * (i_end - 1 - v0) + NEW0 % DE + DE - 1) / DE + 1;

  // Loads are issued by thread 0 from inside its own step loop (no producer warp: 256 threads leave 128 registers per
  // thread, the 288-thread form was capped at 96 and spilled): group gN goes into the slot of group gN - G, which every
  // warp released kPhLag - 1 steps before thread 0 asks for it, so thread 0 only blocks when a warp lags that far.
  int gN = 0, sN = 0;                                             // next group to issue, its slot
  uint32_t parN = 1;                                              // parity of the slot's previous release (first lap: none)
  auto issue_next = [&]() {
    if (gN >= kPhGroups) low::mbar_wait(&empty[sN], parN);
    low::mbar_expect_tx(&full[sN], sN == 0 ? 2 * kGroupBytes : kGroupBytes);
    tma_load_2d(ring + sN * DE * kRowC, &in_map, a0 * 4, (int)(a.row0 + (g0 + gN) * DE), &full[sN]);
    // slot 0 is written twice (also behind the last slot) so that a thread's DE consecutive rows never wrap
    if (sN == 0) tma_load_2d(ring + kPhGroups * DE * kRowC, &in_map, a0 * 4, (int)(a.row0 + (g0 + gN) * DE), &full[sN]);
    ++gN;
    if (++sN == kPhGroups) { sN = 0; parN ^= 1u; }
  };
  const int ngroups_i = (int)ngroups;
  if (tid == 0)
    while (gN < kPhGroups - kPhLag && gN < ngroups_i) issue_next();

  const int k0 = tid / (2 * kPhArms);                             // block phase
  const int al = (tid % (2 * kPhArms)) >> 1, pol = tid & 1;                  // sample arm inside the group, polarisation
  const int jj = a0 + al + k0 * (a.N - a.S);
  const int c = jj >= a.N ? 1 : 0;
  const int j = jj - c * a.N;                                     // output arm
  const int s = k0 - c;                                           // row shift of this phase: 0 .. NU - 1
  float tap[P];
#pragma unroll
  for (int m = 0; m < P; ++m) tap[m] = __ldg(a.taps + j + m * a.N);
  const int slot = pmod(a.phase0 + a.start0 + (long long)k0 * a.S + j, a.N);
  const long long ystride = (long long)a.nu * a.N * 2;
  C* __restrict__ yo = reinterpret_cast<C*>(a.y) + ((long long)k0 * a.N + slot) * 2 + pol;
  const C* myring = ring + al * 2 + pol;

  C w[W];
  const int nv = (int)(i_end - v0);                               // virtual steps of this CTA (32-bit loop state)
  const int vskip = PER - 1;                                      // fill-only steps
  const int kmax = (int)min((long long)nv, (a.nblocks - k0 + a.nu - 1) / a.nu - v0);   // steps that produce a block
  yo += v0 * ystride;
  // the DE new rows of a step start r0 rows into group gA and gA advances by exactly one group per step, so the
  // ring position, the barrier addresses and their parities are carried incrementally as shared-memory byte
  // addresses (round 2, ncu r2y: with the positions recomputed per step only 20 % of the executed instructions were
  // FIR arithmetic)
  const int q0 = s + NEW0 % DE;
  const int gA0 = q0 / DE;                                        // relative group of the first new row
  const int r0 = q0 - gA0 * DE;
  int sA = gA0 % kPhGroups;
  uint32_t parA = (uint32_t)((gA0 / kPhGroups) & 1);
  constexpr uint32_t kGroupStride = DE * kRowC * sizeof(C);
  const uint32_t full0 = low::smem_u32(full), empty0 = low::smem_u32(empty);
  const uint32_t row_base = low::smem_u32(myring) + (uint32_t)r0 * kRowC * sizeof(C);
  uint32_t offA = (uint32_t)sA * kGroupStride;                     // byte offset of group sA in the ring
  bool release = gA0 >= 1;                                        // group gA - 1 exists (false at most on the first step)
  const bool lane0 = lane == 0;
  auto wait_addr = [](uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITA_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONEA_%=;\n"
        "bra WAITA_%=;\n"
        "DONEA_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
  };
  // One barrier per step, probed a whole FIR ahead: a step needs groups gA and (r0 != 0) gA + 1; every group before the
  // LAST one it needs was already waited for by the previous step, so the step only has to see group
  // gL = gA + (r0 != 0), and the (non-blocking) probe of the next step's gL is issued before this step's FIR, whose
  // FFMA2 chain hides the probe's latency (ncu r2y2: the two back-to-back try_waits at the top of a step collected twice
  // the stall samples of its 25 FFMA2s although they never had to spin)
  auto probe = [](uint32_t bar, uint32_t parity) -> bool {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
  };
  int sL = sA;                                                    // slot / parity of the last group the NEXT step needs
  uint32_t parL = parA;
  if (r0 != 0) {
    wait_addr(full0 + 8u * (uint32_t)sA, parA);                   // group gA0 of the first step
    if (++sL == kPhGroups) { sL = 0; parL ^= 1u; }
  }
  bool okL = false;
  // MAIN: every step of the rotation produces a block and releases a group (no predicates)
  auto step = [&](auto sidx_c, auto main_c, int vv) {
    constexpr int sidx = decltype(sidx_c)::value;
    constexpr bool MAIN = decltype(main_c)::value;
    const bool last = sA + 1 == kPhGroups;
    if (!okL) wait_addr(full0 + 8u * (uint32_t)sL, parL);
    if (++sL == kPhGroups) { sL = 0; parL ^= 1u; }
    okL = probe(full0 + 8u * (uint32_t)sL, parL);                 // next step's barrier, consumed after this step's FIR
    const uint32_t pr = row_base + offA;                          // rows pr[0 .. DE): slot 0 is mirrored behind slot G - 1
#pragma unroll
    for (int m = 0; m < DE; ++m) {
      float2 v;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(pr + (uint32_t)(m * kRowC * sizeof(C))));
      w[(DE * sidx + NEW0 + m) % W] = mk<float>(v.x, v.y);
    }
    __syncwarp();
    // group gA - 1 was last touched by this warp one step ago: release it (every group exactly once)
    if (lane0 && (MAIN || release)) {
      const uint32_t eb = empty0 + 8u * (uint32_t)(sA == 0 ? kPhGroups - 1 : sA - 1);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(eb) : "memory");
    }
    release = true;
    if (tid == 0 && gN < ngroups_i) issue_next();
    if (MAIN || (vv >= vskip && vv < kmax)) {
      // packed fp32x2 FMA with a scalar tap (FFMA2): 5 register reads per 2 MACs instead of 6
      float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int m = 0; m < P; ++m) {
        const C x = w[(DE * sidx + m) % W];
        if (m & 1) a1 = __ffma2_rn(make_float2(tap[m], tap[m]), make_float2(x.x, x.y), a1);
        else a0 = __ffma2_rn(make_float2(tap[m], tap[m]), make_float2(x.x, x.y), a0);
      }
      *yo = mk<float>(a0.x + a1.x, a0.y + a1.y);
    }
    yo += ystride;
    if (last) { sA = 0; offA = 0; parA ^= 1u; }
    else { ++sA; offA += kGroupStride; }
  };
  static_assert(PER == 4, "the unrolled step sequence below is written for a 4-step rotation");
  const std::integral_constant<bool, false> edge;
  const std::integral_constant<bool, true> mainc;
  int v = 0;
  if (nv >= PER) {                                                // first rotation: window fill, predicates
    step(std::integral_constant<int, 0>(), edge, 0);
    step(std::integral_constant<int, 1>(), edge, 1);
    step(std::integral_constant<int, 2>(), edge, 2);
    step(std::integral_constant<int, 3>(), edge, 3);
    v = PER;
#pragma unroll 1
    for (; v + PER <= kmax; v += PER) {                           // whole producing rotations: straight-line code
      step(std::integral_constant<int, 0>(), mainc, v);
      step(std::integral_constant<int, 1>(), mainc, v + 1);
      step(std::integral_constant<int, 2>(), mainc, v + 2);
      step(std::integral_constant<int, 3>(), mainc, v + 3);
    }
  }
#pragma unroll 1
  for (; v + PER <= nv; v += PER) {                               // tail rotations (blocks beyond the last one: consume only)
    step(std::integral_constant<int, 0>(), edge, v);
    step(std::integral_constant<int, 1>(), edge, v + 1);
    step(std::integral_constant<int, 2>(), edge, v + 2);
    step(std::integral_constant<int, 3>(), edge, v + 3);
  }
  if (v < nv) step(std::integral_constant<int, 0>(), edge, v);
  if (v + 1 < nv) step(std::integral_constant<int, 1>(), edge, v + 1);
  if (v + 2 < nv) step(std::integral_constant<int, 2>(), edge, v + 2);
}

template <int P, int DE> constexpr size_t fir_phases_smem() {
  return (size_t)(kPhGroups + 1) * DE * kPhArms * 2 * sizeof(C) + 2 * kPhGroups * sizeof(uint64_t);
}

// tensor map of the input stream seen as rows of N samples: [rows][N * 4 fp32], box = 16 arms x 2 pols x `de` rows
inline int make_in_map(const void* in, int N, long long rows, int de, CUtensorMap* tm) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e0 = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e0 != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return (int)cudaErrorNotSupported;
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)N * 4, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)N * 16};
  const cuuint32_t box[2] = {kPhArms * 4, (cuuint32_t)de};
  const cuuint32_t estr[2] = {1, 1};
  if (encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(in), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return (int)cudaErrorInvalidValue;
  return 0;
}

}  // namespace mid
}  // namespace pfb
