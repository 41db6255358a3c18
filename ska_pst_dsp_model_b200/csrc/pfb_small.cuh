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
//    two kernels, the serial FFT tail of every CTA costs more than the second launch; so was letting the last of a span's 8 arm
//    CTAs run the span's FFTs after an atomic count-in: 24.2 us.)
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

// ---------------------------------------------------------------------------
// Tensor-core form of the same FIR (the experiment SURVEY 7 hard part 2 / VERDICT r1 item 8 ask for): per (arm j,
// phase p) the outputs i = 8 I + a are Y[I][a] = sum_t U[I][t] F[t][a] with the Hankel matrix U[I][t] = u[56 I + t] read
// in place from the de-interleaved tile and the banded tap matrix F[t][a] = f_j[t - 7 a] (440 x 8, 384 non-zeros per
// column).  One warp = one 16 x 8 output tile = 128 outputs, 55 k-steps of mma.sync.m16n8k8 TF32; real and imaginary
// parts are separate products and every product is the error-compensated 3xTF32 split (hi*hi + lo*hi + hi*lo), i.e.
// 6 MMAs per k-step.  (tcgen05 needs its A operand in a canonical shared-memory layout; a Hankel matrix is not one, and
// materialising it costs 8x the shared-memory traffic of reading the samples.)
constexpr int kMmaSpan = 1024;                               // blocks per CTA: 8 phases x 128 outputs
constexpr int kMmaRow = 1288;                                // de-interleaved row pitch (samples of one residue mod 8)
constexpr int kMmaKSteps = 55;                               // ceil((7 * 7 + 384) / 8)

struct Fir8MmaArgs {
  const C* in;
  C* y;                     // [k][8] rotated arm sums (pfb_fft8_store finishes)
  const float4* btab;       // [8 arms][55 k-steps][32 lanes] (b0_hi, b1_hi, b0_lo, b1_lo), TF32-rounded on the host
  long long n_dat, nblocks, start0;
  int rot;
};

constexpr size_t fir8_mma_smem_bytes() { return (size_t)2 * 8 * kMmaRow * sizeof(float); }

__device__ __forceinline__ void mma_tf32(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 2) pfb_fir8_mma(const Fir8MmaArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tre = reinterpret_cast<float*>(smem_raw);           // [8 residues][kMmaRow]
  float* tim = tre + 8 * kMmaRow;
  const int j = blockIdx.y;
  const long long k0 = (long long)blockIdx.x * kMmaSpan;
  const long long s0 = a.start0 + 7 * k0;
#pragma unroll 8
  for (int t = threadIdx.x; t < 8 * kMmaRow; t += 256) {
    const long long s = s0 + t;
    float2 r = make_float2(0.f, 0.f);
    if (s < a.n_dat) r = __ldg(reinterpret_cast<const float2*>(a.in + s));
    tre[(t & 7) * kMmaRow + (t >> 3)] = r.x;
    tim[(t & 7) * kMmaRow + (t >> 3)] = r.y;
  }
  __syncthreads();
  const int p = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  const int q0 = 7 * p + j;                                  // first sample of the sub-stream: residue q0 & 7, offset q0 >> 3
  const float* ure = tre + (q0 & 7) * kMmaRow + (q0 >> 3) + 56 * gid + tig;
  const float* uim = tim + (q0 & 7) * kMmaRow + (q0 >> 3) + 56 * gid + tig;
  float dre[4] = {0.f, 0.f, 0.f, 0.f}, dim[4] = {0.f, 0.f, 0.f, 0.f};
  const float4* bt = a.btab + (size_t)j * kMmaKSteps * 32 + lane;
#pragma unroll 5
  for (int ks = 0; ks < kMmaKSteps; ++ks) {
    const float4 b = __ldg(bt + ks * 32);
    const uint32_t bh0 = __float_as_uint(b.x), bh1 = __float_as_uint(b.y), bl0 = __float_as_uint(b.z), bl1 = __float_as_uint(b.w);
    // A fragment: (row gid, col tig), (gid + 8, tig), (gid, tig + 4), (gid + 8, tig + 4); U[I][t] = u[56 I + t]
    const int o = 8 * ks;
    const float xr[4] = {ure[o], ure[o + 448], ure[o + 4], ure[o + 452]};
    const float xi[4] = {uim[o], uim[o + 448], uim[o + 4], uim[o + 452]};
    // split x = hi + lo with hi = x truncated to TF32 (the tensor core ignores the 13 low mantissa bits, so the raw
    // fp32 bits ARE hi) and lo = x - hi, exact in fp32: two instructions per element instead of cvt / sub / cvt
    uint32_t rh[4], rl[4], ih[4], il[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      rh[e] = __float_as_uint(xr[e]);
      rl[e] = __float_as_uint(xr[e] - __uint_as_float(rh[e] & 0xffffe000u));
      ih[e] = __float_as_uint(xi[e]);
      il[e] = __float_as_uint(xi[e] - __uint_as_float(ih[e] & 0xffffe000u));
    }
    mma_tf32(dre, rl, bh0, bh1);
    mma_tf32(dre, rh, bl0, bl1);
    mma_tf32(dre, rh, bh0, bh1);
    mma_tf32(dim, il, bh0, bh1);
    mma_tf32(dim, ih, bl0, bl1);
    mma_tf32(dim, ih, bh0, bh1);
  }
  // D: (row gid, col 2 tig), (gid, 2 tig + 1), (gid + 8, 2 tig), (gid + 8, 2 tig + 1); output i = 8 I + a
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int I = gid + 8 * (e >> 1), aa = 2 * tig + (e & 1);
    const long long k = k0 + p + 8 * (8 * I + aa);
    if (k < a.nblocks) a.y[k * 8 + ((a.rot + 7 * k + j) & 7)] = mk<float>(dre[e], dim[e]);
  }
}

inline int fir8_mma_launch(const Fir8MmaArgs& a, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(pfb_fir8_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fir8_mma_smem_bytes());
  if (e != cudaSuccess) return (int)e;
  const long long spans = (a.nblocks + kMmaSpan - 1) / kMmaSpan;
  if (spans > 2147483647ll) return (int)cudaErrorInvalidValue;
  pfb_fir8_mma<<<dim3((unsigned)spans, 8), 256, fir8_mma_smem_bytes(), st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace small8
}  // namespace pfb
