// Shape-specialised fast kernels for the SKA-Low PST shape (fp32):
//   K1  pfb_analysis_low      : 256 channels, OS 4/3 (step 192), 12 taps/arm (3072-tap filter)
//   K2  pfb_chan_idft_low_*   : 256-point inverse DFT across channels (TMA-fed; optionally writes the grouped,
//                               bank-swizzled scratch layout that K3 streams)
//   K3  pfb_resample_low_ws   : Nf=256 -> FN_width=192 overlap-save resampling per column, warp-autonomous and
//                               TMA-streamed (dual-pol, ov=48); pfb_resample_low is the block-synchronous form
//                               for single-pol / critically sampled / batched plans
// Selected by the plan when the shape matches; anything else runs pfb_generic.cuh.
//
// Design notes (DESIGN.md has the derivations):
//  * K1 exploits gcd(step, N) = 64: input samples with index == r (mod 64) only ever
//    meet taps with index == r (mod 64).  One thread owns one residue class (and one
//    polarisation): its 48 taps stay in registers for the whole launch, and a 48-sample
//    sliding window of its class stream lives in registers too, advancing 3 samples per
//    output block.  Every input sample is read from shared memory exactly once and used
//    in 16 complex MACs; the window rotation is resolved statically by unrolling 16 blocks.
//    The 4 arms a thread produces are exactly the inputs of one radix-4 DIF butterfly of
//    the 256-point FFT, and the oversampling phase rotation (a multiple of 64 slots) is a
//    static renaming of those 4 registers.  The remaining 64-point FFTs run one per thread
//    after one exchange through shared memory.
//  * input is streamed by the TMA bulk-copy engine (cp.async.bulk + mbarrier) into a
//    3-stage shared-memory ring; outputs go straight from registers to global memory in
//    fully used 32-byte sectors.
//  * K2/K3 split the 256-point FFT 16 x 16 (and the 192-point inverse 12 x 16) over 16
//    lanes per transform with in-register radix-16/12 butterflies and compile-time twiddles;
//    in K2 and the warp-autonomous K3 the 16 lanes share a warp, so the passes exchange data
//    with __syncwarp() only.
#pragma once
#include <cuda.h>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "pfb_generic.cuh"
#include "twiddle_tables.cuh"

namespace pfb {
namespace low {

typedef Cx<float> C;

// a * exp(DIR * 2 pi i k / N), k known after unrolling
template <int N, int DIR> __device__ __forceinline__ C mul_lit(C a, int k) {
  k %= N;
  if (k == 0) return a;
  if (2 * k == N) return mk<float>(-a.x, -a.y);
  if (4 * k == N) return mul_i<DIR>(a);
  if (4 * k == 3 * N) return mul_i<-DIR>(a);
  const float c = TwLit<N>::c(k), s = DIR * TwLit<N>::s(k);
  return mk<float>(a.x * c - a.y * s, a.x * s + a.y * c);
}

// Two-pass DIF FFT of N1*N2 points held in registers.  X[k] ends at slot
// N2*(k % N1) + k / N1  (see slot()).
template <int N1, int N2, int DIR> __device__ __forceinline__ void fft_reg(C* a) {
  constexpr int N = N1 * N2;
#pragma unroll
  for (int q = 0; q < N2; ++q) {
    C t[N1];
#pragma unroll
    for (int i = 0; i < N1; ++i) t[i] = a[q + N2 * i];
    Dft<N1, DIR, float>::run(t);
#pragma unroll
    for (int o = 0; o < N1; ++o) a[q + N2 * o] = mul_lit<N, DIR>(t[o], q * o);
  }
#pragma unroll
  for (int o = 0; o < N1; ++o) Dft<N2, DIR, float>::run(a + N2 * o);
}
template <int N1, int N2> __host__ __device__ constexpr int slot(int k) { return N2 * (k % N1) + k / N1; }

__device__ __forceinline__ C ld_nc(const C* p) {
  float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return mk<float>(v.x, v.y);
}
__device__ __forceinline__ void st_cs(C* p, C v) {   // streaming store (written once, read by a later kernel)
  __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y));
}

// ---- mbarrier / TMA bulk copy helpers --------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// (try_wait already suspends the warp in hardware for a bounded time; a longer suspend-time hint was measured
//  in round 2 and made all three kernels slower: K1 +5 %, K2 +17 %, K3 +2 %)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// exp(-2 pi i k / 64) for the looped 64-point FFT passes of K1 (indexed by a warp-uniform loop counter: LDC / uniform
// datapath, no register or shared-memory cost)
__constant__ float2 kW64[64] = {
    {1.0f, 0.0f}, {0.9951847266721969f, -0.0980171403295606f}, {0.9807852804032304f, -0.19509032201612825f}, {0.9569403357322088f, -0.29028467725446233f},
    {0.9238795325112867f, -0.3826834323650898f}, {0.881921264348355f, -0.47139673682599764f}, {0.8314696123025452f, -0.5555702330196022f}, {0.773010453362737f, -0.6343932841636455f},
    {0.7071067811865476f, -0.7071067811865475f}, {0.6343932841636455f, -0.773010453362737f}, {0.5555702330196023f, -0.8314696123025452f}, {0.4713967368259978f, -0.8819212643483549f},
    {0.38268343236508984f, -0.9238795325112867f}, {0.29028467725446233f, -0.9569403357322089f}, {0.19509032201612833f, -0.9807852804032304f}, {0.09801714032956077f, -0.9951847266721968f},
    {0.0f, -1.0f}, {-0.09801714032956065f, -0.9951847266721969f}, {-0.1950903220161282f, -0.9807852804032304f}, {-0.29028467725446216f, -0.9569403357322089f},
    {-0.3826834323650897f, -0.9238795325112867f}, {-0.4713967368259977f, -0.881921264348355f}, {-0.555570233019602f, -0.8314696123025455f}, {-0.6343932841636454f, -0.7730104533627371f},
    {-0.7071067811865475f, -0.7071067811865476f}, {-0.773010453362737f, -0.6343932841636455f}, {-0.8314696123025453f, -0.5555702330196022f}, {-0.8819212643483549f, -0.47139673682599786f},
    {-0.9238795325112867f, -0.3826834323650899f}, {-0.9569403357322088f, -0.2902846772544624f}, {-0.9807852804032304f, -0.1950903220161286f}, {-0.9951847266721968f, -0.09801714032956083f},
    {-1.0f, 0.0f}, {-0.9951847266721969f, 0.09801714032956059f}, {-0.9807852804032304f, 0.19509032201612836f}, {-0.9569403357322089f, 0.2902846772544621f},
    {-0.9238795325112868f, 0.38268343236508967f}, {-0.881921264348355f, 0.47139673682599764f}, {-0.8314696123025455f, 0.555570233019602f}, {-0.7730104533627371f, 0.6343932841636453f},
    {-0.7071067811865477f, 0.7071067811865475f}, {-0.6343932841636459f, 0.7730104533627367f}, {-0.5555702330196022f, 0.8314696123025452f}, {-0.47139673682599786f, 0.8819212643483549f},
    {-0.38268343236509034f, 0.9238795325112865f}, {-0.29028467725446244f, 0.9569403357322088f}, {-0.19509032201612866f, 0.9807852804032303f}, {-0.09801714032956045f, 0.9951847266721969f},
    {0.0f, 1.0f}, {0.09801714032956009f, 0.9951847266721969f}, {0.1950903220161283f, 0.9807852804032304f}, {0.29028467725446205f, 0.9569403357322089f},
    {0.38268343236509f, 0.9238795325112866f}, {0.4713967368259976f, 0.881921264348355f}, {0.5555702330196018f, 0.8314696123025455f}, {0.6343932841636456f, 0.7730104533627369f},
    {0.7071067811865474f, 0.7071067811865477f}, {0.7730104533627367f, 0.6343932841636459f}, {0.8314696123025452f, 0.5555702330196022f}, {0.8819212643483548f, 0.4713967368259979f},
    {0.9238795325112865f, 0.3826834323650904f}, {0.9569403357322088f, 0.2902846772544625f}, {0.9807852804032303f, 0.19509032201612872f}, {0.9951847266721969f, 0.0980171403295605f}};

// ===========================================================================
// K1: analysis
// ===========================================================================
struct AnaArgs {
  const C* in;           // [t][pol]
  C* out;                // [k][chan][pol]
  const float* taps;     // [48][64]: taps[u][r] = scale * f[r + 64 u]
  const float2* tw;      // [3][64]:  exp(-2 pi i r o / 256), o = 1..3
  const C* in2;          // continuation of the stream: samples n1.. live here (a neighbour GPU's buffer mapped over
                         // NVLink in time-sharded runs; null: everything is in `in`)
  long long n1;          // samples in `in` (== n_local when in2 is null)
  long long s0;          // window start of block 0 (>= 0, rotation-free: (phase0+s0) % 256 == 0)
  long long n_local;     // samples resident in `in` (+ `in2`)
  long long nblk;        // blocks to produce
  long long kout0;       // first output block slot
  int L;                 // blocks per CTA (multiple of 16)
  int chan_first, nchan_out;
  // fused round trip (EMIT > 0): the rotated arm sums ARE the channel-inverse-DFT of the channelized block
  // (sum_c out[c] W^(-c t1) = N * acc_rot[t1]), so the analysis writes them straight into the resample kernel's grouped
  // scratch and the channel IDFT kernel is not run at all
  char* v;               // grouped scratch (grouped_unit), rows = output block slots
  long long v_rows_cap;
};

#ifndef K1_FFT_UNROLL
#define K1_FFT_UNROLL 2
#endif
#ifndef K1_FFT_UNROLL2
#define K1_FFT_UNROLL2 K1_FFT_UNROLL
#endif
constexpr int kK1FftUnroll = K1_FFT_UNROLL, kK1FftUnroll2 = K1_FFT_UNROLL2;
constexpr int kAnaStages = 3;
constexpr int kTileBlocks = 4;                 // output blocks fed by one TMA tile
constexpr int kTileSamples = kTileBlocks * 192;
constexpr int kExStride = 65;                  // odd: conflict-free per-task rows

template <int NPOL> constexpr size_t ana_smem_bytes() {
  return (size_t)kAnaStages * kTileSamples * NPOL * sizeof(C) + (size_t)64 * NPOL * kExStride * sizeof(C) + 64;
}

__device__ __forceinline__ long long grouped_unit(long long rows_cap, long long row, int t1);

// FULL: all 256 channels are written in natural order (Bunton); otherwise the runtime channel
// window (chan_first, nchan_out) of the lowcbf variant applies.
// EMIT: 0 = channelized output only; 1 = channelized output + the inversion's intermediate v (dual-pol, fused
// analysis -> inversion in one process); 2 = v only (the channelized data is not wanted: the 64-point FFT stage and
// its stores are compiled out).
template <int NPOL, bool FULL, int EMIT = 0>
__global__ void __launch_bounds__(64 * NPOL, NPOL == 2 ? 2 : 4) pfb_analysis_low(const AnaArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* ring = reinterpret_cast<C*>(smem_raw);                                  // [stage][768][NPOL]
  C* ex = ring + kAnaStages * kTileSamples * NPOL;                           // [64*NPOL tasks][65]
  uint64_t* full = reinterpret_cast<uint64_t*>(ex + 64 * NPOL * kExStride);  // [stages]
  uint64_t* empty = full + kAnaStages;

  const int tid = threadIdx.x;
  const int r = tid / NPOL, pol = tid % NPOL;
  const int lane = tid & 31;
  const long long kA = (long long)blockIdx.x * a.L;
  const long long kend = min(kA + (long long)a.L, a.nblk);
  const int niter = (int)((kend - kA + 15) / 16);
  const int ntiles = niter * 4;
  constexpr uint32_t kTileBytes = kTileSamples * NPOL * sizeof(C);
  const long long tile0 = a.s0 + 192 * kA + 2880;       // first sample of tile 0
  const long long total_bytes = a.n_local * NPOL * (long long)sizeof(C);

  auto issue_tile = [&](int i) {
    const long long b0 = (tile0 + (long long)i * kTileSamples) * NPOL * (long long)sizeof(C);
    long long nbytes = total_bytes - b0;
    if (nbytes > (long long)kTileBytes) nbytes = kTileBytes;
    uint64_t* bar = &full[i % kAnaStages];
    if (nbytes > 0) {
      mbar_expect_tx(bar, (uint32_t)nbytes);
      // the tile may straddle the end of `in`: the rest comes straight from the neighbour's memory (peer pointer,
      // NVLink) - the halo of a time shard is never copied, the TMA engine reads it where it lives
      const long long own_bytes = a.n1 * NPOL * (long long)sizeof(C);
      long long na = own_bytes - b0;
      if (na > nbytes) na = nbytes;
      if (na < 0) na = 0;
      char* dst = reinterpret_cast<char*>(ring + (size_t)(i % kAnaStages) * kTileSamples * NPOL);
      if (na > 0) tma_load_1d(dst, reinterpret_cast<const char*>(a.in) + b0, (uint32_t)na, bar);
      if (nbytes > na)
        tma_load_1d(dst + na, reinterpret_cast<const char*>(a.in2) + (b0 + na - own_bytes), (uint32_t)(nbytes - na), bar);
    } else {
      mbar_arrive(bar);
    }
  };

  if (tid == 0) {
    for (int s = 0; s < kAnaStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2 * NPOL);   // one arrive per warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0)
    for (int i = 0; i < kAnaStages && i < ntiles; ++i) issue_tile(i);

  // taps and first-stage twiddles of this residue class: registers for the whole launch
  float tap[48];
#pragma unroll
  for (int u = 0; u < 48; ++u) tap[u] = __ldg(a.taps + u * 64 + r);
  C tw[3];
#pragma unroll
  for (int o = 0; o < 3; ++o) {
    float2 v = __ldg(a.tw + o * 64 + r);
    tw[o] = mk<float>(v.x, v.y);
  }
  // window prologue: elements 0..44 of block kA's window straight from global
  C w[48];
  {
    const long long i0 = a.s0 + 192 * kA + r;
#pragma unroll
    for (int u = 0; u < 45; ++u) {
      const long long idx = i0 + 64ll * u;
      w[u] = ld_nc(idx < a.n1 ? a.in + idx * NPOL + pol : a.in2 + (idx - a.n1) * NPOL + pol);
    }
  }

  // fused emit: byte offsets of (group of t1 = r + 64 i, row of block kA + 16 it) in the grouped scratch (< 4 GiB,
  // checked by the host)
  uint32_t voff[4] = {0u, 0u, 0u, 0u};
  const int vrow7 = (int)((a.kout0 + kA) & 7);
  const int nloc = (int)(kend - kA);
  if (EMIT > 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      voff[i] = (uint32_t)((((long long)((r >> 3) + 8 * i) * a.v_rows_cap + a.kout0 + kA) << 7) + pol * 8);
  }
  for (int it = 0; it < niter; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int tile = it * 4 + (j >> 2);
      if ((j & 3) == 0) {
        if (tid == 0 && tile >= 1 && tile - 1 + kAnaStages < ntiles) {
          // refill the stage freed by the previous tile
          mbar_wait(&empty[(tile - 1) % kAnaStages], ((tile - 1) / kAnaStages) & 1);
          issue_tile(tile - 1 + kAnaStages);
        }
        mbar_wait(&full[tile % kAnaStages], (tile / kAnaStages) & 1);
      }
      // the 3 new samples of this block: window elements 45..47
      {
        const C* sp = ring + ((size_t)(tile % kAnaStages) * kTileSamples + (j & 3) * 192 + r) * NPOL + pol;
#pragma unroll
        for (int e = 0; e < 3; ++e) w[(3 * j + 45 + e) % 48] = sp[e * 64 * NPOL];
      }
      if ((j & 3) == 3) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[tile % kAnaStages]);
      }
      // polyphase arms: acc[a] = sum_m tap[a+4m] * window[a+4m]
      C acc[4];
      // packed fp32x2 FMA: (sx, sy) += tap * (v.x, v.y) is ONE instruction with a scalar tap operand (SASS FFMA2
      // Rd, Ra.F32, Rb.F32x2, Rc.F32x2), i.e. 5 register reads per 2 MACs instead of 6 - the FIR is bound by
      // register-file bandwidth, not by the FMA pipe
#pragma unroll
      for (int ar = 0; ar < 4; ++ar) {
        float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < 12; ++m) {
          const int u = ar + 4 * m;
          const C v = w[(3 * j + u) % 48];
          s2 = __ffma2_rn(make_float2(tap[u], tap[u]), make_float2(v.x, v.y), s2);
        }
        acc[ar] = mk<float>(s2.x, s2.y);
      }
      // phase rotation: arm ar feeds FFT slot r + 64*((ar + 3j) mod 4)  ->  z[i] = acc[(i + j) mod 4]
      C z[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) z[i] = acc[(i + j) & 3];
      if (EMIT > 0) {
        // v[row][t1 = r + 64 i][pol] = N * z[i] at grouped_unit(row, t1): 16 consecutive r of a warp fill two whole
        // 128-byte group rows.  32-bit byte offsets from the (uniform) scratch base, carried per iteration; the row's
        // swizzle (row & 7) depends on j only because an iteration advances 16 rows
        if (16 * it + j < nloc) {
          const uint32_t swz = (uint32_t)(((r & 7) ^ ((vrow7 + j) & 7)) << 4) + (uint32_t)(j * 128);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<C*>(a.v + (size_t)(voff[i] + swz)) = mk<float>(256.f * z[i].x, 256.f * z[i].y);
        }
      }
      if (EMIT == 2) continue;
      Dft<4, -1, float>::run(z);
      z[1] = z[1] * tw[0];
      z[2] = z[2] * tw[1];
      z[3] = z[3] * tw[2];
#pragma unroll
      // exchange row T = ((j >> 1) * NPOL + pol) * 8 + (j & 1) * 4 + o: the two polarisations of a warp's 64-bit stores
      // sit 8 rows = 8 (mod 16) units apart, i.e. on the other half of the banks (the natural numbering (j, pol, o)
      // put them 4 rows apart: 2-way conflicts on half of the units, 34 % of K1's shared wavefronts in ncu r1d)
#pragma unroll
      for (int o = 0; o < 4; ++o)
        ex[(size_t)((((j >> 1) * NPOL + pol) * 8 + (j & 1) * 4) + o) * kExStride + r] = z[o];
    }
    if (EMIT > 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) voff[i] += 16u * 128u;
    }
    if (EMIT == 2) continue;
    __syncthreads();
    {
      // one 64-point FFT per thread (task = block j, pol, o), two radix-8 passes over the
      // thread's private row in shared memory (the 48-sample window keeps the registers)
      const int o = tid & 3, tp = (tid >> 3) % NPOL, tj = 2 * (tid / (8 * NPOL)) + ((tid >> 2) & 1);   // row T = tid
      C* row = ex + (size_t)tid * kExStride;
      // both passes are LOOPS (unrolled by two), not straight-line code: the fully unrolled form made the loop body of
      // this kernel 58 KB, beyond the 32 KB L1.5 instruction cache, and ncu charged 25 % of the stall samples to
      // instruction fetch (no_inst); the inter-pass twiddles W64^(q oo) come from the constant bank, indexed by the
      // warp-uniform loop counter
#pragma unroll kK1FftUnroll
      for (int q = 0; q < 8; ++q) {
        C t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = row[q + 8 * i];
        Dft<8, -1, float>::run(t);
        row[q] = t[0];
#pragma unroll
        for (int oo = 1; oo < 8; ++oo) {
          const float2 w = kW64[q * oo];
          row[q + 8 * oo] = mk<float>(t[oo].x * w.x - t[oo].y * w.y, t[oo].x * w.y + t[oo].y * w.x);
        }
      }
      const long long kb = kA + 16ll * it + tj;
      C* op = a.out + ((a.kout0 + kb) * (FULL ? 256 : a.nchan_out)) * NPOL + tp + (FULL ? o * NPOL : 0);
      const bool live = kb < kend;
#pragma unroll kK1FftUnroll2
      for (int o1 = 0; o1 < 8; ++o1) {
        C t[8];
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) t[qq] = row[8 * o1 + qq];
        Dft<8, -1, float>::run(t);
#pragma unroll
        for (int o2 = 0; o2 < 8; ++o2) {
          if (FULL) {
            if (live) op[4 * (o1 + 8 * o2) * NPOL] = t[o2];   // X256[o + 4 k'], k' = o1 + 8 o2
          } else {
            const int c = o + 4 * (o1 + 8 * o2);
            const int ci = (c - a.chan_first) & 255;
            if (live && ci < a.nchan_out) op[(long long)ci * NPOL] = t[o2];
          }
        }
      }
    }
    __syncthreads();
  }
}

// ===========================================================================
// K2: 256-point inverse DFT across channels, 16 lanes x 16 points
// ===========================================================================
struct ChanArgs {
  const C* in;          // [row][256][pol]
  C* out;               // [row][256][pol]  (or the grouped layout below)
  const float2* tw256;  // [256] exp(-2 pi i k/256)
  long long nrows;
  long long rows_cap;   // grouped layout: rows per column group
};

// Grouped scratch layout read by pfb_resample_low_ws: the 256 columns t1 are split into 32 groups of
// 8; group g holds its rows contiguously, one row = 8 x 16 B (both polarisations of one t1), and
// the 16-byte unit of t1 sits at position (t1 & 7) ^ (row & 7), so that a plain bulk copy of
// whole rows lands in shared memory already bank-swizzled for the column reads.
__device__ __forceinline__ long long grouped_unit(long long rows_cap, long long row, int t1) {
  return (((long long)(t1 >> 3) * rows_cap + row) << 3) + ((t1 & 7) ^ (int)(row & 7));
}

constexpr int kColStride = 273;   // 16*17 + 1, odd
constexpr int kQStride = 17;

template <int NPOL>
__global__ void __launch_bounds__(256, 3) pfb_chan_idft_low(const ChanArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* ex = reinterpret_cast<C*>(smem_raw);    // [2][16 ffts][273]
  const int tid = threadIdx.x;
  // lanes: (q, pol) contiguous in memory; 16 FFTs per CTA pass = 16/NPOL rows
  const int lane32 = tid & 31, grp = tid >> 5;       // 8 groups of 32 lanes
  int q, pol, rsub;
  if (NPOL == 2) { q = lane32 >> 1; pol = lane32 & 1; rsub = grp; }
  else { q = lane32 & 15; pol = 0; rsub = grp * 2 + (lane32 >> 4); }
  constexpr int ROWS = 16 / NPOL;
  const int fft = rsub * NPOL + pol;
  C tw[15];
#pragma unroll
  for (int o = 1; o < 16; ++o) {
    float2 v = __ldg(a.tw256 + ((q * o) & 255));
    tw[o - 1] = mk<float>(v.x, -v.y);          // inverse: conj
  }
  const long long npass = (a.nrows + ROWS - 1) / ROWS;
  int buf = 0;
  for (long long ps = blockIdx.x; ps < npass; ps += gridDim.x, buf ^= 1) {
    const long long row = ps * ROWS + rsub;
    C* e = ex + (size_t)buf * 16 * kColStride + (size_t)fft * kColStride;
    C x[16];
    if (row < a.nrows) {
      const C* p = a.in + (row * 256 + q) * NPOL + pol;
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = ld_nc(p + i * 16 * NPOL);
      fft_reg<4, 4, +1>(x);
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        C v = x[slot<4, 4>(o)];
        if (o) v = v * tw[o - 1];
        e[o * kQStride + q] = v;
      }
    }
    __syncthreads();
    if (row < a.nrows) {
      const int o = q;   // second pass: this lane owns output residue o
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = e[o * kQStride + i];
      fft_reg<4, 4, +1>(x);
      C* p = a.out + (row * 256 + o) * NPOL + pol;
#pragma unroll
      for (int o2 = 0; o2 < 16; ++o2) p[o2 * 16 * NPOL] = x[slot<4, 4>(o2)];
    }
  }
}

// TMA-fed dual-pol kernel (one lane carries both polarisations of its 16 channels, every access 128 bits
// wide): whole rows (4 KiB, contiguous in TFP order) are
// streamed by cp.async.bulk into a 3-stage ring, so no warp ever waits on a global load, and
// because the 16 lanes of a transform live in one warp the two passes exchange data through the
// ring slot itself with __syncwarp() only - the CTA never executes a block-wide barrier.
constexpr int kChanStages = 3;
constexpr int kChanRows = 8;                      // rows per stage = 16 lanes x 8 = 128 threads
template <bool GROUPED>
__global__ void __launch_bounds__(128, 2) pfb_chan_idft_low_tma(const ChanArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* ring = reinterpret_cast<float4*>(smem_raw);                    // [stage][8 rows][273]
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + kChanStages * kChanRows * kColStride);
  uint64_t* empty = full + kChanStages;
  const int tid = threadIdx.x, lane = tid & 31;
  const int q = tid & 15, rsub = tid >> 4;
  const long long npass = (a.nrows + kChanRows - 1) / kChanRows;
  // passes of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const long long my = npass > blockIdx.x ? (npass - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  auto issue = [&](long long i) {                 // i-th pass of this CTA
    const long long row0 = (blockIdx.x + i * gridDim.x) * kChanRows;
    const int st = (int)(i % kChanStages);
    long long nr = a.nrows - row0;
    if (nr > kChanRows) nr = kChanRows;
    mbar_expect_tx(&full[st], (uint32_t)(nr * 4096));
    for (int r = 0; r < nr; ++r)
      tma_load_1d(ring + ((size_t)st * kChanRows + r) * kColStride,
                  reinterpret_cast<const char*>(a.in) + (row0 + r) * 4096, 4096, &full[st]);
  };
  if (tid == 0) {
    for (int s = 0; s < kChanStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0)
    for (long long i = 0; i < kChanStages && i < my; ++i) issue(i);

  C tw[15];
#pragma unroll
  for (int o = 1; o < 16; ++o) {
    float2 v = __ldg(a.tw256 + ((q * o) & 255));
    tw[o - 1] = mk<float>(v.x, -v.y);
  }
  for (long long i = 0; i < my; ++i) {
    const int st = (int)(i % kChanStages);
    if (tid == 0 && i >= 1 && i - 1 + kChanStages < my) {
      mbar_wait(&empty[(i - 1) % kChanStages], (uint32_t)(((i - 1) / kChanStages) & 1));
      issue(i - 1 + kChanStages);
    }
    mbar_wait(&full[st], (uint32_t)((i / kChanStages) & 1));
    const long long row = (blockIdx.x + i * gridDim.x) * kChanRows + rsub;
    float4* e = ring + ((size_t)st * kChanRows + rsub) * kColStride;
    C x0[16], x1[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 v = e[q + 16 * k];
      x0[k] = mk<float>(v.x, v.y);
      x1[k] = mk<float>(v.z, v.w);
    }
    __syncwarp();                                  // both rows of this warp fully read
    fft_reg<4, 4, +1>(x0);
    fft_reg<4, 4, +1>(x1);
#pragma unroll
    for (int o = 0; o < 16; ++o) {
      C u0 = x0[slot<4, 4>(o)], u1 = x1[slot<4, 4>(o)];
      if (o) { u0 = u0 * tw[o - 1]; u1 = u1 * tw[o - 1]; }
      e[o * kQStride + q] = make_float4(u0.x, u0.y, u1.x, u1.y);
    }
    __syncwarp();
    {
      const int o = q;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float4 v = e[o * kQStride + k];
        x0[k] = mk<float>(v.x, v.y);
        x1[k] = mk<float>(v.z, v.w);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // our generic-proxy writes vs the next TMA fill
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);      // slot free for the TMA again
      fft_reg<4, 4, +1>(x0);
      fft_reg<4, 4, +1>(x1);
      if (row < a.nrows) {
        float4* p = reinterpret_cast<float4*>(a.out) + (GROUPED ? grouped_unit(a.rows_cap, row, o) : row * 256 + o);
        const long long step = GROUPED ? 16 * a.rows_cap : 16;      // t1 += 16: two groups further
#pragma unroll
        for (int o2 = 0; o2 < 16; ++o2) {
          const C u0 = x0[slot<4, 4>(o2)], u1 = x1[slot<4, 4>(o2)];
          p[o2 * step] = make_float4(u0.x, u0.y, u1.x, u1.y);
        }
      }
    }
  }
}

// ===========================================================================
// K3: per column: taper -> FFT256 -> keep 192 bins x weights -> IFFT192 -> keep middle
// ===========================================================================
struct ResArgs {
  const C* v;           // [tau][nchan][pol]
  C* out;               // [T2][nchan][pol]
  const float* taper;   // [256]
  const C* wt;          // [nchan][192] in FN order (r)
  const float2* tw256;  // [256] exp(-2 pi i k/256)
  const float2* tw192;  // [192] exp(+2 pi i k/192)
  int nchan, keep, ov2, keep2, flip;   // flip: critical sampling => y *= (-1)^t2
  int nblk;             // overlap-save blocks in this launch
  int nsplit;           // CTAs per column group along n
  int wt_mod;           // channels per group (weight row = t1 % wt_mod)
};

template <int COLS> constexpr size_t res_smem_bytes() {
  return (size_t)2 * COLS * kColStride * sizeof(C) + (size_t)COLS * 193 * sizeof(C) + (256 + 192) * sizeof(C) +
         256 * sizeof(float);
}

// NCHAN > 0 / OV2 >= 0 bake the row pitch and the kept output range into the instruction stream
// (immediate address offsets, pruned last butterfly stage); 0 / -1 read them from the arguments.
// COLS = columns (t1, pol pairs) per CTA; the CTA has 16 * COLS threads.
template <int NPOL, int NCHAN, int OV2, bool FLIP, int COLS>
__global__ void __launch_bounds__(16 * COLS, 512 / (16 * COLS)) pfb_resample_low(const ResArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* exA = reinterpret_cast<C*>(smem_raw);            // [COLS][273]  phase 1 -> 2
  C* exB = exA + COLS * kColStride;                   // [COLS][273]  phase 2 -> 3 ([m][o] stride 17)
  C* swt = exB + COLS * kColStride;                   // [COLS][193]  weights of this column group
  C* t1s = swt + COLS * 193;                          // [16 o][16 q]  exp(-2 pi i q o / 256)
  C* t2s = t1s + 256;                                 // [12 m][16 o]  exp(+2 pi i o m / 192)
  float* stp = reinterpret_cast<float*>(t2s + 192);
  const int tid = threadIdx.x;
  const int c = tid % COLS, q = tid / COLS;           // column, lane-in-transform
  constexpr int T1G = COLS / NPOL;                    // t1 values per CTA
  const int nchan = NCHAN > 0 ? NCHAN : a.nchan;
  const int ov2 = OV2 >= 0 ? OV2 : a.ov2;
  const int keep = OV2 >= 0 ? 256 - 2 * (OV2 * 4 / 3) : a.keep;
  const int keep2 = OV2 >= 0 ? 192 - 2 * OV2 : a.keep2;
  const int groups = nchan / T1G;
  const int g = blockIdx.x % groups, sp = blockIdx.x / groups;
  const int t1_0 = g * T1G;
  const int rowlen = nchan * NPOL;
  const int per = (a.nblk + a.nsplit - 1) / a.nsplit;
  const int n_begin = sp * per, n_end = min(a.nblk, n_begin + per);

  for (int i = tid; i < COLS * 192; i += 16 * COLS) {
    const int cc = i / 192, rr = i % 192;
    swt[cc * 193 + rr] = a.wt[(size_t)((t1_0 + cc / NPOL) % a.wt_mod) * 192 + rr];
  }
  for (int i = tid; i < 256; i += 16 * COLS) stp[i] = a.taper[i];
  // inter-pass twiddles live in shared tables (all lanes of a warp that share q read one address:
  // broadcast), which frees the registers that carry the next block's samples across phases 2 and 3
  for (int i = tid; i < 256; i += 16 * COLS) {
    const float2 v = __ldg(a.tw256 + (((i >> 4) * (i & 15)) & 255));
    t1s[i] = mk<float>(v.x, v.y);
  }
  for (int i = tid; i < 192; i += 16 * COLS) {
    const float2 v = __ldg(a.tw192 + (((i >> 4) * (i & 15)) % 192));
    t2s[i] = mk<float>(v.x, v.y);
  }
  __syncthreads();

  const C* vcol = a.v + (long long)t1_0 * NPOL + c;
  C* ocol = a.out + (long long)t1_0 * NPOL + c;
  C* eA = exA + c * kColStride;
  C* eB = exB + c * kColStride;

  C xn[16];                                           // samples of the next block, in flight during phases 2-3
  if (n_begin < n_end) {
    const C* p = vcol + ((long long)n_begin * keep + q) * rowlen;
#pragma unroll
    for (int i = 0; i < 16; ++i) xn[i] = ld_nc(p + i * 16 * rowlen);
  }
  for (int n = n_begin; n < n_end; ++n) {
    C x[16];
    {
      // phase 1: lane q holds tau = q + 16 i
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = xn[i] * stp[q + 16 * i];
      if (n + 1 < n_end) {
        const C* p = vcol + ((long long)(n + 1) * keep + q) * rowlen;
#pragma unroll
        for (int i = 0; i < 16; ++i) xn[i] = ld_nc(p + i * 16 * rowlen);
      }
      fft_reg<4, 4, -1>(x);
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        C v = x[slot<4, 4>(o)];
        if (o) v = v * t1s[o * 16 + q];
        eA[o * kQStride + q] = v;
      }
    }
    __syncthreads();
    {
      // phase 2: lane o=q: bins b = o + 16 o2; keep o2 in {0..5} U {10..15}
      const int o = q;
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = eA[o * kQStride + i];
      fft_reg<4, 4, -1>(x);
      C z[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const int o2 = k < 6 ? k : k + 4;
        const int b = o + 16 * o2;                        // FFT bin
        const int rr = b < 96 ? b + 96 : b - 160;         // FN-order index of that bin
        z[k] = x[slot<4, 4>(o2)] * swt[c * 193 + rr];
      }
      // z[k] sits at inverse-transform index i' = o + 16 k: 12-point inverse DFT over k
      fft_reg<3, 4, +1>(z);
#pragma unroll
      for (int m = 0; m < 12; ++m) {
        C v = z[slot<3, 4>(m)];
        if (m) v = v * t2s[m * 16 + o];                   // exp(+2 pi i o m / 192)
        eB[m * kQStride + o] = v;
      }
    }
    __syncthreads();
    if (q < 12) {
      // phase 3: lane m=q: y[m + 12 j] = sum_o eB[m][o] exp(+2 pi i o j / 16)
      const int m = q;
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = eB[m * kQStride + i];
      fft_reg<4, 4, +1>(x);
      C* p = ocol + ((long long)n * keep2 - ov2 + m) * rowlen;
      const bool odd = FLIP && (m & 1);          // t2 = m + 12 j has the parity of m
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        // OV2 multiple of 12: the kept range is whole j's, decided at compile time
        const bool keep_j = (OV2 >= 0 && OV2 % 12 == 0) ? (12 * j >= OV2 && 12 * j < 192 - OV2)
                                                        : (m + 12 * j >= ov2 && m + 12 * j < 192 - ov2);
        if (keep_j) {
          C v = x[slot<4, 4>(j)];
          if (odd) v = mk<float>(-v.x, -v.y);
          p[12 * j * rowlen] = v;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// K3 (SKA-Low shape, dual-pol): warp-autonomous variant.
//
// One warp owns one t1 (both polarisations): lane = (q, pol), so the 16 lanes of a transform share
// a warp and the two exchanges of the 16x16 / 12x16 passes need __syncwarp() only - no block-wide
// barrier couples the 8 compute warps of a CTA.  Global memory is touched by the TMA engine alone:
// a 9th (producer) warp streams the column group's rows from the grouped scratch layout
// (grouped_unit above; one bulk copy per 160-row overlap-save step, double-buffered) and writes the
// finished output rows back with bulk stores from a staging tile.  The 96-row overlap between
// consecutive blocks never leaves the register file: 160 = 16 * 10, so elements i = 10..15 of
// lane q in block n are elements i = 0..5 of the same lane in block n + 1.
// ---------------------------------------------------------------------------
constexpr int kWsWarps = 8;                           // compute warps = t1 per CTA
constexpr int kWsRowBytes = kWsWarps * 16;            // 128 B: 8 t1 x 2 pol
constexpr int kWsStepRows = 160;                      // keep = Nf - 2 ov
constexpr int kWsSlots = 2;
constexpr int kWsColStride = 280;                     // per-pol exchange buffer pitch: 560 words = 16 (mod 32)
constexpr int kWsOutRows = 120;                       // keep2
// Compile-time tuning knobs (tools/build_variants2.sh builds variants side by side; values below are the measured best):
// loop-invariant per-lane table values kept in registers instead of re-read from shared memory every block
#ifndef K3_TAPER_REGS
#define K3_TAPER_REGS 1      // the six edge taper values of a lane
#endif
#ifndef K3_TW1_REGS
#define K3_TW1_REGS 6        // first-exchange twiddles o = 1..6 (8, 10: no further gain)
#endif
#ifndef K3_TW2_REGS
#define K3_TW2_REGS 0        // second-exchange twiddles (4, 6: slightly slower)
#endif

struct WsArgs {
  const char* v;        // grouped scratch
  char* out;            // [T2][256][2] complex64
  const float* taper;   // [256]
  const C* wt;          // [nchan][192]
  const float2* tw256;
  const float2* tw192;
  long long rows_cap;   // rows per group in the scratch
  long long split_rows; // scratch rows between the starts of consecutive time splits (stream: per * 160; batch of
                        // independent vectors: the row pitch of one vector)
  int nchan;            // output row pitch in t1
  int nblk, nsplit;
};

constexpr size_t ws_smem_bytes() {
  return (size_t)kWsSlots * kWsStepRows * kWsRowBytes         // ring
         + (size_t)kWsWarps * 2 * kWsColStride * sizeof(C)   // exchange
         + (size_t)kWsOutRows * kWsRowBytes                  // output staging (tensor-map box, 128B swizzle)
         + (size_t)kWsWarps * 192 * sizeof(C)                // weights
         + (256 + 192) * sizeof(C) + 256 * sizeof(float)     // twiddles, taper
         + 64;                                               // mbarriers
}

// one box [rows][128 B] of the staging tile -> global rows (x in fp32 elements, y in rows)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* ssrc, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(x), "r"(y),
               "r"(smem_u32(ssrc))
               : "memory");
}

// UNITY_MID: the temporal taper is exactly 1 on [ov, Nf - ov) (Tukey, top-hat interior, no window), so only
// the 6 edge samples of a lane are scaled.
template <bool UNITY_MID>
__global__ void __launch_bounds__(32 * (kWsWarps + 1), 2)
    pfb_resample_low_ws(const WsArgs a, const __grid_constant__ CUtensorMap out_map) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* ring = smem_raw;                                              // [slot][160][128 B]
  C* exch = reinterpret_cast<C*>(ring + kWsSlots * kWsStepRows * kWsRowBytes); // [warp][pol][280]
  unsigned char* outs = reinterpret_cast<unsigned char*>(exch + kWsWarps * 2 * kWsColStride);   // [120][128 B], 1 KiB aligned
  C* swt = reinterpret_cast<C*>(outs + kWsOutRows * kWsRowBytes);              // [8 t1][192]
  C* t1s = swt + kWsWarps * 192;                                               // [16 o][16 q]
  C* t2s = t1s + 256;                                                          // [12 m][16 o]
  float* stp = reinterpret_cast<float*>(t2s + 192);                            // [256]
  uint64_t* full = reinterpret_cast<uint64_t*>(stp + 256);                     // [2]
  uint64_t* empty = full + kWsSlots;                                           // [2]
  uint64_t* out_full = empty + kWsSlots;
  uint64_t* out_free = out_full + 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int groups = a.nchan >> 3;
  const int g = blockIdx.x % groups, sp = blockIdx.x / groups;
  const int per = (a.nblk + a.nsplit - 1) / a.nsplit;
  const int n_begin = sp * per, n_end = min(a.nblk, n_begin + per);
  const int nb = n_end - n_begin;
  if (nb <= 0) return;

  if (tid == 0) {
    for (int s = 0; s < kWsSlots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWsWarps); }
    mbar_init(out_full, kWsWarps);
    mbar_init(out_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const char* vgrp = a.v + ((long long)g * a.rows_cap + (long long)sp * a.split_rows) * kWsRowBytes;
  // load k: k = 0 is the 96-row head of the first block, k >= 1 the 160 new rows of block k - 1
  auto issue = [&](int k) {
    const uint32_t bytes = (k == 0 ? 96u : 160u) * kWsRowBytes;
    const long long row0 = k == 0 ? 0 : (long long)(k - 1) * kWsStepRows + 96;
    uint64_t* bar = &full[k & 1];
    mbar_expect_tx(bar, bytes);
    tma_load_1d(ring + (size_t)(k & 1) * kWsStepRows * kWsRowBytes, vgrp + row0 * kWsRowBytes, bytes, bar);
  };
  if (warp == kWsWarps && lane == 0) { issue(0); issue(1); }

  // tables (all warps help)
  for (int i = tid; i < kWsWarps * 192; i += blockDim.x) swt[i] = a.wt[(size_t)g * kWsWarps * 192 + i];
  for (int i = tid; i < 256; i += blockDim.x) {
    stp[i] = a.taper[i];
    const float2 v = __ldg(a.tw256 + (((i >> 4) * (i & 15)) & 255));
    t1s[i] = mk<float>(v.x, v.y);
  }
  for (int i = tid; i < 192; i += blockDim.x) {
    const float2 v = __ldg(a.tw192 + (((i >> 4) * (i & 15)) % 192));
    t2s[i] = mk<float>(v.x, v.y);
  }
  __syncthreads();

  if (warp == kWsWarps) {
    // ---------------- producer warp ----------------
    if (lane != 0) return;
    for (int b = 0; b < nb; ++b) {
      if (b + 2 <= nb) {
        mbar_wait(&empty[b & 1], (uint32_t)((b >> 1) & 1));      // load b consumed by every compute warp
        issue(b + 2);
      }
      mbar_wait(out_full, (uint32_t)(b & 1));
      tma_store_2d(&out_map, outs, g * (kWsRowBytes / 4), (n_begin + b) * kWsOutRows);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      mbar_arrive(out_free);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    return;
  }

  // ---------------- compute warps ----------------
  const int q = lane >> 1, pol = lane & 1;
  C* ex = exch + (size_t)(warp * 2 + pol) * kWsColStride;
  const C* wrow = swt + warp * 192;
  // byte offset of this lane's column inside a ring row: unit (t1l ^ (row & 7)), row & 7 == q & 7
  const int coloff = ((warp ^ (q & 7)) << 4) + (pol << 3);

#if K3_TAPER_REGS
  float tpr[6];                                       // the lane's six edge taper values: loop-invariant, kept in registers
#pragma unroll
  for (int i = 0; i < 6; ++i) tpr[i] = stp[q + 16 * (i < 3 ? i : i + 10)];
#endif
  C tw1r[K3_TW1_REGS > 0 ? K3_TW1_REGS : 1];          // first-exchange twiddles of o = 1 .. K3_TW1_REGS: loop-invariant per lane
#pragma unroll
  for (int o = 1; o <= K3_TW1_REGS; ++o) tw1r[o - 1] = t1s[o * 16 + q];
  C tw2r[K3_TW2_REGS > 0 ? K3_TW2_REGS : 1];          // second-exchange twiddles of m = 1 .. K3_TW2_REGS (lane o = q)
#pragma unroll
  for (int m = 1; m <= K3_TW2_REGS; ++m) tw2r[m - 1] = t2s[m * 16 + q];
  C carry[6];
  {
    mbar_wait(&full[0], 0);
    const unsigned char* rp = ring + (size_t)q * kWsRowBytes + coloff;
#pragma unroll
    for (int i = 0; i < 6; ++i) carry[i] = *reinterpret_cast<const C*>(rp + (size_t)i * 16 * kWsRowBytes);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[0]);
  }

  for (int b = 0; b < nb; ++b) {
    C x[16];
    {
      // phase 1: lane q holds tau = q + 16 i
      const int k = b + 1;
      mbar_wait(&full[k & 1], (uint32_t)((k >> 1) & 1));
      const unsigned char* rp = ring + ((size_t)(k & 1) * kWsStepRows + q) * kWsRowBytes + coloff;
#pragma unroll
      for (int i = 0; i < 6; ++i) x[i] = carry[i];
#pragma unroll
      for (int i = 6; i < 16; ++i) x[i] = *reinterpret_cast<const C*>(rp + (size_t)(i - 6) * 16 * kWsRowBytes);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[k & 1]);
#pragma unroll
      for (int i = 0; i < 6; ++i) carry[i] = x[i + 10];
#if K3_TAPER_REGS
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (!UNITY_MID) x[i] = x[i] * stp[q + 16 * i];
        else if (i < 3) x[i] = x[i] * tpr[i];
        else if (i >= 13) x[i] = x[i] * tpr[i - 10];
#else
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (!UNITY_MID || i < 3 || i >= 13) x[i] = x[i] * stp[q + 16 * i];      // ov = 48 = 3 * 16
#endif
      fft_reg<4, 4, -1>(x);
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        C v = x[slot<4, 4>(o)];
        if (o) v = v * (o <= K3_TW1_REGS ? tw1r[o - 1 < K3_TW1_REGS ? o - 1 : 0] : t1s[o * 16 + q]);
        ex[o * kQStride + q] = v;
      }
    }
    __syncwarp();
    {
      // phase 2: lane o = q: bins b = o + 16 o2; keep o2 in {0..5} U {10..15}
      const int o = q;
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = ex[o * kQStride + i];
      __syncwarp();
      fft_reg<4, 4, -1>(x);
      C z[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const int o2 = k < 6 ? k : k + 4;
        const int bb = o + 16 * o2;
        const int rr = bb < 96 ? bb + 96 : bb - 160;
        z[k] = x[slot<4, 4>(o2)] * wrow[rr];
      }
      fft_reg<3, 4, +1>(z);
#pragma unroll
      for (int m = 0; m < 12; ++m) {
        C v = z[slot<3, 4>(m)];
        if (m) v = v * (m <= K3_TW2_REGS ? tw2r[m - 1 < K3_TW2_REGS ? m - 1 : 0] : t2s[m * 16 + o]);
        ex[m * kQStride + o] = v;
      }
    }
    __syncwarp();
    if (q < 12) {
      // phase 3: lane m = q: y[m + 12 j] = sum_o ex[m][o] exp(+2 pi i o j / 16); keep j in [3, 13)
      const int m = q;
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = ex[m * kQStride + i];
      fft_reg<4, 4, +1>(x);
      if (b > 0) mbar_wait(out_free, (uint32_t)((b - 1) & 1));
      // staging row r = m + 12 j - 36 (12 j - 36 = 0 (mod 8) for odd j, 4 for even j); the tensor map's 128B
      // swizzle puts 16-byte unit u of a row at u ^ (address bits 7..9)
      const int rb = m + (int)((smem_u32(outs) >> 7) & 7);
      unsigned char* op0 = outs + (size_t)m * kWsRowBytes + ((warp ^ (rb & 7)) << 4) + (pol << 3);
      unsigned char* op1 = outs + (size_t)m * kWsRowBytes + ((warp ^ ((rb + 4) & 7)) << 4) + (pol << 3);
#pragma unroll
      for (int j = 3; j < 13; ++j)
        *reinterpret_cast<C*>(((j & 1) ? op0 : op1) + (12 * j - 36) * kWsRowBytes) = x[slot<4, 4>(j)];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(out_full);
  }
}

// ===========================================================================
// host side
// ===========================================================================
struct LowInverse {
  const float* taper;
  const C* wt;
  const float2* tw256;
  const float2* tw192;
  int nchan, keep, ov2, keep2, flip, chan_fast;
  bool taper_unity_mid;   // taper == 1 on [ov, Nf - ov)
};

template <typename U> inline int upload_raw(const std::vector<U>& h, std::vector<void*>* allocs, const U** out) {
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, h.size() * sizeof(U) + 16);
  if (e != cudaSuccess) return (int)e;
  allocs->push_back(d);
  e = cudaMemcpy(d, h.data(), h.size() * sizeof(U), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return (int)e;
  *out = reinterpret_cast<const U*>(d);
  return 0;
}

inline void unit_circle(long long m, long long n, int sign, double* c, double* s) {
  m %= n;
  if (m < 0) m += n;
  const double ang = 2.0 * 3.14159265358979323846 * ((double)m / (double)n);
  *c = cos(ang);
  *s = sign * sin(ang);
  if (m == 0) { *c = 1; *s = 0; }
  else if (2 * m == n) { *c = -1; *s = 0; }
  else if (4 * m == n) { *c = 0; *s = sign; }
  else if (4 * m == 3 * n) { *c = 0; *s = -sign; }
}

inline std::vector<float2> unit_table(int n, int sign) {
  std::vector<float2> t(n);
  for (int k = 0; k < n; ++k) {
    double c, s;
    unit_circle(k, n, sign, &c, &s);
    t[k] = make_float2((float)c, (float)s);
  }
  return t;
}

struct LowAnalysis {
  const float* taps;
  const float2* tw;
};

}  // namespace low

// ---- analysis --------------------------------------------------------------
inline int low_analysis_plan_init(int N, int nu, int de, int P, int dtype, const double* taps, int ntaps, double scale,
                                  std::vector<void*>* allocs, const void** d_low, bool* enabled) {
  *enabled = false;
  if (!(N == 256 && nu == 4 && de == 3 && P == 12 && dtype == 0)) return 0;
  std::vector<float> t(48 * 64, 0.f);
  for (int u = 0; u < 48; ++u)
    for (int r = 0; r < 64; ++r) {
      const int n = r + 64 * u;
      t[u * 64 + r] = n < ntaps ? (float)(scale * taps[n]) : 0.f;
    }
  std::vector<float2> tw(3 * 64);
  for (int o = 1; o < 4; ++o)
    for (int r = 0; r < 64; ++r) {
      double c, s;
      low::unit_circle((long long)r * o, 256, -1, &c, &s);
      tw[(o - 1) * 64 + r] = make_float2((float)c, (float)s);
    }
  low::LowAnalysis* h = new low::LowAnalysis();
  int rc = low::upload_raw(t, allocs, &h->taps);
  if (!rc) rc = low::upload_raw(tw, allocs, &h->tw);
  if (rc) { delete h; return rc; }
  *d_low = h;   // small host struct, leaked with the plan (two pointers)
  *enabled = true;
  return 0;
}

// returns a cudaError (0 ok); *nfast = number of leading-aligned blocks handled by the fast kernel,
// the caller runs the generic kernel for blocks [0,*k_lo) and [*k_hi, nblocks)
template <typename T>
int low_analysis_launch(const AnalysisArgs& g, const void* d_low, cudaStream_t st, long long* k_lo, long long* k_hi) {
  *k_lo = *k_hi = 0;
  return 0;
}

template <>
inline int low_analysis_launch<float>(const AnalysisArgs& g, const void* d_low, cudaStream_t st, long long* k_lo,
                                      long long* k_hi) {
  *k_lo = *k_hi = 0;
  const low::LowAnalysis* h = reinterpret_cast<const low::LowAnalysis*>(d_low);
  if (g.kout_mod != 0) return 0;
  // fast blocks: window fully resident and rotation-free block 0
  // block k window = [start0 + 192k, start0 + 192k + 3072)
  long long lo = 0;
  if (g.start0 < 0) lo = (-g.start0 + 191) / 192;
  // rotation of block lo must be 0 (mod 256) and lo a multiple of 4 keeps (3k mod 4) static
  {
    int tries = 0;
    while (tries < 4 && (((g.phase0 + g.start0 + 192 * lo) % 256 + 256) % 256 != 0)) { ++lo; ++tries; }
    if (tries == 4) return 0;
  }
  const long long n_eff = g.npol == 1 ? (g.n_dat & ~1ll) : g.n_dat;   // whole 16-byte units for the TMA
  long long hi = (n_eff - 3072 - g.start0) / 192 + 1;        // blocks with window end <= n_dat
  if (n_eff - 3072 - g.start0 < 0) hi = 0;
  if (hi > g.nblocks) hi = g.nblocks;
  if (hi - lo < 64) return 0;
  const long long s0 = g.start0 + 192 * lo;
  // TMA source alignment: 16 bytes
  if (((uintptr_t)g.in + (uintptr_t)((s0 + 2880) * g.npol * 8)) % 16 != 0) return 0;
  if (g.npol == 1 && (((uintptr_t)g.in) % 16 != 0 || (s0 % 2) != 0)) return 0;
  if (g.in2 && (((uintptr_t)g.in2) % 16 != 0 || ((uintptr_t)g.in) % 16 != 0 || (g.n_own * g.npol * 8) % 16 != 0)) return 0;
  const long long nblk = hi - lo;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long resident = (long long)sms * (g.npol == 2 ? 2 : 4);
  long long waves = (nblk + resident * 512 - 1) / (resident * 512);
  if (waves < 1) waves = 1;
  long long L = (nblk + resident * waves - 1) / (resident * waves);
  L = (L + 15) / 16 * 16;
  if (L < 16) L = 16;
  const long long grid = (nblk + L - 1) / L;
  low::AnaArgs a;
  a.in = reinterpret_cast<const low::C*>(g.in);
  a.in2 = reinterpret_cast<const low::C*>(g.in2);
  a.n1 = g.in2 ? g.n_own : n_eff;
  a.out = reinterpret_cast<low::C*>(g.out);
  a.taps = h->taps; a.tw = h->tw;
  a.s0 = s0; a.n_local = n_eff; a.nblk = nblk; a.kout0 = g.kout0 + lo; a.L = (int)L;
  a.chan_first = g.chan_first; a.nchan_out = g.nchan_out;
  cudaError_t e = cudaSuccess;
  const bool full = g.chan_first == 0 && g.nchan_out == 256;
  auto go = [&](auto kernel, size_t smem, int threads) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return;
    kernel<<<(unsigned)grid, threads, smem, st>>>(a);
  };
  a.v = reinterpret_cast<char*>(g.vout);
  a.v_rows_cap = g.v_rows_cap;
  if (g.emit_mode > 0) {
    if (!(g.npol == 2 && full && g.vout)) return (int)cudaErrorInvalidValue;
    if (g.emit_mode == 1) go(low::pfb_analysis_low<2, true, 1>, low::ana_smem_bytes<2>(), 128);
    else go(low::pfb_analysis_low<2, true, 2>, low::ana_smem_bytes<2>(), 128);
  } else if (g.npol == 2) {
    if (full) go(low::pfb_analysis_low<2, true>, low::ana_smem_bytes<2>(), 128);
    else go(low::pfb_analysis_low<2, false>, low::ana_smem_bytes<2>(), 128);
  } else {
    if (full) go(low::pfb_analysis_low<1, true>, low::ana_smem_bytes<1>(), 64);
    else go(low::pfb_analysis_low<1, false>, low::ana_smem_bytes<1>(), 64);
  }
  if (e != cudaSuccess) return (int)e;
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  *k_lo = lo;
  *k_hi = hi;
  return 0;
}

// ---- inversion ---------------------------------------------------------------
template <typename T>
int low_inverse_plan_init(int nchan, int nu, int de, int Nf, int ov, int spans, int combine,
                          const std::vector<T>& taper, const void* d_wt, std::vector<void*>* allocs,
                          const void** d_low, bool* enabled) {
  *enabled = false;
  return 0;
}

template <>
inline int low_inverse_plan_init<float>(int nchan, int nu, int de, int Nf, int ov, int spans, int combine,
                                        const std::vector<float>& taper, const void* d_wt,
                                        std::vector<void*>* allocs, const void** d_low, bool* enabled) {
  *enabled = false;
  if (!(Nf == 256 && nu == 4 && de == 3 && nchan % 16 == 0)) return 0;
  low::LowInverse* h = new low::LowInverse();
  int rc = low::upload_raw(taper, allocs, &h->taper);
  std::vector<float2> t256 = low::unit_table(256, -1), t192 = low::unit_table(192, +1);
  if (!rc) rc = low::upload_raw(t256, allocs, &h->tw256);
  if (!rc) rc = low::upload_raw(t192, allocs, &h->tw192);
  if (rc) { delete h; return rc; }
  h->wt = reinterpret_cast<const low::C*>(d_wt);
  h->nchan = nchan; h->keep = Nf - 2 * ov; h->ov2 = ov * de / nu; h->keep2 = 192 - 2 * h->ov2;
  h->flip = spans ? 0 : 1;
  h->chan_fast = (nchan == 256 && combine == 1) ? 1 : 0;
  h->taper_unity_mid = true;
  for (int i = ov; i < Nf - ov; ++i)
    if (taper[(size_t)i] != 1.0f) h->taper_unity_mid = false;
  *d_low = h;
  *enabled = true;
  return 0;
}

template <typename T>
int low_chan_idft_launch(const void* d_low, const void* in, void* out, int npol, long long nrows, cudaStream_t st, bool* done,
                         bool grouped = false) {
  *done = false;
  return 0;
}

// true when the warp-autonomous resample kernel (and the grouped scratch layout it reads) applies
inline bool low_ws_ok(const void* d_low, int npol, int groups_b, const void* in, const void* out) {
  const low::LowInverse* h = reinterpret_cast<const low::LowInverse*>(d_low);
  return h->chan_fast && npol == 2 && groups_b == 1 && h->ov2 == 36 && h->keep == 160 && !h->flip &&
         ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0);
}

template <>
inline int low_chan_idft_launch<float>(const void* d_low, const void* in, void* out, int npol, long long nrows,
                                       cudaStream_t st, bool* done, bool grouped) {
  *done = false;
  const low::LowInverse* h = reinterpret_cast<const low::LowInverse*>(d_low);
  if (!h->chan_fast) return 0;
  low::ChanArgs a;
  a.in = reinterpret_cast<const low::C*>(in);
  a.out = reinterpret_cast<low::C*>(out);
  a.tw256 = h->tw256;
  a.nrows = nrows;
  const size_t smem = (size_t)2 * 16 * low::kColStride * sizeof(low::C);
  const int rows = 16 / npol;
  long long npass = (nrows + rows - 1) / rows;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = npass < (long long)sms * 3 ? npass : (long long)sms * 3;
  cudaError_t e;
  if (npol == 2 && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0)) {
    const size_t smem2 = (size_t)low::kChanStages * low::kChanRows * low::kColStride * sizeof(float4) + 64;
    const long long npass2 = (nrows + low::kChanRows - 1) / low::kChanRows;
    const long long grid2 = npass2 < (long long)sms * 2 ? npass2 : (long long)sms * 2;
    a.rows_cap = nrows;
    if (grouped) {
      e = cudaFuncSetAttribute(low::pfb_chan_idft_low_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      if (e != cudaSuccess) return (int)e;
      low::pfb_chan_idft_low_tma<true><<<(unsigned)grid2, 128, smem2, st>>>(a);
    } else {
      e = cudaFuncSetAttribute(low::pfb_chan_idft_low_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      if (e != cudaSuccess) return (int)e;
      low::pfb_chan_idft_low_tma<false><<<(unsigned)grid2, 128, smem2, st>>>(a);
    }
  } else if (grouped) {
    return (int)cudaErrorInvalidValue;
  } else if (npol == 2) {
    e = cudaFuncSetAttribute(low::pfb_chan_idft_low<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    low::pfb_chan_idft_low<2><<<(unsigned)grid, 256, smem, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(low::pfb_chan_idft_low<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    low::pfb_chan_idft_low<1><<<(unsigned)grid, 256, smem, st>>>(a);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  *done = true;
  return 0;
}

// preferred divisor of the chunk length (blocks) so every CTA of the resample wave gets equal work
inline long long low_resample_nsplit(const void* d_low, int npol, int groups_b = 1) {
  const low::LowInverse* h = reinterpret_cast<const low::LowInverse*>(d_low);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int groups = h->nchan * groups_b / (16 / npol);   // 16 columns per CTA, 2 CTAs per SM (either resample kernel)
  long long n = ((long long)sms * 2) / groups;
  return n < 1 ? 1 : n;
}

template <typename T>
int low_resample_launch(const void* d_low, const void* v, void* out, int npol, long long nblk, int groups_b, cudaStream_t st, bool* done,
                        bool grouped = false, long long rows_cap = 0, long long vec_rows = 0, long long vec_blocks = 0) {
  *done = false;
  return 0;
}

template <>
inline int low_resample_launch<float>(const void* d_low, const void* v, void* out, int npol, long long nblk,
                                      int groups_b, cudaStream_t st, bool* done, bool grouped, long long rows_cap,
                                      long long vec_rows, long long vec_blocks) {
  *done = false;
  const low::LowInverse* h = reinterpret_cast<const low::LowInverse*>(d_low);
  if (nblk > 2000000000ll) return grouped ? (int)cudaErrorInvalidValue : 0;
  if (grouped) {
    low::WsArgs w;
    w.v = reinterpret_cast<const char*>(v);
    w.out = reinterpret_cast<char*>(out);
    w.taper = h->taper; w.wt = h->wt; w.tw256 = h->tw256; w.tw192 = h->tw192;
    w.rows_cap = rows_cap; w.nchan = h->nchan; w.nblk = (int)nblk;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int groups = h->nchan / low::kWsWarps;
    long long ns = ((long long)sms * 2) / groups;     // one resident wave, 2 CTAs per SM
    if (ns > nblk) ns = nblk;
    if (ns < 1) ns = 1;
    if (vec_blocks > 0) ns = nblk / vec_blocks;       // independent vectors: one time split per vector
    w.nsplit = (int)ns;
    w.split_rows = vec_blocks > 0 ? vec_rows : ((nblk + ns - 1) / ns) * (long long)low::kWsStepRows;
    // tensor map of the output rows: [nblk * 120 rows][nchan * 4 fp32], box = 120 rows x 128 B
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
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {(cuuint64_t)h->nchan * 4, (cuuint64_t)nblk * low::kWsOutRows};
    const cuuint64_t gstride[1] = {(cuuint64_t)h->nchan * 16};
    const cuuint32_t box[2] = {low::kWsRowBytes / 4, low::kWsOutRows};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
    auto kern = h->taper_unity_mid ? low::pfb_resample_low_ws<true> : low::pfb_resample_low_ws<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)low::ws_smem_bytes());
    if (e != cudaSuccess) return (int)e;
    kern<<<(unsigned)(ns * groups), 32 * (low::kWsWarps + 1), low::ws_smem_bytes(), st>>>(w, tm);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    *done = true;
    return 0;
  }
  low::ResArgs a;
  a.v = reinterpret_cast<const low::C*>(v);
  a.out = reinterpret_cast<low::C*>(out);
  a.taper = h->taper; a.wt = h->wt; a.tw256 = h->tw256; a.tw192 = h->tw192;
  const int nchan_total = h->nchan * groups_b;        // batched groups are extra columns
  a.nchan = nchan_total; a.wt_mod = h->nchan; a.keep = h->keep; a.ov2 = h->ov2; a.keep2 = h->keep2; a.flip = h->flip;
  a.nblk = (int)nblk;
  const int groups = nchan_total / (16 / npol);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long nsplit = ((long long)sms * 2) / groups;    // at most one resident wave (2 CTAs per SM)
  if (nsplit > nblk) nsplit = nblk;
  if (nsplit < 1) nsplit = 1;
  a.nsplit = (int)nsplit;
  const long long grid = nsplit * groups;
  cudaError_t e = cudaSuccess;
  constexpr int COLS = 16;
  {
    const int groupsv = nchan_total / (COLS / npol);
    long long ns = ((long long)sms * (512 / (16 * COLS))) / groupsv;
    if (ns > nblk) ns = nblk;
    if (ns < 1) ns = 1;
    a.nsplit = (int)ns;
  }
  const long long gridv = (long long)a.nsplit * (nchan_total / (COLS / npol));
  auto go = [&](auto kernel, size_t smem) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return;
    kernel<<<(unsigned)gridv, 16 * COLS, smem, st>>>(a);
  };
  const bool low_shape = nchan_total == 256 && h->ov2 == 36;   // the SKA-Low PST configuration
  if (npol == 2) {
    if (low_shape && !h->flip) go(low::pfb_resample_low<2, 256, 36, false, COLS>, low::res_smem_bytes<COLS>());
    else if (h->flip) go(low::pfb_resample_low<2, 0, -1, true, COLS>, low::res_smem_bytes<COLS>());
    else go(low::pfb_resample_low<2, 0, -1, false, COLS>, low::res_smem_bytes<COLS>());
  } else {
    if (low_shape && !h->flip) go(low::pfb_resample_low<1, 256, 36, false, COLS>, low::res_smem_bytes<COLS>());
    else if (h->flip) go(low::pfb_resample_low<1, 0, -1, true, COLS>, low::res_smem_bytes<COLS>());
    else go(low::pfb_resample_low<1, 0, -1, false, COLS>, low::res_smem_bytes<COLS>());
  }
  if (e != cudaSuccess) return (int)e;
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  *done = true;
  return 0;
}

}  // namespace pfb
