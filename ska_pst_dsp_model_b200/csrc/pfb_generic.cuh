// Generic (any N / taps / OS ratio / precision) kernels for the PFB hot path.
// These are the always-correct path; the Low-specific fast kernels live in
// pfb_low.cuh and are selected by the plan when the shape matches.
#pragma once
#include <type_traits>
#include "fft_smem.cuh"

namespace pfb {

// ---------------------------------------------------------------------------
// K1g: polyphase FIR commutator + N-point forward FFT + phase rotation.
// Covers polyphase_analysis.m:83-122, polyphase_analysis_padded.m:106-156 and
// PSTFilterbank.m:25-45 through one closed form:
//   out[c,k] = scale * sum_n f[n] x[t0_k+n] W^{c (phase0+t0_k+n)},  t0_k = start0 + k S
// The W^{c*shift} rotation is free: the arm sum for tap index n is written to
// FFT input slot (phase0+t0_k+n) mod N.
// ---------------------------------------------------------------------------
struct AnalysisArgs {
  const void* in;        // [t][pol] complex
  void* out;             // [k][chan][pol] complex
  const void* taps;      // T[P*N] (zero padded; reversed for the padded variant)
  FftDev fft;            // forward N
  int N, S, P, npol;
  long long n_dat;       // valid samples in `in`
  long long start0;      // window start of block 0 (relative to in[0], may be < 0)
  long long phase0;      // absolute index of in[0] + extra rotation
  long long nblocks;     // blocks computed by this launch
  long long kout0;       // output block slot = (k + kout0) mod kout_mod
  long long kout_mod;
  int chan_first;        // FFT bin of output channel 0
  int nchan_out;
  double scale;
  int kb;                // blocks per CTA
  // batched form: `in` holds in_pitch independent streams per time sample; blockIdx.y selects a tile of
  // npol of them; output element (k, chan, stream) lives at k*out_sk + chan*out_sc + stream*out_sp
  long long in_pitch, out_sk, out_sc, out_sp;
  // two-kernel form: arm sums already computed by pfb_fir_slide into ysrc[k][slot][pol]
  const void* ysrc;
  // two-buffer form (fast SKA-Low kernel only): samples n_own.. of the stream live in in2 (a peer GPU's memory)
  const void* in2;
  long long n_own;
  // fused round trip (fast SKA-Low kernel only): 1 = also write the inversion intermediate v, 2 = v only
  int emit_mode;
  void* vout;
  long long v_rows_cap;
};

template <typename T>
__global__ void __launch_bounds__(256) pfb_analysis_generic(const AnalysisArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cx<T>* s = reinterpret_cast<Cx<T>*>(smem_raw);
  const int N = a.N, npol = a.npol, ld = padded_len<T>(N);
  const long long k0 = (long long)blockIdx.x * a.kb;
  const int kb = (int)min((long long)a.kb, a.nblocks - k0);
  const int pol0 = blockIdx.y * npol;
  const Cx<T>* __restrict__ x = reinterpret_cast<const Cx<T>*>(a.in) + pol0;
  const T* __restrict__ f = reinterpret_cast<const T*>(a.taps);

  // FIR: work item = (kk, j, pol), pol fastest -> coalesced over [t][pol]
  const int items = kb * N * npol;
  if (a.ysrc) {
    const Cx<T>* __restrict__ y = reinterpret_cast<const Cx<T>*>(a.ysrc) + k0 * (long long)N * npol;
    for (int e = threadIdx.x; e < items; e += blockDim.x) {
      const int pol = e % npol;
      const int slot = (e / npol) % N;
      const int kk = e / (npol * N);
      s[(size_t)(kk * npol + pol) * ld + padi<T>(slot)] = ldg_cx<T>(y + e);
    }
  } else
  for (int e = threadIdx.x; e < items; e += blockDim.x) {
    const int pol = e % npol;
    const int j = (e / npol) % N;
    const int kk = e / (npol * N);
    const long long t0 = a.start0 + (k0 + kk) * (long long)a.S;
    // taps whose sample lies inside the resident stream: m in [m_lo, m_hi); four independent chains
    const long long tj = t0 + j;
    long long m_lo = tj >= 0 ? 0 : (-tj + N - 1) / N;
    long long m_hi = tj < a.n_dat ? (a.n_dat - tj + N - 1) / N : 0;
    if (m_hi > a.P) m_hi = a.P;
    if (m_lo > m_hi) m_lo = m_hi;
    T ax[4] = {0, 0, 0, 0}, ay[4] = {0, 0, 0, 0};
    const Cx<T>* __restrict__ xp = x + (tj + m_lo * N) * a.in_pitch + pol;
    const T* __restrict__ fp = f + j + m_lo * N;
    const long long xs = (long long)N * a.in_pitch;
    int m = (int)m_lo;
    for (; m + 4 <= (int)m_hi; m += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const Cx<T> v = ldg_cx<T>(xp + u * xs);
        const T c = __ldg(fp + u * N);
        ax[u] += c * v.x;
        ay[u] += c * v.y;
      }
      xp += 4 * xs;
      fp += 4 * N;
    }
    for (; m < (int)m_hi; ++m) {
      const Cx<T> v = ldg_cx<T>(xp);
      const T c = __ldg(fp);
      ax[0] += c * v.x;
      ay[0] += c * v.y;
      xp += xs;
      fp += N;
    }
    const T accx = (ax[0] + ax[1]) + (ax[2] + ax[3]), accy = (ay[0] + ay[1]) + (ay[2] + ay[3]);
    const int slot = pmod(a.phase0 + t0 + j, N);
    s[(size_t)(kk * npol + pol) * ld + padi<T>(slot)] = mk<T>(accx, accy);
  }
  __syncthreads();
  fft_smem<T, -1, false>(s, ld, kb * npol, a.fft);

  Cx<T>* __restrict__ out = reinterpret_cast<Cx<T>*>(a.out);
  const T scale = (T)a.scale;
  const int nco = a.nchan_out;
  const int oitems = kb * nco * npol;
  for (int e = threadIdx.x; e < oitems; e += blockDim.x) {
    const int pol = e % npol;
    const int ci = (e / npol) % nco;
    const int kk = e / (npol * nco);
    int c = ci + a.chan_first;
    if (c >= N) c -= N;
    const Cx<T> v = s[(size_t)(kk * npol + pol) * ld + padi<T>(__ldg(a.fft.perm + c))];
    long long ko = k0 + kk + a.kout0;
    if (a.kout_mod > 0) ko %= a.kout_mod;
    out[ko * a.out_sk + ci * a.out_sc + (pol0 + pol) * a.out_sp] = v * scale;
  }
}

// ---------------------------------------------------------------------------
// K1s: sliding-window polyphase FIR for long filters (SKA-Mid: 25 taps/arm over 4096 arms, the
// taps do not fit shared memory).  Blocks k and k + nu read the same sub-stream
// x[k0*S + j + N*q] shifted by `de` elements (nu*S = de*N), so a thread that owns (arm j, pol,
// k mod nu) keeps its P taps and a P-element window in registers and fetches only `de` new
// samples per output (de/P loads per MAC instead of 2).  The rotated arm sums go to a scratch
// array y[k][slot][pol]; pfb_analysis_generic (ysrc != null) then runs the FFT + store.
// Measured bound (SKA-Mid, ncu): the nu = 8 block phases read the same sub-streams at different offsets, so
// the input crosses L2 -> SM eight times (537 MB per 2^22 x 2 pol launch, ~5 TB/s); a TMA-fed ring per
// (64 arms, phase) CTA was tried and is no faster for that reason.  Sharing the rows between the phases of
// one CTA (32 arms x 8 phases x 2 pols) is the next step.
// ---------------------------------------------------------------------------
struct FirSlideArgs {
  const void* in;
  void* y;
  const void* taps;      // T[P*N]
  int N, S, npol, nu;
  long long n_dat, start0, phase0, nblocks;
  long long steps_per_cta;
  long long i0, i1;      // step range [i0, i1) of this launch (a step = one block of every phase k0)
};

__host__ __device__ constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }

template <typename T, int P, int DE>
__global__ void __launch_bounds__(128, sizeof(T) == 4 ? 3 : 2) pfb_fir_slide(const FirSlideArgs a) {
  // The window lives in a circular register file of W >= P slots with W a multiple of DE, so the
  // rotation repeats every W / DE steps (4 for the SKA-Mid shape) and only that many steps are unrolled.
  constexpr int NNEW = DE < P ? DE : P;                    // new window elements per step
  constexpr int NEW0 = P - NNEW;                           // logical window elements [NEW0, P) are new each step
  constexpr int W = DE < P ? (P + DE - 1) / DE * DE : P;
  constexpr int PER = DE < P ? W / DE : 1;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.N * a.npol) return;
  const int pol = e % a.npol, j = e / a.npol;
  const int k0 = blockIdx.y;
  const long long nsteps = (a.nblocks - k0 + a.nu - 1) / a.nu;
  const long long i_begin = a.i0 + (long long)blockIdx.z * a.steps_per_cta;
  const long long i_end = min(min(i_begin + a.steps_per_cta, a.i1), nsteps);
  if (i_begin >= i_end) return;
  const T* __restrict__ f = reinterpret_cast<const T*>(a.taps);
  const Cx<T>* __restrict__ x = reinterpret_cast<const Cx<T>*>(a.in);
  Cx<T>* __restrict__ y = reinterpret_cast<Cx<T>*>(a.y);
  T tap[P];
#pragma unroll
  for (int m = 0; m < P; ++m) tap[m] = __ldg(f + j + m * a.N);
  // sub-stream element q sits at sample tbase + N*q; valid (inside the resident samples) for q in [q_lo, q_hi)
  const long long tbase = a.start0 + (long long)k0 * a.S + j;
  const long long q_lo = tbase >= 0 ? 0 : (-tbase + a.N - 1) / a.N;
  const long long q_hi = tbase < a.n_dat ? (a.n_dat - tbase + a.N - 1) / a.N : 0;
  const Cx<T>* __restrict__ xs = x + tbase * a.npol + pol;
  const long long qstride = (long long)a.N * a.npol;
  // every element this thread will touch is resident: unchecked loads off a running pointer
  const bool interior = (long long)DE * i_begin >= q_lo && (long long)DE * (i_end - 1) + P <= q_hi;
  auto ld = [&](long long q) {
    return (q >= q_lo && q < q_hi) ? ldg_cx<T>(xs + q * qstride) : mk<T>((T)0, (T)0);
  };
  // nu*S = de*N: every block of this thread lands in the same FFT slot
  const int slot = pmod(a.phase0 + a.start0 + (long long)k0 * a.S + j, a.N);
  Cx<T>* __restrict__ yo = y + ((long long)k0 * a.N + slot) * a.npol + pol + i_begin * ((long long)a.nu * a.N * a.npol);
  const long long ystride = (long long)a.nu * a.N * a.npol;
  Cx<T> w[W];
#pragma unroll
  for (int m = 0; m < NEW0; ++m) w[m] = ld((long long)DE * i_begin + m);
  const Cx<T>* __restrict__ pn = xs + ((long long)DE * i_begin + NEW0) * qstride;     // first new element of the step
  const long long pstep = (long long)DE * qstride;

  auto run = [&](auto checked) {
    constexpr bool CHECK = decltype(checked)::value;
    for (long long i = i_begin; i < i_end; i += PER) {
#pragma unroll
      for (int sidx = 0; sidx < PER; ++sidx) {
        if (i + sidx < i_end) {
#pragma unroll
          for (int m = 0; m < NNEW; ++m) {
            if (CHECK) w[(DE * sidx + NEW0 + m) % W] = ld((long long)DE * (i + sidx) + NEW0 + m);
            else w[(DE * sidx + NEW0 + m) % W] = ldg_cx<T>(pn + m * qstride);
          }
          pn += pstep;
          T ax0 = 0, ay0 = 0, ax1 = 0, ay1 = 0;              // two independent chains per component
#pragma unroll
          for (int m = 0; m < P; ++m) {
            const Cx<T> v = w[(DE * sidx + m) % W];
            if (m & 1) { ax1 += tap[m] * v.x; ay1 += tap[m] * v.y; }
            else { ax0 += tap[m] * v.x; ay0 += tap[m] * v.y; }
          }
          *yo = mk<T>(ax0 + ax1, ay0 + ay1);
          yo += ystride;
        }
      }
    }
  };
  if (interior) run(std::false_type());
  else run(std::true_type());
}

// ---------------------------------------------------------------------------
// K2g: the nchan-point inverse DFT across channels, hoisted in front of the
// per-channel FFTs (polyphase_synthesis.m:285 factored as nchan x FN_width,
// SURVEY §7 design note).  One row per fine-channel time sample.
//   v[tau][t1][pol] = sum_ch u[tau][jmap[ch]][pol] exp(+2 pi i ch t1 / nchan)
// ---------------------------------------------------------------------------
struct ChanIdftArgs {
  const void* in;       // [tau][chan][pol]
  void* out;            // [tau][t1][pol]
  FftDev fft;           // nchan (used as inverse, DIT)
  const int* jmap;      // combine re-order (polyphase_synthesis.m:198-238) or null
  int nchan, npol;
  long long nrows;
  int rb;               // rows per CTA
};

template <typename T>
__global__ void __launch_bounds__(256) pfb_chan_idft_generic(const ChanIdftArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cx<T>* s = reinterpret_cast<Cx<T>*>(smem_raw);
  const int n = a.nchan, npol = a.npol, ld = padded_len<T>(n);
  const long long r0 = (long long)blockIdx.x * a.rb;
  const int rb = (int)min((long long)a.rb, a.nrows - r0);
  const Cx<T>* __restrict__ u = reinterpret_cast<const Cx<T>*>(a.in);
  const int items = rb * n * npol;
  for (int e = threadIdx.x; e < items; e += blockDim.x) {
    const int pol = e % npol;
    const int ch = (e / npol) % n;
    const int r = e / (npol * n);
    const int src = a.jmap ? __ldg(a.jmap + ch) : ch;
    const Cx<T> v = ldg_cx<T>(u + ((r0 + r) * n + src) * npol + pol);
    s[(size_t)(r * npol + pol) * ld + padi<T>(__ldg(a.fft.perm + ch))] = v;
  }
  __syncthreads();
  fft_smem<T, +1, true>(s, ld, rb * npol, a.fft);
  Cx<T>* __restrict__ out = reinterpret_cast<Cx<T>*>(a.out);
  for (int e = threadIdx.x; e < items; e += blockDim.x) {
    const int pol = e % npol;
    const int t1 = (e / npol) % n;
    const int r = e / (npol * n);
    out[((r0 + r) * n + t1) * npol + pol] = s[(size_t)(r * npol + pol) * ld + padi<T>(t1)];
  }
}

// ---------------------------------------------------------------------------
// K3g: per output phase t1 (column of v) overlap-save resampling:
// taper -> Nf-point FFT -> keep FN_width bins (fftshift + discard) -> deripple x
// stitch twiddle x 1/(OS*oN) -> FN_width-point inverse FFT -> drop the output
// overlap.  polyphase_synthesis.m:163-302 for one column.
//   out[(n*keep2 + t2 - ov2)*nchan + t1] = sum_r Z[r] exp(2 pi i (r-off) t2/FNw),
//   Z[r] = W[t1][r] * FFT_Nf(taper * v[n*keep : n*keep+Nf, t1])[(r+d2+Nf/2) mod Nf]
// ---------------------------------------------------------------------------
struct ResampleArgs {
  const void* v;        // [tau][t1][pol]  (local chunk, row 0 = first tau of block 0)
  void* out;            // [T2][t1][pol]   (local chunk)
  FftDev fwd;           // Nf forward (DIF)
  FftDev bwd;           // FNw inverse (DIT)
  const void* taper;    // T[Nf]
  const void* wt;       // Cx<T>[nchan][FNw]
  int nchan, npol, Nf, FNw, keep, ov2, keep2, off, d2;
  int tg;               // t1 values per CTA
  int nblk;             // overlap-save blocks in this launch
  int wt_mod;           // channels per group: column t1 uses weight row t1 % wt_mod (batched groups)
};

template <typename T>
__global__ void __launch_bounds__(256) pfb_resample_generic(const ResampleArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int ncol = a.tg * a.npol;
  const int ldA = padded_len<T>(a.Nf), ldB = padded_len<T>(a.FNw);
  Cx<T>* sA = reinterpret_cast<Cx<T>*>(smem_raw);
  Cx<T>* sB = sA + (size_t)ncol * ldA;
  const int groups = a.nchan / a.tg;
  const int g = blockIdx.x % groups;
  const long long nb = blockIdx.x / groups;
  const int t1_0 = g * a.tg;
  const long long rowlen = (long long)a.nchan * a.npol;
  const Cx<T>* __restrict__ v = reinterpret_cast<const Cx<T>*>(a.v) + nb * a.keep * rowlen + (long long)t1_0 * a.npol;
  const T* __restrict__ taper = reinterpret_cast<const T*>(a.taper);

  for (int e = threadIdx.x; e < a.Nf * ncol; e += blockDim.x) {
    const int col = e % ncol, tau = e / ncol;
    const Cx<T> val = ldg_cx<T>(v + tau * rowlen + col);
    sA[(size_t)col * ldA + padi<T>(tau)] = val * __ldg(taper + tau);
  }
  __syncthreads();
  fft_smem<T, -1, false>(sA, ldA, ncol, a.fwd);

  const Cx<T>* __restrict__ wt = reinterpret_cast<const Cx<T>*>(a.wt);
  for (int e = threadIdx.x; e < a.FNw * ncol; e += blockDim.x) {
    const int r = e % a.FNw, col = e / a.FNw;
    const int t1 = (t1_0 + col / a.npol) % a.wt_mod;
    int b = r + a.d2 + a.Nf / 2;
    if (b >= a.Nf) b -= a.Nf;
    int rp = r - a.off;
    if (rp < 0) rp += a.FNw;
    const Cx<T> z = sA[(size_t)col * ldA + padi<T>(__ldg(a.fwd.perm + b))] * ldg_cx<T>(wt + (size_t)t1 * a.FNw + r);
    sB[(size_t)col * ldB + padi<T>(__ldg(a.bwd.perm + rp))] = z;
  }
  __syncthreads();
  fft_smem<T, +1, true>(sB, ldB, ncol, a.bwd);

  Cx<T>* __restrict__ out = reinterpret_cast<Cx<T>*>(a.out) + nb * a.keep2 * rowlen + (long long)t1_0 * a.npol;
  for (int e = threadIdx.x; e < a.keep2 * ncol; e += blockDim.x) {
    const int col = e % ncol, i = e / ncol;
    out[i * rowlen + col] = sB[(size_t)col * ldB + padi<T>(i + a.ov2)];
  }
}

// ---------------------------------------------------------------------------
// Literal-order chain, used only when a spectral (frequency-domain) taper is set
// (InverseFilterBank.frequency_taper, polyphase_synthesis.m:282).  The taper multiplies the
// stitched spectrum by a weight that depends on (channel, bin) jointly, so the channel IDFT
// cannot be hoisted in front of the per-channel FFTs:
//   L1 pfb_fwd_select_generic : per channel: taper -> Nf-FFT -> keep FNw bins x D[r] x T[kk]
//   (K2g)                     : nchan-point inverse DFT across channels for every (block, bin)
//   L3 pfb_inv_store_generic  : per output phase t1: x stitch twiddle x gain -> FNw-IFFT -> keep
// ---------------------------------------------------------------------------
struct FwdSelArgs {
  const void* in;       // [tau][chan][pol]
  void* z;              // [n][r][chan][pol]
  FftDev fwd;
  const void* taper;    // T[Nf]
  const void* wa;       // T[nchan][FNw]  deripple x spectral taper (real)
  const int* jmap;
  int nchan, npol, Nf, FNw, keep, d2;
  int tg, nblk;
};

template <typename T>
__global__ void __launch_bounds__(256) pfb_fwd_select_generic(const FwdSelArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cx<T>* sA = reinterpret_cast<Cx<T>*>(smem_raw);
  const int ncol = a.tg * a.npol, ldA = padded_len<T>(a.Nf);
  const int groups = a.nchan / a.tg;
  const int g = blockIdx.x % groups;
  const long long nb = blockIdx.x / groups;
  const int ch0 = g * a.tg;
  const long long rowlen = (long long)a.nchan * a.npol;
  const Cx<T>* __restrict__ u = reinterpret_cast<const Cx<T>*>(a.in) + nb * a.keep * rowlen;
  const T* __restrict__ taper = reinterpret_cast<const T*>(a.taper);
  for (int e = threadIdx.x; e < a.Nf * ncol; e += blockDim.x) {
    const int col = e % ncol, tau = e / ncol;
    const int ch = ch0 + col / a.npol, pol = col % a.npol;
    const int src = a.jmap ? __ldg(a.jmap + ch) : ch;
    sA[(size_t)col * ldA + padi<T>(tau)] = ldg_cx<T>(u + tau * rowlen + (long long)src * a.npol + pol) * __ldg(taper + tau);
  }
  __syncthreads();
  fft_smem<T, -1, false>(sA, ldA, ncol, a.fwd);
  const T* __restrict__ wa = reinterpret_cast<const T*>(a.wa);
  Cx<T>* __restrict__ z = reinterpret_cast<Cx<T>*>(a.z) + nb * a.FNw * rowlen;
  for (int e = threadIdx.x; e < a.FNw * ncol; e += blockDim.x) {
    const int col = e % ncol, r = e / ncol;
    const int ch = ch0 + col / a.npol, pol = col % a.npol;
    int b = r + a.d2 + a.Nf / 2;
    if (b >= a.Nf) b -= a.Nf;
    z[(long long)r * rowlen + (long long)ch * a.npol + pol] =
        sA[(size_t)col * ldA + padi<T>(__ldg(a.fwd.perm + b))] * __ldg(wa + (size_t)ch * a.FNw + r);
  }
}

struct InvStoreArgs {
  const void* A;        // [n][r][t1][pol]
  void* out;            // [T2][t1][pol]
  FftDev bwd;
  const void* wb;       // Cx<T>[nchan][FNw]: gain x exp(2 pi i (r-off) t1 / oN)
  int nchan, npol, FNw, ov2, keep2, off;
  int tg, nblk;
};

template <typename T>
__global__ void __launch_bounds__(256) pfb_inv_store_generic(const InvStoreArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cx<T>* sB = reinterpret_cast<Cx<T>*>(smem_raw);
  const int ncol = a.tg * a.npol, ldB = padded_len<T>(a.FNw);
  const int groups = a.nchan / a.tg;
  const int g = blockIdx.x % groups;
  const long long nb = blockIdx.x / groups;
  const int t1_0 = g * a.tg;
  const long long rowlen = (long long)a.nchan * a.npol;
  const Cx<T>* __restrict__ A = reinterpret_cast<const Cx<T>*>(a.A) + nb * a.FNw * rowlen + (long long)t1_0 * a.npol;
  const Cx<T>* __restrict__ wb = reinterpret_cast<const Cx<T>*>(a.wb);
  for (int e = threadIdx.x; e < a.FNw * ncol; e += blockDim.x) {
    const int col = e % ncol, r = e / ncol;
    const int t1 = t1_0 + col / a.npol;
    int rp = r - a.off;
    if (rp < 0) rp += a.FNw;
    sB[(size_t)col * ldB + padi<T>(__ldg(a.bwd.perm + rp))] =
        ldg_cx<T>(A + (long long)r * rowlen + col) * ldg_cx<T>(wb + (size_t)t1 * a.FNw + r);
  }
  __syncthreads();
  fft_smem<T, +1, true>(sB, ldB, ncol, a.bwd);
  Cx<T>* __restrict__ out = reinterpret_cast<Cx<T>*>(a.out) + nb * a.keep2 * rowlen + (long long)t1_0 * a.npol;
  for (int e = threadIdx.x; e < a.keep2 * ncol; e += blockDim.x) {
    const int col = e % ncol, i = e / ncol;
    out[i * rowlen + col] = sB[(size_t)col * ldB + padi<T>(i + a.ov2)];
  }
}

}  // namespace pfb
