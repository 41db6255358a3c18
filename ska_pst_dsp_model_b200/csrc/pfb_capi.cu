// C ABI (include/pfb_b200.h) + host-side plan construction for the B200 PFB
// analysis / inversion kernels.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/pfb_b200.h"
#include "pfb_generic.cuh"
#include "pfb_low.cuh"
#include "pfb_mid.cuh"
#include "pfb_payload.cuh"
#include "pfb_purity.cuh"
#include "pfb_small.cuh"

using namespace pfb;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) return fail(PFB_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

const double kPi = 3.14159265358979323846264338327950288;

struct DeviceTables {
  std::vector<void*> ptrs;
  ~DeviceTables() {
    for (void* p : ptrs) cudaFree(p);
  }
  template <typename U> int upload(const std::vector<U>& h, const U** out) {
    void* d = nullptr;
    CU(cudaMalloc(&d, h.size() * sizeof(U) + 16));
    ptrs.push_back(d);
    CU(cudaMemcpy(d, h.data(), h.size() * sizeof(U), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<const U*>(d);
    return PFB_OK;
  }
};

// exact-as-possible unit-circle point exp(sign * 2 pi i m / n)
void unit(long long m, long long n, int sign, double* c, double* s) {
  m %= n;
  if (m < 0) m += n;
  const double a = 2.0 * kPi * ((double)m / (double)n);
  *c = cos(a);
  *s = sign * sin(a);
  if (4 * m == n) { *c = 0; *s = sign; }
  else if (2 * m == n) { *c = -1; *s = 0; }
  else if (4 * m == 3 * n) { *c = 0; *s = -sign; }
  else if (m == 0) { *c = 1; *s = 0; }
}

int factorize(int n, std::vector<int>& radix) {
  radix.clear();
  if (n < 1) return PFB_ERR_ARG;
  int m = n;
  const int odd[3] = {7, 5, 3};
  for (int p : odd)
    while (m % p == 0) { radix.push_back(p); m /= p; }
  int e = 0;
  while (m % 2 == 0) { m /= 2; ++e; }
  if (m != 1) return PFB_ERR_ARG;
  if (e % 3 == 1) radix.push_back(2);
  if (e % 3 == 2) radix.push_back(4);
  for (int i = 0; i < e / 3; ++i) radix.push_back(8);
  if (radix.empty()) radix.push_back(1);
  if ((int)radix.size() > kMaxPass) return PFB_ERR_ARG;
  return PFB_OK;
}

template <typename T> int build_fft(int n, FftDev* dev, DeviceTables* tabs) {
  std::vector<int> radix;
  if (factorize(n, radix) != PFB_OK) return fail(PFB_ERR_ARG, "FFT length %d is not of the form 2^a 3^b 5^c 7^d", n);
  dev->n = n;
  if (n == 1) { dev->npass = 0; }
  else {
    dev->npass = (int)radix.size();
    for (int i = 0; i < dev->npass; ++i) dev->radix[i] = radix[i];
  }
  std::vector<Cx<T>> tw(n);
  for (int k = 0; k < n; ++k) {
    double c, s;
    unit(k, n, -1, &c, &s);
    tw[k].x = (T)c;
    tw[k].y = (T)s;
  }
  std::vector<int> perm(n);
  for (int k = 0; k < n; ++k) {
    int kk = k, pos = 0, stride = n;
    for (int l = 0; l < dev->npass; ++l) {
      const int R = dev->radix[l];
      stride /= R;
      pos += (kk % R) * stride;
      kk /= R;
    }
    perm[k] = pos;
  }
  const Cx<T>* dtw;
  const int* dperm;
  int rc = tabs->upload(tw, &dtw);
  if (rc) return rc;
  rc = tabs->upload(perm, &dperm);
  if (rc) return rc;
  dev->tw = dtw;
  dev->perm = dperm;
  return PFB_OK;
}

template <typename K> int set_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) return fail(PFB_ERR_ARG, "kernel needs %zu B of shared memory (> 227 KiB)", bytes);
  CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return PFB_OK;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PFB_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  return PFB_OK;
}

}  // namespace

// ===========================================================================
struct pfb_analysis_plan {
  int variant, N, nu, de, S, P, fl, ntaps, dtype, device;
  int nchan_out, chan_first, sds;
  double scale;
  FftDev fft;
  const void* d_taps = nullptr;     // T[fl]
  const void* d_low_taps = nullptr; // fast Low kernel tap layout
  bool low_fast = false;
  bool mid_fast = true;             // shape-specialised SKA-Mid kernels allowed (PFB_NO_MID, read once at plan creation)
  small8::Fir8Taps* fir8 = nullptr; // 8-channel 8/7 384-taps-per-arm shape (pfb_small.cuh): arm-major taps for the kernel parameters
  const float4* d_fir8_btab = nullptr;   // tensor-core form of the same FIR (PFB_FIR8_MMA=1 at plan creation): B fragments
  void* d_yscratch = nullptr;       // arm sums of the sliding-window FIR kernel
  size_t yscratch_cap = 0;
  DeviceTables tabs;
  ~pfb_analysis_plan() {
    if (d_yscratch) cudaFree(d_yscratch);
    delete fir8;
  }
};

struct pfb_inverse_plan {
  int nchan, nu, de, Nf, ov, keep, FNw, ov2, keep2, d2, off, spans, combine, dtype, device;
  FftDev fft_chan, fft_fwd, fft_bwd;
  const int* d_jmap = nullptr;
  const void* d_taper = nullptr;
  const void* d_wt = nullptr;
  const void* d_wa = nullptr;        // literal chain: T[nchan][FNw] deripple x spectral taper
  const void* d_wb = nullptr;        // literal chain: Cx<T>[nchan][FNw] stitch twiddle x gain
  bool spectral = false;
  std::vector<double> deripple_w;    // FN-order deripple weights (kept for set_spectral_taper)
  const void* d_low = nullptr;       // fast-path tables (pfb_low.cuh)
  bool low_fast = false;
  bool mid_fast = true;              // PFB_NO_MID / PFB_NO_WS: debugging switches, read once at plan creation
  bool ws_fast = true;
  bool taper_unity_mid = false;      // temporal taper == 1 on [ov, Nf - ov)
  long long scratch_bytes = 1024ll << 20;   // one chunk for streams up to ~2^25 samples/pol at the SKA-Low shape
  void* d_scratch = nullptr;
  size_t scratch_cap = 0;
  DeviceTables tabs;
  // optional per-kernel timing (pfb_inverse_set_profiling): event pairs per stage
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<std::pair<int, int>> ev_used[2];   // (start, stop) indices into ev_pool
  size_t ev_next = 0;
  int prof_begin(int stage, cudaStream_t st) {
    if (!profiling) return -1;
    while (ev_pool.size() < ev_next + 2) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return -1;
      ev_pool.push_back(e);
    }
    const int a = (int)ev_next, b = (int)ev_next + 1;
    ev_next += 2;
    ev_used[stage].push_back({a, b});
    cudaEventRecord(ev_pool[a], st);
    return b;
  }
  void prof_end(int tok, cudaStream_t st) {
    if (tok >= 0) cudaEventRecord(ev_pool[tok], st);
  }
  ~pfb_inverse_plan() {
    if (d_scratch) cudaFree(d_scratch);
    for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
  }
};

// (C linkage comes from the declarations in pfb_b200.h)
const char* pfb_last_error(void) { return g_err.c_str(); }
int pfb_version(void) { return 100; }
long long pfb_launch_count(void) { return g_launches.load(); }

// ---------------------------------------------------------------------------
template <typename T> static int analysis_build(pfb_analysis_plan* p, const double* taps) {
  int rc = build_fft<T>(p->N, &p->fft, &p->tabs);
  if (rc) return rc;
  std::vector<T> f(p->fl, (T)0);
  for (int i = 0; i < p->ntaps; ++i) f[i] = (T)taps[i];
  if (p->variant == PFB_VARIANT_PADDED) {
    std::vector<T> r(p->fl);
    for (int i = 0; i < p->fl; ++i) r[i] = f[p->fl - 1 - i];
    f.swap(r);
  }
  const T* d;
  rc = p->tabs.upload(f, &d);
  if (rc) return rc;
  p->d_taps = d;
  return PFB_OK;
}

int pfb_analysis_plan_create(pfb_analysis_plan** plan, int variant, int nchan, int os_nu, int os_de,
                             const double* taps, int ntaps, int dtype, int device) {
  if (!plan || !taps || ntaps < 1 || nchan < 1 || os_nu < 1 || os_de < 1 || os_de > os_nu)
    return fail(PFB_ERR_ARG, "pfb_analysis_plan_create: bad argument");
  if (dtype != PFB_DTYPE_C64 && dtype != PFB_DTYPE_C128) return fail(PFB_ERR_ARG, "dtype must be 0 (c64) or 1 (c128)");
  if (variant < 0 || variant > 2) return fail(PFB_ERR_ARG, "unknown analysis variant %d", variant);
  CU(cudaSetDevice(device));
  pfb_analysis_plan* p = new pfb_analysis_plan();
  p->variant = variant; p->N = nchan; p->nu = os_nu; p->de = os_de; p->dtype = dtype; p->device = device;
  p->ntaps = ntaps;
  p->mid_fast = getenv("PFB_NO_MID") == nullptr;
  p->S = (int)(((long long)nchan * os_de) / os_nu);          // polyphase_analysis.m:56
  p->P = (ntaps + nchan - 1) / nchan;                         // pad_filter.m:10
  p->fl = p->P * nchan;
  p->nchan_out = nchan; p->chan_first = 0; p->scale = nchan; p->sds = 0;
  if (p->S < 1) { delete p; return fail(PFB_ERR_ARG, "step < 1"); }
  if (variant == PFB_VARIANT_PADDED) {
    p->sds = (int)std::ceil((ntaps - 1) / 2.0 / p->S);        // polyphase_analysis_padded.m:89
  } else if (variant == PFB_VARIANT_LOWCBF) {
    if (nchan != 256 || os_nu != 4 || os_de != 3 || ntaps != 3072) {
      delete p;
      return fail(PFB_ERR_ARG, "lowcbf variant is fixed at 256 channels, 4/3, 3072 taps (PSTFilterbank.m:7-16)");
    }
    p->nchan_out = 216; p->chan_first = 148; p->scale = 4096.0;
  }
  int rc = dtype == PFB_DTYPE_C64 ? analysis_build<float>(p, taps) : analysis_build<double>(p, taps);
  if (rc == PFB_OK && variant != PFB_VARIANT_PADDED)
    rc = low_analysis_plan_init(p->N, p->nu, p->de, p->P, p->dtype, taps, p->ntaps, p->scale, &p->tabs.ptrs, &p->d_low_taps, &p->low_fast);
  if (rc > 0) rc = fail(PFB_ERR_CUDA, "fast-path table upload: %s", cudaGetErrorString((cudaError_t)rc));
  if (rc) { delete p; return rc; }
  if (variant == PFB_VARIANT_BUNTON && dtype == PFB_DTYPE_C64 && nchan == 8 && os_nu == 8 && os_de == 7 &&
      p->P == small8::kTaps && getenv("PFB_NO_SMALL8") == nullptr) {
    p->fir8 = new small8::Fir8Taps();
    for (int j = 0; j < 8; ++j)
      for (int m = 0; m < small8::kTaps; ++m) {
        const int n = j + 8 * m;
        p->fir8->t[j][m] = n < ntaps ? (float)taps[n] : 0.f;
      }
    if (getenv("PFB_FIR8_MMA") != nullptr) {
      // B fragments of mma.m16n8k8 (col-major 8 x 8): lane holds B[k = tig][n = gid] and B[tig + 4][gid];
      // B[t][a] = f_j[t - 7 a]; hi = TF32(x) (round to nearest, ties away: keep 10 mantissa bits), lo = TF32(x - hi)
      auto tf32 = [](float x) {
        uint32_t u;
        memcpy(&u, &x, 4);
        u = (u + 0x1000u) & 0xffffe000u;
        float r;
        memcpy(&r, &u, 4);
        return r;
      };
      std::vector<float4> bt((size_t)8 * small8::kMmaKSteps * 32);
      for (int j = 0; j < 8; ++j)
        for (int ks = 0; ks < small8::kMmaKSteps; ++ks)
          for (int lane = 0; lane < 32; ++lane) {
            const int gid = lane >> 2, tig = lane & 3;
            float v[2];
            for (int h = 0; h < 2; ++h) {
              const int m = 8 * ks + tig + 4 * h - 7 * gid;
              v[h] = (m >= 0 && m < small8::kTaps) ? p->fir8->t[j][m] : 0.f;
            }
            const float h0 = tf32(v[0]), h1 = tf32(v[1]);
            bt[((size_t)j * small8::kMmaKSteps + ks) * 32 + lane] = make_float4(h0, h1, tf32(v[0] - h0), tf32(v[1] - h1));
          }
      int rc2 = p->tabs.upload(bt, &p->d_fir8_btab);
      if (rc2) { delete p; return rc2; }
    }
  }
  *plan = p;
  return PFB_OK;
}

int pfb_analysis_plan_destroy(pfb_analysis_plan* plan) {
  if (plan) { cudaSetDevice(plan->device); delete plan; }
  return PFB_OK;
}

long long pfb_analysis_nblocks(const pfb_analysis_plan* p, long long n_dat, int first_call) {
  if (!p) return PFB_ERR_ARG;
  long long nb;
  if (p->variant == PFB_VARIANT_PADDED) nb = n_dat / p->S;
  else if (p->variant == PFB_VARIANT_LOWCBF) nb = (n_dat + (first_call ? 1536 : 0) - p->fl) / p->S;
  else nb = (n_dat - p->fl) / p->S;
  if (n_dat + (first_call ? 1536 : 0) < p->fl && p->variant != PFB_VARIANT_PADDED) nb = 0;
  return nb < 0 ? 0 : nb;
}

int pfb_analysis_nchan_out(const pfb_analysis_plan* p) { return p ? p->nchan_out : PFB_ERR_ARG; }

template <typename T> static int analysis_launch_generic(AnalysisArgs a, long long b0, long long b1, cudaStream_t st,
                                                          int nstreams = 0) {
  if (b1 <= b0) return PFB_OK;
  a.start0 += b0 * (long long)a.S;
  a.kout0 += b0;
  a.nblocks = b1 - b0;
  unsigned gy = 1;
  if (nstreams > 0) {               // batched: tile the streams over blockIdx.y
    int tile = 16;
    while (tile > 1 && (nstreams % tile || (size_t)tile * padded_len<T>(a.N) * sizeof(Cx<T>) > 96 * 1024)) tile /= 2;
    a.npol = tile;
    gy = (unsigned)(nstreams / tile);
  } else {
    a.in_pitch = a.npol;
    a.out_sk = (long long)a.nchan_out * a.npol;
    a.out_sc = a.npol;
    a.out_sp = 1;
  }
  const int ld = padded_len<T>(a.N);
  int kb = 4096 / (a.N * a.npol);
  if (kb < 1) kb = 1;
  // few-channel filterbanks (cfg1: 8 channels, 384 taps per arm): shrink the CTA's block run until the grid
  // covers the machine (the FIR loop is the cost, not the tiny FFTs)
  {
    const long long want = 148ll * 4;
    const long long cap = (a.nblocks + want - 1) / want;
    if (kb > cap) kb = (int)(cap < 1 ? 1 : cap);
  }
  a.kb = kb;
  const size_t smem = (size_t)kb * a.npol * ld * sizeof(Cx<T>);
  int rc = set_smem(pfb_analysis_generic<T>, smem);
  if (rc) return rc;
  const long long grid = (a.nblocks + kb - 1) / kb;
  if (grid > 2147483647ll || gy > 65535u) return fail(PFB_ERR_ARG, "too many blocks for one launch");
  pfb_analysis_generic<T><<<dim3((unsigned)grid, gy), 256, smem, st>>>(a);
  return check_launch("pfb_analysis_generic");
}

// phase-sharing FIR (pfb_mid.cuh): instantiated for the SKA-Mid shape (25 taps per arm, 8/7)
template <typename T, int P, int DE> static bool fir_phases_applies(int nu) {
  return std::is_same<T, float>::value && P == 25 && DE == 7 && nu == 8;
}
template <int P, int DE>
static int fir_phases_launch(int nu, const AnalysisArgs& a, void* y, long long row0, long long rows_full, long long i1,
                             cudaStream_t st) {
  return (int)cudaErrorNotSupported;
}
template <>
int fir_phases_launch<25, 7>(int nu, const AnalysisArgs& a, void* y, long long row0, long long rows_full, long long i1,
                             cudaStream_t st) {
  CUtensorMap tm;
  int rc = mid::make_in_map(a.in, a.N, rows_full, 7, &tm);
  if (rc) return rc;
  mid::FirPhasesArgs f;
  f.y = y; f.taps = reinterpret_cast<const float*>(a.taps);
  f.N = a.N; f.S = a.S; f.nu = nu; f.start0 = a.start0; f.phase0 = a.phase0; f.nblocks = a.nblocks;
  f.i0 = 0; f.i1 = i1; f.row0 = row0;
  const int gx = a.N / mid::kPhArms;
  long long nseg = (148ll * 2 * 4 + gx - 1) / gx;          // ~4 waves of two CTAs per SM
  if (nseg < 1) nseg = 1;
  long long spc = (i1 + nseg - 1) / nseg;
  if (spc < 16) spc = 16;
  nseg = (i1 + spc - 1) / spc;
  f.steps_per_cta = spc;
  constexpr size_t smem = mid::fir_phases_smem<25, 7>();
  cudaError_t e = cudaFuncSetAttribute(mid::pfb_fir_phases<25, 7, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  mid::pfb_fir_phases<25, 7, 8><<<dim3((unsigned)gx, (unsigned)nseg), 2 * mid::kPhArms * 8, smem, st>>>(f, tm);
  return (int)cudaGetLastError();
}

// sliding-window FIR + FFT (two kernels) for the (taps per arm, de) pairs that have an instantiation
template <typename T, int P, int DE>
static int slide_go(pfb_analysis_plan* p, const AnalysisArgs& a, cudaStream_t st) {
  const size_t need = (size_t)a.nblocks * a.N * a.npol * sizeof(Cx<T>);
  if (need > p->yscratch_cap) {
    if (p->d_yscratch) CU(cudaFree(p->d_yscratch));
    p->d_yscratch = nullptr; p->yscratch_cap = 0;
    CU(cudaMalloc(&p->d_yscratch, need));
    p->yscratch_cap = need;
  }
  FirSlideArgs f;
  memset(&f, 0, sizeof(f));
  f.in = a.in; f.y = p->d_yscratch; f.taps = a.taps;
  f.N = a.N; f.S = a.S; f.npol = a.npol; f.nu = p->nu;
  f.n_dat = a.n_dat; f.start0 = a.start0; f.phase0 = a.phase0; f.nblocks = a.nblocks;
  constexpr int PER = DE < P ? (P + DE - 1) / DE : 1;     // unroll period of the kernel's circular window
  const long long nsteps = (a.nblocks + p->nu - 1) / p->nu;
  const unsigned gx = (unsigned)((a.N * a.npol + 127) / 128);
  auto launch_slide = [&](long long i0, long long i1) -> int {
    if (i1 <= i0) return PFB_OK;
    const long long n = i1 - i0;
    long long nseg = (148ll * 12) / ((long long)gx * p->nu);
    if (nseg < 1) nseg = 1;
    long long spc = (n + nseg - 1) / nseg;
    spc = (spc + PER - 1) / PER * PER;
    nseg = (n + spc - 1) / spc;
    if (nseg > 65535) return fail(PFB_ERR_ARG, "too many segments");
    f.steps_per_cta = spc; f.i0 = i0; f.i1 = i1;
    pfb_fir_slide<T, P, DE><<<dim3(gx, (unsigned)p->nu, (unsigned)nseg), 128, 0, st>>>(f);
    return check_launch("pfb_fir_slide");
  };
  // fp32 dual-pol long filters whose block-0 window starts on a multiple of N samples: the phase-sharing kernel
  // for every step whose rows lie in the whole-row part of the stream; the last few steps keep the checked loads
  long long i_fast = 0;
  if (fir_phases_applies<T, P, DE>(p->nu) && a.npol == 2 && a.N % mid::kPhArms == 0 && a.start0 % a.N == 0 &&
      ((uintptr_t)a.in % 16 == 0) && p->mid_fast) {
    const long long rows_full = a.n_dat / a.N;
    const long long row0 = a.start0 / a.N;                       // exact (may be negative)
    // step i reads rows row0 + Q, Q <= (nu - 1) + DE * i + P - 1
    const long long qmax = rows_full - 1 - row0 - (p->nu - 1) - (P - 1);
    i_fast = qmax >= 0 ? qmax / DE + 1 : 0;
    if (i_fast > nsteps) i_fast = nsteps;
    if (i_fast < 8) i_fast = 0;
    if (i_fast > 0) {
      int rc2 = fir_phases_launch<P, DE>(p->nu, a, p->d_yscratch, row0, rows_full, i_fast, st);
      if (rc2) return fail(PFB_ERR_CUDA, "pfb_fir_phases launch: %s", cudaGetErrorString((cudaError_t)rc2));
      rc2 = check_launch("pfb_fir_phases");
      if (rc2) return rc2;
    }
  }
  int rc = launch_slide(i_fast, nsteps);
  if (rc) return rc;
  if (std::is_same<T, float>::value && a.N == 4096 && a.npol == 2 && ((uintptr_t)a.out % 16 == 0) && p->mid_fast) {
    // 4096-channel fast path: one register-resident 16x16x16 transform per (block, both pols)
    mid::RowFftArgs m;
    m.in = p->d_yscratch; m.out = a.out; m.tw = reinterpret_cast<const float2*>(a.fft.tw); m.jmap = nullptr;
    m.nrows = a.nblocks; m.kout0 = a.kout0; m.kout_mod = a.kout_mod;
    m.chan_first = a.chan_first; m.nchan_out = a.nchan_out; m.scale = (float)a.scale;
    rc = mid::rowfft4096_launch<-1, true>(m, st);
    if (rc) return fail(PFB_ERR_CUDA, "pfb_rowfft4096 launch: %s", cudaGetErrorString((cudaError_t)rc));
    return check_launch("pfb_rowfft4096");
  }
  AnalysisArgs b = a;
  b.ysrc = p->d_yscratch;
  return analysis_launch_generic<T>(b, 0, a.nblocks, st);
}

template <typename T>
static int analysis_launch_slide(pfb_analysis_plan* p, const AnalysisArgs& a, cudaStream_t st, bool* done) {
  *done = true;
  if (a.nblocks < 4 * p->nu) { *done = false; return PFB_OK; }
  const int P = p->P, de = p->de;
  if ((long long)p->nu * p->S != (long long)de * p->N) { *done = false; return PFB_OK; }   // needs N divisible by nu
  if (P == 25 && de == 7) return slide_go<T, 25, 7>(p, a, st);       // SKA-Mid: 4096 ch, 8/7, 100 353 taps
  if (P == 12 && de == 3) return slide_go<T, 12, 3>(p, a, st);       // SKA-Low shape outside the fast kernel (fp64, misaligned)
  if (P == 13 && de == 3) return slide_go<T, 13, 3>(p, a, st);       // 3073-tap 'low' design
  if (P == 24 && de == 27) return slide_go<T, 24, 27>(p, a, st);     // 'sps': 256 ch, 32/27, 6144 taps
  *done = false;
  return PFB_OK;
}

// 8-channel, 384-taps-per-arm shape: register-sliding FIR with CTA-uniform taps, then the generic FFT stage
template <typename T> static int analysis_launch_small8(pfb_analysis_plan* p, const AnalysisArgs& a, cudaStream_t st, bool* done) {
  *done = false;
  return PFB_OK;
}
template <> int analysis_launch_small8<float>(pfb_analysis_plan* p, const AnalysisArgs& a, cudaStream_t st, bool* done) {
  *done = false;
  if (!p->fir8 || a.npol != 1 || a.in2 || a.start0 < 0 || a.kout_mod != 0 || a.nblocks < 64 || ((uintptr_t)a.in % 8)) return PFB_OK;
  const size_t need = (size_t)a.nblocks * 8 * sizeof(Cx<float>);
  if (need > p->yscratch_cap) {
    if (p->d_yscratch) CU(cudaFree(p->d_yscratch));
    p->d_yscratch = nullptr; p->yscratch_cap = 0;
    CU(cudaMalloc(&p->d_yscratch, need));
    p->yscratch_cap = need;
  }
  small8::Fir8Args f;
  f.in = reinterpret_cast<const small8::C*>(a.in);
  f.y = reinterpret_cast<small8::C*>(p->d_yscratch);
  f.n_dat = a.n_dat; f.nblocks = a.nblocks; f.start0 = a.start0;
  f.rot = (int)((((a.phase0 + a.start0) % 8) + 8) % 8);
  int rc;
  if (p->d_fir8_btab) {
    small8::Fir8MmaArgs g;
    g.in = f.in; g.y = f.y; g.btab = p->d_fir8_btab; g.n_dat = f.n_dat; g.nblocks = f.nblocks; g.start0 = f.start0; g.rot = f.rot;
    rc = small8::fir8_mma_launch(g, st);
    if (rc) return fail(PFB_ERR_CUDA, "pfb_fir8_mma launch: %s", cudaGetErrorString((cudaError_t)rc));
    rc = check_launch("pfb_fir8_mma");
  } else {
    rc = small8::fir8_launch(f, *p->fir8, st);
    if (rc) return fail(PFB_ERR_CUDA, "pfb_fir8_slide launch: %s", cudaGetErrorString((cudaError_t)rc));
    rc = check_launch("pfb_fir8_slide");
  }
  if (rc) return rc;
  *done = true;
  if (a.chan_first == 0 && a.nchan_out == 8 && ((uintptr_t)a.out % 16 == 0)) {
    small8::pfb_fft8_store<<<(unsigned)((a.nblocks + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const small8::C*>(p->d_yscratch), reinterpret_cast<small8::C*>(a.out), a.nblocks, a.kout0, (float)a.scale);
    return check_launch("pfb_fft8_store");
  }
  AnalysisArgs b = a;
  b.ysrc = p->d_yscratch;
  return analysis_launch_generic<float>(b, 0, a.nblocks, st);
}

template <typename T> static int analysis_launch(pfb_analysis_plan* p, AnalysisArgs& a, cudaStream_t st) {
  if (a.nblocks <= 0) return PFB_OK;
  {
    bool done = false;
    int rc = analysis_launch_small8<T>(p, a, st, &done);
    if (rc || done) return rc;
  }
  long long k_lo = 0, k_hi = 0;
  if (p->low_fast) {
    int rc = low_analysis_launch<T>(a, p->d_low_taps, st, &k_lo, &k_hi);
    if (rc) return fail(PFB_ERR_CUDA, "pfb_analysis_low launch: %s", cudaGetErrorString((cudaError_t)rc));
    if (k_hi > k_lo) {
      rc = check_launch("pfb_analysis_low");
      if (rc) return rc;
    }
  }
  if (a.in2) {
    // peer-halo form: only the fast kernel reads two buffers; the generic kernel may serve head blocks whose
    // windows lie inside the own buffer, nothing else
    if (k_hi <= k_lo || k_hi < a.nblocks)
      return fail(PFB_ERR_UNSUPPORTED, "peer-halo analysis needs the SKA-Low fast path over all blocks of the segment");
    if (k_lo > 0 && a.start0 + (k_lo - 1) * (long long)a.S + (long long)a.P * a.N > a.n_own)
      return fail(PFB_ERR_UNSUPPORTED, "peer-halo analysis: head blocks reach into the peer buffer");
    AnalysisArgs h = a;
    h.in2 = nullptr; h.n_dat = a.n_own;
    return analysis_launch_generic<T>(h, 0, k_lo, st);
  }
  if (k_hi <= k_lo) {
    bool done = false;
    int rc = analysis_launch_slide<T>(p, a, st, &done);
    if (rc || done) return rc;
    return analysis_launch_generic<T>(a, 0, a.nblocks, st);
  }
  // head / tail blocks whose window leaves the resident samples (zero fill) stay on the generic kernel
  int rc = analysis_launch_generic<T>(a, 0, k_lo, st);
  if (rc) return rc;
  return analysis_launch_generic<T>(a, k_hi, a.nblocks, st);
}

static int analysis_run(pfb_analysis_plan* p, const void* in, long long n_local, long long sample0, void* out,
                        int npol, long long k_first, long long nblk, long long kout0, long long kout_mod,
                        int first_call, void* stream, const void* in2 = nullptr, long long n_own = 0) {
  if (!p) return fail(PFB_ERR_ARG, "null plan");
  if (npol < 1 || npol > 2) return fail(PFB_ERR_ARG, "npol must be 1 or 2");
  if (nblk <= 0) return PFB_OK;                      // nothing to do (empty / too short input)
  if (!in || !out) return fail(PFB_ERR_ARG, "null pointer");
  CU(cudaSetDevice(p->device));
  AnalysisArgs a;
  memset(&a, 0, sizeof(a));
  a.in = in; a.out = out; a.taps = p->d_taps; a.fft = p->fft;
  a.N = p->N; a.S = p->S; a.P = p->P; a.npol = npol;
  a.n_dat = n_local;
  a.in2 = in2; a.n_own = in2 ? n_own : n_local;
  long long start_abs0 = 0, rot = 0;
  if (p->variant == PFB_VARIANT_PADDED) { start_abs0 = -(long long)p->fl; rot = 1; }
  if (p->variant == PFB_VARIANT_LOWCBF) { start_abs0 = first_call ? -1536 : 0; rot = first_call ? 1536 : 0; }
  a.start0 = start_abs0 + k_first * (long long)p->S - sample0;
  a.phase0 = sample0 + rot;
  a.nblocks = nblk; a.kout0 = kout0; a.kout_mod = kout_mod;
  a.chan_first = p->chan_first; a.nchan_out = p->nchan_out; a.scale = p->scale;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return p->dtype == PFB_DTYPE_C64 ? analysis_launch<float>(p, a, st) : analysis_launch<double>(p, a, st);
}

int pfb_analysis_execute(pfb_analysis_plan* plan, const void* in_dev, void* out_dev, int npol, long long n_dat,
                         int first_call, void* stream) {
  if (!plan) return fail(PFB_ERR_ARG, "null plan");
  const long long nb = pfb_analysis_nblocks(plan, n_dat, first_call);
  long long kout0 = 0, kmod = 0;
  if (plan->variant == PFB_VARIANT_PADDED && nb > 0) {      // circshift(out,[0,0,-shift]) :156
    kmod = nb;
    kout0 = (nb - (plan->sds % nb)) % nb;
  }
  return analysis_run(plan, in_dev, n_dat, 0, out_dev, npol, 0, nb, kout0, kmod, first_call, stream);
}

int pfb_analysis_execute_batched(pfb_analysis_plan* p, const void* in_dev, void* out_dev, int nstreams, long long n_dat,
                                 int first_call, int out_layout, void* stream) {
  if (!p) return fail(PFB_ERR_ARG, "null plan");
  if (nstreams < 1 || (out_layout != 0 && out_layout != 1)) return fail(PFB_ERR_ARG, "pfb_analysis_execute_batched: bad argument");
  const long long nb = pfb_analysis_nblocks(p, n_dat, first_call);
  if (nb <= 0) return PFB_OK;
  if (!in_dev || !out_dev) return fail(PFB_ERR_ARG, "null pointer");
  CU(cudaSetDevice(p->device));
  AnalysisArgs a;
  memset(&a, 0, sizeof(a));
  a.in = in_dev; a.out = out_dev; a.taps = p->d_taps; a.fft = p->fft;
  a.N = p->N; a.S = p->S; a.P = p->P; a.n_dat = n_dat;
  long long start_abs0 = 0, rot = 0;
  if (p->variant == PFB_VARIANT_PADDED) { start_abs0 = -(long long)p->fl; rot = 1; }
  if (p->variant == PFB_VARIANT_LOWCBF) { start_abs0 = first_call ? -1536 : 0; rot = first_call ? 1536 : 0; }
  a.start0 = start_abs0; a.phase0 = rot; a.nblocks = nb;
  if (p->variant == PFB_VARIANT_PADDED) { a.kout_mod = nb; a.kout0 = (nb - (p->sds % nb)) % nb; }
  a.chan_first = p->chan_first; a.nchan_out = p->nchan_out; a.scale = p->scale;
  a.in_pitch = nstreams;
  if (out_layout == 0) { a.out_sk = (long long)p->nchan_out * nstreams; a.out_sc = nstreams; a.out_sp = 1; }       // [k][chan][stream]
  else { a.out_sk = (long long)p->nchan_out * nstreams; a.out_sc = 1; a.out_sp = p->nchan_out; }                    // [k][stream][chan]
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return p->dtype == PFB_DTYPE_C64 ? analysis_launch_generic<float>(a, 0, nb, st, nstreams)
                                   : analysis_launch_generic<double>(a, 0, nb, st, nstreams);
}

int pfb_analysis_execute_segment(pfb_analysis_plan* plan, const void* in_dev, long long n_local, long long sample0,
                                 void* out_dev, int npol, long long k_first, long long nblk, int first_call,
                                 void* stream) {
  if (!plan) return fail(PFB_ERR_ARG, "null plan");
  if (plan->variant == PFB_VARIANT_PADDED)
    return fail(PFB_ERR_UNSUPPORTED, "padded variant wraps its output circularly; segment form not defined");
  if (k_first < 0 || nblk < 0) return fail(PFB_ERR_ARG, "negative block range");
  return analysis_run(plan, in_dev, n_local, sample0, out_dev, npol, k_first, nblk, 0, 0, first_call, stream);
}

int pfb_analysis_execute_segment_peer(pfb_analysis_plan* plan, const void* in_dev, long long n_own, const void* next_dev,
                                      long long n_next, long long sample0, void* out_dev, int npol, long long k_first,
                                      long long nblk, int first_call, void* stream) {
  if (!plan) return fail(PFB_ERR_ARG, "null plan");
  if (plan->variant == PFB_VARIANT_PADDED)
    return fail(PFB_ERR_UNSUPPORTED, "padded variant wraps its output circularly; segment form not defined");
  if (k_first < 0 || nblk < 0 || n_own < 0 || n_next < 0) return fail(PFB_ERR_ARG, "negative size");
  if (n_next > 0 && !next_dev) return fail(PFB_ERR_ARG, "null peer pointer");
  if (n_next == 0) return analysis_run(plan, in_dev, n_own, sample0, out_dev, npol, k_first, nblk, 0, 0, first_call, stream);
  return analysis_run(plan, in_dev, n_own + n_next, sample0, out_dev, npol, k_first, nblk, 0, 0, first_call, stream,
                      next_dev, n_own);
}

// ---------------------------------------------------------------------------
static void freqz_half_host(const double* b, int nb, long long npoints, int neval, std::vector<double>* mag) {
  mag->resize(neval);
  for (int k = 0; k < neval; ++k) {
    long double re = 0, im = 0;
    for (int m = 0; m < nb; ++m) {
      double c, s;
      unit((long long)k * m, 2 * npoints, -1, &c, &s);   // exp(-i pi k m / npoints)
      re += (long double)b[m] * c;
      im += (long double)b[m] * s;
    }
    (*mag)[k] = (double)sqrtl(re * re + im * im);
  }
}

template <typename T>
static int inverse_build(pfb_inverse_plan* p, const double* taper, const double* dtaps, int ndtaps) {
  int rc;
  if ((rc = build_fft<T>(p->nchan, &p->fft_chan, &p->tabs))) return rc;
  if ((rc = build_fft<T>(p->Nf, &p->fft_fwd, &p->tabs))) return rc;
  if ((rc = build_fft<T>(p->FNw, &p->fft_bwd, &p->tabs))) return rc;
  std::vector<T> tp(p->Nf);
  for (int i = 0; i < p->Nf; ++i) tp[i] = taper ? (T)taper[i] : (T)1;
  const T* dtp;
  if ((rc = p->tabs.upload(tp, &dtp))) return rc;
  p->d_taper = dtp;
  p->taper_unity_mid = true;
  for (int i = p->ov; i < p->Nf - p->ov; ++i)
    if (tp[(size_t)i] != (T)1) p->taper_unity_mid = false;
  // deripple weights in FN order (polyphase_synthesis.m:142-149, 244-250)
  const int pb = p->FNw / 2;
  std::vector<double> w(p->FNw, 1.0);
  if (ndtaps > 0) {
    std::vector<double> mag;
    freqz_half_host(dtaps, ndtaps, (long long)p->nchan * pb, pb + 1, &mag);
    for (int i = 0; i < pb; ++i) {
      w[i] = 1.0 / mag[pb - i];
      w[pb + i] = 1.0 / mag[i];
    }
  }
  // W[t1][r] = D[r] (de/nu)/oN exp(+2 pi i (r-off) t1 / oN)
  const long long oN = (long long)p->nchan * p->FNw;
  const double g = ((double)p->de / p->nu) / (double)oN;
  std::vector<Cx<T>> wt((size_t)p->nchan * p->FNw);
  for (int t1 = 0; t1 < p->nchan; ++t1)
    for (int r = 0; r < p->FNw; ++r) {
      double c, s;
      unit((long long)(r - p->off) * t1, oN, +1, &c, &s);
      wt[(size_t)t1 * p->FNw + r].x = (T)(w[r] * g * c);
      wt[(size_t)t1 * p->FNw + r].y = (T)(w[r] * g * s);
    }
  const Cx<T>* dwt;
  if ((rc = p->tabs.upload(wt, &dwt))) return rc;
  p->d_wt = dwt;
  p->deripple_w = w;
  if (p->combine > 1) {                                       // polyphase_synthesis.m:198-238
    std::vector<int> jm(p->nchan);
    const int per = p->nchan / p->combine;
    for (int chan = 0; chan < p->nchan; ++chan) {
      int fine = (chan + per / 2) % p->nchan;
      int coarse = fine / per;
      fine -= coarse * per;
      coarse = (coarse + p->combine / 2) % p->combine;
      fine = (fine + per / 2) % per;
      jm[chan] = coarse * per + fine;
    }
    const int* djm;
    if ((rc = p->tabs.upload(jm, &djm))) return rc;
    p->d_jmap = djm;
  }
  rc = low_inverse_plan_init<T>(p->nchan, p->nu, p->de, p->Nf, p->ov, p->spans, p->combine, tp, p->d_wt,
                                &p->tabs.ptrs, &p->d_low, &p->low_fast);
  if (rc > 0) return fail(PFB_ERR_CUDA, "fast-path table upload: %s", cudaGetErrorString((cudaError_t)rc));
  return rc;
}

int pfb_inverse_plan_create(pfb_inverse_plan** plan, int nchan, int os_nu, int os_de, int nfft, int overlap,
                            const double* temporal_taper, const double* deripple_taps, int n_deripple_taps,
                            int spans_nyquist, int combine, int dtype, int device) {
  if (!plan || nchan < 1 || os_nu < 1 || os_de < 1 || os_de > os_nu || nfft < 2 || overlap < 0 || combine < 1)
    return fail(PFB_ERR_ARG, "pfb_inverse_plan_create: bad argument");
  if (dtype != PFB_DTYPE_C64 && dtype != PFB_DTYPE_C128) return fail(PFB_ERR_ARG, "dtype must be 0 (c64) or 1 (c128)");
  if (((long long)nfft * os_de) % os_nu || ((long long)overlap * os_de) % os_nu)
    return fail(PFB_ERR_ARG, "input_fft_length and input_overlap must be divisible by os_factor.nu");
  if (nfft - 2 * overlap < 1) return fail(PFB_ERR_ARG, "input_overlap too large");
  if (nchan % combine) return fail(PFB_ERR_ARG, "nchan must be a multiple of combine");
  if (n_deripple_taps > 0 && !deripple_taps) return fail(PFB_ERR_ARG, "deripple taps missing");
  CU(cudaSetDevice(device));
  pfb_inverse_plan* p = new pfb_inverse_plan();
  p->nchan = nchan; p->nu = os_nu; p->de = os_de; p->Nf = nfft; p->ov = overlap; p->dtype = dtype; p->device = device;
  p->keep = nfft - 2 * overlap;
  p->mid_fast = getenv("PFB_NO_MID") == nullptr;
  p->ws_fast = getenv("PFB_NO_WS") == nullptr;
  p->FNw = (int)(((long long)nfft * os_de) / os_nu);          // polyphase_synthesis.m:133
  p->ov2 = (int)(((long long)overlap * os_de) / os_nu);
  p->keep2 = p->FNw - 2 * p->ov2;
  p->d2 = (nfft - p->FNw) / 2;                                 // :136
  p->spans = spans_nyquist ? 1 : 0;
  p->combine = combine;
  p->off = p->spans ? p->FNw / 2 : 0;
  if (p->FNw % 2 || (nfft - p->FNw) % 2) { delete p; return fail(PFB_ERR_ARG, "FN_width must be even"); }
  int rc = dtype == PFB_DTYPE_C64 ? inverse_build<float>(p, temporal_taper, deripple_taps, n_deripple_taps)
                                  : inverse_build<double>(p, temporal_taper, deripple_taps, n_deripple_taps);
  if (rc) { delete p; return rc; }
  *plan = p;
  return PFB_OK;
}

int pfb_inverse_plan_destroy(pfb_inverse_plan* plan) {
  if (plan) { cudaSetDevice(plan->device); delete plan; }
  return PFB_OK;
}

long long pfb_inverse_nblocks(const pfb_inverse_plan* p, long long n_dat) {
  if (!p) return PFB_ERR_ARG;
  const long long nb = (n_dat - 2 * p->ov) / p->keep;          // polyphase_synthesis.m:114
  return (n_dat < 2 * p->ov || nb < 0) ? 0 : nb;
}

long long pfb_inverse_out_len(const pfb_inverse_plan* p, long long n_dat) {
  if (!p) return PFB_ERR_ARG;
  return pfb_inverse_nblocks(p, n_dat) * (long long)p->keep2 * p->nchan;
}

int pfb_inverse_set_scratch_bytes(pfb_inverse_plan* p, long long bytes) {
  if (!p || bytes < 0) return fail(PFB_ERR_ARG, "bad scratch size");
  p->scratch_bytes = bytes;
  return PFB_OK;
}

template <typename T>
static int launch_chan_idft(pfb_inverse_plan* p, const void* in, void* out, int npol, long long nrows, cudaStream_t st,
                            bool grouped = false) {
  if (nrows <= 0) return PFB_OK;
  const int tok = p->prof_begin(0, st);
  if (p->low_fast) {
    bool done = false;
    int rc = low_chan_idft_launch<T>(p->d_low, in, out, npol, nrows, st, &done, grouped);
    if (rc) return fail(PFB_ERR_CUDA, "pfb_chan_idft_low launch: %s", cudaGetErrorString((cudaError_t)rc));
    if (done) { p->prof_end(tok, st); return check_launch("pfb_chan_idft_low"); }
  }
  if (std::is_same<T, float>::value && p->nchan == 4096 && npol == 2 && ((uintptr_t)in % 16 == 0) &&
      ((uintptr_t)out % 16 == 0) && p->mid_fast) {
    mid::RowFftArgs m;
    memset(&m, 0, sizeof(m));
    m.in = in; m.out = out; m.tw = reinterpret_cast<const float2*>(p->fft_chan.tw); m.jmap = p->d_jmap; m.nrows = nrows;
    int rc = mid::rowfft4096_launch<+1, false>(m, st);
    if (rc) return fail(PFB_ERR_CUDA, "pfb_rowfft4096 launch: %s", cudaGetErrorString((cudaError_t)rc));
    p->prof_end(tok, st);
    return check_launch("pfb_rowfft4096");
  }
  ChanIdftArgs a;
  memset(&a, 0, sizeof(a));
  a.in = in; a.out = out; a.fft = p->fft_chan; a.jmap = p->d_jmap; a.nchan = p->nchan; a.npol = npol; a.nrows = nrows;
  int rb = 4096 / (p->nchan * npol);
  if (rb < 1) rb = 1;
  a.rb = rb;
  const size_t smem = (size_t)rb * npol * padded_len<T>(p->nchan) * sizeof(Cx<T>);
  int rc = set_smem(pfb_chan_idft_generic<T>, smem);
  if (rc) return rc;
  const long long grid = (nrows + rb - 1) / rb;
  if (grid > 2147483647ll) return fail(PFB_ERR_ARG, "too many rows for one launch");
  pfb_chan_idft_generic<T><<<(unsigned)grid, 256, smem, st>>>(a);
  p->prof_end(tok, st);
  return check_launch("pfb_chan_idft_generic");
}

template <typename T>
static int launch_resample(pfb_inverse_plan* p, const void* v, void* out, int npol, long long nblk, cudaStream_t st,
                           int G = 1, bool grouped = false, long long rows_cap = 0, long long vec_rows = 0,
                           long long vec_blocks = 0) {
  if (nblk <= 0) return PFB_OK;
  const int tok = p->prof_begin(1, st);
  if (p->low_fast) {
    bool done = false;
    int rc = low_resample_launch<T>(p->d_low, v, out, npol, nblk, G, st, &done, grouped, rows_cap, vec_rows, vec_blocks);
    if (rc) return fail(PFB_ERR_CUDA, "pfb_resample_low launch: %s", cudaGetErrorString((cudaError_t)rc));
    if (done) { p->prof_end(tok, st); return check_launch("pfb_resample_low"); }
  }
  if (std::is_same<T, float>::value && p->mid_fast && p->Nf == 512 && p->ov == 128 && p->nu == 8 && p->de == 7 && p->spans &&
      npol == 2 && G == 1 && p->nchan % 8 == 0 && ((uintptr_t)v % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
      nblk * (long long)(p->nchan / 8) < 2147483647ll) {
    // SKA-Mid inversion shape: register-resident 512 -> 448 resample (pfb_mid.cuh)
    mid::MidResArgs m;
    m.v = reinterpret_cast<const mid::C*>(v); m.out = reinterpret_cast<mid::C*>(out);
    m.taper = reinterpret_cast<const float*>(p->d_taper); m.wt = reinterpret_cast<const mid::C*>(p->d_wt);
    m.tw512 = reinterpret_cast<const float2*>(p->fft_fwd.tw); m.tw448 = reinterpret_cast<const float2*>(p->fft_bwd.tw);
    m.nchan = p->nchan; m.nblk = (int)nblk;
    int rc = mid::resample_mid_launch(m, p->taper_unity_mid, st);
    if (rc) return fail(PFB_ERR_CUDA, "pfb_resample_mid launch: %s", cudaGetErrorString((cudaError_t)rc));
    p->prof_end(tok, st);
    return check_launch("pfb_resample_mid");
  }
  ResampleArgs a;
  memset(&a, 0, sizeof(a));
  a.v = v; a.out = out; a.fwd = p->fft_fwd; a.bwd = p->fft_bwd; a.taper = p->d_taper; a.wt = p->d_wt;
  const int nchan_total = p->nchan * G;
  a.nchan = nchan_total; a.wt_mod = p->nchan; a.npol = npol; a.Nf = p->Nf; a.FNw = p->FNw; a.keep = p->keep; a.ov2 = p->ov2;
  a.keep2 = p->keep2; a.off = p->off; a.d2 = p->d2;
  const size_t percol = (size_t)(padded_len<T>(p->Nf) + padded_len<T>(p->FNw)) * sizeof(Cx<T>);
  int tg = 1;
  while (nchan_total % (tg * 2) == 0 && (size_t)(tg * 2) * npol * percol <= 110 * 1024) tg *= 2;
  a.tg = tg;
  const size_t smem = (size_t)tg * npol * percol;
  int rc = set_smem(pfb_resample_generic<T>, smem);
  if (rc) return rc;
  const long long groups = nchan_total / tg;
  // keep each launch's grid within 2^31 and the block count in an int
  long long done = 0;
  while (done < nblk) {
    long long nb = nblk - done;
    const long long maxb = 2000000000ll / groups;
    if (nb > maxb) nb = maxb;
    a.nblk = (int)nb;
    const long long rowlen = (long long)nchan_total * npol;
    a.v = reinterpret_cast<const Cx<T>*>(v) + done * p->keep * rowlen;
    a.out = reinterpret_cast<Cx<T>*>(out) + done * p->keep2 * rowlen;
    pfb_resample_generic<T><<<(unsigned)(nb * groups), 256, smem, st>>>(a);
    rc = check_launch("pfb_resample_generic");
    if (rc) return rc;
    done += nb;
  }
  p->prof_end(tok, st);
  return PFB_OK;
}

template <typename T>
static int inverse_run(pfb_inverse_plan* p, const void* in, void* out, int npol, long long n_dat, cudaStream_t st,
                       int G = 1, bool dry_run = false) {
  const long long nblocks = pfb_inverse_nblocks(p, n_dat);
  if (nblocks <= 0) return PFB_OK;
  const long long rowbytes = (long long)p->nchan * G * npol * sizeof(Cx<T>);
  long long B = (p->scratch_bytes - (long long)2 * p->ov * rowbytes) / ((long long)p->keep * rowbytes);
  if (B < 1) B = 1;
  if (p->low_fast) {                    // equal work per CTA of the one-wave resample grid
    const long long ns = low_resample_nsplit(p->d_low, npol, G);
    if (B >= ns) B = B / ns * ns;
  }
  if (B > nblocks) B = nblocks;
  const size_t need = (size_t)(((B - 1) * p->keep + p->Nf) * rowbytes);
  if (need > p->scratch_cap) {
    // (a CUDA graph captured before this call would hold the old scratch pointer: size the plan with
    //  pfb_inverse_reserve before capturing)
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs != cudaStreamCaptureStatusNone)
      return fail(PFB_ERR_ARG, "inversion scratch must grow (%zu > %zu bytes) during stream capture: call pfb_inverse_reserve first", need, p->scratch_cap);
    if (p->d_scratch) CU(cudaFree(p->d_scratch));
    p->d_scratch = nullptr; p->scratch_cap = 0;
    CU(cudaMalloc(&p->d_scratch, need));
    p->scratch_cap = need;
  }
  if (dry_run) return PFB_OK;
  const long long rowlen = (long long)p->nchan * G * npol;
  for (long long n0 = 0; n0 < nblocks; n0 += B) {
    const long long nb = (nblocks - n0 < B) ? nblocks - n0 : B;
    const long long rows = (nb - 1) * p->keep + p->Nf;
    const Cx<T>* src = reinterpret_cast<const Cx<T>*>(in) + n0 * p->keep * rowlen;
    Cx<T>* dst = reinterpret_cast<Cx<T>*>(out) + n0 * p->keep2 * rowlen;
    // SKA-Low dual-pol shape: grouped scratch layout + warp-autonomous resample kernel
    const bool ws = p->low_fast && p->ws_fast && std::is_same<T, float>::value && low_ws_ok(p->d_low, npol, G, src, dst);
    int rc = launch_chan_idft<T>(p, src, p->d_scratch, npol, rows * G, st, ws);  // every (tau, group) is one row
    if (rc) return rc;
    rc = launch_resample<T>(p, p->d_scratch, dst, npol, nb, st, G, ws, rows);
    if (rc) return rc;
  }
  return PFB_OK;
}

// A batch of nvec independent channelized series stored back to back (row pitch vec_rows >= n_dat fine samples each):
// the purity sweep (python/verify/purity.py:165-207 runs one synthesize call per test vector).  On the SKA-Low
// dual-pol fast path the whole batch is ONE channel-IDFT launch over all rows and ONE resample launch in which every
// vector is a time split of the grid; other shapes run vector by vector.
template <typename T>
static int inverse_run_vectors(pfb_inverse_plan* p, const void* in, void* out, int npol, int nvec, long long vec_rows,
                               long long n_dat, cudaStream_t st) {
  const long long nblocks = pfb_inverse_nblocks(p, n_dat);
  if (nblocks <= 0 || nvec <= 0) return PFB_OK;
  const long long rowlen = (long long)p->nchan * npol;
  const long long rowbytes = rowlen * (long long)sizeof(Cx<T>);
  const long long out_per = nblocks * p->keep2 * rowlen;                 // output elements per vector
  const bool ws = p->low_fast && p->ws_fast && std::is_same<T, float>::value && low_ws_ok(p->d_low, npol, 1, in, out) &&
                  ((vec_rows * rowbytes) % 16 == 0);
  if (!ws) {
    for (int v = 0; v < nvec; ++v) {
      int rc = inverse_run<T>(p, reinterpret_cast<const Cx<T>*>(in) + (long long)v * vec_rows * rowlen,
                              reinterpret_cast<Cx<T>*>(out) + (long long)v * out_per, npol, n_dat, st);
      if (rc) return rc;
    }
    return PFB_OK;
  }
  long long vb = p->scratch_bytes / (vec_rows * rowbytes);                // vectors per launch pair
  if (vb < 1) vb = 1;
  if (vb > nvec) vb = nvec;
  const size_t need = (size_t)(vb * vec_rows * rowbytes);
  if (need > p->scratch_cap) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs != cudaStreamCaptureStatusNone)
      return fail(PFB_ERR_ARG, "inversion scratch must grow during stream capture: run the batch once (or reserve) first");
    if (p->d_scratch) CU(cudaFree(p->d_scratch));
    p->d_scratch = nullptr; p->scratch_cap = 0;
    CU(cudaMalloc(&p->d_scratch, need));
    p->scratch_cap = need;
  }
  for (long long v0 = 0; v0 < nvec; v0 += vb) {
    const long long nv = (nvec - v0 < vb) ? nvec - v0 : vb;
    const Cx<T>* src = reinterpret_cast<const Cx<T>*>(in) + v0 * vec_rows * rowlen;
    Cx<T>* dst = reinterpret_cast<Cx<T>*>(out) + v0 * out_per;
    int rc = launch_chan_idft<T>(p, src, p->d_scratch, npol, nv * vec_rows, st, true);
    if (rc) return rc;
    rc = launch_resample<T>(p, p->d_scratch, dst, npol, nv * nblocks, st, 1, true, nv * vec_rows, vec_rows, nblocks);
    if (rc) return rc;
  }
  return PFB_OK;
}

template <typename T> static int spectral_build(pfb_inverse_plan* p, const double* taper) {
  const long long oN = (long long)p->nchan * p->FNw;
  std::vector<T> wa((size_t)oN);
  std::vector<Cx<T>> wb((size_t)oN);
  const double g = ((double)p->de / p->nu) / (double)oN;
  for (int ch = 0; ch < p->nchan; ++ch)
    for (int r = 0; r < p->FNw; ++r) {
      long long kk = ((long long)ch * p->FNw + r - p->off) % oN;          // polyphase_synthesis.m:253-278
      if (kk < 0) kk += oN;
      wa[(size_t)ch * p->FNw + r] = (T)(p->deripple_w[r] * taper[kk]);
      double c, s;
      unit((long long)(r - p->off) * ch, oN, +1, &c, &s);                  // here `ch` plays the role of t1
      wb[(size_t)ch * p->FNw + r].x = (T)(g * c);
      wb[(size_t)ch * p->FNw + r].y = (T)(g * s);
    }
  const T* dwa;
  const Cx<T>* dwb;
  int rc = p->tabs.upload(wa, &dwa);
  if (rc) return rc;
  if ((rc = p->tabs.upload(wb, &dwb))) return rc;
  p->d_wa = dwa;
  p->d_wb = dwb;
  p->spectral = true;
  return PFB_OK;
}

int pfb_inverse_set_spectral_taper(pfb_inverse_plan* p, const double* taper, long long n) {
  if (!p) return fail(PFB_ERR_ARG, "null plan");
  if (!taper) { p->spectral = false; return PFB_OK; }
  if (n != (long long)p->nchan * p->FNw) return fail(PFB_ERR_ARG, "spectral taper must have nchan*FN_width = %lld points", (long long)p->nchan * p->FNw);
  CU(cudaSetDevice(p->device));
  return p->dtype == PFB_DTYPE_C64 ? spectral_build<float>(p, taper) : spectral_build<double>(p, taper);
}

// literal-order chain (spectral taper set): L1 -> K2g -> L3, chunked over blocks
template <typename T>
static int inverse_run_literal(pfb_inverse_plan* p, const void* in, void* out, int npol, long long n_dat, cudaStream_t st) {
  const long long nblocks = pfb_inverse_nblocks(p, n_dat);
  if (nblocks <= 0) return PFB_OK;
  const long long rowlen = (long long)p->nchan * npol;
  const long long blk_bytes = (long long)p->FNw * rowlen * sizeof(Cx<T>);
  long long B = p->scratch_bytes / blk_bytes;
  if (B < 1) B = 1;
  if (B > nblocks) B = nblocks;
  const size_t need = (size_t)(B * blk_bytes);
  if (need > p->scratch_cap) {
    if (p->d_scratch) CU(cudaFree(p->d_scratch));
    p->d_scratch = nullptr; p->scratch_cap = 0;
    CU(cudaMalloc(&p->d_scratch, need));
    p->scratch_cap = need;
  }
  auto pick_tg = [&](size_t percol) {
    int tg = 1;
    while (p->nchan % (tg * 2) == 0 && (size_t)(tg * 2) * npol * percol <= 96 * 1024) tg *= 2;
    return tg;
  };
  for (long long n0 = 0; n0 < nblocks; n0 += B) {
    const long long nb = (nblocks - n0 < B) ? nblocks - n0 : B;
    {
      FwdSelArgs a;
      memset(&a, 0, sizeof(a));
      a.in = reinterpret_cast<const Cx<T>*>(in) + n0 * p->keep * rowlen;
      a.z = p->d_scratch; a.fwd = p->fft_fwd; a.taper = p->d_taper; a.wa = p->d_wa; a.jmap = p->d_jmap;
      a.nchan = p->nchan; a.npol = npol; a.Nf = p->Nf; a.FNw = p->FNw; a.keep = p->keep; a.d2 = p->d2;
      const size_t percol = (size_t)padded_len<T>(p->Nf) * sizeof(Cx<T>);
      a.tg = pick_tg(percol); a.nblk = (int)nb;
      const size_t smem = (size_t)a.tg * npol * percol;
      int rc = set_smem(pfb_fwd_select_generic<T>, smem);
      if (rc) return rc;
      pfb_fwd_select_generic<T><<<(unsigned)(nb * (p->nchan / a.tg)), 256, smem, st>>>(a);
      if ((rc = check_launch("pfb_fwd_select_generic"))) return rc;
    }
    {
      ChanIdftArgs a;
      memset(&a, 0, sizeof(a));
      a.in = p->d_scratch; a.out = p->d_scratch; a.fft = p->fft_chan; a.jmap = nullptr;
      a.nchan = p->nchan; a.npol = npol; a.nrows = nb * p->FNw;
      int rb = 4096 / (p->nchan * npol);
      if (rb < 1) rb = 1;
      a.rb = rb;
      const size_t smem = (size_t)rb * npol * padded_len<T>(p->nchan) * sizeof(Cx<T>);
      int rc = set_smem(pfb_chan_idft_generic<T>, smem);
      if (rc) return rc;
      pfb_chan_idft_generic<T><<<(unsigned)((a.nrows + rb - 1) / rb), 256, smem, st>>>(a);
      if ((rc = check_launch("pfb_chan_idft_generic"))) return rc;
    }
    {
      InvStoreArgs a;
      memset(&a, 0, sizeof(a));
      a.A = p->d_scratch;
      a.out = reinterpret_cast<Cx<T>*>(out) + n0 * p->keep2 * rowlen;
      a.bwd = p->fft_bwd; a.wb = p->d_wb;
      a.nchan = p->nchan; a.npol = npol; a.FNw = p->FNw; a.ov2 = p->ov2; a.keep2 = p->keep2; a.off = p->off;
      const size_t percol = (size_t)padded_len<T>(p->FNw) * sizeof(Cx<T>);
      a.tg = pick_tg(percol); a.nblk = (int)nb;
      const size_t smem = (size_t)a.tg * npol * percol;
      int rc = set_smem(pfb_inv_store_generic<T>, smem);
      if (rc) return rc;
      pfb_inv_store_generic<T><<<(unsigned)(nb * (p->nchan / a.tg)), 256, smem, st>>>(a);
      if ((rc = check_launch("pfb_inv_store_generic"))) return rc;
    }
  }
  return PFB_OK;
}

int pfb_inverse_reserve(pfb_inverse_plan* p, int npol, long long n_dat, int ngroups) {
  if (!p) return fail(PFB_ERR_ARG, "null plan");
  if (npol < 1 || npol > 2 || ngroups < 1) return fail(PFB_ERR_ARG, "pfb_inverse_reserve: bad argument");
  if (p->spectral) return PFB_OK;     // the literal chain sizes its scratch per call (not graph-capturable while growing)
  CU(cudaSetDevice(p->device));
  return p->dtype == PFB_DTYPE_C64 ? inverse_run<float>(p, nullptr, nullptr, npol, n_dat, nullptr, ngroups, true)
                                   : inverse_run<double>(p, nullptr, nullptr, npol, n_dat, nullptr, ngroups, true);
}

int pfb_inverse_set_profiling(pfb_inverse_plan* p, int on) {
  if (!p) return fail(PFB_ERR_ARG, "null plan");
  p->profiling = on != 0;
  p->ev_used[0].clear(); p->ev_used[1].clear(); p->ev_next = 0;
  return PFB_OK;
}

int pfb_inverse_get_profile(pfb_inverse_plan* p, double* ms, long long* launches) {
  if (!p || !ms || !launches) return fail(PFB_ERR_ARG, "null pointer");
  CU(cudaSetDevice(p->device));
  for (int s = 0; s < 2; ++s) {
    ms[s] = 0; launches[s] = (long long)p->ev_used[s].size();
    for (auto& pr : p->ev_used[s]) {
      CU(cudaEventSynchronize(p->ev_pool[pr.second]));
      float t = 0;
      CU(cudaEventElapsedTime(&t, p->ev_pool[pr.first], p->ev_pool[pr.second]));
      ms[s] += t;
    }
    p->ev_used[s].clear();
  }
  p->ev_next = 0;
  return PFB_OK;
}

int pfb_inverse_execute(pfb_inverse_plan* p, const void* in_dev, void* out_dev, int npol, long long n_dat,
                        void* stream) {
  if (!p) return fail(PFB_ERR_ARG, "null plan");
  if (npol < 1 || npol > 2) return fail(PFB_ERR_ARG, "npol must be 1 or 2");
  if (pfb_inverse_nblocks(p, n_dat) <= 0) return PFB_OK;     // nothing to do
  if (!in_dev || !out_dev) return fail(PFB_ERR_ARG, "null pointer");
  CU(cudaSetDevice(p->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p->spectral)
    return p->dtype == PFB_DTYPE_C64 ? inverse_run_literal<float>(p, in_dev, out_dev, npol, n_dat, st)
                                     : inverse_run_literal<double>(p, in_dev, out_dev, npol, n_dat, st);
  return p->dtype == PFB_DTYPE_C64 ? inverse_run<float>(p, in_dev, out_dev, npol, n_dat, st)
                                   : inverse_run<double>(p, in_dev, out_dev, npol, n_dat, st);
}

int pfb_inverse_execute_batched(pfb_inverse_plan* p, const void* in_dev, void* out_dev, int ngroups, int npol,
                                long long n_dat, void* stream) {
  if (!p) return fail(PFB_ERR_ARG, "null plan");
  if (npol < 1 || npol > 2 || ngroups < 1) return fail(PFB_ERR_ARG, "pfb_inverse_execute_batched: bad argument");
  if (p->spectral) return fail(PFB_ERR_UNSUPPORTED, "batched inversion with a spectral taper is not implemented");
  if (pfb_inverse_nblocks(p, n_dat) <= 0) return PFB_OK;
  if (!in_dev || !out_dev) return fail(PFB_ERR_ARG, "null pointer");
  CU(cudaSetDevice(p->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return p->dtype == PFB_DTYPE_C64 ? inverse_run<float>(p, in_dev, out_dev, npol, n_dat, st, ngroups)
                                   : inverse_run<double>(p, in_dev, out_dev, npol, n_dat, st, ngroups);
}

int pfb_inverse_execute_vectors(pfb_inverse_plan* p, const void* in_dev, void* out_dev, int npol, int nvec,
                                long long vec_rows, long long n_dat, void* stream) {
  if (!p) return fail(PFB_ERR_ARG, "null plan");
  if (npol < 1 || npol > 2 || nvec < 0 || n_dat < 0 || vec_rows < n_dat)
    return fail(PFB_ERR_ARG, "pfb_inverse_execute_vectors: bad argument (need vec_rows >= n_dat)");
  if (p->spectral) return fail(PFB_ERR_UNSUPPORTED, "vector batches with a spectral taper are not implemented");
  if (nvec == 0 || pfb_inverse_nblocks(p, n_dat) <= 0) return PFB_OK;
  if (!in_dev || !out_dev) return fail(PFB_ERR_ARG, "null pointer");
  CU(cudaSetDevice(p->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return p->dtype == PFB_DTYPE_C64 ? inverse_run_vectors<float>(p, in_dev, out_dev, npol, nvec, vec_rows, n_dat, st)
                                   : inverse_run_vectors<double>(p, in_dev, out_dev, npol, nvec, vec_rows, n_dat, st);
}

// ---------------------------------------------------------------------------
// Fused analysis -> inversion (one process, no channelized file in between): sum_c out[c] W^(-c t1) = N * acc_rot[t1],
// i.e. the rotated arm sums of the analysis ARE the inversion's channel-IDFT intermediate.  The analysis kernel writes
// them into the inverse plan's grouped scratch (and, optionally, the channelized block as usual) and the resample
// kernel follows; the channel IDFT kernel is not run.  SKA-Low dual-pol fast path only.
static int roundtrip_run(pfb_analysis_plan* ap, pfb_inverse_plan* ip, const void* in, void* chan, void* out, int npol,
                         long long n_total, int nvec, long long vec_samples, cudaStream_t st) {
  if (!ap->low_fast || ap->variant != PFB_VARIANT_BUNTON || ap->dtype != PFB_DTYPE_C64 || !ip->low_fast || !ip->ws_fast ||
      ip->dtype != PFB_DTYPE_C64 || ip->spectral || npol != 2 || ip->nchan != ap->N || !low_ws_ok(ip->d_low, 2, 1, in, out) ||
      ap->device != ip->device)
    return fail(PFB_ERR_UNSUPPORTED, "fused round trip: SKA-Low dual-pol complex64 fast path only");
  const long long nb_total = pfb_analysis_nblocks(ap, n_total, 1);
  const long long nb_vec = nvec > 1 ? pfb_analysis_nblocks(ap, vec_samples, 1) : nb_total;
  const long long nbi = pfb_inverse_nblocks(ip, nb_vec);
  if (nb_total <= 0 || nbi <= 0) return PFB_OK;
  if (nvec > 1 && (vec_samples % ap->S || vec_samples % ap->N))
    return fail(PFB_ERR_ARG, "fused round trip of a vector batch: the vector length must be a multiple of step and nchan");
  const long long vec_rows = nvec > 1 ? vec_samples / ap->S : nb_total;
  const long long rows_cap = nvec > 1 ? (long long)nvec * vec_rows : nb_total;
  const size_t need = (size_t)rows_cap * 256 * 16;
  if (need >= ((size_t)1 << 32))
    return fail(PFB_ERR_UNSUPPORTED, "fused round trip: the intermediate (%zu bytes) must stay below 4 GiB (32-bit offsets in the analysis kernel)", need);
  if ((long long)need > ip->scratch_bytes)
    return fail(PFB_ERR_UNSUPPORTED, "fused round trip: %zu bytes of scratch exceed the plan's cap (pfb_inverse_set_scratch_bytes)", need);
  if (need > ip->scratch_cap) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs != cudaStreamCaptureStatusNone) return fail(PFB_ERR_ARG, "fused round trip: scratch must grow during stream capture; run once first");
    if (ip->d_scratch) CU(cudaFree(ip->d_scratch));
    ip->d_scratch = nullptr; ip->scratch_cap = 0;
    CU(cudaMalloc(&ip->d_scratch, need));
    ip->scratch_cap = need;
  }
  AnalysisArgs a;
  memset(&a, 0, sizeof(a));
  a.in = in; a.out = chan; a.taps = ap->d_taps; a.fft = ap->fft;
  a.N = ap->N; a.S = ap->S; a.P = ap->P; a.npol = npol; a.n_dat = n_total; a.n_own = n_total;
  a.start0 = 0; a.phase0 = 0; a.nblocks = nb_total; a.kout0 = 0; a.kout_mod = 0;
  a.chan_first = ap->chan_first; a.nchan_out = ap->nchan_out; a.scale = ap->scale;
  a.emit_mode = chan ? 1 : 2; a.vout = ip->d_scratch; a.v_rows_cap = rows_cap;
  long long k_lo = 0, k_hi = 0;
  int rc = low_analysis_launch<float>(a, ap->d_low_taps, st, &k_lo, &k_hi);
  if (rc) return fail(PFB_ERR_CUDA, "pfb_analysis_low (fused) launch: %s", cudaGetErrorString((cudaError_t)rc));
  if (k_lo != 0 || k_hi != nb_total)
    return fail(PFB_ERR_UNSUPPORTED, "fused round trip: the fast analysis kernel does not cover the whole stream (alignment / length)");
  rc = check_launch("pfb_analysis_low");
  if (rc) return rc;
  return launch_resample<float>(ip, ip->d_scratch, out, npol, (long long)nvec * nbi, st, 1, true, rows_cap,
                                nvec > 1 ? vec_rows : 0, nvec > 1 ? nbi : 0);
}

int pfb_roundtrip_execute(pfb_analysis_plan* ap, pfb_inverse_plan* ip, const void* in_dev, void* chan_dev, void* out_dev,
                          int npol, long long n_dat, void* stream) {
  if (!ap || !ip) return fail(PFB_ERR_ARG, "null plan");
  if (!in_dev || !out_dev || n_dat < 0) return fail(PFB_ERR_ARG, "pfb_roundtrip_execute: bad argument");
  CU(cudaSetDevice(ap->device));
  return roundtrip_run(ap, ip, in_dev, chan_dev, out_dev, npol, n_dat, 1, n_dat, reinterpret_cast<cudaStream_t>(stream));
}

int pfb_roundtrip_execute_vectors(pfb_analysis_plan* ap, pfb_inverse_plan* ip, const void* in_dev, void* out_dev, int npol,
                                  int nvec, long long vec_samples, void* stream) {
  if (!ap || !ip) return fail(PFB_ERR_ARG, "null plan");
  if (!in_dev || !out_dev || nvec < 1 || vec_samples < 1) return fail(PFB_ERR_ARG, "pfb_roundtrip_execute_vectors: bad argument");
  CU(cudaSetDevice(ap->device));
  return roundtrip_run(ap, ip, in_dev, nullptr, out_dev, npol, (long long)nvec * vec_samples, nvec, vec_samples,
                       reinterpret_cast<cudaStream_t>(stream));
}

int pfb_inverse_execute_stage(pfb_inverse_plan* p, int stage, const void* in_dev, void* out_dev, int npol,
                              long long n_dat, void* stream) {
  if (!p || !in_dev || !out_dev) return fail(PFB_ERR_ARG, "null pointer");
  if (npol < 1 || npol > 2) return fail(PFB_ERR_ARG, "npol must be 1 or 2");
  CU(cudaSetDevice(p->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool f32 = p->dtype == PFB_DTYPE_C64;
  if (stage == 0) return f32 ? launch_chan_idft<float>(p, in_dev, out_dev, npol, n_dat, st)
                             : launch_chan_idft<double>(p, in_dev, out_dev, npol, n_dat, st);
  if (stage == 1) {
    const long long nb = pfb_inverse_nblocks(p, n_dat);
    return f32 ? launch_resample<float>(p, in_dev, out_dev, npol, nb, st)
               : launch_resample<double>(p, in_dev, out_dev, npol, nb, st);
  }
  return fail(PFB_ERR_ARG, "stage must be 0 or 1");
}

// ---------------------------------------------------------------------------
// test vectors + purity reductions on device
static int generate_run(int kind, void* out_dev, long long n, int npol, long long stride, int dtype, double p0, double p1,
                        double p2, void* stream) {
  if (!out_dev || n < 0 || npol < 1 || stride < npol || kind < 0 || kind > 2)
    return fail(PFB_ERR_ARG, "pfb_generate_test_vector: bad argument");
  if (n == 0) return PFB_OK;
  GenArgs a;
  a.out = out_dev; a.n = n; a.npol = npol; a.stride = stride; a.kind = kind; a.p0 = p0; a.p1 = p1; a.p2 = p2;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long grid = (n + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  if (dtype == PFB_DTYPE_C64) pfb_generate_kernel<float><<<(unsigned)grid, 256, 0, st>>>(a);
  else if (dtype == PFB_DTYPE_C128) pfb_generate_kernel<double><<<(unsigned)grid, 256, 0, st>>>(a);
  else return fail(PFB_ERR_ARG, "dtype must be 0 (c64) or 1 (c128)");
  return check_launch("pfb_generate_kernel");
}

int pfb_generate_test_vector(int kind, void* out_dev, long long n, int npol, int dtype, double p0, double p1, double p2,
                             void* stream) {
  return generate_run(kind, out_dev, n, npol, npol, dtype, p0, p1, p2, stream);
}

int pfb_generate_test_vector_strided(int kind, void* out_dev, long long n, long long stride, int dtype, double p0,
                                     double p1, double p2, void* stream) {
  return generate_run(kind, out_dev, n, 1, stride, dtype, p0, p1, p2, stream);
}

int pfb_generate_test_vectors(int kind, void* out_dev, long long n, int nvec, int group, long long group_stride, int dtype,
                              const double* params_dev, void* stream) {
  if (!out_dev || !params_dev || n < 0 || nvec < 0 || group < 1 || kind < 0 || kind > 2)
    return fail(PFB_ERR_ARG, "pfb_generate_test_vectors: bad argument");
  if (n == 0 || nvec == 0) return PFB_OK;
  if (nvec > 65535) return fail(PFB_ERR_ARG, "pfb_generate_test_vectors: at most 65535 vectors per call");
  GenBatchArgs a;
  a.out = out_dev; a.n = n; a.group = group; a.group_stride = group_stride; a.kind = kind; a.params = params_dev;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long gx = kind == 0 ? (n + 255) / 256 : (n + kGenSpan - 1) / kGenSpan;   // tones / combs: one CTA per 65 536-sample span
  const long long cap = nvec >= 148 ? 8 : 148 * 4;
  if (gx > cap) gx = cap;
  if (dtype == PFB_DTYPE_C64 && group == 2 && nvec % 2 == 0 && group_stride % 2 == 0 && (uintptr_t)out_dev % 16 == 0) {
    // both vectors of a [n][2] series from one CTA column: whole 16-byte elements
    long long gp = kind == 0 ? (n + 255) / 256 : (n + kPairSpan - 1) / kPairSpan;
    const long long capp = nvec >= 296 ? 8 : 148 * 8;
    if (gp > capp) gp = capp;
    pfb_generate_pairs_kernel<<<dim3((unsigned)gp, (unsigned)(nvec / 2)), 256, 0, st>>>(a);
    return check_launch("pfb_generate_pairs_kernel");
  }
  const dim3 grid((unsigned)gx, (unsigned)nvec);
  if (dtype == PFB_DTYPE_C64) pfb_generate_batch_kernel<float><<<grid, 256, 0, st>>>(a);
  else if (dtype == PFB_DTYPE_C128) pfb_generate_batch_kernel<double><<<grid, 256, 0, st>>>(a);
  else return fail(PFB_ERR_ARG, "dtype must be 0 (c64) or 1 (c128)");
  return check_launch("pfb_generate_batch_kernel");
}

// 98 304-point spectra of npairs dual-pol series (four-step: pfb_spectrum24_step1 + pfb_rowfft4096)
int pfb_spectrum_vectors(const void* in_dev, long long pitch, void* scratch_dev, void* out_dev, int npairs, long long nfft,
                         void* stream) {
  if (npairs < 0 || pitch < 0) return fail(PFB_ERR_ARG, "pfb_spectrum_vectors: bad argument");
  if (nfft != 24 * 4096) return fail(PFB_ERR_UNSUPPORTED, "pfb_spectrum_vectors: only nfft = 98304 = 24 x 4096 is implemented (got %lld)", nfft);
  if (npairs == 0) return PFB_OK;
  if (!in_dev || !scratch_dev || !out_dev) return fail(PFB_ERR_ARG, "null pointer");
  if ((uintptr_t)in_dev % 16 || (uintptr_t)scratch_dev % 16 || (uintptr_t)out_dev % 16)
    return fail(PFB_ERR_ARG, "pfb_spectrum_vectors: buffers must be 16-byte aligned");
  if (npairs > 65535) return fail(PFB_ERR_ARG, "pfb_spectrum_vectors: at most 65535 pairs per call");
  // 4096-point twiddle table, one per device (tiny, lives for the process)
  static const float2* tw_dev[64] = {nullptr};
  int dev = 0;
  CU(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(PFB_ERR_ARG, "device index out of range");
  if (!tw_dev[dev]) {
    std::vector<float2> tw = low::unit_table(4096, -1);
    void* d = nullptr;
    CU(cudaMalloc(&d, tw.size() * sizeof(float2)));
    CU(cudaMemcpy(d, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    tw_dev[dev] = reinterpret_cast<const float2*>(d);
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  mid::SpecArgs sa;
  sa.in = reinterpret_cast<const float4*>(in_dev);
  sa.s1 = reinterpret_cast<float4*>(scratch_dev);
  sa.pitch = pitch;
  mid::pfb_spectrum24_step1<<<dim3(16, (unsigned)npairs), 256, 0, st>>>(sa);
  int rc = check_launch("pfb_spectrum24_step1");
  if (rc) return rc;
  mid::RowFftArgs m;
  memset(&m, 0, sizeof(m));
  m.in = scratch_dev; m.out = out_dev; m.tw = tw_dev[dev]; m.jmap = nullptr; m.nrows = 24ll * npairs;
  rc = mid::rowfft4096_launch<-1, false>(m, st);
  if (rc) return fail(PFB_ERR_CUDA, "pfb_rowfft4096 launch: %s", cudaGetErrorString((cudaError_t)rc));
  return check_launch("pfb_rowfft4096");
}

int pfb_purity_metrics(const void* args_dev, int nvec, int dtype, void* stream) {
  if (!args_dev || nvec < 0) return fail(PFB_ERR_ARG, "pfb_purity_metrics: bad argument");
  if (nvec == 0) return PFB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const MetricsArgs* a = reinterpret_cast<const MetricsArgs*>(args_dev);
  if (dtype == PFB_DTYPE_C64) pfb_metrics_kernel<float><<<nvec, 1024, 0, st>>>(a);
  else if (dtype == PFB_DTYPE_C128) pfb_metrics_kernel<double><<<nvec, 1024, 0, st>>>(a);
  else return fail(PFB_ERR_ARG, "dtype must be 0 (c64) or 1 (c128)");
  return check_launch("pfb_metrics_kernel");
}

// ---------------------------------------------------------------------------
// integer DADA payloads (NBIT 8 / 16) <-> the arithmetic type, on the device
int pfb_unpack_payload(const void* in_dev, void* out_dev, long long nvalues, int nbit, int dtype, void* stream) {
  if (nvalues < 0 || (nbit != 8 && nbit != 16)) return fail(PFB_ERR_ARG, "pfb_unpack_payload: nbit must be 8 or 16");
  if (dtype != PFB_DTYPE_C64 && dtype != PFB_DTYPE_C128) return fail(PFB_ERR_ARG, "dtype must be 0 (c64) or 1 (c128)");
  if (nvalues == 0) return PFB_OK;
  if (!in_dev || !out_dev) return fail(PFB_ERR_ARG, "null pointer");
  if ((uintptr_t)in_dev % 16 || (uintptr_t)out_dev % 16) return fail(PFB_ERR_ARG, "pfb_unpack_payload: buffers must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long grid = (nvalues / 8 + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  if (nbit == 8 && dtype == PFB_DTYPE_C64) pfb_unpack_kernel<signed char, float><<<(unsigned)grid, 256, 0, st>>>((const signed char*)in_dev, (float*)out_dev, nvalues);
  else if (nbit == 8) pfb_unpack_kernel<signed char, double><<<(unsigned)grid, 256, 0, st>>>((const signed char*)in_dev, (double*)out_dev, nvalues);
  else if (dtype == PFB_DTYPE_C64) pfb_unpack_kernel<short, float><<<(unsigned)grid, 256, 0, st>>>((const short*)in_dev, (float*)out_dev, nvalues);
  else pfb_unpack_kernel<short, double><<<(unsigned)grid, 256, 0, st>>>((const short*)in_dev, (double*)out_dev, nvalues);
  return check_launch("pfb_unpack_kernel");
}

int pfb_pack_payload(const void* in_dev, void* out_dev, long long nvalues, int nbit, int dtype, double scale, void* stream) {
  if (nvalues < 0 || (nbit != 8 && nbit != 16)) return fail(PFB_ERR_ARG, "pfb_pack_payload: nbit must be 8 or 16");
  if (dtype != PFB_DTYPE_C64 && dtype != PFB_DTYPE_C128) return fail(PFB_ERR_ARG, "dtype must be 0 (c64) or 1 (c128)");
  if (nvalues == 0) return PFB_OK;
  if (!in_dev || !out_dev) return fail(PFB_ERR_ARG, "null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long grid = (nvalues + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  if (nbit == 8 && dtype == PFB_DTYPE_C64) pfb_pack_kernel<float, signed char><<<(unsigned)grid, 256, 0, st>>>((const float*)in_dev, (signed char*)out_dev, nvalues, (float)scale);
  else if (nbit == 8) pfb_pack_kernel<double, signed char><<<(unsigned)grid, 256, 0, st>>>((const double*)in_dev, (signed char*)out_dev, nvalues, scale);
  else if (dtype == PFB_DTYPE_C64) pfb_pack_kernel<float, short><<<(unsigned)grid, 256, 0, st>>>((const float*)in_dev, (short*)out_dev, nvalues, (float)scale);
  else pfb_pack_kernel<double, short><<<(unsigned)grid, 256, 0, st>>>((const double*)in_dev, (short*)out_dev, nvalues, scale);
  return check_launch("pfb_pack_kernel");
}

int pfb_purity_metrics_pairs(const void* args_dev, int npairs, int nparts, void* stream) {
  if (!args_dev || npairs < 0 || nparts < 1 || nparts > 64) return fail(PFB_ERR_ARG, "pfb_purity_metrics_pairs: bad argument");
  if (npairs == 0) return PFB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  pfb_metrics_pair_kernel<<<dim3((unsigned)npairs, (unsigned)nparts), 1024, 0, st>>>(reinterpret_cast<const MetricsPairArgs*>(args_dev));
  return check_launch("pfb_metrics_pair_kernel");
}
