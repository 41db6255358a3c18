// Shape-specialised fast kernels (SKA-Low: 256 channels, OS 4/3, 12 taps/arm;
// inversion Nf=256 -> FN_width=192).  Selected by the plan when the shape matches.
#pragma once
#include <vector>
#include "pfb_generic.cuh"

namespace pfb {

inline int low_analysis_plan_init(int N, int nu, int de, int P, int dtype, const double* taps, int ntaps,
                                  std::vector<void*>* allocs, const void** d_low, bool* enabled) {
  *enabled = false;
  return 0;
}

template <typename T>
int low_analysis_launch(const AnalysisArgs& a, const void* d_low, cudaStream_t st, bool* done) {
  *done = false;
  return 0;
}

template <typename T>
int low_inverse_plan_init(int nchan, int nu, int de, int Nf, int ov, int spans, int combine,
                          const std::vector<T>& taper, const std::vector<double>& w, double gain,
                          std::vector<void*>* allocs, const void** d_low, bool* enabled) {
  *enabled = false;
  return 0;
}

template <typename T>
int low_chan_idft_launch(const void* d_low, const void* in, void* out, int npol, long long nrows, cudaStream_t st, bool* done) {
  *done = false;
  return 0;
}

template <typename T>
int low_resample_launch(const void* d_low, const void* v, void* out, int npol, long long nblk, cudaStream_t st, bool* done) {
  *done = false;
  return 0;
}

}  // namespace pfb
