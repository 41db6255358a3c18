/* pfb_b200.h — C ABI of the B200-native oversampled PFB analysis / inversion.
 *
 * This is the drop-in boundary for the hot path of ska-telescope/ska-pst-dsp-model.
 * The reference has no FFI of its own (its "native" side is Matlab compiled by
 * mcc and a third-party numpy package), so each entry point below names the
 * reference interface it replaces; INTEGRATION.md shows the ctypes binding the
 * reference's python/data_gen would add.
 *
 * Conventions
 *  - every function returns 0 on success or a negative PFB_ERR_* code;
 *    pfb_last_error() returns a thread-local message for the last failure.
 *  - all data pointers are DEVICE pointers owned by the caller (e.g. a torch
 *    tensor's data_ptr()); plans own only their tables and scratch.
 *  - data layout is the DADA "TFP" order of the reference files
 *    (matlab/write_dada_data.m:32-50): time slowest, then channel, then
 *    polarisation, interleaved (re,im); dtype 0 = complex64, 1 = complex128.
 *  - calls are stream-ordered on `stream` (a cudaStream_t cast to void*; NULL =
 *    legacy default stream); one plan must not be used from two streams at once.
 */
#ifndef PFB_B200_H
#define PFB_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PFB_OK 0
#define PFB_ERR_ARG (-1)         /* invalid argument / unsupported shape */
#define PFB_ERR_CUDA (-2)        /* CUDA runtime error (see pfb_last_error) */
#define PFB_ERR_UNSUPPORTED (-3) /* valid in the reference but not implemented here */

#define PFB_DTYPE_C64 0
#define PFB_DTYPE_C128 1

/* analysis variants == config "analysis_function" (config/test.config.json) */
#define PFB_VARIANT_BUNTON 0 /* matlab/polyphase_analysis.m        */
#define PFB_VARIANT_PADDED 1 /* matlab/polyphase_analysis_padded.m */
#define PFB_VARIANT_LOWCBF 2 /* matlab/polyphase_analysis_lowcbf.m + PSTFilterbank.m */

typedef struct pfb_analysis_plan pfb_analysis_plan;
typedef struct pfb_inverse_plan pfb_inverse_plan;

const char* pfb_last_error(void);
int pfb_version(void);

/* ---- analysis: replaces out = polyphase_analysis[_padded|_lowcbf](in, filt, block, os_factor)
 *      (matlab/polyphase_analysis.m:1-7, FilterBank.m:91 plugin slot) and the
 *      pfb.format_handler.PSRFormatChannelizer call of python/data_gen/channelize.py:81-91.
 *      taps: ntaps doubles exactly as read from the coefficient file (padding,
 *      reversal and precision cast happen inside). */
int pfb_analysis_plan_create(pfb_analysis_plan** plan, int variant, int nchan, int os_nu, int os_de,
                             const double* taps, int ntaps, int dtype, int device);
int pfb_analysis_plan_destroy(pfb_analysis_plan* plan);
/* number of output time samples ("nblocks", polyphase_analysis.m:62 / _padded.m:75 /
 * PSTFilterbank.m:15) for n_dat input samples; first_call models the persistent
 * pre-padding flag of polyphase_analysis_lowcbf.m:27-34 (ignored otherwise). */
long long pfb_analysis_nblocks(const pfb_analysis_plan* plan, long long n_dat, int first_call);
int pfb_analysis_nchan_out(const pfb_analysis_plan* plan); /* block, or 216 for lowcbf */
/* in: [n_dat][npol]; out: [nblocks][nchan_out][npol] */
int pfb_analysis_execute(pfb_analysis_plan* plan, const void* in_dev, void* out_dev, int npol,
                         long long n_dat, int first_call, void* stream);
/* Batched form (TwoStageFilterBank.m:92-96: one stage-2 filterbank per coarse channel, here one launch):
 * in: [n_dat][nstreams] independent single-pol streams; out_layout 0: [nblocks][nchan_out][nstreams],
 * 1: [nblocks][nstreams][nchan_out] (coarse-major, the channel order TwoStageFilterBank produces). */
int pfb_analysis_execute_batched(pfb_analysis_plan* plan, const void* in_dev, void* out_dev, int nstreams,
                                 long long n_dat, int first_call, int out_layout, void* stream);
/* Time-segment form (multi-GPU sharding / streaming, cf. FilterBank.m:85-88,119-126):
 * compute output blocks [k_first, k_first+nblk) of the stream. in_dev[0] is absolute
 * stream sample `sample0`, n_local samples are resident; samples outside read as 0.
 * out_dev[0] is block k_first.  Bunton and lowcbf variants only. */
int pfb_analysis_execute_segment(pfb_analysis_plan* plan, const void* in_dev, long long n_local,
                                 long long sample0, void* out_dev, int npol, long long k_first,
                                 long long nblk, int first_call, void* stream);
/* The same, with the stream held in TWO buffers: samples [0, n_own) in in_dev and the following n_next samples in
 * next_dev - in a time-sharded run the neighbour GPU's input buffer, mapped into this process (CUDA IPC /
 * torch symmetric memory) and read over NVLink by the analysis kernel's own TMA copies.  This is the halo exchange of
 * SURVEY 8(e) (the tail carry of FilterBank.m:85-88,119-126) without an exchange step: nothing is copied, no
 * collective is called.  SKA-Low fast path only (PFB_ERR_UNSUPPORTED otherwise: exchange the halo instead);
 * the caller guarantees the neighbour's samples are in place (one barrier after loading, not one per call). */
int pfb_analysis_execute_segment_peer(pfb_analysis_plan* plan, const void* in_dev, long long n_own,
                                      const void* next_dev, long long n_next, long long sample0, void* out_dev,
                                      int npol, long long k_first, long long nblk, int first_call, void* stream);

/* ---- inversion: replaces out = polyphase_synthesis(in, spans_nyquist, input_fft_length,
 *      os_factor, deripple, sample_offset, input_overlap, temporal_taper, spectral_taper,
 *      combine) (matlab/polyphase_synthesis.m:1-13) and the PSRFormatSynthesizer call of
 *      python/data_gen/synthesize.py:84-95.
 *      temporal_taper: nfft doubles (the window itself, PFBWindow.m) or NULL;
 *      deripple_taps: the analysis filter (n_deripple_taps > 0 applies the
 *      polyphase_synthesis.m:138-150,242-251 correction) or NULL/0;
 *      sample_offset is applied by the caller as a pointer offset. */
int pfb_inverse_plan_create(pfb_inverse_plan** plan, int nchan, int os_nu, int os_de, int nfft,
                            int overlap, const double* temporal_taper, const double* deripple_taps,
                            int n_deripple_taps, int spans_nyquist, int combine, int dtype, int device);
int pfb_inverse_plan_destroy(pfb_inverse_plan* plan);
long long pfb_inverse_nblocks(const pfb_inverse_plan* plan, long long n_dat);
long long pfb_inverse_out_len(const pfb_inverse_plan* plan, long long n_dat); /* n_blocks*output_keep */
/* in: [n_dat][nchan][npol]; out: [out_len][npol] */
int pfb_inverse_execute(pfb_inverse_plan* plan, const void* in_dev, void* out_dev, int npol,
                        long long n_dat, void* stream);
/* Batched form (TwoStageInverseFilterBank.m:117-151: one inversion per output coarse channel):
 * in: [n_dat][ngroups][nchan][npol]; out: [nblocks*keep2][ngroups][nchan][npol] where keep2 =
 * FN_width - 2*overlap*de/nu; output sample T2*nchan + t1 of group g is element [T2][g][t1]. */
int pfb_inverse_execute_batched(pfb_inverse_plan* plan, const void* in_dev, void* out_dev, int ngroups,
                                int npol, long long n_dat, void* stream);
/* spectral (frequency-domain) taper of polyphase_synthesis.m:282 / InverseFilterBank.frequency_taper:
 * n = nchan*FN_width doubles multiplying the stitched spectrum (NULL clears it).  With a taper set
 * the inversion runs the literal-order chain (per-channel FFT -> weights -> channel IDFT -> bin IFFT). */
/* ---- fused analysis -> inversion in one process (python/data_gen/pipeline.py:62-84 followed by synthesize, or the
 *      generate -> channelize -> synthesize pass of verify/purity.py:165-207, WITHOUT a channelized file in between).
 *      The inverse DFT across channels of a channelized block equals N x the rotated polyphase arm sums the analysis
 *      kernel holds before its FFT, so the analysis writes the inversion's intermediate directly and the channel-IDFT
 *      kernel is not run.  chan_dev may be NULL (the channelized block is not wanted: the analysis then skips its FFT
 *      stage as well).  SKA-Low dual-pol complex64 fast path only, intermediate (n_blocks x 4096 bytes) below 4 GiB:
 *      PFB_ERR_UNSUPPORTED otherwise (call pfb_analysis_execute + pfb_inverse_execute instead).  NOT a replacement for the file-based drop-in:
 *      synthesize() reads channelized files and uses pfb_inverse_execute. */
int pfb_roundtrip_execute(pfb_analysis_plan* aplan, pfb_inverse_plan* iplan, const void* in_dev, void* chan_dev,
                          void* out_dev, int npol, long long n_dat, void* stream);
/* the same for nvec independent vectors of vec_samples samples laid back to back (vec_samples a multiple of step and
 * nchan): in [nvec * vec_samples][npol] -> out [nvec][n_blocks * output_keep][npol] */
int pfb_roundtrip_execute_vectors(pfb_analysis_plan* aplan, pfb_inverse_plan* iplan, const void* in_dev, void* out_dev,
                                  int npol, int nvec, long long vec_samples, void* stream);
/* nvec independent channelized series stored back to back, row pitch vec_rows (>= n_dat) fine samples each:
 * in [nvec][vec_rows][nchan][npol] -> out [nvec][n_blocks*output_keep][npol].  Replaces the per-vector synthesize
 * call of the purity sweep (python/verify/purity.py:175-177, matlab/current_performance.m:261-299); one launch per
 * kernel for the whole batch on the SKA-Low fast path. */
int pfb_inverse_execute_vectors(pfb_inverse_plan* plan, const void* in_dev, void* out_dev, int npol, int nvec,
                                long long vec_rows, long long n_dat, void* stream);
int pfb_inverse_set_spectral_taper(pfb_inverse_plan* plan, const double* taper, long long n);
/* bytes of scratch between the two inversion kernels (default 1 GiB: the whole stream in one chunk up to ~2^25
 * samples/pol at the SKA-Low shape; longer streams are processed in chunks of whole overlap-save blocks) */
int pfb_inverse_set_scratch_bytes(pfb_inverse_plan* plan, long long bytes);
/* Sizes the plan's scratch for streams of up to n_dat fine-channel samples (ngroups = 1 unless the batched call is
 * used) WITHOUT launching anything.  A plan is single-stream and not thread-safe; its scratch grows on demand with
 * cudaFree + cudaMalloc, which is illegal during stream capture and would leave an earlier captured graph with a
 * stale pointer - reserve first, then capture.  (The reference has no counterpart: Matlab allocates per call,
 * polyphase_synthesis.m:152-156.) */
int pfb_inverse_reserve(pfb_inverse_plan* plan, int npol, long long n_dat, int ngroups);
/* run only one stage of the chain over the whole input (profiling / roofline):
 * stage 0 = channel IDFT into the caller's v buffer [n_dat][nchan][npol],
 * stage 1 = overlap-save resampling from such a v buffer. */
int pfb_inverse_execute_stage(pfb_inverse_plan* plan, int stage, const void* in_dev, void* out_dev,
                              int npol, long long n_dat, void* stream);

/* per-kernel device timing of the inversion chain (CUDA events recorded around every
 * launch on the caller's stream).  get_profile synchronises on the recorded events,
 * returns summed milliseconds and launch counts for stage 0 (channel IDFT) and
 * stage 1 (resampling) since the last call, and resets the counters. */
int pfb_inverse_set_profiling(pfb_inverse_plan* plan, int on);
int pfb_inverse_get_profile(pfb_inverse_plan* plan, double* ms2, long long* launches2);

/* ---- on-device test vectors and purity reductions (python/data_gen/generate_test_vector.py:24-71,
 *      python/verify/util.py:15-43, purity.py:154-160).
 * kind 0: time_domain_impulse(n, offset=p0, width=p1); kind 1: complex_sinusoid(n, freq=p0 [bins],
 * phase=p1 [rad], bin_offset=p2); kind 2: the 32-tone frequency comb of matlab/FrequencyComb.m:22-46 with the
 * amplitudes / spacing of sgcht.m:394-432 (a_k = linspace(1, sqrt 2, 32), f_k = -0.5 + 1/128 + p2 + p0/n + k/32
 * cycles per sample, common phase p1); every polarisation gets the same signal; out: [n][npol]. */
int pfb_generate_test_vector(int kind, void* out_dev, long long n, int npol, int dtype, double p0,
                             double p1, double p2, void* stream);
/* same signal written to out[t * stride] only: lets the batch driver place two independent test vectors in the
 * two polarisation slots of one [n][2] stream, so that a pair runs through the dual-pol kernels in one pass */
int pfb_generate_test_vector_strided(int kind, void* out_dev, long long n, long long stride, int dtype,
                                     double p0, double p1, double p2, void* stream);
/* nvec vectors in one launch, parameters (p0, p1, p2 per vector) read from device memory so that a captured CUDA
 * graph of a whole batch can be replayed with new parameters: vector i is written to
 * out[(i / group) * group_stride + (i % group) + t * group], i.e. `group` vectors interleaved sample by sample
 * (group = 2: the two polarisation slots of one [n][2] stream). */
int pfb_generate_test_vectors(int kind, void* out_dev, long long n, int nvec, int group, long long group_stride,
                              int dtype, const double* params_dev, void* stream);
/* args_dev: device array of nvec records {const void* a; const void* b; long long n, stride_a, stride_b;
 * double scale_b; double* out;} (56 bytes each).  For every record one CTA reduces the series b (and the
 * reference a, may be NULL) and writes out[0..7] = sum|b-a|, max|b-a|, sum|b|^2, largest |b|^2,
 * second largest |b|^2, argmax, sum|b|, n. */
int pfb_purity_metrics(const void* args_dev, int nvec, int dtype, void* stream);
/* Pair form for series that carry two independent vectors in the polarisation slots of one [n][2] complex64 stream
 * (the batched sweep): records {const void* a; const void* b; long long n; double scale_b; double* out0; double* out1;}
 * (48 bytes); out0 / out1 get the 8 numbers above for slot 0 / slot 1.  nparts CTAs share a record: part y reduces
 * elements [n y / nparts, n (y+1) / nparts) and writes its 8 numbers at out + 8 y (out0 / out1 hold nparts rows);
 * the caller merges the rows. */
int pfb_purity_metrics_pairs(const void* args_dev, int npairs, int nparts, void* stream);
/* Spectra of npairs dual-pol series of nfft = 98304 (= 2 x 192 x 256, the purity report's fft_size for SKA-Low:
 * python/verify/purity.py:84,150,236-263; matlab/DomainPerformance.m:58-100) points each: in [pair][pitch][2] complex64
 * (pitch in samples), scratch and out [pair][nfft][2].  The bins are stored transposed, out[k1][k2] = X[k1 + 24 k2];
 * the report's reductions do not depend on the order.  Other lengths: PFB_ERR_UNSUPPORTED. */
int pfb_spectrum_vectors(const void* in_dev, long long pitch, void* scratch_dev, void* out_dev, int npairs,
                         long long nfft, void* stream);

/* kernels launched by this library since load (bench.py "gpu_launches") */
long long pfb_launch_count(void);

/* ---- integer DADA payloads: replaces the int8 / int16 -> floating-point cast of matlab/DADARead.m:41-53,66-70
 *      (and psr_formats.DADAFile.load_data on NBIT 8/16 files) and the integer write of
 *      matlab/write_dada_header.m:15-22 + FilterBank.m:106-113 (round(scale*x), ties away from zero, saturating
 *      like Matlab's int16()).  nvalues counts REAL values (2 per complex sample); buffers are 16-byte aligned
 *      device pointers.  The conversion is exact, so results equal a host-side cast bit for bit. */
int pfb_unpack_payload(const void* in_dev, void* out_dev, long long nvalues, int nbit, int dtype, void* stream);
int pfb_pack_payload(const void* in_dev, void* out_dev, long long nvalues, int nbit, int dtype, double scale,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PFB_B200_H */
