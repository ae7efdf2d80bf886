/*
 * dspeed_b200 -- C ABI of the B200 (sm_100a) implementation of dspeed's
 * ProcessingChain block-execution hot path.
 *
 * This is the drop-in boundary: one `extern "C"` launcher per processor of
 * `dspeed.processors` on the hot path (reference: src/dspeed/processors/, cited per
 * function), plus the fused waveform-resident chain program (dspb_chain_*), which is
 * what `ProcessingChain.execute` (src/dspeed/processing_chain.py:665-673,1144-1163)
 * launches once per block instead of one numba gufunc call per processor.
 *
 * Conventions (all launchers):
 *  - every pointer is a DEVICE pointer owned by the caller (PyTorch tensors); the
 *    library never allocates, frees or retains them.  Launches are asynchronous on
 *    `stream` (a cudaStream_t passed as void*); the caller keeps buffers alive until
 *    the stream reaches the launch.
 *  - a waveform operand is (ptr, row_stride, dtype): `n_rows` rows of `n` samples,
 *    consecutive rows `row_stride` ELEMENTS apart (views/slices of a larger block are
 *    expressed by offsetting ptr), element type `dtype` = DSPB_F32/F64/U16/I16/I32/U32.
 *    Integer waveforms are converted on load exactly like numpy casts them into the
 *    reference's float loop.
 *  - the suffix _f32 / _f64 is the reference's type loop ("f" / "d"): the dtype of all
 *    outputs and of every per-row scalar array.
 *  - a per-row scalar argument is (ptr, stride, imm): if ptr != NULL the value for row r
 *    is ptr[r*stride] (stride 0 broadcasts one device value), else the immediate `imm`.
 *  - outputs are written in place into caller-allocated arrays (gufunc convention,
 *    reference processors/__init__.py:47-59); waveform outputs take (ptr, row_stride).
 *  - NaN convention of the reference (docs/source/manuals/build_dsp.rst:152-175): if any
 *    input sample or scalar of a row is NaN all outputs of that row are NaN.
 *  - errors: return 0 on success; > 0 = a DSPFatal condition detectable from the
 *    arguments alone (DSPB_FATAL_*, same messages as the reference); < 0 = -cudaError_t.
 *    Data-dependent DSPFatal conditions are recorded on the device in `fatal`
 *    (int32[4]: code, row low 31 bits, row high bits, reserved; first writer wins;
 *    may be NULL) and raised by the host wrapper after the block completes, mirroring
 *    processing_chain.py:1156-1159.
 *  - re-entrant: no global mutable state; any stream, any device (current device =
 *    the one owning the pointers).
 */
#ifndef DSPEED_B200_H
#define DSPEED_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element types of waveform operands */
enum { DSPB_F32 = 0, DSPB_F64 = 1, DSPB_U16 = 2, DSPB_I16 = 3, DSPB_I32 = 4, DSPB_U32 = 5 };

/* DSPFatal conditions (reference file:line) */
enum {
  DSPB_OK = 0,
  DSPB_FATAL_PZ_NAN = 1,            /* pole_zero.py:76-77 */
  DSPB_FATAL_DPZ_SHORT = 2,         /* pole_zero.py:145-148 */
  DSPB_FATAL_RISE_NEG = 3,          /* trap_filters.py:53-54 */
  DSPB_FATAL_FLAT_NEG = 4,          /* trap_filters.py:56-57 */
  DSPB_FATAL_FALL_NEG = 5,          /* trap_filters.py:205-206 */
  DSPB_FATAL_TRAP_WIDE = 6,         /* trap_filters.py:59-60 */
  DSPB_FATAL_PICKOFF_NONINT = 7,    /* trap_filters.py:278-279 */
  DSPB_FATAL_MW_RANGE = 8,          /* moving_windows.py:52-55 */
  DSPB_FATAL_MWM_LEN_NONINT = 9,    /* moving_windows.py:167-168 */
  DSPB_FATAL_MWM_NUM_NONINT = 10,   /* moving_windows.py:170-171 */
  DSPB_FATAL_MWM_RANGE = 11,        /* moving_windows.py:173-174 */
  DSPB_FATAL_MWM_NUM_NEG = 12,      /* moving_windows.py:176-177 */
  DSPB_FATAL_TSTART_NONINT = 13,    /* time_point_thresh.py:67-68 */
  DSPB_FATAL_WALK_NONINT = 14,      /* time_point_thresh.py:70-71 */
  DSPB_FATAL_TSTART_RANGE = 15,     /* time_point_thresh.py:73-74 */
  DSPB_FATAL_INTERP_MODE = 16,      /* time_point_thresh.py:222, fixed_time_pickoff.py:125 */
  DSPB_FATAL_POLARITY_ZERO = 17,    /* time_point_thresh.py:314 */
  DSPB_FATAL_FTP_INT = 18,          /* fixed_time_pickoff.py:85 */
  DSPB_FATAL_WINDOWER_LEN = 19,     /* windower.py:42-43 */
  DSPB_FATAL_UPSAMPLE = 20,         /* upsampler.py:41-42 */
  DSPB_FATAL_CONV_KERNEL_LONG = 21, /* convolutions.py:48-49 */
  DSPB_FATAL_CONV_MODE = 22,        /* convolutions.py:70 */
  DSPB_FATAL_CONV_OUTLEN = 23,      /* convolutions.py:53-68 */
  DSPB_FATAL_GMLE_LEN = 24,         /* get_multi_local_extrema.py:126-129 */
  DSPB_FATAL_GMLE_DELTA = 25,       /* get_multi_local_extrema.py:130-131 */
  DSPB_FATAL_GMLE_DIR = 26,         /* get_multi_local_extrema.py:305-306 */
  DSPB_FATAL_RF_B_SCALAR = 27,      /* recursive_filter.py:66-67 */
  DSPB_FATAL_RF_SHORT = 28,         /* recursive_filter.py:68-71 */
  DSPB_FATAL_SHAPE = 29,            /* gufunc core-dimension mismatch (numpy raises) */
  DSPB_FATAL_KERNEL_ARGS = 30,      /* energy_kernels.py:51-61, kernels.py:49-56 */
  DSPB_FATAL_HIST_LEN = 31,         /* histogram.py:64-65, 152-153 */
  DSPB_FATAL_HIST_NAN = 32,         /* histogram.py:157-158 */
  DSPB_FATAL_HPS_NAN = 33,          /* histogram_stats.py:84-85 */
  DSPB_FATAL_HPS_LEN = 34,          /* histogram_stats.py:88-89, 221-222 */
  DSPB_FATAL_HPS_WIDTH_TYPE = 35,   /* histogram_stats.py:142-143 */
  DSPB_FATAL_RCCR2_NAN = 36,        /* rc_cr2.py:92-93 */
  DSPB_FATAL_INJ_FRAC = 37,         /* pole_zero.py:316-317 */
  DSPB_ERR_ROW_TOO_LONG = 100,      /* waveform does not fit the shared-memory resident layout */
  DSPB_ERR_UNSUPPORTED = 101
};

/* library / device information */
int dspb_version(void);
/* largest waveform length (samples of the compute dtype) `n_slots` resident copies allow */
int64_t dspb_max_row_len(int elem_bytes, int n_slots);
const char* dspb_fatal_message(int code);

#define DSPB_WAVE_IN(name) const void* name, int64_t name##_row_stride, int32_t name##_dtype
#define DSPB_WAVE_OUT(name) void* name, int64_t name##_row_stride
#define DSPB_SCALAR(name) const void* name, int64_t name##_stride, double name##_imm
#define DSPB_TAIL int32_t* fatal, void* stream

#define DSPB_DECLARE(SFX)                                                                          \
  /* bl_subtract.py:11-46   w_out = w_in - a_baseline */                                           \
  int dspb_bl_subtract##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, DSPB_SCALAR(a_baseline), \
                            DSPB_WAVE_OUT(w_out), DSPB_TAIL);                                      \
  /* min_max.py:11-82       first arg-min / arg-max (as floats) and the extreme values */          \
  int dspb_min_max##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, void* t_min, void* t_max,   \
                        void* a_min, void* a_max, DSPB_TAIL);                                      \
  /* numpy.amax(w, axis=1) as used by the configs (icpc-dsp-config.json:123-129) */                \
  int dspb_amax##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, void* a_max, DSPB_TAIL);       \
  /* min_max.py:85-140 */                                                                          \
  int dspb_min_max_norm##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, DSPB_SCALAR(a_min),    \
                             DSPB_SCALAR(a_max), DSPB_WAVE_OUT(w_out), DSPB_TAIL);                 \
  /* linear_slope_fit.py:11-90   mean, stdev(ddof=1), slope, intercept */                          \
  int dspb_linear_slope_fit##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, void* mean,        \
                                 void* stdev, void* slope, void* intercept, DSPB_TAIL);            \
  /* linear_slope_fit.py:93-158 */                                                                 \
  int dspb_linear_slope_diff##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,                   \
                                  DSPB_SCALAR(slope), DSPB_SCALAR(intercept), void* mean,          \
                                  void* rms, DSPB_TAIL);                                           \
  /* arithmetic.py:9-62 */                                                                         \
  int dspb_mean_below_threshold##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,                \
                                     DSPB_SCALAR(threshold), void* result, DSPB_TAIL);             \
  /* pole_zero.py:24-77 */                                                                         \
  int dspb_pole_zero##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, DSPB_SCALAR(t_tau),       \
                          DSPB_WAVE_OUT(w_out), DSPB_TAIL);                                        \
  /* pole_zero.py:82-198 */                                                                        \
  int dspb_double_pole_zero##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,                    \
                                 DSPB_SCALAR(t_tau1), DSPB_SCALAR(t_tau2), DSPB_SCALAR(frac),      \
                                 DSPB_WAVE_OUT(w_out), DSPB_TAIL);                                 \
  /* trap_filters.py:12-76 (norm=0) and :79-149 (norm=1) */                                        \
  int dspb_trap_filter##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, int32_t rise,           \
                            int32_t flat, int32_t norm, DSPB_WAVE_OUT(w_out), DSPB_TAIL);          \
  /* trap_filters.py:152-227 */                                                                    \
  int dspb_asym_trap_filter##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, int32_t rise,      \
                                 int32_t flat, int32_t fall, DSPB_WAVE_OUT(w_out), DSPB_TAIL);     \
  /* trap_filters.py:230-301 */                                                                    \
  int dspb_trap_pickoff##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, int32_t rise,          \
                             int32_t flat, DSPB_SCALAR(t_pickoff), void* a_out, DSPB_TAIL);        \
  /* moving_windows.py:12-61 (kind 0 = left), :64-114 (kind 1 = right) */                          \
  int dspb_moving_window##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, double length,        \
                              int32_t kind, DSPB_WAVE_OUT(w_out), DSPB_TAIL);                      \
  /* moving_windows.py:117-203 */                                                                  \
  int dspb_moving_window_multi##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, double length,  \
                                    double num_mw, int32_t mw_type, DSPB_WAVE_OUT(w_out),          \
                                    DSPB_TAIL);                                                    \
  /* moving_windows.py:206-249 ; w_out has n_out = n - int(length) samples */                      \
  int dspb_avg_current##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, double length,          \
                            DSPB_WAVE_OUT(w_out), int64_t n_out, DSPB_TAIL);                       \
  /* time_point_thresh.py:12-92 */                                                                 \
  int dspb_time_point_thresh##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,                   \
                                  DSPB_SCALAR(a_threshold), DSPB_SCALAR(t_start),                  \
                                  DSPB_SCALAR(walk_forward), void* t_out, DSPB_TAIL);              \
  /* time_point_thresh.py:95-222 ; mode_in is the interpolation character */                       \
  int dspb_interpolated_time_point_thresh##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,      \
                                               DSPB_SCALAR(a_threshold), DSPB_SCALAR(t_start),     \
                                               int64_t walk_forward, int32_t mode_in, void* t_out, \
                                               DSPB_TAIL);                                         \
  /* time_point_thresh.py:225-401 ; thresholds [n_rows or 1, m], t_out [n_rows, m] */              \
  int dspb_multi_time_point_thresh##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,             \
                                        const void* a_threshold, int64_t m, int64_t thr_row_stride,\
                                        DSPB_SCALAR(t_start), double polarity, int32_t mode_in,    \
                                        void* t_out, DSPB_TAIL);                                   \
  /* fixed_time_pickoff.py:12-125 */                                                               \
  int dspb_fixed_time_pickoff##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,                  \
                                   DSPB_SCALAR(t_in), int32_t mode_in, void* a_out, DSPB_TAIL);    \
  /* windower.py:12-54 ; w_out has m samples */                                                    \
  int dspb_windower##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, DSPB_SCALAR(t0_in),        \
                         DSPB_WAVE_OUT(w_out), int64_t m, DSPB_TAIL);                              \
  /* upsampler.py:14-49 ; w_out has m samples */                                                   \
  int dspb_upsampler##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, double upsample,          \
                          DSPB_WAVE_OUT(w_out), int64_t m, DSPB_TAIL);                             \
  /* convolutions.py:14-72 (convolve_wf) and :75-119 (fft_convolve_wf: same sums, evaluated        \
   * directly); kernel [m] of the output dtype on the device; mode_in 'f'|'v'|'s'; w_out has p. */ \
  int dspb_convolve_wf##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, const void* kernel,     \
                            int64_t m, int32_t mode_in, DSPB_WAVE_OUT(w_out), int64_t p,           \
                            DSPB_TAIL);                                                            \
  /* get_multi_local_extrema.py:12-306 ; vt_max/vt_min [n_rows, m], n_max/n_min uint32 [n_rows] */ \
  int dspb_get_multi_local_extrema##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,             \
                                        double a_delta_max, double a_delta_min,                    \
                                        double search_direction, DSPB_SCALAR(a_abs_max),           \
                                        DSPB_SCALAR(a_abs_min), void* vt_max, void* vt_min,        \
                                        int64_t m,                                                 \
                                        uint32_t* n_max, uint32_t* n_min, DSPB_TAIL);              \
  /* recursive_filter.py:12-93 ; a[p], b[q] host doubles (q <= 3) */                               \
  int dspb_recursive_filter##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, const double* a,   \
                                 int64_t p, const double* b, int64_t q, DSPB_SCALAR(init_in),      \
                                 DSPB_SCALAR(init_out), DSPB_WAVE_OUT(w_out), DSPB_TAIL);          \
  /* set-up time kernel synthesis on the device (const-folded by the chain compiler,               \
   * processing_chain.py:2775-2820): energy_kernels.py:12-73, :76-157, kernels.py:12-61 */         \
  int dspb_cusp_filter##SFX(double sigma, double flat, double decay, void* kernel, int64_t length, \
                            void* stream);                                                         \
  int dspb_zac_filter##SFX(double sigma, double flat, double decay, void* kernel, int64_t length,  \
                           void* stream);                                                          \
  int dspb_t0_filter##SFX(double rise, double fall, void* kernel, int64_t length, void* stream);

DSPB_DECLARE(_f32)
DSPB_DECLARE(_f64)

/* ---- the rest of the SiPM / LAr chain (tests/configs/sipm-dsp-config.json), csrc/sipm.cu ------------------------
 * List-valued operands (histogram weights / borders, index lists) are [n_rows, m] arrays of the loop's type with a row
 * stride in elements. */
#define DSPB_DECLARE_SIPM(SFX)                                                                                           \
  /* histogram.py:14-89   weights[m], borders[m + 1] of the row's min..max range */                                      \
  int dspb_histogram##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, void* weights, int64_t weights_rs,              \
                          int64_t n_bins, void* borders, int64_t borders_rs, int64_t n_borders, DSPB_TAIL);             \
  /* histogram.py:92-204   bins of width bin_width centred on `center` (NaN: on the mode of the row) */                  \
  int dspb_histogram_around_mode##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, DSPB_SCALAR(center),                \
                                      DSPB_SCALAR(bin_width), void* weights, int64_t weights_rs, int64_t n_bins,        \
                                      void* borders, int64_t borders_rs, int64_t n_borders, DSPB_TAIL);                 \
  /* histogram_stats.py:146-261   mode index, left edge of the mode bin, half width at half maximum */                   \
  int dspb_histogram_stats##SFX(const void* weights, int64_t weights_rs, int64_t n_bins, const void* edges,              \
                                int64_t edges_rs, int64_t n_edges, int64_t n_rows, void* mode_out, void* max_out,       \
                                void* fwhm_out, DSPB_SCALAR(max_in), DSPB_TAIL);                                        \
  /* histogram_stats.py:12-143   centre of the mode bin, FWHM / HWHM variants */                                         \
  int dspb_histogram_peakstats##SFX(const void* weights, int64_t weights_rs, int64_t n_bins, const void* edges,          \
                                    int64_t edges_rs, int64_t n_edges, int64_t n_rows, DSPB_SCALAR(max_in),             \
                                    int32_t skip_zeroes, int32_t width_type, void* mode_out, void* width_out,           \
                                    DSPB_TAIL);                                                                         \
  /* peak_snr_threshold.py:11-71   candidates whose local minimum / value ratio is below ratio_in; count uint32 */       \
  int dspb_peak_snr_threshold##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, const void* idx_in,                    \
                                   int64_t idx_in_rs, int64_t m, DSPB_SCALAR(ratio_in), DSPB_SCALAR(width_in),          \
                                   void* idx_out, int64_t idx_out_rs, void* n_idx_out, DSPB_TAIL);                      \
  /* multi_a_filter.py:11-57   waveform values at the (integer) times of the list, NaN padded */                         \
  int dspb_multi_a_filter##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, const void* vt_maxs_in, int64_t vt_rs,     \
                               int64_t m, void* va_max_out, int64_t va_rs, DSPB_TAIL);                                  \
  /* convolutions.py:122-182   'same' convolution of the reflect-padded waveform */                                      \
  int dspb_reflected_convolve_wf##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, const void* kernel, int64_t m,      \
                                      DSPB_WAVE_OUT(w_out), DSPB_TAIL);                                                 \
  /* recursive_filter.py:12-93 for any order (a[p], b[q] host doubles, p, q <= 16): the sequential float64 recursion */   \
  int dspb_recursive_filter_general##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, const double* a, int64_t p,      \
                                         const double* b, int64_t q, DSPB_SCALAR(init_in), DSPB_SCALAR(init_out),       \
                                         DSPB_WAVE_OUT(w_out), DSPB_TAIL);                                              \
  /* rc_cr2.py:11-93 */                                                                                                 \
  int dspb_rc_cr2##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, DSPB_SCALAR(t_tau), DSPB_WAVE_OUT(w_out),          \
                       DSPB_TAIL);                                                                                      \
  /* set-up time kernel generators: gaussian_filter1d.py:46-82, kernels.py:64-103, kernels.py:106-142 */                \
  int dspb_gaussian_filter1d##SFX(double sigma, double truncate, void* weights, int64_t length, void* stream);          \
  int dspb_moving_slope##SFX(void* kernel, int64_t length, void* stream);                                               \
  int dspb_step##SFX(double weight_pos, void* kernel, int64_t length, void* stream);

DSPB_DECLARE_SIPM(_f32)
DSPB_DECLARE_SIPM(_f64)

/* ---- numpy-level glue (csrc/glue.cu): the ufuncs of the expression parser (processing_chain.py:46-59), `where`
 * (where.py:12-54), round / floor / ceil / trunc to nearest (round_to_nearest.py:11-200), unit conversion
 * (unit_conversion.py:16-78), astype (processing_chain.py:1269-1300) -- out[r, j] = op(a[r, j], b[r, j], c[r, j]).
 *   op    0 add 1 subtract 2 multiply 3 divide 4 floor_divide 5 negative 6 absolute 7 sqrt 8 maximum 9 minimum
 *         16 equal 17 not_equal 18 less 19 less_equal 20 greater 21 greater_equal 22 isnan 23 isfinite
 *         32 where(a ? b : c) 33 copy (astype) 40..43 b * {rint, floor, ceil, trunc}(a / b) 48 (a + b) * ratio - c
 *   loop  the type operands are converted to and the operation runs in: 0 float32, 1 float64, 2 int64
 *   dtype codes: DSPB_F32 .. DSPB_U32 (0..5), 6 int64, 7 bool, 8 int8, 9 uint8, 10 uint64
 *   an operand is (ptr, row stride, column stride, dtype, immediate): ptr NULL = the immediate; stride 0 broadcasts
 *   mode (op 48): 0 none 1 rint 2 floor 3 ceil 4 trunc */
int dspb_glue(int32_t op, int32_t loop, int64_t rows, int64_t inner, void* out, int64_t out_rs, int32_t out_dt,
              const void* a, int64_t a_rs, int64_t a_cs, int32_t a_dt, double a_imm,
              const void* b, int64_t b_rs, int64_t b_cs, int32_t b_dt, double b_imm,
              const void* c, int64_t c_rs, int64_t c_cs, int32_t c_dt, double c_imm, double ratio, int32_t mode,
              void* stream);
/* get.py:10-91: out[r] = a[r, idx[r]] (negative indices wrap); use_default: `dflt` replaces out-of-range / NaN entries,
 * else an out-of-range index records fatal code 32 */
int dspb_glue_get(int32_t loop, int64_t rows, const void* a, int64_t a_rs, int32_t a_dt, int64_t n, const void* idx,
                  int64_t idx_rs, int32_t idx_dt, double idx_imm, const void* dflt, int64_t dflt_rs, int32_t dflt_dt,
                  double dflt_imm, int32_t use_default, void* out, int32_t out_dt, int32_t* fatal, void* stream);

/* VectorOfVectors output compaction on the device (LGDOVectorOfVectorsIOManager.write, processing_chain.py:2230-2260):
 * lens uint32[n_rows] (clamped to `width`) -> absolute end offsets int64[n_rows] starting from `base` and the column's
 * cumulative_length uint32[n_rows]; then the padded block [n_rows, width] of elem_bytes (2 / 4 / 8) byte elements ->
 * flat[0 .. ends[n_rows - 1] - base). */
int dspb_vov_offsets(const uint32_t* lens, int64_t n_rows, int64_t width, int64_t base, int64_t* ends,
                     uint32_t* cumulative_length, void* stream);
int dspb_vov_compact(const void* block, int64_t row_stride, int32_t elem_bytes, const uint32_t* lens, int64_t width,
                     const int64_t* ends, int64_t n_rows, int64_t base, void* flat, void* stream);

/* ---- fused waveform-resident chain program (see DESIGN.md "chain compiler") ---------
 * A program is a flat int32/double blob produced by the host chain compiler
 * (dspeed_b200/fusion.py); the kernel interprets it with one CTA per waveform: the raw
 * row is read from HBM once, all intermediates live in shared-memory slots / a
 * per-row scalar file, only requested outputs are written back. */
/* Convolution of a block of float32 waveforms with one generic kernel on the tensor cores
 * (csrc/conv_tc.cu: banded Toeplitz GEMM, 3xTF32 on tcgen05 / TMEM, operands staged by TMA) -- the
 * tensor-core variant of convolve_wf / fft_convolve_wf (convolutions.py:14-119) for long kernels.
 * x [n_rows, L] (row pitch x_stride elements, 16-byte aligned rows), kern [K], mode_in 'f'|'v'|'s',
 * out [n_rows, p] with p = L + K - 1 | L - K + 1 | L, workspace: dspb_convolve_tc_workspace(K) floats; all on
 * the device.  A NaN in a waveform or in the kernel gives an all-NaN output row (convolutions.py:44-46). */
int64_t dspb_convolve_tc_workspace(int64_t K);
int dspb_convolve_tc_f32(const float* x, int64_t x_stride, int64_t n_rows, int64_t L, const float* kern, int64_t K,
                         int32_t mode_in, float* out, int64_t out_stride, int64_t p, float* workspace,
                         int64_t workspace_floats, void* stream);

typedef struct dspb_chain dspb_chain;
int dspb_chain_create(const int32_t* code, int64_t n_code, const double* consts, int64_t n_consts,
                      dspb_chain** out);
/* ptrs: device pointer table indexed by the program (inputs, outputs, kernels) */
int dspb_chain_launch(dspb_chain* chain, const void* const* ptrs, int64_t n_ptrs, int64_t n_rows,
                      int32_t* fatal, void* stream);
int64_t dspb_chain_smem_bytes(const dspb_chain* chain);
/* tracing (reference: the per-processor timers of processing_chain.py:1778-1781,1188-1190):
 * per-instruction SM-cycle counters of CTA 0.  enable != 0 (re)starts counting; when `out`
 * is non-NULL the n_instr counters accumulated so far are copied to it first. */
int dspb_chain_profile(dspb_chain* chain, int enable, int64_t* out);
void dspb_chain_destroy(dspb_chain* chain);

#ifdef __cplusplus
}
#endif
#endif /* DSPEED_B200_H */
