/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * Row loops (OpenMP over waveforms) around the single-waveform restatements in
 * dsp_oracle_impl.h.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Conventions: waveform arrays are C-contiguous [n_rows, n]; every per-row
 * scalar argument is a pointer plus an element stride (0 = one value broadcast
 * to all rows).  Return value: 0, or the first (lowest-row) DSPFatal code
 * (dsp_oracle.h); *bad_row receives that row index when non-NULL.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "dsp_oracle.h"

#define REAL float
#define SUFFIX _f32
#define SQRT_REAL sqrtf
#define FLOOR_REAL floorf
#define ABS_REAL fabsf
#include "dsp_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef SQRT_REAL
#undef FLOOR_REAL
#undef ABS_REAL

#define REAL double
#define SUFFIX _f64
#define SQRT_REAL sqrt
#define FLOOR_REAL floor
#define ABS_REAL fabs
#include "dsp_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef SQRT_REAL
#undef FLOOR_REAL
#undef ABS_REAL

static void note_fatal(int rc, int64_t r, int *first_rc, int64_t *first_row) {
  if (!rc) return;
#pragma omp critical(orc_fatal)
  {
    if (*first_rc == 0 || r < *first_row) {
      *first_rc = rc;
      *first_row = r;
    }
  }
}

int orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

#define ROWLOOP_BEGIN                 \
  int first_rc = 0;                   \
  int64_t first_row = -1;             \
  _Pragma("omp parallel for schedule(static)") for (int64_t r = 0; r < n_rows; r++) {
#define ROWLOOP_END                   \
  }                                   \
  if (bad_row) *bad_row = first_row;  \
  return first_rc;

#define GEN(T, S)                                                                                  \
  int orc_bl_subtract##S(const T *w_in, int64_t n_rows, int64_t n, const T *bl, int64_t bl_s,      \
                         T *w_out, int64_t *bad_row) {                                             \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_bl_subtract##S(w_in + r * n, n, bl[r * bl_s], w_out + r * n), r, &first_rc,    \
               &first_row);                                                                        \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_linear_slope_fit##S(const T *w_in, int64_t n_rows, int64_t n, int64_t row_stride,        \
                              T *mean, T *stdev, T *slope, T *intercept, int64_t *bad_row) {       \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_linear_slope_fit##S(w_in + r * row_stride, n, mean + r, stdev + r, slope + r,  \
                                        intercept + r),                                            \
               r, &first_rc, &first_row);                                                          \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_linear_slope_diff##S(const T *w_in, int64_t n_rows, int64_t n, int64_t row_stride,       \
                               const T *slope, int64_t s_s, const T *icpt, int64_t i_s, T *mean,   \
                               T *rms, int64_t *bad_row) {                                         \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_linear_slope_diff##S(w_in + r * row_stride, n, slope[r * s_s], icpt[r * i_s],  \
                                         mean + r, rms + r),                                       \
               r, &first_rc, &first_row);                                                          \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_mean_below_threshold##S(const T *w_in, int64_t n_rows, int64_t n, const T *thr,          \
                                  int64_t thr_s, T *res, int64_t *bad_row) {                       \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_mean_below_threshold##S(w_in + r * n, n, thr[r * thr_s], res + r), r,          \
               &first_rc, &first_row);                                                             \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_pole_zero##S(const T *w_in, int64_t n_rows, int64_t n, const T *tau, int64_t tau_s,      \
                       T *w_out, int64_t *bad_row) {                                               \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_pole_zero##S(w_in + r * n, n, tau[r * tau_s], w_out + r * n), r, &first_rc,    \
               &first_row);                                                                        \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_double_pole_zero##S(const T *w_in, int64_t n_rows, int64_t n, const T *tau1,             \
                              int64_t s1, const T *tau2, int64_t s2, const T *frac, int64_t s3,    \
                              T *w_out, int64_t *bad_row) {                                        \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_double_pole_zero##S(w_in + r * n, n, tau1[r * s1], tau2[r * s2],               \
                                        frac[r * s3], w_out + r * n),                              \
               r, &first_rc, &first_row);                                                          \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_trap_filter##S(const T *w_in, int64_t n_rows, int64_t n, int32_t rise, int32_t flat,     \
                         T *w_out, int64_t *bad_row) {                                             \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_trap_filter##S(w_in + r * n, n, rise, flat, w_out + r * n), r, &first_rc,      \
               &first_row);                                                                        \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_trap_norm##S(const T *w_in, int64_t n_rows, int64_t n, int32_t rise, int32_t flat,       \
                       T *w_out, int64_t *bad_row) {                                               \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_trap_norm##S(w_in + r * n, n, rise, flat, w_out + r * n), r, &first_rc,        \
               &first_row);                                                                        \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_asym_trap_filter##S(const T *w_in, int64_t n_rows, int64_t n, int32_t rise,              \
                              int32_t flat, int32_t fall, T *w_out, int64_t *bad_row) {            \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_asym_trap_filter##S(w_in + r * n, n, rise, flat, fall, w_out + r * n), r,      \
               &first_rc, &first_row);                                                             \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_trap_pickoff##S(const T *w_in, int64_t n_rows, int64_t n, int32_t rise, int32_t flat,    \
                          const T *t, int64_t t_s, T *a_out, int64_t *bad_row) {                   \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_trap_pickoff##S(w_in + r * n, n, rise, flat, t[r * t_s], a_out + r), r,        \
               &first_rc, &first_row);                                                             \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_moving_window_left##S(const T *w_in, int64_t n_rows, int64_t n, T length, T *w_out,      \
                                int64_t *bad_row) {                                                \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_moving_window_left##S(w_in + r * n, n, length, w_out + r * n), r, &first_rc,   \
               &first_row);                                                                        \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_moving_window_right##S(const T *w_in, int64_t n_rows, int64_t n, T length, T *w_out,     \
                                 int64_t *bad_row) {                                               \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_moving_window_right##S(w_in + r * n, n, length, w_out + r * n), r, &first_rc,  \
               &first_row);                                                                        \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_moving_window_multi##S(const T *w_in, int64_t n_rows, int64_t n, T length, T num_mw,     \
                                 int32_t mw_type, T *w_out, int64_t *bad_row) {                    \
    ROWLOOP_BEGIN                                                                                  \
    T *scr = (T *)malloc((size_t)n * sizeof(T));                                                   \
    note_fatal(orc1_moving_window_multi##S(w_in + r * n, n, length, num_mw, mw_type,               \
                                           w_out + r * n, scr),                                    \
               r, &first_rc, &first_row);                                                          \
    free(scr);                                                                                     \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_avg_current##S(const T *w_in, int64_t n_rows, int64_t n, T length, T *w_out,             \
                         int64_t n_out, int64_t *bad_row) {                                        \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_avg_current##S(w_in + r * n, n, length, w_out + r * n_out, n_out), r,          \
               &first_rc, &first_row);                                                             \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_time_point_thresh##S(const T *w, int64_t n_rows, int64_t n, const T *thr, int64_t thr_s, \
                               const T *ts, int64_t ts_s, T walk, T *t_out, int64_t *bad_row) {    \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_time_point_thresh##S(w + r * n, n, thr[r * thr_s], ts[r * ts_s], walk,         \
                                         t_out + r),                                               \
               r, &first_rc, &first_row);                                                          \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_interpolated_time_point_thresh##S(const T *w, int64_t n_rows, int64_t n, const T *thr,   \
                                            int64_t thr_s, const T *ts, int64_t ts_s,              \
                                            int64_t walk, int8_t mode, T *t_out,                   \
                                            int64_t *bad_row) {                                    \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_interpolated_time_point_thresh##S(w + r * n, n, thr[r * thr_s], ts[r * ts_s],  \
                                                      walk, mode, t_out + r),                      \
               r, &first_rc, &first_row);                                                          \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_multi_time_point_thresh##S(const T *w, int64_t n_rows, int64_t n, const T *thr,          \
                                     int64_t m, int64_t thr_row_s, const T *ts, int64_t ts_s,      \
                                     T polarity, int8_t mode, T *t_out, int64_t *bad_row) {        \
    ROWLOOP_BEGIN                                                                                  \
    int64_t *srt = (int64_t *)malloc((size_t)(m > 0 ? m : 1) * sizeof(int64_t));                   \
    note_fatal(orc1_multi_time_point_thresh##S(w + r * n, n, thr + r * thr_row_s, m, ts[r * ts_s], \
                                               polarity, mode, t_out + r * m, srt),                \
               r, &first_rc, &first_row);                                                          \
    free(srt);                                                                                     \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_fixed_time_pickoff##S(const T *w, int64_t n_rows, int64_t n, const T *t, int64_t t_s,    \
                                int8_t mode, T *a_out, int64_t *bad_row) {                         \
    ROWLOOP_BEGIN                                                                                  \
    double *scr = mode == 's' ? (double *)malloc((size_t)(2 * n) * sizeof(double)) : NULL;         \
    note_fatal(orc1_fixed_time_pickoff##S(w + r * n, n, t[r * t_s], mode, a_out + r, scr), r,      \
               &first_rc, &first_row);                                                             \
    free(scr);                                                                                     \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_min_max##S(const T *w, int64_t n_rows, int64_t n, T *t_min, T *t_max, T *a_min,          \
                     T *a_max, int64_t *bad_row) {                                                 \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_min_max##S(w + r * n, n, t_min + r, t_max + r, a_min + r, a_max + r), r,       \
               &first_rc, &first_row);                                                             \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_min_max_norm##S(const T *w, int64_t n_rows, int64_t n, const T *a_min, int64_t mn_s,     \
                          const T *a_max, int64_t mx_s, T *w_out, int64_t *bad_row) {              \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_min_max_norm##S(w + r * n, n, a_min[r * mn_s], a_max[r * mx_s],                \
                                    w_out + r * n),                                                \
               r, &first_rc, &first_row);                                                          \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_windower##S(const T *w_in, int64_t n_rows, int64_t n, const T *t0, int64_t t0_s,         \
                      T *w_out, int64_t m, int64_t *bad_row) {                                     \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_windower##S(w_in + r * n, n, t0[r * t0_s], w_out + r * m, m), r, &first_rc,    \
               &first_row);                                                                        \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_upsampler##S(const T *w_in, int64_t n_rows, int64_t n, T upsample, T *w_out, int64_t m,  \
                       int64_t *bad_row) {                                                         \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_upsampler##S(w_in + r * n, n, upsample, w_out + r * m, m), r, &first_rc,       \
               &first_row);                                                                        \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_convolve_wf##S(const T *w_in, int64_t n_rows, int64_t n, int64_t row_stride,             \
                         const T *kern, int64_t m, int8_t mode, T *w_out, int64_t p,               \
                         int64_t *bad_row) {                                                       \
    ROWLOOP_BEGIN                                                                                  \
    note_fatal(orc1_convolve_wf##S(w_in + r * row_stride, n, kern, m, mode, w_out + r * p, p), r,  \
               &first_rc, &first_row);                                                             \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_get_multi_local_extrema##S(const T *w, int64_t n_rows, int64_t n, T d_max, T d_min,      \
                                     T dir, T abs_max, T abs_min, T *vt_max, T *vt_min,            \
                                     int64_t m, uint32_t *n_max, uint32_t *n_min,                  \
                                     int64_t *bad_row) {                                           \
    ROWLOOP_BEGIN                                                                                  \
    double *scr = (double *)malloc((size_t)(6 * m + 1) * sizeof(double));                          \
    note_fatal(orc1_get_multi_local_extrema##S(w + r * n, n, d_max, d_min, dir, abs_max, abs_min,  \
                                               vt_max + r * m, vt_min + r * m, m, n_max + r,       \
                                               n_min + r, scr),                                    \
               r, &first_rc, &first_row);                                                          \
    free(scr);                                                                                     \
    ROWLOOP_END                                                                                    \
  }                                                                                                \
  int orc_recursive_filter##S(const T *w_in, int64_t n_rows, int64_t n, const double *a,           \
                              int64_t p, const double *b, int64_t q, const T *init_in,             \
                              int64_t ii_s, const T *init_out, int64_t io_s, T *w_out,             \
                              int64_t *bad_row) {                                                  \
    ROWLOOP_BEGIN                                                                                  \
    double *circ = (double *)malloc((size_t)(q > 0 ? q : 1) * sizeof(double));                     \
    note_fatal(orc1_recursive_filter##S(w_in + r * n, n, a, p, b, q, init_in[r * ii_s],            \
                                        init_out[r * io_s], w_out + r * n, circ),                  \
               r, &first_rc, &first_row);                                                          \
    free(circ);                                                                                    \
    ROWLOOP_END                                                                                    \
  }

GEN(float, _f32)
GEN(double, _f64)
