/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU restatement ("oracle") of the legend-exp/dspeed hot-path processors.
 * This file is a type-generic body: it is included twice by dsp_oracle.c, once
 * with REAL=float (the reference's "f" type loop) and once with REAL=double
 * (the "d" loop).  Every function restates ONE reference processor for ONE
 * waveform (the gufunc core dimensions); the row loops live in dsp_oracle.c.
 *
 * The restatement follows the reference's *sequential* arithmetic including
 * numba's type promotion (float32 op int -> float64, in-place store rounds to
 * the array dtype), so that on identical inputs it reproduces the numba
 * processors bit-for-bit where the reference itself is deterministic
 * (see tests/test_oracle_vs_golden.py for which outputs are pinned bit-exact
 * and which to a tolerance).  Compile with -ffp-contract=off.
 *
 * Citations are file:line under /root/reference/src/dspeed/processors/.
 *
 * Return value of every function: 0 = ok, >0 = the DSPFatal the reference
 * would raise (codes in dsp_oracle.h).
 */

#ifndef REAL
#error "include from dsp_oracle.c"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)
/* Python negative-index wrap: w_out[i-1] at i == 0 reads the last element
 * (only reachable with degenerate rise/length == 0 arguments). */
#define PREV(i) ((i) > 0 ? (i)-1 : n - 1)

static int FN(any_nan)(const REAL *w, int64_t n) {
  for (int64_t i = 0; i < n; i++)
    if (isnan(w[i])) return 1;
  return 0;
}
static void FN(fill_nan)(REAL *w, int64_t n) {
  for (int64_t i = 0; i < n; i++) w[i] = (REAL)NAN;
}

/* bl_subtract.py:11-46 */
static int FN(orc1_bl_subtract)(const REAL *w_in, int64_t n, REAL a_baseline, REAL *w_out) {
  FN(fill_nan)(w_out, n);
  if (FN(any_nan)(w_in, n) || isnan(a_baseline)) return 0;
  for (int64_t i = 0; i < n; i++) w_out[i] = (REAL)(w_in[i] - a_baseline);
  return 0;
}

/* linear_slope_fit.py:11-90.  mean/stdev are 1-element REAL arrays in the
 * reference, so "mean += temp/(i+1)" and "stdev /= isum-1" are array-with-
 * scalar operations that numba evaluates in the array dtype (REAL); the
 * regression sums are int64 / float64 scalars.  Pinned bit-exact against the
 * reference by tests/test_oracle_vs_golden.py. */
static int FN(orc1_linear_slope_fit)(const REAL *w_in, int64_t n, REAL *mean_o, REAL *stdev_o,
                                     REAL *slope_o, REAL *intercept_o) {
  *mean_o = *stdev_o = *slope_o = *intercept_o = (REAL)NAN;
  if (FN(any_nan)(w_in, n)) return 0;
  REAL mean = 0, stdev = 0;
  int64_t sum_x = 0, sum_x2 = 0;
  double sum_xy = 0, sum_y = 0;
  for (int64_t i = 0; i < n; i++) {
    REAL temp = (REAL)(w_in[i] - mean);
    REAL q = (REAL)(temp / (REAL)(i + 1));
    mean = (REAL)(mean + q);
    REAL d2 = (REAL)(w_in[i] - mean);
    REAL pr = (REAL)(temp * d2);
    stdev = (REAL)(stdev + pr);
    sum_x += i;
    sum_x2 += i * i;
    sum_xy += (double)w_in[i] * (double)i;
    sum_y += (double)w_in[i];
  }
  stdev = (REAL)(stdev / (REAL)(n - 1));
  stdev = (REAL)SQRT_REAL(stdev);
  *mean_o = mean;
  *stdev_o = stdev;
  REAL slope = (REAL)(((double)n * sum_xy - (double)sum_x * sum_y) /
                      (double)(n * sum_x2 - sum_x * sum_x));
  *slope_o = slope;
  *intercept_o = (REAL)((sum_y - (double)sum_x * (double)slope) / (double)n);
  return 0;
}

/* linear_slope_fit.py:93-158 */
static int FN(orc1_linear_slope_diff)(const REAL *w_in, int64_t n, REAL slope, REAL intercept,
                                      REAL *mean_o, REAL *rms_o) {
  *mean_o = *rms_o = (REAL)NAN;
  if (FN(any_nan)(w_in, n) || isnan(slope) || isnan(intercept)) return 0;
  REAL mean = 0, rms = 0;
  for (int64_t i = 0; i < n; i++) {
    /* slope*i: REAL*int64 -> float64 */
    double temp = (double)w_in[i] - ((double)slope * (double)i + (double)intercept);
    /* float64 scalars are rounded to REAL before the in-place array add */
    mean = (REAL)(mean + (REAL)(temp / (double)(i + 1)));
    rms = (REAL)(rms + (REAL)(temp * temp));
  }
  rms = (REAL)(rms / (REAL)(n - 1));
  rms = (REAL)SQRT_REAL(rms);
  *mean_o = mean;
  *rms_o = rms;
  return 0;
}

/* arithmetic.py:9-62 */
static int FN(orc1_mean_below_threshold)(const REAL *w_in, int64_t n, REAL thr, REAL *res) {
  *res = (REAL)NAN;
  if (FN(any_nan)(w_in, n) || isnan(thr)) return 0;
  double total = 0.0;
  int64_t count = 0;
  for (int64_t i = 0; i < n; i++)
    if (w_in[i] < thr) {
      total += (double)w_in[i];
      count++;
    }
  if (count == 0) return 0;
  *res = (REAL)(total / (double)count);
  return 0;
}

/* pole_zero.py:24-77: float64 state, REAL store; "-1 / t_tau" is int/REAL ->
 * float64 so the exponential is evaluated in double. */
static int FN(orc1_pole_zero)(const REAL *w_in, int64_t n, REAL t_tau, REAL *w_out) {
  FN(fill_nan)(w_out, n);
  if (FN(any_nan)(w_in, n) || isnan(t_tau)) return 0;
  double c = exp(-1.0 / (double)t_tau);
  double prev = (double)w_in[0];
  w_out[0] = w_in[0];
  for (int64_t i = 1; i < n; i++) {
    double prod = (double)w_in[i - 1] * c;
    double cur = (prev + (double)w_in[i]) - prod;
    w_out[i] = (REAL)cur;
    prev = cur;
  }
  if (FN(any_nan)(w_out, n)) return ORC_FATAL_PZ_NAN;
  return 0;
}

/* pole_zero.py:82-198 */
static int FN(orc1_double_pole_zero)(const REAL *w_in, int64_t n, REAL t_tau1, REAL t_tau2,
                                     REAL frac, REAL *w_out) {
  FN(fill_nan)(w_out, n);
  if (FN(any_nan)(w_in, n) || isnan(t_tau1) || isnan(t_tau2) || isnan(frac)) return 0;
  if (n <= 3) return ORC_FATAL_DPZ_SHORT;
  double a = exp(-1.0 / (double)t_tau1);
  double b = exp(-1.0 / (double)t_tau2);
  double f = (double)frac;
  double den1 = f * b - f * a - b - 1.0;
  double den2 = -1.0 * (f * b - f * a - b);
  double num1 = -1.0 * (a + b);
  double num2 = a * b;
  double t0 = (double)w_in[0], t1 = (double)w_in[1];
  w_out[0] = w_in[0];
  w_out[1] = w_in[1];
  for (int64_t i = 2; i < n; i++) {
    double t2 = (double)w_in[i] + num1 * (double)w_in[i - 1];
    t2 = t2 + num2 * (double)w_in[i - 2];
    t2 = t2 - den1 * t1;
    t2 = t2 - den2 * t0;
    w_out[i] = (REAL)t2;
    t0 = t1;
    t1 = t2;
  }
  return 0;
}

static int FN(trap_check)(int64_t n, int32_t rise, int32_t flat) {
  if (rise < 0) return ORC_FATAL_RISE_NEG;
  if (flat < 0) return ORC_FATAL_FLAT_NEG;
  if (2 * (int64_t)rise + flat > n) return ORC_FATAL_TRAP_WIDE;
  return 0;
}

/* trap_filters.py:12-76: pure REAL running sum, left-to-right adds */
static int FN(orc1_trap_filter)(const REAL *w_in, int64_t n, int32_t rise, int32_t flat,
                                REAL *w_out) {
  FN(fill_nan)(w_out, n);
  if (FN(any_nan)(w_in, n)) return 0;
  int rc = FN(trap_check)(n, rise, flat);
  if (rc) return rc;
  w_out[0] = w_in[0];
  int64_t i;
  for (i = 1; i < rise; i++) w_out[i] = (REAL)(w_out[PREV(i)] + w_in[i]);
  for (i = rise; i < rise + flat; i++) {
    REAL t = (REAL)(w_out[PREV(i)] + w_in[i]);
    w_out[i] = (REAL)(t - w_in[i - rise]);
  }
  for (i = rise + flat; i < 2 * rise + flat; i++) {
    REAL t = (REAL)(w_out[PREV(i)] + w_in[i]);
    t = (REAL)(t - w_in[i - rise]);
    w_out[i] = (REAL)(t - w_in[i - rise - flat]);
  }
  for (i = 2 * rise + flat; i < n; i++) {
    REAL t = (REAL)(w_out[PREV(i)] + w_in[i]);
    t = (REAL)(t - w_in[i - rise]);
    t = (REAL)(t - w_in[i - rise - flat]);
    w_out[i] = (REAL)(t + w_in[i - 2 * rise - flat]);
  }
  return 0;
}

/* trap_filters.py:79-149: the division by the int32 rise promotes every step
 * to float64; the store rounds to REAL. */
static int FN(orc1_trap_norm)(const REAL *w_in, int64_t n, int32_t rise, int32_t flat,
                              REAL *w_out) {
  FN(fill_nan)(w_out, n);
  if (FN(any_nan)(w_in, n)) return 0;
  int rc = FN(trap_check)(n, rise, flat);
  if (rc) return rc;
  double r = (double)rise;
  w_out[0] = (REAL)((double)w_in[0] / r);
  int64_t i;
  for (i = 1; i < rise; i++) w_out[i] = (REAL)((double)w_out[PREV(i)] + (double)w_in[i] / r);
  for (i = rise; i < rise + flat; i++) {
    REAL d = (REAL)(w_in[i] - w_in[i - rise]);
    w_out[i] = (REAL)((double)w_out[PREV(i)] + (double)d / r);
  }
  for (i = rise + flat; i < 2 * rise + flat; i++) {
    REAL d = (REAL)(w_in[i] - w_in[i - rise]);
    d = (REAL)(d - w_in[i - rise - flat]);
    w_out[i] = (REAL)((double)w_out[PREV(i)] + (double)d / r);
  }
  for (i = 2 * rise + flat; i < n; i++) {
    REAL d = (REAL)(w_in[i] - w_in[i - rise]);
    d = (REAL)(d - w_in[i - rise - flat]);
    d = (REAL)(d + w_in[i - 2 * rise - flat]);
    w_out[i] = (REAL)((double)w_out[PREV(i)] + (double)d / r);
  }
  return 0;
}

/* trap_filters.py:152-227 */
static int FN(orc1_asym_trap_filter)(const REAL *w_in, int64_t n, int32_t rise, int32_t flat,
                                     int32_t fall, REAL *w_out) {
  FN(fill_nan)(w_out, n);
  if (FN(any_nan)(w_in, n)) return 0;
  if (rise < 0) return ORC_FATAL_RISE_NEG;
  if (flat < 0) return ORC_FATAL_FLAT_NEG;
  if (fall < 0) return ORC_FATAL_FALL_NEG;
  if ((int64_t)rise + flat + fall > n) return ORC_FATAL_TRAP_WIDE;
  double r = (double)rise, fl = (double)fall;
  w_out[0] = (REAL)((double)w_in[0] / r);
  int64_t i;
  for (i = 1; i < rise; i++) w_out[i] = (REAL)((double)w_out[PREV(i)] + (double)w_in[i] / r);
  for (i = rise; i < rise + flat; i++) {
    REAL d = (REAL)(w_in[i] - w_in[i - rise]);
    w_out[i] = (REAL)((double)w_out[PREV(i)] + (double)d / r);
  }
  for (i = rise + flat; i < rise + flat + fall; i++) {
    REAL d = (REAL)(w_in[i] - w_in[i - rise]);
    double t = (double)w_out[PREV(i)] + (double)d / r;
    w_out[i] = (REAL)(t - (double)w_in[i - rise - flat] / fl);
  }
  for (i = rise + flat + fall; i < n; i++) {
    REAL d = (REAL)(w_in[i] - w_in[i - rise]);
    REAL e = (REAL)(w_in[i - rise - flat] - w_in[i - rise - flat - fall]);
    double t = (double)w_out[PREV(i)] + (double)d / r;
    w_out[i] = (REAL)(t - (double)e / fl);
  }
  return 0;
}

/* trap_filters.py:230-301: float64 window sums */
static int FN(orc1_trap_pickoff)(const REAL *w_in, int64_t n, int32_t rise, int32_t flat,
                                 REAL t_pickoff, REAL *a_out) {
  *a_out = (REAL)NAN;
  if (FN(any_nan)(w_in, n) || isnan(t_pickoff)) return 0;
  if (FLOOR_REAL(t_pickoff) != t_pickoff) return ORC_FATAL_PICKOFF_NONINT;
  int rc = FN(trap_check)(n, rise, flat);
  if (rc) return rc;
  double i_1 = 0.0, i_2 = 0.0;
  int64_t start = (int64_t)(t_pickoff + (REAL)1);
  if (!(n >= start && start >= 2 * (int64_t)rise + flat)) return 0;
  for (int64_t i = start - rise; i < start; i++) i_1 += (double)w_in[i];
  for (int64_t i = start - 2 * rise - flat; i < start - rise - flat; i++) i_2 += (double)w_in[i];
  *a_out = (REAL)((i_1 - i_2) / (double)rise);
  return 0;
}

/* moving_windows.py:12-61 (all REAL arithmetic; length is a REAL) */
static void FN(mw_left_core)(const REAL *w_in, int64_t n, REAL length, REAL *w_out) {
  int64_t L = (int64_t)length;
  w_out[0] = w_in[0];
  int64_t i;
  for (i = 1; i < L; i++) {
    REAL d = (REAL)(w_in[i] - w_in[0]);
    d = (REAL)(d / length);
    w_out[i] = (REAL)(w_out[PREV(i)] + d);
  }
  for (i = (L > 0 ? L : 0); i < n; i++) {
    REAL d = (REAL)(w_in[i] - w_in[i - L]);
    d = (REAL)(d / length);
    w_out[i] = (REAL)(w_out[PREV(i)] + d);
  }
}
/* moving_windows.py:64-114 */
static void FN(mw_right_core)(const REAL *w_in, int64_t n, REAL length, REAL *w_out) {
  int64_t L = (int64_t)length;
  w_out[n - 1] = w_in[n - 1];
  int64_t i;
  for (i = 1; i < L; i++) {
    REAL d = (REAL)(w_in[n - 1 - i] - w_out[n - 1]);
    d = (REAL)(d / length);
    w_out[n - 1 - i] = (REAL)(w_out[n - i] + d);
  }
  for (i = (L > 0 ? L : 1); i < n; i++) { /* L == 0: the reference reads w_out[n] (out of bounds) */
    REAL d = (REAL)(w_in[n - 1 - i] - w_in[n - 1 - i + L]);
    d = (REAL)(d / length);
    w_out[n - 1 - i] = (REAL)(w_out[n - i] + d);
  }
}
static int FN(orc1_moving_window_left)(const REAL *w_in, int64_t n, REAL length, REAL *w_out) {
  FN(fill_nan)(w_out, n);
  if (FN(any_nan)(w_in, n)) return 0;
  if (!(length >= 0) || !(length < (REAL)n)) return ORC_FATAL_MW_RANGE;
  FN(mw_left_core)(w_in, n, length, w_out);
  return 0;
}
static int FN(orc1_moving_window_right)(const REAL *w_in, int64_t n, REAL length, REAL *w_out) {
  FN(fill_nan)(w_out, n);
  if (FN(any_nan)(w_in, n)) return 0;
  if (!(length >= 0) || !(length < (REAL)n)) return ORC_FATAL_MW_RANGE;
  FN(mw_right_core)(w_in, n, length, w_out);
  return 0;
}
/* moving_windows.py:117-203 ; scratch must hold n REALs */
static int FN(orc1_moving_window_multi)(const REAL *w_in, int64_t n, REAL length, REAL num_mw,
                                        int32_t mw_type, REAL *w_out, REAL *scratch) {
  FN(fill_nan)(w_out, n);
  if (FN(any_nan)(w_in, n)) return 0;
  if (FLOOR_REAL(length) != length) return ORC_FATAL_MWM_LEN_NONINT;
  if (FLOOR_REAL(num_mw) != num_mw) return ORC_FATAL_MWM_NUM_NONINT;
  if ((int64_t)length < 0 || (int64_t)length >= n) return ORC_FATAL_MWM_RANGE;
  if ((int64_t)num_mw < 0) return ORC_FATAL_MWM_NUM_NEG;
  memcpy(scratch, w_in, (size_t)n * sizeof(REAL));
  int64_t nm = (int64_t)num_mw;
  for (int64_t k = 0; k < nm; k++) {
    if (((k % 2 == 1) && (mw_type == 0)) || (mw_type == 2))
      FN(mw_right_core)(scratch, n, length, w_out);
    else
      FN(mw_left_core)(scratch, n, length, w_out);
    memcpy(scratch, w_out, (size_t)n * sizeof(REAL));
  }
  return 0;
}
/* moving_windows.py:206-249 ; w_out has n_out samples */
static int FN(orc1_avg_current)(const REAL *w_in, int64_t n, REAL length, REAL *w_out,
                                int64_t n_out) {
  FN(fill_nan)(w_out, n_out);
  if (FN(any_nan)(w_in, n)) return 0;
  if (!(length >= 0) || !(length < (REAL)n)) return ORC_FATAL_MW_RANGE;
  int64_t L = (int64_t)length;
  if (n_out != n - L) return ORC_FATAL_SHAPE;
  for (int64_t i = 0; i < n_out; i++) {
    REAL d = (REAL)(w_in[i + L] - w_in[i]);
    w_out[i] = (REAL)(d / length);
  }
  return 0;
}

/* time_point_thresh.py:12-92 */
static int FN(orc1_time_point_thresh)(const REAL *w, int64_t n, REAL thr, REAL t_start,
                                      REAL walk_forward, REAL *t_out) {
  *t_out = (REAL)NAN;
  if (FN(any_nan)(w, n) || isnan(thr) || isnan(t_start) || isnan(walk_forward)) return 0;
  if (FLOOR_REAL(t_start) != t_start) return ORC_FATAL_TSTART_NONINT;
  if (FLOOR_REAL(walk_forward) != walk_forward) return ORC_FATAL_WALK_NONINT;
  int64_t s = (int64_t)t_start;
  if (s < 0 || s >= n) return ORC_FATAL_TSTART_RANGE;
  if ((int64_t)walk_forward == 1) {
    for (int64_t i = s; i < n - 1; i++)
      if ((w[i] <= thr && thr < w[i + 1]) || (w[i] >= thr && thr > w[i + 1])) {
        *t_out = (REAL)i;
        return 0;
      }
  } else {
    for (int64_t i = s; i > 0; i--)
      if ((w[i - 1] < thr && thr <= w[i]) || (w[i - 1] > thr && thr >= w[i])) {
        *t_out = (REAL)i;
        return 0;
      }
  }
  return 0;
}

/* time_point_thresh.py:95-222 */
static int FN(orc1_interpolated_time_point_thresh)(const REAL *w, int64_t n, REAL thr,
                                                   REAL t_start, int64_t walk_forward,
                                                   int8_t mode, REAL *t_out) {
  *t_out = (REAL)NAN;
  if (FN(any_nan)(w, n) || isnan(thr) || isnan(t_start)) return 0;
  if (t_start < 0 || t_start >= (REAL)n) return 0;
  int64_t ic = -1;
  int64_t s = (int64_t)t_start;
  if (walk_forward > 0) {
    for (int64_t i = s; i < n - 1; i++)
      if ((w[i] <= thr && thr < w[i + 1]) || (w[i] >= thr && thr > w[i + 1])) {
        ic = i;
        break;
      }
  } else {
    for (int64_t i = s; i > 1; i--)
      if ((w[i - 1] < thr && thr <= w[i]) || (w[i - 1] > thr && thr >= w[i])) {
        ic = i - 1;
        break;
      }
  }
  if (ic == -1) return 0;
  switch (mode) {
    case 'i': *t_out = (REAL)ic; break;
    case 'a': case 'f': *t_out = (REAL)(ic + 1); break;
    case 'b': case 'c': *t_out = (REAL)ic; break;
    case 'r':
      if (ABS_REAL((REAL)(thr - w[ic])) < ABS_REAL((REAL)(thr - w[ic + 1])))
        *t_out = (REAL)ic;
      else
        *t_out = (REAL)(ic + 1);
      break;
    case 'n': *t_out = (REAL)((double)ic + 0.5); break;
    case 'l': {
      REAL num = (REAL)(thr - w[ic]);
      REAL den = (REAL)(w[ic + 1] - w[ic]);
      REAL q = (REAL)(num / den);
      *t_out = (REAL)((double)ic + (double)q);
      break;
    }
    default: return ORC_FATAL_INTERP_MODE;
  }
  return 0;
}

/* time_point_thresh.py:225-401.  Negative indices wrap as in Python. */
static REAL FN(wrap_get)(const REAL *w, int64_t n, int64_t i) { return w[i < 0 ? i + n : i]; }
static int FN(mtp_set)(REAL *t_out, int64_t idx, const REAL *w, int64_t n, const REAL *thr,
                       int64_t i_wf, int64_t pol, int8_t mode) {
  switch (mode) {
    case 'i': t_out[idx] = (REAL)i_wf; break;
    case 'a': case 'f': t_out[idx] = (REAL)(pol < 0 ? i_wf : i_wf + 1); break;
    case 'b': case 'c': t_out[idx] = (REAL)(pol > 0 ? i_wf : i_wf - 1); break;
    case 'r':
      if ((REAL)(thr[idx] - FN(wrap_get)(w, n, i_wf)) <
          (REAL)(FN(wrap_get)(w, n, i_wf + pol) - thr[idx]))
        t_out[idx] = (REAL)i_wf;
      else
        t_out[idx] = (REAL)(i_wf + pol);
      break;
    case 'n': t_out[idx] = (REAL)((double)i_wf + 0.5 * (double)pol); break;
    case 'l': {
      REAL num = (REAL)(thr[idx] - FN(wrap_get)(w, n, i_wf));
      REAL den = (REAL)(FN(wrap_get)(w, n, i_wf + pol) - FN(wrap_get)(w, n, i_wf));
      *(&t_out[idx]) = (REAL)((double)i_wf + (double)(REAL)(num / den));
      break;
    }
    default: return ORC_FATAL_INTERP_MODE;
  }
  return 0;
}
static int FN(orc1_multi_time_point_thresh)(const REAL *w, int64_t n, const REAL *thr, int64_t m,
                                            REAL t_start_r, REAL polarity, int8_t mode,
                                            REAL *t_out, int64_t *sorted /* m scratch */) {
  FN(fill_nan)(t_out, m);
  if (FN(any_nan)(w, n) || FN(any_nan)(thr, m) || isnan(t_start_r)) return 0;
  if (t_start_r < 0 || t_start_r >= (REAL)n) return 0;
  int64_t pol;
  if (polarity > 0) pol = 1;
  else if (polarity < 0) pol = -1;
  else return ORC_FATAL_POLARITY_ZERO;
  /* stable insertion argsort */
  for (int64_t i = 0; i < m; i++) sorted[i] = i;
  for (int64_t i = 1; i < m; i++) {
    int64_t k = sorted[i], j = i - 1;
    while (j >= 0 && thr[sorted[j]] > thr[k]) { sorted[j + 1] = sorted[j]; j--; }
    sorted[j + 1] = k;
  }
  int64_t t_start = (int64_t)t_start_r;
  REAL a_start = w[t_start];
  int64_t i_start = m;
  for (int64_t i = 0; i < m; i++)
    if (thr[sorted[i]] >= a_start) { i_start = i; break; }
  int64_t i_tp = i_start;
  if (i_tp < m) {
    int64_t idx = sorted[i_tp];
    int64_t stop = pol > 0 ? n - 1 : -1;
    for (int64_t i_wf = t_start; pol > 0 ? i_wf < stop : i_wf > stop; i_wf += pol) {
      if (i_tp >= m) break;
      while (FN(wrap_get)(w, n, i_wf) <= thr[idx] && thr[idx] < FN(wrap_get)(w, n, i_wf + pol)) {
        int rc = FN(mtp_set)(t_out, idx, w, n, thr, i_wf, pol, mode);
        if (rc) return rc;
        i_tp++;
        if (i_tp >= m) break;
        idx = sorted[i_tp];
      }
    }
  }
  i_tp = i_start - 1;
  if (i_tp >= 0) {
    int64_t idx = sorted[i_tp];
    int64_t stop = pol < 0 ? n - 1 : -1;
    int64_t step = -pol;
    for (int64_t i_wf = t_start - 1; step > 0 ? i_wf < stop : i_wf > stop; i_wf += step) {
      if (i_tp < 0) break;
      while (FN(wrap_get)(w, n, i_wf) <= thr[idx] && thr[idx] < FN(wrap_get)(w, n, i_wf + pol)) {
        int rc = FN(mtp_set)(t_out, idx, w, n, thr, i_wf, pol, mode);
        if (rc) return rc;
        i_tp--;
        if (i_tp < 0) break;
        idx = sorted[i_tp];
      }
    }
  }
  return 0;
}

/* fixed_time_pickoff.py:12-125 ; scratch = 2*n doubles (spline mode only) */
static int FN(orc1_fixed_time_pickoff)(const REAL *w, int64_t n, REAL t_in, int8_t mode,
                                       REAL *a_out, double *scratch) {
  *a_out = (REAL)NAN;
  if (FN(any_nan)(w, n) || isnan(t_in)) return 0;
  if (t_in < 0 || t_in > (REAL)(n - 1)) return 0;
  int64_t i_in = (int64_t)t_in;
  if ((REAL)i_in == t_in) {
    *a_out = w[i_in];
    return 0;
  }
  double t0 = (double)t_in - (double)i_in;
  double t1 = 1.0 - t0;
  switch (mode) {
    case 'i': return ORC_FATAL_FTP_INT;
    case 'n': *a_out = t0 < 0.5 ? w[i_in] : w[i_in + 1]; break;
    case 'f': *a_out = w[i_in]; break;
    case 'c': *a_out = w[i_in + 1]; break;
    case 'l': *a_out = (REAL)(t1 * (double)w[i_in] + t0 * (double)w[i_in + 1]); break;
    case 'h': {
      double m0 = i_in == 0 ? (double)(REAL)(w[1] - w[0])
                            : (double)(REAL)(w[i_in + 1] - w[i_in - 1]) / 2.0;
      double m1 = i_in == n - 2 ? (double)(REAL)(w[n - 1] - w[n - 2])
                                : (double)(REAL)(w[i_in + 2] - w[i_in]) / 2.0;
      double t1_2 = t1 * t1, t1_3 = t1_2 * t1, t0_2 = t0 * t0, t0_3 = t0_2 * t0;
      double r = (-2.0 * t1_3 + 3.0 * t1_2) * (double)w[i_in];
      r = r + (-2.0 * t0_3 + 3.0 * t0_2) * (double)w[i_in + 1];
      r = r - (t1_3 - t1_2) * m0;
      r = r + (t0_3 - t0_2) * m1;
      *a_out = (REAL)r;
      break;
    }
    case 's': {
      double *u = scratch, *w2 = scratch + n;
      for (int64_t i = 0; i < n; i++) u[i] = w2[i] = 0.0;
      for (int64_t i = 1; i < n - 1; i++) {
        double p = 0.5 * w2[i - 1] + 2.0;
        w2[i] = -0.5 / p;
        /* 2 * w[i] is int64 * REAL -> float64 in numba */
        u[i] = ((double)w[i + 1] - 2.0 * (double)w[i]) + (double)w[i - 1];
        u[i] = (3.0 * u[i] - 0.5 * u[i - 1]) / p;
      }
      for (int64_t i = n - 2; i > i_in - 1; i--) w2[i] = w2[i] * w2[i + 1] + u[i];
      double t1_3 = t1 * t1 * t1, t0_3 = t0 * t0 * t0;
      double r = t1 * (double)w[i_in] + t0 * (double)w[i_in + 1];
      r = r + ((t1_3 - t1) * w2[i_in] + (t0_3 - t0) * w2[i_in + 1]) / 6.0;
      *a_out = (REAL)r;
      break;
    }
    default: return ORC_FATAL_INTERP_MODE;
  }
  return 0;
}

/* min_max.py:11-82 */
static int FN(orc1_min_max)(const REAL *w, int64_t n, REAL *t_min, REAL *t_max, REAL *a_min,
                            REAL *a_max) {
  *t_min = *t_max = *a_min = *a_max = (REAL)NAN;
  if (FN(any_nan)(w, n)) return 0;
  int64_t imin = 0, imax = 0;
  for (int64_t i = 0; i < n; i++) {
    if (w[i] < w[imin]) imin = i;
    if (w[i] > w[imax]) imax = i;
  }
  *a_min = w[imin];
  *a_max = w[imax];
  *t_min = (REAL)imin;
  *t_max = (REAL)imax;
  return 0;
}

/* min_max.py:85-140 */
static int FN(orc1_min_max_norm)(const REAL *w, int64_t n, REAL a_min, REAL a_max, REAL *w_out) {
  FN(fill_nan)(w_out, n);
  if (FN(any_nan)(w, n)) return 0;
  REAL amx = ABS_REAL(a_max), amn = ABS_REAL(a_min);
  if (amx == 0 || amn == 0) {
    for (int64_t i = 0; i < n; i++) w_out[i] = w[i];
  } else if (amx >= amn) {
    for (int64_t i = 0; i < n; i++) w_out[i] = (REAL)(w[i] / amx);
  } else if (amx < amn) {
    for (int64_t i = 0; i < n; i++) w_out[i] = (REAL)(w[i] / amn);
  }
  return 0;
}

/* windower.py:12-54 */
static int FN(orc1_windower)(const REAL *w_in, int64_t n, REAL t0_in, REAL *w_out, int64_t m) {
  FN(fill_nan)(w_out, m);
  if (FN(any_nan)(w_in, n) || isnan(t0_in)) return 0;
  if (m >= n) return ORC_FATAL_WINDOWER_LEN;
  int64_t beg = (int64_t)t0_in;
  if (beg > n) beg = n;
  int64_t end = beg + m;
  if (end < 0) end = 0;
  if (beg < 0) {
    for (int64_t i = 0; i < end; i++) w_out[m - end + i] = w_in[i];
  } else if (end < n) {
    for (int64_t i = 0; i < m; i++) w_out[i] = w_in[beg + i];
  } else {
    for (int64_t i = 0; i < n - beg; i++) w_out[i] = w_in[beg + i];
  }
  return 0;
}

/* upsampler.py:14-49 */
static int FN(orc1_upsampler)(const REAL *w_in, int64_t n, REAL upsample, REAL *w_out, int64_t m) {
  FN(fill_nan)(w_out, m);
  if (FN(any_nan)(w_in, n)) return 0;
  if (!(upsample > 0)) return ORC_FATAL_UPSAMPLE;
  double half = floor((double)upsample / 2.0);
  int64_t reps = (int64_t)upsample;
  for (int64_t t_in = 0; t_in < n; t_in++) {
    int64_t t_out = (int64_t)((double)t_in * (double)upsample - half);
    for (int64_t k = 0; k < reps; k++) {
      if (t_out >= 0 && t_out < m) w_out[t_out] = w_in[t_in];
      t_out++;
    }
  }
  return 0;
}

/* numpy.convolve(a, v, mode) as used by convolutions.py:72,118,180 and
 * energy_kernels.py:73,157.  Accumulates in float64 (the reference accumulates
 * in REAL through numpy's dot / pocketfft; see DESIGN.md "tolerances"), rounds
 * the result to REAL.  full: len n+m-1; same: len max(n,m), offset
 * (min(n,m)-1)/2 into full; valid: len max(n,m)-min(n,m)+1, offset min(n,m)-1. */
static void FN(conv_core)(const REAL *a, int64_t n, const REAL *v, int64_t m, int8_t mode,
                          REAL *out) {
  int64_t lo = n < m ? n : m, hi = n < m ? m : n;
  int64_t off, len;
  if (mode == 'f') { off = 0; len = n + m - 1; }
  else if (mode == 's') { off = (lo - 1) / 2; len = hi; }
  else { off = lo - 1; len = hi - lo + 1; }
  for (int64_t k = 0; k < len; k++) {
    int64_t kk = k + off; /* index into full */
    int64_t j0 = kk - (m - 1) > 0 ? kk - (m - 1) : 0;
    int64_t j1 = kk < n - 1 ? kk : n - 1;
    double acc = 0.0;
    for (int64_t j = j0; j <= j1; j++) acc += (double)a[j] * (double)v[kk - j];
    out[k] = (REAL)acc;
  }
}

/* convolutions.py:14-72 (also the arithmetic of fft_convolve_wf :75-119) */
static int FN(orc1_convolve_wf)(const REAL *w_in, int64_t n, const REAL *kern, int64_t m,
                                int8_t mode, REAL *w_out, int64_t p) {
  FN(fill_nan)(w_out, p);
  if (FN(any_nan)(w_in, n)) return 0;
  if (FN(any_nan)(kern, m)) return 0;
  if (m > n) return ORC_FATAL_CONV_KERNEL_LONG;
  int64_t expect;
  if (mode == 'f') expect = n + m - 1;
  else if (mode == 'v') expect = n - m + 1;
  else if (mode == 's') expect = n;
  else return ORC_FATAL_CONV_MODE;
  if (p != expect) return ORC_FATAL_CONV_OUTLEN;
  FN(conv_core)(w_in, n, kern, m, mode, w_out);
  return 0;
}

/* get_multi_local_extrema.py:12-306.  scratch: 4*m doubles. */
static int FN(dcmp)(const void *a, const void *b) {
  double x = *(const double *)a, y = *(const double *)b;
  if (isnan(x)) return isnan(y) ? 0 : 1;
  if (isnan(y)) return -1;
  return x < y ? -1 : (x > y ? 1 : 0);
}
static int64_t FN(gmle_and)(const double *left, int64_t n_left, const double *right_sorted,
                            int64_t n_right, REAL *out) {
  /* coincidences: left entries (in order) that also appear in right_sorted */
  if (n_left <= 0 || n_right <= 0) return 0;
  int64_t cnt = 0;
  for (int64_t i = 0; i < n_left; i++) {
    int64_t v = (int64_t)left[i];
    for (int64_t j = 0; j < n_right; j++)
      if ((int64_t)right_sorted[j] == v) { out[cnt++] = (REAL)v; break; }
  }
  return cnt;
}
static int64_t FN(gmle_or)(const double *left, const double *right, int64_t m, REAL *out,
                           double *tmp /* 2m */) {
  for (int64_t i = 0; i < m; i++) { tmp[i] = left[i]; tmp[m + i] = right[i]; }
  qsort(tmp, (size_t)(2 * m), sizeof(double), FN(dcmp));
  /* unique, NaNs collapse to one trailing NaN */
  int64_t nu = 0;
  for (int64_t i = 0; i < 2 * m; i++) {
    if (nu > 0) {
      double p = tmp[nu - 1];
      if ((isnan(p) && isnan(tmp[i])) || p == tmp[i]) continue;
    }
    tmp[nu++] = tmp[i];
  }
  int64_t cnt = 0;
  int64_t lim = m <= nu ? m : nu;
  for (int64_t i = 0; i < lim; i++) {
    out[i] = (REAL)tmp[i];
    if (!isnan(tmp[i])) cnt++;
  }
  return cnt;
}
static int FN(orc1_get_multi_local_extrema)(const REAL *w, int64_t n, REAL d_max, REAL d_min,
                                            REAL search_direction, REAL abs_max, REAL abs_min,
                                            REAL *vt_max, REAL *vt_min, int64_t m,
                                            uint32_t *n_max_out, uint32_t *n_min_out,
                                            double *scratch /* 6*m doubles */) {
  FN(fill_nan)(vt_max, m);
  FN(fill_nan)(vt_min, m);
  *n_max_out = 0;
  *n_min_out = 0;
  double *l_max = scratch, *l_min = scratch + m, *r_max = scratch + 2 * m,
         *r_min = scratch + 3 * m, *tmp = scratch + 4 * m;
  for (int64_t i = 0; i < m; i++) l_max[i] = l_min[i] = r_max[i] = r_min[i] = NAN;
  int64_t nl_max = 0, nl_min = 0, nr_max = 0, nr_min = 0;
  if (FN(any_nan)(w, n) || isnan(d_max) || isnan(d_min)) return 0;
  if (!(m < n)) return ORC_FATAL_GMLE_LEN;
  if (!(d_max >= 0) || !(d_min >= 0)) return ORC_FATAL_GMLE_DELTA;
  if (search_direction == 0 || search_direction > 1) {
    int find_max = 1;
    int64_t imax = 0, imin = 0;
    for (int64_t i = 0; i < n; i++) {
      if (w[i] > w[imax]) imax = i;
      if (w[i] < w[imin]) imin = i;
      if (find_max) {
        if (w[i] < (REAL)(w[imax] - d_max) && nl_max < m && w[imax] > abs_max) {
          l_max[nl_max++] = (double)imax;
          imin = i;
          find_max = 0;
        }
      } else {
        if (w[i] > (REAL)(w[imin] + d_min) && nl_min < m && w[imin] < abs_min) {
          l_min[nl_min++] = (double)imin;
          imax = i;
          find_max = 1;
        }
      }
    }
  }
  if (search_direction > 0) {
    int find_max = 1;
    int64_t imax = n - 1, imin = n - 1;
    for (int64_t i = n - 1; i >= 0; i--) {
      if (w[i] > w[imax]) imax = i;
      if (w[i] < w[imin]) imin = i;
      if (find_max) {
        if (w[i] < (REAL)(w[imax] - d_max) && nr_max < m && w[imax] > abs_max) {
          r_max[nr_max++] = (double)imax;
          imin = i;
          find_max = 0;
        }
      } else {
        if (w[i] > (REAL)(w[imin] + d_min) && nr_min < m && w[imin] < abs_min) {
          r_min[nr_min++] = (double)imin;
          imax = i;
          find_max = 1;
        }
      }
    }
  }
  if (search_direction == 0) {
    *n_max_out = (uint32_t)nl_max;
    *n_min_out = (uint32_t)nl_min;
    for (int64_t i = 0; i < m; i++) { vt_max[i] = (REAL)l_max[i]; vt_min[i] = (REAL)l_min[i]; }
  } else if (search_direction == 1) {
    *n_max_out = (uint32_t)nr_max;
    *n_min_out = (uint32_t)nr_min;
    for (int64_t i = 0; i < m; i++) { vt_max[i] = (REAL)r_max[i]; vt_min[i] = (REAL)r_min[i]; }
  } else if (search_direction == 2) {
    qsort(r_max, (size_t)m, sizeof(double), FN(dcmp));
    qsort(r_min, (size_t)m, sizeof(double), FN(dcmp));
    *n_max_out = (uint32_t)FN(gmle_and)(l_max, nl_max, r_max, nr_max, vt_max);
    /* reference :255-256 masks the *max* lists with the NaN pattern of the min
     * lists: i.e. it takes the first n_min entries of each max list. */
    *n_min_out = (uint32_t)FN(gmle_and)(l_max, nl_min, r_max, nr_min, vt_min);
  } else if (search_direction == 3) {
    *n_max_out = (uint32_t)FN(gmle_or)(l_max, r_max, m, vt_max, tmp);
    *n_min_out = (uint32_t)FN(gmle_or)(l_min, r_min, m, vt_min, tmp);
  } else {
    return ORC_FATAL_GMLE_DIR;
  }
  return 0;
}

/* recursive_filter.py:12-93 ; circ = scratch of q doubles */
static int FN(orc1_recursive_filter)(const REAL *w_in, int64_t n, const double *a, int64_t p,
                                     const double *b, int64_t q, REAL init_in, REAL init_out,
                                     REAL *w_out, double *circ) {
  FN(fill_nan)(w_out, n);
  if (FN(any_nan)(w_in, n) || isnan(init_in) || isnan(init_out)) return 0;
  for (int64_t j = 0; j < p; j++) if (isnan(a[j])) return 0;
  for (int64_t j = 0; j < q; j++) if (isnan(b[j])) return 0;
  if (q == 0) return ORC_FATAL_RF_B_SCALAR;
  if (n <= q) return ORC_FATAL_RF_SHORT;
  for (int64_t j = 0; j < q; j++) circ[j] = (double)init_out;
  for (int64_t i = 0; i < n; i++) {
    int64_t ib = i % q;
    circ[ib] = 0;
    for (int64_t j = 0; j < p; j++) {
      if (j <= i) circ[ib] += a[j] * (double)w_in[i - j];
      else circ[ib] += a[j] * (double)init_in;
    }
    for (int64_t j = 1; j < q; j++) {
      int64_t k = ib - j;
      if (k < 0) k += q;
      circ[ib] -= b[j] * circ[k];
    }
    circ[ib] /= b[0];
    w_out[i] = (REAL)circ[ib];
  }
  return 0;
}

#undef PREV
#undef FN
#undef CAT
#undef CAT_
