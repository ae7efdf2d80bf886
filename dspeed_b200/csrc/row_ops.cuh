// dspeed_b200 -- block-wide processor routines over shared-memory resident waveforms.
//
// Every routine restates one dspeed processor (reference file:line cited at each
// one) in a parallel form: what the reference evaluates as a sequential float32
// recursion is evaluated here as float64 prefix sums / reductions and rounded once to
// the output dtype, so results agree with the reference to its own rounding drift
// (<= 1e-5 of the waveform maximum; indices, extrema and copies are bit-exact).
//
// Conventions: `in`/`out` are shared-memory slots in the padded layout (sidx);
// all NT threads call every routine; routines end with the data visible to the whole
// CTA (they finish with a barrier when they write a slot).  NaN propagation
// ("any NaN in -> all NaN out") is handled by the caller through per-slot flags.
#pragma once
#include "common.cuh"
#include "dspeed_b200.h"

namespace dspb {

// ---------------------------------------------------------------------------------
// generic cumulative sums:  out[i] = sum_{j<=i} d(j)   (forward)
//                           out[i] = sum_{j>=i} d(j)   (reverse)
// d(j) is evaluated twice (once for the chunk totals, once for the running sums);
// accumulation is float64, `post` maps the float64 running sum to the stored value.
// Returns (per thread) 1 if a NaN was stored.
// ---------------------------------------------------------------------------------
// When a thread's chunk has at most CUM_REG samples the per-sample terms are evaluated ONCE
// and their running sums kept in registers across the block scan (single pass); longer
// chunks fall back to evaluating d(j) twice.
constexpr int CUM_REG = 16;

template <typename T, class D, class Post>
__device__ __forceinline__ int cumsum_fwd(D d, Post post, T* out, int n, Scratch* sc, int shift = 0) {
  int lo, hi;
  chunk_range(n, lo, hi);
  const int c = (n + NT - 1) / NT;
  int bad = 0;
  double tot;
  if (c <= CUM_REG) {
    double v[CUM_REG];
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < CUM_REG; k++) {
      const int i = lo + k;
      if (i < hi) run += d(i);
      v[k] = run;
    }
    const double ex = block_excl_scan(run, tot, sc);
#pragma unroll
    for (int k = 0; k < CUM_REG; k++) {
      const int i = lo + k;
      if (i < hi && i >= shift) {
        const T o = post(ex + v[k]);
        bad |= (o != o);
        out[sidx(i - shift)] = o;
      }
    }
  } else {
    double loc = 0.0;
    for (int i = lo; i < hi; i++) loc += d(i);
    double run = block_excl_scan(loc, tot, sc);
    for (int i = lo; i < hi; i++) {
      run += d(i);
      if (i >= shift) {
        const T o = post(run);
        bad |= (o != o);
        out[sidx(i - shift)] = o;
      }
    }
  }
  __syncthreads();
  return bad;
}

template <typename T, class D, class Post>
__device__ __forceinline__ int cumsum_rev(D d, Post post, T* out, int n, Scratch* sc) {
  int lo, hi;
  chunk_range(n, lo, hi);
  const int c = (n + NT - 1) / NT;
  int bad = 0;
  double tot;
  if (c <= CUM_REG) {
    double v[CUM_REG];
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < CUM_REG; k++) {
      const int i = hi - 1 - k;
      if (i >= lo) run += d(i);
      v[k] = run;
    }
    const double ex = block_excl_scan_rev(run, tot, sc);
#pragma unroll
    for (int k = 0; k < CUM_REG; k++) {
      const int i = hi - 1 - k;
      if (i >= lo) {
        const T o = post(ex + v[k]);
        bad |= (o != o);
        out[sidx(i)] = o;
      }
    }
  } else {
    double loc = 0.0;
    for (int i = hi - 1; i >= lo; i--) loc += d(i);
    double run = block_excl_scan_rev(loc, tot, sc);
    for (int i = hi - 1; i >= lo; i--) {
      run += d(i);
      const T o = post(run);
      bad |= (o != o);
      out[sidx(i)] = o;
    }
  }
  __syncthreads();
  return bad;
}

// bl_subtract.py:11-46
template <typename T>
__device__ __forceinline__ void op_bl_subtract(const T* in, T* out, int n, T bl) {
  for (int i = threadIdx.x; i < n; i += NT) out[sidx(i)] = in[sidx(i)] - bl;
  __syncthreads();
}

// min_max.py:11-82 -- first occurrence wins; indices returned as ints.
template <typename T>
__device__ __forceinline__ void op_min_max(const T* in, int n, int& imin, int& imax, T& vmin, T& vmax,
                                           Scratch* sc) {
  T mn = in[sidx(0)], mx = mn;
  int a = 0, b = 0;
  for (int i = threadIdx.x; i < n; i += NT) {  // ascending i per thread: strict compare keeps the first
    T v = in[sidx(i)];
    if (v < mn || (i < a && v == mn)) { mn = v; a = i; }
    if (v > mx || (i < b && v == mx)) { mx = v; b = i; }
  }
  // a thread that never improved on element 0 keeps index 0, which is correct.
  block_argminmax<T>(mn, a, mx, b, sc);
  imin = a; imax = b; vmin = mn; vmax = mx;
}

// numpy.amax over the core dimension (processing_chain.py configs, e.g.
// icpc-dsp-config.json:123-129)
template <typename T>
__device__ __forceinline__ T op_amax(const T* in, int n, Scratch* sc) {
  int a, b;
  T mn, mx;
  op_min_max<T>(in, n, a, b, mn, mx, sc);
  return mx;
}

// linear_slope_fit.py:11-90 : mean, sample standard deviation, least-squares slope and
// intercept of in[0..n).  The reference's Welford update is mathematically
// M2 = sum (x - mean)^2; evaluated here two-pass in float64.
template <typename T>
__device__ __forceinline__ void op_linear_slope_fit(const T* in, int n, T& mean_o, T& stdev_o, T& slope_o,
                                                    T& icpt_o, Scratch* sc) {
  double sy = 0.0, sxy = 0.0;
  for (int i = threadIdx.x; i < n; i += NT) {
    double v = (double)in[sidx(i)];
    sy += v;
    sxy += v * (double)i;
  }
  block_sum2(sy, sxy, sc);
  const double mean = sy / (double)n;
  double m2 = 0.0;
  for (int i = threadIdx.x; i < n; i += NT) {
    double dv = (double)in[sidx(i)] - mean;
    m2 += dv * dv;
  }
  m2 = block_sum(m2, sc);
  const long long nn = n;
  const long long sx = nn * (nn - 1) / 2;
  const long long sx2 = (nn - 1) * nn * (2 * nn - 1) / 6;
  mean_o = (T)mean;
  stdev_o = (T)sqrt(m2 / (double)(n - 1));
  const T slope = (T)(((double)nn * sxy - (double)sx * sy) / (double)(nn * sx2 - sx * sx));
  slope_o = slope;
  icpt_o = (T)((sy - (double)sx * (double)slope) / (double)nn);
}

// linear_slope_fit.py:93-158
template <typename T>
__device__ __forceinline__ void op_linear_slope_diff(const T* in, int n, T slope, T icpt, T& mean_o, T& rms_o,
                                                     Scratch* sc) {
  // mean_k = mean_{k-1} + (t_k - mean_{k-1})?  No: the reference accumulates
  // mean += temp/(i+1) (NOT a running mean) and rms += temp^2; restated exactly.
  double sm = 0.0, sq = 0.0;
  for (int i = threadIdx.x; i < n; i += NT) {
    double t = (double)in[sidx(i)] - ((double)slope * (double)i + (double)icpt);
    sm += t / (double)(i + 1);
    sq += t * t;
  }
  block_sum2(sm, sq, sc);
  mean_o = (T)sm;
  rms_o = (T)sqrt(sq / (double)(n - 1));
}

// arithmetic.py:9-62
template <typename T>
__device__ __forceinline__ T op_mean_below_threshold(const T* in, int n, T thr, Scratch* sc) {
  double tot = 0.0, cnt = 0.0;
  for (int i = threadIdx.x; i < n; i += NT) {
    T v = in[sidx(i)];
    if (v < thr) { tot += (double)v; cnt += 1.0; }
  }
  block_sum2(tot, cnt, sc);
  return cnt == 0.0 ? nan_of<T>() : (T)(tot / cnt);
}

// pole_zero.py:24-77.  y[i] = y[i-1] + x[i] - c*x[i-1], y[0] = x[0]  has the closed form
// y[i] = x[i] + (1-c) * S[i-1],  S = inclusive prefix sum of x: an affine-map scan whose
// linear part is the identity, i.e. a plain float64 prefix sum.
template <typename T>
__device__ __forceinline__ int op_pole_zero(const T* in, T* out, int n, T tau, Scratch* sc) {
  const double omc = -expm1(-1.0 / (double)tau);  // 1 - exp(-1/tau), evaluated without cancellation
  int lo, hi;
  chunk_range(n, lo, hi);
  double loc = 0.0;
  for (int i = lo; i < hi; i++) loc += (double)in[sidx(i)];
  double tot;
  double run = block_excl_scan(loc, tot, sc);  // S[lo-1]
  int bad = 0;
  for (int i = lo; i < hi; i++) {
    const double x = (double)in[sidx(i)];
    T v = (T)(x + omc * run);
    bad |= (v != v);
    out[sidx(i)] = v;
    run += x;
  }
  __syncthreads();
  return bad;
}

// 2x2 affine map  s -> M s + p  used by second-order recursions
struct Aff2 {
  double m00, m01, m10, m11, p0, p1;
};
// (b after a): s -> Mb (Ma s + pa) + pb
__device__ __forceinline__ Aff2 aff2_compose(const Aff2& a, const Aff2& b) {
  Aff2 r;
  r.m00 = b.m00 * a.m00 + b.m01 * a.m10;
  r.m01 = b.m00 * a.m01 + b.m01 * a.m11;
  r.m10 = b.m10 * a.m00 + b.m11 * a.m10;
  r.m11 = b.m10 * a.m01 + b.m11 * a.m11;
  r.p0 = b.m00 * a.p0 + b.m01 * a.p1 + b.p0;
  r.p1 = b.m10 * a.p0 + b.m11 * a.p1 + b.p1;
  return r;
}
__device__ __forceinline__ Aff2 aff2_shfl_up(const Aff2& a, int o) {
  Aff2 r;
  r.m00 = __shfl_up_sync(0xffffffffu, a.m00, o);
  r.m01 = __shfl_up_sync(0xffffffffu, a.m01, o);
  r.m10 = __shfl_up_sync(0xffffffffu, a.m10, o);
  r.m11 = __shfl_up_sync(0xffffffffu, a.m11, o);
  r.p0 = __shfl_up_sync(0xffffffffu, a.p0, o);
  r.p1 = __shfl_up_sync(0xffffffffu, a.p1, o);
  return r;
}
// Exclusive scan of affine maps over the block: returns the composition of the maps of
// all lower threads applied to `init` (state entering this thread's chunk).
__device__ __forceinline__ void block_aff2_excl(const Aff2& mine, double init0, double init1, double& s0,
                                                double& s1, Aff2* warp_tot /* smem [NW] */) {
  const int lane = lane_id(), w = warp_id();
  Aff2 incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Aff2 lower = aff2_shfl_up(incl, o);
    if (lane >= o) incl = aff2_compose(lower, incl);
  }
  __syncthreads();
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  // state entering this warp
  double a0 = init0, a1 = init1;
  for (int k = 0; k < w; k++) {
    const Aff2 t = warp_tot[k];
    const double n0 = t.m00 * a0 + t.m01 * a1 + t.p0;
    const double n1 = t.m10 * a0 + t.m11 * a1 + t.p1;
    a0 = n0;
    a1 = n1;
  }
  // exclusive within the warp
  Aff2 ex = aff2_shfl_up(incl, 1);
  if (lane > 0) {
    const double n0 = ex.m00 * a0 + ex.m01 * a1 + ex.p0;
    const double n1 = ex.m10 * a0 + ex.m11 * a1 + ex.p1;
    a0 = n0;
    a1 = n1;
  }
  s0 = a0;
  s1 = a1;
}

// pole_zero.py:82-198 : w[i] = u[i] - d1 w[i-1] - d2 w[i-2],  u[i] = x[i] + n1 x[i-1] + n2 x[i-2],
// w[0] = x[0], w[1] = x[1].  Parallel affine-map scan with 2-vector state (w[i], w[i-1]).
template <typename T>
__device__ __forceinline__ int op_double_pole_zero(const T* in, T* out, int n, T tau1, T tau2, T frac,
                                                   Scratch* sc, Aff2* warp_tot) {
  const double a = exp(-1.0 / (double)tau1), b = exp(-1.0 / (double)tau2), f = (double)frac;
  const double d1 = f * b - f * a - b - 1.0;
  const double d2 = -1.0 * (f * b - f * a - b);
  const double n1 = -1.0 * (a + b), n2 = a * b;
  // recurrence steps cover i in [2, n); chunk them
  const int m = n - 2;
  const int c = (m + NT - 1) / NT;
  const int lo = 2 + min(m, (int)threadIdx.x * c), hi = 2 + min(m, (int)threadIdx.x * c + c);
  // particular solution from zero state + homogeneous basis responses
  Aff2 mine = {1.0, 0.0, 0.0, 1.0, 0.0, 0.0};
  {
    double p0 = 0.0, p1 = 0.0;  // (w[i-1], w[i-2]) with zero init
    double e00 = 1.0, e01 = 0.0, e10 = 0.0, e11 = 1.0;  // columns of A^k
    for (int i = lo; i < hi; i++) {
      const double u = (double)in[sidx(i)] + n1 * (double)in[sidx(i - 1)] + n2 * (double)in[sidx(i - 2)];
      const double w = u - d1 * p0 - d2 * p1;
      p1 = p0;
      p0 = w;
      const double r00 = -d1 * e00 - d2 * e10, r01 = -d1 * e01 - d2 * e11;
      e10 = e00; e11 = e01; e00 = r00; e01 = r01;
    }
    mine.m00 = e00; mine.m01 = e01; mine.m10 = e10; mine.m11 = e11;
    mine.p0 = p0; mine.p1 = p1;
  }
  double s0, s1;
  block_aff2_excl(mine, (double)in[sidx(1)], (double)in[sidx(0)], s0, s1, warp_tot);
  int bad = 0;
  if (threadIdx.x == 0) {
    out[sidx(0)] = in[sidx(0)];
    out[sidx(1)] = in[sidx(1)];
  }
  for (int i = lo; i < hi; i++) {
    const double u = (double)in[sidx(i)] + n1 * (double)in[sidx(i - 1)] + n2 * (double)in[sidx(i - 2)];
    const double w = u - d1 * s0 - d2 * s1;
    s1 = s0;
    s0 = w;
    T v = (T)w;
    bad |= (v != v);
    out[sidx(i)] = v;
  }
  __syncthreads();
  (void)sc;
  return bad;
}

// trap_filters.py:12-76 / 79-149 : symmetric trapezoid, optionally divided by rise.
template <typename T>
__device__ __forceinline__ int op_trap(const T* in, T* out, int n, int rise, int flat, bool norm, Scratch* sc) {
  const int o1 = rise, o2 = rise + flat, o3 = 2 * rise + flat;
  auto d = [&](int i) -> double {
    double v = (double)in[sidx(i)];
    if (i >= o1) v -= (double)in[sidx(i - o1)];
    if (i >= o2) v -= (double)in[sidx(i - o2)];
    if (i >= o3) v += (double)in[sidx(i - o3)];
    return v;
  };
  if (norm) {
    const double ir = 1.0 / (double)rise;
    return cumsum_fwd<T>(d, [ir](double s) { return (T)(s * ir); }, out, n, sc);
  }
  return cumsum_fwd<T>(d, [](double s) { return (T)s; }, out, n, sc);
}

// trap_filters.py:152-227
template <typename T>
__device__ __forceinline__ int op_asym_trap(const T* in, T* out, int n, int rise, int flat, int fall,
                                            Scratch* sc) {
  const int o1 = rise, o2 = rise + flat, o3 = rise + flat + fall;
  const double ir = 1.0 / (double)rise, ifl = 1.0 / (double)fall;
  auto d = [&](int i) -> double {
    double a = (double)in[sidx(i)];
    if (i >= o1) a -= (double)in[sidx(i - o1)];
    double v = a * ir;
    if (i >= o2) {
      double b = (double)in[sidx(i - o2)];
      if (i >= o3) b -= (double)in[sidx(i - o3)];
      v -= b * ifl;
    }
    return v;
  };
  return cumsum_fwd<T>(d, [](double s) { return (T)s; }, out, n, sc);
}

// trap_filters.py:230-301 : value of the normalised trapezoid at one index by direct
// window sums.  Returns NaN when the windows do not fit (:294-295).
template <typename T>
__device__ __forceinline__ T op_trap_pickoff(const T* in, int n, int rise, int flat, T t_pickoff, int& fatal,
                                             Scratch* sc) {
  fatal = 0;
  if (floor((double)t_pickoff) != (double)t_pickoff) {
    fatal = DSPB_FATAL_PICKOFF_NONINT;
    return nan_of<T>();
  }
  const long long start = (long long)(t_pickoff + (T)1);
  if (!((long long)n >= start && start >= 2LL * rise + flat)) return nan_of<T>();
  const int s = (int)start;
  double i1 = 0.0, i2 = 0.0;
  for (int k = threadIdx.x; k < rise; k += NT) {
    i1 += (double)in[sidx(s - rise + k)];
    i2 += (double)in[sidx(s - 2 * rise - flat + k)];
  }
  block_sum2(i1, i2, sc);
  return (T)((i1 - i2) / (double)rise);
}

// moving_windows.py:12-61 : out[0] = x[0]; out[i] = out[i-1] + (x[i] - x[max(i-L,0)]) / length
template <typename T>
__device__ __forceinline__ int op_mw_left(const T* in, T* out, int n, T length, Scratch* sc) {
  const int L = (int)length;
  const double ilen = 1.0 / (double)length;
  auto d = [&](int i) -> double {
    if (i == 0) return (double)in[sidx(0)];
    const int j = i >= L ? i - L : 0;
    return ((double)in[sidx(i)] - (double)in[sidx(j)]) * ilen;
  };
  return cumsum_fwd<T>(d, [](double s) { return (T)s; }, out, n, sc);
}

// moving_windows.py:64-114 : mirror image (pads with the last sample)
template <typename T>
__device__ __forceinline__ int op_mw_right(const T* in, T* out, int n, T length, Scratch* sc) {
  const int L = (int)length;
  const double ilen = 1.0 / (double)length;
  auto d = [&](int i) -> double {
    if (i == n - 1) return (double)in[sidx(n - 1)];
    const int j = i + L <= n - 1 ? i + L : n - 1;
    return ((double)in[sidx(i)] - (double)in[sidx(j)]) * ilen;
  };
  return cumsum_rev<T>(d, [](double s) { return (T)s; }, out, n, sc);
}

// moving_windows.py:117-203 : num_mw successive moving averages; `tmp` is a second slot.
// The result always ends in `out`.
template <typename T>
__device__ __forceinline__ int op_mw_multi(const T* in, T* out, T* tmp, int n, T length, int num_mw,
                                           int mw_type, Scratch* sc) {
  if (num_mw <= 0) {  // reference: w_out stays NaN (only the initial fill happened)
    fill_slot_nan<T>(out, n);
    __syncthreads();
    return 1;
  }
  int bad = 0;
  const T* src = in;
  for (int k = 0; k < num_mw; k++) {
    // ping-pong so that the last pass writes `out`
    T* dst = ((num_mw - 1 - k) & 1) ? tmp : out;
    const bool right = ((k & 1) && mw_type == 0) || mw_type == 2;
    bad = right ? op_mw_right<T>(src, dst, n, length, sc) : op_mw_left<T>(src, dst, n, length, sc);
    src = dst;
  }
  return bad;
}

// moving_windows.py:206-249 : (x[i+L] - x[i]) / length, same float ops as the reference
template <typename T>
__device__ __forceinline__ void op_avg_current(const T* in, T* out, int n_out, T length) {
  const int L = (int)length;
  for (int i = threadIdx.x; i < n_out; i += NT) {
    T dlt = in[sidx(i + L)] - in[sidx(i)];
    out[sidx(i)] = dlt / length;
  }
  __syncthreads();
}

// time_point_thresh.py:12-92 (mode < 0) and :95-222 (mode = interpolation character).
// Returns the crossing sample index as used by the reference (before interpolation)
// or -1; *fatal receives a DSPB_FATAL_* code for the plain variant's argument errors.
template <typename T>
__device__ __forceinline__ int search_crossing(const T* in, int n, T thr, int s, bool forward, int stop_back,
                                               Scratch* sc) {
  if (forward) {
    // first i in [s, n-2] with  w[i] <= thr < w[i+1]  or  w[i] >= thr > w[i+1]
    for (int base = s; base < n - 1; base += NT) {
      const int i = base + threadIdx.x;
      int hit = 0x7fffffff;
      if (i < n - 1) {
        const T a = in[sidx(i)], b = in[sidx(i + 1)];
        if ((a <= thr && thr < b) || (a >= thr && thr > b)) hit = i;
      }
      hit = block_min_int(hit, sc);
      if (hit != 0x7fffffff) return hit;
    }
    return -1;
  }
  // backward: first i walking down from s to stop_back with
  //   w[i-1] < thr <= w[i]  or  w[i-1] > thr >= w[i]
  for (int base = s; base >= stop_back; base -= NT) {
    const int i = base - threadIdx.x;
    int hit = -1;
    if (i >= stop_back) {
      const T a = in[sidx(i - 1)], b = in[sidx(i)];
      if ((a < thr && thr <= b) || (a > thr && thr >= b)) hit = i;
    }
    hit = block_max_int(hit, sc);
    if (hit >= 0) return hit;
  }
  return -1;
}

template <typename T>
__device__ __forceinline__ T op_time_point_thresh(const T* in, int n, T thr, T t_start, T walk, int& fatal,
                                                  Scratch* sc) {
  fatal = 0;
  if (thr != thr || t_start != t_start || walk != walk) return nan_of<T>();
  if (floor((double)t_start) != (double)t_start) { fatal = DSPB_FATAL_TSTART_NONINT; return nan_of<T>(); }
  if (floor((double)walk) != (double)walk) { fatal = DSPB_FATAL_WALK_NONINT; return nan_of<T>(); }
  const long long s = (long long)t_start;
  if (s < 0 || s >= n) { fatal = DSPB_FATAL_TSTART_RANGE; return nan_of<T>(); }
  const int hit = search_crossing<T>(in, n, thr, (int)s, (long long)walk == 1, 1, sc);
  return hit < 0 ? nan_of<T>() : (T)hit;
}

template <typename T>
__device__ __forceinline__ T op_interp_time_point_thresh(const T* in, int n, T thr, T t_start, long long walk,
                                                         int mode, int& fatal, Scratch* sc) {
  fatal = 0;
  if (thr != thr || t_start != t_start) return nan_of<T>();
  if (t_start < (T)0 || t_start >= (T)n) return nan_of<T>();
  const int s = (int)t_start;
  int ic;
  if (walk > 0) {
    ic = search_crossing<T>(in, n, thr, s, true, 0, sc);
  } else {
    ic = search_crossing<T>(in, n, thr, s, false, 2, sc);  // the reference's loop stops at 2 (:192)
    if (ic >= 0) ic -= 1;
  }
  if (ic < 0) return nan_of<T>();
  switch (mode) {
    case 'i': case 'b': case 'c': return (T)ic;
    case 'a': case 'f': return (T)(ic + 1);
    case 'r': {
      const T a = fabs(thr - in[sidx(ic)]), b = fabs(thr - in[sidx(ic + 1)]);
      return a < b ? (T)ic : (T)(ic + 1);
    }
    case 'n': return (T)((double)ic + 0.5);
    case 'l': {
      const T q = (thr - in[sidx(ic)]) / (in[sidx(ic + 1)] - in[sidx(ic)]);
      return (T)((double)ic + (double)q);
    }
  }
  fatal = DSPB_FATAL_INTERP_MODE;
  return nan_of<T>();
}

// cubic-spline mode of fixed_time_pickoff (fixed_time_pickoff.py:107-123), kept out of line:
// it needs a small local window and is rarely used.
template <typename T>
__device__ __noinline__ T ftp_spline(const T* in, int n, int i_in, double t0, double t1) {
  constexpr int HALO = 64;
  // forward sweep needs u[i], w2f[i] for i in [i_in, top]; start HALO before that
  const int top = min(n - 2, i_in + 1 + HALO);       // backward sweep starts here
  const int first = max(1, i_in - HALO);
  // forward coefficients w2f[i] = -0.5 / (0.5 w2f[i-1] + 2), w2f[0] = 0 (fixed point sqrt(3)-2)
  double w2f = first == 1 ? 0.0 : -0.2679491924311227;
  double u = 0.0;
  // we need u[i], w2f[i] on [i_in, top] for the backward sweep: keep them in a small
  // local window (top - i_in + 1 <= HALO + 2)
  double uw[HALO + 3], cw[HALO + 3];
  for (int k = 0; k < HALO + 3; k++) { uw[k] = 0.0; cw[k] = 0.0; }
  for (int i = first; i <= min(top, n - 2); i++) {
    const double p = 0.5 * w2f + 2.0;
    w2f = -0.5 / p;
    const double sec = ((double)in[sidx(i + 1)] - 2.0 * (double)in[sidx(i)]) + (double)in[sidx(i - 1)];
    u = (3.0 * sec - 0.5 * u) / p;
    if (i >= i_in) { uw[i - i_in] = u; cw[i - i_in] = w2f; }
  }
  // i_in == 0: u[0] = w2[0] = 0 in the reference
  // backward: w2[i] = w2f[i] * w2[i+1] + u[i], from top down to i_in; w2[n-1] = 0
  double w2n = 0.0;   // w2[i+1]
  double w2_i = 0.0, w2_ip1 = 0.0;
  for (int i = top; i >= i_in; i--) {
    double v;
    if (i == 0) v = 0.0 * w2n + 0.0;  // w2[0]*w2[1] + u[0] with w2[0] = u[0] = 0
    else v = cw[i - i_in] * w2n + uw[i - i_in];
    if (i == i_in + 1) w2_ip1 = v;
    if (i == i_in) w2_i = v;
    w2n = v;
  }
  if (i_in + 1 > top) w2_ip1 = 0.0;  // i_in + 1 == n - 1
  const double a3 = t1 * t1 * t1, b3 = t0 * t0 * t0;
  double r = t1 * (double)in[sidx(i_in)] + t0 * (double)in[sidx(i_in + 1)];
  r += ((a3 - t1) * w2_i + (b3 - t0) * w2_ip1) / 6.0;
  return (T)r;
}

// fixed_time_pickoff.py:12-125.  Executed by every thread redundantly (O(1) work; the
// spline mode uses the exponentially decaying influence of far samples: the tridiagonal
// recursions contract by 2-sqrt(3) = 0.268 per step, so a 64-sample halo reproduces the
// full-length solve to float64 round-off).
template <typename T>
__device__ __forceinline__ T op_fixed_time_pickoff(const T* in, int n, T t_in, int mode, int& fatal) {
  fatal = 0;
  if (t_in != t_in) return nan_of<T>();
  if (t_in < (T)0 || t_in > (T)(n - 1)) return nan_of<T>();
  const int i_in = (int)t_in;
  if ((T)i_in == t_in) return in[sidx(i_in)];
  const double t0 = (double)t_in - (double)i_in, t1 = 1.0 - t0;
  switch (mode) {
    case 'i': fatal = DSPB_FATAL_FTP_INT; return nan_of<T>();
    case 'n': return t0 < 0.5 ? in[sidx(i_in)] : in[sidx(i_in + 1)];
    case 'f': return in[sidx(i_in)];
    case 'c': return in[sidx(i_in + 1)];
    case 'l': return (T)(t1 * (double)in[sidx(i_in)] + t0 * (double)in[sidx(i_in + 1)]);
    case 'h': {
      const double m0 = i_in == 0 ? (double)(T)(in[sidx(1)] - in[sidx(0)])
                                  : (double)(T)(in[sidx(i_in + 1)] - in[sidx(i_in - 1)]) / 2.0;
      const double m1 = i_in == n - 2 ? (double)(T)(in[sidx(n - 1)] - in[sidx(n - 2)])
                                      : (double)(T)(in[sidx(i_in + 2)] - in[sidx(i_in)]) / 2.0;
      const double a2 = t1 * t1, a3 = a2 * t1, b2 = t0 * t0, b3 = b2 * t0;
      double r = (-2.0 * a3 + 3.0 * a2) * (double)in[sidx(i_in)];
      r += (-2.0 * b3 + 3.0 * b2) * (double)in[sidx(i_in + 1)];
      r -= (a3 - a2) * m0;
      r += (b3 - b2) * m1;
      return (T)r;
    }
    case 's': return ftp_spline<T>(in, n, i_in, t0, t1);
  }
  fatal = DSPB_FATAL_INTERP_MODE;
  return nan_of<T>();
}

// min_max.py:85-140
template <typename T>
__device__ __forceinline__ void op_min_max_norm(const T* in, T* out, int n, T a_min, T a_max) {
  const T amx = fabs(a_max), amn = fabs(a_min);
  T div = (T)1;
  bool copy = false;
  if (amx == (T)0 || amn == (T)0) copy = true;
  else if (amx >= amn) div = amx;
  else div = amn;
  for (int i = threadIdx.x; i < n; i += NT) {
    const T v = in[sidx(i)];
    // NaN a_min/a_max: none of the reference's branches fire -> output stays NaN
    out[sidx(i)] = (amx != amx || amn != amn) ? nan_of<T>() : (copy ? v : v / div);
  }
  __syncthreads();
}

// windower.py:12-54 ; returns 1 if any output sample is NaN padding
template <typename T>
__device__ __forceinline__ int op_windower(const T* in, T* out, int n, int m, T t0_in) {
  long long beg = (long long)t0_in;
  if (beg > n) beg = n;
  int padded = 0;
  for (int i = threadIdx.x; i < m; i += NT) {
    const long long j = beg + i;
    T v = nan_of<T>();
    if (j >= 0 && j < n) v = in[sidx((int)j)];
    else padded = 1;
    out[sidx(i)] = v;
  }
  __syncthreads();
  return padded;
}

// upsampler.py:14-49 : output t receives input sample t_in when
// t in [trunc(t_in*up - floor(up/2)), +int(up)); later t_in overwrite earlier ones.
template <typename T>
__device__ __forceinline__ int op_upsampler(const T* in, T* out, int n, int m, T upsample) {
  const double up = (double)upsample;
  const double half = floor(up / 2.0);
  const int reps = (int)upsample;
  int holes = 0;
  for (int t = threadIdx.x; t < m; t += NT) {
    // candidates: the largest t_in whose start <= t; check a small neighbourhood because
    // of the truncation towards zero of negative starts
    int guess = (int)floor(((double)t + half) / up) + 1;
    T v = nan_of<T>();
    bool found = false;
    for (int ti = min(guess, n - 1); ti >= 0 && ti >= guess - 3; ti--) {
      const long long st = (long long)((double)ti * up - half);  // C cast truncates like int()
      if (st <= t && t < st + reps) { v = in[sidx(ti)]; found = true; break; }
    }
    if (!found) holes = 1;
    out[sidx(t)] = v;
  }
  __syncthreads();
  return holes;
}

}  // namespace dspb
