// dspeed_b200 -- one kernel + extern "C" launcher per hot-path processor of
// dspeed.processors (un-fused path: any JSON chain made of these processors runs on
// the device, one launch per processor and block).  The fused path (fused.cu) calls
// the same block routines (row_ops.cuh) on shared-memory resident data.
#include <cmath>
#include <cstdio>

#include "row_ops.cuh"

using namespace dspb;

namespace {

constexpr size_t MAX_SMEM = 227 * 1024;

template <typename T, class Body>
__global__ void __launch_bounds__(DEFAULT_THREADS) k_rows(const Body body, const long long n_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Scratch* sc = reinterpret_cast<Scratch*>(smem_raw);
  T* slots = reinterpret_cast<T*>(smem_raw + SCRATCH_BYTES);
  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    body(row, slots, sc);
    __syncthreads();
  }
}

template <typename T, class Body>
int launch_rows(const Body& body, long long n_rows, int n_slots, long long max_len, void* stream,
                size_t extra_bytes = 0) {
  if (n_rows <= 0) return 0;
  if (max_len > (1 << 24)) return DSPB_ERR_ROW_TOO_LONG;
  const size_t smem = SCRATCH_BYTES + (size_t)n_slots * slot_words((int)max_len) * sizeof(T) + extra_bytes;
  if (smem > MAX_SMEM) return DSPB_ERR_ROW_TOO_LONG;
  auto kern = k_rows<T, Body>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(int)e;
  const long long max_grid = 148LL * 16;
  const int grid = (int)(n_rows < max_grid ? n_rows : max_grid);
  kern<<<grid, DEFAULT_THREADS, smem, (cudaStream_t)stream>>>(body, n_rows);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

template <typename T>
Scalar<T> mk_scalar(const void* p, int64_t stride, double imm) {
  Scalar<T> s;
  s.ptr = reinterpret_cast<const T*>(p);
  s.stride = stride;
  s.imm = (T)imm;
  return s;
}
inline Wave mk_wave(const void* p, int64_t rs, int32_t dt) {
  Wave w;
  w.ptr = p;
  w.row_stride = rs;
  w.dtype = dt;
  return w;
}
template <typename T>
struct WOut {
  T* ptr;
  long long row_stride;
  __device__ __forceinline__ T* row(long long r) const { return ptr + r * row_stride; }
};
template <typename T>
WOut<T> mk_out(void* p, int64_t rs) {
  WOut<T> o;
  o.ptr = reinterpret_cast<T*>(p);
  o.row_stride = rs;
  return o;
}

template <class U, class V>
__device__ __forceinline__ U* align8(V* p) {
  return reinterpret_cast<U*>((reinterpret_cast<uintptr_t>(p) + 7) & ~(uintptr_t)7);
}

// stage input row into slot 0 and OR the NaN flags over the block
template <typename T>
__device__ __forceinline__ int stage_in(T* slot, const Wave& w, long long row, int n) {
  const int f = stage_row<T>(slot, w, row, n);
  return block_or(f);  // also makes the slot visible
}

// ---------------------------------------------------------------------------------
template <typename T>
struct BlSubtract {
  Wave in; int n; Scalar<T> bl; WOut<T> out;
  __device__ void operator()(long long row, T* s, Scratch*) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    const T b = bl.get(row);
    T* o = out.row(row);
    if (nan_in || b != b) { store_row_nan<T>(o, n); return; }
    for (int i = threadIdx.x; i < n; i += NT) o[i] = s[sidx(i)] - b;
  }
};

template <typename T>
struct MinMax {
  Wave in; int n; T *t_min, *t_max, *a_min, *a_max; bool only_max;
  __device__ void operator()(long long row, T* s, Scratch* sc) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    int imin = 0, imax = 0; T vmin = nan_of<T>(), vmax = nan_of<T>();
    if (!nan_in) op_min_max<T>(s, n, imin, imax, vmin, vmax, sc);
    if (threadIdx.x == 0) {
      if (only_max) { a_max[row] = vmax; return; }
      t_min[row] = nan_in ? nan_of<T>() : (T)imin;
      t_max[row] = nan_in ? nan_of<T>() : (T)imax;
      a_min[row] = vmin;
      a_max[row] = vmax;
    }
  }
};

template <typename T>
struct MinMaxNorm {
  Wave in; int n; Scalar<T> a_min, a_max; WOut<T> out;
  __device__ void operator()(long long row, T* s, Scratch*) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    T* o = out.row(row);
    if (nan_in) { store_row_nan<T>(o, n); return; }
    T* s2 = s + slot_words(n);
    op_min_max_norm<T>(s, s2, n, a_min.get(row), a_max.get(row));
    store_row<T>(o, s2, n);
  }
};

template <typename T>
struct LinearSlopeFit {
  Wave in; int n; T *mean, *stdev, *slope, *icpt;
  __device__ void operator()(long long row, T* s, Scratch* sc) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    T m = nan_of<T>(), sd = m, sl = m, ic = m;
    if (!nan_in) op_linear_slope_fit<T>(s, n, m, sd, sl, ic, sc);
    if (threadIdx.x == 0) { mean[row] = m; stdev[row] = sd; slope[row] = sl; icpt[row] = ic; }
  }
};

template <typename T>
struct LinearSlopeDiff {
  Wave in; int n; Scalar<T> slope, icpt; T *mean, *rms;
  __device__ void operator()(long long row, T* s, Scratch* sc) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    const T sl = slope.get(row), ic = icpt.get(row);
    T m = nan_of<T>(), r = m;
    if (!nan_in && sl == sl && ic == ic) op_linear_slope_diff<T>(s, n, sl, ic, m, r, sc);
    if (threadIdx.x == 0) { mean[row] = m; rms[row] = r; }
  }
};

template <typename T>
struct MeanBelowThreshold {
  Wave in; int n; Scalar<T> thr; T* res;
  __device__ void operator()(long long row, T* s, Scratch* sc) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    const T th = thr.get(row);
    T r = nan_of<T>();
    if (!nan_in && th == th) r = op_mean_below_threshold<T>(s, n, th, sc);
    if (threadIdx.x == 0) res[row] = r;
  }
};

template <typename T>
struct PoleZero {
  Wave in; int n; Scalar<T> tau; WOut<T> out; int* fatal;
  __device__ void operator()(long long row, T* s, Scratch* sc) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    const T t = tau.get(row);
    T* o = out.row(row);
    if (nan_in || t != t) { store_row_nan<T>(o, n); return; }
    T* s2 = s + slot_words(n);
    const int bad = block_or(op_pole_zero<T>(s, s2, n, t, sc));
    if (bad && threadIdx.x == 0) raise_fatal(fatal, DSPB_FATAL_PZ_NAN, row);
    store_row<T>(o, s2, n);
  }
};

template <typename T>
struct DoublePoleZero {
  Wave in; int n; Scalar<T> tau1, tau2, frac; WOut<T> out;
  __device__ void operator()(long long row, T* s, Scratch* sc) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    const T a = tau1.get(row), b = tau2.get(row), f = frac.get(row);
    T* o = out.row(row);
    if (nan_in || a != a || b != b || f != f) { store_row_nan<T>(o, n); return; }
    T* s2 = s + slot_words(n);
    Aff2* wt = align8<Aff2>(s2 + slot_words(n));
    op_double_pole_zero<T>(s, s2, n, a, b, f, sc, wt);
    store_row<T>(o, s2, n);
  }
};

template <typename T>
struct Trap {
  Wave in; int n; int rise, flat, fall; int kind;  // 0 trap_filter, 1 trap_norm, 2 asym
  WOut<T> out;
  __device__ void operator()(long long row, T* s, Scratch* sc) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    T* o = out.row(row);
    if (nan_in) { store_row_nan<T>(o, n); return; }
    T* s2 = s + slot_words(n);
    if (kind == 2) op_asym_trap<T>(s, s2, n, rise, flat, fall, sc);
    else op_trap<T>(s, s2, n, rise, flat, kind == 1, sc);
    store_row<T>(o, s2, n);
  }
};

template <typename T>
struct TrapPickoff {
  Wave in; int n; int rise, flat; Scalar<T> t; T* a_out; int* fatal;
  __device__ void operator()(long long row, T* s, Scratch* sc) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    const T tp = t.get(row);
    T r = nan_of<T>();
    int f = 0;
    if (!nan_in && tp == tp) r = op_trap_pickoff<T>(s, n, rise, flat, tp, f, sc);
    if (threadIdx.x == 0) { a_out[row] = r; raise_fatal(fatal, f, row); }
  }
};

template <typename T>
struct MovingWindow {
  Wave in; int n; T length; int kind; int num_mw; int mw_type;  // kind 0 left, 1 right, 2 multi
  WOut<T> out;
  __device__ void operator()(long long row, T* s, Scratch* sc) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    T* o = out.row(row);
    if (nan_in) { store_row_nan<T>(o, n); return; }
    T* s2 = s + slot_words(n);
    if (kind == 0) op_mw_left<T>(s, s2, n, length, sc);
    else if (kind == 1) op_mw_right<T>(s, s2, n, length, sc);
    else op_mw_multi<T>(s, s2, s2 + slot_words(n), n, length, num_mw, mw_type, sc);
    store_row<T>(o, s2, n);
  }
};

template <typename T>
struct AvgCurrent {
  Wave in; int n; T length; WOut<T> out; int n_out;
  __device__ void operator()(long long row, T* s, Scratch*) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    T* o = out.row(row);
    if (nan_in) { store_row_nan<T>(o, n_out); return; }
    const int L = (int)length;
    for (int i = threadIdx.x; i < n_out; i += NT) {
      const T d = s[sidx(i + L)] - s[sidx(i)];
      o[i] = d / length;
    }
  }
};

template <typename T>
struct TimePointThresh {
  Wave in; int n; Scalar<T> thr, ts, walk; T* t_out; int* fatal;
  bool interp; long long walk_i; int mode;
  __device__ void operator()(long long row, T* s, Scratch* sc) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    T r = nan_of<T>();
    int f = 0;
    if (!nan_in) {
      if (interp) r = op_interp_time_point_thresh<T>(s, n, thr.get(row), ts.get(row), walk_i, mode, f, sc);
      else r = op_time_point_thresh<T>(s, n, thr.get(row), ts.get(row), walk.get(row), f, sc);
    }
    if (threadIdx.x == 0) { t_out[row] = r; raise_fatal(fatal, f, row); }
  }
};

template <typename T>
struct FixedTimePickoff {
  Wave in; int n; Scalar<T> t; int mode; T* a_out; int* fatal;
  __device__ void operator()(long long row, T* s, Scratch*) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    if (threadIdx.x == 0) {
      T r = nan_of<T>();
      int f = 0;
      if (!nan_in) r = op_fixed_time_pickoff<T>(s, n, t.get(row), mode, f);
      a_out[row] = r;
      raise_fatal(fatal, f, row);
    }
  }
};

template <typename T>
struct Windower {
  Wave in; int n; Scalar<T> t0; WOut<T> out; int m;
  __device__ void operator()(long long row, T* s, Scratch*) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    const T t = t0.get(row);
    T* o = out.row(row);
    if (nan_in || t != t) { store_row_nan<T>(o, m); return; }
    long long beg = (long long)t;
    if (beg > n) beg = n;
    for (int i = threadIdx.x; i < m; i += NT) {
      const long long j = beg + i;
      o[i] = (j >= 0 && j < n) ? s[sidx((int)j)] : nan_of<T>();
    }
  }
};

template <typename T>
struct Upsampler {
  Wave in; int n; T up; WOut<T> out; int m;
  __device__ void operator()(long long row, T* s, Scratch*) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    T* o = out.row(row);
    if (nan_in) { store_row_nan<T>(o, m); return; }
    T* s2 = s + slot_words(n);
    op_upsampler<T>(s, s2, n, m, up);
    store_row<T>(o, s2, m);
  }
};

// time_point_thresh.py:225-401 -- sequential by construction (one sweep serving all
// thresholds); one thread walks the shared-memory resident row.
template <typename T>
struct MultiTimePointThresh {
  Wave in; int n; const T* thr; int m; long long thr_row_stride; Scalar<T> ts; int pol; int mode;
  T* t_out; int* fatal;
  __device__ static T wrapget(const T* s, int n, int i) { return s[sidx(i < 0 ? i + n : i)]; }
  __device__ int set(T* to, int idx, const T* s, const T* th, int i_wf) const {
    switch (mode) {
      case 'i': to[idx] = (T)i_wf; break;
      case 'a': case 'f': to[idx] = (T)(pol < 0 ? i_wf : i_wf + 1); break;
      case 'b': case 'c': to[idx] = (T)(pol > 0 ? i_wf : i_wf - 1); break;
      case 'r':
        if ((T)(th[idx] - wrapget(s, n, i_wf)) < (T)(wrapget(s, n, i_wf + pol) - th[idx])) to[idx] = (T)i_wf;
        else to[idx] = (T)(i_wf + pol);
        break;
      case 'n': to[idx] = (T)((double)i_wf + 0.5 * (double)pol); break;
      case 'l': {
        const T q = (th[idx] - wrapget(s, n, i_wf)) / (wrapget(s, n, i_wf + pol) - wrapget(s, n, i_wf));
        to[idx] = (T)((double)i_wf + (double)q);
        break;
      }
      default: return DSPB_FATAL_INTERP_MODE;
    }
    return 0;
  }
  __device__ void operator()(long long row, T* s, Scratch*) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    T* to = t_out + row * m;
    for (int i = threadIdx.x; i < m; i += NT) to[i] = nan_of<T>();
    __syncthreads();
    if (threadIdx.x != 0 || nan_in) return;
    const T* th = thr + row * thr_row_stride;
    const T tsv = ts.get(row);
    if (tsv != tsv) return;
    for (int i = 0; i < m; i++) if (th[i] != th[i]) return;
    if (tsv < (T)0 || tsv >= (T)n) return;
    // stable argsort (insertion; m is small) kept in the int scratch behind the row slot
    int* srt = reinterpret_cast<int*>(s + slot_words(n));
    for (int i = 0; i < m; i++) srt[i] = i;
    for (int i = 1; i < m; i++) {
      int k = srt[i], j = i - 1;
      while (j >= 0 && th[srt[j]] > th[k]) { srt[j + 1] = srt[j]; j--; }
      srt[j + 1] = k;
    }
    const int t_start = (int)tsv;
    const T a_start = s[sidx(t_start)];
    int i_start = m;
    for (int i = 0; i < m; i++) if (th[srt[i]] >= a_start) { i_start = i; break; }
    int i_tp = i_start;
    if (i_tp < m) {
      int idx = srt[i_tp];
      const int stop = pol > 0 ? n - 1 : -1;
      for (int i_wf = t_start; pol > 0 ? i_wf < stop : i_wf > stop; i_wf += pol) {
        if (i_tp >= m) break;
        while (wrapget(s, n, i_wf) <= th[idx] && th[idx] < wrapget(s, n, i_wf + pol)) {
          const int rc = set(to, idx, s, th, i_wf);
          if (rc) { raise_fatal(fatal, rc, row); return; }
          i_tp++;
          if (i_tp >= m) break;
          idx = srt[i_tp];
        }
      }
    }
    i_tp = i_start - 1;
    if (i_tp >= 0) {
      int idx = srt[i_tp];
      const int stop = pol < 0 ? n - 1 : -1;
      const int step = -pol;
      for (int i_wf = t_start - 1; step > 0 ? i_wf < stop : i_wf > stop; i_wf += step) {
        if (i_tp < 0) break;
        while (wrapget(s, n, i_wf) <= th[idx] && th[idx] < wrapget(s, n, i_wf + pol)) {
          const int rc = set(to, idx, s, th, i_wf);
          if (rc) { raise_fatal(fatal, rc, row); return; }
          i_tp--;
          if (i_tp < 0) break;
          idx = srt[i_tp];
        }
      }
    }
  }
};

// recursive_filter.py:12-93 for len(b) <= 3 (order <= 2), as an affine-map scan.
template <typename T>
struct RecursiveFilter {
  Wave in; int n; double a[8]; int p; double b[3]; int q; Scalar<T> init_in, init_out; WOut<T> out;
  __device__ double src(const T* s, int i, double ii) const {
    double u = 0.0;
    for (int j = 0; j < p; j++) u += a[j] * (j <= i ? (double)s[sidx(i - j)] : ii);
    return u;
  }
  __device__ void operator()(long long row, T* s, Scratch* sc) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    const T iiv = init_in.get(row), iov = init_out.get(row);
    T* o = out.row(row);
    if (nan_in || iiv != iiv || iov != iov) { store_row_nan<T>(o, n); return; }
    T* s2 = s + slot_words(n);
    Aff2* wt = align8<Aff2>(s2 + slot_words(n));
    const double ii = (double)iiv;
    const double b0 = b[0], b1 = q > 1 ? b[1] : 0.0, b2 = q > 2 ? b[2] : 0.0;
    // y[i] = (u[i] - b1 y[i-1] - b2 y[i-2]) / b0 ; state entering i = 0 is (init_out, init_out)
    int lo, hi;
    chunk_range(n, lo, hi);
    Aff2 mine = {1.0, 0.0, 0.0, 1.0, 0.0, 0.0};
    {
      double p0 = 0.0, p1 = 0.0, e00 = 1.0, e01 = 0.0, e10 = 0.0, e11 = 1.0;
      for (int i = lo; i < hi; i++) {
        const double y = (src(s, i, ii) - b1 * p0 - b2 * p1) / b0;
        p1 = p0; p0 = y;
        const double r00 = (-b1 * e00 - b2 * e10) / b0, r01 = (-b1 * e01 - b2 * e11) / b0;
        e10 = e00; e11 = e01; e00 = r00; e01 = r01;
      }
      mine.m00 = e00; mine.m01 = e01; mine.m10 = e10; mine.m11 = e11; mine.p0 = p0; mine.p1 = p1;
    }
    double s0, s1;
    block_aff2_excl(mine, (double)iov, (double)iov, s0, s1, wt);
    for (int i = lo; i < hi; i++) {
      const double y = (src(s, i, ii) - b1 * s0 - b2 * s1) / b0;
      s1 = s0; s0 = y;
      s2[sidx(i)] = (T)y;
    }
    __syncthreads();
    store_row<T>(o, s2, n);
    (void)sc;
  }
};

// ---------------------------------------------------------------------------------
// get_multi_local_extrema.py:12-306.  The peak-detection state machine is sequential
// per waveform: lane 0 of warp 0 walks left-to-right, lane 0 of warp 1 right-to-left,
// concurrently, over the shared-memory resident row; the AND / OR merges follow.
// ---------------------------------------------------------------------------------
template <typename T>
struct MultiLocalExtrema {
  Wave in; int n; T d_max, d_min; T dir; Scalar<T> abs_max_s, abs_min_s; T *vt_max, *vt_min; int m;   // (thresholds may be per-row)
  uint32_t *n_max, *n_min;
  __device__ static void walk(const T* s, int n, bool fwd, T d_max, T d_min, T abs_max, T abs_min,
                              float* v_max, float* v_min, int m, int& c_max, int& c_min) {
    bool find_max = true;
    int imax = fwd ? 0 : n - 1, imin = imax;
    T vmax = s[sidx(imax)], vmin = vmax;
    int cm = 0, cn = 0;
    for (int k = 0; k < n; k++) {
      const int i = fwd ? k : n - 1 - k;
      const T v = s[sidx(i)];
      if (v > vmax) { vmax = v; imax = i; }
      if (v < vmin) { vmin = v; imin = i; }
      if (find_max) {
        if (v < (T)(vmax - d_max) && cm < m && vmax > abs_max) {
          v_max[cm++] = (float)imax;
          imin = i; vmin = v;
          find_max = false;
        }
      } else {
        if (v > (T)(vmin + d_min) && cn < m && vmin < abs_min) {
          v_min[cn++] = (float)imin;
          imax = i; vmax = v;
          find_max = true;
        }
      }
    }
    c_max = cm; c_min = cn;
  }
  __device__ void operator()(long long row, T* s, Scratch* sc) const {
    const int nan_in = stage_in<T>(s, in, row, n);
    const T abs_max = abs_max_s.get(row), abs_min = abs_min_s.get(row);
    T* omax = vt_max + row * m;
    T* omin = vt_min + row * m;
    for (int i = threadIdx.x; i < m; i += NT) { omax[i] = nan_of<T>(); omin[i] = nan_of<T>(); }
    if (threadIdx.x == 0) { n_max[row] = 0; n_min[row] = 0; }
    if (nan_in || d_max != d_max || d_min != d_min) return;
    // index lists (as floats; indices < 2^24) behind the row slot: L max, L min, R max, R min
    float* lst = reinterpret_cast<float*>(s + slot_words(n));
    float *l_max = lst, *l_min = lst + m, *r_max = lst + 2 * m, *r_min = lst + 3 * m;
    int* cnt = sc->i + 2 * MAXW;  // 4 counters
    if (threadIdx.x < 4) cnt[threadIdx.x] = 0;
    __syncthreads();
    const bool do_left = (dir == (T)0) || (dir > (T)1);
    const bool do_right = dir > (T)0;
    if (threadIdx.x == 0 && do_left) {
      int a, b;
      walk(s, n, true, d_max, d_min, abs_max, abs_min, l_max, l_min, m, a, b);
      cnt[0] = a; cnt[1] = b;
    }
    if (threadIdx.x == 32 && do_right) {
      int a, b;
      walk(s, n, false, d_max, d_min, abs_max, abs_min, r_max, r_min, m, a, b);
      cnt[2] = a; cnt[3] = b;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int nl_max = cnt[0], nl_min = cnt[1], nr_max = cnt[2], nr_min = cnt[3];
    if (dir == (T)0) {
      for (int i = 0; i < nl_max; i++) omax[i] = (T)l_max[i];
      for (int i = 0; i < nl_min; i++) omin[i] = (T)l_min[i];
      n_max[row] = nl_max; n_min[row] = nl_min;
    } else if (dir == (T)1) {
      for (int i = 0; i < nr_max; i++) omax[i] = (T)r_max[i];
      for (int i = 0; i < nr_min; i++) omin[i] = (T)r_min[i];
      n_max[row] = nr_max; n_min[row] = nr_min;
    } else if (dir == (T)2) {
      // AND: left maxima (in order) also found from the right.  Right lists are found in
      // descending position; membership tests do not need them sorted.
      int c = 0;
      for (int i = 0; i < nl_max; i++)
        for (int j = 0; j < nr_max; j++)
          if (r_max[j] == l_max[i]) { omax[c++] = (T)l_max[i]; break; }
      n_max[row] = c;
      // the reference (:255-256) builds the "minima" from the first n_min entries of the
      // (sorted) right maxima and of the left maxima -- reproduced as is.
      c = 0;
      // first nr_min entries of the ascending-sorted right maxima = the nr_min smallest =
      // the last nr_min found (found in descending order)
      for (int i = 0; i < nl_min; i++)
        for (int j = nr_max - nr_min; j < nr_max; j++)
          if (j >= 0 && r_max[j] == l_max[i]) { omin[c++] = (T)l_max[i]; break; }
      n_min[row] = (nl_min > 0 && nr_min > 0) ? c : 0;
    } else if (dir == (T)3) {
      // OR: sorted unique union, truncated to m slots
      for (int pass = 0; pass < 2; pass++) {
        const float* L = pass == 0 ? l_max : l_min;
        const float* R = pass == 0 ? r_max : r_min;  // descending
        const int nl = pass == 0 ? nl_max : nl_min, nr = pass == 0 ? nr_max : nr_min;
        T* o = pass == 0 ? omax : omin;
        int i = 0, j = nr - 1, c = 0;
        float last = -1.0f;
        while ((i < nl || j >= 0) && c < m) {
          float v;
          if (j < 0 || (i < nl && L[i] <= R[j])) v = L[i++];
          else v = R[j--];
          if (v != last) { o[c++] = (T)v; last = v; }
        }
        (pass == 0 ? n_max : n_min)[row] = c;
      }
    }
  }
};

// set-up time kernel synthesis (float64 math, rounded like the reference; see oracle)
template <typename T>
__global__ void k_cusp_stage1(double sigma, int flat_int, int lt, T* kernel, int length) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= length) return;
  double v;
  if (i < lt) v = sinh((double)i / sigma) / sinh((double)lt / sigma);
  else if (i <= lt + flat_int) v = 1.0;
  else v = sinh((double)(length - i) / sigma) / sinh((double)lt / sigma);
  kernel[i] = (T)v;  // rounded to the kernel dtype BEFORE the differencing, as in energy_kernels.py:66-73
}
// np.convolve(k, [1, -c], "same")[i] = k[i] - c * k[i-1]  (k[-1] = 0)
template <typename T>
__global__ void k_diff_same(const T* in, double c, T* out, int length) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= length) return;
  const double a = (double)in[i];
  const double b = i > 0 ? (double)in[i - 1] : 0.0;
  out[i] = (T)__dadd_rn(a, __dmul_rn(b, -c));  // unfused, like numpy's two-tap dot
}
template <typename T>
__global__ void k_zac(double sigma, int flat_int, int lt, double decay, T* kernel, int length) {
  // single CTA: float64 cusp and parabola, sequential area sums by thread 0 (bit-faithful
  // to the reference's python loop), then the [1, -c] differencing.
  extern __shared__ double zs[];
  double* cusp = zs;
  double* par = zs + length;
  const double half = (double)lt / 2.0;
  for (int i = threadIdx.x; i < length; i += blockDim.x) {
    double cv = 0.0, pv = 0.0;
    if (i < lt) {
      cv = sinh((double)i / sigma) / sinh((double)lt / sigma);
      pv = ((double)i - half) * ((double)i - half) - half * half;
    } else if (i <= lt + flat_int) {
      cv = 1.0;
    } else {
      cv = sinh((double)(length - i) / sigma) / sinh((double)lt / sigma);
      pv = ((double)(length - i) - half) * ((double)(length - i) - half) - half * half;
    }
    cusp[i] = cv;
    par[i] = pv;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ap = 0.0, ac = 0.0;
    for (int i = 0; i < length; i++) { ap += par[i]; ac += cusp[i]; }
    zs[2 * length] = ap;
    zs[2 * length + 1] = ac;
  }
  __syncthreads();
  const double ap = zs[2 * length], ac = zs[2 * length + 1];
  const double c = exp(-1.0 / decay);
  for (int i = threadIdx.x; i < length; i += blockDim.x) {
    const double z1 = cusp[i] + (-par[i] / ap * ac);
    const double z0 = i > 0 ? cusp[i - 1] + (-par[i - 1] / ap * ac) : 0.0;
    kernel[i] = (T)__dadd_rn(z1, __dmul_rn(z0, -c));
  }
}
template <typename T>
__global__ void k_t0(double rise, double fall, T* kernel, int length) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= length) return;
  const int ri = (int)rise;
  kernel[i] = i < ri ? (T)(2.0 * (double)(ri - i) / (rise * (rise + 1.0))) : (T)(-1.0 / fall);
}

int trap_static_check(int64_t n, int32_t rise, int32_t flat) {
  if (rise < 0) return DSPB_FATAL_RISE_NEG;
  if (flat < 0) return DSPB_FATAL_FLAT_NEG;
  if (2 * (int64_t)rise + flat > n) return DSPB_FATAL_TRAP_WIDE;
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------
// extern "C" launchers
// ------------------------------------------------------------------------------------
#define WIN(name) mk_wave(name, name##_row_stride, name##_dtype)
#define WOUT(name) mk_out<T>(name, name##_row_stride)
#define SC(name) mk_scalar<T>(name, name##_stride, name##_imm)

#define DEFINE_ALL(T_, SFX)                                                                        \
  extern "C" int dspb_bl_subtract##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,              \
                                       DSPB_SCALAR(a_baseline), DSPB_WAVE_OUT(w_out), DSPB_TAIL) { \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    BlSubtract<T> b{WIN(w_in), (int)n, SC(a_baseline), WOUT(w_out)};                               \
    return launch_rows<T>(b, n_rows, 1, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_min_max##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, void* t_min,     \
                                   void* t_max, void* a_min, void* a_max, DSPB_TAIL) {             \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    MinMax<T> b{WIN(w_in), (int)n, (T*)t_min, (T*)t_max, (T*)a_min, (T*)a_max, false};             \
    return launch_rows<T>(b, n_rows, 1, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_amax##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, void* a_max,        \
                                DSPB_TAIL) {                                                       \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    MinMax<T> b{WIN(w_in), (int)n, nullptr, nullptr, nullptr, (T*)a_max, true};                    \
    return launch_rows<T>(b, n_rows, 1, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_min_max_norm##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,             \
                                        DSPB_SCALAR(a_min), DSPB_SCALAR(a_max),                    \
                                        DSPB_WAVE_OUT(w_out), DSPB_TAIL) {                         \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    MinMaxNorm<T> b{WIN(w_in), (int)n, SC(a_min), SC(a_max), WOUT(w_out)};                         \
    return launch_rows<T>(b, n_rows, 2, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_linear_slope_fit##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,         \
                                            void* mean, void* stdev, void* slope,                  \
                                            void* intercept, DSPB_TAIL) {                          \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    LinearSlopeFit<T> b{WIN(w_in), (int)n, (T*)mean, (T*)stdev, (T*)slope, (T*)intercept};         \
    return launch_rows<T>(b, n_rows, 1, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_linear_slope_diff##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,        \
                                             DSPB_SCALAR(slope), DSPB_SCALAR(intercept),           \
                                             void* mean, void* rms, DSPB_TAIL) {                   \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    LinearSlopeDiff<T> b{WIN(w_in), (int)n, SC(slope), SC(intercept), (T*)mean, (T*)rms};          \
    return launch_rows<T>(b, n_rows, 1, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_mean_below_threshold##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,     \
                                                DSPB_SCALAR(threshold), void* result,              \
                                                DSPB_TAIL) {                                       \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    MeanBelowThreshold<T> b{WIN(w_in), (int)n, SC(threshold), (T*)result};                         \
    return launch_rows<T>(b, n_rows, 1, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_pole_zero##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,                \
                                     DSPB_SCALAR(t_tau), DSPB_WAVE_OUT(w_out), DSPB_TAIL) {        \
    using T = T_;                                                                                  \
    PoleZero<T> b{WIN(w_in), (int)n, SC(t_tau), WOUT(w_out), fatal};                               \
    return launch_rows<T>(b, n_rows, 2, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_double_pole_zero##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,         \
                                            DSPB_SCALAR(t_tau1), DSPB_SCALAR(t_tau2),              \
                                            DSPB_SCALAR(frac), DSPB_WAVE_OUT(w_out), DSPB_TAIL) {  \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    if (n <= 3) return DSPB_FATAL_DPZ_SHORT;                                                       \
    DoublePoleZero<T> b{WIN(w_in), (int)n, SC(t_tau1), SC(t_tau2), SC(frac), WOUT(w_out)};         \
    return launch_rows<T>(b, n_rows, 2, n, stream, MAXW * sizeof(Aff2) + 16);                        \
  }                                                                                                \
  extern "C" int dspb_trap_filter##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,              \
                                       int32_t rise, int32_t flat, int32_t norm,                   \
                                       DSPB_WAVE_OUT(w_out), DSPB_TAIL) {                          \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    if (int rc = trap_static_check(n, rise, flat)) return rc;                                      \
    Trap<T> b{WIN(w_in), (int)n, rise, flat, 0, norm ? 1 : 0, WOUT(w_out)};                        \
    return launch_rows<T>(b, n_rows, 2, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_asym_trap_filter##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,         \
                                            int32_t rise, int32_t flat, int32_t fall,              \
                                            DSPB_WAVE_OUT(w_out), DSPB_TAIL) {                     \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    if (rise < 0) return DSPB_FATAL_RISE_NEG;                                                      \
    if (flat < 0) return DSPB_FATAL_FLAT_NEG;                                                      \
    if (fall < 0) return DSPB_FATAL_FALL_NEG;                                                      \
    if ((int64_t)rise + flat + fall > n) return DSPB_FATAL_TRAP_WIDE;                              \
    Trap<T> b{WIN(w_in), (int)n, rise, flat, fall, 2, WOUT(w_out)};                                \
    return launch_rows<T>(b, n_rows, 2, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_trap_pickoff##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,             \
                                        int32_t rise, int32_t flat, DSPB_SCALAR(t_pickoff),        \
                                        void* a_out, DSPB_TAIL) {                                  \
    using T = T_;                                                                                  \
    if (int rc = trap_static_check(n, rise, flat)) return rc;                                      \
    TrapPickoff<T> b{WIN(w_in), (int)n, rise, flat, SC(t_pickoff), (T*)a_out, fatal};              \
    return launch_rows<T>(b, n_rows, 1, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_moving_window##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,            \
                                         double length, int32_t kind, DSPB_WAVE_OUT(w_out),        \
                                         DSPB_TAIL) {                                              \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    const T len = (T)length;                                                                       \
    if (!(len >= (T)0) || !(len < (T)n)) return DSPB_FATAL_MW_RANGE;                               \
    MovingWindow<T> b{WIN(w_in), (int)n, len, kind ? 1 : 0, 0, 0, WOUT(w_out)};                    \
    return launch_rows<T>(b, n_rows, 2, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_moving_window_multi##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,      \
                                               double length, double num_mw, int32_t mw_type,      \
                                               DSPB_WAVE_OUT(w_out), DSPB_TAIL) {                  \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    const T len = (T)length, num = (T)num_mw;                                                      \
    if (std::floor((double)len) != (double)len) return DSPB_FATAL_MWM_LEN_NONINT;                  \
    if (std::floor((double)num) != (double)num) return DSPB_FATAL_MWM_NUM_NONINT;                  \
    if ((long long)len < 0 || (long long)len >= n) return DSPB_FATAL_MWM_RANGE;                    \
    if ((long long)num < 0) return DSPB_FATAL_MWM_NUM_NEG;                                         \
    MovingWindow<T> b{WIN(w_in), (int)n, len, 2, (int)num, mw_type, WOUT(w_out)};                  \
    return launch_rows<T>(b, n_rows, 3, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_avg_current##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,              \
                                       double length, DSPB_WAVE_OUT(w_out), int64_t n_out,         \
                                       DSPB_TAIL) {                                                \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    const T len = (T)length;                                                                       \
    if (!(len >= (T)0) || !(len < (T)n)) return DSPB_FATAL_MW_RANGE;                               \
    if (n_out != n - (long long)len) return DSPB_FATAL_SHAPE;                                      \
    AvgCurrent<T> b{WIN(w_in), (int)n, len, WOUT(w_out), (int)n_out};                              \
    return launch_rows<T>(b, n_rows, 1, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_time_point_thresh##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,        \
                                             DSPB_SCALAR(a_threshold), DSPB_SCALAR(t_start),       \
                                             DSPB_SCALAR(walk_forward), void* t_out, DSPB_TAIL) {  \
    using T = T_;                                                                                  \
    TimePointThresh<T> b{WIN(w_in), (int)n, SC(a_threshold), SC(t_start), SC(walk_forward),        \
                         (T*)t_out, fatal, false, 0, 0};                                           \
    return launch_rows<T>(b, n_rows, 1, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_interpolated_time_point_thresh##SFX(                                         \
      DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, DSPB_SCALAR(a_threshold),                     \
      DSPB_SCALAR(t_start), int64_t walk_forward, int32_t mode_in, void* t_out, DSPB_TAIL) {       \
    using T = T_;                                                                                  \
    TimePointThresh<T> b{WIN(w_in), (int)n, SC(a_threshold), SC(t_start), SC(t_start),             \
                         (T*)t_out, fatal, true, walk_forward, mode_in};                           \
    return launch_rows<T>(b, n_rows, 1, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_multi_time_point_thresh##SFX(                                                \
      DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, const void* a_threshold, int64_t m,           \
      int64_t thr_row_stride, DSPB_SCALAR(t_start), double polarity, int32_t mode_in,              \
      void* t_out, DSPB_TAIL) {                                                                    \
    using T = T_;                                                                                  \
    if (polarity == 0.0) return DSPB_FATAL_POLARITY_ZERO;                                          \
    MultiTimePointThresh<T> b{WIN(w_in), (int)n, (const T*)a_threshold, (int)m, thr_row_stride,    \
                              SC(t_start), polarity > 0 ? 1 : -1, mode_in, (T*)t_out, fatal};      \
    return launch_rows<T>(b, n_rows, 1, n, stream, (size_t)m * sizeof(int) + 16);                  \
  }                                                                                                \
  extern "C" int dspb_fixed_time_pickoff##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,       \
                                              DSPB_SCALAR(t_in), int32_t mode_in, void* a_out,     \
                                              DSPB_TAIL) {                                         \
    using T = T_;                                                                                  \
    FixedTimePickoff<T> b{WIN(w_in), (int)n, SC(t_in), mode_in, (T*)a_out, fatal};                 \
    return launch_rows<T>(b, n_rows, 1, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_windower##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,                 \
                                    DSPB_SCALAR(t0_in), DSPB_WAVE_OUT(w_out), int64_t m,           \
                                    DSPB_TAIL) {                                                   \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    if (m >= n) return DSPB_FATAL_WINDOWER_LEN;                                                    \
    Windower<T> b{WIN(w_in), (int)n, SC(t0_in), WOUT(w_out), (int)m};                              \
    return launch_rows<T>(b, n_rows, 1, n, stream);                                                \
  }                                                                                                \
  extern "C" int dspb_upsampler##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,                \
                                     double upsample, DSPB_WAVE_OUT(w_out), int64_t m,             \
                                     DSPB_TAIL) {                                                  \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    const T up = (T)upsample;                                                                      \
    if (!(up > (T)0)) return DSPB_FATAL_UPSAMPLE;                                                  \
    Upsampler<T> b{WIN(w_in), (int)n, up, WOUT(w_out), (int)m};                                    \
    return launch_rows<T>(b, n_rows, 2, n > m ? n : m, stream);                                    \
  }                                                                                                \
  extern "C" int dspb_get_multi_local_extrema##SFX(                                                \
      DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, double a_delta_max, double a_delta_min,       \
      double search_direction, DSPB_SCALAR(a_abs_max), DSPB_SCALAR(a_abs_min), void* vt_max,       \
      void* vt_min,                                                                                \
      int64_t m, uint32_t* n_max, uint32_t* n_min, DSPB_TAIL) {                                    \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    const T dir = (T)search_direction;                                                             \
    if (a_delta_max == a_delta_max && a_delta_min == a_delta_min) {                                \
      if (!(m < n)) return DSPB_FATAL_GMLE_LEN;                                                    \
      if (!((T)a_delta_max >= 0) || !((T)a_delta_min >= 0)) return DSPB_FATAL_GMLE_DELTA;          \
      if (!(dir == 0 || dir == 1 || dir == 2 || dir == 3)) return DSPB_FATAL_GMLE_DIR;             \
    }                                                                                              \
    MultiLocalExtrema<T> b{WIN(w_in), (int)n, (T)a_delta_max, (T)a_delta_min, dir, SC(a_abs_max),  \
                           SC(a_abs_min), (T*)vt_max, (T*)vt_min, (int)m, n_max, n_min};           \
    return launch_rows<T>(b, n_rows, 1, n, stream, 4 * (size_t)m * sizeof(float) + 16);            \
  }                                                                                                \
  extern "C" int dspb_recursive_filter##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n,         \
                                            const double* a, int64_t p, const double* b_,          \
                                            int64_t q, DSPB_SCALAR(init_in),                       \
                                            DSPB_SCALAR(init_out), DSPB_WAVE_OUT(w_out),           \
                                            DSPB_TAIL) {                                           \
    using T = T_;                                                                                  \
    (void)fatal;                                                                                   \
    if (q == 0) return DSPB_FATAL_RF_B_SCALAR;                                                     \
    if (n <= q) return DSPB_FATAL_RF_SHORT;                                                        \
    if (q > 3 || p > 8) return DSPB_ERR_UNSUPPORTED;                                               \
    RecursiveFilter<T> b{};                                                                        \
    b.in = WIN(w_in);                                                                              \
    b.n = (int)n;                                                                                  \
    for (int j = 0; j < p; j++) b.a[j] = a[j];                                                     \
    b.p = (int)p;                                                                                  \
    for (int j = 0; j < q; j++) b.b[j] = b_[j];                                                    \
    b.q = (int)q;                                                                                  \
    b.init_in = SC(init_in);                                                                       \
    b.init_out = SC(init_out);                                                                     \
    b.out = WOUT(w_out);                                                                           \
    return launch_rows<T>(b, n_rows, 2, n, stream, MAXW * sizeof(Aff2) + 16);                        \
  }                                                                                                \
  extern "C" int dspb_cusp_filter##SFX(double sigma, double flat, double decay, void* kernel,      \
                                       int64_t length, void* stream) {                             \
    using T = T_;                                                                                  \
    if (sigma < 0 || flat < 0 || std::floor(flat) != flat || decay < 0)                            \
      return DSPB_FATAL_KERNEL_ARGS;                                                               \
    const int lt = (int)(((double)length - flat) / 2.0);                                           \
    T* tmp = nullptr;                                                                              \
    cudaError_t e = cudaMallocAsync(&tmp, sizeof(T) * (size_t)length, (cudaStream_t)stream);       \
    if (e != cudaSuccess) return -(int)e;                                                          \
    const int g = (int)((length + 255) / 256);                                                     \
    k_cusp_stage1<T><<<g, 256, 0, (cudaStream_t)stream>>>(sigma, (int)flat, lt, tmp, (int)length); \
    k_diff_same<T><<<g, 256, 0, (cudaStream_t)stream>>>(tmp, std::exp(-1.0 / decay), (T*)kernel,   \
                                                        (int)length);                              \
    cudaFreeAsync(tmp, (cudaStream_t)stream);                                                      \
    e = cudaGetLastError();                                                                        \
    return e == cudaSuccess ? 0 : -(int)e;                                                         \
  }                                                                                                \
  extern "C" int dspb_zac_filter##SFX(double sigma, double flat, double decay, void* kernel,       \
                                      int64_t length, void* stream) {                              \
    using T = T_;                                                                                  \
    if (sigma < 0 || flat < 0 || std::floor(flat) != flat || decay < 0)                            \
      return DSPB_FATAL_KERNEL_ARGS;                                                               \
    const int lt = (int)(((double)length - flat) / 2.0);                                           \
    const size_t smem = sizeof(double) * (2 * (size_t)length + 2);                                 \
    if (smem > MAX_SMEM) return DSPB_ERR_ROW_TOO_LONG;                                             \
    cudaError_t e = cudaFuncSetAttribute(k_zac<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                         (int)smem);                                               \
    if (e != cudaSuccess) return -(int)e;                                                          \
    k_zac<T><<<1, 1024, smem, (cudaStream_t)stream>>>(sigma, (int)flat, lt, decay, (T*)kernel,     \
                                                      (int)length);                                \
    e = cudaGetLastError();                                                                        \
    return e == cudaSuccess ? 0 : -(int)e;                                                         \
  }                                                                                                \
  extern "C" int dspb_t0_filter##SFX(double rise, double fall, void* kernel, int64_t length,       \
                                     void* stream) {                                               \
    using T = T_;                                                                                  \
    if (rise < 0 || fall < 0 || (double)length != rise + fall) return DSPB_FATAL_KERNEL_ARGS;      \
    k_t0<T><<<(int)((length + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rise, fall, (T*)kernel, \
                                                                           (int)length);           \
    cudaError_t e = cudaGetLastError();                                                            \
    return e == cudaSuccess ? 0 : -(int)e;                                                         \
  }

DEFINE_ALL(float, _f32)
DEFINE_ALL(double, _f64)

extern "C" int dspb_version(void) { return 100; }

extern "C" int64_t dspb_max_row_len(int elem_bytes, int n_slots) {
  if (elem_bytes <= 0 || n_slots <= 0) return 0;
  const int64_t words = (int64_t)(MAX_SMEM - SCRATCH_BYTES) / elem_bytes / n_slots;
  return (words - 1) * 32 / 33;
}

extern "C" const char* dspb_fatal_message(int code) {
  switch (code) {
    case DSPB_FATAL_PZ_NAN: return "Pole-zero filter produced nans in output.";
    case DSPB_FATAL_DPZ_SHORT: return "The length of the waveform must be larger than 3 for the filter to work safely";
    case DSPB_FATAL_RISE_NEG: return "The number of samples in the rise section must be positive";
    case DSPB_FATAL_FLAT_NEG: return "The number of samples in the flat section must be positive";
    case DSPB_FATAL_FALL_NEG: return "The number of samples in the fall section must be positive";
    case DSPB_FATAL_TRAP_WIDE: return "The trapezoid width is wider than the waveform";
    case DSPB_FATAL_PICKOFF_NONINT: return "The pick-off index must be an integer";
    case DSPB_FATAL_MW_RANGE: return "length is out of range, must be between 0 and the length of the waveform";
    case DSPB_FATAL_MWM_LEN_NONINT: return "The length of the moving window must be an integer";
    case DSPB_FATAL_MWM_NUM_NONINT: return "The number of moving windows must be an integer";
    case DSPB_FATAL_MWM_RANGE: return "The length of the moving window is out of range";
    case DSPB_FATAL_MWM_NUM_NEG: return "The number of moving windows much be positive";
    case DSPB_FATAL_TSTART_NONINT: return "The starting index must be an integer";
    case DSPB_FATAL_WALK_NONINT: return "The search direction must be an integer";
    case DSPB_FATAL_TSTART_RANGE: return "The starting index is out of range";
    case DSPB_FATAL_INTERP_MODE: return "Unrecognized interpolation mode";
    case DSPB_FATAL_POLARITY_ZERO: return "polarity cannot be 0";
    case DSPB_FATAL_FTP_INT: return "fixed_time_pickoff requires integer t_in when using mode 'i'";
    case DSPB_FATAL_WINDOWER_LEN: return "The windowed waveform must be smaller than the input waveform";
    case DSPB_FATAL_UPSAMPLE: return "Upsample must be greater than 0";
    case DSPB_FATAL_CONV_KERNEL_LONG: return "The filter is longer than the input waveform";
    case DSPB_FATAL_CONV_MODE: return "Invalid mode";
    case DSPB_FATAL_CONV_OUTLEN: return "Output waveform has the wrong length";
    case DSPB_FATAL_GMLE_LEN: return "The length of your return array must be smaller than the length of your waveform";
    case DSPB_FATAL_GMLE_DELTA: return "Delta must be positive";
    case DSPB_FATAL_GMLE_DIR: return "search direction type not found.";
    case DSPB_FATAL_RF_B_SCALAR: return "b cannot be scalar";
    case DSPB_FATAL_RF_SHORT: return "The length of the waveform must be larger than len(b) for the filter to work safely";
    case DSPB_FATAL_SHAPE: return "array shapes do not match the processor signature";
    case DSPB_FATAL_KERNEL_ARGS: return "invalid kernel-generator arguments";
    case DSPB_FATAL_HIST_LEN: return "length borders_out must be exactly 1 + length of weights_out";
    case DSPB_FATAL_HIST_NAN: return "input data contains nan";
    case DSPB_FATAL_HPS_NAN: return "nan in input weights";
    case DSPB_FATAL_HPS_LEN: return "length edges_in must be exactly 1 + length of weights_in";
    case DSPB_FATAL_HPS_WIDTH_TYPE: return "Unknown width_type, must be [0...4]";
    case DSPB_FATAL_RCCR2_NAN: return "RC-CR^2 filter produced nans in output.";
    case DSPB_FATAL_INJ_FRAC: return "frac must be between zero and one.";
    case DSPB_ERR_ROW_TOO_LONG: return "waveform too long for the shared-memory resident layout";
    case DSPB_ERR_UNSUPPORTED: return "argument combination not supported by the device implementation";
  }
  return "unknown";
}
