// dspeed_b200 -- the remaining processors of the reference's SiPM / LAr chain (tests/configs/sipm-dsp-config.json):
// reflected_convolve_wf, histogram, histogram_around_mode, histogram_stats, histogram_peakstats,
// peak_snr_threshold, multi_a_filter, and the set-up time kernel generators gaussian_filter1d / moving_slope / step.
//
// These run on short waveforms and small per-event lists (a few hundred bins, <= 20 peak candidates): one CTA per
// row for the two passes over the samples (histograms, convolution), one thread per row for the serial scans over
// bins / candidates.  Type handling mirrors what numba compiles from the reference (pinned by the float32 goldens):
// float32 (-) float32 stays float32, a float32 / int quotient and np.linspace are float64, results are rounded when
// they are stored into the loop's output type.
#include <cmath>

#include "common.cuh"

using namespace dspb;

namespace {

constexpr int TPB = 128;

template <typename T>
__device__ __forceinline__ T ldw(const void* p, int dt, long long i) {
  switch (dt) {
    case DSPB_F32: return (T) reinterpret_cast<const float*>(p)[i];
    case DSPB_F64: return (T) reinterpret_cast<const double*>(p)[i];
    case DSPB_U16: return (T) reinterpret_cast<const uint16_t*>(p)[i];
    case DSPB_I16: return (T) reinterpret_cast<const int16_t*>(p)[i];
    case DSPB_I32: return (T) reinterpret_cast<const int32_t*>(p)[i];
    case DSPB_U32: return (T) reinterpret_cast<const uint32_t*>(p)[i];
  }
  return (T)0;
}

template <typename T>
struct SIn {   // per-row scalar argument: device array (stride 0 broadcasts) or immediate
  const T* p;
  long long stride;
  T imm;
  __device__ __forceinline__ T get(long long row) const { return p ? p[row * stride] : imm; }
};
template <typename T>
SIn<T> mk_sin(const void* p, int64_t stride, double imm) {
  return SIn<T>{reinterpret_cast<const T*>(p), stride, (T)imm};
}

// start + i * step with both roundings (np.linspace / numba evaluate the product and the sum separately; the compiler
// would contract them into one FMA)
template <typename T>
__device__ __forceinline__ double lin(T start, int i, double step) {
  return __dadd_rn((double)start, __dmul_rn((double)i, step));
}

// block-wide min / max / any-NaN of a row (TPB threads); results broadcast through shared memory
template <typename T>
__device__ void row_minmax(const void* w, int dt, long long base, int n, T& mn, T& mx, int& has_nan, T* sh) {
  T lo = (T)INFINITY, hi = (T)-INFINITY;
  int bad = 0;
  for (int i = threadIdx.x; i < n; i += TPB) {
    const T v = ldw<T>(w, dt, base + i);
    bad |= (v != v);
    lo = v < lo ? v : lo;
    hi = v > hi ? v : hi;
  }
  __shared__ int sbad;
  if (threadIdx.x == 0) sbad = 0;
  __syncthreads();
  if (bad) atomicOr(&sbad, 1);
  sh[threadIdx.x] = lo;
  sh[TPB + threadIdx.x] = hi;
  __syncthreads();
  for (int s = TPB / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      const T a = sh[threadIdx.x + s], b = sh[TPB + threadIdx.x + s];
      if (a < sh[threadIdx.x]) sh[threadIdx.x] = a;
      if (b > sh[TPB + threadIdx.x]) sh[TPB + threadIdx.x] = b;
    }
    __syncthreads();
  }
  mn = sh[0];
  mx = sh[TPB];
  has_nan = sbad;
  __syncthreads();
}

// histogram.py:14-89 (MODE = false) and histogram.py:92-204 (MODE = true); counts in shared memory
template <typename T, bool MODE>
__global__ void __launch_bounds__(TPB) k_histogram(const void* w, long long w_rs, int w_dt, long long n_rows, int n,
                                                   SIn<T> center_in, SIn<T> bw_in, T* weights, long long wt_rs, T* borders,
                                                   long long bd_rs, int nb, int* fatal) {
  extern __shared__ __align__(16) unsigned char smem[];
  int* cnt = reinterpret_cast<int*>(smem);
  T* sh = reinterpret_cast<T*>(cnt + ((nb + 2) & ~1));   // (8-byte aligned for the float64 loop)
  __shared__ double s_center;
  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const long long base = row * w_rs;
    T* wo = weights + row * wt_rs;
    T* bo = borders + row * bd_rs;
    T mn, mx;
    int has_nan;
    const T c_in = MODE ? center_in.get(row) : (T)NAN;
    const bool need_pass1 = !MODE || (c_in != c_in);
    row_minmax<T>(w, w_dt, base, n, mn, mx, has_nan, sh);
    if (has_nan) {
      if (MODE) {
        if (threadIdx.x == 0) raise_fatal(fatal, DSPB_FATAL_HIST_NAN, row);
      } else {
        for (int i = threadIdx.x; i < nb; i += TPB) wo[i] = (T)0;
        for (int i = threadIdx.x; i <= nb; i += TPB) bo[i] = (T)NAN;
      }
      __syncthreads();
      continue;
    }
    for (int i = threadIdx.x; i <= nb; i += TPB) cnt[i] = 0;
    __syncthreads();
    const double delta = (double)(T)(mx - mn) / (double)nb;
    // np.linspace(wf_min, wf_max, nb + 1): float64 end points, last element = stop
    const double step = ((double)mx - (double)mn) / (double)nb;
    const T b0 = mn;   // borders[0] = (T)(double)mn
    if (need_pass1) {
      if (!MODE)
        for (int i = threadIdx.x; i <= nb; i += TPB) bo[i] = i == nb ? mx : (T)lin(mn, i, step);
      if (delta != 0.0) {
        for (int i = threadIdx.x; i < n; i += TPB) {
          const T v = ldw<T>(w, w_dt, base + i);
          if (v == mx) continue;
          const double q = floor((double)(T)(v - b0) / delta);
          if (q >= 0.0 && q < (double)nb) atomicAdd(&cnt[(int)q], 1);
        }
      }
      __syncthreads();
    }
    if (!MODE) {
      for (int i = threadIdx.x; i < nb; i += TPB) wo[i] = (T)cnt[i];
      __syncthreads();
      continue;
    }
    // ---- histogram around the mode ----------------------------------------------------------------
    const T bw = bw_in.get(row);
    if (threadIdx.x == 0) {
      double c = (double)c_in;
      if (c_in != c_in) {
        if (delta == 0.0) c = (double)mn;
        else {
          int am = 0;
          for (int i = 1; i < nb; i++)
            if (cnt[i] > cnt[am]) am = i;
          const T bam = am == nb ? mx : (T)lin(mn, am, step);
          c = (double)bam + 0.5 * delta;
          c = rint(c / (double)bw) * (double)bw;
        }
      }
      s_center = c;
    }
    __syncthreads();
    // (explicit roundings: numpy evaluates every product and sum separately, the compiler would contract them into FMAs)
    const double hist_min = __dsub_rn(__dsub_rn(s_center, __dmul_rn((double)bw, (double)(nb / 2))), 0.5 * (double)bw);
    for (int i = threadIdx.x; i <= nb; i += TPB) {
      cnt[i] = 0;
      bo[i] = (T)__dadd_rn(hist_min, __dmul_rn((double)bw, (double)i));
    }
    __syncthreads();
    const T e0 = (T)hist_min;
    for (int i = threadIdx.x; i < n; i += TPB) {
      const T v = ldw<T>(w, w_dt, base + i);
      const T q = floor((T)((T)(v - e0) / bw));
      if (q >= (T)0 && q < (T)nb) atomicAdd(&cnt[(int)q], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += TPB) wo[i] = (T)cnt[i];
    __syncthreads();
  }
}

// histogram_stats.py:146-261 -- one thread per row
template <typename T>
__global__ void k_histogram_stats(const T* wts, long long w_rs, const T* edg, long long e_rs, long long n_rows, int nb,
                                  SIn<T> max_in, T* mode_out, T* max_out, T* fwhm_out) {
  const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const T* w = wts + row * w_rs;
  const T* e = edg + row * e_rs;
  T fw = (T)NAN;
  mode_out[row] = (T)NAN;
  max_out[row] = (T)NAN;
  fwhm_out[row] = (T)NAN;
  for (int i = 0; i < nb; i++)
    if (w[i] != w[i]) return;
  const T mi_in = max_in.get(row);
  int mi = 0;
  if (mi_in != mi_in) {
    for (int i = 0; i < nb; i++)
      if (w[i] > w[mi]) mi = i;
  } else if (mi_in > e[nb - 1]) {
    mi = nb - 1;
  } else {
    for (int i = 0; i < nb; i++)
      if (fabs((T)(mi_in - e[i])) < fabs((T)(mi_in - e[mi]))) mi = i;
  }
  const T mo = e[mi];
  const double half = 0.5 * (double)w[mi];
  for (int i = mi; i < nb; i++)
    if ((double)w[i] <= half && w[i] != (T)0) { fw = fabs((T)(mo - e[i])); break; }
  for (int i = 0; i < mi; i++)
    if ((double)w[i] >= half && w[i] != (T)0) {
      const T d = fabs((T)(mo - e[i]));
      if (fw < d) fw = d;
      break;
    }
  mode_out[row] = (T)mi;
  max_out[row] = mo;
  fwhm_out[row] = fw;
}

// histogram_stats.py:12-143 -- one thread per row
template <typename T>
__global__ void k_histogram_peakstats(const T* wts, long long w_rs, const T* edg, long long e_rs, long long n_rows, int nb,
                                      SIn<T> max_in, int skip_zeroes, int width_type, T* mode_out, T* width_out, int* fatal) {
  const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const T* w = wts + row * w_rs;
  const T* e = edg + row * e_rs;
  mode_out[row] = (T)NAN;
  width_out[row] = (T)NAN;
  for (int i = 0; i < nb; i++)
    if (w[i] != w[i]) { raise_fatal(fatal, DSPB_FATAL_HPS_NAN, row); return; }
  const T mi_in = max_in.get(row);
  int mi = 0;
  if (mi_in != mi_in) {
    for (int i = 0; i < nb; i++)
      if (w[i] > w[mi]) mi = i;
  } else if (mi_in > e[nb]) {
    mi = nb - 1;
  } else if (mi_in < e[0]) {
    mi = 0;
  } else {
    for (int i = 0; i < nb; i++)
      if (e[i] <= mi_in && mi_in < e[i + 1]) { mi = i; break; }
  }
  const T mode = (T)((double)e[mi] + 0.5 * (double)(T)(e[mi + 1] - e[mi]));
  const double half = 0.5 * (double)w[mi];
  T right = fabs((T)(mode - e[nb])), left = fabs((T)(mode - e[0]));
  for (int i = mi; i < nb; i++) {
    if (skip_zeroes && w[i] == (T)0) continue;
    if ((double)w[i] <= half) { right = fabs((T)(mode - e[i])); break; }
  }
  for (int i = mi; i >= 0; i--) {
    if (skip_zeroes && w[i] == (T)0) continue;
    if ((double)w[i] <= half) { left = fabs((T)(mode - e[i + 1])); break; }
  }
  T wd;
  switch (width_type) {
    case 0: wd = (T)((double)left + (double)right); break;
    case 1: wd = left < right ? left : right; break;
    case 2: wd = left > right ? left : right; break;
    case 3: wd = left; break;
    default: wd = right; break;
  }
  mode_out[row] = mode;
  width_out[row] = wd;
}

// peak_snr_threshold.py:11-71 -- one thread per row (<= a few dozen candidates, windows of 2 * width samples)
template <typename T>
__global__ void k_peak_snr(const void* w, long long w_rs, int w_dt, long long n_rows, int n, const T* idx_in, long long i_rs,
                           int m, SIn<T> ratio_in, SIn<T> width_in, T* idx_out, long long o_rs, uint32_t* n_out) {
  const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const long long base = row * w_rs;
  const T* ii = idx_in + row * i_rs;
  T* oo = idx_out + row * o_rs;
  const T ratio = ratio_in.get(row);
  const int width = (int)width_in.get(row);
  for (int i = 0; i < m; i++) oo[i] = (T)NAN;
  int k = 0;
  for (int i = 0; i < m; i++) {
    const T t = ii[i];
    if (t != t) continue;
    const int c = (int)t;
    if (c < 0 || c >= n) continue;   // (the reference would index out of bounds)
    int a = c - width, b = c + width;
    if (a < 0) a = 0;
    if (b >= n) b = n - 1;
    T vmin = ldw<T>(w, w_dt, base + a);
    for (int j = a; j < b; j++) {
      const T v = ldw<T>(w, w_dt, base + j);
      if (v < vmin) vmin = v;
    }
    const T vc = ldw<T>(w, w_dt, base + c);
    if (fabs((T)(vmin / vc)) < ratio) oo[k++] = t;
  }
  n_out[row] = (uint32_t)k;
}

// multi_a_filter.py:11-57 -- one thread per row; NaN waveform -> all NaN (checked by the block of the row)
template <typename T>
__global__ void __launch_bounds__(TPB) k_multi_a(const void* w, long long w_rs, int w_dt, long long n_rows, int n, const T* vt,
                                                 long long v_rs, int m, T* out, long long o_rs, int* fatal) {
  __shared__ int sbad;
  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const long long base = row * w_rs;
    if (threadIdx.x == 0) sbad = 0;
    __syncthreads();
    int bad = 0;
    for (int i = threadIdx.x; i < n; i += TPB) {
      const T v = ldw<T>(w, w_dt, base + i);
      bad |= (v != v);
    }
    if (bad) atomicOr(&sbad, 1);
    __syncthreads();
    const T* t_in = vt + row * v_rs;
    T* o = out + row * o_rs;
    // first_nan: where the NaN padding starts, if everything behind it is NaN as well (else the whole list is used)
    int first = m;
    if (!sbad) {
      int f = -1;
      for (int i = 0; i < m; i++)
        if (t_in[i] != t_in[i]) { f = i; break; }
      if (f >= 0) {
        bool tail_nan = true;
        for (int i = f; i < m; i++) tail_nan &= (t_in[i] != t_in[i]);
        if (tail_nan) first = f;
      }
    }
    for (int i = threadIdx.x; i < m; i += TPB) {
      T r = (T)NAN;
      if (!sbad && i < first) {
        const T t = t_in[i];
        if (t == t && t >= (T)0 && t <= (T)(n - 1)) {
          const int k = (int)t;
          if ((T)k == t) r = ldw<T>(w, w_dt, base + k);
          else raise_fatal(fatal, DSPB_FATAL_FTP_INT, row);
        }
      }
      o[i] = r;
    }
    __syncthreads();
  }
}

// convolutions.py:122-182 -- y[o] = sum_j kernel[j] * e[o + ext + (m-1)/2 - j], e = reflect-padded waveform
template <typename T>
__global__ void __launch_bounds__(TPB) k_reflected_conv(const void* w, long long w_rs, int w_dt, long long n_rows, int n,
                                                        const T* kern, int m, T* out, long long o_rs) {
  extern __shared__ __align__(16) unsigned char smem[];
  T* ks = reinterpret_cast<T*>(smem);
  T* xs = ks + m;
  __shared__ int sbad;
  const int ext = m / 2 + 1, np_ = n + 2 * ext;
  int kbad = 0;
  for (int j = threadIdx.x; j < m; j += TPB) {
    ks[j] = kern[j];
    kbad |= (ks[j] != ks[j]);
  }
  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const long long base = row * w_rs;
    if (threadIdx.x == 0) sbad = 0;
    __syncthreads();
    int bad = kbad;
    for (int t = threadIdx.x; t < np_; t += TPB) {
      int i = t - ext;             // numpy 'reflect': ... 2 1 | 0 1 2 ... n-1 | n-2 n-3 ...
      while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
      const T v = ldw<T>(w, w_dt, base + i);
      bad |= (v != v);
      xs[t] = v;
    }
    if (bad) atomicOr(&sbad, 1);
    __syncthreads();
    T* o = out + row * o_rs;
    const int off = ext + (m - 1) / 2;
    for (int q = threadIdx.x; q < n; q += TPB) {
      T r = (T)NAN;
      if (!sbad) {
        double acc = 0.0;
        for (int j = 0; j < m; j++) acc += (double)ks[j] * (double)xs[q + off - j];
        r = (T)acc;
      }
      o[q] = r;
    }
    __syncthreads();
  }
}

// ---- set-up time kernel generators ------------------------------------------------------------------
// gaussian_filter1d.py:46-82 (sigma, truncate arrive rounded to the loop's type; the weights are float64 math)
template <typename T>
__global__ void k_gaussian(double coef, int lw, T* out, int n) {
  double s = 0.0;
  for (int i = 0; i < n; i++) {
    const double x = (double)(i - lw);
    s += exp(coef * (x * x));
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = (double)(i - lw);
    out[i] = (T)(exp(coef * (x * x)) / s);
  }
}
// kernels.py:64-103
template <typename T>
__global__ void k_moving_slope(T* out, int n) {
  const double L = (double)n, sx = L * (L + 1.0) / 2.0, sx2 = L * (L + 1.0) * (2.0 * L + 1.0) / 6.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    T v = (T)((double)(i + 1) * L - sx);          // kernel[:] = ... (stored in the kernel's type)
    v = (T)((double)v / (L * sx2 - sx * sx));      // kernel[:] /= ...
    out[n - 1 - i] = v;                            // kernel[:] = kernel[::-1]
  }
}
// kernels.py:106-142
template <typename T>
__global__ void k_step(T* out, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = (double)i, L = (double)n;
    out[i] = (T)((x >= L / 4.0 && x < 3.0 * L / 4.0) ? 1.0 : -1.0);
  }
}

// ---- VectorOfVectors output compaction (LGDOVectorOfVectorsIOManager.write, processing_chain.py:2230-2260) ---------
// A variable-length output lives in the chain as a padded [rows, width] block plus a length variable; the ragged column
// stores `flattened_data` and `cumulative_length` (end offsets).  One CTA turns the block's lengths into end offsets
// (chunked inclusive scan with a carried total, clamped to the block width like the reference's per-row slice),
// then one warp per row copies its entries to their place -- the padded block never leaves the device.
__global__ void __launch_bounds__(1024) k_vov_offsets(const uint32_t* lens, long long n_rows, int width, long long base,
                                                       long long* ends, uint32_t* cum_out) {
  __shared__ long long warp_tot[32];
  __shared__ long long carry;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = base;
  __syncthreads();
  for (long long r0 = 0; r0 < n_rows; r0 += 1024) {
    const long long r = r0 + threadIdx.x;
    long long v = 0;
    if (r < n_rows) {
      const uint32_t l = lens[r];
      v = l < (uint32_t)width ? l : (uint32_t)width;
    }
    long long inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      long long w = warp_tot[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const long long end = carry + (wid ? warp_tot[wid - 1] : 0) + inc;
    if (r < n_rows) {
      ends[r] = end;
      cum_out[r] = (uint32_t)end;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_tot[31];
    __syncthreads();
  }
}
template <typename E>
__global__ void k_vov_scatter(const E* block, long long row_stride, const uint32_t* lens, int width, const long long* ends,
                              long long n_rows, long long base, E* flat) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  uint32_t l = lens[row];
  if (l > (uint32_t)width) l = (uint32_t)width;
  const long long dst = ends[row] - (long long)l - base;
  for (uint32_t j = lane; j < l; j += 32) flat[dst + j] = block[row * row_stride + j];
}

// ---- generic recursive (IIR) filters of any order, one thread per row ----------------------------------------------
// recursive_filter.py:12-93 for len(b) > 3 (the order <= 2 case is the parallel affine scan of processors.cu): the
// reference's loop itself -- float64 circular buffer initialised with init_out, feed-forward padded with init_in.
// Rows are independent, so a warp works on 32 rows; the recursion is sequential along the waveform by nature.
struct IirCoef {
  double a[16], b[16];
  int p, q;
};
template <typename T>
__global__ void k_iir_general(const void* w, long long w_rs, int w_dt, long long n_rows, int n, IirCoef cf, SIn<T> init_in,
                              SIn<T> init_out, T* out, long long o_rs) {
  const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const long long base = row * w_rs;
  T* o = out + row * o_rs;
  const T ii = init_in.get(row), io = init_out.get(row);
  bool bad = ii != ii || io != io;
  for (int i = 0; i < n && !bad; i++) {
    const T v = ldw<T>(w, w_dt, base + i);
    bad = v != v;
  }
  if (bad) {
    for (int i = 0; i < n; i++) o[i] = (T)NAN;
    return;
  }
  double circ[16];
  for (int j = 0; j < cf.q; j++) circ[j] = (double)io;
  for (int i = 0; i < n; i++) {
    const int ib = i % cf.q;
    double acc = 0.0;
    for (int j = 0; j < cf.p; j++) acc = __dadd_rn(acc, __dmul_rn(cf.a[j], j <= i ? (double)ldw<T>(w, w_dt, base + i - j) : (double)ii));
    for (int j = 1; j < cf.q; j++) {
      int k = ib - j;
      if (k < 0) k += cf.q;
      acc = __dsub_rn(acc, __dmul_rn(cf.b[j], circ[k]));
    }
    acc /= cf.b[0];
    circ[ib] = acc;
    o[i] = (T)acc;
  }
}
// rc_cr2.py:11-93: matched z-transform RC-CR^2 shaper, float64 state seeded with the first three input samples, which
// are also the first three outputs
template <typename T>
__global__ void k_rc_cr2(const void* w, long long w_rs, int w_dt, long long n_rows, int n, SIn<T> tau_in, T* out, long long o_rs,
                         int* fatal) {
  const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const long long base = row * w_rs;
  T* o = out + row * o_rs;
  const T tau = tau_in.get(row);
  const double af = exp(-1.0 / (double)tau);         // (numba: int / float32 is float64, so both loops use float64 here)
  const double d2 = -3.0 * af, d3 = 3.0 * af * af, d4 = -(af * af * af);
  double t0 = (double)ldw<T>(w, w_dt, base), t1 = n > 1 ? (double)ldw<T>(w, w_dt, base + 1) : 0.0,
         t2 = n > 2 ? (double)ldw<T>(w, w_dt, base + 2) : 0.0;
  bool bad = tau != tau;
  for (int i = 0; i < n && !bad; i++) {
    const T v = ldw<T>(w, w_dt, base + i);
    bad = v != v;
  }
  if (bad) {                                   // NaN in, NaN out (rc_cr2.py:46-49)
    for (int i = 0; i < n; i++) o[i] = (T)NAN;
    return;
  }
  for (int i = 0; i < 3 && i < n; i++) o[i] = ldw<T>(w, w_dt, base + i);   // the first three samples pass through
  for (int i = 3; i < n; i++) {
    const double x0 = (double)ldw<T>(w, w_dt, base + i), x1 = (double)ldw<T>(w, w_dt, base + i - 1),
                 x2 = (double)ldw<T>(w, w_dt, base + i - 2);
    const double t3 = __dadd_rn(__dadd_rn(__dadd_rn(__dsub_rn(__dsub_rn(__dmul_rn(-d2, t2), __dmul_rn(d3, t1)), __dmul_rn(d4, t0)), x0),
                                          __dmul_rn(-2.0, x1)), x2);
    o[i] = (T)t3;
    bad |= (t3 != t3);
    t0 = t1; t1 = t2; t2 = t3;
  }
  if (bad) raise_fatal(fatal, DSPB_FATAL_RCCR2_NAN, row);
}

int grid_rows(long long n_rows) { return (int)(n_rows < 148LL * 16 ? n_rows : 148LL * 16); }
int last_error() {
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace

#define SIN(name) mk_sin<T>(name, name##_stride, name##_imm)

#define DSPB_DEFINE_SIPM(SFX, T_)                                                                                        \
  extern "C" int dspb_histogram##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, void* weights, int64_t weights_rs,     \
                                     int64_t n_bins, void* borders, int64_t borders_rs, int64_t n_borders, DSPB_TAIL) {    \
    using T = T_;                                                                                                        \
    if (n_bins + 1 != n_borders) return DSPB_FATAL_HIST_LEN;                                                              \
    if (n_rows <= 0) return 0;                                                                                           \
    const size_t smem = (n_bins + 4) * sizeof(int) + 2 * TPB * sizeof(T) + 16;                                           \
    k_histogram<T, false><<<grid_rows(n_rows), TPB, smem, (cudaStream_t)stream>>>(                                        \
        w_in, w_in_row_stride, w_in_dtype, n_rows, (int)n, SIn<T>{nullptr, 0, (T)0}, SIn<T>{nullptr, 0, (T)1}, (T*)weights, \
        weights_rs, (T*)borders, borders_rs, (int)n_bins, fatal);                                                        \
    return last_error();                                                                                                 \
  }                                                                                                                      \
  extern "C" int dspb_histogram_around_mode##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, DSPB_SCALAR(center),       \
                                                 DSPB_SCALAR(bin_width), void* weights, int64_t weights_rs, int64_t n_bins, \
                                                 void* borders, int64_t borders_rs, int64_t n_borders, DSPB_TAIL) {       \
    using T = T_;                                                                                                        \
    if (n_bins + 1 != n_borders) return DSPB_FATAL_HIST_LEN;                                                              \
    if (n_rows <= 0) return 0;                                                                                           \
    const size_t smem = (n_bins + 4) * sizeof(int) + 2 * TPB * sizeof(T) + 16;                                           \
    k_histogram<T, true><<<grid_rows(n_rows), TPB, smem, (cudaStream_t)stream>>>(                                         \
        w_in, w_in_row_stride, w_in_dtype, n_rows, (int)n, SIN(center), SIN(bin_width), (T*)weights, weights_rs,          \
        (T*)borders, borders_rs, (int)n_bins, fatal);                                                                    \
    return last_error();                                                                                                 \
  }                                                                                                                      \
  extern "C" int dspb_histogram_stats##SFX(const void* weights, int64_t weights_rs, int64_t n_bins, const void* edges,     \
                                           int64_t edges_rs, int64_t n_edges, int64_t n_rows, void* mode_out,            \
                                           void* max_out, void* fwhm_out, DSPB_SCALAR(max_in), DSPB_TAIL) {              \
    using T = T_;                                                                                                        \
    (void)fatal;                                                                                                         \
    if (n_bins + 1 != n_edges) return DSPB_FATAL_HPS_LEN;                                                                 \
    if (n_rows <= 0) return 0;                                                                                           \
    k_histogram_stats<T><<<(unsigned)((n_rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(                             \
        (const T*)weights, weights_rs, (const T*)edges, edges_rs, n_rows, (int)n_bins, SIN(max_in), (T*)mode_out,          \
        (T*)max_out, (T*)fwhm_out);                                                                                      \
    return last_error();                                                                                                 \
  }                                                                                                                      \
  extern "C" int dspb_histogram_peakstats##SFX(const void* weights, int64_t weights_rs, int64_t n_bins,                   \
                                               const void* edges, int64_t edges_rs, int64_t n_edges, int64_t n_rows,     \
                                               DSPB_SCALAR(max_in), int32_t skip_zeroes, int32_t width_type,             \
                                               void* mode_out, void* width_out, DSPB_TAIL) {                             \
    using T = T_;                                                                                                        \
    if (n_bins + 1 != n_edges) return DSPB_FATAL_HPS_LEN;                                                                 \
    if (width_type < 0 || width_type > 4) return DSPB_FATAL_HPS_WIDTH_TYPE;                                               \
    if (n_rows <= 0) return 0;                                                                                           \
    k_histogram_peakstats<T><<<(unsigned)((n_rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(                         \
        (const T*)weights, weights_rs, (const T*)edges, edges_rs, n_rows, (int)n_bins, SIN(max_in), skip_zeroes,          \
        width_type, (T*)mode_out, (T*)width_out, fatal);                                                                 \
    return last_error();                                                                                                 \
  }                                                                                                                      \
  extern "C" int dspb_peak_snr_threshold##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, const void* idx_in,           \
                                              int64_t idx_in_rs, int64_t m, DSPB_SCALAR(ratio_in), DSPB_SCALAR(width_in), \
                                              void* idx_out, int64_t idx_out_rs, void* n_idx_out, DSPB_TAIL) {            \
    using T = T_;                                                                                                        \
    (void)fatal;                                                                                                         \
    if (n_rows <= 0) return 0;                                                                                           \
    k_peak_snr<T><<<(unsigned)((n_rows + 63) / 64), 64, 0, (cudaStream_t)stream>>>(                                       \
        w_in, w_in_row_stride, w_in_dtype, n_rows, (int)n, (const T*)idx_in, idx_in_rs, (int)m, SIN(ratio_in),            \
        SIN(width_in), (T*)idx_out, idx_out_rs, (uint32_t*)n_idx_out);                                                   \
    return last_error();                                                                                                 \
  }                                                                                                                      \
  extern "C" int dspb_multi_a_filter##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, const void* vt_maxs_in,           \
                                          int64_t vt_rs, int64_t m, void* va_max_out, int64_t va_rs, DSPB_TAIL) {         \
    using T = T_;                                                                                                        \
    if (!(m < n)) return DSPB_FATAL_GMLE_LEN;                                                                             \
    if (n_rows <= 0) return 0;                                                                                           \
    k_multi_a<T><<<grid_rows(n_rows), TPB, 0, (cudaStream_t)stream>>>(w_in, w_in_row_stride, w_in_dtype, n_rows, (int)n,   \
                                                                     (const T*)vt_maxs_in, vt_rs, (int)m, (T*)va_max_out, \
                                                                     va_rs, fatal);                                      \
    return last_error();                                                                                                 \
  }                                                                                                                      \
  extern "C" int dspb_reflected_convolve_wf##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, const void* kernel,        \
                                                 int64_t m, DSPB_WAVE_OUT(w_out), DSPB_TAIL) {                            \
    using T = T_;                                                                                                        \
    (void)fatal;                                                                                                         \
    if (m > n) return DSPB_FATAL_CONV_KERNEL_LONG;                                                                        \
    if (n_rows <= 0) return 0;                                                                                           \
    const size_t smem = (size_t)(m + n + 2 * (m / 2 + 1)) * sizeof(T);                                                   \
    if (smem > 200 * 1024) return DSPB_ERR_ROW_TOO_LONG;                                                                  \
    auto kern = k_reflected_conv<T>;                                                                                     \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                  \
    if (e != cudaSuccess) return -(int)e;                                                                                \
    kern<<<grid_rows(n_rows), TPB, smem, (cudaStream_t)stream>>>(w_in, w_in_row_stride, w_in_dtype, n_rows, (int)n,        \
                                                                 (const T*)kernel, (int)m, (T*)w_out, w_out_row_stride);  \
    return last_error();                                                                                                 \
  }                                                                                                                      \
  extern "C" int dspb_recursive_filter_general##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, const double* a,        \
                                                    int64_t p, const double* b, int64_t q, DSPB_SCALAR(init_in),           \
                                                    DSPB_SCALAR(init_out), DSPB_WAVE_OUT(w_out), DSPB_TAIL) {             \
    using T = T_;                                                                                                        \
    (void)fatal;                                                                                                         \
    if (q == 0) return DSPB_FATAL_RF_B_SCALAR;                                                                            \
    if (n <= q) return DSPB_FATAL_RF_SHORT;                                                                               \
    if (p > 16 || q > 16) return DSPB_ERR_UNSUPPORTED;                                                                    \
    if (n_rows <= 0) return 0;                                                                                           \
    IirCoef cf;                                                                                                          \
    cf.p = (int)p; cf.q = (int)q;                                                                                        \
    for (int j = 0; j < 16; j++) { cf.a[j] = j < p ? a[j] : 0.0; cf.b[j] = j < q ? b[j] : 0.0; }                          \
    k_iir_general<T><<<(unsigned)((n_rows + 31) / 32), 32, 0, (cudaStream_t)stream>>>(                                    \
        w_in, w_in_row_stride, w_in_dtype, n_rows, (int)n, cf, SIN(init_in), SIN(init_out), (T*)w_out, w_out_row_stride);  \
    return last_error();                                                                                                 \
  }                                                                                                                      \
  extern "C" int dspb_rc_cr2##SFX(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, DSPB_SCALAR(t_tau), DSPB_WAVE_OUT(w_out), \
                                  DSPB_TAIL) {                                                                           \
    using T = T_;                                                                                                        \
    if (n <= 3) return DSPB_FATAL_DPZ_SHORT;                                                                              \
    if (n_rows <= 0) return 0;                                                                                           \
    k_rc_cr2<T><<<(unsigned)((n_rows + 31) / 32), 32, 0, (cudaStream_t)stream>>>(                                         \
        w_in, w_in_row_stride, w_in_dtype, n_rows, (int)n, SIN(t_tau), (T*)w_out, w_out_row_stride, fatal);               \
    return last_error();                                                                                                 \
  }                                                                                                                      \
  extern "C" int dspb_gaussian_filter1d##SFX(double sigma, double truncate, void* weights, int64_t length,                \
                                             void* stream) {                                                             \
    using T = T_;                                                                                                        \
    const T sg = (T)sigma, tr = (T)truncate;                                                                             \
    const int lw = (int)(T)((T)(tr * sg) + (T)0.5);   /* int(truncate * sd + 0.5) in the loop's type */                    \
    if (length != 2 * lw + 1) return DSPB_FATAL_SHAPE;                                                                    \
    const T coef = (T)-0.5 / (T)(sg * sg);            /* -0.5 / sigma2, rounded to the loop's type */                      \
    k_gaussian<T><<<1, 128, 0, (cudaStream_t)stream>>>((double)coef, lw, (T*)weights, (int)length);                         \
    return last_error();                                                                                                 \
  }                                                                                                                      \
  extern "C" int dspb_moving_slope##SFX(void* kernel, int64_t length, void* stream) {                                     \
    k_moving_slope<T_><<<1, 128, 0, (cudaStream_t)stream>>>((T_*)kernel, (int)length);                                    \
    return last_error();                                                                                                 \
  }                                                                                                                      \
  extern "C" int dspb_step##SFX(double weight_pos, void* kernel, int64_t length, void* stream) {                          \
    (void)weight_pos;                                                                                                    \
    k_step<T_><<<1, 128, 0, (cudaStream_t)stream>>>((T_*)kernel, (int)length);                                            \
    return last_error();                                                                                                 \
  }

DSPB_DEFINE_SIPM(_f32, float)
DSPB_DEFINE_SIPM(_f64, double)

// VectorOfVectors compaction: lens uint32[n_rows] -> ends int64[n_rows] (absolute end offsets, first row starts at
// `base`) and cumulative_length uint32[n_rows]; then block [n_rows, width] (elements of elem_bytes = 4 or 8) ->
// flat[0 .. ends[n_rows-1] - base).  All pointers are device pointers.
extern "C" int dspb_vov_offsets(const uint32_t* lens, int64_t n_rows, int64_t width, int64_t base, int64_t* ends,
                                uint32_t* cumulative_length, void* stream) {
  if (n_rows <= 0) return 0;
  k_vov_offsets<<<1, 1024, 0, (cudaStream_t)stream>>>(lens, n_rows, (int)width, base, (long long*)ends, cumulative_length);
  return last_error();
}
extern "C" int dspb_vov_compact(const void* block, int64_t row_stride, int32_t elem_bytes, const uint32_t* lens,
                                int64_t width, const int64_t* ends, int64_t n_rows, int64_t base, void* flat, void* stream) {
  if (n_rows <= 0) return 0;
  const unsigned grid = (unsigned)((n_rows * 32 + 255) / 256);
  if (elem_bytes == 4)
    k_vov_scatter<uint32_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)block, row_stride, lens, (int)width,
                                                                   (const long long*)ends, n_rows, base, (uint32_t*)flat);
  else if (elem_bytes == 8)
    k_vov_scatter<uint64_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint64_t*)block, row_stride, lens, (int)width,
                                                                   (const long long*)ends, n_rows, base, (uint64_t*)flat);
  else if (elem_bytes == 2)
    k_vov_scatter<uint16_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)block, row_stride, lens, (int)width,
                                                                   (const long long*)ends, n_rows, base, (uint16_t*)flat);
  else
    return DSPB_ERR_UNSUPPORTED;
  return last_error();
}
