// dspeed_b200 -- run-time library of the SPECIALISED chain kernels.
//
// dspeed_b200/codegen.py turns a compiled ProcessingChain (reference: the per-block
// processor loop of processing_chain.py:1144-1163) into one straight-line CUDA kernel and
// compiles it with nvcc for sm_100a.  The generated code is glue: everything that touches
// data is a routine of this header (or of row_ops.cuh for the rarely used processors).
//
// Execution model (one persistent 512-thread CTA per SM, one waveform at a time):
//  * thread t owns samples [16t, 16t+16) of every waveform ("chunk"); chunks stay in
//    registers across consecutive processors of the generated code;
//  * waveforms that other threads must see live in shared-memory slots in the T4 layout
//    (common.cuh): own-chunk and fixed-shift accesses are conflict-free 128-bit loads at
//    immediate offsets;
//  * recursive filters are chunk-local running sums + ONE block scan of the chunk totals;
//  * block collectives (sums, scans, arg-min/max) cost one barrier: partial results go
//    through a double-buffered scratch (`par` toggles every round), and independent
//    collectives of the same round share that barrier;
//  * per-event scalars are uniform registers.
#pragma once
#ifndef DSPB_PSP
#error "DSPB_PSP (padded plane stride of the T4 layout) must be defined before chain_rt.cuh"
#endif
#include "conv_ops.cuh"
#include "row_ops.cuh"

namespace crt {
using namespace dspb;

constexpr int CHK = 16;  // samples per thread chunk
constexpr int NWP = 16;  // warps per CTA
constexpr unsigned FULL = 0xffffffffu;

// double-buffered collective scratch: [parity][collective slot][warp]
struct CScr {
  double d[2][16][NWP];
  int i[2][8][NWP];
};

// ---------------------------------------------------------------------------------------
// chunk access
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldv(const float* slot, int plane, int chunk) {
  return *reinterpret_cast<const float4*>(slot + plane * DSPB_PSP + 4 * chunk);
}
__device__ __forceinline__ void stv(float* slot, int plane, int chunk, float4 v) {
  *reinterpret_cast<float4*>(slot + plane * DSPB_PSP + 4 * chunk) = v;
}
__device__ __forceinline__ float comp(const float4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }

// own chunk (no bounds: slots always hold whole chunks)
__device__ __forceinline__ void ld_chunk(const float* slot, int t, float (&o)[CHK]) {
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const float4 v = ldv(slot, p, t);
    o[4 * p] = v.x; o[4 * p + 1] = v.y; o[4 * p + 2] = v.z; o[4 * p + 3] = v.w;
  }
}
// own chunk of a wave of n samples: zero for threads beyond the last chunk
__device__ __forceinline__ void ld_chunk_n(const float* slot, int t, int n, float (&o)[CHK]) {
  if (CHK * t < n) ld_chunk(slot, t, o);
  else {
#pragma unroll
    for (int j = 0; j < CHK; j++) o[j] = 0.f;
  }
}
__device__ __forceinline__ void st_chunk(float* slot, int t, const float (&v)[CHK]) {
#pragma unroll
  for (int p = 0; p < 4; p++) stv(slot, p, t, make_float4(v[4 * p], v[4 * p + 1], v[4 * p + 2], v[4 * p + 3]));
}
// store with samples >= n forced to zero (keeps the "zero beyond the end" invariant that
// shifted loads rely on); threads whose chunk starts beyond ceil16(n) do nothing
__device__ __forceinline__ void st_chunk_n(float* slot, int t, int n, const float (&v)[CHK]) {
  if (CHK * t >= n + CHK - 1) return;
  float w[CHK];
#pragma unroll
  for (int j = 0; j < CHK; j++) w[j] = (CHK * t + j < n) ? v[j] : 0.0f;
  st_chunk(slot, t, w);
}

__host__ __device__ constexpr int floor_div16(int d) { return d >= 0 ? d / 16 : -((-d + 15) / 16); }

// o[j] = x[16 t + j + D], zero outside [0, nceil) where nceil = n rounded up to whole chunks
// (the tail of the last chunk holds zeros, see st_chunk_n).  D is a compile-time shift.
template <int D>
__device__ __forceinline__ void ld_shift(const float* slot, int t, int n, float (&o)[CHK]) {
  constexpr int qd = floor_div16(D), rd = D - 16 * qd, a = rd >> 2, b = rd & 3;
  constexpr int NV = b ? 5 : 4;
  float4 v[NV];
  const int nchunks = (n + CHK - 1) >> 4;
#pragma unroll
  for (int u = 0; u < NV; u++) {
    const int vv = a + u, plane = vv & 3, c = t + qd + (vv >> 2);
    v[u] = (c >= 0 && c < nchunks) ? ldv(slot, plane, c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int j = 0; j < CHK; j++) o[j] = comp(v[(j + b) >> 2], (j + b) & 3);
}

// raw waveform row from HBM: thread t reads its 16 samples (2 x 128-bit for uint16)
__device__ __forceinline__ uint4 ldg_nc(const void* p) {
  uint4 q;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
               : "l"(p));
  return q;
}
__device__ __forceinline__ void unpack_u16(const uint4& q, float* o) {
  o[0] = (float)(q.x & 0xffffu); o[1] = (float)(q.x >> 16);
  o[2] = (float)(q.y & 0xffffu); o[3] = (float)(q.y >> 16);
  o[4] = (float)(q.z & 0xffffu); o[5] = (float)(q.z >> 16);
  o[6] = (float)(q.w & 0xffffu); o[7] = (float)(q.w >> 16);
}
// n % 16 == 0 and 16-byte aligned rows (checked by the launcher)
__device__ __forceinline__ void ldg_chunk_u16(const uint16_t* g, int t, int n, float (&o)[CHK]) {
  if (CHK * t < n) {
    const uint4 q0 = ldg_nc(g + CHK * t), q1 = ldg_nc(g + CHK * t + 8);
    unpack_u16(q0, o);
    unpack_u16(q1, o + 8);
  } else {
#pragma unroll
    for (int j = 0; j < CHK; j++) o[j] = 0.f;
  }
}
template <typename TIn>
__device__ __forceinline__ int ldg_chunk_any(const TIn* g, int t, int n, float (&o)[CHK]) {
  int has_nan = 0;
#pragma unroll
  for (int j = 0; j < CHK; j++) {
    const int i = CHK * t + j;
    float v = i < n ? (float)g[i] : 0.f;
    has_nan |= (v != v);
    o[j] = v;
  }
  return has_nan;
}
__device__ __forceinline__ void stg_chunk(float* g, int t, int n, const float (&v)[CHK]) {
#pragma unroll
  for (int j = 0; j < CHK; j++)
    if (CHK * t + j < n) g[CHK * t + j] = v[j];
}

// ---------------------------------------------------------------------------------------
// warp-level pieces
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double wscan_incl(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ double wscan_incl_rev(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_down_sync(FULL, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}

// order-preserving key of a float (-0 folded onto +0), and back
__device__ __forceinline__ unsigned fkey(float v) {
  const unsigned u = __float_as_uint(v + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ---------------------------------------------------------------------------------------
// block collectives, split into "put" (before the barrier) and "get" (after it)
// ---------------------------------------------------------------------------------------
// sum of one double per thread
__device__ __forceinline__ void put_sum(CScr* cs, int par, int slot, double v, int lane, int warp) {
  v = wsum(v);
  if (lane == 0) cs->d[par][slot][warp] = v;
}
__device__ __forceinline__ double get_sum(const CScr* cs, int par, int slot, int lane) {
  return wsum(lane < NWP ? cs->d[par][slot][lane] : 0.0);
}
// exclusive forward scan of one double per thread; `incl` (this thread's inclusive warp
// scan) must be kept by the caller for get_excl
__device__ __forceinline__ double put_scan(CScr* cs, int par, int slot, double v, int lane, int warp) {
  const double incl = wscan_incl(v, lane);
  if (lane == 31) cs->d[par][slot][warp] = incl;
  return incl;
}
__device__ __forceinline__ double get_excl(const CScr* cs, int par, int slot, double incl, double v, int lane,
                                           int warp, double& total) {
  const double part = lane < NWP ? cs->d[par][slot][lane] : 0.0;
  const double pin = wscan_incl(part, lane);
  total = __shfl_sync(FULL, pin, NWP - 1);
  return __shfl_sync(FULL, pin - part, warp) + (incl - v);
}
// reverse (suffix) scan
__device__ __forceinline__ double put_scan_rev(CScr* cs, int par, int slot, double v, int lane, int warp) {
  const double incl = wscan_incl_rev(v, lane);
  if (lane == 0) cs->d[par][slot][warp] = incl;
  return incl;
}
__device__ __forceinline__ double get_excl_rev(const CScr* cs, int par, int slot, double incl, double v, int lane,
                                               int warp) {
  const double part = lane < NWP ? cs->d[par][slot][lane] : 0.0;
  const double pin = wscan_incl_rev(part, lane);
  return __shfl_sync(FULL, pin - part, warp) + (incl - v);
}

// first-occurrence arg-max / arg-min of (value, index) pairs: two REDUX per level
__device__ __forceinline__ void put_argmax(CScr* cs, int par, int slot, float v, int idx, int lane, int warp) {
  const unsigned k = fkey(v);
  const unsigned km = __reduce_max_sync(FULL, k);
  const int im = __reduce_min_sync(FULL, k == km ? idx : 0x7fffffff);
  if (lane == 0) { cs->i[par][slot][warp] = (int)km; cs->i[par][slot + 1][warp] = im; }
}
__device__ __forceinline__ void get_argmax(const CScr* cs, int par, int slot, int lane, float& v, int& idx) {
  const unsigned k = lane < NWP ? (unsigned)cs->i[par][slot][lane] : 0u;
  const int ii = lane < NWP ? cs->i[par][slot + 1][lane] : 0x7fffffff;
  const unsigned km = __reduce_max_sync(FULL, k);
  idx = __reduce_min_sync(FULL, k == km ? ii : 0x7fffffff);
  v = fkey_inv(km);
}
__device__ __forceinline__ void put_argmin(CScr* cs, int par, int slot, float v, int idx, int lane, int warp) {
  const unsigned k = fkey(v);
  const unsigned km = __reduce_min_sync(FULL, k);
  const int im = __reduce_min_sync(FULL, k == km ? idx : 0x7fffffff);
  if (lane == 0) { cs->i[par][slot][warp] = (int)km; cs->i[par][slot + 1][warp] = im; }
}
__device__ __forceinline__ void get_argmin(const CScr* cs, int par, int slot, int lane, float& v, int& idx) {
  const unsigned k = lane < NWP ? (unsigned)cs->i[par][slot][lane] : 0xffffffffu;
  const int ii = lane < NWP ? cs->i[par][slot + 1][lane] : 0x7fffffff;
  const unsigned km = __reduce_min_sync(FULL, k);
  idx = __reduce_min_sync(FULL, k == km ? ii : 0x7fffffff);
  v = fkey_inv(km);
}
// max / min of one int per thread
__device__ __forceinline__ void put_imax(CScr* cs, int par, int slot, int v, int lane, int warp) {
  v = __reduce_max_sync(FULL, v);
  if (lane == 0) cs->i[par][slot][warp] = v;
}
__device__ __forceinline__ int get_imax(const CScr* cs, int par, int slot, int lane) {
  return __reduce_max_sync(FULL, lane < NWP ? cs->i[par][slot][lane] : (int)0x80000000);
}
__device__ __forceinline__ void put_imin(CScr* cs, int par, int slot, int v, int lane, int warp) {
  v = __reduce_min_sync(FULL, v);
  if (lane == 0) cs->i[par][slot][warp] = v;
}
__device__ __forceinline__ int get_imin(const CScr* cs, int par, int slot, int lane) {
  return __reduce_min_sync(FULL, lane < NWP ? cs->i[par][slot][lane] : 0x7fffffff);
}

// ---------------------------------------------------------------------------------------
// chunk-local arithmetic of the processors
// ---------------------------------------------------------------------------------------
// min_max.py:11-82 on the samples [lo, hi) of a wave (chunk starts at i0): strict compares,
// ascending order => first occurrence.  Index is relative to `lo`.
struct MinMax {
  float vmin, vmax;
  int imin, imax;
};
__device__ __forceinline__ MinMax minmax_local(const float (&v)[CHK], int i0, int lo, int hi) {
  MinMax m;
  m.vmin = CUDART_INF_F; m.vmax = -CUDART_INF_F; m.imin = 0x7fffffff; m.imax = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < CHK; j++) {
    const int i = i0 + j;
    if (i >= lo && i < hi) {
      if (v[j] < m.vmin) { m.vmin = v[j]; m.imin = i - lo; }
      if (v[j] > m.vmax) { m.vmax = v[j]; m.imax = i - lo; }
    }
  }
  return m;
}

// linear_slope_fit.py:11-90 : sums over [lo, hi), abscissa relative to lo
__device__ __forceinline__ void lsf_local(const float (&v)[CHK], int i0, int lo, int hi, double& sy, double& sxy,
                                          double& syy) {
  sy = 0.0; sxy = 0.0; syy = 0.0;
  if (i0 + CHK <= lo || i0 >= hi) return;
#pragma unroll
  for (int j = 0; j < CHK; j++) {
    const int i = i0 + j;
    if (i >= lo && i < hi) {
      const double y = (double)v[j];
      sy += y;
      sxy = fma(y, (double)(i - lo), sxy);
      syy = fma(y, y, syy);
    }
  }
}
// from the block sums: mean, sample standard deviation, least-squares slope and intercept
__device__ __forceinline__ void lsf_finish(int n, double sy, double sxy, double syy, float& mean, float& stdev,
                                           float& slope, float& icpt) {
  const long long nn = n, sx = nn * (nn - 1) / 2, sx2 = (nn - 1) * nn * (2 * nn - 1) / 6;
  const double m = sy / (double)n;
  double m2 = syy - sy * m;
  if (m2 < 0.0) m2 = 0.0;
  mean = (float)m;
  stdev = (float)sqrt(m2 / (double)(n - 1));
  slope = (float)(((double)nn * sxy - (double)sx * sy) / (double)(nn * sx2 - sx * sx));
  icpt = (float)((sy - (double)sx * (double)slope) / (double)nn);
}

__device__ __forceinline__ double chunk_sum_d(const float (&v)[CHK]) {
  // pairwise in float is exact for integer-valued samples; general samples: accumulate in double
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < CHK; j++) s += (double)v[j];
  return s;
}

// pole_zero.py:24-77 : y[i] = x[i] + (1-c) * S[i-1], S = inclusive prefix sum of x
// (`run` enters as S[i0-1] and leaves as S[i0+15])
__device__ __forceinline__ int pz_chunk(const float (&x)[CHK], double run, double omc, float (&y)[CHK]) {
  int bad = 0;
#pragma unroll
  for (int j = 0; j < CHK; j++) {
    const double xv = (double)x[j];
    const float o = (float)fma(omc, run, xv);
    bad |= (o != o);
    y[j] = o;
    run += xv;
  }
  return bad;
}

// inclusive running sum inside the chunk (float), returns the chunk total
__device__ __forceinline__ float cumsum_local(float (&d)[CHK]) {
#pragma unroll
  for (int j = 1; j < CHK; j++) d[j] += d[j - 1];
  return d[CHK - 1];
}
__device__ __forceinline__ float cumsum_local_rev(float (&d)[CHK]) {
#pragma unroll
  for (int j = CHK - 2; j >= 0; j--) d[j] += d[j + 1];
  return d[0];
}

// d[j] += c * x[16 t + j - TS]   (one tap of a sparse FIR; zero outside the wave)
template <int TS>
__device__ __forceinline__ void fir_tap(const float* slot, int t, int n, float c, float (&d)[CHK]) {
  float v[CHK];
  ld_shift<-TS>(slot, t, n, v);
#pragma unroll
  for (int j = 0; j < CHK; j++) d[j] = fmaf(c, v[j], d[j]);
}
template <int TS>
__device__ __forceinline__ void fir_tap_add(const float* slot, int t, int n, float (&d)[CHK]) {
  float v[CHK];
  ld_shift<-TS>(slot, t, n, v);
#pragma unroll
  for (int j = 0; j < CHK; j++) d[j] += v[j];
}
template <int TS>
__device__ __forceinline__ void fir_tap_sub(const float* slot, int t, int n, float (&d)[CHK]) {
  float v[CHK];
  ld_shift<-TS>(slot, t, n, v);
#pragma unroll
  for (int j = 0; j < CHK; j++) d[j] -= v[j];
}

// value of sample i of a slot (any thread)
__device__ __forceinline__ float at(const float* slot, int i) { return slot[sidx(i)]; }

// time_point_thresh.py:12-92 : block-wide search with ONE barrier per 512-sample window.
// Thread (lane, warp) of window k looks at sample s -/+ (16 * lane + warp + 512 k): lanes of a
// warp touch consecutive chunks at the same in-chunk offset (conflict-free in the T4 layout).
__device__ __forceinline__ int search_cross(const float* w, int n, float thr, int s, bool forward, int stop_back,
                                            CScr* cs, int& par, int lane, int warp) {
  const int u = 16 * lane + warp;
  if (forward) {
    for (int base = s; base < n - 1; base += 512) {
      const int i = base + u;
      int hit = 0x7fffffff;
      if (i < n - 1) {
        const float a = at(w, i), b = at(w, i + 1);
        if ((a <= thr && thr < b) || (a >= thr && thr > b)) hit = i;
      }
      put_imin(cs, par, 0, hit, lane, warp);
      __syncthreads();
      hit = get_imin(cs, par, 0, lane);
      par ^= 1;
      if (hit != 0x7fffffff) return hit;
    }
    return -1;
  }
  for (int base = s; base >= stop_back; base -= 512) {
    const int i = base - u;
    int hit = -1;
    if (i >= stop_back) {
      const float a = at(w, i - 1), b = at(w, i);
      if ((a < thr && thr <= b) || (a > thr && thr >= b)) hit = i;
    }
    put_imax(cs, par, 0, hit, lane, warp);
    __syncthreads();
    hit = get_imax(cs, par, 0, lane);
    par ^= 1;
    if (hit >= 0) return hit;
  }
  return -1;
}

__device__ __forceinline__ float tpt(const float* w, int n, float thr, float t_start, float walk, int& fatal,
                                     CScr* cs, int& par, int lane, int warp) {
  fatal = 0;
  if (thr != thr || t_start != t_start || walk != walk) return CUDART_NAN_F;
  if (floorf(t_start) != t_start) { fatal = DSPB_FATAL_TSTART_NONINT; return CUDART_NAN_F; }
  if (floorf(walk) != walk) { fatal = DSPB_FATAL_WALK_NONINT; return CUDART_NAN_F; }
  const long long s = (long long)t_start;
  if (s < 0 || s >= n) { fatal = DSPB_FATAL_TSTART_RANGE; return CUDART_NAN_F; }
  const int hit = search_cross(w, n, thr, (int)s, (long long)walk == 1, 1, cs, par, lane, warp);
  return hit < 0 ? CUDART_NAN_F : (float)hit;
}

}  // namespace crt
