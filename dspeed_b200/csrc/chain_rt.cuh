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

// tracing build only: cycle stamps inside a routine (thread 0 of CTA 0), accumulated next to the per-node
// stamps of the generated kernel (prof_ts[96 + k], offset DSPB_PROF_OFF of the dynamic shared memory)
#if defined(DSPB_PROFILE) && defined(DSPB_PROF_OFF)
#ifndef DSPB_PROF_TID
#define DSPB_PROF_TID 0   // the observing thread of CTA 0 (-DDSPB_PROF_TID=n: another warp's view)
#endif
#define PROF_SUB_BEGIN() long long psub_prev_ = clock64()
#define PROF_SUB_RESET() psub_prev_ = clock64()
#define PROF_SUB(k)                                                                        \
  do {                                                                                     \
    if (blockIdx.x == 0 && threadIdx.x == DSPB_PROF_TID) {                                 \
      extern __shared__ __align__(16) unsigned char psub_smem_[];                          \
      long long* p_ = reinterpret_cast<long long*>(psub_smem_ + DSPB_PROF_OFF);            \
      const long long t_ = clock64();                                                      \
      p_[96 + (k)] += t_ - psub_prev_;                                                     \
      psub_prev_ = t_;                                                                     \
    }                                                                                      \
  } while (0)
#else
#define PROF_SUB_BEGIN()
#define PROF_SUB_RESET()
#define PROF_SUB(k)
#endif

// Barrier of the 16 block warps.  Kernels that run a separate scalar warp define it as a named
// barrier over the 512 block threads before including this header.
#ifndef BSYNC
#define BSYNC() __syncthreads()
#endif

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
  if ((n & (CHK - 1)) == 0) {  // whole chunks only (n is a literal at every call site)
    st_chunk(slot, t, v);
    return;
  }
  float w[CHK];
#pragma unroll
  for (int j = 0; j < CHK; j++) w[j] = (CHK * t + j < n) ? v[j] : 0.0f;
  st_chunk(slot, t, w);
}

__host__ __device__ constexpr int floor_div16(int d) { return d >= 0 ? d / 16 : -((-d + 15) / 16); }

// o[m] = x[16 t + D + m] for m < CNT, zero outside [0, nceil) where nceil = n rounded up to whole
// chunks (the tail of the last chunk holds zeros, see st_chunk_n).  D is a compile-time shift:
// the vectors are read with 128-bit loads at immediate offsets.  Out-of-range chunks are
// redirected to the always-zero pad column `zc` of the slot instead of being predicated.
template <int D, int CNT>
__device__ __forceinline__ void ld_span(const float* slot, int t, int n, int zc, float (&o)[CNT]) {
  constexpr int qd = floor_div16(D), rd = D - 16 * qd, a = rd >> 2, b = rd & 3;
  constexpr int NV = (b + CNT + 3) >> 2;
  float4 v[NV];
  const int nchunks = (n + CHK - 1) >> 4;
#pragma unroll
  for (int u = 0; u < NV; u++) {
    const int vv = a + u, plane = vv & 3, c = t + qd + (vv >> 2);
    v[u] = ldv(slot, plane, ((unsigned)c < (unsigned)nchunks) ? c : zc);
  }
#pragma unroll
  for (int m = 0; m < CNT; m++) o[m] = comp(v[(m + b) >> 2], (m + b) & 3);
}
template <int D>
__device__ __forceinline__ void ld_shift(const float* slot, int t, int n, int zc, float (&o)[CHK]) {
  ld_span<D, CHK>(slot, t, n, zc, o);
}

// the two pad columns at the end of every plane of slots [k0, k1) are kept zero
__device__ __forceinline__ void zero_pads(float* slots, int slot_words, int nch, int k0, int k1, int tid) {
  for (int q = tid; q < (k1 - k0) * 32; q += 512) {
    const int k = k0 + (q >> 5), plane = (q >> 3) & 3, e = q & 7;
    slots[k * slot_words + plane * DSPB_PSP + 4 * nch + e] = 0.f;
  }
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
__device__ __forceinline__ void put_sum(double* p, double v, int lane, int warp) {
  v = wsum(v);
  if (lane == 0) p[warp] = v;
}
__device__ __forceinline__ double get_sum(const double* p, int lane) {
  return wsum(lane < NWP ? p[lane] : 0.0);
}
// exclusive forward scan of one double per thread; `incl` (this thread's inclusive warp
// scan) must be kept by the caller for get_excl
__device__ __forceinline__ double put_scan(double* p, double v, int lane, int warp) {
  const double incl = wscan_incl(v, lane);
  if (lane == 31) p[warp] = incl;
  return incl;
}
__device__ __forceinline__ double get_excl(const double* p, double incl, double v, int lane, int warp,
                                           double& total) {
  const double part = lane < NWP ? p[lane] : 0.0;
  const double pin = wscan_incl(part, lane);
  total = __shfl_sync(FULL, pin, NWP - 1);
  return __shfl_sync(FULL, pin - part, warp) + (incl - v);
}
// Geometric (first-order IIR) scans, s_l = v_l + R s_{l-1}: Rp[k] = R^(2^k) over lanes, Rw[k] = (R^32)^(2^k) over
// warps, Rl = R^lane.  get_excl_geo returns the state entering this thread's chunk.
__device__ __forceinline__ double wscan_geo(double v, const double (&Rp)[5], int lane) {
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const double t = __shfl_up_sync(FULL, v, 1 << k);
    if (lane >= (1 << k)) v = fma(Rp[k], t, v);
  }
  return v;
}
__device__ __forceinline__ double put_scan_geo(double* p, double v, const double (&Rp)[5], int lane, int warp) {
  const double incl = wscan_geo(v, Rp, lane);
  if (lane == 31) p[warp] = incl;
  return incl;
}
__device__ __forceinline__ double get_excl_geo(const double* p, double incl, const double (&Rw)[5], double Rl, int lane,
                                               int warp) {
  const double part = lane < NWP ? p[lane] : 0.0;
  const double pin = wscan_geo(part, Rw, lane);   // inclusive over the warps
  const double win = __shfl_sync(FULL, pin, warp > 0 ? warp - 1 : 0);
  const double prev = __shfl_up_sync(FULL, incl, 1);
  return (lane > 0 ? prev : 0.0) + (warp > 0 ? Rl * win : 0.0);
}

// double_pole_zero (pole_zero.py:82-198).  The denominator 1 + d1 z^-1 + d2 z^-2 has d1 + d2 = -1, i.e. it factors
// into (1 - z^-1)(1 - r z^-1) with r = d2: the filter is a first-order recursion v[i] = u[i] + r v[i-1] on
// u = x[i] + n1 x[i-1] + n2 x[i-2] followed by a running sum w = cumsum(v) (v[0] = x[0], v[1] = x[1] - x[0] reproduce
// the reference's w[0] = x[0], w[1] = x[1]).  Chunk-local pass: cl[j] = running sum of the zero-state response,
// vtot = its last value; the carry-in v_in of the chunk adds v_in * (r + .. + r^(j+1)) to cl[j].
__device__ __forceinline__ void dpz_local(const float (&x)[CHK], float xm1, float xm2, int i0, int n, double r, double n1,
                                          double n2, double (&cl)[CHK], double& vtot) {
  double v = 0.0, c = 0.0, p1 = (double)xm1, p2 = (double)xm2;
#pragma unroll
  for (int j = 0; j < CHK; j++) {
    const int i = i0 + j;
    const double xj = (double)x[j];
    double u = xj + n1 * p1 + n2 * p2;
    if (i == 0) u = xj;
    if (i == 1) u = xj - p1 - r * p1;
    if (i >= n) u = 0.0;
    v = fma(r, v, u);
    c += v;
    cl[j] = c;
    p2 = p1;
    p1 = xj;
  }
  vtot = v;
}

// float variants for running sums whose magnitude (< 2^24 times the output tolerance) allows it:
// half the shuffles of the float64 scan
__device__ __forceinline__ float wscan_incl_f(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ float wscan_incl_rev_f(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_down_sync(FULL, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}
__device__ __forceinline__ float put_scan_f(double* p, float v, int lane, int warp) {
  const float incl = wscan_incl_f(v, lane);
  if (lane == 31) p[warp] = (double)incl;
  return incl;
}
// exclusive prefix of this thread: float64 across the 16 warp totals, float inside the warp
__device__ __forceinline__ double get_excl_f(const double* p, float incl, float v, int lane, int warp, int nw = NWP) {
  const double part = lane < nw ? p[lane] : 0.0;   // nw: warps that took part in the scan
  double pin = part;
#pragma unroll
  for (int o = 1; o < NWP; o <<= 1) {
    const double t = __shfl_up_sync(FULL, pin, o);
    if (lane >= o) pin += t;
  }
  return __shfl_sync(FULL, pin - part, warp) + (double)(incl - v);
}
// all-float32 variants for running sums that ARE the float32 result (filter outputs: every partial sum is a value of
// the output waveform, so rounding it to float32 costs half an ulp of that output, like the reference's own
// sequential float32 accumulation): one shuffle per level instead of two, no conversions.  The scratch cell is the
// first half of the collective's double[NWP] row.
__device__ __forceinline__ float put_scan_ff(double* p, float v, int lane, int warp) {
  const float incl = wscan_incl_f(v, lane);
  if (lane == 31) reinterpret_cast<float*>(p)[warp] = incl;
  return incl;
}
__device__ __forceinline__ float get_excl_ff(const double* p, float incl, float v, int lane, int warp, int nw = NWP) {
  const float part = lane < nw ? reinterpret_cast<const float*>(p)[lane] : 0.f;
  float pin = part;
#pragma unroll
  for (int o = 1; o < NWP; o <<= 1) {
    const float t = __shfl_up_sync(FULL, pin, o);
    if (lane >= o) pin += t;
  }
  return __shfl_sync(FULL, pin - part, warp) + (incl - v);
}
__device__ __forceinline__ float put_scan_rev_ff(double* p, float v, int lane, int warp) {
  const float incl = wscan_incl_rev_f(v, lane);
  if (lane == 0) reinterpret_cast<float*>(p)[warp] = incl;
  return incl;
}
__device__ __forceinline__ float get_excl_rev_ff(const double* p, float incl, float v, int lane, int warp, int nw = NWP) {
  const float part = lane < nw ? reinterpret_cast<const float*>(p)[lane] : 0.f;
  float pin = part;
#pragma unroll
  for (int o = 1; o < NWP; o <<= 1) {
    const float t = __shfl_down_sync(FULL, pin, o);
    if (lane + o < NWP) pin += t;
  }
  return __shfl_sync(FULL, pin - part, warp) + (incl - v);
}
__device__ __forceinline__ float put_scan_rev_f(double* p, float v, int lane, int warp) {
  const float incl = wscan_incl_rev_f(v, lane);
  if (lane == 0) p[warp] = (double)incl;
  return incl;
}
__device__ __forceinline__ double get_excl_rev_f(const double* p, float incl, float v, int lane, int warp, int nw = NWP) {
  const double part = lane < nw ? p[lane] : 0.0;
  double pin = part;
#pragma unroll
  for (int o = 1; o < NWP; o <<= 1) {
    const double t = __shfl_down_sync(FULL, pin, o);
    if (lane + o < NWP) pin += t;
  }
  return __shfl_sync(FULL, pin - part, warp) + (double)(incl - v);
}
// reverse (suffix) scan
__device__ __forceinline__ double put_scan_rev(double* p, double v, int lane, int warp) {
  const double incl = wscan_incl_rev(v, lane);
  if (lane == 0) p[warp] = incl;
  return incl;
}
__device__ __forceinline__ double get_excl_rev(const double* p, double incl, double v, int lane, int warp) {
  const double part = lane < NWP ? p[lane] : 0.0;
  const double pin = wscan_incl_rev(part, lane);
  return __shfl_sync(FULL, pin - part, warp) + (incl - v);
}

// first-occurrence arg-max / arg-min of (value, index) pairs: two REDUX per level
__device__ __forceinline__ void put_argmax(int* p, float v, int idx, int lane, int warp) {
  const unsigned k = fkey(v);
  const unsigned km = __reduce_max_sync(FULL, k);
  const int im = __reduce_min_sync(FULL, k == km ? idx : 0x7fffffff);
  if (lane == 0) { p[warp] = (int)km; p[NWP + warp] = im; }
}
__device__ __forceinline__ void get_argmax(const int* p, int lane, float& v, int& idx, int w0 = 0, int w1 = NWP) {
  const bool in = lane >= w0 && lane < w1;   // warps that deposited a partial result
  const unsigned k = in ? (unsigned)p[lane] : 0u;
  const int ii = in ? p[NWP + lane] : 0x7fffffff;
  const unsigned km = __reduce_max_sync(FULL, k);
  idx = __reduce_min_sync(FULL, k == km ? ii : 0x7fffffff);
  v = fkey_inv(km);
}
__device__ __forceinline__ void put_argmin(int* p, float v, int idx, int lane, int warp) {
  const unsigned k = fkey(v);
  const unsigned km = __reduce_min_sync(FULL, k);
  const int im = __reduce_min_sync(FULL, k == km ? idx : 0x7fffffff);
  if (lane == 0) { p[warp] = (int)km; p[NWP + warp] = im; }
}
__device__ __forceinline__ void get_argmin(const int* p, int lane, float& v, int& idx, int w0 = 0, int w1 = NWP) {
  const bool in = lane >= w0 && lane < w1;
  const unsigned k = in ? (unsigned)p[lane] : 0xffffffffu;
  const int ii = in ? p[NWP + lane] : 0x7fffffff;
  const unsigned km = __reduce_min_sync(FULL, k);
  idx = __reduce_min_sync(FULL, k == km ? ii : 0x7fffffff);
  v = fkey_inv(km);
}
// max / min of one int per thread
__device__ __forceinline__ void put_imax(int* p, int v, int lane, int warp) {
  v = __reduce_max_sync(FULL, v);
  if (lane == 0) p[warp] = v;
}
__device__ __forceinline__ int get_imax(const int* p, int lane, int nw = NWP) {
  return __reduce_max_sync(FULL, lane < nw ? p[lane] : (int)0x80000000);
}
__device__ __forceinline__ void put_imin(int* p, int v, int lane, int warp) {
  v = __reduce_min_sync(FULL, v);
  if (lane == 0) p[warp] = v;
}
__device__ __forceinline__ int get_imin(const int* p, int lane) {
  return __reduce_min_sync(FULL, lane < NWP ? p[lane] : 0x7fffffff);
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

// value-only extrema (numpy.amax / unused index outputs): one FMNMX per sample
template <bool FULL_RANGE>
__device__ __forceinline__ float max_local(const float (&v)[CHK], int i0, int lo, int hi) {
  float m = -CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < CHK; j++)
    if (FULL_RANGE || (i0 + j >= lo && i0 + j < hi)) m = fmaxf(m, v[j]);
  return m;
}
template <bool FULL_RANGE>
__device__ __forceinline__ float min_local(const float (&v)[CHK], int i0, int lo, int hi) {
  float m = CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < CHK; j++)
    if (FULL_RANGE || (i0 + j >= lo && i0 + j < hi)) m = fminf(m, v[j]);
  return m;
}
// first index (relative to lo) inside the chunk where v == m, or INT_MAX
template <bool FULL_RANGE>
__device__ __forceinline__ int first_eq_local(const float (&v)[CHK], float m, int i0, int lo, int hi) {
  int idx = 0x7fffffff;
#pragma unroll
  for (int j = CHK - 1; j >= 0; j--)
    if ((FULL_RANGE || (i0 + j >= lo && i0 + j < hi)) && v[j] == m) idx = i0 + j - lo;
  return idx;
}
// Extremum AND first position of a whole chunk in one FMNMX pass, for integer-valued samples with |x| < 2^19 (the
// raw ADC words): 16 x + (15 - j) is exact in float32, its maximum is the largest sample at the smallest j
// (16 x + j: the smallest sample at the smallest j).  `j` = position inside the chunk.
__device__ __forceinline__ float argmax_packed(const float (&v)[CHK], int& j) {
  float k = -CUDART_INF_F;
#pragma unroll
  for (int q = 0; q < CHK; q++) k = fmaxf(k, fmaf(v[q], 16.f, (float)(CHK - 1 - q)));
  const float x = floorf(k * 0.0625f);
  j = CHK - 1 - (int)fmaf(x, -16.f, k);
  return x;
}
__device__ __forceinline__ float argmin_packed(const float (&v)[CHK], int& j) {
  float k = CUDART_INF_F;
#pragma unroll
  for (int q = 0; q < CHK; q++) k = fminf(k, fmaf(v[q], 16.f, (float)q));
  const float x = floorf(k * 0.0625f);
  j = (int)fmaf(x, -16.f, k);
  return x;
}
// value-only block max / min through the order-preserving key
__device__ __forceinline__ void put_fmax(int* p, float v, int lane, int warp) {
  const unsigned km = __reduce_max_sync(FULL, fkey(v));
  if (lane == 0) p[warp] = (int)km;
}
__device__ __forceinline__ float get_fmax(const int* p, int lane, int w0 = 0, int w1 = NWP) {
  return fkey_inv(__reduce_max_sync(FULL, (lane >= w0 && lane < w1) ? (unsigned)p[lane] : 0u));
}
__device__ __forceinline__ void put_fmin(int* p, float v, int lane, int warp) {
  const unsigned km = __reduce_min_sync(FULL, fkey(v));
  if (lane == 0) p[warp] = (int)km;
}
__device__ __forceinline__ float get_fmin(const int* p, int lane, int w0 = 0, int w1 = NWP) {
  return fkey_inv(__reduce_min_sync(FULL, (lane >= w0 && lane < w1) ? (unsigned)p[lane] : 0xffffffffu));
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
// a / b for a divisor known when the kernel is generated (rb = 1 / b, correctly rounded by the host compiler): the
// product with the reciprocal plus ONE correction with the exact remainder (fused multiply-add) is the correctly
// rounded quotient (Markstein's division; b's significand is not all ones for the sample counts used here) -- three
// float64 instructions in every block warp instead of the ~20 of the division sequence.
__device__ __forceinline__ double div_by(double a, double b, double rb) {
  const double q = a * rb;
  const double r = fma(-b, q, a);
  return fabs(q) <= 1.7976931348623157e308 ? fma(r, rb, q) : q;   // (inf / NaN sums pass through)
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

// ---------------------------------------------------------------------------------------
// float32-local variants (round 2).  Inside a 16-sample chunk the sums run in float32 -- exact for the
// integer-valued waveforms of the DAQ (|x| < 2^16: every partial sum stays below 2^24), rounding-level
// (<= 16 ulp of a 16-term sum) for general data -- and float64 only where chunks are combined.  One
// FADD / FFMA per sample instead of F2F + DADD / DFMA (4 + 2 issue cycles each on this part).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float chunk_sum_f(const float (&v)[CHK]) {
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; j++) a[j] = v[2 * j] + v[2 * j + 1];
#pragma unroll
  for (int j = 0; j < 4; j++) a[j] = a[2 * j] + a[2 * j + 1];
  return (a[0] + a[1]) + (a[2] + a[3]);
}
// pole_zero.py:24-77 : y[i] = x[i] + (1-c) * S[i-1] with S[i-1] = E + r, E = sum of all samples in front of the
// chunk (float64, from the block scan) and r = running sum inside the chunk (float32).  (1-c) E is rounded to
// float32 once per chunk: <= 1.5 ulp of the output in total (the reference stores float32: 0.5 ulp).
__device__ __forceinline__ void pz_chunk_f(const float (&x)[CHK], double excl, double omc, float (&y)[CHK]) {
  const float oe = (float)(omc * excl), omcf = (float)omc;
  float run = 0.f;
#pragma unroll
  for (int j = 0; j < CHK; j++) {
    y[j] = x[j] + fmaf(omcf, run, oe);
    run += x[j];
  }
}
// linear_slope_fit.py:11-90 : sums over [lo, hi) with the abscissa relative to lo.  The samples are centred on one
// sample of the chunk (c0): u = y - c0 is exact in float32 whenever it matters (a pedestal much larger than
// the fluctuations is exactly the case in which u needs few bits), so sum u, sum j u and sum u^2 over 16 samples
// carry no cancellation; the pedestal terms are added in float64.  FULL: the whole chunk lies inside [lo, hi).
template <bool FULL>
__device__ __forceinline__ void lsf_local_f(const float (&v)[CHK], int i0, int lo, int hi, double& sy, double& sxy,
                                            double& syy) {
  float c0 = v[0];
  if (!FULL) {   // first sample of the chunk that lies inside the range
    const int jf = lo - i0;
#pragma unroll
    for (int j = 1; j < CHK; j++) c0 = (j == jf) ? v[j] : c0;
  }
  float su = 0.f, sju = 0.f, suu = 0.f;
  int m = 0, sj = 0;
#pragma unroll
  for (int j = 0; j < CHK; j++) {
    if (FULL || (i0 + j >= lo && i0 + j < hi)) {
      const float u = v[j] - c0;
      su += u;
      sju = fmaf((float)j, u, sju);
      suu = fmaf(u, u, suu);
      m++;
      sj += j;
    }
  }
  const double c = (double)c0, dm = (double)m, dsu = (double)su;
  sy = fma(dm, c, dsu);
  sxy = fma((double)(i0 - lo), sy, fma((double)sj, c, (double)sju));
  syy = fma(c, fma(dm, c, dsu + dsu), (double)suu);
}
// block sum over the warps [w0, w1) only (the other warps hold no part of the range and deposit nothing)
__device__ __forceinline__ double get_sum_r(const double* p, int lane, int w0, int w1) {
  return wsum((lane >= w0 && lane < w1) ? p[lane] : 0.0);
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
__device__ __forceinline__ void fir_tap(const float* slot, int t, int n, int zc, float c, float (&d)[CHK]) {
  float v[CHK];
  ld_shift<-TS>(slot, t, n, zc, v);
#pragma unroll
  for (int j = 0; j < CHK; j++) d[j] = fmaf(c, v[j], d[j]);
}
template <int TS>
__device__ __forceinline__ void fir_tap_add(const float* slot, int t, int n, int zc, float (&d)[CHK]) {
  float v[CHK];
  ld_shift<-TS>(slot, t, n, zc, v);
#pragma unroll
  for (int j = 0; j < CHK; j++) d[j] += v[j];
}
template <int TS>
__device__ __forceinline__ void fir_tap_sub(const float* slot, int t, int n, int zc, float (&d)[CHK]) {
  float v[CHK];
  ld_shift<-TS>(slot, t, n, zc, v);
#pragma unroll
  for (int j = 0; j < CHK; j++) d[j] -= v[j];
}
// LEN consecutive taps TS0 .. TS0+LEN-1 with one coefficient: d[j] += c * sum_s x[i - TS0 - s],
// evaluated as a sliding window over one span of 16 + LEN - 1 samples
template <int TS0, int LEN>
__device__ __forceinline__ void fir_run(const float* slot, int t, int n, int zc, float c, float (&d)[CHK]) {
  float v[CHK + LEN - 1];
  ld_span<-(TS0 + LEN - 1), CHK + LEN - 1>(slot, t, n, zc, v);  // v[m] = x[16t + m - TS0 - LEN + 1]
  float w = 0.f;
#pragma unroll
  for (int m = 0; m < LEN; m++) w += v[m];
#pragma unroll
  for (int j = 0; j < CHK; j++) {
    d[j] = fmaf(c, w, d[j]);
    if (j + 1 < CHK) w += v[j + LEN] - v[j];
  }
}

// value of sample i of a slot (any thread)
__device__ __forceinline__ float at(const float* slot, int i) { return slot[sidx(i)]; }

// ---------------------------------------------------------------------------------------
// threshold search with a two-level summary (time_point_thresh.py:12-92)
//
// A waveform that is searched by the scalar warp carries a summary built by the block warps
// when they store it: min / max of every 16-sample chunk (level 1, 512 entries) and of every
// 512-sample stretch owned by one block warp (level 2, 16 entries).  A crossing of `thr`
// between samples i-1 and i needs  min <= thr <= max  over the chunks that hold the pair, so
// the search looks at level 2 (one ballot), then at the 32 chunks of the nearest candidate
// stretch (one ballot), then checks the 16 samples of the nearest candidate chunk exactly --
// three dependent steps whether the crossing is 5 or 5000 samples away.  Candidates are a
// superset (equality cases, pairs across chunk borders), the final check is the reference's
// exact condition, so the result is identical to the linear walk.
// ---------------------------------------------------------------------------------------
struct WaveSummary {
  float mn1[512], mx1[512];  // per chunk
  float mn2[16], mx2[16];    // per block warp (32 chunks)
};

// block side: called by every block thread with its own chunk in registers
__device__ __forceinline__ void put_summary(WaveSummary* sm, const float (&v)[CHK], int n, int tid, int lane,
                                            int warp) {
  float mn = CUDART_INF_F, mx = -CUDART_INF_F;
  if (n == CHK * 512) {   // (n is a literal at every call site)
#pragma unroll
    for (int j = 0; j < CHK; j++) { mn = fminf(mn, v[j]); mx = fmaxf(mx, v[j]); }
  } else {
#pragma unroll
    for (int j = 0; j < CHK; j++)
      if (CHK * tid + j < n) { mn = fminf(mn, v[j]); mx = fmaxf(mx, v[j]); }
  }
  sm->mn1[tid] = mn;
  sm->mx1[tid] = mx;
  const unsigned kmn = __reduce_min_sync(FULL, fkey(mn)), kmx = __reduce_max_sync(FULL, fkey(mx));
  if (lane == 0) { sm->mn2[warp] = fkey_inv(kmn); sm->mx2[warp] = fkey_inv(kmx); }
}

// exact check of the pairs whose upper (backward) / lower (forward) sample lies in chunk c
__device__ __forceinline__ int check_chunk(const float* w, int n, float thr, int c, int s, bool forward, int stop_back,
                                           int lane) {
  const int i = 16 * c + (lane & 15);
  bool hit = false;
  if (lane < 16) {
    if (forward) {
      if (i >= s && i < n - 1) {
        const float a = at(w, i), b = at(w, i + 1);
        hit = (a <= thr && thr < b) || (a >= thr && thr > b);
      }
    } else if (i <= s && i >= stop_back && i < n) {
      const float a = at(w, i - 1), b = at(w, i);
      hit = (a < thr && thr <= b) || (a > thr && thr >= b);
    }
  }
  const unsigned m = __ballot_sync(FULL, hit);
  if (!m) return -1;
  return 16 * c + (forward ? __ffs(m) - 1 : 31 - __clz(m));
}

__device__ __noinline__ int search_cross_far(const float* w, const WaveSummary* sm, int n, float thr, int s,
                                             bool forward, int stop_back, int lane);

// The common case -- the crossing lies within 32 samples of the start -- is one plain window
// (the scalar warp runs alone: every instruction here is ~5 cycles of critical path); longer
// walks go through the summary.
__device__ __forceinline__ int search_cross_w(const float* w, const WaveSummary* sm, int n, float thr, int s,
                                              bool forward, int stop_back, int lane) {
  constexpr int NEAR = 3;  // plain 32-sample windows before the summary is consulted
  if (forward) {
    int base = s;
#pragma unroll 1
    for (int k = 0; k < NEAR && base < n - 1; k++, base += 32) {
      const int i = base + lane;
      bool hit = false;
      if (i < n - 1) {
        const float a = at(w, i), b = at(w, i + 1);
        hit = (a <= thr && thr < b) || (a >= thr && thr > b);
      }
      const unsigned m = __ballot_sync(FULL, hit);
      if (m) return base + __ffs(m) - 1;
    }
    return base < n - 1 ? search_cross_far(w, sm, n, thr, base, true, stop_back, lane) : -1;
  }
  int base = s;
#pragma unroll 1
  for (int k = 0; k < NEAR && base >= stop_back; k++, base -= 32) {
    const int i = base - lane;
    bool hit = false;
    if (i >= stop_back) {
      const float a = at(w, i - 1), b = at(w, i);
      hit = (a < thr && thr <= b) || (a > thr && thr >= b);
    }
    const unsigned m = __ballot_sync(FULL, hit);
    if (m) return base - (__ffs(m) - 1);
  }
  return base >= stop_back ? search_cross_far(w, sm, n, thr, base, false, stop_back, lane) : -1;
}

__device__ __noinline__ int search_cross_far(const float* w, const WaveSummary* sm, int n, float thr, int s,
                                             bool forward, int stop_back, int lane) {
  const int nch = (n + 15) >> 4;
  const int cs = s >> 4;  // chunk of the start sample
  // level 2: stretch g is a candidate if thr lies within [min, max] of stretches g-1 .. g+1 clipped
  // (the neighbours cover pairs across stretch borders)
  unsigned cand2;
  {
    const int g = lane & 15;
    const float lo = fminf(sm->mn2[g], fminf(sm->mn2[max(g - 1, 0)], sm->mn2[min(g + 1, 15)]));
    const float hi = fmaxf(sm->mx2[g], fmaxf(sm->mx2[max(g - 1, 0)], sm->mx2[min(g + 1, 15)]));
    cand2 = __ballot_sync(FULL, lane < 16 && lo <= thr && thr <= hi);
  }
  const int gs = cs >> 5;
  // stretches on the far side of the start are irrelevant
  cand2 &= forward ? (0xffffffffu << gs) : (0xffffffffu >> (31 - gs));
  while (cand2) {
    const int g = forward ? __ffs(cand2) - 1 : 31 - __clz(cand2);
    cand2 &= ~(1u << g);
    // level 1: the 32 chunks of stretch g
    const int c = 32 * g + lane;
    bool cnd = false;
    if (c < nch && (forward ? c >= cs : c <= cs)) {
      const int cn = forward ? min(c + 1, nch - 1) : max(c - 1, 0);
      const float lo = fminf(sm->mn1[c], sm->mn1[cn]), hi = fmaxf(sm->mx1[c], sm->mx1[cn]);
      cnd = lo <= thr && thr <= hi;
    }
    unsigned cand1 = __ballot_sync(FULL, cnd);
    while (cand1) {
      // the two nearest candidate chunks at once: lanes 0-15 check the nearer one, 16-31 the next
      const int l0 = forward ? __ffs(cand1) - 1 : 31 - __clz(cand1);
      cand1 &= ~(1u << l0);
      int l1 = -1;
      if (cand1) {
        l1 = forward ? __ffs(cand1) - 1 : 31 - __clz(cand1);
        cand1 &= ~(1u << l1);
      }
      const int cc = 32 * g + (lane < 16 ? l0 : l1);
      const int i = 16 * cc + (lane & 15);
      bool hit = false;
      if (lane < 16 || l1 >= 0) {
        if (forward) {
          if (i >= s && i < n - 1) {
            const float a = at(w, i), b = at(w, i + 1);
            hit = (a <= thr && thr < b) || (a >= thr && thr > b);
          }
        } else if (i <= s && i >= stop_back && i < n) {
          const float a = at(w, i - 1), b = at(w, i);
          hit = (a < thr && thr <= b) || (a > thr && thr >= b);
        }
      }
      const unsigned m = __ballot_sync(FULL, hit);
      const unsigned m0 = m & 0xffffu, m1 = m >> 16;
      if (m0) return 16 * (32 * g + l0) + (forward ? __ffs(m0) - 1 : 31 - __clz(m0));
      if (m1) return 16 * (32 * g + l1) + (forward ? __ffs(m1) - 1 : 31 - __clz(m1));
    }
  }
  return -1;
}

// sum of w[a .. b) clipped to the wave, by one warp (lazy evaluation of a windowed filter at a
// single position, e.g. a trapezoid that is only picked off at one time)
__device__ __noinline__ float wrange_sum(const float* w, int n, int a, int b, int lane) {
  a = max(a, 0);
  b = min(b, n);
  float acc = 0.f;
  // 128 samples per step: lane l reads samples [g + 4l, g + 4l + 4) with one 128-bit load
  for (int g = a & ~3; g < b; g += 128) {
    const int i = g + 4 * lane;
    if (i < b) {
      const float4 q = *reinterpret_cast<const float4*>(w + sidx(i));
      acc += (i >= a && i < b ? q.x : 0.f) + (i + 1 >= a && i + 1 < b ? q.y : 0.f) +
             (i + 2 >= a && i + 2 < b ? q.z : 0.f) + (i + 3 >= a && i + 3 < b ? q.w : 0.f);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
  return acc;
}
__device__ __forceinline__ float at0(const float* w, int n, int i) { return (i >= 0 && i < n) ? at(w, i) : 0.f; }

// time_point_thresh.py:12-92 evaluated by one warp
__device__ __forceinline__ float tpt_w(const float* w, const WaveSummary* sm, int n, float thr, float t_start,
                                       float walk, int& fatal, int lane) {
  fatal = 0;
  if (thr != thr || t_start != t_start || walk != walk) return CUDART_NAN_F;
  if (floorf(t_start) != t_start) { fatal = DSPB_FATAL_TSTART_NONINT; return CUDART_NAN_F; }
  if (floorf(walk) != walk) { fatal = DSPB_FATAL_WALK_NONINT; return CUDART_NAN_F; }
  if (!(t_start >= 0.f && t_start < (float)n)) { fatal = DSPB_FATAL_TSTART_RANGE; return CUDART_NAN_F; }
  const int hit = search_cross_w(w, sm, n, thr, (int)t_start, walk == 1.0f, 1, lane);
  return hit < 0 ? CUDART_NAN_F : (float)hit;
}

// K backward searches on one waveform, each starting where the previous one ended (time_point_thresh.py:84-92 applied
// K times: tp_95 from tp_99, tp_90 from tp_95 ... in the LEGEND chains).  The crossings of a rising edge lie within a
// few samples of each other, so one walk over 32-sample windows serves all thresholds: the window's sample pairs stay
// in registers while consecutive thresholds are resolved in it (compare + ballot per threshold, no further loads),
// and only an empty window moves on (three plain windows, then the two-level summary).  Semantics per search are
// those of tpt_w: the first start is validated like any t_start, a search without a crossing (or with a NaN
// threshold) yields NaN and so do all later ones (their start is NaN).
template <int K>
__device__ __forceinline__ void tpt_chain_bwd(const float* w, const WaveSummary* sm, int n, const float (&thr)[K],
                                              float t_start, float (&out)[K], int& fatal, int lane) {
  fatal = 0;
#pragma unroll
  for (int k = 0; k < K; k++) out[k] = CUDART_NAN_F;
  if (t_start != t_start) return;
  if (floorf(t_start) != t_start) { fatal = DSPB_FATAL_TSTART_NONINT; return; }
  if (!(t_start >= 0.f && t_start < (float)n)) { fatal = DSPB_FATAL_TSTART_RANGE; return; }
  int pos = (int)t_start;   // start of the current search (inclusive)
  int base = pos;           // lane l of the current window looks at the pair (base - l - 1, base - l)
  bool have = false;
  float a = 0.f, b = 0.f;
  int i = 0;
#pragma unroll
  for (int k = 0; k < K; k++) {
    const float t = thr[k];
    if (pos < 1 || t != t) return;
    int found = -1, tries = 0;
#pragma unroll 1
    while (true) {
      if (!have) {
        i = base - lane;
        a = i >= 1 ? at(w, i - 1) : 0.f;
        b = i >= 1 ? at(w, i) : 0.f;
        have = true;
      }
      const bool hit = i >= 1 && i <= pos && ((a < t && t <= b) || (a > t && t >= b));
      const unsigned m = __ballot_sync(FULL, hit);
      if (m) { found = base - (__ffs(m) - 1); break; }
      base -= 32;
      have = false;
      if (base < 1) break;
      if (++tries >= 3) {
        found = search_cross_far(w, sm, n, t, base, false, 1, lane);
        if (found >= 1) base = found;
        break;
      }
    }
    if (found < 1) return;
    out[k] = (float)found;
    pos = found;
  }
}

// =========================================================================================
// 'valid' convolution with cusp / zac kernels (energy_kernels.py:12-157), chunked form.
//
// Same mathematics as op_conv_seg (conv_seg.cuh): with z[j] = x[j] - c x[j-1] the kernel is
// sinh ramps + flat top (+ beta * parabolas for zac), so every output is a combination of the
// prefix sums  Em = sum e^{-j/s} z,  Ep = sum e^{+j/s} z,  P0 = sum z  (and M1 = sum jc z,
// M2 = sum jc^2 z with jc = j - N/2 for the parabolas) at four window bounds
//   hiA = L + o,  loA = L - lt + o,  loB = L - 1 - lt - fl + o,  loC = o      (o = output index).
// Here: (1) one pass over the thread's chunk for the chunk sums and ONE scan round; (2) only
// the threads whose chunk meets one of the four p-wide bands replay it (running sums only)
// and deposit the prefix values at the band positions; (3) the p outputs are combined in
// parallel, one thread each, with the powers q^o = e^{+-o/s} read from a table built by the
// chain compiler.  Two kernels that share (sigma, lt, fl, L, c) -- the ICPC chain's cusp and
// zac -- are evaluated together (TWO): sums, scan, replay and exponentials are shared.
// tab: NQ * (4 * 16 * CW + 544) doubles of scratch (NQ = 5 with parabolas, else 3; CW = odd number of
// chunk columns >= ceil(p / 16) + 1): entry (o % 16) * CW + o / 16 of a band, so that the lanes of a
// warp (consecutive chunks, positions 16 apart) store to consecutive addresses.
// pw: [2][p] doubles, pw[0][o] = e^{o/s}, pw[1][o] = e^{-o/s}.
// Ends with a barrier; out* are complete slots.
// =========================================================================================
struct SegOut {
  double kL, beta, h;  // k[L-1], parabola weight (0: cusp), parabola vertex
};
// Where an output of the convolution goes.  `slot` != null: the waveform is materialised.  When
// its only consumers are numpy.amax and pick-offs at fixed integer positions (the ICPC chain's
// cuspEmax / cuspEftp), the block warps hand the scalar warp the per-warp maxima (`mx`, 16 ints:
// float keys) and the picked sample (`pick_dst[0]`) through the mailbox and no slot is needed.
struct SegSink {
  float* slot;
  int* mx;
  double* pick_dst;
  int pick;
};

// passes 2 and 3 of the chunked cusp / zac convolution (shared by both pass-1 variants): `s` = this
// thread's chunk sums, the band table `tab` holds the chunk-local running sums at the band positions
template <bool POLY, bool TWO>
__device__ __forceinline__ void conv_seg_finish(const float* X, int N, int lt, int fl, int L, double c, double inv2S,
                                                double qm, double qp, double eA, const double* __restrict__ pw,
                                                SegOut s0, SegOut s1, SegSink k0, SegSink k1, double* tab,
                                                const double (&s)[POLY ? 5 : 3], int tid, int lane, int warp) {
  constexpr int NQ = POLY ? 5 : 3;
  const int p = N - L + 1;
  // operands of pass 3 that come from global memory: requested now, consumed two barriers later
  const double po_pf = tid < p ? pw[tid] : 0.0, mo_pf = tid < p ? pw[p + tid] : 0.0;
  const int CW = (((p + CHK - 1) >> 4) + 1) | 1, PP = CHK * CW;
  const double j0 = 0.5 * (double)N;
  const int base[4] = {L, L - lt, L - 1 - lt - fl, 0};
  PROF_SUB_BEGIN();
  // ---- pass 2: exclusive scan of the chunk sums over the 512 threads, through shared memory:
  // every thread deposits its NQ sums (one store each), warp q scans quantity q (lane l owns
  // entries [16l, 16l+16), skewed by one word per 16 so that both access patterns are
  // conflict-free), and pass 3 picks up the prefix in front of whatever chunk it needs.  Costs
  // the block NQ stores per thread instead of NQ float64 shuffle scans per thread.
  double* otab = tab + NQ * 4 * PP;
  constexpr int OT = 512 + 32;
#pragma unroll
  for (int q = 0; q < NQ; q++) otab[q * OT + tid + (tid >> 4)] = s[q];
  BSYNC();
  PROF_SUB(1);   // wait for the slowest pass-1 thread
  // (warps 1,2,3,5,6: the scalar warp shares scheduler partition 0 with warps 0,4,8,12)
  const int sq = warp < 4 ? warp - 1 : (warp == 5 ? 3 : (warp == 6 ? 4 : -1));
  if (sq >= 0 && sq < NQ) {
    double* o = otab + sq * OT + 17 * lane;
    double v[CHK];
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < CHK; k++) {
      v[k] = run;       // exclusive inside the lane
      run += o[k];
    }
    const double incl = wscan_incl(run, lane);
    const double base_l = incl - run;
#pragma unroll
    for (int k = 0; k < CHK; k++) o[k] = v[k] + base_l;
  }
  BSYNC();
  PROF_SUB(2);   // pass 2 (scan warps)
  // ---- pass 3: one thread per output ------------------------------------------------------------
  const int pceil = (p + CHK - 1) & ~(CHK - 1);
  const double eAm = 1.0 / eA;  // eA = e^{(L-1)/s}: e^{+-n/s} = eA^{+-1} * q^{+-o}
  float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
  for (int o = tid; o < pceil; o += 512) {
    float y0 = 0.f, y1 = 0.f;
    if (o < p) {
      const int ow[4] = {(base[0] + o) >> 4, (base[1] + o) >> 4, (base[2] + o) >> 4, (base[3] + o) >> 4};
#define TB(q, b) (tab[((q)*4 + (b)) * PP + (o & 15) * CW + (o >> 4)] + otab[(q)*OT + ow[b] + (ow[b] >> 4)])
      const double po = o == tid ? po_pf : pw[o], mo = o == tid ? mo_pf : pw[p + o];
      const double en = eA * po, enm = eAm * mo;      // e^{+-n/s}, n = L - 1 + o
      const double eLn = qp * mo, eLnm = qm * po;     // e^{+-(L-n)/s} = e^{+-(1-o)/s}
      const double yA = (en * (TB(0, 0) - TB(0, 1)) - enm * (TB(1, 0) - TB(1, 1))) * inv2S;
      const double yB = TB(2, 1) - TB(2, 2);
      const double yC = (eLn * (TB(1, 2) - TB(1, 3)) - eLnm * (TB(0, 2) - TB(0, 3))) * inv2S;
      const double xm = o >= 1 ? (double)at(X, o - 1) : 0.0;
      double ysh = yA + yB + yC;
      double v0 = ysh + c * s0.kL * xm, v1 = ysh + c * s1.kL * xm;
      if (POLY) {
        const double n = (double)(L - 1 + o), nc = n - j0, a = ((double)L - n) + j0;
        const double dPa = TB(2, 0) - TB(2, 1), dM1a = TB(3, 0) - TB(3, 1), dM2a = TB(4, 0) - TB(4, 1);
        const double s2a = nc * nc * dPa - 2.0 * nc * dM1a + dM2a, s1a = nc * dPa - dM1a;
        const double dPc = TB(2, 2) - TB(2, 3), dM1c = TB(3, 2) - TB(3, 3), dM2c = TB(4, 2) - TB(4, 3);
        const double s2c = a * a * dPc + 2.0 * a * dM1c + dM2c, s1c = a * dPc + dM1c;
        v0 += s0.beta * ((s2a - 2.0 * s0.h * s1a) + (s2c - 2.0 * s0.h * s1c));
        v1 += s1.beta * ((s2a - 2.0 * s1.h * s1a) + (s2c - 2.0 * s1.h * s1c));
      }
#undef TB
      y0 = (float)v0;
      y1 = (float)v1;
      mx0 = fmaxf(mx0, y0);
      mx1 = fmaxf(mx1, y1);
      if (k0.pick_dst && o == k0.pick) k0.pick_dst[0] = (double)y0;
      if (TWO && k1.pick_dst && o == k1.pick) k1.pick_dst[0] = (double)y1;
    }
    if (k0.slot) k0.slot[sidx(o)] = y0;
    if (TWO && k1.slot) k1.slot[sidx(o)] = y1;
  }
  PROF_SUB(3);   // pass 3 (this thread)
  if (k0.mx) put_fmax(k0.mx, mx0, lane, warp);
  if (TWO && k1.mx) put_fmax(k1.mx, mx1, lane, warp);
  BSYNC();
  PROF_SUB(4);   // wait for the slowest pass-3 thread
}

template <bool POLY, bool TWO>
__device__ __forceinline__ void conv_seg_chunked(const float* X, const float (&x)[CHK], int N, double sigma, int lt,
                                                 int fl, int L, double c, double inv2S, double qm, double qp,
                                                 double eA, const double* __restrict__ pw, SegOut s0, SegOut s1,
                                                 SegSink k0, SegSink k1, double* tab, int tid, int lane, int warp) {
  constexpr int NQ = POLY ? 5 : 3;
  const int p = N - L + 1;
  const int CW = (((p + CHK - 1) >> 4) + 1) | 1, PP = CHK * CW;
  const int i0 = CHK * tid;
  const double j0 = 0.5 * (double)N;
  const float xm1 = (i0 > 0 && i0 < N) ? at(X, i0 - 1) : 0.f;
  const int base[4] = {L, L - lt, L - 1 - lt - fl, 0};
  // ---- pass 1: chunk sums; threads whose chunk meets a band also deposit the chunk-local
  // running sums (sum over the chunk's elements before position j) at the band positions ----
  PROF_SUB_BEGIN();
  double s[NQ];
#pragma unroll
  for (int q = 0; q < NQ; q++) s[q] = 0.0;
  bool hit = false;
#pragma unroll
  for (int b = 0; b < 4; b++) hit |= (i0 + CHK - 1 >= base[b] && i0 <= base[b] + p - 1);
  hit &= i0 <= N;
  if (i0 < N || hit) {
    const double wm0 = exp(-(double)i0 / sigma);
    double wm = wm0, wp = 1.0 / wm0, jc = (double)i0 - j0;
    float prev = xm1;
    auto step = [&](int k) {
      const double z = i0 + k < N ? fma(-c, (double)prev, (double)x[k]) : 0.0;
      prev = x[k];
      s[0] = fma(wm, z, s[0]);
      s[1] = fma(wp, z, s[1]);
      s[2] += z;
      if (POLY) {
        const double jz = jc * z;
        s[3] += jz;
        s[4] = fma(jc, jz, s[4]);
        jc += 1.0;
      }
      wm *= qm;
      wp *= qp;
    };
    if (!hit) {
      // the common case (3 of 4 warps): no band position in this chunk, sums only
#pragma unroll
      for (int k = 0; k < CHK; k++) step(k);
    } else {
      // o of the chunk's first position in every band; position j = i0 + k lies in band b iff
      // 0 <= ob[b] + k < p (and j <= N)
      const int ob[4] = {i0 - base[0], i0 - base[1], i0 - base[2], i0 - base[3]};
#pragma unroll
      for (int k = 0; k < CHK; k++) {
        if (i0 + k <= N) {
#pragma unroll
          for (int b = 0; b < 4; b++) {
            const int o = ob[b] + k;
            if ((unsigned)o < (unsigned)p) {
              double* t = tab + b * PP + (o & 15) * CW + (o >> 4);
#pragma unroll
              for (int q = 0; q < NQ; q++) t[q * 4 * PP] = s[q];
            }
          }
        }
        step(k);
      }
    }
  }
  PROF_SUB(0);   // pass 1 (this thread)
  conv_seg_finish<POLY, TWO>(X, N, lt, fl, L, c, inv2S, qm, qp, eA, pw, s0, s1, k0, k1, tab, s, tid, lane, warp);
}

// Running sums of (band, chunk) pair `h` at the chunk's positions [K0, K1): the chunk is re-read from shared
// memory, the sums of the positions before K0 are accumulated without deposits.  Branch-free: positions outside
// the band write to a column of the band table that no band position uses (CW - 1 > (p - 1) / 16), so the
// positions form ONE basic block and the conversions / FP64 products / stores of consecutive positions overlap.
template <bool POLY, int K0, int K1>
__device__ __forceinline__ void seg_deposit(const float* X, int N, int p, int CW, int PP, const int (&base)[4], int h,
                                            double j0, float omc, const double* __restrict__ pwc, int NPW,
                                            const float (&wm)[CHK], const float (&wp)[CHK], double* tab) {
  int b = -1, ch = 0, bbase = 0;
#pragma unroll
  for (int bb = 0; bb < 4; bb++) {
    const int c_lo = base[bb] >> 4, cnt = ((base[bb] + p - 1) >> 4) - c_lo + 1;
    if (b < 0) {
      if (h < cnt) { b = bb; ch = c_lo + h; bbase = base[bb]; }
      else h -= cnt;
    }
  }
  if (b < 0 || CHK * ch > N) return;
  const int i0 = CHK * ch;
  float xs[CHK];
  ld_chunk(X, ch, xs);
  float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f, r4 = 0.f;
  float prev = i0 > 0 ? at(X, i0 - 1) : 0.f;
  const double w0 = pwc[ch], w0i = pwc[NPW + ch];
  const double jc0 = (double)i0 - j0, jc02 = jc0 * jc0, jc2 = 2.0 * jc0;
  const int ob = i0 - bbase;
  double* tb = tab + b * PP;
#pragma unroll
  for (int k = 0; k < K1; k++) {
    if (k >= K0) {
      const int o = ob + k;
      const bool dep = (unsigned)o < (unsigned)p && i0 + k <= N;
      double* t = tb + (dep ? (o & 15) * CW + (o >> 4) : CW - 1);
      const double d2 = (double)r2, d3 = (double)r3;
      t[0] = w0 * (double)r0;
      t[4 * PP] = w0i * (double)r1;
      t[2 * 4 * PP] = d2;
      if (POLY) {
        t[3 * 4 * PP] = fma(jc0, d2, d3);
        t[4 * 4 * PP] = fma(jc02, d2, fma(jc2, d3, (double)r4));
      }
    }
    float z = (xs[k] - prev) + omc * prev;
    z = i0 + k < N ? z : 0.f;
    prev = xs[k];
    r0 = fmaf(wm[k], z, r0);
    r1 = fmaf(wp[k], z, r1);
    r2 += z;
    if (POLY) {
      r3 = fmaf((float)k, z, r3);
      r4 = fmaf((float)(k * k), z, r4);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Pass 1 with helper threads.  The chunk sums are float32 inside a 16-sample chunk (weights relative
// to the chunk start, compile-time tables `wm` / `wp`; z = (x[k] - x[k-1]) + (1 - c) x[k-1] keeps the
// pole-zero difference exact) and float64 across chunks (`pwc[t]` = e^{-16 t / sigma}, `pwc[NPW + t]`
// its inverse).  The (band, chunk) pairs that need the running sums at every position are handed to the
// threads beyond the end of the input (tid >= HB, idle otherwise; first half of the positions) and to the
// four owner warps below them (second half, after their own chunk), which re-read the chunk from shared
// memory -- the 4 warps that used to be the critical path (16 x 4 predicated bands x 5 stores) are gone.
// ---------------------------------------------------------------------------------------
template <bool POLY, bool TWO, int HB, int NPW>
__device__ __forceinline__ void conv_seg_chunked_h(const float* X, const float (&x)[CHK], int N, double sigma, int lt,
                                                   int fl, int L, double c, double inv2S, double qm, double qp,
                                                   double eA, const double* __restrict__ pw,
                                                   const double* __restrict__ pwc, const float (&wm)[CHK],
                                                   const float (&wp)[CHK], SegOut s0, SegOut s1, SegSink k0,
                                                   SegSink k1, double* tab, int tid, int lane, int warp) {
  constexpr int NQ = POLY ? 5 : 3;
  const int p = N - L + 1;
  const int CW = (((p + CHK - 1) >> 4) + 1) | 1, PP = CHK * CW;
  const double j0 = 0.5 * (double)N;
  const float omc = (float)(1.0 - c);
  const int base[4] = {L, L - lt, L - 1 - lt - fl, 0};
  (void)sigma;
  PROF_SUB_BEGIN();
  double s[NQ];
#pragma unroll
  for (int q = 0; q < NQ; q++) s[q] = 0.0;
  if (tid < HB) {
    // ---- owner threads: chunk sums of the own chunk (registers) ------------------------------------
    const int i0 = CHK * tid;
    if (i0 < N) {
      float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f, r4 = 0.f;
      float prev = i0 > 0 ? at(X, i0 - 1) : 0.f;
      const bool whole = i0 + CHK <= N;
#pragma unroll
      for (int k = 0; k < CHK; k++) {
        float z = (x[k] - prev) + omc * prev;
        if (!whole) z = i0 + k < N ? z : 0.f;
        prev = x[k];
        r0 = fmaf(wm[k], z, r0);
        r1 = fmaf(wp[k], z, r1);
        r2 += z;
        if (POLY) {
          r3 = fmaf((float)k, z, r3);
          r4 = fmaf((float)(k * k), z, r4);
        }
      }
      const double jc0 = (double)i0 - j0;
      s[0] = pwc[tid] * (double)r0;
      s[1] = pwc[NPW + tid] * (double)r1;
      s[2] = (double)r2;
      if (POLY) {
        s[3] = fma(jc0, s[2], (double)r3);
        s[4] = fma(jc0 * jc0, s[2], fma(2.0 * jc0, (double)r3, (double)r4));
      }
    }
  }
  // ---- band deposits: (band, chunk) pair h, positions [0, 8) by the helper threads (tid >= HB, idle otherwise),
  // positions [8, 16) by the owner threads of the four warps below them once their own chunk is done ------------
  if (tid >= HB) seg_deposit<POLY, 0, CHK / 2>(X, N, p, CW, PP, base, tid - HB, j0, omc, pwc, NPW, wm, wp, tab);
  else if (tid >= HB - 128) seg_deposit<POLY, CHK / 2, CHK>(X, N, p, CW, PP, base, tid - (HB - 128), j0, omc, pwc, NPW, wm, wp, tab);
  PROF_SUB(0);   // pass 1 (this thread)
  conv_seg_finish<POLY, TWO>(X, N, lt, fl, L, c, inv2S, qm, qp, eA, pw, s0, s1, k0, k1, tab, s, tid, lane, warp);
}

}  // namespace crt
