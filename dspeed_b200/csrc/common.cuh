// dspeed_b200 -- device-side building blocks shared by the per-processor kernels and
// the fused waveform-resident chain kernel (sm_100a).
//
// Execution model: one CTA owns one waveform ("row") at a time.  The row is staged
// once from HBM into shared memory (padded layout below) and every processor is a
// block-wide routine over shared-memory "slots"; scalars travel in registers /
// a small shared scalar file.  Reductions and searches use warp shuffles, the
// recursive filters of the reference (pole-zero, trapezoids, moving averages) are
// evaluated as prefix sums: each thread owns a contiguous chunk (sequential, in
// registers, float64 accumulator), chunks are stitched with one shuffle scan.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "dspeed_b200.h"

namespace dspb {

// Row-resident kernels are written for any 1-D CTA size that is a multiple of 32 (the
// per-processor kernels launch 256 threads, the fused chain kernel 512): NT / NW read the
// launch configuration.
#ifdef DSPB_PSP
// specialised chain kernels (csrc/chain_rt.cuh) are always launched with 512 threads
#define NT 512
#define NW 16
#else
#define NT ((int)blockDim.x)
#define NW ((int)(blockDim.x >> 5))
#endif
constexpr int MAXW = 32;             // max warps per CTA
constexpr int SCRATCH_BYTES = 2048;  // block-primitive scratch at the start of dynamic smem
constexpr int DEFAULT_THREADS = 256;

// Padded shared-memory index: one pad word per 32 elements.  With a thread-contiguous
// chunk of C in {4,8,16,32} elements, lane l at step j touches bank (l*C + j + (l*C+j)/32) % 32,
// which is a permutation of the 32 banks; unit-stride (thread-strided) access stays
// conflict-free as well.
#ifdef DSPB_PSP
// "T4" layout of the specialised chain kernels: the waveform is cut into 16-sample chunks
// (chunk c = samples [16c, 16c+16) belongs to thread c); float4 number p of every chunk
// lives in plane p, planes are DSPB_PSP floats apart (= 4 * chunks + 8 pad words).  A thread
// reads or writes its own chunk -- or the chunk a fixed number of samples away -- with four
// or five conflict-free 128-bit accesses at immediate offsets; unit-stride access is
// conflict-free as well thanks to the 8-word plane skew.
__device__ __forceinline__ int sidx(int i) { return ((i >> 2) & 3) * DSPB_PSP + ((i >> 4) << 2) + (i & 3); }
__host__ __device__ __forceinline__ int slot_words(int n) { return 4 * DSPB_PSP; }
#else
__device__ __forceinline__ int sidx(int i) { return i + (i >> 5); }
__host__ __device__ __forceinline__ int slot_words(int n) { return n + (n >> 5) + 1; }
#endif

template <typename T> __device__ __forceinline__ T nan_of();
template <> __device__ __forceinline__ float nan_of<float>() { return CUDART_NAN_F; }
template <> __device__ __forceinline__ double nan_of<double>() { return CUDART_NAN; }

// Per-row scalar argument: device array (stride 0 = broadcast, 1 = per row) or immediate.
template <typename T>
struct Scalar {
  const T* ptr;
  long long stride;
  T imm;
  __device__ __forceinline__ T get(long long row) const { return ptr ? ptr[row * stride] : imm; }
};

// A waveform operand in global memory.
struct Wave {
  const void* ptr;
  long long row_stride;  // elements
  int dtype;             // DSPB_F32 / DSPB_F64 / DSPB_U16 / ...
};

struct Scratch {
  double d[MAXW * 4];
  int i[MAXW * 4];
};
static_assert(sizeof(Scratch) <= SCRATCH_BYTES, "scratch too large");

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// ---- block-wide primitives (all threads of the CTA must call) -------------------
// Two-level: shuffles inside each warp, one shared-memory round trip for the (at most 32)
// warp partials, then every warp combines the partials again with shuffles (lane k holds
// the partial of warp k), so the cross-warp phase costs ~5 shuffle steps, not NW loads.

// Exclusive prefix sum of one double per thread; returns the exclusive prefix,
// writes the block total.
__device__ __forceinline__ double block_excl_scan(double v, double& total, Scratch* sc) {
  const int lane = lane_id(), w = warp_id();
  double incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();  // protect scratch reuse
  if (lane == 31) sc->d[w] = incl;
  __syncthreads();
  double part = lane < NW ? sc->d[lane] : 0.0;
  double pin = part;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, pin, o);
    if (lane >= o) pin += t;
  }
  total = __shfl_sync(0xffffffffu, pin, 31);
  const double woff = __shfl_sync(0xffffffffu, pin - part, w);
  return woff + (incl - v);
}

// Same, scanning from the last thread towards the first (suffix sums).
__device__ __forceinline__ double block_excl_scan_rev(double v, double& total, Scratch* sc) {
  const int lane = lane_id(), w = warp_id();
  double incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += t;
  }
  __syncthreads();
  if (lane == 0) sc->d[w] = incl;
  __syncthreads();
  double part = lane < NW ? sc->d[lane] : 0.0;
  double pin = part;  // suffix-inclusive over warps
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_down_sync(0xffffffffu, pin, o);
    if (lane + o < 32) pin += t;
  }
  total = __shfl_sync(0xffffffffu, pin, 0);
  const double woff = __shfl_sync(0xffffffffu, pin - part, w);
  return woff + (incl - v);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double block_sum(double v, Scratch* sc) {
  v = warp_sum(v);
  __syncthreads();
  if (lane_id() == 0) sc->d[warp_id()] = v;
  __syncthreads();
  return warp_sum(lane_id() < NW ? sc->d[lane_id()] : 0.0);
}

// two sums at once (saves barriers)
__device__ __forceinline__ void block_sum2(double& a, double& b, Scratch* sc) {
  a = warp_sum(a);
  b = warp_sum(b);
  __syncthreads();
  if (lane_id() == 0) {
    sc->d[warp_id()] = a;
    sc->d[MAXW + warp_id()] = b;
  }
  __syncthreads();
  a = warp_sum(lane_id() < NW ? sc->d[lane_id()] : 0.0);
  b = warp_sum(lane_id() < NW ? sc->d[MAXW + lane_id()] : 0.0);
}

__device__ __forceinline__ int block_or(int pred) { return __syncthreads_or(pred); }

// Smallest int over the block (INT_MAX = none).
__device__ __forceinline__ int block_min_int(int v, Scratch* sc) {
  v = __reduce_min_sync(0xffffffffu, v);
  __syncthreads();
  if (lane_id() == 0) sc->i[warp_id()] = v;
  __syncthreads();
  return __reduce_min_sync(0xffffffffu, lane_id() < NW ? sc->i[lane_id()] : 0x7fffffff);
}
__device__ __forceinline__ int block_max_int(int v, Scratch* sc) {
  v = __reduce_max_sync(0xffffffffu, v);
  __syncthreads();
  if (lane_id() == 0) sc->i[warp_id()] = v;
  __syncthreads();
  return __reduce_max_sync(0xffffffffu, lane_id() < NW ? sc->i[lane_id()] : (int)0x80000000);
}

// First-occurrence arg-min and arg-max (strict comparisons, as min_max.py:73-77).
template <typename T>
__device__ __forceinline__ void warp_argminmax(T& vmin, int& imin, T& vmax, int& imax) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T ov = __shfl_xor_sync(0xffffffffu, vmin, o);
    int oi = __shfl_xor_sync(0xffffffffu, imin, o);
    if (ov < vmin || (ov == vmin && oi < imin)) { vmin = ov; imin = oi; }
    ov = __shfl_xor_sync(0xffffffffu, vmax, o);
    oi = __shfl_xor_sync(0xffffffffu, imax, o);
    if (ov > vmax || (ov == vmax && oi < imax)) { vmax = ov; imax = oi; }
  }
}
template <typename T>
__device__ __forceinline__ void block_argminmax(T& vmin, int& imin, T& vmax, int& imax, Scratch* sc) {
  warp_argminmax<T>(vmin, imin, vmax, imax);
  __syncthreads();
  if (lane_id() == 0) {
    sc->d[warp_id()] = (double)vmin;
    sc->d[MAXW + warp_id()] = (double)vmax;
    sc->i[warp_id()] = imin;
    sc->i[MAXW + warp_id()] = imax;
  }
  __syncthreads();
  // lanes >= NW replicate warp 0's partial (harmless for first-occurrence semantics)
  const int k = lane_id() < NW ? lane_id() : 0;
  vmin = (T)sc->d[k];
  vmax = (T)sc->d[MAXW + k];
  imin = sc->i[k];
  imax = sc->i[MAXW + k];
  warp_argminmax<T>(vmin, imin, vmax, imax);
}

// ---- row staging -----------------------------------------------------------------

template <typename T, typename TIn>
__device__ __forceinline__ int stage_row_typed(T* s, const TIn* g, int n) {
  int has_nan = 0;
  for (int i = threadIdx.x; i < n; i += NT) {
    T v = (T)g[i];
    if (v != v) has_nan = 1;
    s[sidx(i)] = v;
  }
  return has_nan;
}

// 128-bit path for uint16 rows whose start is 16-byte aligned (always true for the
// 8192-sample LEGEND records): 8 samples per load, coalesced.
template <typename T>
__device__ __forceinline__ void stage_row_u16_vec(T* s, const uint16_t* g, int n) {
  const int nv = n >> 3;
  const uint4* gv = reinterpret_cast<const uint4*>(g);
  for (int v = threadIdx.x; v < nv; v += NT) {
    uint4 q;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                 : "l"(gv + v));
    const int i = v << 3;
    const int b = sidx(i);  // the 8 samples share one 32-group (i % 8 == 0)
    s[b + 0] = (T)(q.x & 0xffffu);
    s[b + 1] = (T)(q.x >> 16);
    s[b + 2] = (T)(q.y & 0xffffu);
    s[b + 3] = (T)(q.y >> 16);
    s[b + 4] = (T)(q.z & 0xffffu);
    s[b + 5] = (T)(q.z >> 16);
    s[b + 6] = (T)(q.w & 0xffffu);
    s[b + 7] = (T)(q.w >> 16);
  }
  for (int i = (nv << 3) + threadIdx.x; i < n; i += NT) s[sidx(i)] = (T)g[i];
}

// Stage one row of a global waveform operand into a shared slot, converting to T.
// Returns (per thread) whether this thread saw a NaN; caller ORs over the block.
template <typename T>
__device__ __forceinline__ int stage_row(T* s, const Wave& w, long long row, int n) {
  switch (w.dtype) {
    case DSPB_U16: {
      const uint16_t* g = reinterpret_cast<const uint16_t*>(w.ptr) + row * w.row_stride;
      if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) stage_row_u16_vec<T>(s, g, n);
      else stage_row_typed<T, uint16_t>(s, g, n);
      return 0;
    }
    case DSPB_F32: return stage_row_typed<T, float>(s, reinterpret_cast<const float*>(w.ptr) + row * w.row_stride, n);
    case DSPB_F64: return stage_row_typed<T, double>(s, reinterpret_cast<const double*>(w.ptr) + row * w.row_stride, n);
    case DSPB_I16: return stage_row_typed<T, int16_t>(s, reinterpret_cast<const int16_t*>(w.ptr) + row * w.row_stride, n);
    case DSPB_I32: return stage_row_typed<T, int32_t>(s, reinterpret_cast<const int32_t*>(w.ptr) + row * w.row_stride, n);
    case DSPB_U32: return stage_row_typed<T, uint32_t>(s, reinterpret_cast<const uint32_t*>(w.ptr) + row * w.row_stride, n);
  }
  return 0;
}

template <typename T>
__device__ __forceinline__ void store_row(T* g, const T* s, int n) {
  for (int i = threadIdx.x; i < n; i += NT) g[i] = s[sidx(i)];
}
template <typename T>
__device__ __forceinline__ void store_row_nan(T* g, int n) {
  const T v = nan_of<T>();
  for (int i = threadIdx.x; i < n; i += NT) g[i] = v;
}
template <typename T>
__device__ __forceinline__ void fill_slot_nan(T* s, int n) {
  const T v = nan_of<T>();
  for (int i = threadIdx.x; i < n; i += NT) s[sidx(i)] = v;
}

// Record the first data-dependent fatal condition (the reference's DSPFatal) of a launch.
__device__ __forceinline__ void raise_fatal(int* fatal, int code, long long row) {
  if (fatal && code) {
    if (atomicCAS(fatal, 0, code) == 0) {
      fatal[1] = (int)(row & 0x7fffffff);
      fatal[2] = (int)(row >> 31);
    }
  }
}

// Thread-contiguous chunk of a row of n samples.
__device__ __forceinline__ void chunk_range(int n, int& lo, int& hi) {
  const int c = (n + NT - 1) / NT;
  lo = min(n, (int)threadIdx.x * c);
  hi = min(n, lo + c);
}

}  // namespace dspb
