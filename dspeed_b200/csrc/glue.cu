// dspeed_b200 -- the numpy-level glue of a processing chain as ONE element-wise kernel family: the ufuncs the
// reference's expression parser emits (processing_chain.py:46-59: add, subtract, multiply, divide, floor_divide,
// negative, comparisons, isnan, isfinite, ...), `where` (processors/where.py:12-54), the round / floor / ceil / trunc
// to-nearest helpers (round_to_nearest.py:11-200), the unit conversion of coordinates (unit_conversion.py:16-78),
// `astype` (processing_chain.py:1269-1300) and `get` / `get_default` (get.py:10-91).
//
// In the fused chain kernels these are scalar epilogues; this file is the per-processor tier: every operand is a
// strided [rows, inner] view (row / column stride 0 broadcasts: per-event scalars against waveforms, constants) or an
// immediate, loaded in its own dtype and converted to the LOOP type numpy's type resolution picks -- float32,
// float64 or int64 (integer loops are evaluated in 64 bits and wrapped by the store, like the native width would) --
// so that results are bit-identical to the IEEE basic operations numpy performs.
#include <cmath>

#include "common.cuh"

using namespace dspb;

namespace {

enum { GT_F32 = 0, GT_F64 = 1, GT_U16 = 2, GT_I16 = 3, GT_I32 = 4, GT_U32 = 5, GT_I64 = 6, GT_BOOL = 7, GT_I8 = 8, GT_U8 = 9, GT_U64 = 10 };

enum {
  GOP_ADD = 0, GOP_SUB, GOP_MUL, GOP_DIV, GOP_FLOORDIV, GOP_NEG, GOP_ABS, GOP_SQRT, GOP_MAX, GOP_MIN,
  GOP_EQ = 16, GOP_NE, GOP_LT, GOP_LE, GOP_GT, GOP_GE, GOP_ISNAN, GOP_ISFINITE,
  GOP_WHERE = 32, GOP_COPY,
  GOP_ROUND = 40, GOP_FLOOR, GOP_CEIL, GOP_TRUNC,          // to_nearest * f(val / to_nearest)
  GOP_CONVERT = 48,                                        // (a + b) * ratio - c, optional rounding mode
};

struct Opnd {
  const void* p;       // null: immediate
  long long rs, cs;    // row / column stride in elements (0 broadcasts)
  int dt;
  double imm;
};

template <typename C>
__device__ __forceinline__ C ld(const Opnd& o, long long r, long long j) {
  if (!o.p) return (C)o.imm;
  const long long i = r * o.rs + j * o.cs;
  switch (o.dt) {
    case GT_F32: return (C) reinterpret_cast<const float*>(o.p)[i];
    case GT_F64: return (C) reinterpret_cast<const double*>(o.p)[i];
    case GT_U16: return (C) reinterpret_cast<const uint16_t*>(o.p)[i];
    case GT_I16: return (C) reinterpret_cast<const int16_t*>(o.p)[i];
    case GT_I32: return (C) reinterpret_cast<const int32_t*>(o.p)[i];
    case GT_U32: return (C) reinterpret_cast<const uint32_t*>(o.p)[i];
    case GT_I64: return (C) reinterpret_cast<const long long*>(o.p)[i];
    case GT_U64: return (C) reinterpret_cast<const unsigned long long*>(o.p)[i];
    case GT_BOOL: case GT_U8: return (C) reinterpret_cast<const uint8_t*>(o.p)[i];
    case GT_I8: return (C) reinterpret_cast<const int8_t*>(o.p)[i];
  }
  return (C)0;
}

template <typename C>
__device__ __forceinline__ void st(void* p, int dt, long long i, C v) {
  switch (dt) {
    case GT_F32: reinterpret_cast<float*>(p)[i] = (float)v; break;
    case GT_F64: reinterpret_cast<double*>(p)[i] = (double)v; break;
    case GT_U16: reinterpret_cast<uint16_t*>(p)[i] = (uint16_t)(long long)v; break;
    case GT_I16: reinterpret_cast<int16_t*>(p)[i] = (int16_t)(long long)v; break;
    case GT_I32: reinterpret_cast<int32_t*>(p)[i] = (int32_t)(long long)v; break;
    case GT_U32: reinterpret_cast<uint32_t*>(p)[i] = (uint32_t)(long long)v; break;
    case GT_I64: reinterpret_cast<long long*>(p)[i] = (long long)v; break;
    case GT_U64: reinterpret_cast<unsigned long long*>(p)[i] = (unsigned long long)(long long)v; break;
    case GT_BOOL: reinterpret_cast<uint8_t*>(p)[i] = v != (C)0 ? 1 : 0; break;
    case GT_U8: reinterpret_cast<uint8_t*>(p)[i] = (uint8_t)(long long)v; break;
    case GT_I8: reinterpret_cast<int8_t*>(p)[i] = (int8_t)(long long)v; break;
  }
}

// numpy's floor division of floats (npy_divmod): exact for the cases where floor(a / b) would round the wrong way
template <typename F>
__device__ __forceinline__ F np_floordiv(F a, F b) {
  if (b == (F)0) return a / b;
  F mod = fmod(a, b);
  F div = (a - mod) / b;
  if (mod != (F)0 && ((b < (F)0) != (mod < (F)0))) div -= (F)1;
  if (div != (F)0) {
    F fl = floor(div);
    if (div - fl > (F)0.5) fl += (F)1;
    return fl;
  }
  return copysign((F)0, a / b);
}
__device__ __forceinline__ long long np_floordiv(long long a, long long b) {
  if (b == 0) return 0;   // numpy: 0 with a warning
  long long q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) q--;
  return q;
}
template <typename C> __device__ __forceinline__ C c_div(C a, C b) { return a / b; }
template <> __device__ __forceinline__ long long c_div(long long a, long long b) { return b ? a / b : 0; }   // (integer loops: numpy yields 0)
template <typename C> __device__ __forceinline__ C c_sqrt(C a) { return sqrt(a); }
template <> __device__ __forceinline__ long long c_sqrt(long long a) { return (long long)sqrt((double)a); }
template <typename C> __device__ __forceinline__ C c_abs(C a) { return fabs(a); }
template <> __device__ __forceinline__ long long c_abs(long long a) { return a < 0 ? -a : a; }
template <typename C> __device__ __forceinline__ bool c_isnan(C a) { return a != a; }
template <typename C> __device__ __forceinline__ bool c_isfinite(C a) { return isfinite(a); }
template <> __device__ __forceinline__ bool c_isfinite(long long) { return true; }
// numpy maximum / minimum propagate NaN
template <typename C> __device__ __forceinline__ C c_max(C a, C b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
template <typename C> __device__ __forceinline__ C c_min(C a, C b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
template <typename C> __device__ __forceinline__ C c_rnd(int op, C q) {
  return op == GOP_ROUND ? rint(q) : (op == GOP_FLOOR ? floor(q) : (op == GOP_CEIL ? ceil(q) : trunc(q)));
}
template <> __device__ __forceinline__ long long c_rnd(int, long long q) { return q; }

template <typename C>
__global__ void k_glue(int op, long long rows, long long inner, void* out, long long out_rs, int out_dt, Opnd a, Opnd b, Opnd c,
                       double ratio, int mode) {
  const long long total = rows * inner;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / inner, j = e - r * inner;
    const long long o = r * out_rs + j;
    const C x = ld<C>(a, r, j);
    if (op < 16) {
      C y = (C)0, res;
      if (op != GOP_NEG && op != GOP_ABS && op != GOP_SQRT) y = ld<C>(b, r, j);
      switch (op) {
        case GOP_ADD: res = x + y; break;
        case GOP_SUB: res = x - y; break;
        case GOP_MUL: res = x * y; break;
        case GOP_DIV: res = c_div<C>(x, y); break;
        case GOP_FLOORDIV: res = np_floordiv(x, y); break;
        case GOP_NEG: res = -x; break;
        case GOP_ABS: res = c_abs<C>(x); break;
        case GOP_SQRT: res = c_sqrt<C>(x); break;
        case GOP_MAX: res = c_max<C>(x, y); break;
        default: res = c_min<C>(x, y); break;
      }
      st<C>(out, out_dt, o, res);
    } else if (op < 32) {
      bool res;
      if (op == GOP_ISNAN) res = c_isnan<C>(x);
      else if (op == GOP_ISFINITE) res = c_isfinite<C>(x);
      else {
        const C y = ld<C>(b, r, j);
        res = op == GOP_EQ ? x == y : op == GOP_NE ? x != y : op == GOP_LT ? x < y : op == GOP_LE ? x <= y
              : op == GOP_GT ? x > y : x >= y;
      }
      st<C>(out, out_dt, o, (C)(res ? 1 : 0));
    } else if (op == GOP_WHERE) {     // a: condition (any dtype, non-zero = true), b / c: the two choices
      st<C>(out, out_dt, o, x != (C)0 ? ld<C>(b, r, j) : ld<C>(c, r, j));
    } else if (op == GOP_COPY) {
      st<C>(out, out_dt, o, x);
    } else if (op < 48) {             // to_nearest * f(val / to_nearest); NaN stays NaN
      const C tn = ld<C>(b, r, j);
      C res = tn * c_rnd<C>(op, x / tn);
      if (x != x) res = x;
      st<C>(out, out_dt, o, res);
    } else {                          // unit conversion, float64: (buf + offset_in) * ratio - offset_out
      const double v = __dsub_rn(__dmul_rn(__dadd_rn((double)x, (double)ld<C>(b, r, j)), ratio), (double)ld<C>(c, r, j));
      const double w = mode == 1 ? rint(v) : mode == 2 ? floor(v) : mode == 3 ? ceil(v) : mode == 4 ? trunc(v) : v;
      st<double>(out, out_dt, o, w);
    }
  }
}

// get.py:10-91 : out[r] = a[r, i[r]] (negative indices wrap); default / NaN handling by the caller's `dflt`
template <typename C>
__global__ void k_get(long long rows, Opnd a, long long n, Opnd idx, Opnd dflt, int use_default, void* out, int out_dt, int* fatal) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  long long i = ld<long long>(idx, r, 0);
  const bool valid = i >= -n && i < n;
  if (i < 0) i += n;
  C v = (C)0;
  if (valid) v = ld<C>(a, r, i);
  if (use_default) {
    if (!valid || v != v) v = ld<C>(dflt, r, 0);
  } else if (!valid) {
    raise_fatal(fatal, 32, r);
  }
  st<C>(out, out_dt, r, v);
}

}  // namespace

// Element-wise glue: out[r, j] = op(a[r, j], b[r, j], c[r, j]) for r < rows, j < inner.
//   loop        0 float32, 1 float64, 2 int64: the type the operands are converted to and the operation runs in
//   operands    (ptr, row stride, column stride, dtype, immediate): ptr NULL = the immediate; strides in elements
//   out         contiguous along the inner axis, row stride out_rs, dtype out_dt (bool for comparisons / predicates)
extern "C" int dspb_glue(int32_t op, int32_t loop, int64_t rows, int64_t inner, void* out, int64_t out_rs, int32_t out_dt,
                         const void* a, int64_t a_rs, int64_t a_cs, int32_t a_dt, double a_imm,
                         const void* b, int64_t b_rs, int64_t b_cs, int32_t b_dt, double b_imm,
                         const void* c, int64_t c_rs, int64_t c_cs, int32_t c_dt, double c_imm,
                         double ratio, int32_t mode, void* stream) {
  if (rows <= 0 || inner <= 0) return 0;
  const Opnd A{a, a_rs, a_cs, a_dt, a_imm}, B{b, b_rs, b_cs, b_dt, b_imm}, Cc{c, c_rs, c_cs, c_dt, c_imm};
  const long long total = rows * inner;
  const int threads = 256;
  const int grid = (int)((total + threads - 1) / threads < 148LL * 32 ? (total + threads - 1) / threads : 148LL * 32);
  cudaStream_t s = (cudaStream_t)stream;
  if (loop == 0) k_glue<float><<<grid, threads, 0, s>>>(op, rows, inner, out, out_rs, out_dt, A, B, Cc, ratio, mode);
  else if (loop == 1) k_glue<double><<<grid, threads, 0, s>>>(op, rows, inner, out, out_rs, out_dt, A, B, Cc, ratio, mode);
  else if (loop == 2) k_glue<long long><<<grid, threads, 0, s>>>(op, rows, inner, out, out_rs, out_dt, A, B, Cc, ratio, mode);
  else return DSPB_ERR_UNSUPPORTED;
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

extern "C" int dspb_glue_get(int32_t loop, int64_t rows, const void* a, int64_t a_rs, int32_t a_dt, int64_t n,
                             const void* idx, int64_t idx_rs, int32_t idx_dt, double idx_imm,
                             const void* dflt, int64_t dflt_rs, int32_t dflt_dt, double dflt_imm, int32_t use_default,
                             void* out, int32_t out_dt, int32_t* fatal, void* stream) {
  if (rows <= 0) return 0;
  const Opnd A{a, a_rs, 1, a_dt, 0.0}, I{idx, idx_rs, 0, idx_dt, idx_imm}, D{dflt, dflt_rs, 0, dflt_dt, dflt_imm};
  const int grid = (int)((rows + 127) / 128);
  cudaStream_t s = (cudaStream_t)stream;
  if (loop == 0) k_get<float><<<grid, 128, 0, s>>>(rows, A, n, I, D, use_default, out, out_dt, fatal);
  else if (loop == 1) k_get<double><<<grid, 128, 0, s>>>(rows, A, n, I, D, use_default, out, out_dt, fatal);
  else if (loop == 2) k_get<long long><<<grid, 128, 0, s>>>(rows, A, n, I, D, use_default, out, out_dt, fatal);
  else return DSPB_ERR_UNSUPPORTED;
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
