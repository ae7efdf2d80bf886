// dspeed_b200 -- the fused, waveform-resident chain kernel.
//
// The host chain compiler (dspeed_b200/fusion.py) lowers a compiled ProcessingChain
// (reference: the per-block processor loop of processing_chain.py:1144-1163) into a flat
// program.  One persistent CTA per SM interprets it for one waveform at a time: the raw
// uint16 row is read from HBM ONCE (128-bit loads), every intermediate waveform lives in a
// shared-memory slot, per-event scalars live in a small shared scalar file, and only the
// requested outputs are written back.  All threads execute the same instruction stream
// (the program is uniform), so there is no divergence between warps; block routines are
// the ones of row_ops.cuh, i.e. exactly the arithmetic of the per-processor kernels.
//
// Convolutions have three lowerings, chosen by the compiler from the kernel array itself:
//   CONV_RUNS  kernels whose first difference is sparse (piecewise-constant, e.g. the t0
//              ramp+flat kernel): a sparse FIR followed by a float64 cumulative sum -- exact;
//   CONV_SEG   cusp / zac kernels (sinh ramps + flat top [+ parabolas], differenced by
//              [1,-c]): exponentially / polynomially weighted prefix sums, O(L) instead of
//              O(L*K); used only after the compiler verified the analytic model against the
//              actual kernel array;
//   CONV_DIRECT register-tiled direct convolution for everything else.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "conv_ops.cuh"
#include "conv_seg.cuh"
#include "row_ops.cuh"

using namespace dspb;

namespace {

constexpr size_t MAX_SMEM = 227 * 1024;
constexpr int FUSED_THREADS = 512;
constexpr int MAX_PTRS = 64;
constexpr int MAX_SREG = 128;
constexpr int MAX_SLOTS = 8;
constexpr int IARGS = 15;
constexpr int MAX_PROG_SMEM = 160;    // instructions staged in shared memory (64 B each)
constexpr int MAX_CONST_SMEM = 384;   // float64 constants staged in shared memory

enum Op {
  OP_END = 0,
  OP_LOAD_WAVE = 1,
  OP_LOAD_SCALAR = 2,
  OP_STORE_SCALAR = 3,
  OP_STORE_WAVE = 4,
  OP_BL_SUB = 5,
  OP_MIN_MAX = 6,
  OP_LSF = 7,
  OP_POLE_ZERO = 8,
  OP_DPZ = 9,
  OP_TRAP = 10,
  OP_ASYM = 11,
  OP_TRAP_PICKOFF = 12,
  OP_MW = 13,
  OP_AVG_CURRENT = 14,
  OP_TPT = 15,
  OP_ITPT = 16,
  OP_FTP = 17,
  OP_WINDOWER = 18,
  OP_UPSAMPLER = 19,
  OP_CONV_DIRECT = 20,
  OP_CONV_RUNS = 21,
  OP_CONV_SEG = 22,
  OP_SC_BIN = 23,
  OP_SC_CONVERT = 24,
  OP_SC_UNARY = 25,
  OP_MIN_MAX_NORM = 26,
  OP_LSD = 27,
  OP_MBT = 28,
  OP_PREFIX = 29,
  OP_PFIR = 30
};

struct Instr {
  int op;
  int a[IARGS];
};

struct PtrTable {
  const void* p[MAX_PTRS];
  long long s[MAX_PTRS];  // row strides (elements); 0 = broadcast
};

struct ChainParams {
  const Instr* prog;
  int n_instr;
  const double* consts;
  int n_consts;
  int slot_words;  // floats per slot
  int n_slots;
  int* fatal;      // int32[n][4] table (may be null)
  long long row0;  // global index of the first row of this launch (for fatal records)
  long long* prof; // optional [n_instr] cycle counters (CTA 0 only; profiling builds of a chain)
};

// slots are padded to a multiple of 4 words so that float64 scratch carved out of a slot
// (CONV_SEG tables, CONV_DIRECT partial sums) is 16-byte aligned
inline int slot_words_aligned(int n) { return (slot_words(n) + 3) & ~3; }

__device__ __forceinline__ double load_scalar_as_double(const void* p, long long idx, int dtype) {
  switch (dtype) {
    case DSPB_F32: return (double)reinterpret_cast<const float*>(p)[idx];
    case DSPB_F64: return reinterpret_cast<const double*>(p)[idx];
    case DSPB_U16: return (double)reinterpret_cast<const uint16_t*>(p)[idx];
    case DSPB_I16: return (double)reinterpret_cast<const int16_t*>(p)[idx];
    case DSPB_I32: return (double)reinterpret_cast<const int32_t*>(p)[idx];
    case DSPB_U32: return (double)reinterpret_cast<const uint32_t*>(p)[idx];
    case 6: return (double)reinterpret_cast<const long long*>(p)[idx];
  }
  return 0.0;
}

// out[i] = sum_s c_s * P[i + shift - t_s] with the (few) taps held in registers
template <typename T, int MAXT>
__device__ __forceinline__ int pfir_run(const double* __restrict__ Pq, const double* taps, int nt, int n, int shift,
                                        int p, T* out) {
  int tt[MAXT];
  double cc[MAXT];
#pragma unroll
  for (int s = 0; s < MAXT; s++) {
    tt[s] = s < nt ? shift - (int)taps[2 * s] : 0;
    cc[s] = s < nt ? taps[2 * s + 1] : 0.0;
  }
  int bad = 0;
  for (int i = threadIdx.x; i < p; i += NT) {
    double v = 0.0;
#pragma unroll
    for (int s = 0; s < MAXT; s++) {
      if (MAXT > 12 && s >= nt) break;
      const int idx = min(i + tt[s], n - 1);
      if (idx >= 0) v = fma(cc[s], Pq[sidx(idx)], v);
    }
    const T o = (T)v;
    bad |= (o != o);
    out[sidx(i)] = o;
  }
  return bad;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
k_chain(const ChainParams cp, const __grid_constant__ PtrTable pt, const long long n_rows) {
  using T = float;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Scratch* sc = reinterpret_cast<Scratch*>(smem_raw);
  double* sreg = reinterpret_cast<double*>(smem_raw + SCRATCH_BYTES);
  int* slot_nan = reinterpret_cast<int*>(sreg + MAX_SREG);
  Aff2* aff = reinterpret_cast<Aff2*>(slot_nan + 16);
  double* Ks = reinterpret_cast<double*>(aff + MAXW);
  int4* prog_s = reinterpret_cast<int4*>(Ks + MAX_CONST_SMEM);
  T* slots = reinterpret_cast<T*>(prog_s + 4 * MAX_PROG_SMEM);
  const int SW = cp.slot_words;
  // stage the (small, read-only) program and constant pool once per CTA
  const bool prog_in_smem = cp.n_instr <= MAX_PROG_SMEM;
  const bool k_in_smem = cp.n_consts <= MAX_CONST_SMEM;
  if (prog_in_smem)
    for (int i = threadIdx.x; i < cp.n_instr * 4; i += NT) prog_s[i] = reinterpret_cast<const int4*>(cp.prog)[i];
  if (k_in_smem)
    for (int i = threadIdx.x; i < cp.n_consts; i += NT) Ks[i] = cp.consts[i];
  __syncthreads();
  const double* K = k_in_smem ? Ks : cp.consts;
  const int4* prog = prog_in_smem ? prog_s : reinterpret_cast<const int4*>(cp.prog);

  auto slot = [&](int s) -> T* { return slots + (size_t)s * SW; };
  auto sget = [&](int kind, int idx) -> double { return kind == 0 ? sreg[idx] : K[idx]; };
  auto fatal_at = [&](int fi) -> int* { return cp.fatal ? cp.fatal + 4 * fi : nullptr; };

  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    if (threadIdx.x < 16) slot_nan[threadIdx.x] = 0;
    __syncthreads();
    long long t_prev = cp.prof ? clock64() : 0;
    for (int pc = 0; pc < cp.n_instr; pc++) {
      if (cp.prof && pc > 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        const long long t_now = clock64();
        cp.prof[pc - 1] += t_now - t_prev;
        t_prev = t_now;
      }
      // one instruction = 4 x 128-bit loads into registers (no re-reads inside the op)
      const int4 w0 = prog[4 * pc], w1 = prog[4 * pc + 1], w2 = prog[4 * pc + 2], w3 = prog[4 * pc + 3];
      const int op = w0.x;
      const int a[IARGS] = {w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
      switch (op) {
        case OP_LOAD_WAVE: {  // a0 slot, a1 ptr, a2 n, a3 dtype
          Wave w;
          w.ptr = pt.p[a[1]]; w.row_stride = pt.s[a[1]]; w.dtype = a[3];
          const int f = block_or(stage_row<T>(slot(a[0]), w, row, a[2]));
          if (threadIdx.x == 0) slot_nan[a[0]] = f;
          __syncthreads();
          break;
        }
        case OP_LOAD_SCALAR: {  // a0 reg, a1 ptr, a2 dtype, a3 round_to_f32
          double v = load_scalar_as_double(pt.p[a[1]], row * pt.s[a[1]], a[2]);
          if (a[3]) v = (double)(float)v;
          sreg[a[0]] = v;
          break;
        }
        case OP_STORE_SCALAR: {  // a0 reg, a1 ptr, a2 dtype
          if (threadIdx.x == 0) {
            const double v = sreg[a[0]];
            void* p = const_cast<void*>(pt.p[a[1]]);
            switch (a[2]) {
              case DSPB_F32: reinterpret_cast<float*>(p)[row] = (float)v; break;
              case DSPB_F64: reinterpret_cast<double*>(p)[row] = v; break;
              case DSPB_I32: reinterpret_cast<int*>(p)[row] = (int)v; break;
              case DSPB_U32: reinterpret_cast<unsigned*>(p)[row] = (unsigned)v; break;
            }
          }
          break;
        }
        case OP_STORE_WAVE: {  // a0 slot, a1 off, a2 n, a3 ptr
          T* g = reinterpret_cast<T*>(const_cast<void*>(pt.p[a[3]])) + row * pt.s[a[3]];
          const T* s = slot(a[0]);
          if (slot_nan[a[0]] != 0) store_row_nan<T>(g, a[2]);
          else for (int i = threadIdx.x; i < a[2]; i += NT) g[i] = s[sidx(a[1] + i)];
          break;
        }
        case OP_BL_SUB: {  // a0 in, a1 off, a2 n, a3 out, a4/a5 scalar
          const T b = (T)sget(a[4], a[5]);
          const int nanf = slot_nan[a[0]] != 0 || b != b;
          if (nanf) fill_slot_nan<T>(slot(a[3]), a[2]);
          else {
            const T* in = slot(a[0]);
            T* out = slot(a[3]);
            for (int i = threadIdx.x; i < a[2]; i += NT) out[sidx(i)] = in[sidx(a[1] + i)] - b;
          }
          __syncthreads();
          if (threadIdx.x == 0) slot_nan[a[3]] = nanf;
          __syncthreads();
          break;
        }
        case OP_MIN_MAX: {  // a0 in, a1 off, a2 n, a3..a6 regs (tmin,tmax,vmin,vmax; -1 unused)
          const T* in = slot(a[0]);
          double r[4] = {CUDART_NAN, CUDART_NAN, CUDART_NAN, CUDART_NAN};
          if (slot_nan[a[0]] == 0) {
            const int off = a[1], n = a[2];
            T mn = in[sidx(off)], mx = mn;
            int ia = 0, ib = 0;
            for (int i = threadIdx.x; i < n; i += NT) {
              const T v = in[sidx(off + i)];
              if (v < mn) { mn = v; ia = i; }
              if (v > mx) { mx = v; ib = i; }
            }
            block_argminmax<T>(mn, ia, mx, ib, sc);
            r[0] = (double)ia; r[1] = (double)ib; r[2] = (double)mn; r[3] = (double)mx;
          }
#pragma unroll
          for (int k = 0; k < 4; k++) if (a[3 + k] >= 0) sreg[a[3 + k]] = r[k];
          break;
        }
        case OP_LSF: {  // a0 in, a1 off, a2 n, a3..a6 regs
          double r[4] = {CUDART_NAN, CUDART_NAN, CUDART_NAN, CUDART_NAN};
          if (slot_nan[a[0]] == 0) {
            const T* in = slot(a[0]);
            const int off = a[1], n = a[2];
            double sy = 0.0, sxy = 0.0;
            for (int i = threadIdx.x; i < n; i += NT) {
              const double v = (double)in[sidx(off + i)];
              sy += v; sxy += v * (double)i;
            }
            block_sum2(sy, sxy, sc);
            const double mean = sy / (double)n;
            double m2 = 0.0;
            for (int i = threadIdx.x; i < n; i += NT) {
              const double dv = (double)in[sidx(off + i)] - mean;
              m2 += dv * dv;
            }
            m2 = block_sum(m2, sc);
            const long long nn = n, sx = nn * (nn - 1) / 2, sx2 = (nn - 1) * nn * (2 * nn - 1) / 6;
            const float slope = (float)(((double)nn * sxy - (double)sx * sy) / (double)(nn * sx2 - sx * sx));
            r[0] = (double)(float)mean;
            r[1] = (double)(float)sqrt(m2 / (double)(n - 1));
            r[2] = (double)slope;
            r[3] = (double)(float)((sy - (double)sx * (double)slope) / (double)nn);
          }
#pragma unroll
          for (int k = 0; k < 4; k++) if (a[3 + k] >= 0) sreg[a[3 + k]] = r[k];
          break;
        }
        case OP_POLE_ZERO: {  // a0 in, a2 n, a3 out, a4/a5 tau, a14 fatal idx
          const T tau = (T)sget(a[4], a[5]);
          int nanf = slot_nan[a[0]] != 0 || tau != tau;
          if (nanf) { fill_slot_nan<T>(slot(a[3]), a[2]); __syncthreads(); }
          else {
            const int bad = block_or(op_pole_zero<T>(slot(a[0]), slot(a[3]), a[2], tau, sc));
            if (bad) { nanf = 2; if (threadIdx.x == 0) raise_fatal(fatal_at(a[14]), DSPB_FATAL_PZ_NAN, cp.row0 + row); }
          }
          if (threadIdx.x == 0) slot_nan[a[3]] = nanf;
          __syncthreads();
          break;
        }
        case OP_DPZ: {  // a0 in, a2 n, a3 out, a4..a9 three scalars
          const T t1 = (T)sget(a[4], a[5]), t2 = (T)sget(a[6], a[7]), fr = (T)sget(a[8], a[9]);
          int nanf = slot_nan[a[0]] != 0 || t1 != t1 || t2 != t2 || fr != fr;
          if (nanf) { fill_slot_nan<T>(slot(a[3]), a[2]); __syncthreads(); }
          else nanf = block_or(op_double_pole_zero<T>(slot(a[0]), slot(a[3]), a[2], t1, t2, fr, sc, aff)) ? 2 : 0;
          if (threadIdx.x == 0) slot_nan[a[3]] = nanf;
          __syncthreads();
          break;
        }
        case OP_TRAP: {  // a0 in, a2 n, a3 out, a4 rise, a5 flat, a6 norm
          int nanf = slot_nan[a[0]] != 0;
          if (nanf) { fill_slot_nan<T>(slot(a[3]), a[2]); __syncthreads(); }
          else nanf = block_or(op_trap<T>(slot(a[0]), slot(a[3]), a[2], a[4], a[5], a[6] != 0, sc)) ? 2 : 0;
          if (threadIdx.x == 0) slot_nan[a[3]] = nanf;
          __syncthreads();
          break;
        }
        case OP_ASYM: {  // a0 in, a2 n, a3 out, a4 rise, a5 flat, a6 fall
          int nanf = slot_nan[a[0]] != 0;
          if (nanf) { fill_slot_nan<T>(slot(a[3]), a[2]); __syncthreads(); }
          else nanf = block_or(op_asym_trap<T>(slot(a[0]), slot(a[3]), a[2], a[4], a[5], a[6], sc)) ? 2 : 0;
          if (threadIdx.x == 0) slot_nan[a[3]] = nanf;
          __syncthreads();
          break;
        }
        case OP_TRAP_PICKOFF: {  // a0 in, a2 n, a3 rise, a4 flat, a5/a6 t, a7 out reg, a14 fatal
          const T t = (T)sget(a[5], a[6]);
          double r = CUDART_NAN;
          int f = 0;
          if (slot_nan[a[0]] == 0 && t == t) r = (double)op_trap_pickoff<T>(slot(a[0]), a[2], a[3], a[4], t, f, sc);
          if (f && threadIdx.x == 0) raise_fatal(fatal_at(a[14]), f, cp.row0 + row);
          sreg[a[7]] = r;
          break;
        }
        case OP_MW: {  // a0 in, a2 n, a3 out, a4 tmp, a5 length const idx, a6 kind, a7 num, a8 type
          int nanf = slot_nan[a[0]] != 0;
          const T len = (T)K[a[5]];
          if (nanf) { fill_slot_nan<T>(slot(a[3]), a[2]); __syncthreads(); }
          else if (a[6] == 0) nanf = block_or(op_mw_left<T>(slot(a[0]), slot(a[3]), a[2], len, sc)) ? 2 : 0;
          else if (a[6] == 1) nanf = block_or(op_mw_right<T>(slot(a[0]), slot(a[3]), a[2], len, sc)) ? 2 : 0;
          else nanf = block_or(op_mw_multi<T>(slot(a[0]), slot(a[3]), slot(a[4]), a[2], len, a[7], a[8], sc)) ? 2 : 0;
          if (threadIdx.x == 0) slot_nan[a[3]] = nanf;
          __syncthreads();
          break;
        }
        case OP_AVG_CURRENT: {  // a0 in, a3 out, a4 n_out, a5 length const
          const int nanf = slot_nan[a[0]] != 0;
          if (nanf) { fill_slot_nan<T>(slot(a[3]), a[4]); __syncthreads(); }
          else op_avg_current<T>(slot(a[0]), slot(a[3]), a[4], (T)K[a[5]]);
          if (threadIdx.x == 0) slot_nan[a[3]] = nanf;
          __syncthreads();
          break;
        }
        case OP_TPT: {  // a0 in, a2 n, a3/a4 thr, a5/a6 start, a7/a8 walk, a9 out reg, a14 fatal
          double r = CUDART_NAN;
          int f = 0;
          if (slot_nan[a[0]] == 0)
            r = (double)op_time_point_thresh<T>(slot(a[0]), a[2], (T)sget(a[3], a[4]), (T)sget(a[5], a[6]),
                                                (T)sget(a[7], a[8]), f, sc);
          if (f && threadIdx.x == 0) raise_fatal(fatal_at(a[14]), f, cp.row0 + row);
          sreg[a[9]] = r;
          break;
        }
        case OP_ITPT: {  // a0 in, a2 n, a3/a4 thr, a5/a6 start, a7 walk, a8 mode, a9 out reg
          double r = CUDART_NAN;
          int f = 0;
          if (slot_nan[a[0]] == 0)
            r = (double)op_interp_time_point_thresh<T>(slot(a[0]), a[2], (T)sget(a[3], a[4]), (T)sget(a[5], a[6]),
                                                       (long long)a[7], a[8], f, sc);
          if (f && threadIdx.x == 0) raise_fatal(fatal_at(a[14]), f, cp.row0 + row);
          sreg[a[9]] = r;
          break;
        }
        case OP_FTP: {  // a0 in, a2 n, a3/a4 t, a5 mode, a6 out reg
          double r = CUDART_NAN;
          int f = 0;
          if (slot_nan[a[0]] == 0) {
            if (a[5] == 's') {  // spline: heavy local state, one thread only
              if (threadIdx.x == 0) {
                sreg[a[6]] = (double)op_fixed_time_pickoff<T>(slot(a[0]), a[2], (T)sget(a[3], a[4]), a[5], f);
                if (f) raise_fatal(fatal_at(a[14]), f, cp.row0 + row);
              }
              __syncthreads();
              break;
            }
            r = (double)op_fixed_time_pickoff<T>(slot(a[0]), a[2], (T)sget(a[3], a[4]), a[5], f);
          }
          if (f && threadIdx.x == 0) raise_fatal(fatal_at(a[14]), f, cp.row0 + row);
          sreg[a[6]] = r;
          break;
        }
        case OP_WINDOWER: {  // a0 in, a2 n, a3 out, a4 m, a5/a6 t0
          const T t0 = (T)sget(a[5], a[6]);
          int nanf = slot_nan[a[0]] != 0 || t0 != t0;
          if (nanf) { fill_slot_nan<T>(slot(a[3]), a[4]); __syncthreads(); }
          else nanf = block_or(op_windower<T>(slot(a[0]), slot(a[3]), a[2], a[4], t0)) ? 1 : 0;
          if (threadIdx.x == 0) slot_nan[a[3]] = nanf;
          __syncthreads();
          break;
        }
        case OP_UPSAMPLER: {  // a0 in, a2 n, a3 out, a4 m, a5 up const
          int nanf = slot_nan[a[0]] != 0;
          if (nanf) { fill_slot_nan<T>(slot(a[3]), a[4]); __syncthreads(); }
          else nanf = block_or(op_upsampler<T>(slot(a[0]), slot(a[3]), a[2], a[4], (T)K[a[5]])) ? 1 : 0;
          if (threadIdx.x == 0) slot_nan[a[3]] = nanf;
          __syncthreads();
          break;
        }
        case OP_MIN_MAX_NORM: {  // a0 in, a2 n, a3 out, a4/a5 a_min, a6/a7 a_max
          const int nanf = slot_nan[a[0]] != 0;
          if (nanf) { fill_slot_nan<T>(slot(a[3]), a[2]); __syncthreads(); }
          else op_min_max_norm<T>(slot(a[0]), slot(a[3]), a[2], (T)sget(a[4], a[5]), (T)sget(a[6], a[7]));
          if (threadIdx.x == 0) slot_nan[a[3]] = nanf;
          __syncthreads();
          break;
        }
        case OP_CONV_DIRECT: {  // a0 in, a1 off, a2 n, a3 out, a4 p, a5 kernel ptr, a6 m, a7 conv off
          // kernel NaNs are checked on the host when the program is built (constants)
          const int nanf = slot_nan[a[0]] != 0;
          T* out = slot(a[3]);
          if (nanf) { fill_slot_nan<T>(out, a[4]); __syncthreads(); }
          else {
            const T* kv = reinterpret_cast<const T*>(pt.p[a[5]]);
            const T* in = slot(a[0]);
            const int n = a[2], m = a[6], p = a[4], coff = a[7], in_off = a[1];
            const int G = (p + R - 1) / R;
            // split the taps over threads when there are fewer output groups than threads
            const int S = a[9] > 0 ? a[9] : 1;  // tap segments (chosen by the host compiler)
            if (S == 1) {
              for (int g = threadIdx.x; g < G; g += NT) {
                double acc[R];
#pragma unroll
                for (int r = 0; r < R; r++) acc[r] = 0.0;
                const int kk0 = g * R + coff;
                const int t_lo = max(0, kk0 - (n - 1)), t_hi = min(m, kk0 + R);
                if (t_lo < t_hi) conv_group_off<T>(in, in_off, n, kv, kk0, t_lo, t_hi, acc);
#pragma unroll
                for (int r = 0; r < R; r++) if (g * R + r < p) out[sidx(g * R + r)] = (T)acc[r];
              }
              __syncthreads();
            } else {
              // partial sums (float64) go through the scratch slot a8
              double* part = reinterpret_cast<double*>(slot(a[8]));
              int Ls = (m + S - 1) / S; Ls = (Ls + 7) & ~7;
              const int g = threadIdx.x % G, seg = threadIdx.x / G;
              if (seg < S) {
                double acc[R];
#pragma unroll
                for (int r = 0; r < R; r++) acc[r] = 0.0;
                const int kk0 = g * R + coff;
                const int t_lo = max(seg * Ls, max(0, kk0 - (n - 1)));
                const int t_hi = min(min(m, (seg + 1) * Ls), kk0 + R);
                if (t_lo < t_hi) conv_group_off<T>(in, in_off, n, kv, kk0, t_lo, t_hi, acc);
#pragma unroll
                for (int r = 0; r < R; r++) part[(seg * G + g) * R + r] = acc[r];
              }
              __syncthreads();
              for (int k = threadIdx.x; k < p; k += NT) {
                double s = 0.0;
                for (int sg = 0; sg < S; sg++) s += part[sg * G * R + k];
                out[sidx(k)] = (T)s;
              }
              __syncthreads();
            }
          }
          if (threadIdx.x == 0) slot_nan[a[3]] = nanf;
          __syncthreads();
          break;
        }
        case OP_CONV_RUNS: {  // a0 in, a1 off, a2 n, a3 out, a4 p, a5 consts idx (pairs t,c), a6 n_taps, a7 conv off
          const int nanf = slot_nan[a[0]] != 0;
          if (nanf) { fill_slot_nan<T>(slot(a[3]), a[4]); __syncthreads(); }
          else {
            const T* in = slot(a[0]);
            const int n = a[2], in_off = a[1], nt = a[6];
            const double* taps = K + a[5];
            auto d = [&](int kk) -> double {
              double v = 0.0;
              for (int s = 0; s < nt; s++) {
                const int idx = kk - (int)taps[2 * s];
                if (idx >= 0 && idx < n) v += taps[2 * s + 1] * (double)in[sidx(in_off + idx)];
              }
              return v;
            };
            cumsum_fwd<T>(d, [](double v) { return (T)v; }, slot(a[3]), a[7] + a[4], sc, a[7]);
          }
          if (threadIdx.x == 0) slot_nan[a[3]] = nanf;
          __syncthreads();
          break;
        }
        case OP_CONV_SEG: {  // a0 in (off must be 0), a2 n, a3 out, a5 consts idx, a8 scratch slot
          const int nanf = slot_nan[a[0]] != 0;
          const int p = a[2] - (int)K[a[5] + 3] + 1;
          if (nanf) { fill_slot_nan<T>(slot(a[3]), p); __syncthreads(); }
          else op_conv_seg<T>(slot(a[0]), a[2], slot(a[3]), reinterpret_cast<double*>(slot(a[8])), K + a[5], sc);
          if (threadIdx.x == 0) slot_nan[a[3]] = nanf;
          __syncthreads();
          break;
        }
        case OP_PREFIX: {  // a0 in, a2 n, a3 first of two slots receiving the float64 inclusive prefix sums
          // P[k] = sum_{j<=k} x[j] (float64, layout pidx): the common operand of every
          // windowed / recursive filter of this waveform (trapezoids, run-structured kernels)
          if (slot_nan[a[0]] == 0) {
            const T* in = slot(a[0]);
            double* Pq = reinterpret_cast<double*>(slot(a[3]));
            const int n = a[2];
            int lo, hi;
            chunk_range(n, lo, hi);
            double loc = 0.0;
            for (int i = lo; i < hi; i++) loc += (double)in[sidx(i)];
            double tot;
            double run = block_excl_scan(loc, tot, sc);
            for (int i = lo; i < hi; i++) {
              run += (double)in[sidx(i)];
              Pq[sidx(i)] = run;
            }
          }
          __syncthreads();
          break;
        }
        case OP_PFIR: {  // a0 src wave (NaN flag), a2 n, a3 out, a4 p, a5 consts idx (pairs t,c), a6 n_taps,
                         // a7 shift, a8 prefix slots
          // out[i] = sum_s c_s * P[i + shift - t_s]   (P[k<0] = 0, P[k>=n] = P[n-1])
          const int nanf = slot_nan[a[0]] != 0;
          T* out = slot(a[3]);
          if (nanf) fill_slot_nan<T>(out, a[4]);
          else {
            const double* Pq = reinterpret_cast<const double*>(slot(a[8]));
            const double* taps = K + a[5];
            const int n = a[2], nt = a[6], shift = a[7], p = a[4];
            int bad = 0;
            if (nt <= 4) bad = pfir_run<T, 4>(Pq, taps, nt, n, shift, p, out);
            else if (nt <= 12) bad = pfir_run<T, 12>(Pq, taps, nt, n, shift, p, out);
            else bad = pfir_run<T, 32>(Pq, taps, nt, n, shift, p, out);
            (void)bad;
          }
          __syncthreads();
          if (threadIdx.x == 0) slot_nan[a[3]] = nanf;
          __syncthreads();
          break;
        }
        case OP_SC_BIN: {  // a0 op, a1/a2 A, a3/a4 B, a5 out, a6 dtype (0 f32, 1 f64)
          const double x = sget(a[1], a[2]), y = sget(a[3], a[4]);
          double r;
          if (a[6] == 0) {
            const float xf = (float)x, yf = (float)y;
            float rf;
            switch (a[0]) {
              case 0: rf = xf + yf; break;
              case 1: rf = xf - yf; break;
              case 2: rf = xf * yf; break;
              case 3: rf = xf / yf; break;
              default: rf = floorf(xf / yf); break;
            }
            r = (double)rf;
          } else {
            switch (a[0]) {
              case 0: r = x + y; break;
              case 1: r = x - y; break;
              case 2: r = x * y; break;
              case 3: r = x / y; break;
              default: r = floor(x / y); break;
            }
          }
          sreg[a[5]] = r;
          break;
        }
        case OP_SC_UNARY: {  // a0 op (0 neg), a1/a2 A, a5 out
          sreg[a[5]] = -sget(a[1], a[2]);
          break;
        }
        case OP_SC_CONVERT: {  // a0 in reg, a1/a2 off_in, a3/a4 off_out, a5 ratio const, a6 out, a7 mode, a8 f32 out
          double v = (sreg[a[0]] + sget(a[1], a[2])) * K[a[5]] - sget(a[3], a[4]);
          switch (a[7]) {
            case 1: v = rint(v); break;
            case 2: v = floor(v); break;
            case 3: v = ceil(v); break;
            case 4: v = trunc(v); break;
          }
          if (a[8]) v = (double)(float)v;
          sreg[a[6]] = v;
          break;
        }
        case OP_LSD: {  // a0 in, a1 off, a2 n, a3/a4 slope, a5/a6 icpt, a7 mean reg, a8 rms reg
          const T sl = (T)sget(a[3], a[4]), ic = (T)sget(a[5], a[6]);
          double m = CUDART_NAN, r = CUDART_NAN;
          if (slot_nan[a[0]] == 0 && sl == sl && ic == ic) {
            const T* in = slot(a[0]);
            double sm = 0.0, sq = 0.0;
            for (int i = threadIdx.x; i < a[2]; i += NT) {
              const double t = (double)in[sidx(a[1] + i)] - ((double)sl * (double)i + (double)ic);
              sm += t / (double)(i + 1);
              sq += t * t;
            }
            block_sum2(sm, sq, sc);
            m = (double)(float)sm;
            r = (double)(float)sqrt(sq / (double)(a[2] - 1));
          }
          sreg[a[7]] = m;
          sreg[a[8]] = r;
          break;
        }
        case OP_MBT: {  // a0 in, a1 off, a2 n, a3/a4 thr, a5 out reg
          const T th = (T)sget(a[3], a[4]);
          double r = CUDART_NAN;
          if (slot_nan[a[0]] == 0 && th == th) {
            const T* in = slot(a[0]);
            double tot = 0.0, cnt = 0.0;
            for (int i = threadIdx.x; i < a[2]; i += NT) {
              const T v = in[sidx(a[1] + i)];
              if (v < th) { tot += (double)v; cnt += 1.0; }
            }
            block_sum2(tot, cnt, sc);
            if (cnt != 0.0) r = (double)(float)(tot / cnt);
          }
          sreg[a[5]] = r;
          break;
        }
        default: break;
      }
    }
    if (cp.prof && cp.n_instr > 0 && blockIdx.x == 0 && threadIdx.x == 0) cp.prof[cp.n_instr - 1] += clock64() - t_prev;
    __syncthreads();
  }
}

}  // namespace

struct dspb_chain {
  Instr* d_prog = nullptr;
  double* d_consts = nullptr;
  int n_instr = 0;
  int n_consts = 0;
  int n_slots = 0;
  int slot_len = 0;
  size_t smem = 0;
  int num_sms = 148;
  int threads = FUSED_THREADS;
  long long* d_prof = nullptr;  // per-instruction cycle counters (dspb_chain_profile)
};

extern "C" int dspb_chain_create(const int32_t* code, int64_t n_code, const double* consts, int64_t n_consts,
                                 dspb_chain** out) {
  // code = [n_slots, max_slot_len, n_instr, then n_instr * (1 + IARGS) ints]
  if (n_code < 3) return DSPB_ERR_UNSUPPORTED;
  dspb_chain* c = new dspb_chain();
  c->n_slots = code[0];
  c->slot_len = code[1];
  c->n_instr = code[2];
  if ((int64_t)c->n_instr * (1 + IARGS) + 3 != n_code || c->n_slots > MAX_SLOTS) { delete c; return DSPB_ERR_UNSUPPORTED; }
  c->smem = SCRATCH_BYTES + MAX_SREG * sizeof(double) + 16 * sizeof(int) + MAXW * sizeof(Aff2) +
            MAX_CONST_SMEM * sizeof(double) + MAX_PROG_SMEM * sizeof(Instr) +
            (size_t)c->n_slots * slot_words_aligned(c->slot_len) * sizeof(float);
  if (c->smem > MAX_SMEM) { delete c; return DSPB_ERR_ROW_TOO_LONG; }
  std::vector<Instr> prog(c->n_instr);
  for (int i = 0; i < c->n_instr; i++) {
    const int32_t* p = code + 3 + (int64_t)i * (1 + IARGS);
    prog[i].op = p[0];
    for (int k = 0; k < IARGS; k++) prog[i].a[k] = p[1 + k];
  }
  cudaError_t e = cudaMalloc(&c->d_prog, sizeof(Instr) * (size_t)c->n_instr);
  if (e == cudaSuccess) e = cudaMemcpy(c->d_prog, prog.data(), sizeof(Instr) * (size_t)c->n_instr, cudaMemcpyHostToDevice);
  c->n_consts = (int)n_consts;
  const size_t nc = n_consts > 0 ? (size_t)n_consts : 1;
  if (e == cudaSuccess) e = cudaMalloc(&c->d_consts, sizeof(double) * nc);
  if (e == cudaSuccess && n_consts > 0) e = cudaMemcpy(c->d_consts, consts, sizeof(double) * nc, cudaMemcpyHostToDevice);
  if (const char* t = getenv("DSPEED_B200_FUSED_THREADS")) c->threads = atoi(t) >= 1024 ? 1024 : FUSED_THREADS;
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_chain<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_chain<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) { dspb_chain_destroy(c); return -(int)e; }
  *out = c;
  return 0;
}

extern "C" int dspb_chain_launch(dspb_chain* c, const void* const* ptrs, int64_t n_ptrs, int64_t n_rows,
                                 int32_t* fatal, void* stream) {
  // ptrs: n_ptrs device pointers followed by n_ptrs row strides (as int64 reinterpret) and row0
  if (!c || n_ptrs > MAX_PTRS) return DSPB_ERR_UNSUPPORTED;
  if (n_rows <= 0) return 0;
  PtrTable pt;
  memset(&pt, 0, sizeof(pt));
  const long long* strides = reinterpret_cast<const long long*>(ptrs + n_ptrs);
  for (int i = 0; i < n_ptrs; i++) { pt.p[i] = ptrs[i]; pt.s[i] = strides[i]; }
  ChainParams cp;
  cp.prog = c->d_prog;
  cp.n_instr = c->n_instr;
  cp.consts = c->d_consts;
  cp.n_consts = c->n_consts;
  cp.slot_words = slot_words_aligned(c->slot_len);
  cp.n_slots = c->n_slots;
  cp.fatal = fatal;
  cp.row0 = strides[n_ptrs];
  cp.prof = c->d_prof;
  const int grid = (int)(n_rows < c->num_sms ? n_rows : c->num_sms);
  if (c->threads == 1024) k_chain<1024><<<grid, 1024, c->smem, (cudaStream_t)stream>>>(cp, pt, n_rows);
  else k_chain<512><<<grid, 512, c->smem, (cudaStream_t)stream>>>(cp, pt, n_rows);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

// Per-instruction cycle profile of CTA 0 (accumulated over its rows and over launches).
// enable != 0 allocates/zeroes the counters; out (host, n_instr int64) receives them when non-null.
extern "C" int dspb_chain_profile(dspb_chain* c, int enable, int64_t* out) {
  if (!c) return DSPB_ERR_UNSUPPORTED;
  cudaError_t e = cudaSuccess;
  if (out && c->d_prof) {
    e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out, c->d_prof, sizeof(long long) * (size_t)c->n_instr, cudaMemcpyDeviceToHost);
  }
  if (enable) {
    if (!c->d_prof && e == cudaSuccess) e = cudaMalloc(&c->d_prof, sizeof(long long) * (size_t)c->n_instr);
    if (e == cudaSuccess) e = cudaMemset(c->d_prof, 0, sizeof(long long) * (size_t)c->n_instr);
  } else if (c->d_prof) {
    cudaFree(c->d_prof);
    c->d_prof = nullptr;
  }
  return e == cudaSuccess ? 0 : -(int)e;
}

extern "C" int64_t dspb_chain_smem_bytes(const dspb_chain* c) { return c ? (int64_t)c->smem : 0; }

extern "C" void dspb_chain_destroy(dspb_chain* c) {
  if (!c) return;
  if (c->d_prog) cudaFree(c->d_prog);
  if (c->d_consts) cudaFree(c->d_consts);
  if (c->d_prof) cudaFree(c->d_prof);
  delete c;
}
