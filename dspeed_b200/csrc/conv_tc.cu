// dspeed_b200 -- convolution ('valid'; 'full' / 'same' by zero fill) of a block of waveforms with ONE generic long kernel on the
// 5th-generation tensor cores (north_star (3), BASELINE.json config 5; reference: convolve_wf /
// fft_convolve_wf, processors/convolutions.py:14-119).
//
//   y[r, o] = sum_k kern[k] * x[r, o + K - 1 - k]                      (o < P = L - K + 1)
//           = sum_j A[o, j] * x[r, j],   A[o, j] = kern[K - 1 - (j - o)]  (0 <= j - o < K)
//
// i.e. a GEMM  Y^T[P, rows] = A[P, L] * X^T[L, rows]  whose left operand is the banded Toeplitz
// matrix of the kernel.  A tile of A depends only on its distance from the diagonal, so the
// (K + 127) / 32 distinct 128 x 32 tiles are built once per kernel (set-up kernel below) and every
// CTA streams exactly the tiles of its band with TMA; zero tiles are never touched.
//
// Precision: 3xTF32.  Both operands are split into hi = the top 19 bits and lo = x - hi (exact in
// float32); D += A_hi B_hi + A_lo B_hi + A_hi B_lo with float32 accumulation in TMEM.  The dropped
// term and the rounding of the lo parts are O(2^-22) relative.  The waveform tile is split by four
// converter warps in shared memory (in place: hi, second buffer: lo), so x is read from HBM once.
// The tensor core adds into its float32 accumulator with truncation, a bias that grows with the length of
// the accumulation chain times the magnitude of the partial sum (measured: 9e-6 of the output scale
// after 36 k-tiles): the accumulator is therefore drained every FLUSH k-tiles into float32 registers
// (round-to-nearest adds) while the MMAs continue on a second TMEM accumulator.
//
// One CTA = one 128 (outputs) x 128 (waveforms) tile:
//   warp 0   TMA producer   (A_hi, A_lo, X tiles -> 3-stage shared-memory ring, mbarrier complete_tx)
//   warp 1   TMEM allocator + MMA issuer (one elected thread, tcgen05.mma.kind::tf32, M = N = 128, K = 8)
//   warps 2-5 converters (hi / lo split of the X tile) and, at the end, the epilogue
//             (tcgen05.ld 32x32b -> coalesced stores of y[r, o0 + lane])
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "dspeed_b200.h"

namespace {

#ifndef DSPB_TC_BN
#define DSPB_TC_BN 128      // waveforms per CTA tile (MMA N)
#endif
#ifndef DSPB_TC_STAGES
#define DSPB_TC_STAGES 3    // shared-memory ring depth
#endif
#ifndef DSPB_TC_CTAS
#define DSPB_TC_CTAS 1      // CTAs per SM the kernel is sized for
#endif
constexpr int BM = 128, BN = DSPB_TC_BN, BK = 32, STAGES = DSPB_TC_STAGES;
constexpr int FLUSH_DEFAULT = 2;                        // k-tiles per accumulation window
constexpr int TILE_BYTES = BM * BK * 4;                 // 16 KB: a tile of the Toeplitz operand
constexpr int BTILE_BYTES = BN * BK * 4;                // a tile of waveforms
constexpr int STAGE_BYTES = 2 * TILE_BYTES + 2 * BTILE_BYTES;   // A_hi, A_lo, B_hi, B_lo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment */ + 256 /* barriers */ + BN * 4 /* pedestals */;
constexpr uint32_t TMEM_COLS = 2 * BN;                  // two accumulator stages
constexpr int NTHREADS = 192;
constexpr uint32_t SPIN_LIMIT = 1u << 28;               // a protocol bug traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; spin++) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// shared-memory matrix descriptor: K-major operand tile, rows of 128 bytes, 128-byte swizzle,
// 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);     // start address
  d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset
  d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
  return d;
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, N = 128, M = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

struct Params {
  int64_t n_rows, L, K, P;
  float* out;
  int64_t out_stride;
  int nk;   // k tiles of a band: ceil((K + BM - 1) / BK)
  int flush;  // k-tiles per accumulation window
  const float* x;       // waveforms (for the pedestal c[r]) and their row pitch
  int64_t x_stride;
  const float* ksum;    // sum of the kernel taps (device scalar), nullptr: no pedestal removal
  int shift;  // column of x under A's diagonal for output 0: 0 ('valid'), -(K-1) ('full'), -(K-1-(K-1)/2) ('same')
};

__global__ void __launch_bounds__(NTHREADS, DSPB_TC_CTAS)
k_conv_valid_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_x,
                const Params prm) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles: 1024-byte aligned
  unsigned char* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  // barriers: full[s] (TMA landed), conv[s] (hi / lo split done), empty[s] (MMAs of the stage retired), accum
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto conv_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  // accumulator stages: tfull[a] (window complete, MMA -> epilogue), tempty[a] (drained, epilogue -> MMA)
  auto tfull_bar = [&](int a) { return bars + 8u * (3 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (3 * STAGES + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * STAGE_BYTES + 8 * (3 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o0 = blockIdx.x * BM;
  const int r0 = blockIdx.y * BN;
  const int nk = prm.nk;
  // Pedestal removal ('valid' mode): y = k * (x - c) + c sum(k) with c[r] = mean(x[r, 0:64]).  The tensor core truncates when
  // it adds into its accumulator, an error proportional to the partial sums: a large pedestal under a kernel of small
  // area would dominate it.  The subtraction happens in the converter warps, the constant returns in the epilogue.
  float* cvals = reinterpret_cast<float*>(base_ptr + STAGES * STAGE_BYTES + 256);
  if (threadIdx.x >= 64 && threadIdx.x - 64 < BN) {
    const long long r = (long long)r0 + (threadIdx.x - 64);
    float c = 0.f;
    if (prm.ksum && r < prm.n_rows) {   // mean of the first 64 samples (a single sample would inject its own noise)
      const float4* row = reinterpret_cast<const float4*>(prm.x + r * prm.x_stride);
      const int nq = (int)(prm.L < 64 ? prm.L : 64) / 4;
      for (int i = 0; i < nq; i++) {
        const float4 v = row[i];
        c += (v.x + v.y) + (v.z + v.w);
      }
      c = nq > 0 ? c / (float)(4 * nq) : 0.f;
      c = (c == c && fabsf(c) < 3.0e38f) ? c : 0.f;   // NaN / Inf rows are overwritten by k_nan_rows anyway
    }
    cvals[threadIdx.x - 64] = c;
  }
  const int FLUSH = prm.flush;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), 128);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: two stages of 128 float32 accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(const_cast<uint32_t*>(tmem_slot))), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kt = 0; kt < nk; kt++) {
        const int s = kt % STAGES, ph = (kt / STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t st = base + s * STAGE_BYTES;
        mbar_expect_tx(full_bar(s), 2 * TILE_BYTES + BTILE_BYTES);
        tma_load_2d(st, &map_a, full_bar(s), 0, (2 * kt) * BM);                  // A_hi tile kt
        tma_load_2d(st + TILE_BYTES, &map_a, full_bar(s), 0, (2 * kt + 1) * BM); // A_lo tile kt
        tma_load_2d(st + 2 * TILE_BYTES, &map_x, full_bar(s), o0 + kt * BK + prm.shift, r0);  // out-of-range columns read as 0
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      for (int kt = 0; kt < nk; kt++) {
        const int s = kt % STAGES, ph = (kt / STAGES) & 1;
        const int w = kt / FLUSH, a = w & 1;                       // accumulation window and its TMEM stage
        if (kt % FLUSH == 0 && w >= 2) mbar_wait(tempty_bar(a), ((w >> 1) - 1) & 1);   // window w - 2 drained
        mbar_wait(conv_bar(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = base + s * STAGE_BYTES;
        const uint32_t acc = tmem_d + (uint32_t)(a * BN);
#pragma unroll
        for (int k = 0; k < BK / 8; k++) {
          const uint64_t a_hi = umma_desc(st + 32 * k), a_lo = umma_desc(st + TILE_BYTES + 32 * k);
          const uint64_t b_hi = umma_desc(st + 2 * TILE_BYTES + 32 * k), b_lo = umma_desc(st + 2 * TILE_BYTES + BTILE_BYTES + 32 * k);
          umma_tf32(acc, a_hi, b_hi, (kt % FLUSH != 0) || (k != 0));
          umma_tf32(acc, a_lo, b_hi, 1);
          umma_tf32(acc, a_hi, b_lo, 1);
        }
        umma_commit(empty_bar(s));          // the stage may be refilled once these MMAs have read it
        if (kt % FLUSH == FLUSH - 1 || kt == nk - 1) umma_commit(tfull_bar(a));   // window complete
      }
    }
  } else {
    // ===== converters: x tile -> hi (in place) + lo; epilogue: drain the accumulation windows =====
    const int t = threadIdx.x - 64;   // 0 .. 127
    const int q = warp & 3;           // the TMEM lane quarter this warp may read: lane = output o0 + 32 q + lane
    float sum[BN];                    // running sums of this output over the 128 waveforms (registers)
#pragma unroll
    for (int c = 0; c < BN; c++) sum[c] = 0.f;
    auto drain = [&](int w) {
      const int a = w & 1;
      mbar_wait(tfull_bar(a), (w >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(a * BN + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 32; c++) sum[c0 + c] += __uint_as_float(v[c]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tempty_bar(a));
    };
    const int nw = (nk + FLUSH - 1) / FLUSH;
    for (int kt = 0; kt < nk; kt++) {
      const int s = kt % STAGES, ph = (kt / STAGES) & 1;
      mbar_wait(full_bar(s), ph);
      float4* hi = reinterpret_cast<float4*>(base_ptr + s * STAGE_BYTES + 2 * TILE_BYTES);
      float4* lo = reinterpret_cast<float4*>(base_ptr + s * STAGE_BYTES + 2 * TILE_BYTES + BTILE_BYTES);
#pragma unroll
      for (int i4 = 0; i4 < BTILE_BYTES / 16 / 128; i4++) {
        const int i = i4 * 128 + t;     // same byte offset in both buffers: the swizzle is irrelevant here
        float4 v = hi[i];
        const float c = cvals[i >> 3];   // a row of the tile = 128 bytes = 8 float4 (the swizzle stays inside the row)
        v.x -= c; v.y -= c; v.z -= c; v.w -= c;
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
        hi[i] = h;
        lo[i] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
      mbar_arrive(conv_bar(s));
      // the window before the one that just received its last tile has long been multiplied: drain it
      if (kt % FLUSH == FLUSH - 1 && kt / FLUSH >= 1) drain(kt / FLUSH - 1);
    }
    // the windows not drained inside the loop: the last complete one (if any) and the tail
    for (int w = (nk / FLUSH >= 1 ? nk / FLUSH - 1 : 0); w < nw; w++) drain(w);
    const long long o = (long long)o0 + 32 * q + lane;
    if (o < prm.P) {
      const float ks = prm.ksum ? prm.ksum[0] : 0.f;
#pragma unroll
      for (int c = 0; c < BN; c++) {
        const long long r = (long long)r0 + c;
        if (r < prm.n_rows) prm.out[r * prm.out_stride + o] = fmaf(cvals[c], ks, sum[c]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
  }
}

// The distinct tiles of the banded Toeplitz matrix of the kernel: tile kt (distance 32 kt from the
// diagonal), element (oo, jj) = kern[K - 1 - (32 kt + jj - oo)] or 0; hi tiles at 2 kt, lo at 2 kt + 1.
__global__ void k_toeplitz_tiles(const float* __restrict__ kern, int K, int nk, int e, float* __restrict__ tiles) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)nk * BM * BK;
  if (i >= total) return;
  const int jj = (int)(i % BK), oo = (int)((i / BK) % BM), kt = (int)(i / (BK * BM));
  const int d = BK * kt + jj - oo - e;   // e: the columns of x start e samples early (16-byte aligned TMA coordinates)
  const float v = (d >= 0 && d < K) ? kern[K - 1 - d] : 0.f;
  const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  tiles[((long long)(2 * kt) * BM + oo) * BK + jj] = h;
  tiles[((long long)(2 * kt + 1) * BM + oo) * BK + jj] = v - h;
}

// convolutions.py:44-46: a NaN anywhere in the waveform or in the kernel leaves the whole output row NaN (in the
// GEMM a NaN sample only reaches the outputs whose band covers it).  One CTA per waveform, runs after the GEMM.
__global__ void k_nan_rows(const float* __restrict__ x, long long x_stride, int L, const float* __restrict__ kern, int K,
                           float* __restrict__ out, long long out_stride, int P) {
  const float* row = x + (long long)blockIdx.x * x_stride;
  int bad = 0;
  for (int i = threadIdx.x; i < L / 4; i += blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(row)[i];
    bad |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
  }
  for (int i = (L / 4) * 4 + threadIdx.x; i < L; i += blockDim.x) bad |= row[i] != row[i];
  for (int i = threadIdx.x; i < K; i += blockDim.x) bad |= kern[i] != kern[i];
  if (__syncthreads_or(bad)) {
    float* o = out + (long long)blockIdx.x * out_stride;
    for (int i = threadIdx.x; i < P; i += blockDim.x) o[i] = __int_as_float(0x7fc00000);
  }
}

// sum of the kernel taps (float64 accumulation) -> one float behind the Toeplitz tiles
__global__ void k_kernel_sum(const float* __restrict__ kern, int K, float* __restrict__ dst) {
  __shared__ double part[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < K; i += 256) s += (double)kern[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) dst[0] = (float)part[0];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D float32 tensor [rows][cols] (row pitch in elements), box = 32 columns x box_rows rows, 128-byte swizzle
int make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return DSPB_ERR_UNSUPPORTED;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {pitch_elems * sizeof(float)};
  const cuuint32_t box[2] = {BK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : DSPB_ERR_UNSUPPORTED;
}

}  // namespace

// floats of workspace the launcher needs for a kernel of length K (the Toeplitz tiles, hi and lo)
extern "C" int64_t dspb_convolve_tc_workspace(int64_t K) {
  const int64_t nk = (K + 3 + BM - 1 + BK - 1) / BK;
  return 2 * nk * BM * BK + 32;
}

// y[r, 0:P] = convolution (numpy.convolve modes 'f' | 'v' | 's', P = L + K - 1 | L - K + 1 | L) of x[r, 0:L] with
// kern[0:K] for n_rows waveforms; all pointers on the device, x / out row pitches in elements (x: 16-byte aligned
// rows).  The zero padding of 'full' / 'same' is TMA's out-of-bounds fill.  Returns 0, a DSPB_FATAL_* code,
// DSPB_ERR_UNSUPPORTED or -cudaError_t.
extern "C" int dspb_convolve_tc_f32(const float* x, int64_t x_stride, int64_t n_rows, int64_t L, const float* kern,
                                    int64_t K, int32_t mode_in, float* out, int64_t out_stride, int64_t p,
                                    float* workspace, int64_t workspace_floats, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (K < 1 || K > L || n_rows <= 0 || n_rows > 65535LL * BN || L > 0x7fffffff) return DSPB_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (x_stride & 3) || (reinterpret_cast<uintptr_t>(workspace) & 127))
    return DSPB_ERR_UNSUPPORTED;
  int64_t P, shift;
  if (mode_in == 'v') { P = L - K + 1; shift = 0; }
  else if (mode_in == 'f') { P = L + K - 1; shift = -(K - 1); }
  else if (mode_in == 's') { P = L; shift = -((K - 1) - (K - 1) / 2); }
  else return DSPB_FATAL_CONV_MODE;
  if (p != P) return DSPB_FATAL_CONV_OUTLEN;
  // TMA wants the innermost coordinate on a 16-byte boundary: start the columns lead = 0..3 samples early and move the
  // Toeplitz band by the same amount
  const int64_t lead = ((shift % 4) + 4) % 4;
  shift -= lead;
  const int64_t nk = (K + lead + BM - 1 + BK - 1) / BK;
  if (workspace_floats < 2 * nk * BM * BK + 32) return DSPB_ERR_UNSUPPORTED;
  float* ksum = workspace + 2 * nk * BM * BK;
  {
    const long long total = (long long)nk * BM * BK;
    k_toeplitz_tiles<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(kern, (int)K, (int)nk, (int)lead, workspace);
    k_kernel_sum<<<1, 256, 0, stream>>>(kern, (int)K, ksum);
  }
  CUtensorMap map_a, map_x;
  int rc = make_map(&map_a, workspace, (uint64_t)(2 * nk * BM), BK, BK, BM);
  if (rc) return rc;
  rc = make_map(&map_x, x, (uint64_t)n_rows, (uint64_t)L, (uint64_t)x_stride, BN);
  if (rc) return rc;
  int flush = FLUSH_DEFAULT;
  if (const char* e = getenv("DSPEED_B200_TC_FLUSH")) flush = atoi(e) > 0 ? atoi(e) : FLUSH_DEFAULT;
  Params prm{n_rows, L, K, P, out, out_stride, (int)nk, flush, x, x_stride, mode_in == 'v' ? ksum : nullptr, (int)shift};
  cudaError_t e = cudaFuncSetAttribute(k_conv_valid_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return -(int)e;
  dim3 grid((unsigned)((P + BM - 1) / BM), (unsigned)((n_rows + BN - 1) / BN));
  k_conv_valid_tc<<<grid, NTHREADS, SMEM_BYTES, stream>>>(map_a, map_x, prm);
  k_nan_rows<<<(unsigned)n_rows, 256, 0, stream>>>(x, x_stride, (int)L, kern, (int)K, out, out_stride, (int)P);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
