// dspeed_b200 -- shared-memory tiled direct convolution (convolve_wf / fft_convolve_wf,
// reference convolutions.py:14-119).
//
// One CTA per waveform: the row and the kernel are staged in shared memory; a thread
// computes R = 8 consecutive output samples over a segment of the taps with a sliding
// register window (8 FMAs per 2 shared loads), partial sums of tap segments are combined
// in float64 in a fixed order (deterministic).  float32 products are accumulated in
// chunks of 64 taps that are flushed into float64 accumulators, so the error does not
// grow with the kernel length (cusp/zac: 5792 taps).
#include "conv_ops.cuh"

using namespace dspb;

namespace {

constexpr size_t MAX_SMEM = 227 * 1024;
template <typename T>
__global__ void __launch_bounds__(DEFAULT_THREADS) k_convolve(Wave in, const T* __restrict__ kern, ConvPlan<T> pl, T* out,
                                                 long long out_row_stride, long long n_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* a = reinterpret_cast<T*>(smem_raw + SCRATCH_BYTES);
  T* kv = a + slot_words(pl.n);
  double* part = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(kv + ((pl.m + 3) & ~3)) + 7) & ~(uintptr_t)7);
  // stage the kernel once per CTA
  int knan = 0;
  for (int i = threadIdx.x; i < pl.m; i += NT) {
    const T v = kern[i];
    knan |= (v != v);
    kv[i] = v;
  }
  knan = block_or(knan);
  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int nan_in = block_or(stage_row<T>(a, in, row, pl.n));
    T* o = out + row * out_row_stride;
    if (nan_in || knan) {
      store_row_nan<T>(o, pl.p);
      __syncthreads();
      continue;
    }
    if (pl.S == 1) {
      for (int g = threadIdx.x; g < pl.G; g += NT) {
        double acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0.0;
        // taps that can touch the row for this group: t in [kk0 - (n-1), kk0 + 7]
        const int kk0 = g * R + pl.off;
        const int t_lo = max(0, kk0 - (pl.n - 1));
        const int t_hi = min(pl.m, kk0 + R);
        if (t_lo < t_hi) conv_group<T>(a, pl.n, kv, kk0, t_lo, t_hi, acc);
#pragma unroll
        for (int r = 0; r < R; r++)
          if (g * R + r < pl.p) o[g * R + r] = (T)acc[r];
      }
    } else {
      const int g = threadIdx.x % pl.G, seg = threadIdx.x / pl.G;
      if (seg < pl.S) {
        double acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0.0;
        const int kk0 = g * R + pl.off;
        const int t_lo = max(seg * pl.L, max(0, kk0 - (pl.n - 1)));
        const int t_hi = min(min(pl.m, (seg + 1) * pl.L), kk0 + R);
        if (t_lo < t_hi) conv_group<T>(a, pl.n, kv, kk0, t_lo, t_hi, acc);
#pragma unroll
        for (int r = 0; r < R; r++) part[(seg * pl.G + g) * R + r] = acc[r];
      }
      __syncthreads();
      for (int k = threadIdx.x; k < pl.p; k += NT) {
        double s = 0.0;
        for (int sg = 0; sg < pl.S; sg++) s += part[sg * pl.G * R + k];
        o[k] = (T)s;
      }
    }
    __syncthreads();
  }
}

template <typename T>
int launch_convolve(const void* w_in, int64_t rs, int32_t dt, int64_t n_rows, int64_t n, const void* kernel,
                    int64_t m, int32_t mode, void* w_out, int64_t out_rs, int64_t p, void* stream) {
  if (m > n) return DSPB_FATAL_CONV_KERNEL_LONG;
  int64_t expect, off;
  if (mode == 'f') { expect = n + m - 1; off = 0; }
  else if (mode == 'v') { expect = n - m + 1; off = m - 1; }
  else if (mode == 's') { expect = n; off = (m - 1) / 2; }
  else return DSPB_FATAL_CONV_MODE;
  if (p != expect) return DSPB_FATAL_CONV_OUTLEN;
  if (n_rows <= 0) return 0;
  ConvPlan<T> pl;
  pl.n = (int)n; pl.m = (int)m; pl.p = (int)p; pl.off = (int)off;
  pl.G = (int)((p + R - 1) / R);
  if (pl.G >= DEFAULT_THREADS) { pl.S = 1; pl.L = (int)m; }
  else {
    int S = DEFAULT_THREADS / pl.G;
    const int max_s = (int)((m + CH - 1) / CH);
    if (S > max_s) S = max_s;
    if (S < 1) S = 1;
    int L = (int)((m + S - 1) / S);
    L = (L + 7) & ~7;
    pl.S = S; pl.L = L;
  }
  const size_t part_bytes = pl.S > 1 ? sizeof(double) * (size_t)pl.S * pl.G * R : 0;
  const size_t smem = SCRATCH_BYTES + sizeof(T) * ((size_t)slot_words((int)n) + ((m + 3) & ~3)) + 16 + part_bytes;
  if (smem > MAX_SMEM) return DSPB_ERR_ROW_TOO_LONG;
  auto kern = k_convolve<T>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(int)e;
  const long long max_grid = 148LL * 8;
  const int grid = (int)(n_rows < max_grid ? n_rows : max_grid);
  Wave in;
  in.ptr = w_in; in.row_stride = rs; in.dtype = dt;
  kern<<<grid, DEFAULT_THREADS, smem, (cudaStream_t)stream>>>(in, reinterpret_cast<const T*>(kernel), pl,
                                                 reinterpret_cast<T*>(w_out), out_rs, n_rows);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace

extern "C" int dspb_convolve_wf_f32(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, const void* kernel,
                                    int64_t m, int32_t mode_in, DSPB_WAVE_OUT(w_out), int64_t p, DSPB_TAIL) {
  (void)fatal;
  return launch_convolve<float>(w_in, w_in_row_stride, w_in_dtype, n_rows, n, kernel, m, mode_in, w_out,
                                w_out_row_stride, p, stream);
}
extern "C" int dspb_convolve_wf_f64(DSPB_WAVE_IN(w_in), int64_t n_rows, int64_t n, const void* kernel,
                                    int64_t m, int32_t mode_in, DSPB_WAVE_OUT(w_out), int64_t p, DSPB_TAIL) {
  (void)fatal;
  return launch_convolve<double>(w_in, w_in_row_stride, w_in_dtype, n_rows, n, kernel, m, mode_in, w_out,
                                 w_out_row_stride, p, stream);
}
