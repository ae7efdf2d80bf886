// dspeed_b200 -- register-tiled direct convolution core shared by the stand-alone
// convolution kernel (conv.cu) and the fused chain kernel (fused.cu).
#pragma once
#include "common.cuh"

namespace dspb {

constexpr int R = 8;    // outputs per thread
constexpr int CH = 64;  // taps per float32 accumulation chunk (multiple of 8)

template <typename T>
struct ConvPlan {
  int n, m, p, off;  // out[k] = sum_j a[j] v[k + off - j]
  int G, S, L;       // output groups, tap segments, taps per segment (multiple of 8)
};

// sample idx of a view starting at element `off` of a slot (0 outside the view)
template <typename T>
__device__ __forceinline__ T lda(const T* a, int off, int n, int idx) {
  return (idx >= 0 && idx < n) ? a[sidx(off + idx)] : (T)0;
}

// accumulate taps [t_lo, t_hi) for outputs k0..k0+7 (kk0 = k0 + off)
template <typename T>
__device__ __forceinline__ void conv_group_off(const T* a, int off, int n, const T* kv, int kk0, int t_lo,
                                               int t_hi, double (&acc)[R]) {
  for (int tb = t_lo; tb < t_hi; tb += CH) {
    const int te = min(tb + CH, t_hi);
    T f[R];
#pragma unroll
    for (int r = 0; r < R; r++) f[r] = (T)0;
    int t = tb;
    // window W[j] = a[kk0 - t - 7 + j], j = 0..14 ; output r at tap t+u reads W[r - u + 7]
    T W[15];
#pragma unroll
    for (int j = 7; j < 15; j++) W[j] = lda<T>(a, off, n, kk0 - t - 7 + j);
    for (; t + 8 <= te; t += 8) {
#pragma unroll
      for (int j = 0; j < 7; j++) W[j] = lda<T>(a, off, n, kk0 - t - 7 + j);
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const T kvv = kv[t + u];
#pragma unroll
        for (int r = 0; r < R; r++) f[r] = fma(W[r - u + 7], kvv, f[r]);
      }
      // next iteration (t+8): W'[j] = a[kk0 - t - 15 + j]; W'[8..14] = W[0..6]; W'[7] = a[kk0-t-8]
#pragma unroll
      for (int j = 14; j >= 8; j--) W[j] = W[j - 8];
      W[7] = lda<T>(a, off, n, kk0 - t - 8);
    }
    for (; t < te; t++) {  // tail (< 8 taps)
      const T kvv = kv[t];
#pragma unroll
      for (int r = 0; r < R; r++) f[r] = fma(lda<T>(a, off, n, kk0 + r - t), kvv, f[r]);
    }
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] += (double)f[r];
  }
}


template <typename T>
__device__ __forceinline__ void conv_group(const T* a, int n, const T* kv, int kk0, int t_lo, int t_hi,
                                           double (&acc)[R]) {
  conv_group_off<T>(a, 0, n, kv, kk0, t_lo, t_hi, acc);
}

}  // namespace dspb
