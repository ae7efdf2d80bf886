// dspeed_b200 -- structure-aware 'valid' convolution with cusp / zac kernels, shared by the
// interpreted chain kernel (fused.cu) and the specialised chain kernels (codegen.py).
#pragma once
#include "row_ops.cuh"

namespace dspb {

// ---------------------------------------------------------------------------------------
// CONV_SEG: valid-mode convolution with a cusp / zac kernel (energy_kernels.py:12-157)
// from weighted prefix sums.  With z[j] = x[j] - c x[j-1] (the kernel's [1,-c] factor moved
// onto the input) the kernel is sinh ramps (left: i in [0,lt), right: i in (lt+fl, L)), a
// flat top, and for zac additionally beta*(i^2 - 2 h i) on both ramps.  For output
// n = L-1+o the ramps/flat are windows of j = n - i bounded by
//   hiA = L+o, loA = L-lt+o, loB = L-1-lt-fl+o, loC = o
// and every windowed sum of w(j) z[j] is a difference of exclusive prefix sums at those
// bounds.  prm: [sigma, lt, fl, L, c, 1/(2 sinh(lt/sigma)), k[L-1], beta, h, is_zac]
// tab: >= 13*p doubles of scratch.
// ---------------------------------------------------------------------------------------
template <typename T>
__device__ void op_conv_seg(const T* x, int N, T* out, double* tab, const double* prm, Scratch* sc) {
  const double sigma = prm[0];
  const int lt = (int)prm[1], fl = (int)prm[2], L = (int)prm[3];
  const double c = prm[4], inv2S = prm[5], kLm1 = prm[6], beta = prm[7], h = prm[8];
  const bool zac = prm[9] != 0.0;
  const int p = N - L + 1;
  const int base[4] = {L, L - lt, L - 1 - lt - fl, 0};  // hiA, loA, loB, loC for o = 0
  double* ysave = tab + 12 * p;
  int lo, hi;
  chunk_range(N, lo, hi);
  auto zval = [&](int j) -> double {
    const double xj = (double)x[sidx(j)];
    return j > 0 ? xj - c * (double)x[sidx(j - 1)] : xj;
  };
  auto put = [&](int m, double r0, double r1, double r2) {
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int o = m - base[b];
      if (o >= 0 && o < p) {
        tab[(b * 3 + 0) * p + o] = r0;
        tab[(b * 3 + 1) * p + o] = r1;
        tab[(b * 3 + 2) * p + o] = r2;
      }
    }
  };
  // ---- exponential pass: weights e^{-j/sigma}, e^{+j/sigma}, 1 ------------------------------
  {
    const double qm = exp(-1.0 / sigma), qp = exp(1.0 / sigma);
    double wm = exp(-(double)lo / sigma), wp = exp((double)lo / sigma);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int j = lo; j < hi; j++) {
      const double z = zval(j);
      s0 += wm * z; s1 += wp * z; s2 += z;
      wm *= qm; wp *= qp;
    }
    double t0, t1, t2;
    double r0 = block_excl_scan(s0, t0, sc);
    double r1 = block_excl_scan(s1, t1, sc);
    double r2 = block_excl_scan(s2, t2, sc);
    wm = exp(-(double)lo / sigma); wp = exp((double)lo / sigma);
    for (int j = lo; j < hi; j++) {
      put(j, r0, r1, r2);
      const double z = zval(j);
      r0 += wm * z; r1 += wp * z; r2 += z;
      wm *= qm; wp *= qp;
    }
    if (hi == N && lo < hi) put(N, r0, r1, r2);
    __syncthreads();
    for (int o = threadIdx.x; o < p; o += NT) {
      const double n = (double)(L - 1 + o);
      const double en = exp(n / sigma), eLn = exp(((double)L - n) / sigma);
      const double enm = 1.0 / en, eLnm = 1.0 / eLn;
      // tab index: (b*3 + w)*p + o ; b: 0 hiA, 1 loA, 2 loB, 3 loC ; w: 0 Em, 1 Ep, 2 P0
#define TB(b, w) tab[((b)*3 + (w)) * p + o]
      const double yA = (en * (TB(0, 0) - TB(1, 0)) - enm * (TB(0, 1) - TB(1, 1))) * inv2S;
      const double yB = TB(1, 2) - TB(2, 2);
      const double yC = (eLn * (TB(2, 1) - TB(3, 1)) - eLnm * (TB(2, 0) - TB(3, 0))) * inv2S;
      ysave[o] = yA + yB + yC;
    }
    __syncthreads();
  }
  // ---- polynomial pass (zac): weights 1, jc, jc^2 with jc = j - N/2 ----------------------------
  if (zac) {
    const double j0 = (double)N * 0.5;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int j = lo; j < hi; j++) {
      const double z = zval(j), jc = (double)j - j0;
      s0 += z; s1 += jc * z; s2 += jc * jc * z;
    }
    double t0, t1, t2;
    double r0 = block_excl_scan(s0, t0, sc);
    double r1 = block_excl_scan(s1, t1, sc);
    double r2 = block_excl_scan(s2, t2, sc);
    for (int j = lo; j < hi; j++) {
      put(j, r0, r1, r2);
      const double z = zval(j), jc = (double)j - j0;
      r0 += z; r1 += jc * z; r2 += jc * jc * z;
    }
    if (hi == N && lo < hi) put(N, r0, r1, r2);
    __syncthreads();
    for (int o = threadIdx.x; o < p; o += NT) {
      const double n = (double)(L - 1 + o);
      const double nc = n - j0;
      double dP = TB(0, 0) - TB(1, 0), dM1 = TB(0, 1) - TB(1, 1), dM2 = TB(0, 2) - TB(1, 2);
      double s2a = nc * nc * dP - 2.0 * nc * dM1 + dM2, s1a = nc * dP - dM1;
      const double yA2 = s2a - 2.0 * h * s1a;
      dP = TB(2, 0) - TB(3, 0); dM1 = TB(2, 1) - TB(3, 1); dM2 = TB(2, 2) - TB(3, 2);
      const double a = ((double)L - n) + j0;
      s2a = a * a * dP + 2.0 * a * dM1 + dM2; s1a = a * dP + dM1;
      const double yC2 = s2a - 2.0 * h * s1a;
      ysave[o] += beta * (yA2 + yC2);
    }
    __syncthreads();
  }
#undef TB
  for (int o = threadIdx.x; o < p; o += NT) {
    const int n = L - 1 + o;
    const double xm = (n - L >= 0) ? (double)x[sidx(n - L)] : 0.0;
    out[sidx(o)] = (T)(ysave[o] + c * kLm1 * xm);
  }
  __syncthreads();
}

}  // namespace dspb
