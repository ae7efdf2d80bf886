// micro-benchmark: issue cost of float<->double conversions vs FP64 arithmetic on sm_100a
// (cycles per warp instruction per SM sub-partition, 1..8 warps per scheduler)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double f2d_bits(float f) {
  const unsigned u = __float_as_uint(f);
  unsigned hi = (u & 0x80000000u) | (((u >> 3) & 0x0fffffffu) + 0x38000000u);
  unsigned lo = u << 29;
  if (((u >> 23) & 0xffu) == 0u) { hi = u & 0x80000000u; lo = 0u; }
  return __hiloint2double((int)hi, (int)lo);
}
template <int OP>
__global__ void k(float* out, long long* cyc, int iters, float seed) {
  float a[8];
  double d[8];
  for (int j = 0; j < 8; j++) { a[j] = seed + j + threadIdx.x; d[j] = (double)a[j]; }
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (OP == 0) { d[j] = (double)a[j] ; a[j] = (float)__double2hiint(d[j]) * 1e-9f + a[j]; }      // F2F.F64.F32 (+ cheap dependency)
      if (OP == 1) { a[j] = (float)d[j]; d[j] = __hiloint2double(__double2hiint(d[j]) ^ __float_as_int(a[j]) & 1, __double2loint(d[j])); }  // F2F.F32.F64
      if (OP == 2) { d[j] = fma(d[j], 1.0000001, 0.5); }                         // DFMA
      if (OP == 3) { d[j] = d[j] + 1.5; }                                       // DADD
      if (OP == 4) { d[j] = f2d_bits(a[j]); a[j] = (float)__double2hiint(d[j]) * 1e-9f + a[j]; }     // bit-trick conversion
      if (OP == 5) { a[j] = fmaf(a[j], 1.0000001f, 0.5f); }                     // FFMA (reference)
      if (OP == 6) { d[j] = (double)(__float_as_int(a[j]) >> 8); a[j] = (float)__double2hiint(d[j]) * 1e-9f + a[j]; }   // I2F.F64
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < 8; j++) s += a[j] + (float)d[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const char* names[] = {"F2F.F64.F32 (+I2F,FFMA)", "F2F.F32.F64 (+LOP)", "DFMA", "DADD", "bit-trick f32->f64 (+I2F,FFMA)", "FFMA", "I2F.F64.S32 (+I2F,FFMA)"};
  const int iters = 2000;
  for (int op = 0; op < 7; op++) {
    for (int warps = 4; warps <= 32; warps *= 2) {   // warps per CTA (1 CTA on 1 SM): 1, 2, 4, 8 per scheduler
      long long h = 0;
      for (int rep = 0; rep < 2; rep++) {
        switch (op) {
          case 0: k<0><<<1, warps * 32>>>(out, cyc, iters, 1.f); break;
          case 1: k<1><<<1, warps * 32>>>(out, cyc, iters, 1.f); break;
          case 2: k<2><<<1, warps * 32>>>(out, cyc, iters, 1.f); break;
          case 3: k<3><<<1, warps * 32>>>(out, cyc, iters, 1.f); break;
          case 4: k<4><<<1, warps * 32>>>(out, cyc, iters, 1.f); break;
          case 5: k<5><<<1, warps * 32>>>(out, cyc, iters, 1.f); break;
          case 6: k<6><<<1, warps * 32>>>(out, cyc, iters, 1.f); break;
        }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      const double per = (double)h / (iters * 8.0);                 // cycles per (op group) per warp, as seen by one warp
      const double per_sched = per / (warps / 4.0);                 // issue cycles per op group per scheduler
      printf("%-34s warps/sched %d: %.2f cycles per group per warp, %.2f per scheduler\n", names[op], warps / 4, per, per_sched);
    }
  }
  return 0;
}
