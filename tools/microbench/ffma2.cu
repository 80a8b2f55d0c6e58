// Micro-benchmark: issue rate of scalar FFMA/FADD versus packed FFMA2/FADD2 (f32x2) on sm_100a.
// Prints warp-instructions per clock per SM for 1, 2, 4, 8 warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float *out, long long *cyc, int iters) {
  float2 a[8];
  for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.001f, -0.001f);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) {            // scalar FFMA x2 (two instructions per pair)
        a[i].x = fmaf(a[i].x, m.x, c.x);
        a[i].y = fmaf(a[i].y, m.y, c.y);
      } else if (MODE == 1) {     // packed FFMA2 (one instruction per pair)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(*(unsigned long long *) &a[i])
                     : "l"(*(const unsigned long long *) &m), "l"(*(const unsigned long long *) &c));
      } else if (MODE == 2) {     // scalar FADD x2
        a[i].x += c.x;
        a[i].y += c.y;
      } else {                    // packed FADD2
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(*(unsigned long long *) &a[i]) : "l"(*(const unsigned long long *) &c));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMallocManaged(&cyc, 148 * sizeof(long long));
  const int iters = 20000;
  const char *names[4] = {"FFMA  (2 per pair)", "FFMA2 (1 per pair)", "FADD  (2 per pair)", "FADD2 (1 per pair)"};
  for (int threads = 128; threads <= 1024; threads *= 2) {
    for (int mode = 0; mode < 4; mode++) {
      if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
      if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
      if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
      if (mode == 3) k<3><<<148, threads>>>(out, cyc, iters);
      cudaDeviceSynchronize();
      double c = (double) cyc[0];
      double pairs = (double) iters * 8 * (threads / 32);      // warp-level pair-operations per SM
      printf("warps/SMSP %d  %s : %.3f pair-ops/clk/SM  (%.3f warp-instr/clk/SM)\n", threads / 128, names[mode],
             pairs / c, pairs * ((mode & 1) ? 1 : 2) / c);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
