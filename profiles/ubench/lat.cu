// Dependent-chain latencies (cycles per op, one warp per SM) of the operations the K5 critical path is made of.
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
template <int OP>
__global__ void k(double* out, double a, double b, int nwarps_active) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = a;
  __syncthreads();
  double x = a + threadIdx.x * 1e-9, y = b;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, y, y);
    if (OP == 1) x = x + y;
    if (OP == 2) x = x * y;
    if (OP == 3) x = 1.0 / x + y;
    if (OP == 4) x = sqrt(x) + y;
    if (OP == 5) x = rsqrt(x) + y;
    if (OP == 6) { double s, c; sincos(x, &s, &c); x = s + c; }
    if (OP == 7) x = __shfl_xor_sync(0xffffffffu, x, 1) + y;
    if (OP == 8) { x = sm[(__double2loint(x) & 7)] + y; }
    if (OP == 9) { asm volatile("bar.sync 1, 192;" ::: "memory"); }
    if (OP == 10) { float f = (float)x; f = f * 1.0001f + 1.0f; x = (double)f; }
    if (OP == 11) { int q = __double2int_rn(x); q = q / (int)b + 7; x = (double)q; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = (double)(t1 - t0) / N; out[blockIdx.x * 2 + 1] = x; }
}
int main() {
  double* d; cudaMalloc(&d, 1024);
  const char* names[] = {"DFMA", "DADD", "DMUL", "1/x (+add)", "sqrt (+add)", "rsqrt (+add)", "sincos (+add)", "shfl64 (+add)", "LDS.64 dependent (+add)", "bar.sync 192thr", "f64->f32 fmul fadd ->f64", "f64->int idiv int->f64"};
  for (int threads : {32, 192}) {
    printf("threads per block = %d (1 block)\n", threads);
    for (int op = 0; op < 12; ++op) {
      if (op == 9 && threads != 192) continue;
      double h[2];
      switch (op) {
#define C(O) case O: k<O><<<1, threads>>>(d, 1.2345, 1.000001, 0); break;
        C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11)
      }
      cudaDeviceSynchronize();
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("  %-28s %.1f cycles\n", names[op], h[0]);
    }
  }
  return 0;
}
