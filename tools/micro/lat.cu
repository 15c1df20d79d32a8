// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/lat tools/micro/lat.cu ; run on the GPU box: tools/micro/lat
// dependent-issue latencies on one warp: DFMA, DMUL, DADD, MUFU.RSQ64H, 64-bit shuffle, LDS
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double x0) {
  __shared__ double sm[64];
  sm[threadIdx.x] = x0 + threadIdx.x;
  __syncthreads();
  double a = x0, b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c);
  }
  long long t1 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    a = a * b; a = a * b; a = a * b; a = a * b;
  }
  long long t2 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    double y;
    asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    a = y + 1.0;
    asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    a = y + 1.0;
  }
  long long t3 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    a = __shfl_sync(0xffffffffu, a, (i + 1) & 31);
    a = __shfl_sync(0xffffffffu, a, (i + 2) & 31);
    a = __shfl_sync(0xffffffffu, a, (i + 3) & 31);
    a = __shfl_sync(0xffffffffu, a, (i + 4) & 31);
  }
  long long t4 = clock64();
  int idx = threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    idx = (int)sm[idx & 31] & 31; idx = (int)sm[idx & 31] & 31; idx = (int)sm[idx & 31] & 31; idx = (int)sm[idx & 31] & 31;
  }
  long long t5 = clock64();
  if (threadIdx.x == 0) {
    cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4;
  }
  out[threadIdx.x] = a + idx;
}
int main() {
  double* d; long long* c;
  cudaMalloc(&d, 32 * 8); cudaMalloc(&c, 5 * 8);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(d, c, 1.5);
  long long h[5];
  cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
  printf("dependent latency (cycles): DFMA %.1f  DMUL %.1f  rsqrt.approx+DADD %.1f  SHFL64 %.1f  LDS+cvt chain %.1f\n",
         h[0] / 1024.0, h[1] / 1024.0, h[2] / 512.0, h[3] / 1024.0, h[4] / 1024.0);
  return 0;
}
