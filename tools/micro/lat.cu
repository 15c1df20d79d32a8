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
  // issue rate of ONE warp: 8 independent DFMA chains (no dependence stalls: 8 x 2 cycles of issue > 8.8 of latency
  // if a warp's FP64 instruction issues every 2 cycles), then the same mixed 1:1 with integer instructions
  double q0 = x0, q1 = x0 + 1, q2 = x0 + 2, q3 = x0 + 3, q4 = x0 + 4, q5 = x0 + 5, q6 = x0 + 6, q7 = x0 + 7;
  long long t6 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    q0 = fma(q0, b, c); q1 = fma(q1, b, c); q2 = fma(q2, b, c); q3 = fma(q3, b, c);
    q4 = fma(q4, b, c); q5 = fma(q5, b, c); q6 = fma(q6, b, c); q7 = fma(q7, b, c);
  }
  long long t7 = clock64();
  // 16-byte shared loads feeding independent FMAs (the column update of the block factorisation)
  const double2* sm2 = reinterpret_cast<const double2*>(sm);
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    const double2 u0 = sm2[(i + 0) & 15], u1 = sm2[(i + 1) & 15], u2 = sm2[(i + 2) & 15], u3 = sm2[(i + 3) & 15];
    q0 = fma(q0, u0.x, c); q1 = fma(q1, u0.y, c); q2 = fma(q2, u1.x, c); q3 = fma(q3, u1.y, c);
    q4 = fma(q4, u2.x, c); q5 = fma(q5, u2.y, c); q6 = fma(q6, u3.x, c); q7 = fma(q7, u3.y, c);
  }
  long long t8 = clock64();
  a += q0 + q1 + q2 + q3 + q4 + q5 + q6 + q7;
  if (threadIdx.x == 0) {
    cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4;
    cyc[5] = t7 - t6; cyc[6] = t8 - t7;
  }
  out[threadIdx.x] = a + idx;
}
int main() {
  double* d; long long* c;
  cudaMalloc(&d, 32 * 8); cudaMalloc(&c, 7 * 8);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(d, c, 1.5);
  long long h[7];
  cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
  printf("dependent latency (cycles): DFMA %.1f  DMUL %.1f  rsqrt.approx+DADD %.1f  SHFL64 %.1f  LDS+cvt chain %.1f\n",
         h[0] / 1024.0, h[1] / 1024.0, h[2] / 512.0, h[3] / 1024.0, h[4] / 1024.0);
  printf("one warp, 8 independent DFMA chains: %.2f cycles per DFMA; with a 16-byte LDS per two DFMAs: %.2f cycles per DFMA\n",
         h[5] / 2048.0, h[6] / 2048.0);
  return 0;
}
