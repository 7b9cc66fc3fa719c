// Micro-benchmarks of the dependent-issue latencies that bound the pivot chain of the diagonal-block
// Cholesky (stba_chol.cu): nvcc -O3 -gencode arch=compute_100a,code=sm_100a lat.cu -o lat && ./lat
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double* out, long long* cyc, double seed) {
  __shared__ double sb[64];
  const int lane = threadIdx.x & 31;
  double x = seed + lane * 1e-3, y = 1.000001, z = 0.5;
  long long t0, t1;
  // 1. dependent DFMA chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = fma(x, y, z);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 1024;
  // 2. dependent rcp.approx.ftz.f64 chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x));
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = (t1 - t0) / 1024;
  // 3. dependent rsqrt.approx.ftz.f64
  x = fabs(x) + 1.0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x));
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = (t1 - t0) / 1024;
  // 4. STS -> syncwarp -> LDS (other lane) round trip
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      sb[(j & 1) * 32 + lane] = x;
      __syncwarp();
      x = sb[(j & 1) * 32 + ((lane + 1) & 31)];
    }
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = (t1 - t0) / 1024;
  // 5. 64-bit shuffle round trip
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = (t1 - t0) / 1024;
  // 6. dependent DMUL+DADD mix
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = x * y;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = (t1 - t0) / 1024;
  // 7. dependent FFMA
  float f = (float)x, fy = 1.00001f, fz = 0.5f;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f = fmaf(f, fy, fz);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = (t1 - t0) / 1024;
  // 8. dependent DMMA m8n8k4 chain
  double c0 = x, c1 = y;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(y), "d"(z));
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = (t1 - t0) / 1024;
  out[threadIdx.x] = x + f + c0 + c1;
}
// throughput: many warps, independent DFMA / DMMA streams
__global__ void k_tput(double* out, long long* cyc, double seed, int mode) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x * 1e-3;
  const double y = 1.0000001, z = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {
#pragma unroll 1
    for (int it = 0; it < 256; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], y, z);
  } else {
#pragma unroll 1
    for (int it = 0; it < 256; ++it)
#pragma unroll
      for (int i = 0; i < 8; i += 2)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(a[i]), "+d"(a[i + 1]) : "d"(y), "d"(z));
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[threadIdx.x] = s;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64 * 8);
  long long h[8];
  for (int rep = 0; rep < 2; ++rep) {
    k_lat<<<1, 32>>>(out, cyc, 1.5);
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
  }
  printf("latency (cycles): DFMA %lld  rcp.approx.f64 %lld  rsqrt.approx.f64 %lld  STS+syncwarp+LDS %lld  shfl64 %lld  DMUL %lld  FFMA %lld  DMMA %lld\n",
         h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
  for (int threads = 128; threads <= 1024; threads *= 2)
    for (int mode = 0; mode < 2; ++mode) {
      k_tput<<<1, threads>>>(out, cyc, 1.5, mode);
      k_tput<<<1, threads>>>(out, cyc, 1.5, mode);
      cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
      const double ops = mode == 0 ? 256.0 * 8 * threads : 256.0 * 4 * (threads / 32) * 256;   // FMAs
      printf("throughput %s threads %4d: %lld cycles, %.1f FMA/clk/SM\n", mode ? "DMMA" : "DFMA", threads, h[0], ops / h[0]);
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
