// fp64 pipe micro-benchmarks on the B200: DFMA, DMMA m8n8k4, DMMA m16n8k16, cublasDgemm, syrk.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/peaks tools/peaks.cu -lcublas
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void k_dfma(int iters, double* out) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], 1.0000001, 1e-9);
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 1.2345) out[0] = s;
}
__global__ void k_dmma884(int iters, double* out) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 1.2345) out[0] = s;
}
__global__ void k_dmma16816(int iters, double* out) {
  double c[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-4 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  double s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 1.2345) out[0] = s;
}
__global__ void k_dmma1688(int iters, double* out) {
  double c[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
  double a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 2; ++i) b[i] = 1.0 + threadIdx.x * 1e-4 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  double s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 1.2345) out[0] = s;
}
template <typename F> float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (r && ms < best) best = ms; }
  return best;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, 8);
  const int sms = p.multiProcessorCount;
  for (int warps : {4, 8, 16, 32}) {
    const int grid = sms * (warps >= 16 ? 2 : 4), block = warps * 32 / (warps >= 16 ? 1 : 1) > 1024 ? 1024 : warps * 32;
    const int iters = 1 << 14;
    float t = timeit([&] { k_dfma<<<grid, block>>>(iters, out); });
    printf("DFMA        grid %d block %d : %.2f TFLOP/s\n", grid, block, 2.0 * 8 * iters * (double)grid * block / t / 1e9);
    t = timeit([&] { k_dmma884<<<grid, block>>>(iters, out); });
    printf("DMMA m8n8k4 grid %d block %d : %.2f TFLOP/s\n", grid, block, 2.0 * 8 * 256 * iters * (double)grid * (block / 32) / t / 1e9);
    t = timeit([&] { k_dmma1688<<<grid, block>>>(iters, out); });
    printf("DMMA m16n8k8 grid %d block %d : %.2f TFLOP/s\n", grid, block, 2.0 * 4 * 1024 * iters * (double)grid * (block / 32) / t / 1e9);
    t = timeit([&] { k_dmma16816<<<grid, block>>>(iters, out); });
    printf("DMMA m16n8k16 grid %d block %d : %.2f TFLOP/s\n", grid, block, 2.0 * 4 * 2048 * iters * (double)grid * (block / 32) / t / 1e9);
  }
  cublasHandle_t h; cublasCreate(&h);
  for (int n : {2048, 4096, 6144}) {
    double *A, *B, *C; cudaMalloc(&A, 8.0 * n * n); cudaMalloc(&B, 8.0 * n * n); cudaMalloc(&C, 8.0 * n * n);
    cudaMemset(A, 0, 8.0 * n * n); cudaMemset(B, 0, 8.0 * n * n); cudaMemset(C, 0, 8.0 * n * n);
    const double one = 1.0, zero = 0.0, mone = -1.0;
    float t = timeit([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n); });
    printf("cublasDgemm NT n=%d : %.3f ms  %.2f TFLOP/s\n", n, t, 2.0 * n * n * n / t / 1e9);
    t = timeit([&] { cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, 128, &mone, A, n, &one, C, n); });
    printf("cublasDsyrk n=%d k=128 : %.3f ms  %.2f TFLOP/s\n", n, t, 1.0 * n * n * 128 / t / 1e9);
    t = timeit([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, 128, &mone, A, n, B, n, &one, C, n); });
    printf("cublasDgemm NT n=%d k=128 : %.3f ms  %.2f TFLOP/s\n", n, t, 2.0 * n * n * 128 / t / 1e9);
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  return 0;
}
