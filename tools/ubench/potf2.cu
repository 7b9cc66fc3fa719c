// Micro-benchmark of the 32 x 32 pivot block of the diagonal-tile Cholesky (one warp, lane = row):
// variants of the loop in potrf128_prog_dev (stba_chol.cu).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
constexpr int QLD = 132;
__device__ __forceinline__ double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-d * y, y, 1.0);
  y = fma(0.5 * y, e, y);
  e = fma(-d * y, y, 1.0);
  return fma(0.5 * y, e, y);
}
__device__ __forceinline__ double fast_rcp(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}
// V0: the loop as it is in stba_chol.cu
__device__ __forceinline__ void v0(double* D, double* xd, double* cb, int* info, int lane) {
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? D[c * QLD + lane] : 0.0;
  bool bad = false;
  cb[lane] = a[0];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double* col = cb + (j & 1) * 32;
    const double d = col[j];
    if (!(d > 0.0) && !bad) { bad = true; if (lane == 0) atomicCAS(info, 0, j + 1); }
    const double u = (j + 1 < 32) ? a[j] * col[j + 1] : 0.0;
    const double r = fast_rcp(d);
    if (j + 1 < 32) {
      a[j + 1] = fma(-u, r, a[j + 1]);
      cb[((j + 1) & 1) * 32 + lane] = a[j + 1];
    }
    const double t = a[j] * r;
#pragma unroll
    for (int k = j + 2; k < 32; ++k) a[k] = fma(-t, col[k], a[k]);
    const double rs = fast_rsqrt(d);
    a[j] = (lane == j) ? d * rs : a[j] * rs;
    if (lane == j) xd[j] = rs;
    __syncwarp();
  }
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c <= lane) D[c * QLD + lane] = a[c];
}
// V1: no branch in the loop (minimum pivot tracked, checked once), full 32 x 32 exchange buffer (no reuse ->
// a single __syncwarp per step orders everything), scaling of the columns after the loop
__device__ __forceinline__ void v1(double* D, double* xd, double* cb /* 32 x 32 */, int* info, int lane) {
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? D[c * QLD + lane] : 0.0;
  double dmin = 1.0;
  double rr[32];
  cb[lane] = a[0];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double* col = cb + j * 32;
    const double d = col[j];
    dmin = fmin(dmin, d);
    const double u = (j + 1 < 32) ? a[j] * col[j + 1] : 0.0;
    const double r = fast_rcp(d);
    rr[j] = d;
    if (j + 1 < 32) {
      a[j + 1] = fma(-u, r, a[j + 1]);
      cb[(j + 1) * 32 + lane] = a[j + 1];
      __syncwarp();
    }
    const double t = a[j] * r;
#pragma unroll
    for (int k = j + 2; k < 32; ++k) a[k] = fma(-t, col[k], a[k]);
  }
  if (!(dmin > 0.0) && lane == 0) atomicCAS(info, 0, 1);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double rs = fast_rsqrt(rr[j]);
    a[j] = (lane == j) ? rr[j] * rs : a[j] * rs;
    if (lane == j) xd[j] = rs;
  }
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c <= lane) D[c * QLD + lane] = a[c];
}
// V2: as V1 with the column read back as double2 (16 LDS.128 instead of 32 LDS.64 per step)
__device__ __forceinline__ void v2(double* D, double* xd, double* cb, int* info, int lane) {
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? D[c * QLD + lane] : 0.0;
  double dmin = 1.0;
  double rr[32];
  cb[lane] = a[0];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double* col = cb + j * 32;
    const double2 dd = *reinterpret_cast<const double2*>(col + (j & ~1));
    const double d = (j & 1) ? dd.y : dd.x;
    double nx;
    if (j & 1) nx = (j + 1 < 32) ? col[j + 1] : 0.0; else nx = dd.y;
    dmin = fmin(dmin, d);
    const double u = a[j] * nx;
    const double r = fast_rcp(d);
    rr[j] = d;
    if (j + 1 < 32) {
      a[j + 1] = fma(-u, r, a[j + 1]);
      cb[(j + 1) * 32 + lane] = a[j + 1];
      __syncwarp();
    }
    const double t = a[j] * r;
    const int k0 = (j + 2 + 1) & ~1;
    if (((j + 2) & 1) && j + 2 < 32) a[j + 2] = fma(-t, col[j + 2], a[j + 2]);
#pragma unroll
    for (int k = k0; k < 32; k += 2) {
      const double2 c2 = *reinterpret_cast<const double2*>(col + k);
      a[k] = fma(-t, c2.x, a[k]);
      a[k + 1] = fma(-t, c2.y, a[k + 1]);
    }
  }
  if (!(dmin > 0.0) && lane == 0) atomicCAS(info, 0, 1);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double rs = fast_rsqrt(rr[j]);
    a[j] = (lane == j) ? rr[j] * rs : a[j] * rs;
    if (lane == j) xd[j] = rs;
  }
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c <= lane) D[c * QLD + lane] = a[c];
}
// V3: exchange by shuffles (whole warp converged here, unlike the warp-specialised branch of the real kernel)
__device__ __forceinline__ void v3(double* D, double* xd, double* cb, int* info, int lane) {
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? D[c * QLD + lane] : 0.0;
  double dmin = 1.0;
  double rr[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double d = __shfl_sync(0xffffffffu, a[j], j);
    dmin = fmin(dmin, d);
    const double r = fast_rcp(d);
    rr[j] = d;
    const double t = a[j] * r;
#pragma unroll
    for (int k = j + 1; k < 32; ++k) {
      const double ck = __shfl_sync(0xffffffffu, a[j], k);
      a[k] = fma(-t, ck, a[k]);
    }
  }
  if (!(dmin > 0.0) && lane == 0) atomicCAS(info, 0, 1);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double rs = fast_rsqrt(rr[j]);
    a[j] = (lane == j) ? rr[j] * rs : a[j] * rs;
    if (lane == j) xd[j] = rs;
  }
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c <= lane) D[c * QLD + lane] = a[c];
}

// V4: V1 + the bulk of the rank-1 update of step j deferred into step j + 1 (it fills the latency of the next
// pivot's exchange and reciprocal); pivots kept in shared memory instead of registers
__device__ __forceinline__ void v4(double* D, double* xd, double* cb, int* info, int lane) {
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? D[c * QLD + lane] : 0.0;
  double dmin = 1.0;
  cb[lane] = a[0];
  __syncwarp();
  double tp = 0.0;                 // t of the previous step
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double* col = cb + j * 32;
    const double d = col[j];
    const double nx = (j + 1 < 32) ? col[j + 1] : 0.0;
    // deferred: step j - 1's update of a[k], k >= j + 2
    if (j > 0) {
      const double* pc = cb + (j - 1) * 32;
#pragma unroll
      for (int k = j + 2; k < 32; ++k) a[k] = fma(-tp, pc[k], a[k]);
    }
    dmin = fmin(dmin, d);
    const double u = a[j] * nx;
    const double r = fast_rcp(d);
    if (j + 1 < 32) {
      a[j + 1] = fma(-u, r, a[j + 1]);
      cb[(j + 1) * 32 + lane] = a[j + 1];
      __syncwarp();
    }
    tp = a[j] * r;
    if (j + 2 < 32) a[j + 2] = fma(-tp, col[j + 2], a[j + 2]);
  }
  if (!(dmin > 0.0) && lane == 0) atomicCAS(info, 0, 1);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double d = cb[j * 32 + j];
    const double rs = fast_rsqrt(d);
    a[j] = (lane == j) ? d * rs : a[j] * rs;
    if (lane == j) xd[j] = rs;
  }
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c <= lane) D[c * QLD + lane] = a[c];
}
// V5: V4 with the update of a[j + 2] also moved off the issue path before the store (only one FMA between the
// reciprocal and the exchange), and the column of step j scaled right away (no second pass)
__device__ __forceinline__ void v5(double* D, double* xd, double* cb, int* info, int lane) {
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? D[c * QLD + lane] : 0.0;
  double dmin = 1.0;
  cb[lane] = a[0];
  __syncwarp();
  double tp = 0.0, dp = 1.0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double* col = cb + j * 32;
    const double d = col[j];
    const double nx = (j + 1 < 32) ? col[j + 1] : 0.0;
    const double u = a[j] * nx;
    const double r = fast_rcp(d);
    if (j > 0) {
      const double* pc = cb + (j - 1) * 32;
#pragma unroll
      for (int k = j + 2; k < 32; ++k) a[k] = fma(-tp, pc[k], a[k]);
      const double rs = fast_rsqrt(dp);
      a[j - 1] = (lane == j - 1) ? dp * rs : a[j - 1] * rs;
      if (lane == j - 1) xd[j - 1] = rs;
    }
    dmin = fmin(dmin, d);
    if (j + 1 < 32) {
      a[j + 1] = fma(-u, r, a[j + 1]);
      cb[(j + 1) * 32 + lane] = a[j + 1];
      __syncwarp();
    }
    tp = a[j] * r;
    dp = d;
    if (j + 2 < 32) a[j + 2] = fma(-tp, col[j + 2], a[j + 2]);
  }
  {
    const double rs = fast_rsqrt(dp);
    a[31] = (lane == 31) ? dp * rs : a[31] * rs;
    if (lane == 31) xd[31] = rs;
  }
  if (!(dmin > 0.0) && lane == 0) atomicCAS(info, 0, 1);
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c <= lane) D[c * QLD + lane] = a[c];
}
template <int V>
__global__ void k_potf2(const double* A, double* out, long long* cyc, int* info, int n_warps) {
  __shared__ double D[32 * QLD];
  __shared__ double xd[32];
  __shared__ __align__(16) double cb[32 * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) D[(e / 32) * QLD + e % 32] = A[e];
  __syncthreads();
  long long t0 = clock64();
  if (warp == 0) {
    if (V == 0) v0(D, xd, cb, info, lane);
    if (V == 1) v1(D, xd, cb, info, lane);
    if (V == 2) v2(D, xd, cb, info, lane);
    if (V == 3) v3(D, xd, cb, info, lane);
    if (V == 4) v4(D, xd, cb, info, lane);
    if (V == 5) v5(D, xd, cb, info, lane);
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) out[e] = D[(e / 32) * QLD + e % 32];
}

// V6: V4 exactly as it sits in potrf128_prog_dev: progress counter for the follower warps (volatile shared
// store after a block-level fence every four columns), pivot check by ballot, scale through shared memory
__device__ __forceinline__ void v6(double* D, double* xd, double* cb, int* info, int lane, volatile int* s_prog, bool fences) {
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? D[c * QLD + lane] : 0.0;
  cb[lane] = a[0];
  if (fences) __threadfence_block();
  __syncwarp();
  if (lane == 0) *s_prog = 1;
  double tp = 0.0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double* col = cb + j * 32;
    const double d = col[j];
    const double nx = (j + 1 < 32) ? col[j + 1] : 0.0;
    if (j > 0) {
      const double* pc = cb + (j - 1) * 32;
#pragma unroll
      for (int k = j + 2; k < 32; ++k) a[k] = fma(-tp, pc[k], a[k]);
    }
    const double u = a[j] * nx;
    const double r = fast_rcp(d);
    if (j + 1 < 32) {
      a[j + 1] = fma(-u, r, a[j + 1]);
      cb[(j + 1) * 32 + lane] = a[j + 1];
      if (fences && ((j & 3) == 3 || j == 30)) __threadfence_block();
      __syncwarp();
      if (((j & 3) == 3 || j == 30) && lane == 0) *s_prog = j + 2;
    }
    tp = a[j] * r;
    if (j + 2 < 32) a[j + 2] = fma(-tp, col[j + 2], a[j + 2]);
  }
  const double dl = cb[lane * 32 + lane];
  const unsigned badm = __ballot_sync(0xffffffffu, !(dl > 0.0));
  if (badm && lane == 0) atomicCAS(info, 0, __ffs(badm));
  const double rsl = fast_rsqrt(dl);
  xd[lane] = rsl;
  if (fences) __threadfence_block();
  __syncwarp();
  if (lane == 0) *s_prog = 33;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const double v = (c == lane) ? dl * rsl : a[c] * xd[c];
    if (c <= lane) D[c * QLD + lane] = v;
  }
}
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_potf2_big(const double* A, double* out, long long* cyc, int* info) {
  __shared__ double D[32 * QLD];
  __shared__ double xd[32];
  __shared__ __align__(16) double cb[32 * 32];
  __shared__ int s_prog;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) D[(e / 32) * QLD + e % 32] = A[e];
  if (threadIdx.x == 0) s_prog = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp == 0) {
    v6(D, xd, cb, info, lane, &s_prog, MODE != 0);
  } else if (MODE == 2 && warp < 4) {
    // three warps polling the progress counter like the followers do
    volatile int* sp = &s_prog;
    while (*sp < 33) { }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) out[e] = D[(e / 32) * QLD + e % 32];
}
int main() {
  std::vector<double> A(1024), L(1024), R(1024);
  srand(1);
  std::vector<double> B(1024);
  for (auto& b : B) b = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) {
      double s = (i == j) ? 4.0 : 0.0;
      for (int k = 0; k < 32; ++k) s += B[i * 32 + k] * B[j * 32 + k];
      A[j * 32 + i] = s;
    }
  R = A;
  for (int j = 0; j < 32; ++j) {
    const double d = sqrt(R[j * 32 + j]);
    for (int i = j; i < 32; ++i) R[j * 32 + i] /= d;
    for (int k = j + 1; k < 32; ++k)
      for (int i = k; i < 32; ++i) R[k * 32 + i] -= R[j * 32 + i] * R[j * 32 + k];
  }
  double *dA, *dO; long long* cyc; int* info;
  cudaMalloc(&dA, 8192); cudaMalloc(&dO, 8192); cudaMalloc(&cyc, 64); cudaMalloc(&info, 4);
  cudaMemcpy(dA, A.data(), 8192, cudaMemcpyHostToDevice);
  cudaMemset(info, 0, 4);
  for (int v = 0; v < 6; ++v) {
    long long h = 0;
    for (int rep = 0; rep < 3; ++rep) {
      if (v == 0) k_potf2<0><<<1, 128>>>(dA, dO, cyc, info, 16);
      if (v == 1) k_potf2<1><<<1, 128>>>(dA, dO, cyc, info, 16);
      if (v == 2) k_potf2<2><<<1, 128>>>(dA, dO, cyc, info, 16);
      if (v == 3) k_potf2<3><<<1, 128>>>(dA, dO, cyc, info, 16);
      if (v == 4) k_potf2<4><<<1, 128>>>(dA, dO, cyc, info, 16);
      if (v == 5) k_potf2<5><<<1, 128>>>(dA, dO, cyc, info, 16);
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    }
    cudaMemcpy(L.data(), dO, 8192, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int j = 0; j < 32; ++j)
      for (int i = j; i < 32; ++i) err = fmax(err, fabs(L[j * 32 + i] - R[j * 32 + i]));
    printf("variant %d: %lld cycles (%.0f per pivot), max |L - ref| = %.2e  %s\n", v, h, h / 32.0, err, cudaGetErrorString(cudaGetLastError()));
  }
  for (int v = 0; v < 3; ++v) {
    long long h = 0;
    for (int rep = 0; rep < 3; ++rep) {
      if (v == 0) k_potf2_big<0><<<1, 512>>>(dA, dO, cyc, info);
      if (v == 1) k_potf2_big<1><<<1, 512>>>(dA, dO, cyc, info);
      if (v == 2) k_potf2_big<2><<<1, 512>>>(dA, dO, cyc, info);
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    }
    cudaMemcpy(L.data(), dO, 8192, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int j = 0; j < 32; ++j)
      for (int i = j; i < 32; ++i) err = fmax(err, fabs(L[j * 32 + i] - R[j * 32 + i]));
    printf("512-thread kernel, mode %d (0 no fences, 1 fences, 2 fences + 3 polling warps): %lld cycles, max err %.2e %s\n", v, h, err, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
