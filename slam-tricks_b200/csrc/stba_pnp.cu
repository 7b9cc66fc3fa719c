// stba_pnp_gauss_newton — the reference's hand Gauss-Newton for the single-camera reprojection problem,
// `SelfGaussNewton`, st17-ceres/src/include/solver.hpp:387-462, on the B200.
//
// The reference loops on the host: per observation residual (:405-410) and 2 x 6 Jacobian [e_R | e_t] (:412-428),
// H += J^T J, g -= J^T r (:430-431), delta = H.ldlt().solve(g) (:434), R <- R Exp(d_theta), t <- t + d_t
// (:438-439), stop when |d_theta| + |d_t| < 1e-8 or after 10 iterations (:441-449).  Here the WHOLE loop is one
// kernel launch: one CTA per problem (the entry takes a batch of independent PnP problems: per-frame pose
// refinement is the same computation many times), the observations of a problem are strided over the CTA's
// threads, the 21 + 6 sums are reduced in a fixed tree order in shared memory (deterministic), thread 0 solves
// the 6 x 6 system by LDL^T and retracts.  No host round trip between iterations.
//
// jacobian_mode: STBA_PNP_JACOBIAN_REFERENCE reproduces the reference's e_R = Pi' (-R^-1 hat(P_w)) (-R)
// (solver.hpp:195; = Pi' hat(R^T P_w): it omits the translation term) so that iterates and the
// iteration count equal SelfGaussNewton's; STBA_PNP_JACOBIAN_EXACT uses Pi' hat(R^T (P_w - t)).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <vector>

#include "../../include/stba.h"
#include "stba_dev.cuh"

namespace {

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      fprintf(stderr, "[stba] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, \
              __LINE__, cudaGetErrorString(e_));                                              \
      return STBA_ERR_CUDA;                                                                   \
    }                                                                                         \
  } while (0)

constexpr int kGnThreads = 128;
constexpr int kGnAcc = 27;        // 21 unique entries of H (upper triangle, row-major) + 6 of g

__global__ void __launch_bounds__(kGnThreads)
k_pnp_gauss_newton(const int* __restrict__ ptr, const double* __restrict__ points, const double* __restrict__ uv, double* __restrict__ q_io,
                   double* __restrict__ t_io, int max_iterations, double tolerance, int jacobian_mode, int* __restrict__ iterations,
                   double* __restrict__ last_change, int* __restrict__ status) {
  __shared__ double red[kGnAcc][kGnThreads];
  __shared__ double pose[7];
  __shared__ int s_stop;
  const int prob = blockIdx.x, tid = threadIdx.x;
  const int beg = ptr[prob], end = ptr[prob + 1];
  if (tid < 4) pose[tid] = q_io[4 * (size_t)prob + tid];
  if (tid < 3) pose[4 + tid] = t_io[3 * (size_t)prob + tid];
  if (tid == 0) s_stop = 0;
  __syncthreads();
  int it = 0;
  double change = 0.0;
  for (; it != max_iterations; ++it) {            // `for (; i != 10; ++i)`, solver.hpp:401
    double R[9];
    stba::quat_to_rot(pose[0], pose[1], pose[2], pose[3], R);
    const double tx = pose[4], ty = pose[5], tz = pose[6];
    double acc[kGnAcc];
#pragma unroll
    for (int k = 0; k < kGnAcc; ++k) acc[k] = 0.0;
    for (int o = beg + tid; o < end; o += kGnThreads) {
      const double Px = points[3 * (size_t)o], Py = points[3 * (size_t)o + 1], Pz = points[3 * (size_t)o + 2];
      const double dx = Px - tx, dy = Py - ty, dz = Pz - tz;
      // p_c = R^T (P - t)
      const double X = R[0] * dx + R[3] * dy + R[6] * dz, Y = R[1] * dx + R[4] * dy + R[7] * dz, Z = R[2] * dx + R[5] * dy + R[8] * dz;
      const double iz = 1.0 / Z;
      const double r0 = X * iz - uv[2 * (size_t)o], r1 = Y * iz - uv[2 * (size_t)o + 1];
      // Pi' = [[1/Z, 0, -X/Z^2], [0, 1/Z, -Y/Z^2]]
      const double a0 = iz, a2 = -X * iz * iz, b1 = iz, b2 = -Y * iz * iz;
      // vector whose hat() the rotation Jacobian uses: R^T P_w (reference form) or p_c (exact)
      double hx, hy, hz;
      if (jacobian_mode == STBA_PNP_JACOBIAN_REFERENCE) {
        hx = R[0] * Px + R[3] * Py + R[6] * Pz; hy = R[1] * Px + R[4] * Py + R[7] * Pz; hz = R[2] * Px + R[5] * Py + R[8] * Pz;
      } else {
        hx = X; hy = Y; hz = Z;
      }
      // J = [Pi' hat(h) | -Pi' R^T], rows j0, j1
      double j0[6], j1[6];
      // hat(h) = [[0,-hz,hy],[hz,0,-hx],[-hy,hx,0]];  row * hat(h): (a0,0,a2) -> (a2*(-hy), a0*(-hz)+a2*hx, a0*hy)
      j0[0] = -a2 * hy;            j0[1] = -a0 * hz + a2 * hx;  j0[2] = a0 * hy;
      j1[0] = b1 * hz - b2 * hy;   j1[1] = b2 * hx;             j1[2] = -b1 * hx;
      // -Pi' R^T: column k of R^T is row k of R
      j0[3] = -(a0 * R[0] + a2 * R[2]); j0[4] = -(a0 * R[3] + a2 * R[5]); j0[5] = -(a0 * R[6] + a2 * R[8]);
      j1[3] = -(b1 * R[1] + b2 * R[2]); j1[4] = -(b1 * R[4] + b2 * R[5]); j1[5] = -(b1 * R[7] + b2 * R[8]);
      int k = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b, ++k) acc[k] = fma(j0[a], j0[b], fma(j1[a], j1[b], acc[k]));
#pragma unroll
      for (int a = 0; a < 6; ++a) acc[21 + a] -= j0[a] * r0 + j1[a] * r1;      // g = -sum J^T r
    }
#pragma unroll
    for (int k = 0; k < kGnAcc; ++k) red[k][tid] = acc[k];
    __syncthreads();
    for (int s = kGnThreads / 2; s > 0; s >>= 1) {
      if (tid < s)
#pragma unroll
        for (int k = 0; k < kGnAcc; ++k) red[k][tid] += red[k][tid + s];
      __syncthreads();
    }
    if (tid == 0) {
      // delta = H^-1 g by LDL^T (solver.hpp:434)
      double H[6][6], g[6], L[6][6], D[6], d[6];
      int k = 0;
      for (int a = 0; a < 6; ++a) for (int b = a; b < 6; ++b, ++k) H[a][b] = H[b][a] = red[k][0];
      for (int a = 0; a < 6; ++a) g[a] = red[21 + a][0];
      bool ok = true;
      for (int j = 0; j < 6; ++j) {
        double dj = H[j][j];
        for (int m = 0; m < j; ++m) dj -= L[j][m] * L[j][m] * D[m];
        D[j] = dj;
        if (!(fabs(dj) > 0.0) || !isfinite(dj)) ok = false;
        for (int i = j + 1; i < 6; ++i) {
          double s = H[i][j];
          for (int m = 0; m < j; ++m) s -= L[i][m] * L[j][m] * D[m];
          L[i][j] = s / dj;
        }
      }
      for (int i = 0; i < 6; ++i) { d[i] = g[i]; for (int m = 0; m < i; ++m) d[i] -= L[i][m] * d[m]; }
      for (int i = 0; i < 6; ++i) d[i] /= D[i];
      for (int i = 5; i >= 0; --i) for (int m = i + 1; m < 6; ++m) d[i] -= L[m][i] * d[m];
      if (!ok || !isfinite(d[0] + d[1] + d[2] + d[3] + d[4] + d[5])) {
        status[prob] = 1;
        s_stop = 1;
      } else {
        double e[4], qn[4];
        stba::so3_exp_quat(d[0], d[1], d[2], e);
        stba::quat_mul_normalized(pose, e, qn);                    // SO3 <- SO3 * exp(d_theta)   (:438)
        pose[0] = qn[0]; pose[1] = qn[1]; pose[2] = qn[2]; pose[3] = qn[3];
        pose[4] += d[3]; pose[5] += d[4]; pose[6] += d[5];        // POS <- POS + d_t            (:439)
        red[0][0] = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) + sqrt(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);   // change (:441)
        if (red[0][0] < tolerance) s_stop = 1;                     // :447
      }
    }
    __syncthreads();
    change = red[0][0];
    const int stop = s_stop;
    __syncthreads();
    if (stop) break;
  }
  if (tid < 4) q_io[4 * (size_t)prob + tid] = pose[tid];
  if (tid < 3) t_io[3 * (size_t)prob + tid] = pose[4 + tid];
  if (tid == 0) {
    iterations[prob] = it;          // the value the reference logs as "iter num" (the loop index at exit)
    last_change[prob] = change;
  }
}

}  // namespace

extern "C" int stba_pnp_gauss_newton(int device, int32_t n_problems, const int32_t* ptr, const double* points, const double* uv, double* q,
                                     double* t, int32_t max_iterations, double tolerance, int32_t jacobian_mode, int32_t* iterations,
                                     double* last_change, float* kernel_ms) {
  if (n_problems < 0 || max_iterations < 0 || (n_problems && (!ptr || !q || !t)) ||
      (jacobian_mode != STBA_PNP_JACOBIAN_REFERENCE && jacobian_mode != STBA_PNP_JACOBIAN_EXACT))
    return STBA_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < n_problems; ++i)
    if (ptr[i + 1] < ptr[i] || ptr[0] != 0) return STBA_ERR_INVALID_ARGUMENT;
  const int64_t n_obs = n_problems ? ptr[n_problems] : 0;
  if (n_obs && (!points || !uv)) return STBA_ERR_INVALID_ARGUMENT;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return STBA_ERR_NO_DEVICE; }
  if (device < 0 || device >= ndev) return STBA_ERR_INVALID_ARGUMENT;
  if (n_problems == 0) return STBA_OK;
  CK(cudaSetDevice(device));
  cudaStream_t s;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  int *d_ptr = nullptr, *d_it = nullptr, *d_status = nullptr;
  double *d_pts = nullptr, *d_uv = nullptr, *d_q = nullptr, *d_t = nullptr, *d_change = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = STBA_OK;
  std::vector<int> h_status((size_t)n_problems, 0);
#define CKG(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fprintf(stderr, "[stba] CUDA error %s at %s:%d\n", cudaGetErrorName(e__), __FILE__, __LINE__); rc = STBA_ERR_CUDA; goto done; } } while (0)
  CKG(cudaMallocAsync((void**)&d_ptr, (n_problems + 1) * sizeof(int), s));
  CKG(cudaMallocAsync((void**)&d_it, n_problems * sizeof(int), s));
  CKG(cudaMallocAsync((void**)&d_status, n_problems * sizeof(int), s));
  CKG(cudaMallocAsync((void**)&d_pts, std::max<int64_t>(n_obs, 1) * 3 * sizeof(double), s));
  CKG(cudaMallocAsync((void**)&d_uv, std::max<int64_t>(n_obs, 1) * 2 * sizeof(double), s));
  CKG(cudaMallocAsync((void**)&d_q, (size_t)n_problems * 4 * sizeof(double), s));
  CKG(cudaMallocAsync((void**)&d_t, (size_t)n_problems * 3 * sizeof(double), s));
  CKG(cudaMallocAsync((void**)&d_change, (size_t)n_problems * sizeof(double), s));
  CKG(cudaMemcpyAsync(d_ptr, ptr, (n_problems + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
  if (n_obs) {
    CKG(cudaMemcpyAsync(d_pts, points, (size_t)n_obs * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    CKG(cudaMemcpyAsync(d_uv, uv, (size_t)n_obs * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
  }
  CKG(cudaMemcpyAsync(d_q, q, (size_t)n_problems * 4 * sizeof(double), cudaMemcpyHostToDevice, s));
  CKG(cudaMemcpyAsync(d_t, t, (size_t)n_problems * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
  CKG(cudaMemsetAsync(d_status, 0, n_problems * sizeof(int), s));
  CKG(cudaEventCreate(&e0)); CKG(cudaEventCreate(&e1));
  CKG(cudaEventRecord(e0, s));
  k_pnp_gauss_newton<<<n_problems, kGnThreads, 0, s>>>(d_ptr, d_pts, d_uv, d_q, d_t, max_iterations, tolerance, jacobian_mode, d_it, d_change, d_status);
  CKG(cudaEventRecord(e1, s));
  CKG(cudaGetLastError());
  CKG(cudaMemcpyAsync(q, d_q, (size_t)n_problems * 4 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CKG(cudaMemcpyAsync(t, d_t, (size_t)n_problems * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (iterations) CKG(cudaMemcpyAsync(iterations, d_it, n_problems * sizeof(int), cudaMemcpyDeviceToHost, s));
  if (last_change) CKG(cudaMemcpyAsync(last_change, d_change, n_problems * sizeof(double), cudaMemcpyDeviceToHost, s));
  CKG(cudaMemcpyAsync(h_status.data(), d_status, n_problems * sizeof(int), cudaMemcpyDeviceToHost, s));
  CKG(cudaStreamSynchronize(s));
  if (kernel_ms) cudaEventElapsedTime(kernel_ms, e0, e1);
  for (int v : h_status) if (v) rc = STBA_ERR_SOLVER;
done:
#undef CKG
  for (void* p : {(void*)d_ptr, (void*)d_it, (void*)d_status, (void*)d_pts, (void*)d_uv, (void*)d_q, (void*)d_t, (void*)d_change})
    if (p) cudaFreeAsync(p, s);
  cudaStreamSynchronize(s);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaStreamDestroy(s);
  return rc;
}
