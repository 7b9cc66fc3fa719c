// Zhang calibration on the B200 (BASELINE.json configs[3]).
//
//   stba_calib_initialize   CalibSolver::computeHomoMats / reconstructIntriMat / reconstructExtriMat,
//                           st3-calibration/src/src/calib.cpp:49-173 — V tiny SVDs; host code, as in
//                           the reference ("initialisation only; keep on host")
//   stba_calib_optimize     CalibSolver::totalOptimization, calib.cpp:282-422 — the joint Gauss-Newton
//                           over 4 intrinsics + 5 distortion + 6 V pose parameters, on the device.
//
// The reference accumulates a dense (9+6V)^2 matrix with a rank-2 update per corner
// (`H += J * J.transpose()`, calib.cpp:383-389) although only a 15 x 15 sub-block is non-zero, then
// runs a dense LDLT.  Here the arrow structure is used directly: one CTA per view accumulates its
// 15 x 15 block [A_i B_i; B_i^T C_i] in a fixed order (deterministic), one CTA eliminates the pose
// blocks (S = sum A_i - sum B_i C_i^-1 B_i^T, 9 x 9), solves, back-substitutes and applies the
// reference's LEFT-perturbation update T_i <- exp(d_i) T_i (calib.cpp:397-402).  Poses stay on the
// device as (R, t); the se3 logarithm the reference re-takes every iteration is only taken on exit.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/stba.h"
#include "stba_pool.cuh"

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

constexpr double kEps = 1e-10;       // Sophus::Constants<double>::epsilon()
constexpr int kP = 15;               // 9 shared + 6 pose parameters touch one corner
constexpr int kTri = kP * (kP + 1) / 2;
constexpr int kRow = 32;             // doubles per staged corner: Jx[15], Jy[15], ex, ey
constexpr int kTile = 128;           // corners per shared-memory tile
constexpr int kAccThreads = 160;     // >= kTri + kP + 1 accumulating threads
constexpr int kViewOut = kTri + kP + 1;

// ---- host + device SO(3)/SE(3) closed forms (Sophus semantics) ---------------------------------
__host__ __device__ inline void so3_exp_matrix(const double* w, double* R) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double imag, real;
  if (th2 < kEps * kEps) {
    imag = 0.5 - th2 / 48.0 + th2 * th2 / 3840.0;
    real = 1.0 - th2 / 8.0 + th2 * th2 / 384.0;
  } else {
    const double th = sqrt(th2);
    imag = sin(0.5 * th) / th;
    real = cos(0.5 * th);
  }
  const double x = imag * w[0], y = imag * w[1], z = imag * w[2], q = real;
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * q);     R[2] = 2 * (x * z + y * q);
  R[3] = 2 * (x * y + z * q);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * q);
  R[6] = 2 * (x * z - y * q);     R[7] = 2 * (y * z + x * q);     R[8] = 1 - 2 * (x * x + y * y);
}

// V(omega) ups of Sophus::SE3d::exp
__host__ __device__ inline void se3_exp(const double* xi, double* R, double* t) {
  const double* u = xi;
  const double* w = xi + 3;
  so3_exp_matrix(w, R);
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
  if (th < kEps) {
    for (int i = 0; i < 3; ++i) t[i] = R[3 * i] * u[0] + R[3 * i + 1] * u[1] + R[3 * i + 2] * u[2];
    return;
  }
  const double a = (1.0 - cos(th)) / th2, b = (th - sin(th)) / (th2 * th);
  // V u = u + a (w x u) + b (w x (w x u))
  const double c0 = w[1] * u[2] - w[2] * u[1], c1 = w[2] * u[0] - w[0] * u[2], c2 = w[0] * u[1] - w[1] * u[0];
  const double d0 = w[1] * c2 - w[2] * c1, d1 = w[2] * c0 - w[0] * c2, d2 = w[0] * c1 - w[1] * c0;
  t[0] = u[0] + a * c0 + b * d0;
  t[1] = u[1] + a * c1 + b * d1;
  t[2] = u[2] + a * c2 + b * d2;
}

// rotation matrix -> quaternion xyzw (w >= 0) -> Sophus::SO3d::log -> Sophus::SE3d::log
void se3_log_host(const double* R, const double* t, double* xi) {
  double q[4];
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    const double s = sqrt(tr + 1.0) * 2;
    q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s; q[3] = 0.25 * s;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    const double s = sqrt(1.0 + R[4 * i] - R[4 * j] - R[4 * k]) * 2;
    q[3] = (R[3 * k + j] - R[3 * j + k]) / s;
    q[i] = 0.25 * s;
    q[j] = (R[3 * j + i] + R[3 * i + j]) / s;
    q[k] = (R[3 * k + i] + R[3 * i + k]) / s;
  }
  if (q[3] < 0) for (double& v : q) v = -v;
  const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (double& v : q) v /= nq;
  const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
  double f;
  if (n2 < kEps * kEps) {
    f = 2.0 / q[3] - (2.0 / 3.0) * n2 / (q[3] * q[3] * q[3]);
  } else {
    const double n = sqrt(n2);
    f = 2.0 * (q[3] < 0.0 ? atan2(-n, -q[3]) : atan2(n, q[3])) / n;
  }
  double* w = xi + 3;
  w[0] = f * q[0]; w[1] = f * q[1]; w[2] = f * q[2];
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
  // V^-1 t = t - 1/2 (w x t) + c (w x (w x t))
  const double c = th < kEps ? 1.0 / 12.0 : (1.0 - th * cos(0.5 * th) / (2.0 * sin(0.5 * th))) / th2;
  const double c0 = w[1] * t[2] - w[2] * t[1], c1 = w[2] * t[0] - w[0] * t[2], c2 = w[0] * t[1] - w[1] * t[0];
  const double d0 = w[1] * c2 - w[2] * c1, d1 = w[2] * c0 - w[0] * c2, d2 = w[0] * c1 - w[1] * c0;
  xi[0] = t[0] - 0.5 * c0 + c * d0;
  xi[1] = t[1] - 0.5 * c1 + c * d1;
  xi[2] = t[2] - 0.5 * c2 + c * d2;
}

// ---- host: one-sided Jacobi SVD (Hestenes) of a row-major m x n matrix, n <= 9 -------------------
// On exit the columns of A are U * Sigma and V (n x n, row-major) holds the right singular vectors.
void jacobi_svd(std::vector<double>& A, int m, int n, double* V) {
  for (int i = 0; i < n * n; ++i) V[i] = (i / n == i % n) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double al = 0, be = 0, ga = 0;
        for (int r = 0; r < m; ++r) {
          const double a = A[(size_t)r * n + p], b = A[(size_t)r * n + q];
          al += a * a; be += b * b; ga += a * b;
        }
        if (fabs(ga) <= 1e-15 * sqrt(al * be) || ga == 0.0) continue;
        rotated = true;
        const double zeta = (be - al) / (2.0 * ga);
        const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
        for (int r = 0; r < m; ++r) {
          const double a = A[(size_t)r * n + p], b = A[(size_t)r * n + q];
          A[(size_t)r * n + p] = c * a - s * b;
          A[(size_t)r * n + q] = s * a + c * b;
        }
        for (int r = 0; r < n; ++r) {
          const double a = V[r * n + p], b = V[r * n + q];
          V[r * n + p] = c * a - s * b;
          V[r * n + q] = s * a + c * b;
        }
      }
    if (!rotated) break;
  }
}

// right singular vector of the smallest singular value (Eigen sorts descending: `matrixV().col(n-1)`)
void null_vector(std::vector<double>& A, int m, int n, double* v) {
  double V[81];
  jacobi_svd(A, m, n, V);
  int best = 0;
  double bn = 1e300;
  for (int j = 0; j < n; ++j) {
    double s = 0;
    for (int r = 0; r < m; ++r) s += A[(size_t)r * n + j] * A[(size_t)r * n + j];
    if (s < bn) { bn = s; best = j; }
  }
  for (int r = 0; r < n; ++r) v[r] = V[r * n + best];
}

// ---- device: residual and Jacobian of one corner, calib.cpp:318-380 ------------------------------
struct Cam {
  double alpha, beta, u0, v0, k1, k2, k3, p1, p2;
};

__device__ __forceinline__ void corner(const Cam& c, const double* __restrict__ R, const double* __restrict__ t, double X, double Y,
                                       double u, double v, double* __restrict__ row) {
  const double Xp = fma(R[0], X, fma(R[1], Y, t[0])), Yp = fma(R[3], X, fma(R[4], Y, t[1])), Zp = fma(R[6], X, fma(R[7], Y, t[2]));
  const double iz = 1.0 / Zp, xn = Xp * iz, yn = Yp * iz;
  const double r2 = xn * xn + yn * yn, r4 = r2 * r2, r6 = r4 * r2;
  const double rad = 1.0 + c.k1 * r2 + c.k2 * r4 + c.k3 * r6;
  const double xd = xn * rad + 2.0 * c.p1 * xn * yn + c.p2 * (r2 + 2.0 * xn * xn);
  const double yd = yn * rad + 2.0 * c.p2 * xn * yn + c.p1 * (r2 + 2.0 * yn * yn);
  double* Jx = row;
  double* Jy = row + kP;
  row[30] = c.alpha * xd + c.u0 - u;
  row[31] = c.beta * yd + c.v0 - v;
  // intrinsics (:337-339)
  Jx[0] = xd; Jx[1] = 0.0; Jx[2] = 1.0; Jx[3] = 0.0;
  Jy[0] = 0.0; Jy[1] = yd; Jy[2] = 0.0; Jy[3] = 1.0;
  // distortion (:342-348)
  Jx[4] = c.alpha * xn * r2; Jy[4] = c.beta * yn * r2;
  Jx[5] = c.alpha * xn * r4; Jy[5] = c.beta * yn * r4;
  Jx[6] = c.alpha * xn * r6; Jy[6] = c.beta * yn * r6;
  Jx[7] = 2.0 * c.alpha * xn * yn; Jy[7] = c.beta * (r2 + 2.0 * yn * yn);
  Jx[8] = c.alpha * (r2 + 2.0 * xn * xn); Jy[8] = 2.0 * c.beta * xn * yn;
  // pose (:352-380): diag(alpha, beta) * d(dist)/d(n) * Pi' * [I | -hat(P')]
  const double dr = 2.0 * c.k1 + 4.0 * c.k2 * r2 + 6.0 * c.k3 * r4;
  const double a00 = rad + xn * (dr * xn) + 2.0 * c.p1 * yn + 6.0 * c.p2 * xn;
  const double a01 = xn * (dr * yn) + 2.0 * c.p1 * xn + 2.0 * c.p2 * yn;
  const double a10 = yn * (dr * xn) + 2.0 * c.p1 * xn + 2.0 * c.p2 * yn;
  const double a11 = rad + yn * (dr * yn) + 2.0 * c.p2 * xn + 6.0 * c.p1 * yn;
  // M = diag(alpha,beta) * pd_pn * pn_PPrime  (2 x 3)
  const double m00 = c.alpha * a00 * iz, m01 = c.alpha * a01 * iz, m02 = -c.alpha * (a00 * xn + a01 * yn) * iz;
  const double m10 = c.beta * a10 * iz, m11 = c.beta * a11 * iz, m12 = -c.beta * (a10 * xn + a11 * yn) * iz;
  Jx[9] = m00; Jx[10] = m01; Jx[11] = m02;
  Jy[9] = m10; Jy[10] = m11; Jy[11] = m12;
  // M * (-hat(P')):  -hat(P) = [[0, Zp, -Yp], [-Zp, 0, Xp], [Yp, -Xp, 0]]
  Jx[12] = -m01 * Zp + m02 * Yp; Jx[13] = m00 * Zp - m02 * Xp; Jx[14] = -m00 * Yp + m01 * Xp;
  Jy[12] = -m11 * Zp + m12 * Yp; Jy[13] = m10 * Zp - m12 * Xp; Jy[14] = -m10 * Yp + m11 * Xp;
}

// One CTA per view: block[0..120) upper triangle of the view's 15 x 15 J J^T, [120..135) g = -J e,
// [135] cost.  Corners are staged tile by tile; every sum runs in corner order.
__global__ void __launch_bounds__(kAccThreads)
k_calib_accumulate(const int* __restrict__ view_ptr, const double* __restrict__ obj, const double* __restrict__ img,
                   const double* __restrict__ param, const double* __restrict__ pose, double* __restrict__ out) {
  __shared__ double s_rows[kTile][kRow + 1];
  __shared__ unsigned char s_a[kTri], s_b[kTri];
  const int v = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    int e = 0;
    for (int a = 0; a < kP; ++a)
      for (int b = a; b < kP; ++b) { s_a[e] = (unsigned char)a; s_b[e] = (unsigned char)b; ++e; }
  }
  Cam c;
  c.alpha = param[0]; c.beta = param[1]; c.u0 = param[2]; c.v0 = param[3];
  c.k1 = param[4]; c.k2 = param[5]; c.k3 = param[6]; c.p1 = param[7]; c.p2 = param[8];
  double R[9], t[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = pose[12 * v + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) t[k] = pose[12 * v + 9 + k];
  const int beg = view_ptr[v], end = view_ptr[v + 1];
  double acc = 0.0;
  __syncthreads();
  for (int b0 = beg; b0 < end; b0 += kTile) {
    const int nt = min(kTile, end - b0);
    if (tid < nt) corner(c, R, t, obj[2 * (b0 + tid)], obj[2 * (b0 + tid) + 1], img[2 * (b0 + tid)], img[2 * (b0 + tid) + 1], s_rows[tid]);
    __syncthreads();
    if (tid < kTri) {
      const int a = s_a[tid], b = s_b[tid];
      for (int r = 0; r < nt; ++r) acc = fma(s_rows[r][a], s_rows[r][b], fma(s_rows[r][kP + a], s_rows[r][kP + b], acc));
    } else if (tid < kTri + kP) {
      const int a = tid - kTri;
      for (int r = 0; r < nt; ++r) acc -= fma(s_rows[r][a], s_rows[r][30], s_rows[r][kP + a] * s_rows[r][31]);
    } else if (tid == kTri + kP) {
      for (int r = 0; r < nt; ++r) acc = fma(s_rows[r][30], s_rows[r][30], fma(s_rows[r][31], s_rows[r][31], acc));
    }
    __syncthreads();
  }
  if (tid < kViewOut) out[(size_t)v * kViewOut + tid] = (tid == kTri + kP) ? 0.5 * acc : acc;
}

__device__ __forceinline__ int tri15(int a, int b) { return a * kP - (a * (a - 1)) / 2 + (b - a); }   // a <= b

// One CTA: eliminate the pose blocks, solve the 9 x 9 system, back-substitute, update.
// scal[0] = |update|, scal[1] = cost at the linearisation point, scal[2] = failure flag.
__global__ void __launch_bounds__(1024)
k_calib_solve(int V, const double* __restrict__ blocks, double* __restrict__ param, double* __restrict__ pose,
              double* __restrict__ work /* V x 54 */, double* __restrict__ scal) {
  __shared__ double s_S[45], s_g[9], s_d[9], s_red[1024];
  __shared__ int s_fail;
  const int i = threadIdx.x;
  if (i == 0) s_fail = 0;
  __syncthreads();
  double L[21], B[54], g6[6];          // L: Cholesky factor of C_i, row-major lower (r,c) at r(r+1)/2 + c
  if (i < V) {
    const double* blk = blocks + (size_t)i * kViewOut;
    for (int r = 0; r < 6; ++r)
      for (int c2 = 0; c2 <= r; ++c2) L[r * (r + 1) / 2 + c2] = blk[tri15(9 + c2, 9 + r)];
    for (int a = 0; a < 9; ++a)
      for (int c2 = 0; c2 < 6; ++c2) B[a * 6 + c2] = blk[tri15(a, 9 + c2)];
    for (int c2 = 0; c2 < 6; ++c2) g6[c2] = blk[kTri + 9 + c2];
    bool ok = true;
    for (int c2 = 0; c2 < 6; ++c2) {
      double d = L[c2 * (c2 + 1) / 2 + c2];
      for (int k = 0; k < c2; ++k) d -= L[c2 * (c2 + 1) / 2 + k] * L[c2 * (c2 + 1) / 2 + k];
      ok = ok && d > 0.0;
      d = sqrt(d);
      L[c2 * (c2 + 1) / 2 + c2] = d;
      for (int r = c2 + 1; r < 6; ++r) {
        double s = L[r * (r + 1) / 2 + c2];
        for (int k = 0; k < c2; ++k) s -= L[r * (r + 1) / 2 + k] * L[c2 * (c2 + 1) / 2 + k];
        L[r * (r + 1) / 2 + c2] = s / d;
      }
    }
    if (!ok) atomicExch(&s_fail, 1);
    // Y = L^-1 B^T (6 x 9, stored over B as Y[a][c] = row a of B solved), z = L^-1 g6
    for (int a = 0; a < 9; ++a)
      for (int r = 0; r < 6; ++r) {
        double s = B[a * 6 + r];
        for (int k = 0; k < r; ++k) s -= L[r * (r + 1) / 2 + k] * B[a * 6 + k];
        B[a * 6 + r] = s / L[r * (r + 1) / 2 + r];
      }
    for (int r = 0; r < 6; ++r) {
      double s = g6[r];
      for (int k = 0; k < r; ++k) s -= L[r * (r + 1) / 2 + k] * g6[k];
      g6[r] = s / L[r * (r + 1) / 2 + r];
    }
    // contribution  B C^-1 B^T = Y Y^T (45 unique) and B C^-1 g6 = Y z (9)
    double* w = work + (size_t)i * 54;
    int e = 0;
    for (int a = 0; a < 9; ++a)
      for (int b = a; b < 9; ++b) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s = fma(B[a * 6 + k], B[b * 6 + k], s);
        w[e++] = s;
      }
    for (int a = 0; a < 9; ++a) {
      double s = 0.0;
      for (int k = 0; k < 6; ++k) s = fma(B[a * 6 + k], g6[k], s);
      w[45 + a] = s;
    }
  }
  __threadfence_block();
  __syncthreads();
  if (i < 54) {                          // reduced system, summed over views in view order
    double s = 0.0;
    if (i < 45) {
      int a = 0, rem = i;
      while (rem >= 9 - a) { rem -= 9 - a; ++a; }
      const int b = a + rem;
      for (int v = 0; v < V; ++v) s += blocks[(size_t)v * kViewOut + tri15(a, b)] - work[(size_t)v * 54 + i];
      s_S[i] = s;
    } else {
      for (int v = 0; v < V; ++v) s += blocks[(size_t)v * kViewOut + kTri + (i - 45)] - work[(size_t)v * 54 + i];
      s_g[i - 45] = s;
    }
  }
  __syncthreads();
  if (i == 0) {                          // 9 x 9 LDL^T without pivoting (the reference: `H.ldlt().solve(g)`, calib.cpp:393)
    double A[9][9], D[9], y[9];
    int e = 0;
    for (int a = 0; a < 9; ++a)
      for (int b = a; b < 9; ++b) { A[b][a] = s_S[e]; A[a][b] = s_S[e]; ++e; }
    for (int c2 = 0; c2 < 9; ++c2) {
      double d = A[c2][c2];
      for (int k = 0; k < c2; ++k) d -= A[c2][k] * A[c2][k] * D[k];
      D[c2] = d;
      if (!(fabs(d) > 0.0) || !isfinite(d)) s_fail = 1;
      for (int r = c2 + 1; r < 9; ++r) {
        double s = A[r][c2];
        for (int k = 0; k < c2; ++k) s -= A[r][k] * A[c2][k] * D[k];
        A[r][c2] = s / d;
      }
    }
    for (int r = 0; r < 9; ++r) {
      double s = s_g[r];
      for (int k = 0; k < r; ++k) s -= A[r][k] * y[k];
      y[r] = s;
    }
    for (int r = 0; r < 9; ++r) y[r] /= D[r];
    for (int r = 8; r >= 0; --r) {
      double s = y[r];
      for (int k = r + 1; k < 9; ++k) s -= A[k][r] * s_d[k];
      s_d[r] = s;
    }
  }
  __syncthreads();
  double n2 = 0.0;
  if (i < V) {
    // d_i = C^-1 (g6 - B^T d9) = L^-T (z - Y^T d9)
    double d[6];
    for (int k = 0; k < 6; ++k) {
      double s = g6[k];
      for (int a = 0; a < 9; ++a) s -= B[a * 6 + k] * s_d[a];
      d[k] = s;
    }
    for (int r = 5; r >= 0; --r) {
      double s = d[r];
      for (int k = r + 1; k < 6; ++k) s -= L[k * (k + 1) / 2 + r] * d[k];
      d[r] = s / L[r * (r + 1) / 2 + r];
    }
    for (int k = 0; k < 6; ++k) n2 = fma(d[k], d[k], n2);
    if (!s_fail) {                        // T <- exp(d) T   (calib.cpp:397-402)
      double Rd[9], td[3], R[9], t[3], Rn[9], tn[3];
      se3_exp(d, Rd, td);
      for (int k = 0; k < 9; ++k) R[k] = pose[12 * i + k];
      for (int k = 0; k < 3; ++k) t[k] = pose[12 * i + 9 + k];
      for (int r = 0; r < 3; ++r) {
        for (int c2 = 0; c2 < 3; ++c2) Rn[3 * r + c2] = Rd[3 * r] * R[c2] + Rd[3 * r + 1] * R[3 + c2] + Rd[3 * r + 2] * R[6 + c2];
        tn[r] = Rd[3 * r] * t[0] + Rd[3 * r + 1] * t[1] + Rd[3 * r + 2] * t[2] + td[r];
      }
      for (int k = 0; k < 9; ++k) pose[12 * i + k] = Rn[k];
      for (int k = 0; k < 3; ++k) pose[12 * i + 9 + k] = tn[k];
    }
  }
  s_red[i] = n2;
  __syncthreads();
  if (i == 0) {
    double s = 0.0, cost = 0.0;
    for (int a = 0; a < 9; ++a) s = fma(s_d[a], s_d[a], s);
    for (int v = 0; v < V; ++v) { s += s_red[v]; cost += blocks[(size_t)v * kViewOut + kTri + kP]; }
    if (!s_fail)
      for (int a = 0; a < 9; ++a) param[a] += s_d[a];       // calib.cpp:394
    scal[0] = sqrt(s);
    scal[1] = cost;
    scal[2] = s_fail ? 1.0 : 0.0;
  }
}

bool bad_views(int32_t n_views, const int32_t* view_ptr) {
  if (n_views <= 0 || !view_ptr || view_ptr[0] != 0) return true;
  for (int i = 0; i < n_views; ++i)
    if (view_ptr[i + 1] < view_ptr[i]) return true;
  return false;
}

}  // namespace

extern "C" {

int stba_calib_initialize(int32_t n_views, const int32_t* view_ptr, const double* obj_xy, const double* img_uv, double* intrinsics,
                          double* poses, double* homographies) {
  if (bad_views(n_views, view_ptr) || !obj_xy || !img_uv || !intrinsics || !poses) return STBA_ERR_INVALID_ARGUMENT;
  std::vector<double> Hs(9 * (size_t)n_views);
  for (int v = 0; v < n_views; ++v) {                       // computeHomoMat, calib.cpp:55-93
    const int b = view_ptr[v], n = view_ptr[v + 1] - b;
    if (n < 4) return STBA_ERR_INVALID_ARGUMENT;
    std::vector<double> A(2 * (size_t)n * 9, 0.0);
    for (int i = 0; i < n; ++i) {
      const double x = obj_xy[2 * (b + i)], y = obj_xy[2 * (b + i) + 1], u = img_uv[2 * (b + i)], w = img_uv[2 * (b + i) + 1];
      double* r0 = &A[(size_t)(2 * i) * 9];
      double* r1 = &A[(size_t)(2 * i + 1) * 9];
      r0[0] = x; r0[1] = y; r0[2] = 1.0; r0[6] = -u * x; r0[7] = -u * y; r0[8] = -u;
      r1[3] = x; r1[4] = y; r1[5] = 1.0; r1[6] = -w * x; r1[7] = -w * y; r1[8] = -w;
    }
    double h[9];
    null_vector(A, 2 * n, 9, h);
    // The reference keeps the sign Eigen's JacobiSVD happens to return; H and -H give the same K,
    // distortion and cost but mirrored poses.  Fix it so that the board is in front of the camera.
    if (h[8] < 0) for (double& e : h) e = -e;
    memcpy(&Hs[9 * (size_t)v], h, sizeof(h));
  }
  if (homographies) memcpy(homographies, Hs.data(), Hs.size() * sizeof(double));
  {                                                          // reconstructIntriMat, calib.cpp:95-140
    std::vector<double> C(2 * (size_t)n_views * 5);
    auto cof = [&](const double* H, int i, int j, double* o) {
      const double hi[3] = {H[i], H[3 + i], H[6 + i]}, hj[3] = {H[j], H[3 + j], H[6 + j]};
      o[0] = hi[0] * hj[0]; o[1] = hi[2] * hj[0] + hi[0] * hj[2]; o[2] = hi[1] * hj[1];
      o[3] = hi[2] * hj[1] + hi[1] * hj[2]; o[4] = hi[2] * hj[2];
    };
    for (int v = 0; v < n_views; ++v) {
      double c12[5], c11[5], c22[5];
      cof(&Hs[9 * (size_t)v], 0, 1, c12); cof(&Hs[9 * (size_t)v], 0, 0, c11); cof(&Hs[9 * (size_t)v], 1, 1, c22);
      for (int k = 0; k < 5; ++k) { C[(size_t)(2 * v) * 5 + k] = c12[k]; C[(size_t)(2 * v + 1) * 5 + k] = c11[k] - c22[k]; }
    }
    if (2 * n_views < 5) return STBA_ERR_INVALID_ARGUMENT;
    double b[5];
    null_vector(C, 2 * n_views, 5, b);
    const double b11 = b[0], b13 = b[1], b22 = b[2], b23 = b[3], b33 = b[4];
    const double v0 = -b23 / b22;
    const double lambda = b33 - (b13 * b13 - v0 * b11 * b23) / b11;
    const double alpha = sqrt(lambda / b11), beta = sqrt(lambda / b22);
    intrinsics[0] = alpha; intrinsics[1] = beta; intrinsics[2] = -b13 * alpha * alpha / lambda; intrinsics[3] = v0;
    if (!std::isfinite(alpha) || !std::isfinite(beta)) return STBA_ERR_SOLVER;
  }
  const double alpha = intrinsics[0], beta = intrinsics[1], u0 = intrinsics[2], v0 = intrinsics[3];
  for (int v = 0; v < n_views; ++v) {                       // reconstructExtriMat, calib.cpp:142-173
    const double* H = &Hs[9 * (size_t)v];
    auto kinv = [&](int col, double* o) {                   // K^-1 h
      const double a = H[col], b = H[3 + col], c = H[6 + col];
      o[0] = (a - u0 * c) / alpha; o[1] = (b - v0 * c) / beta; o[2] = c;
    };
    double r1[3], r2[3], r3[3], t[3];
    kinv(0, r1); kinv(1, r2); kinv(2, t);
    const double n1 = sqrt(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]), n2 = sqrt(r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]);
    const double lambda = 1.0 / (2.0 * n1) + 1.0 / (2.0 * n2);
    for (int k = 0; k < 3; ++k) { r1[k] /= n1; r2[k] /= n2; t[k] *= lambda; }
    r3[0] = r1[1] * r2[2] - r1[2] * r2[1]; r3[1] = r1[2] * r2[0] - r1[0] * r2[2]; r3[2] = r1[0] * r2[1] - r1[1] * r2[0];
    r1[0] = r2[1] * r3[2] - r2[2] * r3[1]; r1[1] = r2[2] * r3[0] - r2[0] * r3[2]; r1[2] = r2[0] * r3[1] - r2[1] * r3[0];
    std::vector<double> M = {r1[0], r2[0], r3[0], r1[1], r2[1], r3[1], r1[2], r2[2], r3[2]};
    double Vm[9], R[9];
    jacobi_svd(M, 3, 3, Vm);                                 // M = U Sigma V^T  ->  R = U V^T
    for (int j = 0; j < 3; ++j) {
      const double s = sqrt(M[j] * M[j] + M[3 + j] * M[3 + j] + M[6 + j] * M[6 + j]);
      for (int r = 0; r < 3; ++r) M[3 * r + j] /= s;
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) R[3 * r + c] = M[3 * r] * Vm[3 * c] + M[3 * r + 1] * Vm[3 * c + 1] + M[3 * r + 2] * Vm[3 * c + 2];
    se3_log_host(R, t, poses + 6 * (size_t)v);
  }
  return STBA_OK;
}

int stba_calib_optimize(int device, int32_t n_views, const int32_t* view_ptr, const double* obj_xy, const double* img_uv,
                        double* intrinsics, double* distortion, double* poses, int32_t max_iterations, double tolerance,
                        int32_t* iterations_run, double* update_norms, double* costs, int64_t* gpu_launches) {
  return stba_calib_optimize_timed(device, n_views, view_ptr, obj_xy, img_uv, intrinsics, distortion, poses, max_iterations, tolerance,
                                   iterations_run, update_norms, costs, gpu_launches, nullptr, nullptr);
}

int stba_calib_optimize_timed(int device, int32_t n_views, const int32_t* view_ptr, const double* obj_xy, const double* img_uv,
                              double* intrinsics, double* distortion, double* poses, int32_t max_iterations, double tolerance,
                              int32_t* iterations_run, double* update_norms, double* costs, int64_t* gpu_launches,
                              float* loop_ms, float* accumulate_ms) {
  if (bad_views(n_views, view_ptr) || n_views > 1024 || !obj_xy || !img_uv || !intrinsics || !distortion || !poses || max_iterations < 0)
    return STBA_ERR_INVALID_ARGUMENT;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return STBA_ERR_NO_DEVICE; }
  if (device < 0 || device >= ndev) return STBA_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(device));
  const int n = view_ptr[n_views];
  std::vector<double> h_pose(12 * (size_t)n_views), h_param(9);
  for (int v = 0; v < n_views; ++v) se3_exp(poses + 6 * (size_t)v, &h_pose[12 * (size_t)v], &h_pose[12 * (size_t)v + 9]);
  memcpy(h_param.data(), intrinsics, 4 * sizeof(double));
  memcpy(h_param.data() + 4, distortion, 5 * sizeof(double));
  stba::keep_default_pool(device);
  cudaStream_t s;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  int* d_ptr = nullptr;
  double *d_obj = nullptr, *d_img = nullptr, *d_param = nullptr, *d_pose = nullptr, *d_blocks = nullptr, *d_work = nullptr, *d_scal = nullptr;
  auto cleanup = [&]() {
    for (void* p : {(void*)d_ptr, (void*)d_obj, (void*)d_img, (void*)d_param, (void*)d_pose, (void*)d_blocks, (void*)d_work, (void*)d_scal})
      if (p) cudaFreeAsync(p, s);
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
  };
#define CKC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fprintf(stderr, "[stba] CUDA error %s at %s:%d\n", cudaGetErrorName(e__), __FILE__, __LINE__); cleanup(); return STBA_ERR_CUDA; } } while (0)
  CKC(cudaMallocAsync((void**)&d_ptr, (n_views + 1) * sizeof(int), s));
  CKC(cudaMallocAsync((void**)&d_obj, std::max(n, 1) * 2 * sizeof(double), s));
  CKC(cudaMallocAsync((void**)&d_img, std::max(n, 1) * 2 * sizeof(double), s));
  CKC(cudaMallocAsync((void**)&d_param, 9 * sizeof(double), s));
  CKC(cudaMallocAsync((void**)&d_pose, 12 * (size_t)n_views * sizeof(double), s));
  CKC(cudaMallocAsync((void**)&d_blocks, (size_t)n_views * kViewOut * sizeof(double), s));
  CKC(cudaMallocAsync((void**)&d_work, (size_t)n_views * 54 * sizeof(double), s));
  CKC(cudaMallocAsync((void**)&d_scal, 4 * sizeof(double), s));
  CKC(cudaMemcpyAsync(d_ptr, view_ptr, (n_views + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
  CKC(cudaMemcpyAsync(d_obj, obj_xy, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
  CKC(cudaMemcpyAsync(d_img, img_uv, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
  CKC(cudaMemcpyAsync(d_param, h_param.data(), 9 * sizeof(double), cudaMemcpyHostToDevice, s));
  CKC(cudaMemcpyAsync(d_pose, h_pose.data(), h_pose.size() * sizeof(double), cudaMemcpyHostToDevice, s));
  int it = 0;
  int64_t launches = 0;
  int status = STBA_OK;
  const int solve_threads = std::max(64, ((n_views + 31) / 32) * 32);
  // optional device timing (bench.py --workload CALIB): the whole Gauss-Newton loop with the inputs resident,
  // and the accumulation kernel alone
  const bool timed = loop_ms || accumulate_ms;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
  float acc_total = 0.f;
  if (timed) {
    cudaEventCreate(&ev_a); cudaEventCreate(&ev_b); cudaEventCreate(&ev_k0); cudaEventCreate(&ev_k1);
    cudaEventRecord(ev_a, s);
  }
  for (; it < max_iterations;) {                            // calib.cpp:298
    if (timed) cudaEventRecord(ev_k0, s);
    k_calib_accumulate<<<n_views, kAccThreads, 0, s>>>(d_ptr, d_obj, d_img, d_param, d_pose, d_blocks);
    if (timed) cudaEventRecord(ev_k1, s);
    k_calib_solve<<<1, solve_threads, 0, s>>>(n_views, d_blocks, d_param, d_pose, d_work, d_scal);
    launches += 2;
    double sc[3];
    CKC(cudaMemcpyAsync(sc, d_scal, sizeof(sc), cudaMemcpyDeviceToHost, s));
    CKC(cudaStreamSynchronize(s));
    CKC(cudaGetLastError());
    if (timed) { float ms = 0.f; cudaEventElapsedTime(&ms, ev_k0, ev_k1); acc_total += ms; }
    if (update_norms) update_norms[it] = sc[0];
    if (costs) costs[it] = sc[1];
    ++it;
    if (sc[2] != 0.0 || !std::isfinite(sc[0])) { status = STBA_ERR_SOLVER; break; }
    if (sc[0] < tolerance) break;                           // calib.cpp:404
  }
  if (timed) {
    cudaEventRecord(ev_b, s);
    cudaEventSynchronize(ev_b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev_a, ev_b);
    if (loop_ms) *loop_ms = ms;
    if (accumulate_ms) *accumulate_ms = acc_total;
    cudaEventDestroy(ev_a); cudaEventDestroy(ev_b); cudaEventDestroy(ev_k0); cudaEventDestroy(ev_k1);
  }
  CKC(cudaMemcpyAsync(h_param.data(), d_param, 9 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CKC(cudaMemcpyAsync(h_pose.data(), d_pose, h_pose.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
  CKC(cudaStreamSynchronize(s));
#undef CKC
  cleanup();
  memcpy(intrinsics, h_param.data(), 4 * sizeof(double));
  memcpy(distortion, h_param.data() + 4, 5 * sizeof(double));
  for (int v = 0; v < n_views; ++v) se3_log_host(&h_pose[12 * (size_t)v], &h_pose[12 * (size_t)v + 9], poses + 6 * (size_t)v);
  if (iterations_run) *iterations_run = it;
  if (gpu_launches) *gpu_launches = launches;
  return status;
}

}  // extern "C"
