// SE(3) pose graph on the B200 (BASELINE.json configs[4]).
//
// The reference has no pose-graph code (only the chain simulator st4-kalman/src/src/pose_simulation.cpp:17-88,
// its recorded tracks st4-kalman/output/{truth,obs}.csv and the SE(3) Jacobian notes
// st23-lie-group-v2/doc.tex:862-997), so this path is specified by oracle/pg_oracle.py:
//   residual   r_ij = Log(Z_ij^-1 T_i^-1 T_j)  (Sophus order [rho, theta]),  manifold T <- T Exp(delta),
//   Jacobians  dr/d delta_j = J_r^-1(r),  dr/d delta_i = -J_r^-1(r) Ad(T_j^-1 T_i)   (exact),
//   solver     Ceres-faithful trust-region LM, exact solve of the damped normal
//              equations, pose 0 constant.
// Structure: J^T J is block-banded (6 x 6 blocks, half-bandwidth B = max |i - j| over the edges, "block-
// tridiagonal" for a pure chain).  Kernels:
//   k_pg_linearize   one thread per pose gathers its incident edges in a fixed order: diagonal block, the
//                    band blocks below it, gradient, cost share — no atomics, deterministic
//   k_pg_damp        Jacobi scaling + LM diagonal -> scaled damped band (the linearisation is kept for retries)
//   k_pg_band_solve  block-banded Cholesky + both substitutions in ONE CTA: a ring of B + 3 block columns
//                    lives in shared memory, the next column streams in with cp.async two steps ahead, the
//                    factor streams out; per column: 6 x 6 Cholesky + inverse, B block solves, B(B+1)/2
//                    block updates.  Sequential over the columns by nature: used for short graphs and for the
//                    reduced separator system of the partitioned solve (k_pg_part_*, further down), which
//                    factorises ~sqrt(N / B) interiors side by side on as many SMs.
//   k_pg_update / k_pg_cost / k_pg_gradnorm / k_pg_sum   step, candidate cost, norms (fixed-order sums)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <new>
#include <vector>

#include "../../include/stba.h"
#include "stba_chol.cuh"
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

constexpr int kMaxBand = 16;         // half-bandwidth in blocks the shared-memory ring is sized for
constexpr int kMaxClosureBand = 8;   // widest band next to loop closures (the partitioned kernels need 2B - 1 <= 16)
constexpr double kEps = 1e-10;       // Sophus epsilon (exp / log branches)
constexpr double kSmall = 1e-5;      // series branch of the Jacobians (oracle/pg_oracle.py SMALL)

// ---- small fixed-size algebra -----------------------------------------------------------------
struct V3 { double x, y, z; };
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
struct Q4 { double x, y, z, w; };
struct Pose { Q4 q; V3 t; };

__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
__device__ __forceinline__ Q4 qnormalize(Q4 q) {
  const double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return {q.x / n, q.y / n, q.z / n, q.w / n};
}
__device__ __forceinline__ void qrot_matrix(Q4 q, double* R) {
  const double x = q.x, y = q.y, z = q.z, w = q.w;
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}
__device__ __forceinline__ V3 qrot(Q4 q, V3 v) {
  double R[9];
  qrot_matrix(q, R);
  return {R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z};
}
__device__ __forceinline__ Pose compose(Pose a, Pose b) { return {qnormalize(qmul(a.q, b.q)), a.t + qrot(a.q, b.t)}; }
__device__ __forceinline__ Pose inverse(Pose a) {
  const Q4 qi = {-a.q.x, -a.q.y, -a.q.z, a.q.w};
  const V3 r = qrot(qi, a.t);
  return {qi, {-r.x, -r.y, -r.z}};
}
__device__ __forceinline__ Pose load_pose(const double* q, const double* t, int i) {
  return {{q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]}, {t[3 * i], t[3 * i + 1], t[3 * i + 2]}};
}

// Sophus::SO3d::exp / log on quaternions, Sophus::SE3d::exp / log
__device__ __forceinline__ Q4 so3_exp(V3 w) {
  const double th2 = dot(w, w);
  double imag, real;
  if (th2 < kEps * kEps) {
    imag = 0.5 - th2 / 48.0 + th2 * th2 / 3840.0;
    real = 1.0 - th2 / 8.0 + th2 * th2 / 384.0;
  } else {
    const double th = sqrt(th2);
    imag = sin(0.5 * th) / th;
    real = cos(0.5 * th);
  }
  return {imag * w.x, imag * w.y, imag * w.z, real};
}
__device__ __forceinline__ V3 so3_log(Q4 q) {
  const double n2 = q.x * q.x + q.y * q.y + q.z * q.z;
  double f;
  if (n2 < kEps * kEps) {
    f = 2.0 / q.w - (2.0 / 3.0) * n2 / (q.w * q.w * q.w);
  } else {
    const double n = sqrt(n2);
    f = 2.0 * (q.w < 0.0 ? atan2(-n, -q.w) : atan2(n, q.w)) / n;
  }
  return {f * q.x, f * q.y, f * q.z};
}
__device__ __forceinline__ Pose se3_exp(const double* xi) {
  const V3 rho = {xi[0], xi[1], xi[2]}, om = {xi[3], xi[4], xi[5]};
  const double th = sqrt(dot(om, om));
  double a, b;
  if (th < kEps) { a = 0.5; b = 1.0 / 6.0; } else { a = (1.0 - cos(th)) / (th * th); b = (th - sin(th)) / (th * th * th); }
  const V3 c1 = cross(om, rho), c2 = cross(om, c1);
  return {so3_exp(om), rho + a * c1 + b * c2};
}
__device__ __forceinline__ void se3_log(Pose p, double* xi) {
  const V3 om = so3_log(p.q);
  const double th = sqrt(dot(om, om));
  const double c = th < kEps ? 1.0 / 12.0 : (1.0 - th * cos(0.5 * th) / (2.0 * sin(0.5 * th))) / (th * th);
  const V3 c1 = cross(om, p.t), c2 = cross(om, c1);
  const V3 r = p.t - 0.5 * c1 + c * c2;
  xi[0] = r.x; xi[1] = r.y; xi[2] = r.z; xi[3] = om.x; xi[4] = om.y; xi[5] = om.z;
}

__device__ __forceinline__ void hat(V3 v, double* M) {
  M[0] = 0; M[1] = -v.z; M[2] = v.y; M[3] = v.z; M[4] = 0; M[5] = -v.x; M[6] = -v.y; M[7] = v.x; M[8] = 0;
}
__device__ __forceinline__ void mm3(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

// inverse LEFT Jacobian of SE(3), row-major 6 x 6, order [rho, theta] (oracle/pg_oracle.py jl_inv_se3)
__device__ void jl_inv_se3(const double* xi, double* J) {
  const V3 rho = {xi[0], xi[1], xi[2]}, om = {xi[3], xi[4], xi[5]};
  const double th = sqrt(dot(om, om));
  const bool small = th < kSmall;
  const double t = small ? 1.0 : th, s = sin(t), c = cos(t);
  double W[9], P[9], WW[9], A[9], WP[9], PW[9], WPW[9], T1[9], T2[9], Q[9];
  hat(om, W); hat(rho, P);
  mm3(W, W, WW);
  const double ca = small ? 1.0 / 12.0 : 1.0 / (t * t) - (1.0 + c) / (2.0 * t * s);
#pragma unroll
  for (int k = 0; k < 9; ++k) A[k] = ((k % 4 == 0) ? 1.0 : 0.0) - 0.5 * W[k] + ca * WW[k];
  const double q1 = small ? 1.0 / 6.0 : (t - s) / (t * t * t);
  const double q2 = small ? 1.0 / 24.0 : (t * t + 2.0 * c - 2.0) / (2.0 * t * t * t * t);
  const double q3 = small ? 1.0 / 120.0 : (2.0 * t - 3.0 * s + t * c) / (2.0 * t * t * t * t * t);
  mm3(W, P, WP); mm3(P, W, PW); mm3(WP, W, WPW);
  mm3(W, WP, T1);      // W W P
  mm3(PW, W, T2);      // P W W
#pragma unroll
  for (int k = 0; k < 9; ++k) Q[k] = 0.5 * P[k] + q1 * (WP[k] + PW[k] + WPW[k]) + q2 * (T1[k] + T2[k] - 3.0 * WPW[k]);
  mm3(WPW, W, T1);     // W P W W
  mm3(W, WPW, T2);     // W W P W
#pragma unroll
  for (int k = 0; k < 9; ++k) Q[k] += q3 * (T1[k] + T2[k]);
  mm3(A, Q, T1);
  mm3(T1, A, T2);      // A Q A
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      J[6 * i + j] = A[3 * i + j];
      J[6 * (i + 3) + j + 3] = A[3 * i + j];
      J[6 * i + j + 3] = -T2[3 * i + j];
      J[6 * (i + 3) + j] = 0.0;
    }
}

// residual and both Jacobians of one edge (row-major 6 x 6)
__device__ void edge_eval(Pose Ti, Pose Tj, Pose Z, double* r, double* Ji, double* Jj) {
  const Pose rel = compose(inverse(Ti), Tj);
  const Pose E = compose(inverse(Z), rel);
  se3_log(E, r);
  if (!Jj) return;
  double m[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) m[k] = -r[k];
  jl_inv_se3(m, Jj);                                    // J_r^-1(r) = J_l^-1(-r)
  const Pose ji = compose(inverse(Tj), Ti);             // T_j^-1 T_i
  double R[9], tx[9], tR[9];
  qrot_matrix(ji.q, R);
  hat(ji.t, tx);
  mm3(tx, R, tR);
  // Ji = -Jj * Ad,  Ad = [[R, tR], [0, R]]
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        s0 += Jj[6 * a + k] * R[3 * k + b];
        s1 += Jj[6 * a + k] * tR[3 * k + b] + Jj[6 * a + 3 + k] * R[3 * k + b];
      }
      Ji[6 * a + b] = -s0;
      Ji[6 * a + 3 + b] = -s1;
    }
}

// ---- linearisation: one thread per pose c ------------------------------------------------------
// band[c][d][36]: block (c + d, c), row-major (rows = tangent of pose c + d, cols = tangent of pose c)
__global__ void __launch_bounds__(64)
k_pg_linearize(int n, int B, const double* __restrict__ q, const double* __restrict__ t, const int* __restrict__ inc_ptr,
               const int* __restrict__ inc_edge, const int* __restrict__ ei, const int* __restrict__ ej,
               const double* __restrict__ zq, const double* __restrict__ zt, double* __restrict__ band, double* __restrict__ g,
               double* __restrict__ cost_share, const int* __restrict__ edge_cl, double* __restrict__ cl_blk) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double* col = band + (size_t)c * (B + 1) * 36;
  for (int k = 0; k < (B + 1) * 36; ++k) col[k] = 0.0;
  double gc[6] = {0, 0, 0, 0, 0, 0}, cost = 0.0;
  if (c == 0) {                                   // the constant pose: identity block, zero gradient, no coupling
    for (int k = 0; k < 6; ++k) { col[7 * k] = 1.0; g[k] = 0.0; }
  }
  const Pose Tc = load_pose(q, t, c);
  for (int p = inc_ptr[c]; p < inc_ptr[c + 1]; ++p) {
    const int e = inc_edge[p], i = ei[e], j = ej[e];
    const Pose Z = load_pose(zq, zt, e);
    double r[6], Ji[36], Jj[36];
    edge_eval(i == c ? Tc : load_pose(q, t, i), j == c ? Tc : load_pose(q, t, j), Z, r, Ji, Jj);
    if (i == c) {
#pragma unroll
      for (int k = 0; k < 6; ++k) cost += r[k] * r[k];
    }
    if (c == 0) continue;
    const double* Jc = (i == c) ? Ji : Jj;
    for (int a = 0; a < 6; ++a) {
      double s = 0.0;
      for (int k = 0; k < 6; ++k) s += Jc[6 * k + a] * r[k];
      gc[a] += s;
      for (int b = 0; b < 6; ++b) {
        double h = 0.0;
        for (int k = 0; k < 6; ++k) h += Jc[6 * k + a] * Jc[6 * k + b];
        col[6 * a + b] += h;
      }
    }
    if (i == c) {                                  // block (j, c) = Jj^T Ji, owned by the lower-index endpoint
      // inside the band: summed into the band column; a loop closure (j - c > B): its own 6 x 6 slot, written once
      const bool far = j - c > B;
      double* blk = far ? cl_blk + (size_t)edge_cl[e] * 36 : col + (size_t)(j - c) * 36;
      for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) {
          double h = 0.0;
          for (int k = 0; k < 6; ++k) h += Jj[6 * k + a] * Ji[6 * k + b];
          if (far) blk[6 * a + b] = h; else blk[6 * a + b] += h;
        }
    }
  }
  if (c) for (int k = 0; k < 6; ++k) g[6 * c + k] = gc[k];
  cost_share[c] = 0.5 * cost;
}

// Jacobi column scale 1 / (1 + sqrt(H_kk)), computed once at x0.  Its own launch: k_pg_damp reads the scales of the
// NEIGHBOURING poses, so they must all exist before it starts (found by a compute-sanitizer run whose timing
// exposed the race of the first, fused version).
__global__ void k_pg_scale(int n, int B, int jacobi, const double* __restrict__ band, double* __restrict__ scale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const double* col = band + (size_t)c * (B + 1) * 36;
  for (int k = 0; k < 6; ++k) scale[6 * c + k] = (jacobi && c) ? 1.0 / (1.0 + sqrt(col[7 * k])) : 1.0;
}

// LM diagonal, scaled damped band A = S H S + D^2, gs = S g
__global__ void k_pg_damp(int n, int B, int new_diag, double dmin, double dmax, double inv_radius,
                          const double* __restrict__ band, const double* __restrict__ g, const double* __restrict__ scale,
                          double* __restrict__ diag, double* __restrict__ A, double* __restrict__ gs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const double* col = band + (size_t)c * (B + 1) * 36;
  double* out = A + (size_t)c * (B + 1) * 36;
  double sc[6];
  for (int k = 0; k < 6; ++k) sc[k] = scale[6 * c + k];
  for (int d = 0; d <= B; ++d) {
    if (c + d >= n) { for (int k = 0; k < 36; ++k) out[d * 36 + k] = 0.0; continue; }
    for (int a = 0; a < 6; ++a) {
      const double sa = scale[6 * (c + d) + a];
      for (int b = 0; b < 6; ++b) out[d * 36 + 6 * a + b] = sa * col[d * 36 + 6 * a + b] * sc[b];
    }
  }
  for (int k = 0; k < 6; ++k) {
    if (new_diag) diag[6 * c + k] = c ? fmin(fmax(out[7 * k], dmin), dmax) : 0.0;
    out[7 * k] += diag[6 * c + k] * inv_radius;
    gs[6 * c + k] = sc[k] * g[6 * c + k];
  }
}

// 1/sqrt(d): hardware seed + two Newton steps, inline (the library rsqrt() is a CALL with a slow path, which
// forces the register-resident 6 x 6 factor of the pivot thread out to local memory around every pivot)
__device__ __forceinline__ double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-d * y, y, 1.0);
  y = fma(0.5 * y, e, y);
  e = fma(-d * y, y, 1.0);
  return fma(0.5 * y, e, y);
}

// ---- block-banded Cholesky + solve, one CTA -----------------------------------------------------
constexpr int BS_THREADS = 256;
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// A: scaled damped band (overwritten by the factor: slot 0 = inverse of L_cc, slots d >= 1 = L_{c+d,c});
// y: right-hand side in, solution out.  info: first non-positive pivot (1-based scalar index), 0 = ok.
__global__ void __launch_bounds__(BS_THREADS, 1)
k_pg_band_solve(int n, int B, double* __restrict__ A, double* __restrict__ y, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  const int Wn = B + 3, CB = (B + 1) * 36;              // ring size, doubles per block column
  double* W = sm;                                       // W[slot][d][36]
  double* Lc = W + (size_t)Wn * CB;                     // Lc[d][36], d = 1..B: the finished column
  double* Li = Lc + (size_t)(B + 1) * 36;               // inverse of L_cc (lower, row-major)
  double* yw = Li + 36;                                 // rhs window yw[slot][6]
  double* yc = yw + (size_t)Wn * 6;                     // y_c
  const int tid = threadIdx.x;
  auto load_col = [&](int c) {                          // asynchronous: column c of A (+ rhs) into its ring slot
    if (c < n) {
      const int s = c % Wn;
      for (int e = tid; e < CB; e += BS_THREADS) cp_async8(W + (size_t)s * CB + e, A + (size_t)c * CB + e);
      if (tid < 6) cp_async8(yw + s * 6 + tid, y + 6 * (size_t)c + tid);
    }
    cp_async_commit();
  };
  for (int c = 0; c < B + 2; ++c) load_col(c);
  for (int c = 0; c < n; ++c) {
    cp_async_wait<1>();                                 // everything but the newest group has landed
    __syncthreads();
    double* Wc = W + (size_t)(c % Wn) * CB;
    // P1: 6 x 6 Cholesky of the diagonal block and the inverse of its factor (one thread; 6 pivots)
    if (tid == 0) {
      double L[21];                                     // lower triangle in registers (fully unrolled), inverse straight into Li
#define LL(i, j) L[(i) * ((i) + 1) / 2 + (j)]
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) LL(i, j) = Wc[6 * i + j];
#pragma unroll
      for (int k = 0; k < 36; ++k) Li[k] = 0.0;
      bool ok = true;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        double d = LL(j, j);
#pragma unroll
        for (int k = 0; k < j; ++k) d -= LL(j, k) * LL(j, k);
        if (!(d > 0.0) && ok) { ok = false; atomicCAS(info, 0, 6 * c + j + 1); }
        const double inv = fast_rsqrt(d);
        LL(j, j) = inv;                                 // keep 1 / L_jj on the diagonal
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
          double s = LL(i, j);
#pragma unroll
          for (int k = 0; k < j; ++k) s -= LL(i, k) * LL(j, k);
          LL(i, j) = s * inv;
        }
      }
      double X[21];
#pragma unroll
      for (int j = 0; j < 6; ++j) {                     // X = L^-1 by columns
        X[j * (j + 1) / 2 + j] = LL(j, j);
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
          double s = 0.0;
#pragma unroll
          for (int k = j; k < i; ++k) s -= LL(i, k) * X[k * (k + 1) / 2 + j];
          X[i * (i + 1) / 2 + j] = s * LL(i, i);
        }
      }
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) Li[6 * i + j] = X[i * (i + 1) / 2 + j];
#undef LL
    }
    __syncthreads();
    // P2: L_{c+d,c} = A_{c+d,c} L_cc^-T ; y_c = L_cc^-1 b_c
    for (int e = tid; e < 36 * B; e += BS_THREADS) {
      const int d = 1 + e / 36, a = (e % 36) / 6, b = e % 6;
      double s = 0.0;
      for (int k = 0; k <= b; ++k) s += Wc[d * 36 + 6 * a + k] * Li[6 * b + k];
      Lc[d * 36 + 6 * a + b] = s;
    }
    if (tid >= BS_THREADS - 6) {
      const int a = tid - (BS_THREADS - 6);
      double s = 0.0;
      for (int k = 0; k <= a; ++k) s += Li[6 * a + k] * yw[(c % Wn) * 6 + k];
      yc[a] = s;
    }
    __syncthreads();
    // P3: trailing update inside the ring, rhs update, factor column out, next column in
    const int npair = B * (B + 1) / 2;
    for (int e = tid; e < 36 * npair; e += BS_THREADS) {
      int p = e / 36, d2 = 1;
      while (p >= B - d2 + 1) { p -= B - d2 + 1; ++d2; }
      const int d1 = d2 + p, a = (e % 36) / 6, b = e % 6;
      if (c + d1 < n) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += Lc[d1 * 36 + 6 * a + k] * Lc[d2 * 36 + 6 * b + k];
        W[(size_t)((c + d2) % Wn) * CB + (d1 - d2) * 36 + 6 * a + b] -= s;
      }
    }
    for (int e = tid; e < 6 * B; e += BS_THREADS) {
      const int d = 1 + e / 6, a = e % 6;
      if (c + d < n) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += Lc[d * 36 + 6 * a + k] * yc[k];
        yw[((c + d) % Wn) * 6 + a] -= s;
      }
    }
    for (int e = tid; e < CB; e += BS_THREADS) A[(size_t)c * CB + e] = e < 36 ? Li[e] : Lc[e];
    if (tid < 6) y[6 * (size_t)c + tid] = yc[tid];
    __syncthreads();                                    // the slot of column c - 1 is free: nobody reads it any more
    load_col(c + B + 2);
  }
  cp_async_wait<0>();
  __syncthreads();
  // backward substitution: x_c = L_cc^-T (y_c - sum_d L_{c+d,c}^T x_{c+d}); the last B solutions live in yw.
  // The factor columns stream back in through the same ring, two steps ahead.
  auto load_back = [&](int c) {
    if (c >= 0) {
      const int s = c % Wn;
      for (int e = tid; e < CB; e += BS_THREADS) cp_async8(W + (size_t)s * CB + e, A + (size_t)c * CB + e);
    }
    cp_async_commit();
  };
  double* xw = Lc;                                      // x ring: xw[slot][6] (Lc is free now), partial sums behind it
  double* ps = Lc + (size_t)Wn * 6;
  load_back(n - 1);
  load_back(n - 2);
  for (int c = n - 1; c >= 0; --c) {
    cp_async_wait<1>();
    __syncthreads();
    const double* Ac = W + (size_t)(c % Wn) * CB;
    if (tid < 6 * (B + 1)) {                            // partial sums: one thread per (d, a)
      const int d = tid / 6, a = tid % 6;
      double s = 0.0;
      if (d >= 1 && c + d < n) {
        for (int k = 0; k < 6; ++k) s += Ac[d * 36 + 6 * k + a] * xw[((c + d) % Wn) * 6 + k];
      }
      ps[tid] = s;
    }
    __syncthreads();
    if (tid < 6) {
      double s = y[6 * (size_t)c + tid];
      for (int d = 1; d <= B; ++d) s -= ps[d * 6 + tid];
      yc[tid] = s;
    }
    __syncthreads();
    if (tid < 6) {
      double s = 0.0;
      for (int k = tid; k < 6; ++k) s += Ac[6 * k + tid] * yc[k];      // (L^-1)^T
      xw[(c % Wn) * 6 + tid] = s;
      y[6 * (size_t)c + tid] = s;
    }
    load_back(c - 2);                                   // its slot was last read at step c + 1
  }
}

// ---- partitioned band solve (more than one SM) ---------------------------------------------------
// The block columns are cut into P interiors I_p separated by separators S_p of exactly B block columns, so
// that two interiors never couple directly.  With the interiors ordered first the matrix is
// [[A_II, A_IS], [A_SI, A_SS]] with A_II block diagonal, and
//   k_pg_part_rhs      Z_p <- [b_I | A_{I,S_{p-1}} | A_{I,S_p}]                       (right-hand sides + spikes)
//   k_pg_part_factor   banded Cholesky of every interior side by side (one CTA each, the ring kernel above
//                      with 1 + 12 B right-hand-side columns): Z_p <- L_p^-1 Z_p
//   k_pg_part_gram     G_p = Z_p^T Z_p
//   k_pg_part_assemble reduced system over the separators, R = A_SS - sum_p (spike Gram blocks): block banded with
//                      half-bandwidth 2B - 1, in the same band layout
//   k_pg_band_solve    the reduced system, on one SM (P B columns instead of N)
//   k_pg_part_back     x_I = L_p^-T (z_p - W_l x_{S_{p-1}} - W_r x_{S_p}), every interior side by side
// Exact (same factorisation in another elimination order); the serial depth drops from N to N / P + P B.
__device__ __forceinline__ void cp_async8_zfill(void* smem, const void* gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}

__global__ void __launch_bounds__(256)
k_pg_part_rhs(int B, int P, const int* __restrict__ pi0, const int* __restrict__ pi1, const double* __restrict__ A,
              const double* __restrict__ y, double* __restrict__ Z) {
  const int p = blockIdx.x, i0 = pi0[p], i1 = pi1[p], CB = (B + 1) * 36, Rw = 1 + 12 * B;
  const int total = (i1 - i0) * 6 * Rw;
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < total; e += blockDim.x * gridDim.y) {      // grid (P, S): S CTAs share an interior
    const int c = i0 + e / (6 * Rw), a = (e / Rw) % 6, col = e % Rw;
    double v = 0.0;
    if (col == 0) {
      v = y[6 * (size_t)c + a];
    } else if (col < 1 + 6 * B) {
      const int j = col - 1, s = i0 - B + j / 6, k = j % 6;             // left separator column s, block (c, s)[a][k]
      if (p > 0 && c - s <= B) v = A[(size_t)s * CB + (c - s) * 36 + a * 6 + k];
    } else {
      const int j = col - 1 - 6 * B, s = i1 + j / 6, k = j % 6;         // right separator column s, block (s, c)[k][a]
      if (p < P - 1 && s - c <= B) v = A[(size_t)c * CB + (s - c) * 36 + k * 6 + a];
    }
    Z[(size_t)c * 6 * Rw + a * Rw + col] = v;
  }
}

__global__ void __launch_bounds__(BS_THREADS, 1)
k_pg_part_factor(int B, const int* __restrict__ pi0, const int* __restrict__ pi1, double* __restrict__ A, double* __restrict__ Z,
                 int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  const int Wn = B + 3, CB = (B + 1) * 36, Rw = 1 + 12 * B, YB = 6 * Rw;
  double* W = sm;                                       // W[slot][d][36]
  double* Lc = W + (size_t)Wn * CB;                     // Lc[d][36]
  double* Li = Lc + (size_t)(B + 1) * 36;
  double* yw = Li + 36;                                 // rhs ring yw[slot][6][Rw]
  double* yc = yw + (size_t)Wn * YB;                    // y_c [6][Rw]
  const int tid = threadIdx.x, i0 = pi0[blockIdx.x], i1 = pi1[blockIdx.x];
  auto load_col = [&](int c) {
    if (c < i1) {
      const int s = c % Wn;
      // blocks whose row lies outside the interior belong to the spikes, not to A_II: they enter as zeros
      for (int e = tid; e < CB; e += BS_THREADS) cp_async8_zfill(W + (size_t)s * CB + e, A + (size_t)c * CB + e, c + e / 36 < i1);
      for (int e = tid; e < YB; e += BS_THREADS) cp_async8(yw + (size_t)s * YB + e, Z + (size_t)c * YB + e);
    }
    cp_async_commit();
  };
  for (int c = i0; c < i0 + B + 2; ++c) load_col(c);
  for (int c = i0; c < i1; ++c) {
    cp_async_wait<1>();
    __syncthreads();
    double* Wc = W + (size_t)(c % Wn) * CB;
    if (tid == 0) {
      double L[21];
#define LL(i, j) L[(i) * ((i) + 1) / 2 + (j)]
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) LL(i, j) = Wc[6 * i + j];
#pragma unroll
      for (int k = 0; k < 36; ++k) Li[k] = 0.0;
      bool ok = true;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        double d = LL(j, j);
#pragma unroll
        for (int k = 0; k < j; ++k) d -= LL(j, k) * LL(j, k);
        if (!(d > 0.0) && ok) { ok = false; atomicCAS(info, 0, 6 * c + j + 1); }
        const double inv = fast_rsqrt(d);
        LL(j, j) = inv;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
          double s = LL(i, j);
#pragma unroll
          for (int k = 0; k < j; ++k) s -= LL(i, k) * LL(j, k);
          LL(i, j) = s * inv;
        }
      }
      double X[21];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        X[j * (j + 1) / 2 + j] = LL(j, j);
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
          double s = 0.0;
#pragma unroll
          for (int k = j; k < i; ++k) s -= LL(i, k) * X[k * (k + 1) / 2 + j];
          X[i * (i + 1) / 2 + j] = s * LL(i, i);
        }
      }
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) Li[6 * i + j] = X[i * (i + 1) / 2 + j];
#undef LL
    }
    __syncthreads();
    // P2: L_{c+d,c} = A_{c+d,c} L_cc^-T ; y_c = L_cc^-1 Y_c  (6 x Rw)
    for (int e = tid; e < 36 * B; e += BS_THREADS) {
      const int d = 1 + e / 36, a = (e % 36) / 6, b = e % 6;
      double s = 0.0;
      for (int k = 0; k <= b; ++k) s += Wc[d * 36 + 6 * a + k] * Li[6 * b + k];
      Lc[d * 36 + 6 * a + b] = s;
    }
    {
      const double* Yc = yw + (size_t)(c % Wn) * YB;
      for (int e = tid; e < YB; e += BS_THREADS) {
        const int a = e / Rw, col = e % Rw;
        double s = 0.0;
        for (int k = 0; k <= a; ++k) s += Li[6 * a + k] * Yc[k * Rw + col];
        yc[e] = s;
      }
    }
    __syncthreads();
    // P3: trailing update and right-hand-side update inside the interior, factor column and y_c out, next column in
    const int npair = B * (B + 1) / 2;
    for (int e = tid; e < 36 * npair; e += BS_THREADS) {
      int pp = e / 36, d2 = 1;
      while (pp >= B - d2 + 1) { pp -= B - d2 + 1; ++d2; }
      const int d1 = d2 + pp, a = (e % 36) / 6, b = e % 6;
      if (c + d1 < i1) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += Lc[d1 * 36 + 6 * a + k] * Lc[d2 * 36 + 6 * b + k];
        W[(size_t)((c + d2) % Wn) * CB + (d1 - d2) * 36 + 6 * a + b] -= s;
      }
    }
    for (int e = tid; e < B * YB; e += BS_THREADS) {
      const int d = 1 + e / YB, r = e % YB, a = r / Rw, col = r % Rw;
      if (c + d < i1) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += Lc[d * 36 + 6 * a + k] * yc[k * Rw + col];
        yw[(size_t)((c + d) % Wn) * YB + r] -= s;
      }
    }
    for (int e = tid; e < CB; e += BS_THREADS) {
      // rows outside the interior keep their original (spike) blocks in A: only the interior part is overwritten
      if (e < 36) A[(size_t)c * CB + e] = Li[e];
      else if (c + e / 36 < i1) A[(size_t)c * CB + e] = Lc[e];
    }
    for (int e = tid; e < YB; e += BS_THREADS) Z[(size_t)c * YB + e] = yc[e];
    __syncthreads();
    load_col(c + B + 2);
  }
  cp_async_wait<0>();
}

// G_p = Z_p^T Z_p  (Rw x Rw), rows streamed through shared memory in tiles of 32
__global__ void __launch_bounds__(256)
k_pg_part_gram(int B, const int* __restrict__ pi0, const int* __restrict__ pi1, const double* __restrict__ Z, double* __restrict__ G) {
  // grid (P, S): with fewer interiors than SMs the rows of an interior are cut into S = gridDim.y slices, one CTA each; the
  // slices' Gram matrices land in G[(p * S + slice)] and k_pg_gram_sum adds them in slice order (S = 1: G is final)
  extern __shared__ __align__(16) double sm[];          // tile[32][Rw]
  const int p = blockIdx.x, S = gridDim.y, part = blockIdx.y, Rw = 1 + 12 * B, rows_all = 6 * (pi1[p] - pi0[p]);
  const int per = ((rows_all + S - 1) / S + 5) / 6 * 6, r_lo = min(rows_all, part * per), rows = min(rows_all, r_lo + per) - r_lo;
  const double* Zp = Z + (size_t)pi0[p] * 6 * Rw + (size_t)r_lo * Rw;
  const int n_ent = Rw * Rw;
  constexpr int MAXE = 40;                              // entries per thread: Rw^2 / 256 <= 9409 / 256 for B = 8
  double acc[MAXE];
#pragma unroll
  for (int k = 0; k < MAXE; ++k) acc[k] = 0.0;
  for (int r0 = 0; r0 < rows; r0 += 32) {
    const int nr = min(32, rows - r0);
    __syncthreads();
    for (int e = threadIdx.x; e < nr * Rw; e += 256) sm[e] = Zp[(size_t)r0 * Rw + e];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MAXE; ++k) {
      const int e = threadIdx.x + 256 * k;
      if (e < n_ent) {
        const int i = e / Rw, j = e % Rw;
        double s = acc[k];
        for (int r = 0; r < nr; ++r) s = fma(sm[r * Rw + i], sm[r * Rw + j], s);
        acc[k] = s;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < MAXE; ++k) {
    const int e = threadIdx.x + 256 * k;
    if (e < n_ent) G[((size_t)p * S + part) * n_ent + e] = acc[k];
  }
}
__global__ void __launch_bounds__(256) k_pg_gram_sum(int n_ent, int S, const double* __restrict__ Gs, double* __restrict__ G) {
  const int p = blockIdx.x;
  for (int e = threadIdx.x; e < n_ent; e += 256) {
    double t = Gs[((size_t)p * S) * n_ent + e];
    for (int q = 1; q < S; ++q) t += Gs[((size_t)p * S + q) * n_ent + e];
    G[(size_t)p * n_ent + e] = t;
  }
}

// reduced band Rb (block column rc = p B + sl, half-bandwidth Br = 2B - 1) and reduced right-hand side, one CTA per separator
__global__ void __launch_bounds__(256)
k_pg_part_assemble(int B, int P, const int* __restrict__ pi1, const double* __restrict__ A, const double* __restrict__ y,
                   const double* __restrict__ G, double* __restrict__ Rb, double* __restrict__ yR) {
  const int p = blockIdx.x;                             // separator p: between interiors p and p + 1
  const int CB = (B + 1) * 36, Rw = 1 + 12 * B, Br = 2 * B - 1, CBr = (Br + 1) * 36, s0 = pi1[p];
  const double* Gl = G + (size_t)p * Rw * Rw;           // interior p: this separator is its RIGHT one  (columns 1 + 6B ..)
  const double* Gr = G + (size_t)(p + 1) * Rw * Rw;     // interior p + 1: this separator is its LEFT one (columns 1 ..)
  const int oR = 1 + 6 * B, oL = 1;
  // (i) blocks inside the separator: rows (sl2, a), cols (sl1, k), sl2 >= sl1
  for (int e = threadIdx.x; e < B * B * 36; e += 256) {
    const int sl2 = e / (B * 36), sl1 = (e / 36) % B, a = (e % 36) / 6, k = e % 6;
    if (sl2 < sl1) continue;
    const double v = A[(size_t)(s0 + sl1) * CB + (sl2 - sl1) * 36 + a * 6 + k]
                     - Gl[(size_t)(oR + sl2 * 6 + a) * Rw + oR + sl1 * 6 + k] - Gr[(size_t)(oL + sl2 * 6 + a) * Rw + oL + sl1 * 6 + k];
    Rb[(size_t)(p * B + sl1) * CBr + (sl2 - sl1) * 36 + a * 6 + k] = v;
  }
  // (ii) coupling with the previous separator through interior p: rows (p, sl2), cols (p - 1, sl1)
  if (p > 0) {
    for (int e = threadIdx.x; e < B * B * 36; e += 256) {
      const int sl2 = e / (B * 36), sl1 = (e / 36) % B, a = (e % 36) / 6, k = e % 6;
      const double v = -Gl[(size_t)(oR + sl2 * 6 + a) * Rw + oL + sl1 * 6 + k];
      Rb[(size_t)((p - 1) * B + sl1) * CBr + (B + sl2 - sl1) * 36 + a * 6 + k] = v;
    }
  }
  for (int e = threadIdx.x; e < 6 * B; e += 256)
    yR[(size_t)p * 6 * B + e] = y[6 * (size_t)s0 + e] - Gl[(size_t)(oR + e) * Rw] - Gr[(size_t)(oL + e) * Rw];
}

__global__ void __launch_bounds__(BS_THREADS, 1)
k_pg_part_back(int B, int P, const int* __restrict__ pi0, const int* __restrict__ pi1, const double* __restrict__ A,
               const double* __restrict__ Z, const double* __restrict__ yR, double* __restrict__ y) {
  extern __shared__ __align__(16) double sm[];
  const int Wn = B + 3, CB = (B + 1) * 36, Rw = 1 + 12 * B, YB = 6 * Rw;
  double* W = sm;
  double* xw = W + (size_t)Wn * CB;                     // x ring [Wn][6]
  double* ps = xw + (size_t)Wn * 6;                     // partial sums [6 (B + 1)]
  double* yc = ps + 6 * (B + 1);
  double* xs = yc + 6;                                  // [12 B]: solutions of the left and right separators
  const int tid = threadIdx.x, p = blockIdx.x, i0 = pi0[p], i1 = pi1[p];
  for (int e = tid; e < 12 * B; e += BS_THREADS) {
    const bool left = e < 6 * B;
    const int j = left ? e : e - 6 * B;
    double v = 0.0;
    if (left && p > 0) v = yR[(size_t)(p - 1) * 6 * B + j];
    if (!left && p < P - 1) v = yR[(size_t)p * 6 * B + j];
    xs[e] = v;
    if (!left && p < P - 1) y[6 * (size_t)i1 + j] = v;  // this CTA also publishes its right separator's solution
  }
  __syncthreads();
  // t = z - W_l x_l - W_r x_r, written over y (the backward substitution below reads it from there)
  for (int e = tid; e < 6 * (i1 - i0); e += BS_THREADS) {
    const double* z = Z + (size_t)i0 * YB + (size_t)(e / 6) * YB + (e % 6) * Rw;
    double s = z[0];
    for (int j = 0; j < 12 * B; ++j) s = fma(-z[1 + j], xs[j], s);
    y[6 * (size_t)i0 + e] = s;
  }
  __syncthreads();
  auto load_back = [&](int c) {
    if (c >= i0) {
      const int s = c % Wn;
      for (int e = tid; e < CB; e += BS_THREADS) cp_async8(W + (size_t)s * CB + e, A + (size_t)c * CB + e);
    }
    cp_async_commit();
  };
  load_back(i1 - 1);
  load_back(i1 - 2);
  for (int c = i1 - 1; c >= i0; --c) {
    cp_async_wait<1>();
    __syncthreads();
    const double* Ac = W + (size_t)(c % Wn) * CB;
    if (tid < 6 * (B + 1)) {
      const int d = tid / 6, a = tid % 6;
      double s = 0.0;
      if (d >= 1 && c + d < i1) {
        for (int k = 0; k < 6; ++k) s += Ac[d * 36 + 6 * k + a] * xw[((c + d) % Wn) * 6 + k];
      }
      ps[tid] = s;
    }
    __syncthreads();
    if (tid < 6) {
      double s = y[6 * (size_t)c + tid];
      for (int d = 1; d <= B; ++d) s -= ps[d * 6 + tid];
      yc[tid] = s;
    }
    __syncthreads();
    if (tid < 6) {
      double s = 0.0;
      for (int k = tid; k < 6; ++k) s += Ac[6 * k + tid] * yc[k];
      xw[(c % Wn) * 6 + tid] = s;
      y[6 * (size_t)c + tid] = s;
    }
    load_back(c - 2);
  }
  cp_async_wait<0>();
}

// ---- loop closures: the reduced separator system as a DENSE matrix ---------------------------------------------
// Edges longer than the band (|i - j| > B) couple poses that are far apart in the chain.  Their endpoints are put
// into separator groups (B consecutive block columns, so that the interiors in between still never couple), the
// interiors are eliminated exactly as above — any length, including empty — and the reduced system over the groups
// receives (i) the band blocks among separator columns, (ii) the spike Gram blocks of the interiors, (iii) the
// closure blocks.  It is no longer banded: it goes to the dense DAG Cholesky of the BA path (stba_chol.cu).
// R: column-major, lower, leading dimension ld; reduced pose index of group g, slot sl = g B + sl.
__global__ void __launch_bounds__(256)
k_pg_cl_assemble(int B, int P, const int* __restrict__ pi0, const int* __restrict__ pi1, const double* __restrict__ A, const double* __restrict__ y,
                 const double* __restrict__ G, double* __restrict__ R, int ld, double* __restrict__ yR) {
  const int g = blockIdx.x;                             // group g: between interiors g and g + 1
  const int CB = (B + 1) * 36, Rw = 1 + 12 * B, s0 = pi1[g];
  const double* Gl = G + (size_t)g * Rw * Rw;           // interior g: this group is its RIGHT one  (columns 1 + 6B ..)
  const double* Gr = G + (size_t)(g + 1) * Rw * Rw;     // interior g + 1: this group is its LEFT one (columns 1 ..)
  const int oR = 1 + 6 * B, oL = 1;
  for (int e = threadIdx.x; e < B * B * 36; e += 256) {
    const int sl2 = e / (B * 36), sl1 = (e / 36) % B, a = (e % 36) / 6, k = e % 6;
    if (sl2 < sl1 || (sl2 == sl1 && a < k)) continue;     // lower triangle only
    const double v = A[(size_t)(s0 + sl1) * CB + (sl2 - sl1) * 36 + a * 6 + k]
                     - Gl[(size_t)(oR + sl2 * 6 + a) * Rw + oR + sl1 * 6 + k] - Gr[(size_t)(oL + sl2 * 6 + a) * Rw + oL + sl1 * 6 + k];
    R[(size_t)(6 * (g * B + sl1) + k) * ld + 6 * (g * B + sl2) + a] = v;
  }
  if (g > 0) {
    // coupling with the previous group: through interior g (Gram) and, when that interior is shorter than B, directly
    const int sp = pi1[g - 1];
    for (int e = threadIdx.x; e < B * B * 36; e += 256) {
      const int sl2 = e / (B * 36), sl1 = (e / 36) % B, a = (e % 36) / 6, k = e % 6;
      const int c1 = sp + sl1, c2 = s0 + sl2;
      double v = -Gl[(size_t)(oR + sl2 * 6 + a) * Rw + oL + sl1 * 6 + k];
      if (c2 - c1 <= B) v += A[(size_t)c1 * CB + (c2 - c1) * 36 + a * 6 + k];
      R[(size_t)(6 * ((g - 1) * B + sl1) + k) * ld + 6 * (g * B + sl2) + a] = v;
    }
  }
  for (int e = threadIdx.x; e < 6 * B; e += 256)
    yR[(size_t)g * 6 * B + e] = y[6 * (size_t)s0 + e] - Gl[(size_t)(oR + e) * Rw] - Gr[(size_t)(oL + e) * Rw];
}
// closure blocks (j, i), i < j: scaled like the band (A = S H S), added at the endpoints' reduced indices
__global__ void __launch_bounds__(64)
k_pg_cl_add(int n_cl, const int* __restrict__ cl_i, const int* __restrict__ cl_j, const int* __restrict__ sep_of, const double* __restrict__ cl_blk,
            const double* __restrict__ scale, double* __restrict__ R, int ld) {
  const int e = blockIdx.x, t = threadIdx.x;
  if (e >= n_cl || t >= 36) return;
  const int i = cl_i[e], j = cl_j[e], a = t / 6, k = t % 6;
  if (i == 0) return;                                   // the constant pose has no column
  const int ri = sep_of[i], rj = sep_of[j];
  const double v = scale[6 * j + a] * cl_blk[(size_t)e * 36 + t] * scale[6 * i + k];
  atomicAdd(R + (size_t)(6 * ri + k) * ld + 6 * rj + a, v);      // (several closures may join the same pair of poses)
}
__global__ void k_pg_or_info(int* info, const int* other) { if (*other != 0 && *info == 0) *info = *other > 0 ? *other : 1; }

// step: delta = -ys * scale, candidate T+ = T Exp(delta); per-pose shares of |step|^2, |x+|... and of the model cost change
__global__ void k_pg_update(int n, const double* __restrict__ q, const double* __restrict__ t, const double* __restrict__ ys,
                            const double* __restrict__ scale, const double* __restrict__ gs, const double* __restrict__ diag,
                            double inv_radius, double* __restrict__ q2, double* __restrict__ t2, double* __restrict__ share /* n x 3 */) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const Pose T = load_pose(q, t, c);
  Pose Tn = T;
  double mcc = 0.0, step2 = 0.0;
  if (c) {
    double d[6];
    for (int k = 0; k < 6; ++k) {
      const double yk = ys[6 * c + k];
      d[k] = -yk * scale[6 * c + k];
      mcc += yk * (gs[6 * c + k] + diag[6 * c + k] * inv_radius * yk);
    }
    Tn = compose(T, se3_exp(d));
    const double e[7] = {Tn.q.x - T.q.x, Tn.q.y - T.q.y, Tn.q.z - T.q.z, Tn.q.w - T.q.w, Tn.t.x - T.t.x, Tn.t.y - T.t.y, Tn.t.z - T.t.z};
    for (int k = 0; k < 7; ++k) step2 += e[k] * e[k];
  }
  q2[4 * c] = Tn.q.x; q2[4 * c + 1] = Tn.q.y; q2[4 * c + 2] = Tn.q.z; q2[4 * c + 3] = Tn.q.w;
  t2[3 * c] = Tn.t.x; t2[3 * c + 1] = Tn.t.y; t2[3 * c + 2] = Tn.t.z;
  share[3 * c] = 0.5 * mcc;
  share[3 * c + 1] = step2;
  share[3 * c + 2] = c ? (T.q.x * T.q.x + T.q.y * T.q.y + T.q.z * T.q.z + T.q.w * T.q.w + dot(T.t, T.t)) : 0.0;
}

__global__ void k_pg_cost(int64_t m, const double* __restrict__ q, const double* __restrict__ t, const int* __restrict__ ei,
                          const int* __restrict__ ej, const double* __restrict__ zq, const double* __restrict__ zt,
                          double* __restrict__ share) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= m) return;
  double r[6];
  edge_eval(load_pose(q, t, ei[e]), load_pose(q, t, ej[e]), load_pose(zq, zt, (int)e), r, nullptr, nullptr);
  double s = 0.0;
  for (int k = 0; k < 6; ++k) s += r[k] * r[k];
  share[e] = 0.5 * s;
}

// |x - Plus(x, -g)| in ambient coordinates: share[2c] = sum of squares, share[2c+1] = max abs
__global__ void k_pg_gradnorm(int n, const double* __restrict__ q, const double* __restrict__ t, const double* __restrict__ g,
                              double* __restrict__ share) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double s2 = 0.0, mx = 0.0;
  if (c) {
    const Pose T = load_pose(q, t, c);
    double d[6];
    for (int k = 0; k < 6; ++k) d[k] = -g[6 * c + k];
    const Pose Tn = compose(T, se3_exp(d));
    const double e[7] = {T.q.x - Tn.q.x, T.q.y - Tn.q.y, T.q.z - Tn.q.z, T.q.w - Tn.q.w, T.t.x - Tn.t.x, T.t.y - Tn.t.y, T.t.z - Tn.t.z};
    for (int k = 0; k < 7; ++k) { s2 += e[k] * e[k]; mx = fmax(mx, fabs(e[k])); }
  }
  share[2 * c] = s2;
  share[2 * c + 1] = mx;
}

// out[k] = sum (or max when k >= first_max) over i of v[i * stride + k], one block, fixed order
__global__ void __launch_bounds__(256) k_pg_sum(int64_t n, int stride, int first_max, const double* __restrict__ v, double* __restrict__ out) {
  __shared__ double s[256];
  for (int k = 0; k < stride; ++k) {
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) a = (k >= first_max) ? fmax(a, v[i * stride + k]) : a + v[i * stride + k];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
      if ((int)threadIdx.x < w) s[threadIdx.x] = (k >= first_max) ? fmax(s[threadIdx.x], s[threadIdx.x + w]) : s[threadIdx.x] + s[threadIdx.x + w];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[k] = s[0];
    __syncthreads();
  }
}

}  // namespace

struct stba_pg {
  int device = 0, n = 0, B = 1;
  int64_t m = 0, launches = 0;
  cudaStream_t s = nullptr;
  double *q = nullptr, *t = nullptr, *q2 = nullptr, *t2 = nullptr, *zq = nullptr, *zt = nullptr;
  int *ei = nullptr, *ej = nullptr, *inc_ptr = nullptr, *inc_edge = nullptr, *info = nullptr;
  double *band = nullptr, *A = nullptr, *g = nullptr, *gs = nullptr, *ys = nullptr, *scale = nullptr, *diag = nullptr;
  double *share = nullptr, *red = nullptr, *red_host = nullptr;
  double *save_q = nullptr, *save_t = nullptr;
  bool have_scale = false;
  // partitioned band solve (P > 1): interiors [pi0, pi1), separators of B columns in between
  int P = 1, Br = 1;
  int *pi0 = nullptr, *pi1 = nullptr;
  double *Z = nullptr, *G = nullptr, *Rb = nullptr, *yR = nullptr, *Gs = nullptr;
  int gram_split = 1;        // row slices per interior of the spike Gram kernel (fewer interiors than SMs)
  // loop closures (edges longer than the band): endpoints live in separator groups, the reduced system is dense
  bool closure = false;
  int n_cl = 0, n_red = 0, ld_red = 0;
  int *edge_cl = nullptr, *cl_i = nullptr, *cl_j = nullptr, *sep_of = nullptr, *info2 = nullptr;
  double *cl_blk = nullptr, *R = nullptr;
  stba::CholWorkspace chol;
  std::vector<void*> allocs;
  template <typename T>
  int alloc(T** p, size_t count) {
    void* x = nullptr;      // stream-ordered, from the kept default pool: creating a graph per call costs no cudaMalloc / cudaFree stalls
    CK(cudaMallocAsync(&x, (std::max<size_t>(count, 1) * sizeof(T) + 255) & ~(size_t)255, s));
    allocs.push_back(x);
    *p = static_cast<T*>(x);
    return STBA_OK;
  }
  ~stba_pg() {
    chol.reset();
    if (s) {
      for (void* p : allocs) cudaFreeAsync(p, s);
      cudaStreamSynchronize(s);
    }
    stba::pinned_blocks().put(red_host);
    if (s) cudaStreamDestroy(s);
  }
  static int band_smem(int b) { return (int)(((size_t)(b + 3) * (b + 1) * 36 + (size_t)(b + 1) * 36 + 36 + (size_t)(b + 3) * 6 + 6) * sizeof(double)); }
  int smem_bytes() const { return band_smem(std::max(B, Br)); }
  int factor_smem() const { const int Rw = 1 + 12 * B; return (int)(((size_t)(B + 3) * (B + 1) * 36 + (size_t)(B + 1) * 36 + 36 + (size_t)(B + 4) * 6 * Rw) * sizeof(double)); }
  int back_smem() const { return (int)(((size_t)(B + 3) * (B + 1) * 36 + (size_t)(B + 3) * 6 + 6 * (B + 1) + 6 + 12 * B) * sizeof(double)); }

  // solve A ys = ys in place (A = scaled damped band); info <- first non-positive pivot
  int band_solve() {
    if (P <= 1) {
      k_pg_band_solve<<<1, BS_THREADS, band_smem(B), s>>>(n, B, A, ys, info);
      ++launches;
      return STBA_OK;
    }
    const int Rw = 1 + 12 * B;
    k_pg_part_rhs<<<dim3(P, gram_split), 256, 0, s>>>(B, P, pi0, pi1, A, ys, Z);
    k_pg_part_factor<<<P, BS_THREADS, factor_smem(), s>>>(B, pi0, pi1, A, Z, info);
    if (gram_split > 1) {
      k_pg_part_gram<<<dim3(P, gram_split), 256, 32 * Rw * (int)sizeof(double), s>>>(B, pi0, pi1, Z, Gs);
      k_pg_gram_sum<<<P, 256, 0, s>>>(Rw * Rw, gram_split, Gs, G);
      ++launches;
    } else {
      k_pg_part_gram<<<P, 256, 32 * Rw * (int)sizeof(double), s>>>(B, pi0, pi1, Z, G);
    }
    if (closure) {
      CK(cudaMemsetAsync(R, 0, (size_t)ld_red * n_red * sizeof(double), s));
      CK(cudaMemsetAsync(info2, 0, sizeof(int), s));
      k_pg_cl_assemble<<<P - 1, 256, 0, s>>>(B, P, pi0, pi1, A, ys, G, R, ld_red, yR);
      if (n_cl) k_pg_cl_add<<<n_cl, 64, 0, s>>>(n_cl, cl_i, cl_j, sep_of, cl_blk, scale, R, ld_red);
      int nl = 0;
      const int r = stba::chol_factor_solve(chol, R, n_red, ld_red, yR, info2, s, &nl);
      if (r != STBA_OK) return r;
      k_pg_or_info<<<1, 1, 0, s>>>(info, info2);
      k_pg_part_back<<<P, BS_THREADS, back_smem(), s>>>(B, P, pi0, pi1, A, Z, yR, ys);
      launches += 7 + nl;
      return STBA_OK;
    }
    CK(cudaMemsetAsync(Rb, 0, (size_t)(P - 1) * B * (Br + 1) * 36 * sizeof(double), s));
    k_pg_part_assemble<<<P - 1, 256, 0, s>>>(B, P, pi1, A, ys, G, Rb, yR);
    k_pg_band_solve<<<1, BS_THREADS, band_smem(Br), s>>>((P - 1) * B, Br, Rb, yR, info);
    k_pg_part_back<<<P, BS_THREADS, back_smem(), s>>>(B, P, pi0, pi1, A, Z, yR, ys);
    launches += 6;
    return STBA_OK;
  }

  int linearize(double* cost) {
    k_pg_linearize<<<(n + 63) / 64, 64, 0, s>>>(n, B, q, t, inc_ptr, inc_edge, ei, ej, zq, zt, band, g, share, edge_cl, cl_blk);
    k_pg_sum<<<1, 256, 0, s>>>(n, 1, 1, share, red);
    launches += 2;
    CK(cudaMemcpyAsync(red_host, red, sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    *cost = red_host[0];
    return STBA_OK;
  }
  int grad_norms(double* g2, double* gmax) {
    k_pg_gradnorm<<<(n + 127) / 128, 128, 0, s>>>(n, q, t, g, share);
    k_pg_sum<<<1, 256, 0, s>>>(n, 2, 1, share, red);
    launches += 2;
    CK(cudaMemcpyAsync(red_host, red, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *g2 = red_host[0]; *gmax = red_host[1];
    return STBA_OK;
  }
};

extern "C" {

int stba_pg_create(stba_pg** out, int device, int32_t n_poses, int64_t n_edges, const double* q, const double* t,
                   const int32_t* ei, const int32_t* ej, const double* zq, const double* zt) {
  if (!out || n_poses < 1 || n_edges < 0 || !q || !t || (n_edges && (!ei || !ej || !zq || !zt))) return STBA_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int B = 1;
  for (int64_t e = 0; e < n_edges; ++e) {
    if (ei[e] < 0 || ej[e] >= n_poses || ei[e] >= ej[e]) return STBA_ERR_INVALID_ARGUMENT;     // i < j
    B = std::max(B, ej[e] - ei[e]);
  }
  // Loop closures: edges longer than the band.  The band is the largest offset <= 8 that at least 1 % of the poses use
  // (the odometry / local-window edges); every longer edge is a closure whose endpoints become separator columns of
  // the partitioned solve, and the reduced system is solved densely (k_pg_cl_*).
  bool closure = false;
  if (B > kMaxBand || getenv("STBA_PG_CLOSURE_BAND")) {
    std::vector<int64_t> hist(kMaxClosureBand + 1, 0);
    for (int64_t e = 0; e < n_edges; ++e) if (ej[e] - ei[e] <= kMaxClosureBand) ++hist[ej[e] - ei[e]];
    int Bc = 1;
    for (int d = 1; d <= kMaxClosureBand; ++d) if (hist[d] >= std::max<int64_t>(1, n_poses / 100)) Bc = d;
    if (const char* ov = getenv("STBA_PG_CLOSURE_BAND")) Bc = std::max(1, std::min(kMaxClosureBand, atoi(ov)));     // tests
    closure = B > Bc;
    if (closure) B = Bc;
    if (B > kMaxBand) return STBA_ERR_UNSUPPORTED;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return STBA_ERR_NO_DEVICE; }
  if (device < 0 || device >= ndev) return STBA_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(device));
  stba_pg* h = new (std::nothrow) stba_pg();
  if (!h) return STBA_ERR_CUDA;
  h->device = device; h->n = n_poses; h->m = n_edges; h->B = B; h->closure = closure;
  // pose -> incident edges (host index logic; edges in their given order: the summation order of the gather)
  std::vector<int> ptr((size_t)n_poses + 1, 0), inc(2 * (size_t)n_edges);
  for (int64_t e = 0; e < n_edges; ++e) { ++ptr[ei[e] + 1]; ++ptr[ej[e] + 1]; }
  for (int i = 0; i < n_poses; ++i) ptr[i + 1] += ptr[i];
  {
    std::vector<int> cur(ptr.begin(), ptr.end() - 1);
    for (int64_t e = 0; e < n_edges; ++e) { inc[cur[ei[e]]++] = (int)e; inc[cur[ej[e]]++] = (int)e; }
  }
  auto fail = [&](int r) { delete h; return r; };
#define CKH(x) do { int r_ = (x); if (r_ != STBA_OK) return fail(r_); } while (0)
#define CKD(x) do { if ((x) != cudaSuccess) return fail(STBA_ERR_CUDA); } while (0)
  stba::keep_default_pool(device);
  CKD(cudaStreamCreateWithFlags(&h->s, cudaStreamNonBlocking));
  // (with loop closures the last separator group may reach past the last pose: B dummy identity columns behind it)
  const size_t N = (size_t)n_poses + (closure ? B : 0), M = std::max<int64_t>(n_edges, 1), CBn = (size_t)(B + 1) * 36;
  CKH(h->alloc(&h->q, 4 * N)); CKH(h->alloc(&h->t, 3 * N)); CKH(h->alloc(&h->q2, 4 * N)); CKH(h->alloc(&h->t2, 3 * N));
  CKH(h->alloc(&h->zq, 4 * M)); CKH(h->alloc(&h->zt, 3 * M)); CKH(h->alloc(&h->ei, M)); CKH(h->alloc(&h->ej, M));
  CKH(h->alloc(&h->inc_ptr, N + 1)); CKH(h->alloc(&h->inc_edge, 2 * M)); CKH(h->alloc(&h->info, 1));
  CKH(h->alloc(&h->band, N * CBn)); CKH(h->alloc(&h->A, N * CBn)); CKH(h->alloc(&h->g, 6 * N)); CKH(h->alloc(&h->gs, 6 * N));
  CKH(h->alloc(&h->ys, 6 * N)); CKH(h->alloc(&h->scale, 6 * N)); CKH(h->alloc(&h->diag, 6 * N));
  CKH(h->alloc(&h->share, std::max(3 * N, (size_t)M))); CKH(h->alloc(&h->red, 8));
  h->red_host = stba::pinned_blocks().get();
  if (!h->red_host) return fail(STBA_ERR_CUDA);
  CKD(cudaMemsetAsync(h->red, 0, 8 * sizeof(double), h->s));
  CKD(cudaMemsetAsync(h->ys, 0, 6 * N * sizeof(double), h->s));
  CKD(cudaMemsetAsync(h->gs, 0, 6 * N * sizeof(double), h->s));
  CKD(cudaMemsetAsync(h->diag, 0, 6 * N * sizeof(double), h->s));
  CKD(cudaMemsetAsync(h->scale, 0, 6 * N * sizeof(double), h->s));
  CKD(cudaMemcpyAsync(h->q, q, 4 * (size_t)n_poses * sizeof(double), cudaMemcpyHostToDevice, h->s));
  CKD(cudaMemcpyAsync(h->t, t, 3 * (size_t)n_poses * sizeof(double), cudaMemcpyHostToDevice, h->s));
  if (n_edges) {
    CKD(cudaMemcpyAsync(h->zq, zq, 4 * (size_t)n_edges * sizeof(double), cudaMemcpyHostToDevice, h->s));
    CKD(cudaMemcpyAsync(h->zt, zt, 3 * (size_t)n_edges * sizeof(double), cudaMemcpyHostToDevice, h->s));
    CKD(cudaMemcpyAsync(h->ei, ei, (size_t)n_edges * sizeof(int), cudaMemcpyHostToDevice, h->s));
    CKD(cudaMemcpyAsync(h->ej, ej, (size_t)n_edges * sizeof(int), cudaMemcpyHostToDevice, h->s));
    CKD(cudaMemcpyAsync(h->inc_edge, inc.data(), inc.size() * sizeof(int), cudaMemcpyHostToDevice, h->s));
  }
  CKD(cudaMemcpyAsync(h->inc_ptr, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice, h->s));
  if (closure) {
    // ---- loop closures: separator groups at the closure endpoints (and every L columns), interiors of any length ----
    std::vector<int> h_edge_cl((size_t)M, -1), h_cl_i, h_cl_j;
    std::vector<uint8_t> mand((size_t)n_poses, 0);
    for (int64_t e = 0; e < n_edges; ++e)
      if (ej[e] - ei[e] > B) {
        h_edge_cl[e] = (int)h_cl_i.size();
        h_cl_i.push_back(ei[e]); h_cl_j.push_back(ej[e]);
        mand[ei[e]] = 1; mand[ej[e]] = 1;
      }
    h->n_cl = (int)h_cl_i.size();
    const int L = std::max(std::max(B, 8), (int)std::lround(std::sqrt((double)n_poses * B)));      // interior length without closures nearby
    std::vector<int> i0, i1, sep((size_t)N, -1);
    int cur = 0, next_m = 0;
    for (;;) {
      while (next_m < n_poses && (next_m < cur || !mand[next_m])) ++next_m;      // first closure endpoint at or after cur
      const int cut = std::min(next_m, cur + L);
      if (cut >= n_poses) { i0.push_back(cur); i1.push_back(std::max(cur, n_poses)); break; }
      i0.push_back(cur); i1.push_back(cut);
      const int g = (int)i1.size() - 1;
      for (int sl = 0; sl < B; ++sl) sep[cut + sl] = g * B + sl;
      cur = cut + B;
      if (cur >= n_poses) { i0.push_back(cur); i1.push_back(cur); break; }
    }
    // (a final interior that starts inside the padding is empty by construction; one that ends at n_poses is real)
    if (i1.back() > n_poses) i1.back() = i0.back();
    const int P = (int)i0.size();
    h->P = P; h->Br = 1;
    h->n_red = 6 * B * (P - 1);
    h->ld_red = h->n_red + 2;
    const size_t Rw = 1 + 12 * (size_t)B;
    CKH(h->alloc(&h->pi0, P)); CKH(h->alloc(&h->pi1, P));
    CKH(h->alloc(&h->Z, N * 6 * Rw)); CKH(h->alloc(&h->G, (size_t)P * Rw * Rw));
    CKH(h->alloc(&h->yR, (size_t)std::max(h->n_red, 1)));
    CKH(h->alloc(&h->R, (size_t)h->ld_red * std::max(h->n_red, 1)));
    CKH(h->alloc(&h->edge_cl, M)); CKH(h->alloc(&h->cl_i, std::max(h->n_cl, 1))); CKH(h->alloc(&h->cl_j, std::max(h->n_cl, 1)));
    CKH(h->alloc(&h->cl_blk, 36 * (size_t)std::max(h->n_cl, 1))); CKH(h->alloc(&h->sep_of, N)); CKH(h->alloc(&h->info2, 1));
    CKD(cudaMemsetAsync(h->cl_blk, 0, 36 * (size_t)std::max(h->n_cl, 1) * sizeof(double), h->s));
    CKD(cudaMemsetAsync(h->Z, 0, N * 6 * Rw * sizeof(double), h->s));
    CKD(cudaMemcpyAsync(h->pi0, i0.data(), P * sizeof(int), cudaMemcpyHostToDevice, h->s));
    CKD(cudaMemcpyAsync(h->pi1, i1.data(), P * sizeof(int), cudaMemcpyHostToDevice, h->s));
    CKD(cudaMemcpyAsync(h->edge_cl, h_edge_cl.data(), (size_t)M * sizeof(int), cudaMemcpyHostToDevice, h->s));
    CKD(cudaMemcpyAsync(h->cl_i, h_cl_i.data(), h_cl_i.size() * sizeof(int), cudaMemcpyHostToDevice, h->s));
    CKD(cudaMemcpyAsync(h->cl_j, h_cl_j.data(), h_cl_j.size() * sizeof(int), cudaMemcpyHostToDevice, h->s));
    CKD(cudaMemcpyAsync(h->sep_of, sep.data(), N * sizeof(int), cudaMemcpyHostToDevice, h->s));
    // the dummy columns behind the last pose: identity diagonal blocks, no coupling, zero right-hand side
    {
      std::vector<double> pad((size_t)B * CBn, 0.0);
      for (int c = 0; c < B; ++c) for (int k = 0; k < 6; ++k) pad[(size_t)c * CBn + 7 * k] = 1.0;
      CKD(cudaMemcpyAsync(h->A + (size_t)n_poses * CBn, pad.data(), pad.size() * sizeof(double), cudaMemcpyHostToDevice, h->s));
      CKD(cudaMemcpyAsync(h->band + (size_t)n_poses * CBn, pad.data(), pad.size() * sizeof(double), cudaMemcpyHostToDevice, h->s));
      CKD(cudaStreamSynchronize(h->s));       // host vectors above are locals
    }
    CKD(cudaFuncSetAttribute(k_pg_part_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, h->factor_smem()));
    CKD(cudaFuncSetAttribute(k_pg_part_back, cudaFuncAttributeMaxDynamicSharedMemorySize, h->back_smem()));
    CKD(cudaFuncSetAttribute(k_pg_part_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * (int)Rw * (int)sizeof(double)));
  } else
  {
    // partitions: serial depth N / P + P B is smallest near P = sqrt(N / B); every interior must be at least B
    // columns long (so that interiors never couple) and the reduced half-bandwidth 2B - 1 must fit the ring
    int P = 1;
    if (!getenv("STBA_PG_SERIAL") && 2 * B - 1 <= kMaxBand && n_poses >= 64 * B) {
      int sms = 0;                                   // (cudaGetDeviceProperties costs ~1 ms per call)
      CKD(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
      P = (int)std::lround(std::sqrt((double)n_poses / B));
      P = std::max(1, std::min(P, sms));
    }
    if (const char* ov = getenv("STBA_PG_PARTS")) P = (2 * B - 1 <= kMaxBand) ? std::max(1, atoi(ov)) : 1;    // tests: force a partition count
    while (P > 1 && (n_poses - (P - 1) * B) / P < std::max(B, 8)) --P;
    h->P = P;
    h->Br = P > 1 ? 2 * B - 1 : 1;
    if (P > 1) {
      std::vector<int> i0(P), i1(P);
      const int m = n_poses - (P - 1) * B, base = m / P, rem = m % P;
      int c = 0;
      for (int p = 0; p < P; ++p) { i0[p] = c; c += base + (p < rem ? 1 : 0); i1[p] = c; if (p < P - 1) c += B; }
      const size_t Rw = 1 + 12 * (size_t)B;
      CKH(h->alloc(&h->pi0, P)); CKH(h->alloc(&h->pi1, P));
      CKH(h->alloc(&h->Z, N * 6 * Rw)); CKH(h->alloc(&h->G, (size_t)P * Rw * Rw));
      {
        int sms = 0;
        CKD(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        h->gram_split = std::max(1, std::min(4, (sms + P / 2) / P));
        if (h->gram_split > 1) CKH(h->alloc(&h->Gs, (size_t)P * h->gram_split * Rw * Rw));
      }
      CKH(h->alloc(&h->Rb, (size_t)(P - 1) * B * (h->Br + 1) * 36)); CKH(h->alloc(&h->yR, (size_t)(P - 1) * B * 6));
      CKD(cudaMemcpyAsync(h->pi0, i0.data(), P * sizeof(int), cudaMemcpyHostToDevice, h->s));
      CKD(cudaMemcpyAsync(h->pi1, i1.data(), P * sizeof(int), cudaMemcpyHostToDevice, h->s));
      CKD(cudaStreamSynchronize(h->s));       // i0 / i1 are stack vectors
      CKD(cudaFuncSetAttribute(k_pg_part_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, h->factor_smem()));
      CKD(cudaFuncSetAttribute(k_pg_part_back, cudaFuncAttributeMaxDynamicSharedMemorySize, h->back_smem()));
      CKD(cudaFuncSetAttribute(k_pg_part_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * (int)Rw * (int)sizeof(double)));
    }
  }
  CKD(cudaFuncSetAttribute(k_pg_band_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes()));
  CKD(cudaStreamSynchronize(h->s));
#undef CKH
#undef CKD
  *out = h;
  return STBA_OK;
}

void stba_pg_destroy(stba_pg* pg) {
  if (!pg) return;
  cudaSetDevice(pg->device);
  delete pg;
}

int stba_pg_get_state(stba_pg* pg, double* q, double* t) {
  if (!pg || !q || !t) return STBA_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(pg->device));
  CK(cudaMemcpyAsync(q, pg->q, 4 * (size_t)pg->n * sizeof(double), cudaMemcpyDeviceToHost, pg->s));
  CK(cudaMemcpyAsync(t, pg->t, 3 * (size_t)pg->n * sizeof(double), cudaMemcpyDeviceToHost, pg->s));
  CK(cudaStreamSynchronize(pg->s));
  return STBA_OK;
}

int stba_pg_set_state(stba_pg* pg, const double* q, const double* t) {
  if (!pg || !q || !t) return STBA_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(pg->device));
  CK(cudaMemcpyAsync(pg->q, q, 4 * (size_t)pg->n * sizeof(double), cudaMemcpyHostToDevice, pg->s));
  CK(cudaMemcpyAsync(pg->t, t, 3 * (size_t)pg->n * sizeof(double), cudaMemcpyHostToDevice, pg->s));
  CK(cudaStreamSynchronize(pg->s));
  return STBA_OK;
}

// device-side snapshot of the poses (benchmarks restore x0 without touching the host)
int stba_pg_save_state(stba_pg* pg) {
  if (!pg) return STBA_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(pg->device));
  if (!pg->save_q) { int r = pg->alloc(&pg->save_q, 4 * (size_t)pg->n); if (r != STBA_OK) return r; r = pg->alloc(&pg->save_t, 3 * (size_t)pg->n); if (r != STBA_OK) return r; }
  CK(cudaMemcpyAsync(pg->save_q, pg->q, 4 * (size_t)pg->n * sizeof(double), cudaMemcpyDeviceToDevice, pg->s));
  CK(cudaMemcpyAsync(pg->save_t, pg->t, 3 * (size_t)pg->n * sizeof(double), cudaMemcpyDeviceToDevice, pg->s));
  CK(cudaStreamSynchronize(pg->s));
  return STBA_OK;
}
int stba_pg_restore_state(stba_pg* pg) {
  if (!pg || !pg->save_q) return STBA_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(pg->device));
  CK(cudaMemcpyAsync(pg->q, pg->save_q, 4 * (size_t)pg->n * sizeof(double), cudaMemcpyDeviceToDevice, pg->s));
  CK(cudaMemcpyAsync(pg->t, pg->save_t, 3 * (size_t)pg->n * sizeof(double), cudaMemcpyDeviceToDevice, pg->s));
  return STBA_OK;
}

// device time of `reps` launches of the linearisation kernel alone (CUDA events on the solver's stream)
int stba_pg_time_linearize(stba_pg* pg, int reps, float* ms) {
  if (!pg || reps < 1 || !ms) return STBA_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(pg->device));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a, pg->s));
    k_pg_linearize<<<(pg->n + 63) / 64, 64, 0, pg->s>>>(pg->n, pg->B, pg->q, pg->t, pg->inc_ptr, pg->inc_edge, pg->ei, pg->ej, pg->zq, pg->zt, pg->band,
                                                     pg->g, pg->share, pg->edge_cl, pg->cl_blk);
    CK(cudaEventRecord(b, pg->s));
    CK(cudaStreamSynchronize(pg->s));
    CK(cudaEventElapsedTime(&ms[r], a, b));
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  return STBA_OK;
}

int stba_pg_linearize(stba_pg* pg, double* cost, double* g, double* Hdiag, int32_t* bandwidth) {
  if (!pg) return STBA_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(pg->device));
  double c = 0.0;
  const int r = pg->linearize(&c);
  if (r != STBA_OK) return r;
  if (cost) *cost = c;
  if (bandwidth) *bandwidth = pg->B;
  if (g) CK(cudaMemcpy(g, pg->g, 6 * (size_t)pg->n * sizeof(double), cudaMemcpyDeviceToHost));
  if (Hdiag) CK(cudaMemcpy2D(Hdiag, 36 * sizeof(double), pg->band, (size_t)(pg->B + 1) * 36 * sizeof(double), 36 * sizeof(double), pg->n,
                             cudaMemcpyDeviceToHost));
  return STBA_OK;
}

int stba_pg_solve(stba_pg* pg, const stba_options* opt, stba_summary* sum, stba_iteration_callback cb, void* user) {
  if (!pg) return STBA_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(pg->device));
  stba_options o;
  if (opt) o = *opt; else stba_options_init(&o);
  const int n = pg->n, B = pg->B;
  pg->have_scale = false;      // Jacobi scaling is computed at the x0 of THIS solve (as Ceres does)
  const auto t_start = std::chrono::steady_clock::now();
  const int64_t launches0 = pg->launches;
  auto record = [&](const stba_iteration& it) -> int {
    if (sum && sum->iterations && sum->num_iterations < sum->iterations_capacity) sum->iterations[sum->num_iterations] = it;
    if (sum) ++sum->num_iterations;
    return cb ? cb(&it, user) : STBA_SOLVER_CONTINUE;
  };
  if (sum) { stba_iteration* keep = sum->iterations; const int cap = sum->iterations_capacity; memset(sum, 0, sizeof(*sum)); sum->iterations = keep; sum->iterations_capacity = cap; }
  int term = STBA_NO_CONVERGENCE, n_ok = 0, n_bad = 0;
  const char* msg = "";
  // ---- IterationZero
  double x_cost = 0.0, g2 = 0.0, gmax = 0.0;
  int r = pg->linearize(&x_cost);
  if (r != STBA_OK) return r;
  if ((r = pg->grad_norms(&g2, &gmax)) != STBA_OK) return r;
  // |x| over the non-constant poses
  k_pg_update<<<(n + 127) / 128, 128, 0, pg->s>>>(n, pg->q, pg->t, pg->ys, pg->scale, pg->gs, pg->diag, 0.0, pg->q2, pg->t2, pg->share);
  k_pg_sum<<<1, 256, 0, pg->s>>>(n, 3, 3, pg->share, pg->red);
  pg->launches += 2;
  CK(cudaMemcpyAsync(pg->red_host, pg->red, 3 * sizeof(double), cudaMemcpyDeviceToHost, pg->s));
  CK(cudaStreamSynchronize(pg->s));
  double x_norm = std::sqrt(pg->red_host[2]);
  const double initial_cost = x_cost;
  double radius = o.initial_trust_region_radius, dec = 2.0;
  bool reuse = false;
  int n_invalid = 0;
  stba_iteration it;
  memset(&it, 0, sizeof(it));
  it.cost = x_cost; it.gradient_norm = std::sqrt(g2); it.gradient_max_norm = gmax; it.step_is_valid = 1; it.step_is_successful = 1;
  while (true) {
    if (it.step_is_successful) ++n_ok; else ++n_bad;
    it.trust_region_radius = radius;
    const int cbr = record(it);
    if (cbr == STBA_SOLVER_ABORT) { term = STBA_USER_FAILURE; msg = "User callback returned abort."; break; }
    if (cbr == STBA_SOLVER_TERMINATE_SUCCESSFULLY) { term = STBA_USER_SUCCESS; msg = "User callback returned terminate successfully."; break; }
    if (it.iteration >= o.max_num_iterations) { term = STBA_NO_CONVERGENCE; msg = "Maximum number of iterations reached."; break; }
    if (it.step_is_successful && it.gradient_max_norm <= o.gradient_tolerance) { term = STBA_CONVERGENCE; msg = "Gradient tolerance reached."; break; }
    if (radius < o.min_trust_region_radius) { term = STBA_CONVERGENCE; msg = "Minimum trust region radius reached."; break; }
    stba_iteration nx;
    memset(&nx, 0, sizeof(nx));
    nx.iteration = it.iteration + 1; nx.cost = x_cost; nx.gradient_norm = it.gradient_norm; nx.gradient_max_norm = it.gradient_max_norm;
    nx.trust_region_radius = radius;
    it = nx;
    // ---- ComputeTrustRegionStep
    if (!pg->have_scale) {
      k_pg_scale<<<(n + 127) / 128, 128, 0, pg->s>>>(n, B, o.jacobi_scaling, pg->band, pg->scale);
      ++pg->launches;
      pg->have_scale = true;
    }
    k_pg_damp<<<(n + 127) / 128, 128, 0, pg->s>>>(n, B, reuse ? 0 : 1, o.min_lm_diagonal, o.max_lm_diagonal, 1.0 / radius, pg->band, pg->g,
                                                   pg->scale, pg->diag, pg->A, pg->gs);
    CK(cudaMemcpyAsync(pg->ys, pg->gs, 6 * (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, pg->s));
    CK(cudaMemsetAsync(pg->info, 0, sizeof(int), pg->s));
    if ((r = pg->band_solve()) != STBA_OK) return r;
    k_pg_update<<<(n + 127) / 128, 128, 0, pg->s>>>(n, pg->q, pg->t, pg->ys, pg->scale, pg->gs, pg->diag, 1.0 / radius, pg->q2, pg->t2, pg->share);
    k_pg_sum<<<1, 256, 0, pg->s>>>(n, 3, 3, pg->share, pg->red);
    if (pg->m) k_pg_cost<<<(unsigned)((pg->m + 127) / 128), 128, 0, pg->s>>>(pg->m, pg->q2, pg->t2, pg->ei, pg->ej, pg->zq, pg->zt, pg->share);
    k_pg_sum<<<1, 256, 0, pg->s>>>(pg->m, 1, 1, pg->share, pg->red + 4);
    pg->launches += 5;
    int info_h = 0;
    CK(cudaMemcpyAsync(pg->red_host, pg->red, 5 * sizeof(double), cudaMemcpyDeviceToHost, pg->s));
    CK(cudaMemcpyAsync(&info_h, pg->info, sizeof(int), cudaMemcpyDeviceToHost, pg->s));
    CK(cudaStreamSynchronize(pg->s));
    CK(cudaGetLastError());
    reuse = true;
    const double mcc = pg->red_host[0], step_norm = std::sqrt(pg->red_host[1]), cand = pg->red_host[4];
    const bool valid = info_h == 0 && std::isfinite(mcc) && mcc > 0.0;
    if (!valid) {
      if (++n_invalid >= o.max_num_consecutive_invalid_steps) {
        term = STBA_FAILURE; msg = "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps"; break;
      }
      radius /= dec; dec *= 2.0; reuse = false;
      continue;
    }
    n_invalid = 0;
    it.step_is_valid = 1;
    it.step_norm = step_norm;
    const bool cand_ok = std::isfinite(cand);
    if (step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) { term = STBA_CONVERGENCE; msg = "Parameter tolerance reached."; break; }
    if (cand_ok) {
      it.cost_change = x_cost - cand;
      if (std::fabs(it.cost_change) <= o.function_tolerance * x_cost) { term = STBA_CONVERGENCE; msg = "Function tolerance reached."; break; }
    }
    const double rho = cand_ok ? (x_cost - cand) / mcc : -1.7976931348623157e308;
    it.relative_decrease = rho;
    if (rho > o.min_relative_decrease) {
      std::swap(pg->q, pg->q2); std::swap(pg->t, pg->t2);
      if ((r = pg->linearize(&x_cost)) != STBA_OK) return r;
      if ((r = pg->grad_norms(&g2, &gmax)) != STBA_OK) return r;
      // |x| of the new point: share[3c+2] of an update with a zero step
      k_pg_update<<<(n + 127) / 128, 128, 0, pg->s>>>(n, pg->q, pg->t, pg->ys, pg->scale, pg->gs, pg->diag, 0.0, pg->q2, pg->t2, pg->share);
      k_pg_sum<<<1, 256, 0, pg->s>>>(n, 3, 3, pg->share, pg->red);
      pg->launches += 2;
      CK(cudaMemcpyAsync(pg->red_host, pg->red, 3 * sizeof(double), cudaMemcpyDeviceToHost, pg->s));
      CK(cudaStreamSynchronize(pg->s));
      x_norm = std::sqrt(pg->red_host[2]);
      it.cost = x_cost; it.gradient_norm = std::sqrt(g2); it.gradient_max_norm = gmax; it.step_is_successful = 1;
      const double tt = 2.0 * rho - 1.0;
      radius = std::min(o.max_trust_region_radius, radius / std::max(1.0 / 3.0, 1.0 - tt * tt * tt));
      dec = 2.0; reuse = false;
    } else {
      it.cost = cand_ok ? cand : x_cost;
      radius /= dec; dec *= 2.0; reuse = true;
    }
  }
  if (sum) {
    sum->termination_type = term;
    sum->num_successful_steps = n_ok; sum->num_unsuccessful_steps = n_bad;
    sum->initial_cost = initial_cost; sum->final_cost = x_cost;
    sum->total_time_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
    sum->gpu_launches = pg->launches - launches0;
    strncpy(sum->message, msg, sizeof(sum->message) - 1);
    sum->reserved = sum->num_iterations;
    if (sum->iterations && sum->num_iterations > sum->iterations_capacity) sum->num_iterations = sum->iterations_capacity;
  }
  return STBA_OK;
}

}  // extern "C"
