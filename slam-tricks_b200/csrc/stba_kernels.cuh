// Kernels of the bundle-adjustment hot path (fp64, sm_100a).  See DESIGN.md §3 for the data
// layout and the roofline of each kernel.  Reference behaviour restated (never copied):
//   residual      st20-g2o/src/include/test_ceres.h:63-80
//   manifold      test_ceres.h:14-45
//   problem       test_ceres.h:98-152 (landmark-major residual order, constant end cameras)
//   J^T J pattern st20-g2o/src/include/sim_data.h:108-159
#pragma once
#include "stba_dev.cuh"

namespace stba {

constexpr int kBlock = 128;          // threads per CTA of the per-landmark / per-chunk kernels
constexpr int kCamAcc = 23;          // camera-frame accumulators of lin_cam (see below)
constexpr int kDiagAcc = 27;         // 21 (E E^T upper) + 6 (E h) of schur_diag
// E record per observation: 18 doubles padded to 20 (160 B = five 32-byte sectors).  The consumers gather whole records
// at random (k_schur_off: two per pair, one thread per block) and are bound by L1 wavefronts, one per lane and load
// instruction: five 256-bit loads (sm_100 LDG.E.256) per record instead of nine 128-bit ones.
constexpr int kEStride = 20;
__device__ __forceinline__ void ldg_e(const double* __restrict__ rec, double (&e)[20]) {
#pragma unroll
  for (int k = 0; k < 20; k += 4)
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(e[k]), "=d"(e[k + 1]), "=d"(e[k + 2]), "=d"(e[k + 3]) : "l"(rec + k));
}
__device__ __forceinline__ void stg4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------
// camera tiles: q,t -> [R row-major | t]
// ---------------------------------------------------------------------------------------------
__global__ void k_cam_prep(int n_cam, const double* __restrict__ q, const double* __restrict__ t,
                           double* __restrict__ Rt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cam) return;
  double R[9];
  quat_to_rot(q[4 * c], q[4 * c + 1], q[4 * c + 2], q[4 * c + 3], R);
#pragma unroll
  for (int k = 0; k < 9; ++k) Rt[kCamTile * c + k] = R[k];
  Rt[kCamTile * c + 9] = t[3 * c];
  Rt[kCamTile * c + 10] = t[3 * c + 1];
  Rt[kCamTile * c + 11] = t[3 * c + 2];
  Rt[kCamTile * c + 12] = 0.0;
  Rt[kCamTile * c + 13] = 0.0;
}

// The linearisation kernels (lin_lm2 / lin_cam2) live in stba_lin.cuh.  lin_cam2 accumulates in
// the CAMERA frame, where the per-observation blocks collapse to 23 sums
//   T  (6)  = sum J_th^T J_th                       (theta-theta block, already body-frame)
//   K  (7)  = sum J_th^T Pi'   : iz*a, iz*c, iz*b, iz*(a u + c v), iz*(b u + a v), iz*v, iz*u
//   Q  (4)  = sum Pi'^T Pi'    : iz^2, iz^2 u, iz^2 v, iz^2 (u^2+v^2)
//   gth(3)  = sum J_th^T r
//   m  (3)  = sum Pi'^T r      : iz r0, iz r1, -iz (u r0 + v r1)
// with a = uv, b = 1+u^2, c = 1+v^2.  The finish kernel below rotates them into the world-frame
// translation tangent once per camera:  H_tt = R Q R^T, H_th,t = -K R^T, g_t = -R m.

// ---------------------------------------------------------------------------------------------
// Jacobi column scaling, computed once at x0 (Ceres: 1 / (1 + sqrt(squared column norm)))
// ---------------------------------------------------------------------------------------------
__global__ void k_jacobi_scale(int n_cam, int n_lm, int enable, const double* __restrict__ Hcc,
                               const double* __restrict__ Hll, double* __restrict__ sc,
                               double* __restrict__ sl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_cam) {
#pragma unroll
    for (int k = 0; k < 6; ++k)
      sc[(size_t)i * 6 + k] = enable ? 1.0 / (1.0 + sqrt(Hcc[(size_t)i * 21 + tri6(k, k)])) : 1.0;
  }
  if (i < n_lm) {
    const double* H = Hll + 6 * (size_t)i;
    sl[(size_t)i * 3 + 0] = enable ? 1.0 / (1.0 + sqrt(H[0])) : 1.0;
    sl[(size_t)i * 3 + 1] = enable ? 1.0 / (1.0 + sqrt(H[3])) : 1.0;
    sl[(size_t)i * 3 + 2] = enable ? 1.0 / (1.0 + sqrt(H[5])) : 1.0;
  }
}

// LM diagonal in UNSCALED variables.  Ceres solves (Js^T Js + D^2) ys = Js^T r with
// Js = J diag(s), D^2 = clamp(diag(Js^T Js), lo, hi) / radius.  Substituting y = s * ys gives
// (J^T J + diag(D^2 / s^2)) y = J^T r, so the kernels never scale a Jacobian; only this
// diagonal carries s.
__device__ __forceinline__ double lm_diag(double h, double s, double lo, double hi, double inv_radius) {
  const double s2 = s * s;
  return fmin(fmax(s2 * h, lo), hi) * inv_radius / s2;
}

// ---------------------------------------------------------------------------------------------
// schur_lm — per landmark: A = H_ll + D_l^2 = L L^T, Linv = L^-1 (so (A)^-1 = Linv^T Linv),
// h = Linv g_l, and per observation E = W Linv^T with W = Jc^T Jl (6x3).  Then
//   S  = H_cc + D_c^2 - sum E E^T,   rhs = g_c - sum E h,   y_l = Linv^T (h - sum E^T y_c).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
k_schur_lm(int n_lm, const int* __restrict__ lm_ptr, const int* __restrict__ obs_cam,
           const double* __restrict__ obs_uv, const double* __restrict__ Rt,
           const double* __restrict__ lm4, const uint8_t* __restrict__ cam_const,
           const uint8_t* __restrict__ lm_const, const double* __restrict__ Hll,
           const double* __restrict__ gl, const double* __restrict__ sl, double lo, double hi,
           double inv_radius, double* __restrict__ Dl2, double* __restrict__ Linv,
           double* __restrict__ hl, double* __restrict__ E) {
  for (int l = blockIdx.x * kBlock + threadIdx.x; l < n_lm; l += gridDim.x * kBlock) {
    if (lm_const && lm_const[l]) {
#pragma unroll
      for (int k = 0; k < 6; ++k) Linv[6 * (size_t)l + k] = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) { hl[3 * (size_t)l + k] = 0.0; Dl2[3 * (size_t)l + k] = 0.0; }
      continue;
    }
    const double* H = Hll + 6 * (size_t)l;
    const double d0 = lm_diag(H[0], sl[3 * (size_t)l], lo, hi, inv_radius);
    const double d1 = lm_diag(H[3], sl[3 * (size_t)l + 1], lo, hi, inv_radius);
    const double d2 = lm_diag(H[5], sl[3 * (size_t)l + 2], lo, hi, inv_radius);
    Dl2[3 * (size_t)l] = d0; Dl2[3 * (size_t)l + 1] = d1; Dl2[3 * (size_t)l + 2] = d2;
    // Cholesky of the 3x3 A (lower): [l00; l10 l11; l20 l21 l22]
    const double a00 = H[0] + d0, a10 = H[1], a20 = H[2], a11 = H[3] + d1, a21 = H[4], a22 = H[5] + d2;
    const double l00 = sqrt(a00), i00 = 1.0 / l00;
    const double l10 = a10 * i00, l20 = a20 * i00;
    const double l11 = sqrt(a11 - l10 * l10), i11 = 1.0 / l11;
    const double l21 = (a21 - l20 * l10) * i11;
    const double l22 = sqrt(a22 - l20 * l20 - l21 * l21), i22 = 1.0 / l22;
    // Linv (lower): m00; m10 m11; m20 m21 m22
    const double m00 = i00, m11 = i11, m22 = i22;
    const double m10 = -l10 * m00 * i11;
    const double m21 = -l21 * m11 * i22;
    const double m20 = -(l20 * m00 + l21 * m10) * i22;
    double* Lo = Linv + 6 * (size_t)l;
    Lo[0] = m00; Lo[1] = m10; Lo[2] = m11; Lo[3] = m20; Lo[4] = m21; Lo[5] = m22;
    const double g0 = gl[3 * (size_t)l], g1 = gl[3 * (size_t)l + 1], g2 = gl[3 * (size_t)l + 2];
    hl[3 * (size_t)l] = m00 * g0;
    hl[3 * (size_t)l + 1] = m10 * g0 + m11 * g1;
    hl[3 * (size_t)l + 2] = m20 * g0 + m21 * g1 + m22 * g2;

  }
}

// E per observation, one thread per OBSERVATION (ten times the parallelism of walking a landmark's
// observations in its thread: the camera-tile gather is the latency, occupancy hides it):
//   E = W L^-T,  W = Jc^T Jl,  L L^T = H_ll + D_l^2  (Linv from k_schur_lm)
__global__ void __launch_bounds__(kBlock)
k_schur_E(int64_t n_obs, const int* __restrict__ obs_cam, const int* __restrict__ obs_lm, const double* __restrict__ obs_uv,
          const double* __restrict__ Rt, const double* __restrict__ lm4, const uint8_t* __restrict__ cam_const,
          const uint8_t* __restrict__ lm_const, const double* __restrict__ Linv, double* __restrict__ E) {
  for (int64_t o = blockIdx.x * (int64_t)kBlock + threadIdx.x; o < n_obs; o += (int64_t)gridDim.x * kBlock) {
    const int c = __ldg(obs_cam + o), l = __ldg(obs_lm + o);
    double* Eo = E + kEStride * (size_t)o;
    if (cam_const[c] || (lm_const && lm_const[l])) {
#pragma unroll
      for (int k = 0; k < kEStride; k += 4) stg4(Eo + k, 0.0, 0.0, 0.0, 0.0);
      continue;
    }
    const double2 uv = ldg2(obs_uv + 2 * (size_t)o);
    const double2 pxy = ldg2(lm4 + 4 * (size_t)l);
    const double pz = __ldg(lm4 + 4 * (size_t)l + 2);
    const double2 ma = ldg2(Linv + 6 * (size_t)l), mb = ldg2(Linv + 6 * (size_t)l + 2), mc = ldg2(Linv + 6 * (size_t)l + 4);
    const double m00 = ma.x, m10 = ma.y, m11 = mb.x, m20 = mb.y, m21 = mc.x, m22 = mc.y;
    double T[kCamVals];
#pragma unroll
    for (int k = 0; k < kCamVals; k += 2) {
      const double2 x = ldg2(Rt + (size_t)kCamTile * c + k);
      T[k] = x.x; T[k + 1] = x.y;
    }
    const Obs ob = project(T, pxy.x, pxy.y, pz, uv.x, uv.y);
    double J0[3], J1[3], T0[3], T1[3];
    landmark_jacobian(T, ob, J0, J1);
    rotation_jacobian(ob, T0, T1);
    // Z = Jl Linv^T  (2x3):  Z[r][k] = sum_{j<=k} Jl[r][j] Linv[k][j]
    const double z00 = J0[0] * m00, z01 = J0[0] * m10 + J0[1] * m11, z02 = J0[0] * m20 + J0[1] * m21 + J0[2] * m22;
    const double z10 = J1[0] * m00, z11 = J1[0] * m10 + J1[1] * m11, z12 = J1[0] * m20 + J1[1] * m21 + J1[2] * m22;
    double e[20];
    e[18] = 0.0; e[19] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {           // E_theta = J_th^T Z ; E_t = -Jl^T Z
      e[3 * i + 0] = T0[i] * z00 + T1[i] * z10;
      e[3 * i + 1] = T0[i] * z01 + T1[i] * z11;
      e[3 * i + 2] = T0[i] * z02 + T1[i] * z12;
      e[9 + 3 * i + 0] = -(J0[i] * z00 + J1[i] * z10);
      e[9 + 3 * i + 1] = -(J0[i] * z01 + J1[i] * z11);
      e[9 + 3 * i + 2] = -(J0[i] * z02 + J1[i] * z12);
    }
#pragma unroll
    for (int k = 0; k < kEStride; k += 4) stg4(Eo + k, e[k], e[k + 1], e[k + 2], e[k + 3]);
  }
}

// per camera LM diagonal D_c^2 (unscaled form) — tiny
__global__ void k_cam_diag(int n_cam, const double* __restrict__ Hcc, const double* __restrict__ sc,
                           double lo, double hi, double inv_radius, double* __restrict__ Dc2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cam * 6) return;
  const int c = i / 6, k = i % 6;
  Dc2[i] = lm_diag(Hcc[(size_t)c * 21 + tri6(k, k)], sc[i], lo, hi, inv_radius);
}

// schur_diag — camera-major: per chunk  sum_a E_a E_a^T (21 unique) and sum_a E_a h_l(a) (6).
// One warp per CTA (grid = n_chunk): provably convergent, so the final reduction is plain SHFL.BFLY.
__global__ void __launch_bounds__(32)
k_schur_diag(const int* __restrict__ chunk_beg, const int* __restrict__ chunk_end,
             const int* __restrict__ cam_perm, const int* __restrict__ cobs_lm,
             const double* __restrict__ E, const double* __restrict__ hl,
             double* __restrict__ chunk_acc) {
  const int lane = threadIdx.x, ch = blockIdx.x;
  double acc[kDiagAcc];
#pragma unroll
  for (int k = 0; k < kDiagAcc; ++k) acc[k] = 0.0;
  const int end = chunk_end[ch];
  for (int o = chunk_beg[ch] + lane; o < end; o += 32) {
    const int a = __ldg(cam_perm + o);
    const int l = __ldg(cobs_lm + o);
    double e[20];
    ldg_e(E + kEStride * (size_t)a, e);
    const double h0 = __ldg(hl + 3 * (size_t)l), h1 = __ldg(hl + 3 * (size_t)l + 1), h2 = __ldg(hl + 3 * (size_t)l + 2);
    int t = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = i; j < 6; ++j, ++t)
        acc[t] = fma(e[3 * i], e[3 * j], fma(e[3 * i + 1], e[3 * j + 1], fma(e[3 * i + 2], e[3 * j + 2], acc[t])));
#pragma unroll
    for (int i = 0; i < 6; ++i)
      acc[21 + i] = fma(e[3 * i], h0, fma(e[3 * i + 1], h1, fma(e[3 * i + 2], h2, acc[21 + i])));
  }
#pragma unroll
  for (int k = 0; k < kDiagAcc; ++k) {
    const double r = warp_sum(acc[k]);
    if (lane == 0) chunk_acc[(size_t)ch * kDiagAcc + k] = r;
  }
}

// per free camera: S_ii = H_cc + D_c^2 - sum E E^T  (full 6x6 written, both triangles of the
// diagonal block), rhs_i = g_c - sum E h
__global__ void k_schur_diag_finish(int n_cam, const int* __restrict__ cam_chunk_ptr,
                                    const int* __restrict__ free_of, const double* __restrict__ chunk_acc,
                                    const double* __restrict__ Hcc, const double* __restrict__ gc,
                                    const double* __restrict__ Dc2, double* __restrict__ S, int n,
                                    double* __restrict__ rhs, int include_cam, int packed) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cam) return;
  const int f = free_of[c];
  if (f < 0) return;
  double a[kDiagAcc];
#pragma unroll
  for (int k = 0; k < kDiagAcc; ++k) a[k] = 0.0;
  for (int ch = cam_chunk_ptr[c]; ch < cam_chunk_ptr[c + 1]; ++ch)
#pragma unroll
    for (int k = 0; k < kDiagAcc; ++k) a[k] += chunk_acc[(size_t)ch * kDiagAcc + k];
  int t = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 6; ++j, ++t) {
      double v = -a[t];
      if (include_cam) {
        v += Hcc[(size_t)c * 21 + t];
        if (i == j) v += Dc2[(size_t)c * 6 + i];
      }
      if (packed) {          // multi-GPU send buffer: row-major packed lower triangle, (r, c <= r) at r (r + 1) / 2 + c
        const size_t r = 6 * f + j, cc = 6 * f + i;
        S[r * (r + 1) / 2 + cc] = v;
      } else {
        S[(size_t)(6 * f + i) + (size_t)(6 * f + j) * n] = v;
        S[(size_t)(6 * f + j) + (size_t)(6 * f + i) * n] = v;
      }
    }
#pragma unroll
  for (int i = 0; i < 6; ++i) rhs[6 * f + i] = (include_cam ? gc[(size_t)c * 6 + i] : 0.0) - a[21 + i];
}

// multi-GPU: after the all-reduce of the landmark-sharded partial sums, add the (already
// reduced, replicated) camera blocks once:  S_ii += H_cc + D_c^2,  rhs_i += g_c
__global__ void k_add_cam_blocks(int n_cam, const int* __restrict__ free_of, const double* __restrict__ Hcc,
                                 const double* __restrict__ gc, const double* __restrict__ Dc2,
                                 double* __restrict__ S, int n, double* __restrict__ rhs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cam) return;
  const int f = free_of[c];
  if (f < 0) return;
  int t = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 6; ++j, ++t) {
      double v = Hcc[(size_t)c * 21 + t];
      if (i == j) v += Dc2[(size_t)c * 6 + i];
      S[(size_t)(6 * f + i) + (size_t)(6 * f + j) * n] += v;
      if (i != j) S[(size_t)(6 * f + j) + (size_t)(6 * f + i) * n] += v;
    }
#pragma unroll
  for (int i = 0; i < 6; ++i) rhs[6 * f + i] += gc[(size_t)c * 6 + i];
}

// schur_off — one thread per strictly-lower 6x6 block (i > j, free-camera indices):
// S_ij = - sum over the block's incidence list of E_a E_b^T.  The list (pairs of observation
// indices of one landmark seen by both cameras) is static and sorted, so the sum order is fixed:
// no atomics, no zero-fill, every block written exactly once.
__global__ void __launch_bounds__(kBlock)
k_schur_off(int64_t b_begin, int64_t n_blk, const int64_t* __restrict__ blk_ptr, const uint64_t* __restrict__ inc,
            const double* __restrict__ E, double* __restrict__ S, int n, int packed) {
  // blocks [b_begin, n_blk): the multi-GPU path produces the packed buffer block-row range by block-row range so that
  // the all-reduce of a finished range overlaps the computation of the next one
  for (int64_t b = b_begin + (int64_t)blockIdx.x * kBlock + threadIdx.x; b < n_blk; b += (int64_t)gridDim.x * kBlock) {
    // b = i (i-1) / 2 + j,  i > j >= 0
    int64_t i = (int64_t)((1.0 + sqrt(1.0 + 8.0 * (double)b)) * 0.5);
    while (i * (i - 1) / 2 > b) --i;
    while ((i + 1) * i / 2 <= b) ++i;
    const int64_t j = b - i * (i - 1) / 2;
    double acc[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) acc[k] = 0.0;
    for (int64_t p = blk_ptr[b]; p < blk_ptr[b + 1]; ++p) {
      const uint64_t ab = inc[p];
      double ea[20], eb[20];
      ldg_e(E + kEStride * (size_t)(ab >> 32), ea);
      ldg_e(E + kEStride * (size_t)(ab & 0xffffffffu), eb);
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c)
          acc[6 * c + r] = fma(ea[3 * r], eb[3 * c], fma(ea[3 * r + 1], eb[3 * c + 1], fma(ea[3 * r + 2], eb[3 * c + 2], acc[6 * c + r])));
    }
    if (packed) {            // row-major packed lower triangle (multi-GPU send buffer)
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        const size_t row = 6 * i + r;
        double* dst = S + row * (row + 1) / 2 + 6 * j;
#pragma unroll
        for (int c = 0; c < 6; ++c) dst[c] = -acc[6 * c + r];
      }
      continue;
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double2* col = reinterpret_cast<double2*>(S + (size_t)(6 * i) + (size_t)(6 * j + c) * n);
      col[0] = make_double2(-acc[6 * c], -acc[6 * c + 1]);
      col[1] = make_double2(-acc[6 * c + 2], -acc[6 * c + 3]);
      col[2] = make_double2(-acc[6 * c + 4], -acc[6 * c + 5]);
    }
  }
}

// multi-GPU: the reduced packed lower triangle (row-major: (r, c <= r) at r (r + 1) / 2 + c) back into the
// column-major S the factorisation works on, through a 32 x 32 shared-memory transpose (both sides coalesced).
// grid: (tiles, tiles), tile (x = column tile, y = row tile), only y >= x does work.
// schur_off for FEW, LONG incidence lists (config B: 1 128 blocks with ~200 pairs each — one thread per block left the
// build at 0.55 ms, as long as at config C with twenty times the work): one WARP per block, lanes stride over the list,
// the 36 per-lane sums are added in lane order through shared memory (fixed order: deterministic), lanes 0..35 write.
constexpr int kOffWarps = 4;
__global__ void __launch_bounds__(32 * kOffWarps)
k_schur_off_warp(int64_t b_begin, int64_t n_blk, const int64_t* __restrict__ blk_ptr, const uint64_t* __restrict__ inc,
                 const double* __restrict__ E, double* __restrict__ S, int n, int packed) {
  __shared__ double s_red[kOffWarps][36][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t b = b_begin + (int64_t)blockIdx.x * kOffWarps + warp; b < n_blk; b += (int64_t)gridDim.x * kOffWarps) {
    int64_t i = (int64_t)((1.0 + sqrt(1.0 + 8.0 * (double)b)) * 0.5);
    while (i * (i - 1) / 2 > b) --i;
    while ((i + 1) * i / 2 <= b) ++i;
    const int64_t j = b - i * (i - 1) / 2;
    double acc[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) acc[k] = 0.0;
    const int64_t end = blk_ptr[b + 1];
    for (int64_t p = blk_ptr[b] + lane; p < end; p += 32) {
      const uint64_t ab = inc[p];
      double ea[20], eb[20];
      ldg_e(E + kEStride * (size_t)(ab >> 32), ea);
      ldg_e(E + kEStride * (size_t)(ab & 0xffffffffu), eb);
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c)
          acc[6 * c + r] = fma(ea[3 * r], eb[3 * c], fma(ea[3 * r + 1], eb[3 * c + 1], fma(ea[3 * r + 2], eb[3 * c + 2], acc[6 * c + r])));
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 36; ++k) s_red[warp][k][lane] = acc[k];
    __syncwarp();
    for (int k = lane; k < 36; k += 32) {            // k = 6 c + r
      double t = 0.0;
#pragma unroll 8
      for (int l = 0; l < 32; ++l) t += s_red[warp][k][l];
      const int c = k / 6, r = k % 6;
      if (packed) {
        const size_t row = 6 * i + r;
        S[row * (row + 1) / 2 + 6 * j + c] = -t;
      } else {
        S[(size_t)(6 * i + r) + (size_t)(6 * j + c) * n] = -t;
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) k_unpack_lower(const double* __restrict__ Sp, int n, double* __restrict__ S, int ld) {
  __shared__ double tile[32][33];
  const int tc = blockIdx.x, tr = blockIdx.y;
  if (tr < tc) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = tr * 32 + ty + 8 * k, c = tc * 32 + tx;
    tile[ty + 8 * k][tx] = (r < n && c <= r) ? Sp[(size_t)r * (r + 1) / 2 + c] : 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = tc * 32 + ty + 8 * k, r = tr * 32 + tx;
    if (r < n && c <= r) S[(size_t)c * ld + r] = tile[tx][ty + 8 * k];
  }
}

// multi-GPU: scalars of the one collective.  tail = [cost, |x_l|^2, |g_l|^2, max|g_l| per rank (8 slots)] after the
// sum-reduce; cam = camera-side [sum of squares, max] of the gradient (identical on every rank).
__global__ void k_combine_reduced(const double* __restrict__ tail, int nranks, int with_xnorm, double* __restrict__ scal, int sc_cost,
                                  int sc_g2, int sc_gmax, int sc_xnorm2) {
  if (threadIdx.x || blockIdx.x) return;
  scal[sc_cost] = tail[0];
  if (with_xnorm) scal[sc_xnorm2] += tail[1];
  scal[sc_g2] += tail[2];
  double m = scal[sc_gmax];
  for (int r = 0; r < nranks && r < 8; ++r) m = fmax(m, tail[3 + r]);
  scal[sc_gmax] = m;
}
__global__ void k_fill_tail(double* __restrict__ tail, int rank, const double* __restrict__ scal, int sc_cost, int sc_g2, int sc_gmax,
                            int sc_xnorm2) {
  if (threadIdx.x || blockIdx.x) return;
  tail[0] = scal[sc_cost];
  tail[1] = scal[sc_xnorm2];
  tail[2] = scal[sc_g2];
  for (int r = 0; r < 8; ++r) tail[3 + r] = (r == rank) ? scal[sc_gmax] : 0.0;
}

// scatter the reduced solution into per-camera rows (zero for constant cameras)
__global__ void k_scatter_yc(int n_cam, const int* __restrict__ free_of, const double* __restrict__ y,
                             double* __restrict__ yc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cam * 6) return;
  const int f = free_of[i / 6];
  yc[i] = f < 0 ? 0.0 : y[6 * f + i % 6];
}

// ---------------------------------------------------------------------------------------------
// back-substitution + candidate landmarks:  y_l = Linv^T (h - sum_a E_a^T y_c[cam_a]),
// P+ = P - y_l.  Scalars: [0] sum y_l.(g_l + D_l^2 y_l), [1] |dP|^2, [2] |P+|^2
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
k_backsub(int n_lm, const int* __restrict__ lm_ptr, const int* __restrict__ obs_cam,
          const double* __restrict__ E, const double* __restrict__ yc, const double* __restrict__ Linv,
          const double* __restrict__ hl, const double* __restrict__ gl, const double* __restrict__ Dl2,
          const double* __restrict__ lm4, double* __restrict__ yl, double* __restrict__ lm4_new,
          double* partial, unsigned int* counter, double* out) {
  double sc[3] = {0.0, 0.0, 0.0};
  for (int l = blockIdx.x * kBlock + threadIdx.x; l < n_lm; l += gridDim.x * kBlock) {
    double w0 = hl[3 * (size_t)l], w1 = hl[3 * (size_t)l + 1], w2 = hl[3 * (size_t)l + 2];
    const int end = lm_ptr[l + 1];
    for (int o = lm_ptr[l]; o < end; ++o) {
      const int c = __ldg(obs_cam + o);
      double y[6];
#pragma unroll
      for (int k = 0; k < 6; k += 2) {
        const double2 x = ldg2(yc + 6 * (size_t)c + k);
        y[k] = x.x; y[k + 1] = x.y;
      }
      double e[20];
      ldg_e(E + kEStride * (size_t)o, e);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        w0 = fma(-e[3 * i], y[i], w0);
        w1 = fma(-e[3 * i + 1], y[i], w1);
        w2 = fma(-e[3 * i + 2], y[i], w2);
      }
    }
    const double* m = Linv + 6 * (size_t)l;   // m00; m10 m11; m20 m21 m22
    const double y0 = m[0] * w0 + m[1] * w1 + m[3] * w2;
    const double y1 = m[2] * w1 + m[4] * w2;
    const double y2 = m[5] * w2;
    yl[3 * (size_t)l] = y0; yl[3 * (size_t)l + 1] = y1; yl[3 * (size_t)l + 2] = y2;
    const double px = lm4[4 * (size_t)l], py = lm4[4 * (size_t)l + 1], pz = lm4[4 * (size_t)l + 2];
    const double nx = px - y0, ny = py - y1, nz = pz - y2;
    lm4_new[4 * (size_t)l] = nx; lm4_new[4 * (size_t)l + 1] = ny; lm4_new[4 * (size_t)l + 2] = nz;
    lm4_new[4 * (size_t)l + 3] = 0.0;
    sc[0] += y0 * (gl[3 * (size_t)l] + Dl2[3 * (size_t)l] * y0) + y1 * (gl[3 * (size_t)l + 1] + Dl2[3 * (size_t)l + 1] * y1) +
             y2 * (gl[3 * (size_t)l + 2] + Dl2[3 * (size_t)l + 2] * y2);
    // ambient differences are taken exactly as (x+ - x), the way Ceres forms x_plus_delta - x
    const double ex = nx - px, ey = ny - py, ez = nz - pz;
    sc[1] += ex * ex + ey * ey + ez * ez;
    sc[2] += nx * nx + ny * ny + nz * nz;
  }
  grid_reduce<3, kBlock>(sc, 3, partial, counter, out);
}

// candidate cameras: q+ = q * exp(-y_theta), t+ = t - y_t.
// Scalars: [0] sum y_c.(g_c + D_c^2 y_c), [1] |dx|^2 ambient, [2] |x+|^2 ambient (free cameras)
__global__ void __launch_bounds__(kBlock)
k_cam_update(int n_cam, const uint8_t* __restrict__ cam_const, const double* __restrict__ q,
             const double* __restrict__ t, const double* __restrict__ yc, const double* __restrict__ gc,
             const double* __restrict__ Dc2, double* __restrict__ q_new, double* __restrict__ t_new,
             double* partial, unsigned int* counter, double* out) {
  double sc[3] = {0.0, 0.0, 0.0};
  for (int c = blockIdx.x * kBlock + threadIdx.x; c < n_cam; c += gridDim.x * kBlock) {
    double qo[4] = {q[4 * c], q[4 * c + 1], q[4 * c + 2], q[4 * c + 3]};
    double to[3] = {t[3 * c], t[3 * c + 1], t[3 * c + 2]};
    double qn[4] = {qo[0], qo[1], qo[2], qo[3]};
    double tn[3] = {to[0], to[1], to[2]};
    if (!cam_const[c]) {
      const double* y = yc + 6 * (size_t)c;
      double ex[4];
      so3_exp_quat(-y[0], -y[1], -y[2], ex);
      quat_mul_normalized(qo, ex, qn);
#pragma unroll
      for (int k = 0; k < 3; ++k) tn[k] = to[k] - y[3 + k];
#pragma unroll
      for (int k = 0; k < 6; ++k) sc[0] += y[k] * (gc[6 * (size_t)c + k] + Dc2[6 * (size_t)c + k] * y[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const double d = qn[k] - qo[k]; sc[1] += d * d; sc[2] += qn[k] * qn[k]; }
#pragma unroll
      for (int k = 0; k < 3; ++k) { const double d = tn[k] - to[k]; sc[1] += d * d; sc[2] += tn[k] * tn[k]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) q_new[4 * c + k] = qn[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) t_new[3 * c + k] = tn[k];
  }
  grid_reduce<3, kBlock>(sc, 3, partial, counter, out);
}

// gradient norms the way Ceres forms them: | x - Plus(x, -g) | in ambient coordinates.
// Scalars: [0] sum of squares, [1] max abs; both over cameras (free) and landmarks (free).
__global__ void __launch_bounds__(kBlock)
k_grad_norm(int n_cam, int n_lm, const uint8_t* __restrict__ cam_const, const uint8_t* __restrict__ lm_const,
            const double* __restrict__ q, const double* __restrict__ gc, const double* __restrict__ gl,
            double* partial, unsigned int* counter, double* out) {
  double sc[2] = {0.0, 0.0};
  const int total = n_cam + n_lm;
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < total; i += gridDim.x * kBlock) {
    if (i < n_cam) {
      if (cam_const[i]) continue;
      const double* g = gc + 6 * (size_t)i;
      double qo[4] = {q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]}, ex[4], qn[4];
      so3_exp_quat(-g[0], -g[1], -g[2], ex);
      quat_mul_normalized(qo, ex, qn);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const double d = qo[k] - qn[k]; sc[0] += d * d; sc[1] = fmax(sc[1], fabs(d)); }
#pragma unroll
      for (int k = 0; k < 3; ++k) { const double d = g[3 + k]; sc[0] += d * d; sc[1] = fmax(sc[1], fabs(d)); }
    } else {
      const int l = i - n_cam;
      if (lm_const && lm_const[l]) continue;
#pragma unroll
      for (int k = 0; k < 3; ++k) { const double d = gl[3 * (size_t)l + k]; sc[0] += d * d; sc[1] = fmax(sc[1], fabs(d)); }
    }
  }
  grid_reduce<2, kBlock>(sc, 1, partial, counter, out);
}

// |x|^2 over the free blocks, ambient.  Scalars: [0]
__global__ void __launch_bounds__(kBlock)
k_x_norm(int n_cam, int n_lm, const uint8_t* __restrict__ cam_const, const uint8_t* __restrict__ lm_const,
         const double* __restrict__ q, const double* __restrict__ t, const double* __restrict__ lm4,
         double* partial, unsigned int* counter, double* out) {
  double sc[1] = {0.0};
  const int total = n_cam + n_lm;
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < total; i += gridDim.x * kBlock) {
    if (i < n_cam) {
      if (cam_const[i]) continue;
#pragma unroll
      for (int k = 0; k < 4; ++k) sc[0] += q[4 * i + k] * q[4 * i + k];
#pragma unroll
      for (int k = 0; k < 3; ++k) sc[0] += t[3 * i + k] * t[3 * i + k];
    } else {
      const int l = i - n_cam;
      if (lm_const && lm_const[l]) continue;
#pragma unroll
      for (int k = 0; k < 3; ++k) sc[0] += lm4[4 * (size_t)l + k] * lm4[4 * (size_t)l + k];
    }
  }
  grid_reduce<1, kBlock>(sc, 1, partial, counter, out);
}

// ---------------------------------------------------------------------------------------------
// integer preprocessing (bit-exact contract)
// ---------------------------------------------------------------------------------------------
// index ranges + landmark-major order of the observation list (the contract of stba_ba_create)
__global__ void k_validate_obs(int64_t n, const int* __restrict__ oc, const int* __restrict__ ol, int n_cam, int n_lm,
                               int* __restrict__ bad) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = i; k < n; k += stride) {
    const int c = oc[k], l = ol[k];
    if (c < 0 || c >= n_cam || l < 0 || l >= n_lm || (k && ol[k - 1] > l)) *bad = 1;
  }
}

__global__ void k_histogram(int64_t n, const int* __restrict__ key, int* __restrict__ hist) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(hist + key[i], 1);
}

// exclusive scan by ONE block of 1024 threads, chunked; T_in counts -> T_out offsets (n+1 entries)
template <typename TI, typename TO>
__global__ void __launch_bounds__(1024) k_exclusive_scan(int64_t n, const TI* __restrict__ in, TO* __restrict__ out) {
  __shared__ TO warp_tot[32];
  __shared__ TO carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const TO x = i < n ? (TO)in[i] : (TO)0;
    TO s = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const TO y = __shfl_up_sync(0xffffffffu, s, d);
      if (lane >= d) s += y;
    }
    if (lane == 31) warp_tot[warp] = s;
    __syncthreads();
    if (warp == 0) {
      TO w = warp_tot[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const TO y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const TO before = carry + (warp ? warp_tot[warp - 1] : (TO)0) + s - x;
    if (i < n) out[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
}

// Multi-block exclusive scan for long inputs (the pair-block counts: ~500 k entries at C took 0.49 ms in the one-block
// kernel above, 2 % of an end-to-end solve).  Three launches: per-block totals (kScanTile items per block), the
// one-block scan over the totals, then every block scans its tile from its offset.  Integer: order-independent.
constexpr int kScanItems = 8, kScanTile = 1024 * kScanItems;
template <typename TI, typename TO>
__global__ void __launch_bounds__(1024) k_scan_tile_sums(int64_t n, const TI* __restrict__ in, TO* __restrict__ tile_sum) {
  __shared__ TO warp_tot[32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  TO s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) s += base + k < n ? (TO)in[base + k] : (TO)0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    TO w = warp_tot[threadIdx.x];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(0xffffffffu, w, d);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = w;
  }
}
template <typename TI, typename TO>
__global__ void __launch_bounds__(1024) k_scan_tiles(int64_t n, const TI* __restrict__ in, const TO* __restrict__ tile_off, TO* __restrict__ out) {
  __shared__ TO warp_tot[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  TO x[kScanItems], s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { x[k] = base + k < n ? (TO)in[base + k] : (TO)0; s += x[k]; }
  TO incl = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const TO y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    TO w = warp_tot[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const TO y = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += y;
    }
    warp_tot[lane] = w;
  }
  __syncthreads();
  TO run = tile_off[blockIdx.x] + (warp ? warp_tot[warp - 1] : (TO)0) + incl - s;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = run;
    run += x[k];
  }
  if (base <= n - 1 && n - 1 < base + kScanItems) out[n] = run;      // the thread that owns the last item writes the total
}

// unstable bucket scatter; the per-bucket sort below makes the result the STABLE order
__global__ void k_bucket_scatter(int64_t n, const int* __restrict__ key, const int* __restrict__ ptr,
                                 int* __restrict__ cursor, int* __restrict__ perm) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = key[i];
    perm[ptr[k] + atomicAdd(cursor + k, 1)] = (int)i;
  }
}

// in-place ascending sort of `n` keys by a cooperating group of `nthr` threads (thread `tid`),
// "flip" bitonic network — every compare is ascending, so out-of-range partners act as +inf
// and arbitrary n is handled without padding.  `sync` must synchronise the group.
template <typename T, typename Sync>
__device__ __forceinline__ void group_sort(T* a, int64_t n, int tid, int nthr, Sync sync) {
  for (int64_t k = 2; (k >> 1) < n; k <<= 1) {
    for (int64_t i = tid; i < n; i += nthr) {
      const int64_t l = i ^ (k - 1);
      if (l > i && l < n) { const T x = a[i], y = a[l]; if (y < x) { a[i] = y; a[l] = x; } }
    }
    sync();
    for (int64_t j = k >> 2; j > 0; j >>= 1) {
      for (int64_t i = tid; i < n; i += nthr) {
        const int64_t l = i ^ j;
        if (l > i && l < n) { const T x = a[i], y = a[l]; if (y < x) { a[i] = y; a[l] = x; } }
      }
      sync();
    }
  }
}

// one CTA per bucket (camera): sort its slice of cam_perm ascending
__global__ void __launch_bounds__(256) k_sort_buckets_i32(int n_bucket, const int* __restrict__ ptr, int* perm) {
  for (int b = blockIdx.x; b < n_bucket; b += gridDim.x) {
    group_sort<int>(perm + ptr[b], ptr[b + 1] - ptr[b], threadIdx.x, 256, [] { __syncthreads(); });
    __syncthreads();
  }
}

// one warp per S block: sort its incidence slice (slices of at most 16 entries belong to the kernel below when skip_small)
__global__ void __launch_bounds__(kBlock) k_sort_segments_u64(int64_t n_seg, const int64_t* __restrict__ ptr, uint64_t* inc, int skip_small) {
  const int lane = threadIdx.x & 31;
  const int64_t wpg = (int64_t)gridDim.x * (kBlock / 32);
  for (int64_t s = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5); s < n_seg; s += wpg) {
    const int64_t p0 = ptr[s], len = ptr[s + 1] - p0;
    if (len < 2 || (skip_small && len <= 16)) continue;
    group_sort<uint64_t>(inc + p0, len, lane, 32, [] { __syncwarp(); });
  }
}
// Short slices (config C: ~9 pairs per block, 499 500 blocks): one HALF-WARP per slice, the keys in registers, a 16-lane
// bitonic network of shuffles (10 compare-exchange steps) — one streaming pass instead of a 32-lane shared-memory sort per
// slice (0.26 ms of the create path at C).  Unique keys: the sorted list is the same.
__global__ void __launch_bounds__(kBlock) k_sort_segments_16(int64_t n_seg, const int64_t* __restrict__ ptr, uint64_t* inc) {
  const int lane = threadIdx.x & 31, l16 = lane & 15;
  const unsigned hmask = 0xffffu << (lane & 16);
  const int64_t hpg = (int64_t)gridDim.x * (kBlock / 16);
  for (int64_t s = (int64_t)blockIdx.x * (kBlock / 16) + (threadIdx.x >> 4); s < n_seg; s += hpg) {
    const int64_t p0 = ptr[s];
    const int len = (int)(ptr[s + 1] - p0);
    if (len < 2 || len > 16) continue;                       // (uniform over the half-warp)
    unsigned long long key = l16 < len ? inc[p0 + l16] : ~0ull;
#pragma unroll
    for (int k = 2; k <= 16; k <<= 1)
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        const unsigned long long other = __shfl_xor_sync(hmask, key, j);
        const bool up = (l16 & k) == 0, low = (l16 & j) == 0;
        const unsigned long long mn = key < other ? key : other, mx = key < other ? other : key;
        key = (up == low) ? mn : mx;
      }
    if (l16 < len) inc[p0 + l16] = key;
  }
}

// gather the camera-major copies of the observation stream
__global__ void k_gather_cam_major(int64_t n, const int* __restrict__ perm, const int* __restrict__ obs_lm,
                                   const double* __restrict__ obs_uv, int* __restrict__ cobs_lm,
                                   double* __restrict__ cobs_uv) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = perm[i];
    cobs_lm[i] = obs_lm[a];
    cobs_uv[2 * i] = obs_uv[2 * (size_t)a];
    cobs_uv[2 * i + 1] = obs_uv[2 * (size_t)a + 1];
  }
}

// pair structure of the Schur complement: for every landmark, every pair of its observations
// whose cameras are both free: block (i > j) = free indices.  PASS 0 counts, PASS 1 fills.
template <int PASS>
__global__ void k_pair_pass(int n_lm, const int* __restrict__ lm_ptr, const int* __restrict__ obs_cam,
                            const int* __restrict__ free_of, const uint8_t* __restrict__ lm_const,
                            int* __restrict__ cnt, const int64_t* __restrict__ blk_ptr, uint64_t* __restrict__ inc,
                            int* __restrict__ dup_flag) {
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < n_lm; l += gridDim.x * blockDim.x) {
    if (lm_const && lm_const[l]) continue;
    const int beg = lm_ptr[l], end = lm_ptr[l + 1];
    for (int a = beg + 1; a < end; ++a) {
      const int fa = free_of[obs_cam[a]];
      if (fa < 0) continue;
      for (int b = beg; b < a; ++b) {
        const int fb = free_of[obs_cam[b]];
        if (fb == fa) { *dup_flag = 1; continue; }   // same (camera, landmark) observed twice: unsupported
        if (fb < 0) continue;
        const int64_t i = fa > fb ? fa : fb, j = fa > fb ? fb : fa;
        const int64_t blk = i * (i - 1) / 2 + j;
        const uint64_t hi = fa > fb ? (uint64_t)a : (uint64_t)b, lo = fa > fb ? (uint64_t)b : (uint64_t)a;
        if (PASS == 0) atomicAdd(cnt + blk, 1);
        else inc[blk_ptr[blk] + atomicAdd(cnt + blk, 1)] = (hi << 32) | lo;
      }
    }
  }
}

// co-visibility byte map over ALL camera pairs (i > j): map[i (i-1)/2 + j] = 1
__global__ void k_covis_mark(int n_lm, const int* __restrict__ lm_ptr, const int* __restrict__ obs_cam,
                             uint8_t* __restrict__ map) {
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < n_lm; l += gridDim.x * blockDim.x) {
    const int beg = lm_ptr[l], end = lm_ptr[l + 1];
    for (int a = beg + 1; a < end; ++a)
      for (int b = beg; b < a; ++b) {
        const int64_t ca = obs_cam[a], cb = obs_cam[b];
        if (ca == cb) continue;
        const int64_t i = ca > cb ? ca : cb, j = ca > cb ? cb : ca;
        map[i * (i - 1) / 2 + j] = 1;
      }
  }
}

// landmarks f64[n,3] <-> padded f64[n,4] (one 32-byte sector per point)
__global__ void k_pad_lm(int n_lm, const double* __restrict__ lm3, double* __restrict__ lm4) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lm) return;
  lm4[4 * (size_t)l] = lm3[3 * (size_t)l];
  lm4[4 * (size_t)l + 1] = lm3[3 * (size_t)l + 1];
  lm4[4 * (size_t)l + 2] = lm3[3 * (size_t)l + 2];
  lm4[4 * (size_t)l + 3] = 0.0;
}
__global__ void k_unpad_lm(int n_lm, const double* __restrict__ lm4, double* __restrict__ lm3) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lm) return;
  lm3[3 * (size_t)l] = lm4[4 * (size_t)l];
  lm3[3 * (size_t)l + 1] = lm4[4 * (size_t)l + 1];
  lm3[3 * (size_t)l + 2] = lm4[4 * (size_t)l + 2];
}

// write-only sweep of a buffer larger than L2 (used between timed repetitions)
__global__ void k_flush(size_t n, double* __restrict__ buf, double v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = v;
}

}  // namespace stba
