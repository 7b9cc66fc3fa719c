// The two stages in FRONT of the bundle-adjustment path, B200 only:
//
//   stba_visibility   ProblemScene::CreateMeasurements, st20-g2o/src/src/sim_data.cpp:119-142 —
//                     the O(N_c x N_l) visibility predicate with ordered compaction into the
//                     landmark -> [(camera, uv)] and camera -> [landmark] lists.  Index output is a
//                     bit-exact contract: the predicate is evaluated with __dmul_rn/__dadd_rn in
//                     one fixed order (no FMA contraction), the compaction is ballot + popcount
//                     in camera (resp. landmark) order — nothing depends on thread scheduling.
//   stba_triangulate  the per-landmark Ceres solve of ProblemScene::Simulation,
//                     sim_data.cpp:298-311 (functor sim_data.h:165-194): 100k independent
//                     3-parameter trust-region LM problems = ONE kernel, one thread per landmark
//                     running Ceres' control flow in registers.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>

#include "../../include/stba.h"

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

struct DevBuf {          // RAII list of stream-ordered allocations
  cudaStream_t s = nullptr;
  void* p[32];
  int n = 0;
  ~DevBuf() {
    for (int i = 0; i < n; ++i) cudaFreeAsync(p[i], s);
    if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
  }
  template <typename T>
  cudaError_t get(T** out, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMallocAsync(&q, std::max<size_t>(count, 1) * sizeof(T), s);
    if (e == cudaSuccess) p[n++] = q;
    *out = static_cast<T*>(q);
    return e;
  }
};

int open_device(int device, DevBuf& b) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return STBA_ERR_NO_DEVICE; }
  if (device < 0 || device >= ndev) return STBA_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(device));
  CK(cudaStreamCreateWithFlags(&b.s, cudaStreamNonBlocking));
  return STBA_OK;
}

constexpr int kW = 12;             // world->camera tile: rows of R (camera->world) [9] + t_cw [3]
constexpr int kVisCamChunk = 1024; // cameras staged in shared memory at a time (96 KB)
constexpr int kVisThreads = 256;

// OptPose::inverse() (sim_data.h:30-32) without fused multiply-adds — see oracle/front_oracle.py
__global__ void k_world_to_camera(int n_cam, const double* __restrict__ q, const double* __restrict__ t, double* __restrict__ W,
                                  int exact) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cam) return;
  const double x = q[4 * c], y = q[4 * c + 1], z = q[4 * c + 2], w = q[4 * c + 3];
  double R[9];
  const double xx = __dmul_rn(x, x), yy = __dmul_rn(y, y), zz = __dmul_rn(z, z);
  const double xy = __dmul_rn(x, y), xz = __dmul_rn(x, z), yz = __dmul_rn(y, z);
  const double xw = __dmul_rn(x, w), yw = __dmul_rn(y, w), zw = __dmul_rn(z, w);
  R[0] = __dsub_rn(1.0, __dmul_rn(2.0, __dadd_rn(yy, zz)));
  R[1] = __dmul_rn(2.0, __dsub_rn(xy, zw));
  R[2] = __dmul_rn(2.0, __dadd_rn(xz, yw));
  R[3] = __dmul_rn(2.0, __dadd_rn(xy, zw));
  R[4] = __dsub_rn(1.0, __dmul_rn(2.0, __dadd_rn(xx, zz)));
  R[5] = __dmul_rn(2.0, __dsub_rn(yz, xw));
  R[6] = __dmul_rn(2.0, __dsub_rn(xz, yw));
  R[7] = __dmul_rn(2.0, __dadd_rn(yz, xw));
  R[8] = __dsub_rn(1.0, __dmul_rn(2.0, __dadd_rn(xx, yy)));
  const double t0 = t[3 * c], t1 = t[3 * c + 1], t2 = t[3 * c + 2];
  double* o = W + (size_t)kW * c;
#pragma unroll
  for (int k = 0; k < 9; ++k) o[k] = R[k];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    o[9 + i] = -__dadd_rn(__dadd_rn(__dmul_rn(R[i], t0), __dmul_rn(R[3 + i], t1)), __dmul_rn(R[6 + i], t2));
  (void)exact;
}

// p_c = ((row0 P0 + row1 P1) + row2 P2) + t_cw, then sim_data.cpp:129-134
// (W is read with element stride S: 1 for a register tile, kVisCamChunk for the plane-major shared copy)
template <int S>
__device__ __forceinline__ bool visible(const double* __restrict__ W, double p0, double p1, double p2, double hw, double hh,
                                        double& u, double& v) {
  const double z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(W[2 * S], p0), __dmul_rn(W[5 * S], p1)), __dmul_rn(W[8 * S], p2)), W[11 * S]);
  if (z < 0.0) return false;
  const double x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(W[0], p0), __dmul_rn(W[3 * S], p1)), __dmul_rn(W[6 * S], p2)), W[9 * S]);
  const double y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(W[1 * S], p0), __dmul_rn(W[4 * S], p1)), __dmul_rn(W[7 * S], p2)), W[10 * S]);
  u = __ddiv_rn(x, z);
  v = __ddiv_rn(y, z);
  return fabs(u) < hw && fabs(v) < hh;
}

// Landmark-major pass: one warp per landmark, lanes over the cameras of the staged chunk.
// PASS 0 counts (lm_cur += count, cam_deg histogram); PASS 1 writes at lm_ptr[l] + lm_cur[l].
template <int PASS>
__global__ void __launch_bounds__(kVisThreads)
k_vis_lm(int n_cam, int n_lm, const double* __restrict__ W, const double* __restrict__ pts, double hw, double hh, int round_f32,
         int* __restrict__ lm_cur, int* __restrict__ cam_deg, const int64_t* __restrict__ lm_ptr, int* __restrict__ obs_cam,
         int* __restrict__ obs_lm, double* __restrict__ obs_uv) {
  extern __shared__ __align__(16) double s_W[];                 // plane-major [kW][kVisCamChunk]: lanes read consecutive words
  int* s_hist = reinterpret_cast<int*>(s_W + (size_t)kVisCamChunk * kW);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = kVisThreads / 32;
  for (int c0 = 0; c0 < n_cam; c0 += kVisCamChunk) {
    const int nc = min(kVisCamChunk, n_cam - c0);
    __syncthreads();
    for (int e = threadIdx.x; e < nc * kW; e += kVisThreads) s_W[(e % kW) * kVisCamChunk + e / kW] = W[(size_t)c0 * kW + e];
    if (PASS == 0) for (int e = threadIdx.x; e < nc; e += kVisThreads) s_hist[e] = 0;
    __syncthreads();
    for (int l = blockIdx.x * wpb + warp; l < n_lm; l += gridDim.x * wpb) {
      const double p0 = __ldg(pts + 3 * (size_t)l), p1 = __ldg(pts + 3 * (size_t)l + 1), p2 = __ldg(pts + 3 * (size_t)l + 2);
      int run = lm_cur[l];
      const int64_t base = PASS ? lm_ptr[l] : 0;
      for (int cb = 0; cb < nc; cb += 32) {
        const int c = cb + lane;
        double u = 0.0, v = 0.0;
        const bool vis = c < nc && visible<kVisCamChunk>(s_W + c, p0, p1, p2, hw, hh, u, v);
        const unsigned m = __ballot_sync(0xffffffffu, vis);
        if (vis) {
          if (PASS == 0) {
            atomicAdd(&s_hist[c], 1);
          } else {
            const int64_t o = base + run + __popc(m & ((1u << lane) - 1u));
            obs_cam[o] = c0 + c;
            obs_lm[o] = l;
            if (round_f32) { u = (double)(float)u; v = (double)(float)v; }      // pcl::PointXY, sim_data.cpp:135-136
            obs_uv[2 * o] = u;
            obs_uv[2 * o + 1] = v;
          }
        }
        run += __popc(m);
      }
      __syncwarp();
      if (lane == 0) lm_cur[l] = run;
    }
    if (PASS == 0) {
      __syncthreads();
      for (int e = threadIdx.x; e < nc; e += kVisThreads)
        if (s_hist[e]) atomicAdd(&cam_deg[c0 + e], s_hist[e]);
    }
  }
}

// Camera-major pass: one warp per camera, lanes over landmarks in ascending order -> cam_lm
__global__ void __launch_bounds__(kVisThreads)
k_vis_cam(int n_cam, int n_lm, const double* __restrict__ W, const double* __restrict__ pts, double hw, double hh,
          const int64_t* __restrict__ cam_ptr, int* __restrict__ cam_lm) {
  const int lane = threadIdx.x & 31, c = blockIdx.x * (kVisThreads / 32) + (threadIdx.x >> 5);
  if (c >= n_cam) return;
  double Wc[kW];
#pragma unroll
  for (int k = 0; k < kW; ++k) Wc[k] = __ldg(W + (size_t)kW * c + k);
  int64_t o = cam_ptr[c];
  for (int lb = 0; lb < n_lm; lb += 32) {
    const int l = lb + lane;
    double u, v;
    bool vis = false;
    if (l < n_lm) vis = visible<1>(Wc, __ldg(pts + 3 * (size_t)l), __ldg(pts + 3 * (size_t)l + 1), __ldg(pts + 3 * (size_t)l + 2), hw, hh, u, v);
    const unsigned m = __ballot_sync(0xffffffffu, vis);
    if (vis) cam_lm[o + __popc(m & ((1u << lane) - 1u))] = l;
    o += __popc(m);
  }
}

// single-block exclusive scan int32 -> int64 (n + 1 outputs); sizes here are <= a few million
__global__ void __launch_bounds__(1024) k_scan_i32_i64(int64_t n, const int* __restrict__ in, int64_t* __restrict__ out) {
  __shared__ int64_t s_warp[32];
  __shared__ int64_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int64_t b = 0; b < n; b += 1024) {
    const int64_t i = b + tid;
    const int64_t v = i < n ? in[i] : 0;
    int64_t x = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, x, s);
      if (lane >= s) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int64_t w = s_warp[lane];
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, w, s);
        if (lane >= s) w += y;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int64_t before = s_carry + (warp ? s_warp[warp - 1] : 0) + x - v;
    if (i < n) out[i] = before;
    __syncthreads();
    if (tid == 1023) s_carry = before + v;
    __syncthreads();
  }
  if (tid == 0) out[n] = s_carry;
}

__global__ void k_lm_hist(int64_t n, const int* __restrict__ obs_lm, int n_lm, int* __restrict__ deg, int* __restrict__ bad) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int l = obs_lm[i];
  if (l < 0 || l >= n_lm || (i + 1 < n && obs_lm[i + 1] < l)) { atomicExch(bad, 1); return; }
  atomicAdd(&deg[l], 1);
}

// ---- batched triangulation ---------------------------------------------------------------------
struct TriOpt {
  int max_it, max_invalid, jacobi;
  double r0, rmax, rmin, min_rel, dmin, dmax, ftol, gtol, ptol;
};

// cost (and optionally H = J^T J upper triangle, g = J^T r) of one landmark at P; residual
// uv - (R_cw P + t_cw).xy / z, sim_data.h:181-193
template <bool DERIV>
__device__ __forceinline__ double tri_eval(const double* __restrict__ W, const int* __restrict__ oc, const double* __restrict__ uv,
                                           int64_t beg, int64_t end, const double* P, double* H, double* g) {
  double cost = 0.0;
  if (DERIV) {
#pragma unroll
    for (int k = 0; k < 6; ++k) H[k] = 0.0;
    g[0] = g[1] = g[2] = 0.0;
  }
  for (int64_t o = beg; o < end; ++o) {
    const double* w = W + (size_t)kW * __ldg(oc + o);
    double T[kW];
#pragma unroll
    for (int k = 0; k < kW; ++k) T[k] = __ldg(w + k);
    const double x = fma(T[0], P[0], fma(T[3], P[1], fma(T[6], P[2], T[9])));
    const double y = fma(T[1], P[0], fma(T[4], P[1], fma(T[7], P[2], T[10])));
    const double z = fma(T[2], P[0], fma(T[5], P[1], fma(T[8], P[2], T[11])));
    const double iz = 1.0 / z, u = x * iz, v = y * iz;
    const double r0 = __ldg(uv + 2 * o) - u, r1 = __ldg(uv + 2 * o + 1) - v;
    cost = fma(r0, r0, fma(r1, r1, cost));
    if (DERIV) {
      double J0[3], J1[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {                 // J = -Pi' R_cw ; column k of R_cw = (T[3k], T[3k+1], T[3k+2])
        J0[k] = -iz * fma(-u, T[3 * k + 2], T[3 * k]);
        J1[k] = -iz * fma(-v, T[3 * k + 2], T[3 * k + 1]);
      }
      H[0] = fma(J0[0], J0[0], fma(J1[0], J1[0], H[0]));
      H[1] = fma(J0[0], J0[1], fma(J1[0], J1[1], H[1]));
      H[2] = fma(J0[0], J0[2], fma(J1[0], J1[2], H[2]));
      H[3] = fma(J0[1], J0[1], fma(J1[1], J1[1], H[3]));
      H[4] = fma(J0[1], J0[2], fma(J1[1], J1[2], H[4]));
      H[5] = fma(J0[2], J0[2], fma(J1[2], J1[2], H[5]));
#pragma unroll
      for (int k = 0; k < 3; ++k) g[k] = fma(J0[k], r0, fma(J1[k], r1, g[k]));
    }
  }
  return 0.5 * cost;
}

__device__ __forceinline__ double gmax_of(const double* x, const double* g) {
  double m = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) m = fmax(m, fabs(x[k] - (x[k] - g[k])));     // |x - Plus(x, -g)|_inf
  return m;
}

__global__ void __launch_bounds__(128)
k_triangulate(int n_lm, const double* __restrict__ W, const int64_t* __restrict__ lm_ptr, const int* __restrict__ obs_cam,
              const double* __restrict__ obs_uv, double* __restrict__ lm, TriOpt o, int* __restrict__ iters,
              double* __restrict__ final_cost, int* __restrict__ term) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lm) return;
  const int64_t beg = lm_ptr[l], end = lm_ptr[l + 1];
  double x[3] = {lm[3 * (size_t)l], lm[3 * (size_t)l + 1], lm[3 * (size_t)l + 2]};
  if (end == beg) { iters[l] = 0; final_cost[l] = 0.0; term[l] = STBA_CONVERGENCE; return; }
  double H[6], g[3], s[3];
  double x_cost = tri_eval<true>(W, obs_cam, obs_uv, beg, end, x, H, g);
  s[0] = o.jacobi ? 1.0 / (1.0 + sqrt(H[0])) : 1.0;
  s[1] = o.jacobi ? 1.0 / (1.0 + sqrt(H[3])) : 1.0;
  s[2] = o.jacobi ? 1.0 / (1.0 + sqrt(H[5])) : 1.0;
  double gmax = gmax_of(x, g);
  double x_norm = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  double radius = o.r0, dec = 2.0, diag[3] = {0.0, 0.0, 0.0};
  bool reuse = false, successful = true;
  int it = 0, n_rec = 0, n_invalid = 0, tt = STBA_NO_CONVERGENCE;
  while (true) {
    ++n_rec;
    if (it >= o.max_it) { tt = STBA_NO_CONVERGENCE; break; }
    if (successful && gmax <= o.gtol) { tt = STBA_CONVERGENCE; break; }
    if (radius < o.rmin) { tt = STBA_CONVERGENCE; break; }
    ++it;
    successful = false;
    // scaled system Hs = S H S, gs = S g
    const double a00 = s[0] * s[0] * H[0], a01 = s[0] * s[1] * H[1], a02 = s[0] * s[2] * H[2];
    const double a11 = s[1] * s[1] * H[3], a12 = s[1] * s[2] * H[4], a22 = s[2] * s[2] * H[5];
    const double gs0 = s[0] * g[0], gs1 = s[1] * g[1], gs2 = s[2] * g[2];
    if (!reuse) {
      diag[0] = fmin(fmax(a00, o.dmin), o.dmax);
      diag[1] = fmin(fmax(a11, o.dmin), o.dmax);
      diag[2] = fmin(fmax(a22, o.dmin), o.dmax);
    }
    const double d0 = diag[0] / radius, d1 = diag[1] / radius, d2 = diag[2] / radius;
    reuse = true;
    // 3x3 Cholesky of Hs + D^2
    bool valid = true;
    double y0 = 0, y1 = 0, y2 = 0, mcc = 0.0;
    {
      const double p00 = a00 + d0;
      valid = p00 > 0.0;
      const double l00 = sqrt(p00), l10 = a01 / l00, l20 = a02 / l00;
      const double p11 = (a11 + d1) - l10 * l10;
      valid = valid && p11 > 0.0;
      const double l11 = sqrt(p11), l21 = (a12 - l20 * l10) / l11;
      const double p22 = (a22 + d2) - l20 * l20 - l21 * l21;
      valid = valid && p22 > 0.0;
      const double l22 = sqrt(p22);
      const double z0 = gs0 / l00, z1 = (gs1 - l10 * z0) / l11, z2 = (gs2 - l20 * z0 - l21 * z1) / l22;
      y2 = z2 / l22;
      y1 = (z1 - l21 * y2) / l11;
      y0 = (z0 - l10 * y1 - l20 * y2) / l00;
      valid = valid && isfinite(y0) && isfinite(y1) && isfinite(y2);
      // model cost change of step = -y:  1/2 y^T (gs + D^2 y)
      mcc = 0.5 * (y0 * (gs0 + d0 * y0) + y1 * (gs1 + d1 * y1) + y2 * (gs2 + d2 * y2));
      valid = valid && mcc > 0.0;
    }
    if (!valid) {
      if (++n_invalid >= o.max_invalid) { tt = STBA_FAILURE; break; }
      radius /= dec;
      dec *= 2.0;
      reuse = false;
      continue;
    }
    n_invalid = 0;
    const double xc[3] = {x[0] - y0 * s[0], x[1] - y1 * s[1], x[2] - y2 * s[2]};
    const double cand = tri_eval<false>(W, obs_cam, obs_uv, beg, end, xc, nullptr, nullptr);
    const bool cand_ok = isfinite(cand);
    const double e0 = xc[0] - x[0], e1 = xc[1] - x[1], e2 = xc[2] - x[2];
    const double step_norm = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
    if (step_norm <= o.ptol * (x_norm + o.ptol)) { tt = STBA_CONVERGENCE; break; }
    if (cand_ok && fabs(x_cost - cand) <= o.ftol * x_cost) { tt = STBA_CONVERGENCE; break; }
    const double rho = cand_ok ? (x_cost - cand) / mcc : -1.7976931348623157e308;
    if (rho > o.min_rel) {
      x[0] = xc[0]; x[1] = xc[1]; x[2] = xc[2];
      x_norm = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
      x_cost = tri_eval<true>(W, obs_cam, obs_uv, beg, end, x, H, g);
      gmax = gmax_of(x, g);
      successful = true;
      const double t = 2.0 * rho - 1.0;
      radius = fmin(o.rmax, radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
      dec = 2.0;
      reuse = false;
    } else {
      radius /= dec;
      dec *= 2.0;
      reuse = true;
    }
  }
  lm[3 * (size_t)l] = x[0];
  lm[3 * (size_t)l + 1] = x[1];
  lm[3 * (size_t)l + 2] = x[2];
  iters[l] = n_rec;      // = summary.iterations.size(): an iteration that ends on a tolerance test leaves no record
  final_cost[l] = x_cost;
  term[l] = tt;
}

// range check of a device-resident obs_cam (host arrays are checked on the host)
__global__ void k_cam_range(int64_t n, const int* __restrict__ oc, int n_cam, int* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (oc[i] < 0 || oc[i] >= n_cam)) *bad = 1;
}

}  // namespace

extern "C" {

int stba_visibility(int device, int32_t n_cam, int32_t n_lm, const double* cam_q, const double* cam_t, const double* pts,
                    double half_w, double half_h, int32_t round_uv_f32, int64_t capacity, int64_t* n_obs, int32_t* lm_deg,
                    int32_t* cam_deg, int32_t* obs_cam, int32_t* obs_lm, double* obs_uv, int32_t* cam_lm) {
  if (n_cam < 0 || n_lm < 0 || !n_obs || (n_cam && (!cam_q || !cam_t)) || (n_lm && !pts)) return STBA_ERR_INVALID_ARGUMENT;
  DevBuf b;
  const int st = open_device(device, b);
  if (st != STBA_OK) return st;
  *n_obs = 0;
  if (n_cam == 0 || n_lm == 0) {
    if (lm_deg) for (int i = 0; i < n_lm; ++i) lm_deg[i] = 0;
    if (cam_deg) for (int i = 0; i < n_cam; ++i) cam_deg[i] = 0;
    return STBA_OK;
  }
  double *d_q, *d_t, *d_p, *d_W, *d_uv = nullptr;
  int *d_cur, *d_cdeg, *d_oc = nullptr, *d_ol = nullptr, *d_cl = nullptr;
  int64_t *d_lptr, *d_cptr;
  CK(b.get(&d_q, 4 * (size_t)n_cam)); CK(b.get(&d_t, 3 * (size_t)n_cam)); CK(b.get(&d_p, 3 * (size_t)n_lm));
  CK(b.get(&d_W, (size_t)kW * n_cam)); CK(b.get(&d_cur, (size_t)n_lm)); CK(b.get(&d_cdeg, (size_t)n_cam));
  CK(b.get(&d_lptr, (size_t)n_lm + 1)); CK(b.get(&d_cptr, (size_t)n_cam + 1));
  CK(cudaMemcpyAsync(d_q, cam_q, 4 * (size_t)n_cam * sizeof(double), cudaMemcpyDefault, b.s));
  CK(cudaMemcpyAsync(d_t, cam_t, 3 * (size_t)n_cam * sizeof(double), cudaMemcpyDefault, b.s));
  CK(cudaMemcpyAsync(d_p, pts, 3 * (size_t)n_lm * sizeof(double), cudaMemcpyDefault, b.s));
  CK(cudaMemsetAsync(d_cur, 0, (size_t)n_lm * sizeof(int), b.s));
  CK(cudaMemsetAsync(d_cdeg, 0, (size_t)n_cam * sizeof(int), b.s));
  k_world_to_camera<<<(n_cam + 127) / 128, 128, 0, b.s>>>(n_cam, d_q, d_t, d_W, 1);
  static int sm_count_of[64] = {0};       // cudaGetDeviceProperties costs milliseconds: one attribute query per device and process
  if (device < 64 && !sm_count_of[device]) CK(cudaDeviceGetAttribute(&sm_count_of[device], cudaDevAttrMultiProcessorCount, device));
  struct { int multiProcessorCount; } prop = {device < 64 ? sm_count_of[device] : 148};
  const int smem = kVisCamChunk * kW * (int)sizeof(double) + kVisCamChunk * (int)sizeof(int);
  CK(cudaFuncSetAttribute(k_vis_lm<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(k_vis_lm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int wpb = kVisThreads / 32;
  const int grid = std::max(1, std::min((n_lm + wpb - 1) / wpb, 2 * prop.multiProcessorCount));
  k_vis_lm<0><<<grid, kVisThreads, smem, b.s>>>(n_cam, n_lm, d_W, d_p, half_w, half_h, round_uv_f32, d_cur, d_cdeg, nullptr, nullptr,
                                                nullptr, nullptr);
  k_scan_i32_i64<<<1, 1024, 0, b.s>>>(n_lm, d_cur, d_lptr);
  k_scan_i32_i64<<<1, 1024, 0, b.s>>>(n_cam, d_cdeg, d_cptr);
  int64_t total = 0;
  CK(cudaMemcpyAsync(&total, d_lptr + n_lm, sizeof(int64_t), cudaMemcpyDeviceToHost, b.s));
  if (lm_deg) CK(cudaMemcpyAsync(lm_deg, d_cur, (size_t)n_lm * sizeof(int), cudaMemcpyDefault, b.s));
  if (cam_deg) CK(cudaMemcpyAsync(cam_deg, d_cdeg, (size_t)n_cam * sizeof(int), cudaMemcpyDefault, b.s));
  CK(cudaStreamSynchronize(b.s));
  *n_obs = total;
  const bool want_lists = obs_cam || obs_lm || obs_uv || cam_lm;
  if (!want_lists) return STBA_OK;
  if (capacity < total) return STBA_ERR_OVERFLOW;
  if (total == 0) return STBA_OK;
  CK(b.get(&d_oc, (size_t)total)); CK(b.get(&d_ol, (size_t)total)); CK(b.get(&d_uv, 2 * (size_t)total)); CK(b.get(&d_cl, (size_t)total));
  CK(cudaMemsetAsync(d_cur, 0, (size_t)n_lm * sizeof(int), b.s));
  k_vis_lm<1><<<grid, kVisThreads, smem, b.s>>>(n_cam, n_lm, d_W, d_p, half_w, half_h, round_uv_f32, d_cur, d_cdeg, d_lptr, d_oc, d_ol,
                                                d_uv);
  if (cam_lm) k_vis_cam<<<(n_cam + wpb - 1) / wpb, kVisThreads, 0, b.s>>>(n_cam, n_lm, d_W, d_p, half_w, half_h, d_cptr, d_cl);
  CK(cudaGetLastError());
  if (obs_cam) CK(cudaMemcpyAsync(obs_cam, d_oc, (size_t)total * sizeof(int), cudaMemcpyDefault, b.s));
  if (obs_lm) CK(cudaMemcpyAsync(obs_lm, d_ol, (size_t)total * sizeof(int), cudaMemcpyDefault, b.s));
  if (obs_uv) CK(cudaMemcpyAsync(obs_uv, d_uv, 2 * (size_t)total * sizeof(double), cudaMemcpyDefault, b.s));
  if (cam_lm) CK(cudaMemcpyAsync(cam_lm, d_cl, (size_t)total * sizeof(int), cudaMemcpyDefault, b.s));
  CK(cudaStreamSynchronize(b.s));
  return STBA_OK;
}

int stba_triangulate(int device, int32_t n_cam, int32_t n_lm, int64_t n_obs, const double* cam_q, const double* cam_t, double* lm,
                     const int32_t* obs_cam, const int32_t* obs_lm, const double* obs_uv, const stba_options* opt,
                     int32_t* iterations, double* final_cost, int32_t* termination, float* kernel_ms) {
  if (n_cam < 0 || n_lm < 0 || n_obs < 0 || (n_lm && !lm) || (n_obs && (!obs_cam || !obs_lm || !obs_uv || !cam_q || !cam_t)))
    return STBA_ERR_INVALID_ARGUMENT;
  // Array arguments may live in host or in device memory (unified addressing: the copies below are
  // cudaMemcpyDefault).  Host arrays are range-checked here, device arrays by k_cam_range on the device.
  bool oc_on_device = false;
  if (n_obs) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, obs_cam) == cudaSuccess) oc_on_device = pa.type == cudaMemoryTypeDevice;
    else cudaGetLastError();
  }
  if (!oc_on_device)
    for (int64_t i = 0; i < n_obs; ++i)
      if (obs_cam[i] < 0 || obs_cam[i] >= n_cam) return STBA_ERR_INVALID_ARGUMENT;
  DevBuf b;
  const int st = open_device(device, b);
  if (st != STBA_OK) return st;
  if (n_lm == 0) return STBA_OK;
  stba_options o;
  if (opt) o = *opt; else stba_options_init(&o);
  double *d_q, *d_t, *d_W, *d_lm, *d_uv, *d_cost;
  int *d_oc, *d_ol, *d_deg, *d_bad, *d_it, *d_term;
  int64_t* d_ptr;
  CK(b.get(&d_q, 4 * (size_t)n_cam)); CK(b.get(&d_t, 3 * (size_t)n_cam)); CK(b.get(&d_W, (size_t)kW * n_cam));
  CK(b.get(&d_lm, 3 * (size_t)n_lm)); CK(b.get(&d_uv, 2 * (size_t)n_obs)); CK(b.get(&d_cost, (size_t)n_lm));
  CK(b.get(&d_oc, (size_t)n_obs)); CK(b.get(&d_ol, (size_t)n_obs)); CK(b.get(&d_deg, (size_t)n_lm)); CK(b.get(&d_bad, 1));
  CK(b.get(&d_it, (size_t)n_lm)); CK(b.get(&d_term, (size_t)n_lm)); CK(b.get(&d_ptr, (size_t)n_lm + 1));
  if (n_cam) {
    CK(cudaMemcpyAsync(d_q, cam_q, 4 * (size_t)n_cam * sizeof(double), cudaMemcpyDefault, b.s));
    CK(cudaMemcpyAsync(d_t, cam_t, 3 * (size_t)n_cam * sizeof(double), cudaMemcpyDefault, b.s));
  }
  CK(cudaMemcpyAsync(d_lm, lm, 3 * (size_t)n_lm * sizeof(double), cudaMemcpyDefault, b.s));
  if (n_obs) {
    CK(cudaMemcpyAsync(d_oc, obs_cam, (size_t)n_obs * sizeof(int), cudaMemcpyDefault, b.s));
    CK(cudaMemcpyAsync(d_ol, obs_lm, (size_t)n_obs * sizeof(int), cudaMemcpyDefault, b.s));
    CK(cudaMemcpyAsync(d_uv, obs_uv, 2 * (size_t)n_obs * sizeof(double), cudaMemcpyDefault, b.s));
  }
  CK(cudaMemsetAsync(d_deg, 0, (size_t)n_lm * sizeof(int), b.s));
  CK(cudaMemsetAsync(d_bad, 0, sizeof(int), b.s));
  if (n_cam) k_world_to_camera<<<(n_cam + 127) / 128, 128, 0, b.s>>>(n_cam, d_q, d_t, d_W, 0);
  if (n_obs && oc_on_device) k_cam_range<<<(unsigned)((n_obs + 255) / 256), 256, 0, b.s>>>(n_obs, d_oc, n_cam, d_bad);
  if (n_obs) k_lm_hist<<<(unsigned)((n_obs + 255) / 256), 256, 0, b.s>>>(n_obs, d_ol, n_lm, d_deg, d_bad);
  k_scan_i32_i64<<<1, 1024, 0, b.s>>>(n_lm, d_deg, d_ptr);
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, b.s));
  CK(cudaStreamSynchronize(b.s));
  if (bad) return STBA_ERR_INVALID_ARGUMENT;       // observations must be landmark-major (sim_data.cpp:298-306)
  TriOpt t;
  t.max_it = o.max_num_iterations; t.max_invalid = o.max_num_consecutive_invalid_steps; t.jacobi = o.jacobi_scaling;
  t.r0 = o.initial_trust_region_radius; t.rmax = o.max_trust_region_radius; t.rmin = o.min_trust_region_radius;
  t.min_rel = o.min_relative_decrease; t.dmin = o.min_lm_diagonal; t.dmax = o.max_lm_diagonal;
  t.ftol = o.function_tolerance; t.gtol = o.gradient_tolerance; t.ptol = o.parameter_tolerance;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, b.s));
  k_triangulate<<<(n_lm + 127) / 128, 128, 0, b.s>>>(n_lm, d_W, d_ptr, d_oc, d_uv, d_lm, t, d_it, d_cost, d_term);
  CK(cudaEventRecord(e1, b.s));
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(lm, d_lm, 3 * (size_t)n_lm * sizeof(double), cudaMemcpyDefault, b.s));
  if (iterations) CK(cudaMemcpyAsync(iterations, d_it, (size_t)n_lm * sizeof(int), cudaMemcpyDefault, b.s));
  if (final_cost) CK(cudaMemcpyAsync(final_cost, d_cost, (size_t)n_lm * sizeof(double), cudaMemcpyDefault, b.s));
  if (termination) CK(cudaMemcpyAsync(termination, d_term, (size_t)n_lm * sizeof(int), cudaMemcpyDefault, b.s));
  CK(cudaStreamSynchronize(b.s));
  if (kernel_ms) cudaEventElapsedTime(kernel_ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return STBA_OK;
}

}  // extern "C"
