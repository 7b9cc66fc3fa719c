// Linearisation kernels, second generation: the contract kernel of BASELINE.json
// (per-observation reprojection residual + exact Jacobian + J^T J / J^T r block accumulation, J never
// stored; algorithmic bytes 24 N_obs + 96 N_lm + 272 N_cam).
//
// lin_lm2 — landmark-major pass, persistent CTAs (one per SM, 512 threads):
//   * the whole camera-tile table ([R|t], 112 B per camera) is staged ONCE per CTA into shared memory
//     with a TMA bulk copy (cp.async.bulk, SASS UBLKCP) when it fits (n_cam <= kMaxSmemCams);
//   * the observation stream (obs_cam i32 + obs_uv 2 x f64) of each chunk of 256 consecutive
//     landmarks is contiguous; it is double-buffered into shared memory by TMA bulk copies signalled
//     on mbarriers, issued one chunk ahead by thread 0;
//   * two threads per landmark then walk alternate observations out of shared memory, keep their
//     halves of H_ll (6) and g_l (3) in registers, combine them with one shuffle step and write
//     once: no global-memory latency inside the loop, no atomics, deterministic.
// lin_cam2 — camera-major pass, one warp per CTA per chunk of one camera's observations: four
//   observations per lane in flight (index -> landmark gathers are the latency), camera-frame
//   accumulation (23 sums, see stba_kernels.cuh), plain butterfly shuffles (a one-warp CTA is
//   provably convergent, so ptxas emits SHFL.BFLY instead of WARPSYNC.COLLECTIVE call sequences).
#pragma once
#include "stba_kernels.cuh"

namespace stba {

constexpr int kLinLm = 256;            // landmarks per chunk
constexpr int kLinThreads = 512;       // two threads per landmark (alternate observations, combined by one shuffle step)
constexpr int kStageObs = 2688;        // observation capacity of one stage (multiple of 4)
constexpr int kMaxSmemCams = 1024;     // camera tiles that fit next to two stages
constexpr int kStageBytes = kStageObs * 16 + kStageObs * 4;
constexpr int kLinSmemBytes = 64 + 2 * kStageBytes + kMaxSmemCams * kCamTile * 8;   // barriers first

#ifdef STBA_LIN_TIMING
__device__ long long g_lin_clk[64];
#define LTICK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_lin_clk[i] = clock64(); } while (0)
#else
#define LTICK(i) do {} while (0)
#endif

template <bool COST_ONLY, bool CAM_SMEM>
__global__ void __launch_bounds__(kLinThreads, 1)
k_lin_lm2(int n_lm, int n_cam, const int* __restrict__ lm_ptr, const int* __restrict__ obs_cam,
          const double* __restrict__ obs_uv, const double* __restrict__ Rt, const double* __restrict__ lm4,
          double* __restrict__ Hll, double* __restrict__ gl, double* partial, unsigned int* counter,
          double* out_cost) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw);   // [0],[1]: stages, [2]: cameras
  unsigned char* stage_base = smem_raw + 64;
  double* s_cam_tiles = reinterpret_cast<double*>(smem_raw + 64 + 2 * kStageBytes);
  __shared__ int s_bounds[2][2];      // per stage: first staged observation (aligned down to 4), count or -1 (not staged)
  const int tid = threadIdx.x;
  const int n_chunks = (n_lm + kLinLm - 1) / kLinLm;
  LTICK(0);

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // producer: stage the observations of `chunk` into buffer `st`
  auto issue = [&](int chunk, int st) {
    const int l0 = chunk * kLinLm, l1 = min(l0 + kLinLm, n_lm);
    const int o0 = lm_ptr[l0] & ~3, o1 = lm_ptr[l1];
    const int cnt = o1 - o0;
    if (cnt > kStageObs || cnt <= 0) {           // too many observations for a stage: read them from global
      s_bounds[st][0] = o0;
      s_bounds[st][1] = -1;
      return;
    }
    s_bounds[st][0] = o0;
    s_bounds[st][1] = cnt;
    unsigned char* base = stage_base + st * kStageBytes;
    const unsigned uv_bytes = (unsigned)cnt * 16u, cam_bytes = ((unsigned)cnt * 4u + 15u) & ~15u;
    mbar_expect_tx(&bars[st], uv_bytes + cam_bytes);
    bulk_g2s(base, obs_uv + 2 * (size_t)o0, uv_bytes, &bars[st]);
    bulk_g2s(base + kStageObs * 16, obs_cam + o0, cam_bytes, &bars[st]);
  };

  if (tid == 0) {
    if (CAM_SMEM) {
      const unsigned bytes = (unsigned)n_cam * kCamTile * 8u;
      mbar_expect_tx(&bars[2], bytes);
      // one bulk copy is limited in size by the tx-count width; split into <= 32 KB pieces
      for (unsigned off = 0; off < bytes; off += 32768u)
        bulk_g2s(reinterpret_cast<unsigned char*>(s_cam_tiles) + off, reinterpret_cast<const unsigned char*>(Rt) + off,
                 min(32768u, bytes - off), &bars[2]);
    }
    if ((int)blockIdx.x < n_chunks) issue(blockIdx.x, 0);
  }
  __syncthreads();
  LTICK(1);
  if (CAM_SMEM) mbar_wait(&bars[2], 0);
  LTICK(2);

  double cost[1] = {0.0};
  int it = 0;
  unsigned phase0 = 0, phase1 = 0;      // completed uses of each stage barrier (uniform across the CTA)
  for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++it) {
    const int st = it & 1;
    // prefetch the next chunk into the other stage (its previous contents were consumed before the
    // __syncthreads that ended the previous iteration)
    if (tid == 0 && chunk + (int)gridDim.x < n_chunks) issue(chunk + gridDim.x, st ^ 1);
    const int o_base = s_bounds[st][0], staged = s_bounds[st][1];
    LTICK(3 + 4 * it);
    if (staged > 0) {
      mbar_wait(&bars[st], (st ? phase1 : phase0) & 1);
      if (st) ++phase1; else ++phase0;
    }
    LTICK(4 + 4 * it);
    const double2* s_uv = reinterpret_cast<const double2*>(stage_base + st * kStageBytes);
    const int* s_oc = reinterpret_cast<const int*>(stage_base + st * kStageBytes + kStageObs * 16);

    // lanes 2i and 2i+1 share landmark l and take alternate observations
    const int l = chunk * kLinLm + (tid >> 1), par = tid & 1;
    double h0 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0, h5 = 0, g0 = 0, g1 = 0, g2 = 0;
    if (l < n_lm) {
      const double2 pxy = ldg2(lm4 + 4 * (size_t)l);
      const double pz = __ldg(lm4 + 4 * (size_t)l + 2);
      const int beg = lm_ptr[l], end = lm_ptr[l + 1];
      // two observations per trip, as two independent chains (index -> tile -> projection -> 1/z ->
      // Jacobian): with 16 warps per SM the latency of ONE such chain per thread left the FP64 pipe at
      // 20 % (profiles/r1_lin_full.md); the second chain fills its bubbles.
      auto fetch = [&](int o, int& c, double2& uv) {
        if (staged > 0) {
          c = s_oc[o - o_base];
          uv = s_uv[o - o_base];
        } else {
          c = __ldg(obs_cam + o);
          uv = ldg2(obs_uv + 2 * (size_t)o);
        }
      };
      auto tile_of = [&](int c, double* T) {
        if (CAM_SMEM) {
          const double2* tile = reinterpret_cast<const double2*>(s_cam_tiles + (size_t)kCamTile * c);
#pragma unroll
          for (int k = 0; k < kCamVals / 2; ++k) {
            const double2 x = tile[k];
            T[2 * k] = x.x;
            T[2 * k + 1] = x.y;
          }
        } else {
#pragma unroll
          for (int k = 0; k < kCamVals; k += 2) {
            const double2 x = ldg2(Rt + (size_t)kCamTile * c + k);
            T[k] = x.x;
            T[k + 1] = x.y;
          }
        }
      };
      auto accumulate = [&](const double* T, const Obs& ob) {
        cost[0] = fma(ob.r0, ob.r0, fma(ob.r1, ob.r1, cost[0]));
        if (!COST_ONLY) {
          double J0[3], J1[3];
          landmark_jacobian(T, ob, J0, J1);
          h0 = fma(J0[0], J0[0], fma(J1[0], J1[0], h0));
          h1 = fma(J0[0], J0[1], fma(J1[0], J1[1], h1));
          h2 = fma(J0[0], J0[2], fma(J1[0], J1[2], h2));
          h3 = fma(J0[1], J0[1], fma(J1[1], J1[1], h3));
          h4 = fma(J0[1], J0[2], fma(J1[1], J1[2], h4));
          h5 = fma(J0[2], J0[2], fma(J1[2], J1[2], h5));
          g0 = fma(J0[0], ob.r0, fma(J1[0], ob.r1, g0));
          g1 = fma(J0[1], ob.r0, fma(J1[1], ob.r1, g1));
          g2 = fma(J0[2], ob.r0, fma(J1[2], ob.r1, g2));
        }
      };
      int o = beg + par;
      for (; o + 2 < end; o += 4) {
        int ca, cb;
        double2 uva, uvb;
        fetch(o, ca, uva);
        fetch(o + 2, cb, uvb);
        double Ta[kCamVals], Tb[kCamVals];
        tile_of(ca, Ta);
        tile_of(cb, Tb);
        const Obs oa = project(Ta, pxy.x, pxy.y, pz, uva.x, uva.y);
        const Obs ob = project(Tb, pxy.x, pxy.y, pz, uvb.x, uvb.y);
        accumulate(Ta, oa);
        accumulate(Tb, ob);
      }
      if (o < end) {
        int c;
        double2 uv;
        fetch(o, c, uv);
        double T[kCamVals];
        tile_of(c, T);
        const Obs ob = project(T, pxy.x, pxy.y, pz, uv.x, uv.y);
        accumulate(T, ob);
      }
    }
    LTICK(5 + 4 * it);
    if (!COST_ONLY) {
      // combine the two half sums (fixed order: even lane + odd lane); every lane of the CTA gets here
      h0 += __shfl_xor_sync(0xffffffffu, h0, 1); h1 += __shfl_xor_sync(0xffffffffu, h1, 1);
      h2 += __shfl_xor_sync(0xffffffffu, h2, 1); h3 += __shfl_xor_sync(0xffffffffu, h3, 1);
      h4 += __shfl_xor_sync(0xffffffffu, h4, 1); h5 += __shfl_xor_sync(0xffffffffu, h5, 1);
      g0 += __shfl_xor_sync(0xffffffffu, g0, 1); g1 += __shfl_xor_sync(0xffffffffu, g1, 1);
      g2 += __shfl_xor_sync(0xffffffffu, g2, 1);
      if (l < n_lm) {
        if (par == 0) {
          double2* H = reinterpret_cast<double2*>(Hll + 6 * (size_t)l);
          H[0] = make_double2(h0, h1);
          H[1] = make_double2(h2, h3);
          H[2] = make_double2(h4, h5);
        } else {
          gl[3 * (size_t)l] = g0;
          gl[3 * (size_t)l + 1] = g1;
          gl[3 * (size_t)l + 2] = g2;
        }
      }
    }
    __syncthreads();     // everyone is done with stage `st` and with s_bounds[st]
    LTICK(6 + 4 * it);
  }
  cost[0] *= 0.5;
  LTICK(30);
  grid_reduce<1, kLinThreads>(cost, 1, partial, counter, out_cost);
  LTICK(31);
}

// ---------------------------------------------------------------------------------------------
// lin_cam2: grid = n_chunk CTAs of ONE warp.  Same 23 camera-frame sums as k_lin_cam.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cam_accumulate(const double* __restrict__ T, double px, double py, double pz,
                                               double u0, double v0, double (&acc)[kCamAcc]) {
  const Obs ob = project(T, px, py, pz, u0, v0);
  const double u = ob.u, v = ob.v, iz = ob.iz, r0 = ob.r0, r1 = ob.r1;
  const double a = u * v, b = fma(u, u, 1.0), c2 = fma(v, v, 1.0);
  acc[0] = fma(a, a, fma(c2, c2, acc[0]));
  acc[1] = fma(-a, b + c2, acc[1]);
  acc[2] = fma(a, v, fma(-c2, u, acc[2]));
  acc[3] = fma(b, b, fma(a, a, acc[3]));
  acc[4] = fma(a, u, fma(-b, v, acc[4]));
  acc[5] = fma(u, u, fma(v, v, acc[5]));
  const double aucv = fma(a, u, c2 * v), buav = fma(b, u, a * v);
  acc[6] = fma(iz, a, acc[6]);
  acc[7] = fma(iz, c2, acc[7]);
  acc[8] = fma(iz, b, acc[8]);
  acc[9] = fma(iz, aucv, acc[9]);
  acc[10] = fma(iz, buav, acc[10]);
  acc[11] = fma(iz, v, acc[11]);
  acc[12] = fma(iz, u, acc[12]);
  const double w = iz * iz;
  acc[13] += w;
  acc[14] = fma(w, u, acc[14]);
  acc[15] = fma(w, v, acc[15]);
  acc[16] = fma(w, b + c2 - 2.0, acc[16]);
  acc[17] = fma(a, r0, fma(c2, r1, acc[17]));
  acc[18] = fma(-b, r0, fma(-a, r1, acc[18]));
  acc[19] = fma(v, r0, fma(-u, r1, acc[19]));
  const double s = fma(u, r0, v * r1);
  acc[20] = fma(iz, r0, acc[20]);
  acc[21] = fma(iz, r1, acc[21]);
  acc[22] = fma(-iz, s, acc[22]);
}

// rotate the 23 camera-frame sums of one camera into the world-frame tangent and pack:
// H_tt = R Q R^T, H_th,t = -K R^T, g_t = -R m  (one lane does it; ~200 flops per camera)
__device__ __forceinline__ void cam_finish(const double* __restrict__ a, const double* __restrict__ R,
                                           double* __restrict__ Hcc_c, double* __restrict__ gc_c) {
  double H[6][6];
  H[0][0] = a[0]; H[0][1] = a[1]; H[0][2] = a[2]; H[1][1] = a[3]; H[1][2] = a[4]; H[2][2] = a[5];
  const double K[3][3] = {{a[6], a[7], -a[9]}, {-a[8], -a[6], a[10]}, {a[11], -a[12], 0.0}};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      H[i][3 + j] = -(K[i][0] * R[3 * j] + K[i][1] * R[3 * j + 1] + K[i][2] * R[3 * j + 2]);
  const double Q[3][3] = {{a[13], 0.0, -a[14]}, {0.0, a[13], -a[15]}, {-a[14], -a[15], a[16]}};
  double RQ[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      RQ[i][j] = R[3 * i] * Q[0][j] + R[3 * i + 1] * Q[1][j] + R[3 * i + 2] * Q[2][j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j)
      H[3 + i][3 + j] = RQ[i][0] * R[3 * j] + RQ[i][1] * R[3 * j + 1] + RQ[i][2] * R[3 * j + 2];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 6; ++j) Hcc_c[tri6(i, j)] = H[i][j];
  gc_c[0] = a[17];
  gc_c[1] = a[18];
  gc_c[2] = a[19];
#pragma unroll
  for (int i = 0; i < 3; ++i) gc_c[3 + i] = -(R[3 * i] * a[20] + R[3 * i + 1] * a[21] + R[3 * i + 2] * a[22]);
}

// The chunk of a camera that finishes LAST (per-camera ticket) adds that camera's chunk partials in
// chunk order and writes H_cc / g_c: deterministic, no second launch.
__global__ void __launch_bounds__(32)
k_lin_cam2(const int* __restrict__ chunk_cam, const int* __restrict__ chunk_beg, const int* __restrict__ chunk_end,
           const int* __restrict__ cam_chunk_ptr, const int* __restrict__ cobs_lm, const double* __restrict__ cobs_uv,
           const double* __restrict__ Rt, const double* __restrict__ lm4, double* __restrict__ chunk_acc,
           unsigned int* __restrict__ cam_ticket, double* __restrict__ Hcc, double* __restrict__ gc) {
  __shared__ double s_sum[kCamAcc];
  __shared__ int s_last;
  const int ch = blockIdx.x, lane = threadIdx.x;
  const int c = chunk_cam[ch];
  double T[kCamVals];
#pragma unroll
  for (int k = 0; k < kCamVals; k += 2) {
    const double2 x = ldg2(Rt + (size_t)kCamTile * c + k);
    T[k] = x.x;
    T[k + 1] = x.y;
  }
  double acc[kCamAcc];
#pragma unroll
  for (int k = 0; k < kCamAcc; ++k) acc[k] = 0.0;
  const int beg = chunk_beg[ch], end = chunk_end[ch];
  constexpr int U = 4;                      // observations per lane in flight
  for (int o0 = beg + lane; o0 < end; o0 += 32 * U) {
    int l[U];
    double2 uv[U], pxy[U];
    double pz[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int o = o0 + 32 * k;
      l[k] = o < end ? __ldg(cobs_lm + o) : -1;
      uv[k] = o < end ? ldg2(cobs_uv + 2 * (size_t)o) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (l[k] >= 0) {
        pxy[k] = ldg2(lm4 + 4 * (size_t)l[k]);
        pz[k] = __ldg(lm4 + 4 * (size_t)l[k] + 2);
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k)
      if (l[k] >= 0) cam_accumulate(T, pxy[k].x, pxy[k].y, pz[k], uv[k].x, uv[k].y, acc);
  }
#pragma unroll
  for (int k = 0; k < kCamAcc; ++k) {
    const double r = warp_sum(acc[k]);
    if (lane == 0) chunk_acc[(size_t)ch * kCamAcc + k] = r;
  }
  // ---- per-camera ticket: the last chunk to arrive finishes the camera ----
  const int ch0 = cam_chunk_ptr[c], ch1 = cam_chunk_ptr[c + 1];
  if (lane == 0) {
    __threadfence();
    const unsigned int t = atomicInc(cam_ticket + c, (unsigned int)(ch1 - ch0 - 1));   // wraps to 0: re-armed
    s_last = (t == (unsigned int)(ch1 - ch0 - 1));
  }
  __syncwarp();
  if (!s_last) return;
  __threadfence();
  if (lane < kCamAcc) {
    const volatile double* pa = chunk_acc;
    double r = 0.0;
    for (int k = ch0; k < ch1; ++k) r += pa[(size_t)k * kCamAcc + lane];
    s_sum[lane] = r;
  }
  __syncwarp();
  if (lane == 0) cam_finish(s_sum, T, Hcc + (size_t)c * 21, gc + (size_t)c * 6);
}

}  // namespace stba
