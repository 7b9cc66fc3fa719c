// Linearisation, third generation: ONE launch for camera tiles + landmark-major pass + camera-major pass
// (the contract kernel of BASELINE.json: per-observation reprojection residual + exact Jacobian + J^T J / J^T r
// block accumulation, J never stored; algorithmic bytes 24 N_obs + 96 N_lm + 272 N_cam).
//
// Rounds 1-2 ran three launches (k_cam_prep, k_lin_lm2, k_lin_cam2) back to back: each pass alone left the FP64
// pipe at 20-27 % (profiles/r2_lin_full.md: 16 warps per SM, long/short-scoreboard stalls), and they could not
// overlap because lin_lm2 owned the SM (222 KB of shared memory, 512 x 110 registers).  Here the two passes run
// SIDE BY SIDE on every SM, as warp-specialised halves of one persistent 512-thread CTA:
//   * warps 0-7, the landmark group: contiguous ranges of 128-landmark chunks per CTA.  EVERYTHING a chunk reads —
//     its observation stream (uv + camera index), its lm_ptr slice and its landmark points — is staged by TMA bulk
//     copies (cp.async.bulk + mbarrier, one chunk ahead, issued by thread 0 from chunk-table entries it loaded a whole
//     chunk earlier), the camera tiles [R | t] are built in shared memory from (q, t) by the group itself; the inner
//     loop touches no global memory.  With more than 1024 cameras the table holds a WINDOW of 1024 consecutive
//     cameras, re-staged when a chunk's camera range (chunk table, built once per problem) leaves it; a chunk that
//     spans more than a window rebuilds R from (q, t) per observation.
//   * warps 8-15, the camera group: one-warp chunks of one camera's observations, pulled from a self re-arming
//     ticket counter (results do not depend on who processes which chunk: chunk partials are summed in chunk order by
//     the camera's last arriver); the stream two rounds ahead is prefetched into L2; camera-frame accumulation
//     (23 sums) and a transposing butterfly (24 instead of 115 shuffles per chunk, bit-identical to warp_sum).
//   * landmark warps join the camera queue when their range is done: the static split needs no tuning.
// No atomics on floating-point data, fixed summation orders: bit-reproducible.
#pragma once
#include "stba_lin.cuh"

namespace stba {

constexpr int kL3Threads = 512;
constexpr int kL3LmThreads = 256;
constexpr int kL3Lm = 128;                                   // landmarks per chunk, two threads each
constexpr int kL3StageObs = 1536;                            // observation capacity of one stage (multiple of 4)
constexpr int kL3OffCam = kL3StageObs * 16;                  // i32 camera indices behind the uv pairs
constexpr int kL3OffPtr = kL3OffCam + kL3StageObs * 4;       // lm_ptr slice (kL3Lm + 1 ints, rounded up to 16 B)
constexpr int kL3OffLm = kL3OffPtr + (kL3Lm + 4) * 4;        // landmark points (32 B each)
constexpr int kL3StageBytes = (kL3OffLm + kL3Lm * 32 + 127) / 128 * 128;
constexpr int kL3MaxCams = 1024;
constexpr int kL3Head = 128;                                 // mbarriers + the two stage descriptors
constexpr int kL3SmemBytes = kL3Head + 2 * kL3StageBytes + kL3MaxCams * kCamTile * 8;

struct L3Params {
  int n_lm, n_cam, n_lm_chunks, n_cam_chunks;
  const int* lm_ptr; const int* obs_cam; const double* obs_uv; const double* lm4;
  const double* cam_q; const double* cam_t;
  const int4* ctab;                 // per landmark chunk: first staged observation (aligned down to 4), end, camera range
  double* Rt;                       // out: camera tiles for the kernels that follow
  double* Hll; double* gl;
  double* partial; unsigned int* counter; double* out_cost;
  const int* chunk_cam; const int* chunk_beg; const int* chunk_end; const int* cam_chunk_ptr;
  const int* cobs_lm; const double* cobs_uv;
  double* chunk_acc; unsigned int* cam_ticket; double* Hcc; double* gc;
  unsigned int* cam_counter;        // ticket counter of the camera-chunk queue (wraps to 0 after the last pull)
};

// chunk table of the landmark group: one warp per chunk
__global__ void k_l3_chunk_table(int n_lm, int n_chunks, const int* __restrict__ lm_ptr, const int* __restrict__ obs_cam,
                                 int4* __restrict__ ctab) {
  const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (chunk >= n_chunks) return;
  const int l0 = chunk * kL3Lm, l1 = min(l0 + kL3Lm, n_lm);
  const int ob = lm_ptr[l0], oe = lm_ptr[l1];
  int lo = 0x7fffffff, hi = -1;
  for (int o = ob + lane; o < oe; o += 32) {
    const int c = __ldg(obs_cam + o);
    lo = min(lo, c);
    hi = max(hi, c);
  }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if (lane == 0) ctab[chunk] = make_int4(ob & ~3, oe, hi < 0 ? 0 : lo, hi);
}

__device__ __forceinline__ void l3_lm_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kL3LmThreads) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// [R row-major | t] of camera c from its quaternion and position
__device__ __forceinline__ void tile_from_qt(const double* __restrict__ q, const double* __restrict__ t, int c, double* T) {
  const double2 qa = ldg2(q + 4 * (size_t)c), qb = ldg2(q + 4 * (size_t)c + 2);
  quat_to_rot(qa.x, qa.y, qb.x, qb.y, T);
  T[9] = __ldg(t + 3 * (size_t)c);
  T[10] = __ldg(t + 3 * (size_t)c + 1);
  T[11] = __ldg(t + 3 * (size_t)c + 2);
}

// One step of the transposing butterfly: every lane keeps the half of its N values that its lane bit selects and
// adds the partner's copy of the same half.  Pairing order = warp_sum's (xor 16, 8, 4, 2, 1): bit-identical sums.
template <int N>
__device__ __forceinline__ void fold_step(double* v, bool bit, int offset) {
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const double send = bit ? v[i] : v[i + N / 2];
    const double keep = bit ? v[i + N / 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, offset);
  }
}
// after the five steps lane L holds the total of accumulator l3_acc_of(L) (or a padding slot: -1)
__device__ __forceinline__ int l3_acc_of(int lane) {
  const int li = lane & 3;
  const int idx = li + 3 * ((lane >> 2) & 1) + 6 * ((lane >> 3) & 1) + 12 * ((lane >> 4) & 1);
  return (li < 3 && idx < kCamAcc) ? idx : -1;
}
__host__ __device__ constexpr int l3_lane_of(int k) {
  return ((k % 12) % 6) % 3 + 4 * (((k % 12) % 6) / 3) + 8 * ((k % 12) / 6) + 16 * (k / 12);
}

template <bool WINDOWED>
__global__ void __launch_bounds__(kL3Threads, 1) k_lin3(const L3Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw);          // [0], [1]: stages
  int4* s_ent = reinterpret_cast<int4*>(smem_raw + 32);                                 // per stage: o0, staged count | -1, camera range
  double* s_red = reinterpret_cast<double*>(smem_raw + 64);                             // cost partials of the 8 landmark warps
  unsigned char* stage_base = smem_raw + kL3Head;
  double* s_tiles = reinterpret_cast<double*>(smem_raw + kL3Head + 2 * kL3StageBytes);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x, G = gridDim.x;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  // camera tiles for the kernels that follow this one (nothing in this launch reads them)
  for (int c = b * kL3Threads + tid; c < p.n_cam; c += G * kL3Threads) {
    double T[kCamTile];
    tile_from_qt(p.cam_q, p.cam_t, c, T);
    T[12] = 0.0;
    T[13] = 0.0;
    double2* dst = reinterpret_cast<double2*>(p.Rt + (size_t)kCamTile * c);
#pragma unroll
    for (int k = 0; k < kCamTile / 2; ++k) dst[k] = make_double2(T[2 * k], T[2 * k + 1]);
  }
  __syncthreads();

  if (warp < kL3LmThreads / 32) {
    // =============================== landmark group ===============================
    const int c_begin = (int)((long long)b * p.n_lm_chunks / G), c_end = (int)((long long)(b + 1) * p.n_lm_chunks / G);
    int4 e_next = make_int4(0, 0, 0, -1);      // thread 0: table entry of the chunk to be issued next
    auto issue = [&](int chunk, const int4 e, int st) {
      const int l0 = chunk * kL3Lm, nl = min(kL3Lm, p.n_lm - l0);
      const int cnt = e.y - e.x;
      const bool staged = cnt > 0 && cnt <= kL3StageObs;
      s_ent[st] = make_int4(e.x, staged ? cnt : -1, e.z, e.w);
      unsigned char* base = stage_base + st * kL3StageBytes;
      const unsigned ptr_bytes = ((unsigned)(nl + 1) * 4u + 15u) & ~15u, lm_bytes = (unsigned)nl * 32u;
      const unsigned uv_bytes = staged ? (unsigned)cnt * 16u : 0u, cam_bytes = staged ? (((unsigned)cnt * 4u + 15u) & ~15u) : 0u;
      mbar_expect_tx(&bars[st], ptr_bytes + lm_bytes + uv_bytes + cam_bytes);
      bulk_g2s(base + kL3OffPtr, p.lm_ptr + l0, ptr_bytes, &bars[st]);
      bulk_g2s(base + kL3OffLm, p.lm4 + 4 * (size_t)l0, lm_bytes, &bars[st]);
      if (staged) {
        bulk_g2s(base, p.obs_uv + 2 * (size_t)e.x, uv_bytes, &bars[st]);
        bulk_g2s(base + kL3OffCam, p.obs_cam + e.x, cam_bytes, &bars[st]);
      }
    };
    if (tid == 0 && c_begin < c_end) {
      issue(c_begin, p.ctab[c_begin], 0);
      if (c_begin + 1 < c_end) e_next = p.ctab[c_begin + 1];
    }
    int win_lo = 0, win_n = 0;
    auto stage_table = [&](int lo) {
      const int n = min(kL3MaxCams, p.n_cam - lo);
      for (int i = tid; i < n; i += kL3LmThreads) {
        double T[kCamVals];
        tile_from_qt(p.cam_q, p.cam_t, lo + i, T);
        double2* dst = reinterpret_cast<double2*>(s_tiles + (size_t)kCamTile * i);
#pragma unroll
        for (int k = 0; k < kCamVals / 2; ++k) dst[k] = make_double2(T[2 * k], T[2 * k + 1]);
      }
      win_lo = lo;
      win_n = n;
      l3_lm_bar();
    };
    if (!WINDOWED && c_begin < c_end) stage_table(0);
    else l3_lm_bar();                          // s_ent[0] is visible to the group

    double cost = 0.0;
    unsigned ph0 = 0, ph1 = 0;
    int it = 0;
    for (int chunk = c_begin; chunk < c_end; ++chunk, ++it) {
      const int st = it & 1;
      if (tid == 0 && chunk + 1 < c_end) {
        issue(chunk + 1, e_next, st ^ 1);
        if (chunk + 2 < c_end) e_next = p.ctab[chunk + 2];     // consumed one chunk from now
      }
      const int4 ent = s_ent[st];
      const int o_base = ent.x, staged = ent.y;
      bool tiles_ok = true;
      if (WINDOWED) {
        if (ent.w >= ent.z && (ent.z < win_lo || ent.w >= win_lo + win_n)) {
          if (ent.w - ent.z < kL3MaxCams) stage_table(ent.z);     // (everyone left the old window at the last barrier)
          else tiles_ok = false;
        }
      }
      mbar_wait(&bars[st], (st ? ph1 : ph0) & 1);
      if (st) ++ph1; else ++ph0;
      const unsigned char* base = stage_base + st * kL3StageBytes;
      const double2* s_uv = reinterpret_cast<const double2*>(base);
      const int* s_oc = reinterpret_cast<const int*>(base + kL3OffCam);
      const int* s_ptr = reinterpret_cast<const int*>(base + kL3OffPtr);
      const double* s_lm = reinterpret_cast<const double*>(base + kL3OffLm);

      const int ll = tid >> 1, par = tid & 1, l = chunk * kL3Lm + ll;
      double h0 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0, h5 = 0, g0 = 0, g1 = 0, g2 = 0;
      if (l < p.n_lm) {
        const double2 pxy = *reinterpret_cast<const double2*>(s_lm + 4 * ll);
        const double pz = s_lm[4 * ll + 2];
        const int beg = s_ptr[ll], end = s_ptr[ll + 1];
        auto fetch = [&](int o, int& c, double2& uv) {
          if (staged > 0) {
            c = s_oc[o - o_base];
            uv = s_uv[o - o_base];
          } else {
            c = __ldg(p.obs_cam + o);
            uv = ldg2(p.obs_uv + 2 * (size_t)o);
          }
        };
        auto tile_of = [&](int c, double* T) {
          if (!WINDOWED || tiles_ok) {
            const double2* tile = reinterpret_cast<const double2*>(s_tiles + (size_t)kCamTile * (c - win_lo));
#pragma unroll
            for (int k = 0; k < kCamVals / 2; ++k) {
              const double2 x = tile[k];
              T[2 * k] = x.x;
              T[2 * k + 1] = x.y;
            }
          } else {
            tile_from_qt(p.cam_q, p.cam_t, c, T);
          }
        };
        auto accumulate = [&](const double* T, const Obs& ob) {
          cost = fma(ob.r0, ob.r0, fma(ob.r1, ob.r1, cost));
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
        };
        int o = beg + par;
        for (; o + 2 < end; o += 4) {          // two observations per trip: two independent latency chains
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
      // combine the two half sums (fixed order: even lane + odd lane)
      h0 += __shfl_xor_sync(0xffffffffu, h0, 1); h1 += __shfl_xor_sync(0xffffffffu, h1, 1);
      h2 += __shfl_xor_sync(0xffffffffu, h2, 1); h3 += __shfl_xor_sync(0xffffffffu, h3, 1);
      h4 += __shfl_xor_sync(0xffffffffu, h4, 1); h5 += __shfl_xor_sync(0xffffffffu, h5, 1);
      g0 += __shfl_xor_sync(0xffffffffu, g0, 1); g1 += __shfl_xor_sync(0xffffffffu, g1, 1);
      g2 += __shfl_xor_sync(0xffffffffu, g2, 1);
      if (l < p.n_lm) {
        if (par == 0) {
          double2* H = reinterpret_cast<double2*>(p.Hll + 6 * (size_t)l);
          H[0] = make_double2(h0, h1);
          H[1] = make_double2(h2, h3);
          H[2] = make_double2(h4, h5);
        } else {
          p.gl[3 * (size_t)l] = g0;
          p.gl[3 * (size_t)l + 1] = g1;
          p.gl[3 * (size_t)l + 2] = g2;
        }
      }
      l3_lm_bar();       // the group is done with stage `st`, its descriptor and (if it changes next) the window
    }
    // ---- cost: warp sums -> CTA sum (warp order) -> the last CTA to arrive adds the CTA sums in a fixed order ----
    cost = warp_sum(0.5 * cost);
    if (lane == 0) s_red[warp] = cost;
    l3_lm_bar();
    if (warp == 0) {
      unsigned int last = 0;
      if (lane == 0) {
        double r = s_red[0];
#pragma unroll
        for (int w = 1; w < kL3LmThreads / 32; ++w) r += s_red[w];
        p.partial[b] = r;
        __threadfence();
        last = (atomicInc(p.counter, (unsigned int)G - 1) == (unsigned int)G - 1);   // wraps to 0: re-armed
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        __threadfence();
        const volatile double* pv = p.partial;
        double r = 0.0;
        for (int k = lane; k < G; k += 32) r += pv[k];
        r = warp_sum(r);
        if (lane == 0) *p.out_cost = r;
      }
    }
  }

  // =============================== camera group (and everybody who is done) ===============================
  {
    const unsigned int limit = (unsigned int)p.n_cam_chunks + (unsigned int)G * (kL3Threads / 32) - 1u;
    unsigned int ch = 0;
    if (lane == 0) ch = atomicInc(p.cam_counter, limit);
    ch = __shfl_sync(0xffffffffu, ch, 0);
    while (ch < (unsigned int)p.n_cam_chunks) {
      unsigned int nxt = 0;
      if (lane == 0) nxt = atomicInc(p.cam_counter, limit);      // used after this chunk
      const int c = __ldg(p.chunk_cam + ch);
      const int beg = __ldg(p.chunk_beg + ch), end = __ldg(p.chunk_end + ch);
      double T[kCamVals];
      tile_from_qt(p.cam_q, p.cam_t, c, T);
      double acc[kCamAcc + 1];
#pragma unroll
      for (int k = 0; k <= kCamAcc; ++k) acc[k] = 0.0;
      constexpr int U = 4;                      // observations per lane in flight
      for (int o0 = beg + lane; o0 < end; o0 += 32 * U) {
        int l[U];
        double2 uv[U], pxy[U];
        double pz[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int o = o0 + 32 * k;
          l[k] = o < end ? __ldg(p.cobs_lm + o) : -1;
          uv[k] = o < end ? ldg2(p.cobs_uv + 2 * (size_t)o) : make_double2(0.0, 0.0);
        }
        // the stream two rounds ahead -> L2 (a 128-byte line holds 8 uv pairs / 32 indices)
        {
          const int op = o0 + 64 * U;
          if ((lane & 7) == 0) {
#pragma unroll
            for (int k = 0; k < U; ++k)
              if (op + 32 * k < end) prefetch_l2(p.cobs_uv + 2 * (size_t)(op + 32 * k));
          }
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < U; ++k)
              if (op + 32 * k < end) prefetch_l2(p.cobs_lm + op + 32 * k);
          }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
          if (l[k] >= 0) {
            pxy[k] = ldg2(p.lm4 + 4 * (size_t)l[k]);
            pz[k] = __ldg(p.lm4 + 4 * (size_t)l[k] + 2);
          }
        }
#pragma unroll
        for (int k = 0; k < U; ++k)
          if (l[k] >= 0) cam_accumulate(T, pxy[k].x, pxy[k].y, pz[k], uv[k].x, uv[k].y, reinterpret_cast<double(&)[kCamAcc]>(acc));
      }
      __syncwarp();
      fold_step<24>(acc, (lane >> 4) & 1, 16);
      fold_step<12>(acc, (lane >> 3) & 1, 8);
      fold_step<6>(acc, (lane >> 2) & 1, 4);
      acc[3] = 0.0;
      fold_step<4>(acc, (lane >> 1) & 1, 2);
      fold_step<2>(acc, lane & 1, 1);
      const int mine = l3_acc_of(lane);
      if (mine >= 0) p.chunk_acc[(size_t)ch * kCamAcc + mine] = acc[0];
      // ---- per-camera ticket: the last chunk to arrive finishes the camera ----
      const int ch0 = __ldg(p.cam_chunk_ptr + c), ch1 = __ldg(p.cam_chunk_ptr + c + 1);
      __threadfence();
      __syncwarp();
      unsigned int last = 0;
      if (lane == 0) last = (atomicInc(p.cam_ticket + c, (unsigned int)(ch1 - ch0 - 1)) == (unsigned int)(ch1 - ch0 - 1));   // wraps to 0: re-armed
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        __threadfence();
        double r = 0.0;
        if (mine >= 0) {
          const volatile double* pa = p.chunk_acc;
          for (int k = ch0; k < ch1; ++k) r += pa[(size_t)k * kCamAcc + mine];
        }
        double a[kCamAcc];
#pragma unroll
        for (int k = 0; k < kCamAcc; ++k) a[k] = __shfl_sync(0xffffffffu, r, l3_lane_of(k));
        if (lane == 0) cam_finish(a, T, p.Hcc + (size_t)c * 21, p.gc + (size_t)c * 6);
      }
      ch = __shfl_sync(0xffffffffu, nxt, 0);
    }
  }
}

}  // namespace stba
