// Linearisation, third generation: ONE launch for landmark-major pass + camera-major pass (the camera tiles are part
// of the state: k_cam_prep builds them for a candidate point, the engine swaps them in when the step is accepted)
// (the contract kernel of BASELINE.json: per-observation reprojection residual + exact Jacobian + J^T J / J^T r
// block accumulation, J never stored; algorithmic bytes 24 N_obs + 96 N_lm + 272 N_cam).
//
// Rounds 1-2 ran three launches (k_cam_prep, k_lin_lm2, k_lin_cam2) back to back: each pass alone left the FP64
// pipe at 20-27 % (profiles/r2_lin_full.md: 16 warps per SM, long/short-scoreboard stalls), and they could not
// overlap because lin_lm2 owned the SM (222 KB of shared memory, 512 x 110 registers).  Here the two passes run
// SIDE BY SIDE on every SM, as warp-specialised halves of one persistent 512-thread CTA:
//   * warps 0-7, the landmark group: contiguous ranges of 128-landmark chunks per CTA.  EVERYTHING a chunk reads —
//     its observation stream (uv + camera index), its lm_ptr slice and its landmark points — is staged by TMA bulk
//     copies (cp.async.bulk + mbarrier, one chunk ahead, asked into L2 four chunks ahead by cp.async.bulk.prefetch.L2,
//     issued by thread 0 from a shared-memory ring of chunk-table entries), and so is the camera-tile table [R | t];
//     the inner loop touches no global memory.  With more than 1024 cameras the table holds a WINDOW of 1024
//     consecutive cameras, re-staged when a chunk's camera range (chunk table, built once per problem) leaves it; a
//     chunk that spans more than a window gathers its tiles from global memory.
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
constexpr int kL3Stages = 2;                                 // (three stages left 28 KB of L1 to the camera group's gathers: 41 -> 51 us at C)
constexpr int kL3Ahead = 4;                                  // chunks between the L2 prefetch of a chunk and its use
constexpr int kL3Head = 512;                                 // mbarriers + the stage descriptors + cost partials + chunk-table ring
constexpr int kL3SmemBytes = kL3Head + kL3Stages * kL3StageBytes + kL3MaxCams * kCamTile * 8;

#ifdef STBA_L3_TIMING
__device__ long long g_l3_clk[160 * 64];
__device__ __forceinline__ long long l3_now() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define L3TICK(cond, slot) do { if (cond) g_l3_clk[blockIdx.x * 64 + (slot)] = l3_now(); } while (0)
#else
#define L3TICK(cond, slot) do {} while (0)
#endif

struct L3Params {
  int n_lm, n_cam, n_lm_chunks, n_cam_chunks;
  const int* lm_ptr; const int* obs_cam; const double* obs_uv; const double* lm4;
  const int4* ctab;                 // per landmark chunk: first staged observation (aligned down to 4), end, camera range
  const double* Rt;                 // camera tiles [R | t] of the current state (kept valid by the engine: built once, swapped on accept)
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
// TMA prefetch of a contiguous range into L2 (no shared memory, no completion): one instruction per range
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, unsigned bytes) {
  const unsigned long long a = reinterpret_cast<unsigned long long>(p);
  const unsigned long long a0 = a & ~15ull;
  const unsigned n = (bytes + (unsigned)(a - a0) + 15u) & ~15u;
  if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(n) : "memory");
}

// one landmark point (padded to 32 bytes: one sector) with ONE 256-bit load (sm_100: LDG.E.256) — a warp's 32 random
// gathers cost 32 L1 wavefronts instead of the 64 of a 16-byte + an 8-byte load
__device__ __forceinline__ void ldg_point(const double* p, double& x, double& y, double& z) {
  [[maybe_unused]] double w;
  // (no L1 allocation: the points are used once per warp and would only evict each other; measured 41.0 -> 39.3 us at C)
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(p));
  (void)w;
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
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw);          // one per stage
  int4* s_ent = reinterpret_cast<int4*>(smem_raw + 32);                                 // per stage: o0, staged count | -1, camera range
  double* s_red = reinterpret_cast<double*>(smem_raw + 128);                            // cost partials of the 8 landmark warps
  unsigned char* stage_base = smem_raw + kL3Head;
  double* s_tiles = reinterpret_cast<double*>(smem_raw + kL3Head + kL3Stages * kL3StageBytes);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x, G = gridDim.x;
  const int c_begin = (int)((long long)b * p.n_lm_chunks / G), c_end = (int)((long long)(b + 1) * p.n_lm_chunks / G);

  // producer (thread 0): stage the observations, the lm_ptr slice and the points of `chunk`
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
  // everything `chunk` will read -> L2, kL3Ahead chunks before its bulk copy into shared memory is issued: that copy
  // then takes an L2 round trip instead of a cold DRAM access (2-3 us against 1.5 us of work per chunk)
  auto prefetch = [&](int chunk, const int4 e) {
    const int l0 = chunk * kL3Lm, nl = min(kL3Lm, p.n_lm - l0);
    bulk_prefetch_l2(p.lm_ptr + l0, (unsigned)(nl + 1) * 4u);
    bulk_prefetch_l2(p.lm4 + 4 * (size_t)l0, (unsigned)nl * 32u);
    const int cnt = e.y - e.x;
    if (cnt > 0) {
      bulk_prefetch_l2(p.obs_uv + 2 * (size_t)e.x, (unsigned)cnt * 16u);
      bulk_prefetch_l2(p.obs_cam + e.x, (unsigned)cnt * 4u);
    }
  };
  // camera tiles lo .. lo + n - 1 -> the shared-memory table (bulk copies of <= 32 KB, one barrier)
  auto issue_table = [&](int lo) {
    const unsigned bytes = (unsigned)min(kL3MaxCams, p.n_cam - lo) * (kCamTile * 8u);
    mbar_expect_tx(&bars[kL3Stages], bytes);
    for (unsigned off = 0; off < bytes; off += 32768u)
      bulk_g2s(reinterpret_cast<unsigned char*>(s_tiles) + off, reinterpret_cast<const unsigned char*>(p.Rt + (size_t)kCamTile * lo) + off,
               min(32768u, bytes - off), &bars[kL3Stages]);
  };
  int4* s_tab = reinterpret_cast<int4*>(smem_raw + 256);      // ring of chunk-table entries (16)
  int4 e_next = make_int4(0, 0, 0, -1);        // thread 0: table entry on its way into the ring

  L3TICK(tid == 0, 0);
#ifdef STBA_L3_TIMING
  if (tid == 256) for (int k = 56; k < 62; ++k) g_l3_clk[blockIdx.x * 64 + k] = 0;
#endif
  if (tid == 32) {
    // The camera group gathers landmark points at random: 128 gathers per warp and round, and while the points are
    // cold in L2 every round waits for a DRAM miss (measured: 2.7-3.6 k cycles per round on the gathers alone).
    // Each CTA asks for its slice of the array up front (3.2 MB at C in total: 0.5 us of DRAM time).
    const size_t total = (size_t)p.n_lm * 32, per = ((total + G - 1) / G + 127) & ~(size_t)127;
    const size_t lo = min(total, (size_t)b * per), hi = min(total, lo + per);
    for (size_t off = lo; off < hi; off += 32768)
      bulk_prefetch_l2(reinterpret_cast<const char*>(p.lm4) + off, (unsigned)min((size_t)32768, hi - off));
  }
  if (tid == 64) {
    // the chunk descriptors of the camera group are read in dependent chains (ticket -> descriptor -> stream): keep them
    // one L2 round trip away instead of one DRAM access
    auto warm = [&](const void* base, size_t total) {
      const size_t per = ((total + G - 1) / G + 127) & ~(size_t)127;
      const size_t lo = min(total, (size_t)b * per), hi = min(total, lo + per);
      if (hi > lo) bulk_prefetch_l2(reinterpret_cast<const char*>(base) + lo, (unsigned)(hi - lo));
    };
    warm(p.chunk_cam, (size_t)p.n_cam_chunks * 4);
    warm(p.chunk_beg, (size_t)p.n_cam_chunks * 4);
    warm(p.chunk_end, (size_t)p.n_cam_chunks * 4);
    warm(p.cam_chunk_ptr, ((size_t)p.n_cam + 1) * 4);
  }
  if (tid < 32) {
    // warp 0: the first table entries (one cold load for the warp), barriers, first bulk copy, first prefetches
    int4 e_mine = make_int4(0, 0, 0, -1);
    if (tid < 8 && c_begin + tid < c_end) e_mine = p.ctab[c_begin + tid];      // (in flight while thread 0 sets up)
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s <= kL3Stages; ++s) mbar_init(&bars[s], 1);
      mbar_fence_init();
      if (!WINDOWED && c_begin < c_end) issue_table(0);
    }
    if (tid < 8) s_tab[tid] = e_mine;
    __syncwarp();
    if (tid == 0 && c_begin < c_end) {
      issue(c_begin, s_tab[0], 0);
      for (int a = 1; a < kL3Ahead && c_begin + a < c_end; ++a) prefetch(c_begin + a, s_tab[a]);
      if (c_begin + 8 < c_end) e_next = p.ctab[c_begin + 8];
    }
  }
  int win_lo = 0, win_n = 0;
  unsigned tphase = 0;      // completed uses of the table barrier (uniform across the landmark group)

  if (warp < kL3LmThreads / 32) {
    // =============================== landmark group ===============================
    l3_lm_bar();            // mbarriers and first descriptors are visible (the camera group never touches them)
    L3TICK(tid == 0, 1);
    auto stage_table = [&](int lo) {
      if (tid == 0) issue_table(lo);
      win_lo = lo;
      win_n = min(kL3MaxCams, p.n_cam - lo);
      mbar_wait(&bars[kL3Stages], tphase & 1u);
      ++tphase;
    };
    if (!WINDOWED && c_begin < c_end) {       // issued by thread 0 before anything else
      win_n = p.n_cam;
      mbar_wait(&bars[kL3Stages], 0);
      ++tphase;
    }
    L3TICK(tid == 0, 2);

    double cost = 0.0;
    int it = 0;
    for (int chunk = c_begin; chunk < c_end; ++chunk, ++it) {
      const int st = it % kL3Stages;
      if (tid == 0) {
        // ring slot (it + 8) % 16 <- the entry requested one chunk ago; request the next one
        if (chunk + 8 < c_end) s_tab[(it + 8) & 15] = e_next;
        if (chunk + 9 < c_end) e_next = p.ctab[chunk + 9];
        if (chunk + 1 < c_end) issue(chunk + 1, s_tab[(it + 1) & 15], (it + 1) % kL3Stages);     // (that stage was consumed before the last barrier)
        if (chunk + kL3Ahead < c_end) prefetch(chunk + kL3Ahead, s_tab[(it + kL3Ahead) & 15]);
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
      mbar_wait(&bars[st], (unsigned)(it / kL3Stages) & 1u);
      L3TICK(tid == 0 && it < 8, 8 + 2 * it);
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
#pragma unroll
            for (int k = 0; k < kCamVals; k += 2) {
              const double2 x = ldg2(p.Rt + (size_t)kCamTile * c + k);
              T[k] = x.x;
              T[k + 1] = x.y;
            }
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
      L3TICK(tid == 0 && it < 8, 9 + 2 * it);
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
    L3TICK(tid == 0, 3);
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
  L3TICK(tid == 0, 4);
  L3TICK(tid == 256, 32);
  {
#ifdef STBA_L3_TIMING
    int n_done = 0;
#endif
    // chunk ids: the first G * 8 are handed out statically (camera warp w of CTA b starts on chunk w * G + b, without a
    // ticket round trip), the rest through the ticket counter; every warp ends on exactly one failing pull, so the
    // counter wraps to 0 with the last one
    const unsigned int n_static = (unsigned int)G * (kL3Threads / 32 - kL3LmThreads / 32);
    const unsigned int n_dyn = (unsigned int)p.n_cam_chunks > n_static ? (unsigned int)p.n_cam_chunks - n_static : 0u;
    const unsigned int limit = n_dyn + (unsigned int)G * (kL3Threads / 32) - 1u;
    // the whole stream of a chunk (16 + 4 bytes per observation, contiguous) -> L2
    auto prefetch_chunk = [&](int c, int beg, int end) {
      if (lane == 0) {
        bulk_prefetch_l2(p.cobs_uv + 2 * (size_t)beg, (unsigned)(end - beg) * 16u);
        bulk_prefetch_l2(p.cobs_lm + beg, (unsigned)(end - beg) * 4u);
        prefetch_l2(p.Rt + (size_t)kCamTile * c);
      }
    };
    unsigned int ch = 0;
    if (warp >= kL3LmThreads / 32) {
      ch = (unsigned int)(warp - kL3LmThreads / 32) * (unsigned int)G + (unsigned int)b;
      if (ch >= (unsigned int)p.n_cam_chunks) ch = 0xffffffffu;      // fewer chunks than camera warps: go and fail a pull
    } else {
      ch = 0xffffffffu;
    }
    if (ch == 0xffffffffu) {
      if (lane == 0) ch = atomicInc(p.cam_counter, limit) + n_static;
      ch = __shfl_sync(0xffffffffu, ch, 0);
    }
    int c = 0, beg = 0, end = 0;
    if (ch < (unsigned int)p.n_cam_chunks) {
      c = __ldg(p.chunk_cam + ch);
      beg = __ldg(p.chunk_beg + ch);
      end = __ldg(p.chunk_end + ch);
      prefetch_chunk(c, beg, end);
    }
    while (ch < (unsigned int)p.n_cam_chunks) {
      unsigned int nxt = 0;
      if (lane == 0) nxt = atomicInc(p.cam_counter, limit) + n_static;      // the chunk after this one
      int cn = 0, begn = 0, endn = 0, round = 0;
      double T[kCamVals];
#pragma unroll
      for (int k = 0; k < kCamVals; k += 2) {
        const double2 x = ldg2(p.Rt + (size_t)kCamTile * c + k);
        T[k] = x.x;
        T[k + 1] = x.y;
      }
      double acc[kCamAcc + 1];
#pragma unroll
      for (int k = 0; k <= kCamAcc; ++k) acc[k] = 0.0;
      constexpr int U = 4;                      // observations per lane in flight
#ifdef STBA_L3_TIMING
      long long pc0 = 0, pc1 = 0, pc2 = 0, pc3 = 0, ps0 = 0, ps1 = 0, ps2 = 0;
      const long long chunk_t0 = clock64();
#define L3PH(var) var = clock64()
#define L3TOUCHI(x) asm volatile("" ::"r"(x))
#define L3TOUCHD(x) asm volatile("" ::"d"(x))
#else
#define L3PH(var) do {} while (0)
#define L3TOUCHI(x) do {} while (0)
#define L3TOUCHD(x) do {} while (0)
#endif
      for (int ob = beg; ob < end; ob += 32 * U, ++round) {      // (warp-uniform trip count: shuffles inside)
        const int o0 = ob + lane;
        L3PH(pc0);
        int l[U];
        double2 uv[U], pxy[U];
        double pz[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int o = o0 + 32 * k;
          l[k] = o < end ? __ldg(p.cobs_lm + o) : -1;
          uv[k] = o < end ? ldg2(p.cobs_uv + 2 * (size_t)o) : make_double2(0.0, 0.0);
        }
#ifdef STBA_L3_TIMING
#pragma unroll
        for (int k = 0; k < U; ++k) { L3TOUCHI(l[k]); L3TOUCHD(uv[k].x); L3TOUCHD(uv[k].y); }
#endif
        L3PH(pc1);
#pragma unroll
        for (int k = 0; k < U; ++k) {
          if (l[k] >= 0) ldg_point(p.lm4 + 4 * (size_t)l[k], pxy[k].x, pxy[k].y, pz[k]);
        }
        // the next chunk: its descriptor now (the ticket has arrived behind the stream loads), its stream -> L2 after
        // this round: the next chunk starts on L2 hits instead of three dependent DRAM misses
        if (round == 0) {
          nxt = __shfl_sync(0xffffffffu, nxt, 0);
          if (nxt < (unsigned int)p.n_cam_chunks) {
            cn = __ldg(p.chunk_cam + nxt);
            begn = __ldg(p.chunk_beg + nxt);
            endn = __ldg(p.chunk_end + nxt);
          }
        }
#ifdef STBA_L3_TIMING
#pragma unroll
        for (int k = 0; k < U; ++k) { L3TOUCHD(pxy[k].x); L3TOUCHD(pxy[k].y); L3TOUCHD(pz[k]); }
#endif
        L3PH(pc2);
#pragma unroll
        for (int k = 0; k < U; ++k)
          if (l[k] >= 0) cam_accumulate(T, pxy[k].x, pxy[k].y, pz[k], uv[k].x, uv[k].y, reinterpret_cast<double(&)[kCamAcc]>(acc));
#ifdef STBA_L3_TIMING
#pragma unroll
        for (int k = 0; k < kCamAcc; ++k) L3TOUCHD(acc[k]);
        L3PH(pc3);
        ps0 += pc1 - pc0; ps1 += pc2 - pc1; ps2 += pc3 - pc2;
#endif
        if (round == 0 && nxt < (unsigned int)p.n_cam_chunks) prefetch_chunk(cn, begn, endn);
      }
#ifdef STBA_L3_TIMING
      const long long chunk_t1 = clock64();
#endif
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
#ifdef STBA_L3_TIMING
      if (tid == 256) {
        long long* o = g_l3_clk + blockIdx.x * 64;
        o[56] += ps0; o[57] += ps1; o[58] += ps2; o[59] += round; o[60] += chunk_t1 - chunk_t0; o[61] += clock64() - chunk_t1;
      }
#endif
      ch = nxt;
      c = cn;
      beg = begn;
      end = endn;
#ifdef STBA_L3_TIMING
      ++n_done;
#endif
      L3TICK(tid == 256 && n_done < 24, 32 + n_done);
      L3TICK(tid == 0 && n_done < 6, 24 + n_done);
    }
    L3TICK(tid == 0, 5);
    L3TICK(tid == 256, 6);
#ifdef STBA_L3_TIMING
    if (tid == 0) g_l3_clk[blockIdx.x * 64 + 7] = n_done;
    if (tid == 256) g_l3_clk[blockIdx.x * 64 + 63] = n_done;
#endif
  }
}

}  // namespace stba
