// Dense reduced-camera solve for SMALL systems (n <= 320 by default, 512 at most: BASELINE configs[1], 50 cameras ->
// n = 288): ONE launch, ONE CTA.
//
// The 128-block DAG kernel (k_chol_dag2) needs three panel steps at n = 288 and each of them is a chain of one-CTA
// tasks (diagonal block 128 x 128, panel solve, tile update) with a flag round trip in between: 0.21 ms, of which the
// machine is one SM wide almost all the time.  A matrix this small is better served by one CTA that never leaves its
// SM: 32-column panels, the factored panel X kept in shared memory (k-major, so that the DMMA fragments of the
// trailing update are conflict-free 8-byte loads), the trailing matrix updated in place in L2 by 32 x 32 warp tiles
// on the FP64 tensor pipe (mma.sync.m8n8k4.f64), the right-hand side riding along as row n of S (forward substitution
// for free), the backward substitution right-looking by 32-blocks at the end.  Look-ahead: warp 0 factorises the next diagonal block
// while the other warps are still updating the rest of the trailing matrix.
// Fixed summation orders, no atomics on data: bit-reproducible.  Same contract as chol_factor_solve (stba_chol.cuh).
#pragma once

namespace stba {
namespace {

constexpr int CS_B = 32;                       // panel width
constexpr int CS_THREADS = 512;
constexpr int CS_MAXN = 512;                   // largest n served (rows below a panel: n + 1 - 32 <= 481 threads)
constexpr int CS_DEFAULT_N = 320;               // systems up to this order take this path (measured: 0.144 ms against 0.205 ms at n = 288, 0.247 against 0.216 at n = 384; STBA_CHOL_SMALL_N overrides)
constexpr int CS_LDX = CS_MAXN + 4;            // k-major panel: (q * CS_LDX + g) mod 16 distinct for q, g < 4 (8-byte banks)
constexpr int CS_LDD = CS_B + 1;               // diagonal block, row-major, padded
constexpr int CS_LDT = CS_B + 2;               // its factor, column-major, rows 16-byte aligned
constexpr int CS_SMEM = (CS_B * CS_LDX + 2 * CS_B * CS_LDD + 2 * CS_B + CS_MAXN + 1024 + CS_B * CS_LDT + 16) * (int)sizeof(double);      // (the tail: 20 ints)

// Cholesky of the 32 x 32 block in D (row-major, ld CS_LDD, lower part valid, identity-padded) by ONE warp, lane = row,
// the row in registers: the pivot chain of potrf128_prog_dev (stba_chol.cu) — elimination on UNSCALED values, column
// j + 1 exchanged through shared memory one dependent operation after the reciprocal of pivot j, the rest of the
// rank-1 update of step j issued inside step j + 1 where it fills the latency of the exchange and of the reciprocal;
// the columns are scaled by 1 / sqrt(d_j) once at the end.  No shuffles (inside a warp-specialised branch they compile
// to WARPSYNC.COLLECTIVE call sequences).  Output: Lt[c * CS_LDT + r] = L(r, c) (column-major: the panel solve reads a
// column of L with 16-byte broadcasts), invd[j] = 1 / L(j, j).  cb = 32 x 32 doubles of scratch.
// Returns the 1-based index of the first non-positive pivot or 0.
__device__ __forceinline__ int cs_potf2_warp(const double* D, double* Lt, double* invd, double* cb, int lane) {
  double a[CS_B];
#pragma unroll
  for (int c = 0; c < CS_B; ++c) a[c] = (c <= lane) ? D[lane * CS_LDD + c] : 0.0;
  cb[lane] = a[0];
  __syncwarp();
  double tp = 0.0;
#pragma unroll
  for (int j = 0; j < CS_B; ++j) {
    const double* col = cb + j * 32;
    const double d = col[j];
    const double nx = (j + 1 < CS_B) ? col[j + 1] : 0.0;
    if (j > 0) {
      const double* pc = cb + (j - 1) * 32;
#pragma unroll
      for (int k = j + 2; k < CS_B; ++k) a[k] = fma(-tp, pc[k], a[k]);
    }
    const double u = a[j] * nx;
    const double r = fast_rcp(d);
    if (j + 1 < CS_B) {
      a[j + 1] = fma(-u, r, a[j + 1]);
      cb[(j + 1) * 32 + lane] = a[j + 1];
      __syncwarp();
    }
    tp = a[j] * r;
    if (j + 2 < CS_B) a[j + 2] = fma(-tp, col[j + 2], a[j + 2]);
  }
  const double dl = cb[lane * 32 + lane];
  const unsigned badm = __ballot_sync(0xffffffffu, !(dl > 0.0));
  const double rsl = fast_rsqrt(dl);
  invd[lane] = rsl;
  __syncwarp();
#pragma unroll
  for (int c = 0; c < CS_B; ++c) {
    const double v = (c == lane) ? dl * rsl : a[c] * invd[c];
    if (c <= lane) Lt[c * CS_LDT + lane] = v;
  }
  __syncwarp();
  return badm ? __ffs(badm) : 0;
}

#ifdef STBA_CS_TIMING
// phase clocks of warp 0 (every lane ticks, so that the warp stays converged; lane 0's sums are reported)
__device__ long long g_cs_clk[16];
#define CSTICK(slot) do { if (warp == 0) { const long long now_ = clock64(); s_clk[(slot) * 32 + lane] += now_ - t_last; t_last = now_; } } while (0)
constexpr int CS_SMEM_CLK = 16 * 32 * 8;
#else
#define CSTICK(slot) do {} while (0)
constexpr int CS_SMEM_CLK = 0;
#endif

__global__ void __launch_bounds__(CS_THREADS, 1)
k_chol_small(double* __restrict__ S, int ld, int n, double* __restrict__ rhs, int* __restrict__ info) {
  extern __shared__ __align__(16) double cs_sm[];
  double* Xt = cs_sm;                              // [32][CS_LDX]: X^T of the current panel (rows below the diagonal block)
  double* Dbuf = Xt + CS_B * CS_LDX;               // two diagonal blocks (current, next)
  double* invbuf = Dbuf + 2 * CS_B * CS_LDD;       // their inverse diagonals
  double* xs = invbuf + 2 * CS_B;                  // backward substitution: solution so far
  double* cbuf = xs + CS_MAXN;                     // 32 x 32: column exchange of the pivot loop / of the backward solve
  double* Lt = cbuf + 1024;                        // factor of the current diagonal block, column-major
  int* s_bad = reinterpret_cast<int*>(Lt + CS_B * CS_LDT);
  int* s_next = s_bad + 1;
  int* s_tk = s_bad + 4;                           // per-warp ticket broadcast (16)                         // ticket of the trailing-update tiles
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;           // DMMA fragment coordinates

#ifdef STBA_CS_TIMING
  long long* s_clk = reinterpret_cast<long long*>(cs_sm) + CS_SMEM / 8;
  if (warp == 0) for (int k = 0; k < 16; ++k) s_clk[k * 32 + lane] = 0;
  long long t_last = clock64();
#endif
  if (tid == 0) *s_bad = 0;
  for (int c = tid; c < n; c += CS_THREADS) S[(size_t)c * ld + n] = rhs[c];      // the augmented row
  // first diagonal block
  auto load_diag = [&](int k0, double* D, int t0, int nthreads) {
    const int nb = min(CS_B, n - k0);
    for (int idx = t0; idx < CS_B * CS_B; idx += nthreads) {
      const int c = idx >> 5, i = idx & 31;
      double v = (i == c) ? 1.0 : 0.0;
      if (i < nb && c < nb && i >= c) v = S[(size_t)(k0 + c) * ld + k0 + i];
      D[i * CS_LDD + c] = v;
    }
  };
  load_diag(0, Dbuf, tid, CS_THREADS);
  __syncthreads();
  if (warp == 0) {
    const int bad = cs_potf2_warp(Dbuf, Lt, invbuf, cbuf, lane);
    if (bad && lane == 0) { atomicCAS(info, 0, bad); *s_bad = 1; }
  }
  __syncthreads();
  CSTICK(0);

  int cur = 0;
  for (int k0 = 0; k0 < n; k0 += CS_B, cur ^= 1) {
    if (*s_bad) return;                            // (uniform: written before the last barrier)
    const double* invd = invbuf + cur * CS_B;
    const int nb = min(CS_B, n - k0);
    const int r0 = k0 + nb;                        // first row / column of the trailing matrix
    const int mc = n - r0;                         // its order; row mc of the panel (global row n) is the right-hand side
    const int m = mc + 1;

    // ---- panel solve: X = A L^-T, one thread per row (right-looking over the 32 columns), L written back ----
    if (tid < m) {
      const int r = r0 + tid;
      double x[CS_B];
#pragma unroll
      for (int c = 0; c < CS_B; ++c) x[c] = c < nb ? S[(size_t)(k0 + c) * ld + r] : 0.0;
#ifdef STBA_CS_TIMING
      if (x[0] == 1.2345e300) x[1] = 0.0;      // (wait for the loads)
#endif
      CSTICK(10);
#pragma unroll
      for (int c = 0; c < CS_B; ++c) {
        x[c] *= invd[c];
#pragma unroll
        for (int c2 = c + 1; c2 < CS_B; ++c2) x[c2] = fma(-x[c], Lt[c * CS_LDT + c2], x[c2]);
      }
      CSTICK(11);
#pragma unroll
      for (int c = 0; c < CS_B; ++c) {
        if (c < nb) S[(size_t)(k0 + c) * ld + r] = x[c];
        Xt[c * CS_LDX + tid] = x[c];
      }
      CSTICK(12);
    }
    if (tid == 0) *s_next = 1;
    // the factored diagonal block goes back too (lower part)
    for (int idx = tid; idx < CS_B * CS_B; idx += CS_THREADS) {
      const int c = idx >> 5, i = idx & 31;
      if (i < nb && c <= i) S[(size_t)(k0 + c) * ld + k0 + i] = Lt[c * CS_LDT + i];
    }
    __syncthreads();
    CSTICK(1);
    if (mc == 0) break;

    // ---- trailing update A -= X X^T (lower part) + the right-hand-side row ----
    auto update_tile = [&](int ti, int tj, double* Dnext) {
      // 32 x 32 warp tile = 4 x 4 DMMA blocks; C fragment: row g, columns 2q, 2q + 1 of each 8 x 8 block
      const int i0 = ti * 32, j0 = tj * 32;
      double acc[4][4][2];
#pragma unroll
      for (int bi = 0; bi < 4; ++bi)
#pragma unroll
        for (int bj = 0; bj < 4; ++bj)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int i = i0 + 8 * bi + g, j = j0 + 8 * bj + 2 * q + e;
            acc[bi][bj][e] = (i < mc && j <= i) ? S[(size_t)(r0 + j) * ld + r0 + i] : 0.0;
          }
#pragma unroll
      for (int kk = 0; kk < CS_B; kk += 4) {
        double a[4], b[4];
        const double* xr = Xt + (kk + q) * CS_LDX;
#pragma unroll
        for (int bi = 0; bi < 4; ++bi) a[bi] = -xr[i0 + 8 * bi + g];
#pragma unroll
        for (int bj = 0; bj < 4; ++bj) b[bj] = xr[j0 + 8 * bj + g];
#pragma unroll
        for (int bi = 0; bi < 4; ++bi)
#pragma unroll
          for (int bj = 0; bj < 4; ++bj) dmma(acc[bi][bj][0], acc[bi][bj][1], a[bi], b[bj]);
      }
#pragma unroll
      for (int bi = 0; bi < 4; ++bi)
#pragma unroll
        for (int bj = 0; bj < 4; ++bj)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int i = i0 + 8 * bi + g, j = j0 + 8 * bj + 2 * q + e;
            if (Dnext) {      // tile (0, 0) = the next diagonal block: straight into shared memory, identity-padded
              Dnext[i * CS_LDD + j] = (i < mc && j <= i) ? acc[bi][bj][e] : (i == j ? 1.0 : 0.0);
            } else if (i < mc && j <= i) {
              S[(size_t)(r0 + j) * ld + r0 + i] = acc[bi][bj][e];
            }
          }
    };
    const int nt = (mc + 31) >> 5;
    const int n_tiles = nt * (nt + 1) / 2;
    const int nb_next = min(CS_B, n - r0);
    double* Dn = Dbuf + (cur ^ 1) * CS_B * CS_LDD;
    double* invn = invbuf + (cur ^ 1) * CS_B;
    // the right-hand-side row: one thread per column (warps 1 .. 15: warp 0 is on the critical path)
    for (int j = warp ? tid - 32 : mc; j < mc; j += CS_THREADS - 32) {
      double y = S[(size_t)(r0 + j) * ld + n];
#pragma unroll
      for (int kk = 0; kk < CS_B; ++kk) y = fma(-Xt[kk * CS_LDX + mc], Xt[kk * CS_LDX + j], y);
      S[(size_t)(r0 + j) * ld + n] = y;
    }
    if (warp == 0) {
      // look-ahead: tile (0, 0) holds the next diagonal block — update it, factorise it, then join the others
      CSTICK(2);
      update_tile(0, 0, Dn);
      __syncwarp();
      CSTICK(4);
      const int bad = cs_potf2_warp(Dn, Lt, invn, cbuf, lane);
      CSTICK(5);
      if (bad && lane == 0 && bad <= nb_next) { atomicCAS(info, 0, r0 + bad); *s_bad = 1; }
    }
    // tiles 1 .. n_tiles - 1 (tile t of row ti: index ti (ti + 1) / 2 + tj) from a shared ticket; warp 0 joins late.
    // (Keeping the warps that share warp 0's scheduler out of the tile loop until the pivot chain is done halves the
    // chain — 12 k -> 6.6 k cycles per block, their DMMAs queue in front of its dependent operations — but the tiles
    // they do not take arrive later by as much: 0.146 -> 0.152 ms at n = 288, not kept.)
    for (;;) {
      if (lane == 0) s_tk[warp] = atomicAdd(s_next, 1);
      __syncwarp();
      const int t = s_tk[warp];
      __syncwarp();
      if (t >= n_tiles) break;
      int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
      while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
      while (ti * (ti + 1) / 2 > t) --ti;
      update_tile(ti, t - ti * (ti + 1) / 2, nullptr);
    }
    CSTICK(6);
    __syncthreads();
    CSTICK(7);
  }

  // ---- backward substitution L^T x = y (y = row n of S), right-looking by 32-blocks from the end: xs holds y until a
  //      block is solved, then x.  Per block: one warp solves the 32 x 32 triangle, then every thread j < j0 takes the
  //      block's 32 rows out of its y_j (its own 256 contiguous bytes of column j: all loads in flight together) while
  //      the next diagonal block is fetched — one L2 round trip per block hop. ----
  for (int c = tid; c < n; c += CS_THREADS) xs[c] = S[(size_t)c * ld + n];
  auto load_lb = [&](int j0, double* Lb) {
    const int nb = min(CS_B, n - j0);
    for (int idx = tid; idx < CS_B * CS_B; idx += CS_THREADS) {
      const int c = idx >> 5, i = idx & 31;
      Lb[i * CS_LDD + c] = (i < nb && c <= i) ? S[(size_t)(j0 + c) * ld + j0 + i] : (i == c ? 1.0 : 0.0);
    }
  };
  const int j_last = ((n - 1) / CS_B) * CS_B;
  load_lb(j_last, Dbuf);
  __syncthreads();
  int pb = 0;
  for (int j0 = j_last; j0 >= 0; j0 -= CS_B, pb ^= 1) {
    const int nb = min(CS_B, n - j0);
    const double* Lb = Dbuf + pb * CS_B * CS_LDD;
    if (warp == 0) {
      // lane j owns column j of the block: x_i known -> y_j -= L[i][j] x_i for j < i
      double y = lane < nb ? xs[j0 + lane] : 0.0;
      const double dinv = 1.0 / Lb[lane * CS_LDD + lane];
      double xv = 0.0;
#pragma unroll 4
      for (int i = CS_B - 1; i >= 0; --i) {
        const double l = Lb[i * CS_LDD + lane];
        double* slot = cbuf + (i & 1) * 32;
        if (lane == i) { xv = y * dinv; *slot = xv; }      // lane i's y is final at step i
        __syncwarp();
        const double xi = *slot;
        if (lane < i) y = fma(-l, xi, y);
      }
      if (lane < nb) xs[j0 + lane] = xv;
    }
    __syncthreads();
    CSTICK(8);
    if (j0 == 0) break;
    for (int j = tid; j < j0; j += CS_THREADS) {
      const double* col = S + (size_t)j * ld + j0;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      if (nb == CS_B) {
#pragma unroll
        for (int i = 0; i < CS_B; i += 4) {
          s0 = fma(col[i], xs[j0 + i], s0);
          s1 = fma(col[i + 1], xs[j0 + i + 1], s1);
          s2 = fma(col[i + 2], xs[j0 + i + 2], s2);
          s3 = fma(col[i + 3], xs[j0 + i + 3], s3);
        }
      } else {
        for (int i = 0; i < nb; ++i) s0 = fma(col[i], xs[j0 + i], s0);
      }
      xs[j] -= (s0 + s1) + (s2 + s3);
    }
    load_lb(j0 - CS_B, Dbuf + (pb ^ 1) * CS_B * CS_LDD);
    __syncthreads();
    CSTICK(9);
  }
  for (int c = tid; c < n; c += CS_THREADS) rhs[c] = xs[c];
#ifdef STBA_CS_TIMING
  if (tid < 16) g_cs_clk[tid] = s_clk[tid * 32];
#endif
}

}  // namespace
}  // namespace stba
