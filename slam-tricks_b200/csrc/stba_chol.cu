// Own dense reduced-camera solve: blocked right-looking Cholesky (fp64, lower, column-major, in
// place) + block triangular solves.  The trailing-matrix update and the panel solve run on the
// FP64 tensor pipe (mma.sync.m8n8k4.f64 = DMMA); see DESIGN.md §3.5.
//
//   for each 128-wide block column k:
//     k_potrf128     one CTA: Cholesky of the diagonal block in shared memory + explicit inverse
//                    of the triangular factor (Linv_kk, kept in a side buffer for the solves)
//     k_gemm_nt<TRSM> panel:   L_ik = A_ik Linv_kk^T                (DMMA, one CTA per row tile)
//     k_gemm_nt<SYRK> trailing: A_ij -= L_ik L_jk^T for i >= j > k   (DMMA, one CTA per tile)
//   The right-hand side rides along as row n of S (leading dimension >= n + 1): the panel solves
//   and trailing updates turn it into y = L^-1 rhs for free, so only the backward substitution
//   (by 128-blocks, with the stored Linv_kk^T, matrix-vector only) remains.
//
// The block column k+1 is updated first and its panel factorised on a second stream while the
// rest of the trailing update of step k is still running (look-ahead); the whole schedule is
// captured once per (S, n) into a CUDA graph.
#include "stba_chol.cuh"

#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <vector>

namespace stba {

namespace {

constexpr int NB = 128;          // block size
constexpr int KC = 16;           // k-chunk per pipeline stage
constexpr int STAGES = 4;
constexpr int LDS = NB + 4;      // smem leading dimension (doubles): (q*LDS + g) mod 16 distinct
constexpr int GEMM_THREADS = 256;
// per-TM shared memory: A stages are only TM + 4 rows wide, so two 64-row CTAs fit on one SM
constexpr int gemm_smem(int tm) { return STAGES * KC * ((tm + 4) + LDS) * (int)sizeof(double); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 1/sqrt(d): hardware seed (MUFU.RSQ64H, ~2^-22) + two Newton steps, no special-case branches —
// the pivot of the Cholesky recurrence sits on this chain 128 times per diagonal block.
__device__ __forceinline__ double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-d * y, y, 1.0);       // seed error <= 2^-22 -> 2^-43 -> 2^-85
  y = fma(0.5 * y, e, y);
  e = fma(-d * y, y, 1.0);
  return fma(0.5 * y, e, y);
}

enum { MODE_SYRK = 0, MODE_TRSM = 1 };

// One TM x 128 output tile (TM = 128 for the bulk trailing update, 64 for the latency-critical
// panel solve and next-panel strip so that they spread over all SMs):
//   acc = A[i0.., k0..k0+K) * B[j0.., ...)^T  (both "row x k" panels stored column-major), then
//   MODE_SYRK: C[i0.., j0..] -= acc            A = B = the factored panel of S, ldb = lda
//   MODE_TRSM: A[i0.., k0..k0+K) = acc         B = Linv (K x K, ld NB), j0 = 0
// tiles[] lists (i-tile in units of TM rows, j-tile in units of 128 cols); rows >= n_rows and
// cols >= n_cols are masked.
template <int MODE, int TM>
__global__ void __launch_bounds__(GEMM_THREADS, TM == 64 ? 2 : 1)
k_gemm_nt(double* __restrict__ S, int lda, int n_rows, int n_cols, int k0, int K, const double* __restrict__ Bmat, int ldb,
          const int2* __restrict__ tiles) {
  constexpr int WM = TM / 2;                      // warp tile rows (2 warps along M, 4 along N)
  constexpr int MT = WM / 8;                      // 8-row mma tiles per warp
  extern __shared__ __align__(16) double smem[];
  constexpr int LDA = TM + 4;                     // (q*LDA + g) mod 16 distinct for TM = 64, 128
  double* As = smem;                              // [STAGES][KC][LDA]
  double* Bs = smem + STAGES * KC * LDA;          // [STAGES][KC][LDS]
  const int2 tile = tiles[blockIdx.x];
  const int i0 = tile.x * TM, j0 = (MODE == MODE_TRSM) ? 0 : tile.y * NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  const double* Ag = S + (size_t)k0 * lda;        // panel columns k0..k0+K
  const double* Bg = (MODE == MODE_TRSM) ? Bmat : S + (size_t)k0 * lda;
  const int b_rows = (MODE == MODE_TRSM) ? K : n_rows;

  // SYRK: the accumulators START as the C tile (loads issued here overlap the pipeline fill) and
  // the A fragments are negated, so the epilogue is store-only.  A read-modify-write epilogue
  // serialises on load->store aliasing and cost half the kernel (profiles/r1_dense_notes.md).
  double acc[MT][4][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int r = i0 + wm * WM + mt * 8 + g;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = j0 + wn * 32 + nt * 8 + 2 * q + e;
        acc[mt][nt][e] = (MODE == MODE_SYRK && r < n_rows && c < n_cols) ? __ldcg(S + (size_t)c * lda + r) : 0.0;
      }
  }

  const int n_chunks = (K + KC - 1) / KC;
  auto load_stage = [&](int chunk, int stage) {
    double* as = As + stage * KC * LDA;
    double* bs = Bs + stage * KC * LDS;
#pragma unroll
    for (int p = 0; p < KC * (TM / 2) / GEMM_THREADS; ++p) {       // A: KC columns x TM rows
      const int piece = tid + p * GEMM_THREADS;
      const int kk = piece / (TM / 2), r2 = (piece % (TM / 2)) * 2;
      const int k = chunk * KC + kk;
      const bool kin = k < K;
      const int ra = i0 + r2;
      cp_async16(as + kk * LDA + r2, Ag + (size_t)(kin ? k : 0) * lda + (ra < n_rows ? ra : 0), kin && ra < n_rows);
    }
#pragma unroll
    for (int p = 0; p < KC * 64 / GEMM_THREADS; ++p) {             // B: KC columns x 128 rows
      const int piece = tid + p * GEMM_THREADS;
      const int kk = piece >> 6, r2 = (piece & 63) * 2;
      const int k = chunk * KC + kk;
      const bool kin = k < K;
      const int rb = j0 + r2;
      cp_async16(bs + kk * LDS + r2, Bg + (size_t)(kin ? k : 0) * ldb + (rb < b_rows ? rb : 0), kin && rb < b_rows);
    }
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < n_chunks) load_stage(s, s);
    cp_async_commit();
  }
  for (int c = 0; c < n_chunks; ++c) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (c + STAGES - 1 < n_chunks) load_stage(c + STAGES - 1, (c + STAGES - 1) % STAGES);
    cp_async_commit();
    const double* as = As + (c % STAGES) * KC * LDA + wm * WM + g;
    const double* bs = Bs + (c % STAGES) * KC * LDS + wn * 32 + g;
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
      double a[MT], b[4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a[mt] = (MODE == MODE_SYRK) ? -as[(kk + q) * LDA + mt * 8] : as[(kk + q) * LDA + mt * 8];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) b[nt] = bs[(kk + q) * LDS + nt * 8];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();   // TRSM overwrites the panel it read: every warp must be done reading
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int r = i0 + wm * WM + mt * 8 + g;
    if (r >= n_rows) continue;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cc = wn * 32 + nt * 8 + 2 * q + e;
        if (MODE == MODE_SYRK) {
          const int c = j0 + cc;
          if (c < n_cols) S[(size_t)c * lda + r] = acc[mt][nt][e];
        } else {
          if (cc < K) S[(size_t)(k0 + cc) * lda + r] = acc[mt][nt][e];
        }
      }
    }
  }
}

// ---- panel solve with the diagonal inverse blocks only ------------------------------------------
// X = A L_kk^-T for 64 rows of the panel, by forward substitution over the four 32-column blocks:
//   X_j = (A_j - sum_{p<j} X_p L_jp^T) Inv_jj^T,   j = 0..3
// Each warp owns 8 rows and carries them through all four steps on its own (its X rows live in a
// warp-private shared-memory strip, the C fragments are turned into A fragments through it), so
// there is no CTA-wide synchronisation after the operands have landed.  320 DMMAs per warp instead
// of the 512 of a product with the full 128 x 128 inverse — and the inverse's off-diagonal blocks
// are no longer on the critical path.
#ifdef STBA_CHOL_TIMING
__device__ long long g_trsm_clk[16];
#define TTICK(i) do { if (threadIdx.x == 0 && blockIdx.x == (gridDim.x > 100 ? 1 : 0)) g_trsm_clk[i] = clock64(); } while (0)
#else
#define TTICK(i) do {} while (0)
#endif
constexpr int TS_THREADS = 256;
constexpr int TS_ROWS = 64;
constexpr int TS_SMEM = (NB * LDS + 8 * NB * 8) * (int)sizeof(double);

// Device-function form: the first 8 warps of the CTA own the 64 rows i0 .. i0 + 63; every warp of the CTA
// (n_warps of them) helps with the operand loads and must take part in the CTA-wide barriers.  All global
// reads bypass L1 (cp.async.cg / ld.cg): inside the persistent DAG kernel the operands were written by
// other SMs during the same launch.
__device__ __forceinline__ void trsm64_dev(double* __restrict__ S, int ld, int n_rows, int k0, int nb, int i0,
                                           const double* __restrict__ Linv, double* smem, int n_warps) {
  double* Ls = smem;                         // Ls[c * LDS + r]: L(r, c) below the diagonal blocks, Inv(r, c) inside them
  double* Xs = smem + NB * LDS;              // Xs[warp][col][8 rows]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  TTICK(0);
  // Operands arrive in four cp.async groups, one per 32-column block b: the columns 32b..32b+31 of
  // the factor block (Inv_bb inside the diagonal block, L below it) and of the CTA's 64 rows.  Step
  // j only needs groups <= j, so the later groups land while the first steps compute.  One warp per
  // column, lanes along the rows: the index arithmetic is warp-uniform and cheap.
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    for (int c = 32 * b + warp; c < 32 * b + 32; c += n_warps) {
      double* dl = Ls + c * LDS;
      const bool col_ok = c < nb;
      // rows 32b .. 127 of column c, as pairs (r2, r2 + 1)
      for (int r2 = 32 * b + 2 * lane; r2 < NB; r2 += 64) {
        const bool in_diag = r2 < 32 * b + 32;
        const double* src = in_diag ? Linv + (size_t)c * NB + r2 : S + (size_t)(k0 + c) * ld + k0 + r2;
        if (col_ok && r2 >= c && r2 + 1 < nb) {
          cp_async16(dl + r2, src, true);
        } else {                       // diagonal / padding pairs (8-byte cp.async is .ca only: it could read a stale L1 line)
          dl[r2] = (col_ok && r2 >= c && r2 < nb) ? __ldcg(src) : 0.0;
          dl[r2 + 1] = (col_ok && r2 + 1 >= c && r2 + 1 < nb) ? __ldcg(src + 1) : 0.0;
        }
      }
      // the CTA's 64 rows of panel column c
      {
        const int r2 = 2 * lane, row = i0 + r2;
        double* dst = Xs + ((r2 >> 3) * NB + c) * 8 + (r2 & 7);
        if (col_ok && row + 1 < n_rows) {
          cp_async16(dst, S + (size_t)(k0 + c) * ld + row, true);
        } else {
          dst[0] = (col_ok && row < n_rows) ? __ldcg(S + (size_t)(k0 + c) * ld + row) : 0.0;
          dst[1] = 0.0;
        }
      }
    }
    cp_async_commit();
  }
  double* Xw = Xs + (warp & 7) * NB * 8;
  const int row = i0 + warp * 8 + g;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    TTICK(1 + 2 * j);
    if (j == 0) cp_async_wait<3>(); else if (j == 1) cp_async_wait<2>(); else if (j == 2) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    TTICK(2 + 2 * j);
    if (warp >= 8) continue;
    double acc[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) acc[nt][e] = Xw[(32 * j + 8 * nt + 2 * q + e) * 8 + g];
    for (int kk = 0; kk < 32 * j; kk += 4) {
      const double a = -Xw[(kk + q) * 8 + g];
      const double* bs = Ls + (kk + q) * LDS + 32 * j + g;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dmma(acc[nt][0], acc[nt][1], a, bs[nt * 8]);
    }
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) Xw[(32 * j + 8 * nt + 2 * q + e) * 8 + g] = acc[nt][e];
    __syncwarp();
    double out[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int kk = 0; kk < 32; kk += 4) {
      const double a = Xw[(32 * j + kk + q) * 8 + g];
      const double* bs = Ls + (32 * j + kk + q) * LDS + 32 * j + g;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dmma(out[nt][0], out[nt][1], a, bs[nt * 8]);
    }
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = 32 * j + 8 * nt + 2 * q + e;
        Xw[c * 8 + g] = out[nt][e];
        if (row < n_rows && c < nb) S[(size_t)(k0 + c) * ld + row] = out[nt][e];
      }
    __syncwarp();
  }
  TTICK(9);
}

__global__ void __launch_bounds__(TS_THREADS, 1)
k_trsm_sub(double* __restrict__ S, int ld, int n_rows, int k0, int nb, int row0, const double* __restrict__ Linv) {
  extern __shared__ __align__(16) double smem[];
  trsm64_dev(S, ld, n_rows, k0, nb, row0 + (int)blockIdx.x * TS_ROWS, Linv, smem, TS_THREADS / 32);
}

// ---- diagonal block: Cholesky + inverse of the factor, one CTA of 512 threads -----------------
constexpr int PT = 512;          // threads of the diagonal-block kernel
constexpr int PLD = NB + 1;      // odd leading dimension: column reads by consecutive lanes conflict-free
constexpr int TLD = 97;          // leading dimension of the 32 x 96 product scratch
constexpr int POTRF_SMEM = (NB * PLD + NB + 32 * TLD + 32 * 33) * (int)sizeof(double);
constexpr unsigned FULL = 0xffffffffu;
#ifdef STBA_CHOL_TIMING
__device__ long long g_potrf_clk[64];
__device__ int g_tick_k0 = -1;      // panel whose phases are recorded (-1: every panel, i.e. the last one survives)
#define TICK(i) do { if (threadIdx.x == 0 && (g_tick_k0 < 0 || g_tick_k0 == k0)) g_potrf_clk[i] = clock64(); } while (0)
#else
#define TICK(i) do {} while (0)
#endif

__device__ __forceinline__ void potrf128_dev(double* __restrict__ S, int ld, int k0, int nb, double* __restrict__ Linv,
                                             int* __restrict__ info, double* sm) {
  double* D = sm;                 // D[c * PLD + r]: lower triangle + diagonal = the factor L;
                                  // strict upper triangle = the inverse, transposed: X(r,c), r > c, at D[r * PLD + c]
  double* xd = sm + NB * PLD;     // diagonal of the inverse
  double* Tm = xd + NB;           // Tm[rr * TLD + cc]: 32 x (32 bi) product scratch
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  TICK(0);
  for (int e = tid; e < NB * NB; e += PT) {
    const int r = e % NB, c = e / NB;
    D[c * PLD + r] = (r < nb && c < nb && r >= c) ? __ldcg(S + (size_t)(k0 + c) * ld + k0 + r) : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  TICK(1);
  for (int b0 = 0; b0 < NB; b0 += 32) {
    // (1) warp 0: lane = row of the 32 x 32 sub-block, the row lives in registers.  At step j every
    //     lane publishes its (unscaled) column-j entry in shared memory; one __syncwarp later all
    //     lanes read the column back as broadcasts.  (Shuffles inside this warp-specialised branch
    //     compile to slow WARPSYNC.COLLECTIVE call sequences — 4x slower, see profiles/.)
    if (warp == 0) {
      double a[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? D[(b0 + c) * PLD + b0 + lane] : 0.0;
      double* cb = Tm;                 // 2 x 32 column buffers (Tm is free during the factorisation)
      bool bad = false;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        double* col = cb + (j & 1) * 32;
        col[lane] = a[j];
        __syncwarp();
        const double d = col[j];
        if (!(d > 0.0) && !bad) { bad = true; if (lane == 0 && b0 + j < nb) atomicCAS(info, 0, k0 + b0 + j + 1); }
        // one reciprocal square root per pivot (seed + Newton, no special-case branches): both the
        // Cholesky column (a_ij / sqrt(d)) and the update factor (a_ij / d) come from it
        const double inv = fast_rsqrt(d);
        const double lj = a[j] * inv;
        const double t = lj * inv;
#pragma unroll
        for (int k = j + 1; k < 32; ++k) a[k] = fma(-t, col[k], a[k]);   // lanes < k: unused upper values
        a[j] = (lane == j) ? d * inv : lj;
        if (lane == j) xd[b0 + j] = inv;         // 1 / L_jj = diagonal of the inverse
      }
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c <= lane) D[(b0 + c) * PLD + b0 + lane] = a[c];
      TICK(20 + b0 / 32);
    }
    __syncthreads();
    TICK(2 + (b0 / 32) * 3);
    // (2) rows below the sub-block: x L_bb^T = a by forward substitution, one thread per row (the row
    //     lives in registers, L_bb is read as broadcasts).  No inverse is needed on this chain: the
    //     four 32 x 32 inverse blocks are computed side by side after the loop.
    const int below = NB - b0 - 32;
    if (tid < below) {
      const int r = b0 + 32 + tid;
      double sr[32];
#pragma unroll
      for (int p = 0; p < 32; ++p) sr[p] = D[(b0 + p) * PLD + r];
#pragma unroll
      for (int p = 0; p < 32; ++p) {
        const double x = sr[p] * xd[b0 + p];
        sr[p] = x;
        const double* lp = D + (b0 + p) * PLD + b0;
#pragma unroll
        for (int c = p + 1; c < 32; ++c) sr[c] = fma(-x, lp[c], sr[c]);
      }
#pragma unroll
      for (int p = 0; p < 32; ++p) D[(b0 + p) * PLD + r] = sr[p];
    }
    __syncthreads();
    TICK(3 + (b0 / 32) * 3);
    // (3) symmetric rank-32 update of the remaining lower triangle: task = (row, group of 4
    //     columns); consecutive lanes take consecutive rows (conflict-free), columns broadcast
    const int cg = below / 4;
    for (int e = tid; e < below * cg; e += PT) {
      const int rr = e % below, c4 = (e / below) * 4;
      if (rr < c4) continue;
      const int r = b0 + 32 + rr, c = b0 + 32 + c4;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 8
      for (int p = 0; p < 32; ++p) {
        const double* col = D + (b0 + p) * PLD;
        const double lr = col[r];
        s0 = fma(lr, col[c], s0); s1 = fma(lr, col[c + 1], s1); s2 = fma(lr, col[c + 2], s2); s3 = fma(lr, col[c + 3], s3);
      }
      D[c * PLD + r] -= s0;
      if (rr >= c4 + 1) D[(c + 1) * PLD + r] -= s1;
      if (rr >= c4 + 2) D[(c + 2) * PLD + r] -= s2;
      if (rr >= c4 + 3) D[(c + 3) * PLD + r] -= s3;
    }
    __syncthreads();
    TICK(4 + (b0 / 32) * 3);
  }
  // write the factor back (lower triangle incl. diagonal)
  for (int e = tid; e < NB * NB; e += PT) {
    const int r = e % NB, c = e / NB;
    if (r < nb && c < nb && r >= c) S[(size_t)(k0 + c) * ld + k0 + r] = D[c * PLD + r];
  }
  TICK(14);
  // inverses of the four 32 x 32 diagonal factor blocks, one warp each: lane = column, L(i,p) read
  // back as broadcasts; X(b0+i, c), i > c, goes to its transposed slot D[(b0+i) * PLD + c]
  if (warp < 4) {
    const int b0 = 32 * warp;
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
      for (int p = 0; p < i; ++p) {
        const double lip = D[(b0 + p) * PLD + b0 + i];
        if (p & 1) s1 = fma(-lip, x[p], s1); else s0 = fma(-lip, x[p], s0);
      }
      x[i] = (s0 + s1) * xd[b0 + i];
    }
    __syncwarp();          // every lane has finished reading the factor block before the slots above it are filled
    const int c = b0 + lane;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i > lane) D[(b0 + i) * PLD + c] = x[i];
  }
  __syncthreads();
  // The panel solve (k_trsm_sub) only needs the four 32 x 32 diagonal blocks of the inverse; the
  // off-diagonal blocks (needed by the backward substitution alone) are completed by
  // k_inv_offdiag on a side stream, off the critical path of the factorisation.
  for (int e = tid; e < NB * 32; e += PT) {
    const int c = e / 32, r = (c & ~31) + (e % 32);
    if (r < c || r >= nb || c >= nb) continue;
    Linv[(size_t)c * NB + r] = r > c ? D[r * PLD + c] : xd[r];      // column-major
  }
  TICK(15);
  TICK(16);
}

__global__ void __launch_bounds__(PT, 1)
k_potrf128(double* __restrict__ S, int ld, int k0, int nb, double* __restrict__ Linv, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  potrf128_dev(S, ld, k0, nb, Linv, info, sm);
}

// Completes Linv_kk: off-diagonal 32 x 32 blocks X_ib = -X_ii (sum_{p=b}^{i-1} L_ip X_pb), from the
// factor (in S) and the diagonal inverse blocks k_potrf128 stored.  One CTA; runs concurrently with
// the panel solve / trailing update of the same step.
__device__ __forceinline__ void inv_offdiag_dev(const double* __restrict__ S, int ld, int k0, int nb, double* __restrict__ Linv,
                                                double* sm) {
  double* D = sm;
  double* xd = sm + NB * PLD;
  double* Tm = xd + NB;
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += PT) {
    const int r = e % NB, c = e / NB;
    if (r >= c) D[c * PLD + r] = (r < nb && c < nb) ? __ldcg(S + (size_t)(k0 + c) * ld + k0 + r) : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();       // (lower part written before the upper slots of the same columns are filled)
  for (int e = tid; e < NB * 32; e += PT) {
    const int c = e / 32, r = (c & ~31) + (e % 32);
    if (r < c) continue;
    const double v = __ldcg(Linv + (size_t)c * NB + r);
    if (r == c) xd[r] = v; else D[r * PLD + c] = v;
  }
  __syncthreads();
  // ---- off-diagonal blocks of the inverse, one block row at a time:
  //      X_ib = -X_ii (sum_{p=b}^{i-1} L_ip X_pb)
  for (int bi = 1; bi < 4; ++bi) {
    const int w = 32 * bi;
    // T[rr][cc] = sum_{p=cc}^{w-1} L[w+rr][p] X[p][cc]; one thread = one row x 4 columns (5 shared loads per 4 FMAs)
    for (int e = tid; e < 32 * (w / 4); e += PT) {
      const int rr = e % 32, cc0 = (e / 32) * 4;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      {
        // triangular corner p = cc0 .. cc0+3: column c takes part from p >= c, X[c][c] = xd[c]
        const double l0 = D[(cc0 + 0) * PLD + w + rr], l1 = D[(cc0 + 1) * PLD + w + rr];
        const double l2 = D[(cc0 + 2) * PLD + w + rr], l3 = D[(cc0 + 3) * PLD + w + rr];
        s0 = l0 * xd[cc0];
        s0 = fma(l1, D[(cc0 + 1) * PLD + cc0], s0); s1 = l1 * xd[cc0 + 1];
        s0 = fma(l2, D[(cc0 + 2) * PLD + cc0], s0); s1 = fma(l2, D[(cc0 + 2) * PLD + cc0 + 1], s1); s2 = l2 * xd[cc0 + 2];
        s0 = fma(l3, D[(cc0 + 3) * PLD + cc0], s0); s1 = fma(l3, D[(cc0 + 3) * PLD + cc0 + 1], s1);
        s2 = fma(l3, D[(cc0 + 3) * PLD + cc0 + 2], s2); s3 = l3 * xd[cc0 + 3];
      }
#pragma unroll 4
      for (int p = cc0 + 4; p < w; ++p) {
        const double l = D[p * PLD + w + rr];
        const double* xp = D + p * PLD + cc0;
        s0 = fma(l, xp[0], s0); s1 = fma(l, xp[1], s1); s2 = fma(l, xp[2], s2); s3 = fma(l, xp[3], s3);
      }
      double* tp = Tm + rr * TLD + cc0;
      tp[0] = s0; tp[1] = s1; tp[2] = s2; tp[3] = s3;
    }
    __syncthreads();
    // X[w+rr][cc] = -sum_{p<=rr} X_ii[rr][p] T[p][cc]
    for (int e = tid; e < 32 * (w / 4); e += PT) {
      const int rr = e % 32, cc0 = (e / 32) * 4;
      const double xdd = xd[w + rr];
      const double* tr = Tm + rr * TLD + cc0;
      double s0 = xdd * tr[0], s1 = xdd * tr[1], s2 = xdd * tr[2], s3 = xdd * tr[3];
      const double* xi = D + (w + rr) * PLD + w;
      for (int p = 0; p < rr; ++p) {
        const double v = xi[p];
        const double* tq = Tm + p * TLD + cc0;
        s0 = fma(v, tq[0], s0); s1 = fma(v, tq[1], s1); s2 = fma(v, tq[2], s2); s3 = fma(v, tq[3], s3);
      }
      double* o = D + (w + rr) * PLD + cc0;
      o[0] = -s0; o[1] = -s1; o[2] = -s2; o[3] = -s3;
    }
    __syncthreads();
  }
  for (int e = tid; e < NB * NB; e += PT) {
    const int r = e % NB, c = e / NB;
    if ((r >> 5) <= (c >> 5) || r >= nb || c >= nb) continue;
    Linv[(size_t)c * NB + r] = D[r * PLD + c];
  }
}

__global__ void __launch_bounds__(PT, 1)
k_inv_offdiag(const double* __restrict__ S, int ld, int k0, int nb, double* __restrict__ Linv) {
  extern __shared__ __align__(16) double sm[];
  inv_offdiag_dev(S, ld, k0, nb, Linv, sm);
}

// ---- backward substitution, ONE launch --------------------------------------------------------
// x = L^-T y by 128-blocks.  CTA j owns block j: it keeps y_j in shared memory, applies
// y_j -= L_kj^T x_k for k = T-1 ... j+1 as each x_k is published (flag in global memory), then
// computes x_j = Linv_jj^T y_j and publishes it.  L_kj is prefetched into shared memory with
// cp.async while the CTA waits for x_k, Linv_jj^T sits packed in shared memory from the start:
// the dependent chain per block is flag -> 128 x 128 matvec from smem -> matvec -> flag, instead
// of one kernel launch per block (47 launches, 0.8 ms at n = 5988).  CTA j waits only on CTAs
// with a smaller blockIdx, so the kernel cannot deadlock even if not all CTAs are resident.
constexpr int TBA_THREADS = 256;
constexpr int TBA_SMEM = (NB * NB + NB * (NB + 1) / 2 + 2 * NB) * (int)sizeof(double);

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(TBA_THREADS, 1)
k_trsv_bwd_all(const double* __restrict__ S, int ld, int n, int T, const double* __restrict__ Linv, const double* __restrict__ y,
               double* x, int* flags, int* __restrict__ info, int blk0, int poll_data = 0) {
  // poll_data: x was pre-filled with NaNs (0xFF bytes); every element is its own "ready" flag — the consumers spin
  // on the values themselves (one L2 round trip per hop instead of flag + data) and the producer needs neither a
  // fence nor a flag store.  8-byte stores are single-copy atomic; a NaN solution only occurs after info != 0.
  extern __shared__ __align__(16) double sm[];
  double* Lb = sm;                          // Lb[c * NB + r] = L(k0 + r, j0 + c)
  double* U = Lb + NB * NB;                 // U[c (c + 1) / 2 + r] = Linv_jj(c, r), r <= c
  double* yj = U + NB * (NB + 1) / 2;
  double* xk = yj + NB;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int j = T - 1 - (blk0 + (int)blockIdx.x), j0 = j * NB;
  const double* Li = Linv + (size_t)j * NB * NB;     // column-major: Linv(c, r) at Li[r * NB + c]
  for (int e = t; e < NB * (NB + 1) / 2; e += TBA_THREADS) {
    int c = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
    while (c * (c + 1) / 2 > e) --c;
    while ((c + 1) * (c + 2) / 2 <= e) ++c;
    const int r = e - c * (c + 1) / 2;
    U[e] = Li[(size_t)r * NB + c];
  }
  if (t < NB) yj[t] = (j0 + t < n) ? y[j0 + t] : 0.0;
  __syncthreads();
  for (int k = T - 1; k > j; --k) {
    const int k0 = k * NB, nbk = min(NB, n - k0);
    // prefetch L_kj (128 columns x nbk rows) while x_k is still being computed elsewhere
    for (int e = t; e < NB * (NB / 2); e += TBA_THREADS) {
      const int c = e / (NB / 2), r2 = (e % (NB / 2)) * 2;
      const double* src = S + (size_t)(j0 + c) * ld + k0 + r2;
      if (r2 + 1 < nbk) {
        cp_async16(Lb + c * NB + r2, src, true);
      } else {
        Lb[c * NB + r2] = (r2 < nbk) ? __ldg(src) : 0.0;
        Lb[c * NB + r2 + 1] = 0.0;
      }
    }
    cp_async_commit();
    if (poll_data) {
      if (t < NB) {
        double v = 0.0;
        if (t < nbk) {
          long long spins = 0;
          const volatile double* src = x + k0 + t;
          for (;;) {
            v = *src;
            if (v == v) break;
            if (++spins > (1ll << 24)) { atomicCAS(info, 0, -1); v = 0.0; break; }    // never hang the device
          }
        }
        xk[t] = v;
      }
    } else {
      if (t == 0) {
        long long spins = 0;
        while (ld_acquire(flags + k) == 0) {
          if (++spins > (1ll << 26)) { atomicCAS(info, 0, -1); break; }    // never hang the device
        }
      }
      __syncthreads();
      if (t < NB) xk[t] = (t < nbk) ? __ldcg(x + k0 + t) : 0.0;
    }
    cp_async_wait<0>();
    __syncthreads();
    const double x0 = xk[lane], x1 = xk[lane + 32], x2 = xk[lane + 64], x3 = xk[lane + 96];
#pragma unroll 4
    for (int cc = 0; cc < 16; ++cc) {
      const int c = warp * 16 + cc;
      const double* col = Lb + c * NB + lane;
      double acc = fma(col[0], x0, fma(col[32], x1, fma(col[64], x2, col[96] * x3)));
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(FULL, acc, sft);
      if (lane == 0) yj[c] -= acc;
    }
    __syncthreads();
  }
  // x_j = Linv_jj^T y_j : x[r] = sum_{c >= r} Linv(c, r) y[c]; two threads per row (even / odd c)
  {
    const int r = t & 127, h = t >> 7;
    double s0 = 0.0, s1 = 0.0;
    int c = r + h;
    for (; c + 2 < NB; c += 4) {
      s0 = fma(U[c * (c + 1) / 2 + r], yj[c], s0);
      s1 = fma(U[(c + 2) * (c + 3) / 2 + r], yj[c + 2], s1);
    }
    for (; c < NB; c += 2) s0 = fma(U[c * (c + 1) / 2 + r], yj[c], s0);
    xk[r] = 0.0;
    __syncthreads();
    if (h == 1) xk[r] = s0 + s1;
    __syncthreads();
    if (h == 0 && j0 + r < n) {
      double v = (s0 + s1) + xk[r];
      if (poll_data && !(v == v)) { atomicCAS(info, 0, -1); v = 0.0; }      // a NaN would read as "not ready" downstream
      x[j0 + r] = v;
    }
  }
  if (poll_data) return;
  __threadfence();
  __syncthreads();
  if (t == 0) st_release(flags + j, 1);
}

// ---- triangular solves on an existing factor (any potrf) ---------------------------------------
// k_inv128: inverse of the 128 x 128 lower-triangular diagonal block k of a finished factor, one CTA
// per block (all T blocks in parallel): four 32 x 32 inverses on four warps, then the off-diagonal
// blocks exactly as k_inv_offdiag.
__global__ void __launch_bounds__(PT, 1)
k_inv128(const double* __restrict__ S, int ld, int n, double* __restrict__ LinvAll) {
  extern __shared__ __align__(16) double sm[];
  double* D = sm;
  double* xd = sm + NB * PLD;
  double* Tm = xd + NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k0 = (int)blockIdx.x * NB, nb = min(NB, n - k0);
  double* Linv = LinvAll + (size_t)blockIdx.x * NB * NB;
  for (int e = tid; e < NB * NB; e += PT) {
    const int r = e % NB, c = e / NB;
    if (r >= c) D[c * PLD + r] = (r < nb && c < nb) ? S[(size_t)(k0 + c) * ld + k0 + r] : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  if (tid < NB) xd[tid] = 1.0 / D[tid * PLD + tid];
  __syncthreads();
  if (warp < 4) {
    const int b0 = 32 * warp;
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
      for (int p = 0; p < i; ++p) {
        const double lip = D[(b0 + p) * PLD + b0 + i];
        if (p & 1) s1 = fma(-lip, x[p], s1); else s0 = fma(-lip, x[p], s0);
      }
      x[i] = (s0 + s1) * xd[b0 + i];
    }
    __syncwarp();
    const int c = b0 + lane;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i > lane) D[(b0 + i) * PLD + c] = x[i];
  }
  __syncthreads();
  for (int bi = 1; bi < 4; ++bi) {
    const int w = 32 * bi;
    // T[rr][cc] = sum_{p=cc}^{w-1} L[w+rr][p] X[p][cc]; one thread = one row x 4 columns (5 shared loads per 4 FMAs)
    for (int e = tid; e < 32 * (w / 4); e += PT) {
      const int rr = e % 32, cc0 = (e / 32) * 4;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      {
        // triangular corner p = cc0 .. cc0+3: column c takes part from p >= c, X[c][c] = xd[c]
        const double l0 = D[(cc0 + 0) * PLD + w + rr], l1 = D[(cc0 + 1) * PLD + w + rr];
        const double l2 = D[(cc0 + 2) * PLD + w + rr], l3 = D[(cc0 + 3) * PLD + w + rr];
        s0 = l0 * xd[cc0];
        s0 = fma(l1, D[(cc0 + 1) * PLD + cc0], s0); s1 = l1 * xd[cc0 + 1];
        s0 = fma(l2, D[(cc0 + 2) * PLD + cc0], s0); s1 = fma(l2, D[(cc0 + 2) * PLD + cc0 + 1], s1); s2 = l2 * xd[cc0 + 2];
        s0 = fma(l3, D[(cc0 + 3) * PLD + cc0], s0); s1 = fma(l3, D[(cc0 + 3) * PLD + cc0 + 1], s1);
        s2 = fma(l3, D[(cc0 + 3) * PLD + cc0 + 2], s2); s3 = l3 * xd[cc0 + 3];
      }
#pragma unroll 4
      for (int p = cc0 + 4; p < w; ++p) {
        const double l = D[p * PLD + w + rr];
        const double* xp = D + p * PLD + cc0;
        s0 = fma(l, xp[0], s0); s1 = fma(l, xp[1], s1); s2 = fma(l, xp[2], s2); s3 = fma(l, xp[3], s3);
      }
      double* tp = Tm + rr * TLD + cc0;
      tp[0] = s0; tp[1] = s1; tp[2] = s2; tp[3] = s3;
    }
    __syncthreads();
    // X[w+rr][cc] = -sum_{p<=rr} X_ii[rr][p] T[p][cc]
    for (int e = tid; e < 32 * (w / 4); e += PT) {
      const int rr = e % 32, cc0 = (e / 32) * 4;
      const double xdd = xd[w + rr];
      const double* tr = Tm + rr * TLD + cc0;
      double s0 = xdd * tr[0], s1 = xdd * tr[1], s2 = xdd * tr[2], s3 = xdd * tr[3];
      const double* xi = D + (w + rr) * PLD + w;
      for (int p = 0; p < rr; ++p) {
        const double v = xi[p];
        const double* tq = Tm + p * TLD + cc0;
        s0 = fma(v, tq[0], s0); s1 = fma(v, tq[1], s1); s2 = fma(v, tq[2], s2); s3 = fma(v, tq[3], s3);
      }
      double* o = D + (w + rr) * PLD + cc0;
      o[0] = -s0; o[1] = -s1; o[2] = -s2; o[3] = -s3;
    }
    __syncthreads();
  }
  for (int e = tid; e < NB * NB; e += PT) {
    const int r = e % NB, c = e / NB;
    if (r < c || r >= nb || c >= nb) continue;
    Linv[(size_t)c * NB + r] = r > c ? D[r * PLD + c] : xd[r];
  }
}

// Forward substitution y = L^-1 b in ONE launch, mirror image of k_trsv_bwd_all: CTA j keeps b_j in
// shared memory, applies b_j -= L_jk y_k for k = 0 .. j-1 as each y_k is published, then y_j = Linv_jj b_j.
__global__ void __launch_bounds__(TBA_THREADS, 1)
k_trsv_fwd_all(const double* __restrict__ S, int ld, int n, int T, const double* __restrict__ Linv, const double* __restrict__ b,
               double* y, int* flags, int* __restrict__ info, int blk0) {
  extern __shared__ __align__(16) double sm[];
  double* Lb = sm;                          // Lb[c * NB + r] = L(j0 + r, k0 + c)
  double* U = Lb + NB * NB;                 // U[c * NB - c (c - 1) / 2 + (r - c)] = Linv_jj(r, c), r >= c
  double* bj = U + NB * (NB + 1) / 2;
  double* yk = bj + NB;
  const int t = threadIdx.x;
  const int j = blk0 + (int)blockIdx.x, j0 = j * NB, nbj = min(NB, n - j0);
  const double* Li = Linv + (size_t)j * NB * NB;
  for (int c = 0; c < NB; ++c)
    for (int r = c + t; r < NB; r += TBA_THREADS) U[c * NB - c * (c - 1) / 2 + (r - c)] = Li[(size_t)c * NB + r];
  if (t < NB) bj[t] = (t < nbj) ? b[j0 + t] : 0.0;
  __syncthreads();
  for (int k = 0; k < j; ++k) {
    const int k0 = k * NB;
    for (int e = t; e < NB * (NB / 2); e += TBA_THREADS) {
      const int c = e / (NB / 2), r2 = (e % (NB / 2)) * 2;
      const double* src = S + (size_t)(k0 + c) * ld + j0 + r2;
      if (r2 + 1 < nbj) {
        cp_async16(Lb + c * NB + r2, src, true);
      } else {
        if (r2 < nbj) cp_async8(Lb + c * NB + r2, src); else Lb[c * NB + r2] = 0.0;
        Lb[c * NB + r2 + 1] = 0.0;
      }
    }
    cp_async_commit();
    if (t == 0) {
      long long spins = 0;
      while (ld_acquire(flags + k) == 0) {
        if (++spins > (1ll << 26)) { atomicCAS(info, 0, -1); break; }
      }
    }
    __syncthreads();
    if (t < NB) yk[t] = __ldcg(y + k0 + t);             // block k < j is always a full block
    cp_async_wait<0>();
    __syncthreads();
    {
      const int r = t & 127, h = t >> 7;                 // two threads per row: columns [64 h, 64 h + 64)
      double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
      for (int c = 64 * h; c < 64 * h + 64; c += 2) {
        s0 = fma(Lb[c * NB + r], yk[c], s0);
        s1 = fma(Lb[(c + 1) * NB + r], yk[c + 1], s1);
      }
      __syncthreads();
      if (h == 1) yk[r] = s0 + s1;                       // yk is free again: park the upper half's sum
      __syncthreads();
      if (h == 0) bj[r] -= (s0 + s1) + yk[r];
    }
    __syncthreads();
  }
  {
    const int r = t & 127, h = t >> 7;                   // y_j = Linv_jj b_j : sum over c <= r, even / odd c
    double s0 = 0.0;
    for (int c = h; c <= r; c += 2) s0 = fma(U[c * NB - c * (c - 1) / 2 + (r - c)], bj[c], s0);
    __syncthreads();
    if (h == 1) yk[r] = s0;
    __syncthreads();
    if (h == 0 && j0 + r < n) y[j0 + r] = s0 + yk[r];
  }
  __threadfence();
  __syncthreads();
  if (t == 0) st_release(flags + j, 1);
}

// rhs -> row n of S (the augmented row) and back
__global__ void k_put_row(double* __restrict__ S, int ld, int n, const double* __restrict__ rhs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) S[(size_t)c * ld + n] = rhs[c];
}
__global__ void k_get_row(const double* __restrict__ S, int ld, int n, double* __restrict__ y) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) y[c] = S[(size_t)c * ld + n];
}

// =================================================================================================
// DAG-scheduled factorisation: ONE persistent kernel, one CTA per SM, tile tasks with data-flow
// dependencies published through version counters in global memory (STBA_DENSE_OWN).
//
// A right-looking blocked Cholesky launched as a chain of kernels is bound, for the last two thirds of
// the block columns, by the serial chain  potrf(k) -> panel solve -> update of column k+1 -> potrf(k+1)
// (~80 us per 128 columns at n = 5988: launch gaps, kernel tails, whole-panel kernels on the chain; see
// profiles/r1_dense_notes.md).  Here only the tile operations that truly are on the critical path stay on
// it, on two CTAs that do nothing else and hand each other 32-column block columns as they complete:
//     CTA 0   potrf of the diagonal block k, as soon as tile (k, k) has received update k - 1
//             (potrf128_prog_dev: publishes block column b and the inverse of its diagonal block);
//     CTA 1   solve of tile (k+1, k) and update of tile (k+1, k+1), block column by block column (tu_dev);
// every other SM is a worker pulling tile tasks from three queues by ticket (one atomicAdd), polling the
// inputs of the tasks it holds:
//     high priority  64-row panel solves below tile row k+1, each followed on the same worker by the
//                    64-row update of its rows in block column k+1;
//     urgent         single-panel 128 x 128 updates of block columns k+2 .. k+1+W, nearest the diagonal first,
//                    and the partial chunks of a column entering that window;
//     far            the rest of the trailing matrix, G block columns per update (K = 128 G: one C-tile round
//                    trip per G panels), plus the completion of the diagonal inverses for the backward solve.
// No kernel boundaries, no tails; the C tile of a held ticket is prefetched into L2 while the previous task
// runs and staged through shared memory by TMA bulk copies.
//
// Versions: ver[h][j] counts the block-column updates applied to the 64-row half-tile h of block column j;
// j + 1 means "final" (panel solve done).  pdone[k] = 1: L_kk and its diagonal inverse blocks are in memory;
// pb[k] / pinv[k]: block columns / inverse blocks of panel k published so far.  Every counter has a single
// writer at any time (updates of one tile are serialised by the version itself), so publication is a release
// store after a __threadfence; consumers poll with relaxed loads and fence (or re-load with acquire) once.
// All operand reads bypass L1 (cp.async.cg / ld.cg / bulk copies): they were written by other SMs during the
// same launch.  Measured history and the per-CTA cycle / globaltimer profiles: profiles/r2_dense_notes.md.
// =================================================================================================
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TRACE(P, k, slot) do { if ((P).trace && threadIdx.x == 0) (P).trace[(size_t)(k) * 16 + (slot)] = gtime(); } while (0)
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- progressive diagonal-block factorisation (DAG kernel, CTA 0) ---------------------------------
// Same arithmetic as potrf128_dev, reorganised around the pivot chain, which is what bounds the whole
// factorisation once the trailing update runs concurrently:
//   * warp 0 factors the 32 x 32 sub-block with the NEXT column published before the rest of the rank-1
//     update (the remaining FMAs leave the chain) and the pivot's reciprocal from rcp + Newton (the
//     reciprocal square root that scales the column is computed beside the chain, not on it);
//   * the rank-32 update of the rest of the block runs on the FP64 tensor pipe (warps 0..13);
//   * warp 14 copies every finished 32-column block column to global memory and publishes it (pb), warp 15
//     inverts the 32 x 32 diagonal factor blocks as they appear and publishes them (pinv): the consumers of
//     the panel (the solve of the next tile row) start on block column b while b + 1 is being factored.
constexpr int QLD = NB + 4;      // (q QLD + g) mod 16 distinct over a half-warp: conflict-free DMMA fragments
constexpr int PROG_SMEM = (NB * QLD + NB + 32 * 32 + 2 * 32 * QLD) * (int)sizeof(double);

__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, 480;" ::: "memory"); }

// 1 / d: hardware seed (~2^-22) and ONE cubically convergent correction y (1 + e + e^2), e = 1 - d y:
// three dependent FP64 operations instead of the four of two Newton steps (error e^3 ~ 2^-66).
__device__ __forceinline__ double fast_rcp(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}

__device__ __forceinline__ void potrf128_prog_dev(double* __restrict__ S, int ld, int k0, int nb, double* __restrict__ Linv,
                                                  int* __restrict__ info, double* sm, int* pb_flag, int* pinv_flag,
                                                  volatile int* s_sig, volatile int* s_prog,
                                                  const int* xflag = nullptr, int xbase = 0, int n_rows = 0, const int* abort_flag = nullptr,
                                                  const int* xflag2 = nullptr) {
  double* D = sm;                  // D[c * QLD + r], lower triangle
  double* xd = sm + NB * QLD;      // 1 / L_jj
  double* cb = xd + NB;            // 32 x 32: column j of the pivot block as it was when it became the pivot column
  double* Xb = cb + 32 * 32;       // 32 x QLD: one block column of the previous panel's rows of this tile (streamed update)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  TICK(0);
  // ---- load the lower triangle (16-byte coalesced), identity outside the matrix ----
  {
    double2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int e = tid + i * 512;
      const int c = e >> 6, r2 = (e & 63) * 2;
      v[i] = (c < nb && r2 < nb && r2 + 1 >= c) ? __ldcg(reinterpret_cast<const double2*>(S + (size_t)(k0 + c) * ld + k0 + r2)) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int e = tid + i * 512;
      const int c = e >> 6, r2 = (e & 63) * 2;
      double x0 = v[i].x, x1 = v[i].y;
      if (!(r2 < nb && c < nb)) x0 = (r2 == c) ? 1.0 : 0.0;
      if (!(r2 + 1 < nb && c < nb)) x1 = (r2 + 1 == c) ? 1.0 : 0.0;
      D[c * QLD + r2] = x0;
      D[c * QLD + r2 + 1] = x1;
    }
  }
  if (tid == 0) { *s_sig = 0; *s_prog = 0; }
  __syncthreads();
  if (xflag) {
    // ---- the LAST update of this tile, A_kk -= X X^T with X = L(k, k-1), applied here while the solve of tile row k
    //      against panel k-1 is still publishing its block columns (xflag >= xbase + b + 1): the tile never makes the
    //      round trip  update CTA -> global memory -> flag -> this CTA  between the last solve step and the first pivot
    // Two buffers: the block column b + 1 is fetched (cp.async) while block column b is applied, if it has been published
    // already — in the steady state all but the last one have, long before this CTA gets here.
    int have = 0, avail = 0;          // block columns whose load has been issued / that are known to be published (CTA-uniform)
    // the flags are polled only when the next block column is not yet known to be there: a poll costs L2 round trips,
    // and in the steady state all four have been published long before this CTA gets here
    auto poll = [&](int want) {
      if (tid == 0) {
        long long spins = 0;
        int v = 0;
        for (;;) {
          v = ld_relaxed(xflag) - xbase;
          if (xflag2) v = min(v, ld_relaxed(xflag2) - xbase);
          if (v >= want) break;
          if ((++spins & 255) == 0 && abort_flag && ld_relaxed(abort_flag)) { v = 4; break; }
          if (spins > (1ll << 21)) { v = 4; break; }
        }
        __threadfence();
        *s_sig = min(v, 4);
      }
      __syncthreads();
      avail = *s_sig;
      __syncthreads();
    };
    auto issue = [&]() {
      double* Xd = Xb + (have & 1) * 32 * QLD;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = tid + i * 512;
        const int c = e >> 6, r2 = (e & 63) * 2;
        const double* src = S + (size_t)(k0 - NB + 32 * have + c) * ld + k0 + r2;
        if (r2 + 1 < nb) {
          cp_async16(Xd + c * QLD + r2, src, true);
        } else {
          Xd[c * QLD + r2] = (r2 < nb) ? __ldcg(src) : 0.0;
          Xd[c * QLD + r2 + 1] = 0.0;
        }
      }
      cp_async_commit();
      ++have;
    };
    poll(1);
    issue();
    for (int b = 0; b < 4; ++b) {
      if (have == b + 1 && have < avail) issue();          // next block column already published: fetch it behind this one
      if (have == b + 2) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncthreads();
      {
        const double* Xc = Xb + (b & 1) * 32 * QLD;
        int cnt = 0;
        for (int ti = 0; ti < 8; ++ti)
          for (int tj = 0; tj <= ti; ++tj, ++cnt) {
            if ((cnt & 15) != warp) continue;
            const int r0 = 16 * ti, c0 = 16 * tj;
            double acc[2][2][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
              for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) acc[mi][ni][e] = D[(c0 + 8 * ni + 2 * q + e) * QLD + r0 + 8 * mi + g];
#pragma unroll
            for (int kk = 0; kk < 32; kk += 4) {
              const double* colp = Xc + (kk + q) * QLD;
              const double a0 = -colp[r0 + g], a1 = -colp[r0 + 8 + g];
              const double bb0 = colp[c0 + g], bb1 = colp[c0 + 8 + g];
              dmma(acc[0][0][0], acc[0][0][1], a0, bb0);
              dmma(acc[0][1][0], acc[0][1][1], a0, bb1);
              dmma(acc[1][0][0], acc[1][0][1], a1, bb0);
              dmma(acc[1][1][0], acc[1][1][1], a1, bb1);
            }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
              for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) D[(c0 + 8 * ni + 2 * q + e) * QLD + r0 + 8 * mi + g] = acc[mi][ni][e];
          }
      }
      __syncthreads();
      if (have == b + 1 && b + 1 < 4) { poll(have + 1); issue(); }
    }
    // the right-hand-side row (row n of S) lives inside the last diagonal tile when n is not a multiple of 128: its
    // share of this update, one thread per column
    if (n_rows > k0 + nb && nb < NB && tid < nb) {
      const int rr = k0 + nb;          // = n
      double acc = 0.0;
      for (int kk = 0; kk < NB; ++kk) {
        const double* colp = S + (size_t)(k0 - NB + kk) * ld;
        acc = fma(__ldcg(colp + rr), __ldcg(colp + k0 + tid), acc);
      }
      S[(size_t)(k0 + tid) * ld + rr] -= acc;
    }
    if (tid == 0) *s_sig = 0;       // (borrowed as the broadcast slot of try_issue)
    __syncthreads();
  }
  TICK(1);
  if (warp == 15) {
    // ---- inverse of the 32 x 32 diagonal factor blocks, as they are produced ----
    for (int b = 0; b < 4; ++b) {
      const int b0 = 32 * b;
      while (*s_sig < b + 1) { }
      __threadfence_block();
      __syncwarp();
      double x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
        for (int p = 0; p < i; ++p) {
          const double lip = D[(b0 + p) * QLD + b0 + i];
          if (p & 1) s1 = fma(-lip, x[p], s1); else s0 = fma(-lip, x[p], s0);
        }
        x[i] = (s0 + s1) * xd[b0 + i];
      }
      const int c = b0 + lane;
      if (c < nb) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i >= lane && b0 + i < nb) Linv[(size_t)c * NB + b0 + i] = x[i];     // x[lane] = 1 / L_cc
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) st_release_gpu(pinv_flag, b + 1);
    }
  } else {
    for (int b = 0; b < 4; ++b) {
      const int b0 = 32 * b;
      const int base = 64 * b;       // *s_prog = base + number of complete columns of cb; base + 33: xd is complete too
      if (warp == b) {
        // (1) pivot chain, rows b0 .. b0 + 31: lane = row.  Column j + 1 is exchanged (shared memory) one dependent
        //     operation after the reciprocal of pivot j; the rest of the rank-1 update of step j is issued inside
        //     step j + 1, where it fills the latency of the exchange and of the reciprocal (84 cycles of
        //     dependent latency per pivot: tools/ubench/potf2.cu measures 137 for this loop, 243 for the
        //     plain one).  Branch-free: the pivots are checked once after the loop.
        double a[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? D[(b0 + c) * QLD + b0 + lane] : 0.0;
        cb[lane] = a[0];
        __threadfence_block();
        __syncwarp();
        if (lane == 0) *s_prog = base + 1;
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
            if ((j & 3) == 3 || j == 30) __threadfence_block();
            __syncwarp();
            if (((j & 3) == 3 || j == 30) && lane == 0) *s_prog = base + j + 2;      // the followers advance four columns at a time
          }
          tp = a[j] * r;
          if (j + 2 < 32) a[j + 2] = fma(-tp, col[j + 2], a[j + 2]);
        }
        // scale: L_ij = a_ij / sqrt(d_j); lane j owns pivot j
        const double dl = cb[lane * 32 + lane];
        const unsigned badm = __ballot_sync(FULL, !(dl > 0.0) && b0 + lane < nb);
        if (badm && lane == 0) atomicCAS(info, 0, k0 + b0 + __ffs(badm));
        const double rsl = fast_rsqrt(dl);
        xd[b0 + lane] = rsl;
        __threadfence_block();
        __syncwarp();
        if (lane == 0) *s_prog = base + 33;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const double v = (c == lane) ? dl * rsl : a[c] * xd[b0 + c];
          if (c <= lane) D[(b0 + c) * QLD + b0 + lane] = v;
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) *s_sig = b + 1;
        TICK(20 + b);
      } else if (warp > b && warp < 4) {
        // (2) rows below the pivot block follow the same elimination, four columns behind at most: the former
        //     forward substitution of these rows (a separate phase on the chain) is gone
        const int r = 32 * warp + lane;
        double a[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) a[c] = D[(b0 + c) * QLD + r];
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
          const int need = base + (j4 + 4 < 32 ? j4 + 4 : 32);
          while (*s_prog < need) { }
          __threadfence_block();
          double rc[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) rc[u] = fast_rcp(cb[(j4 + u) * 32 + j4 + u]);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = j4 + u;
            const double* col = cb + j * 32;
            const double t = a[j] * rc[u];
#pragma unroll
            for (int k = j + 1; k < 32; ++k) a[k] = fma(-t, col[k], a[k]);
          }
        }
        while (*s_prog < base + 33) { }
        __threadfence_block();
#pragma unroll
        for (int c = 0; c < 32; ++c) D[(b0 + c) * QLD + r] = a[c] * xd[b0 + c];
      }
      bar_workers();
      TICK(2 + 3 * b);
      TICK(3 + 3 * b);
      const int below = NB - b0 - 32;
      // Block column b is final: copy it out (direct stores from the row-owning warps were measured and cost 3-4 k
      // cycles per block on the pivot phase: profiles/r2_dense_notes.md).  While the other warps run the rank-32 update,
      // warp 14 copies and publishes; after the last block there is nothing to update and all 15 warps share the copy.
      if (warp == 14 || below == 0) {
        const int cstep = below == 0 ? 15 : 1, cfirst = below == 0 ? warp : 0;
        for (int c = b0 + cfirst; c < b0 + 32 && c < nb; c += cstep) {
          double* dst = S + (size_t)(k0 + c) * ld + k0;
          const double* srcc = D + c * QLD;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = b0 + lane + 32 * i;
            if (r >= c && r < nb) dst[r] = srcc[r];
          }
        }
        __threadfence();
        __syncwarp();
        if (below > 0 && lane == 0) st_release_gpu(pb_flag, b + 1);
      } else if (below > 0) {
        // (3) rank-32 update of the remaining lower triangle on the FP64 tensor pipe: 16 x 16 macro tiles (four
        //     independent accumulator pairs hide the DMMA latency), lower triangle of the (below / 16)^2 grid
        //     dealt round-robin to the 14 warps
        const int nt = below / 16;
        int cnt = 0;
        for (int ti = 0; ti < nt; ++ti)
          for (int tj = 0; tj <= ti; ++tj, ++cnt) {
            if (cnt % 14 != warp) continue;
            const int r0 = b0 + 32 + 16 * ti, c0 = b0 + 32 + 16 * tj;
            double acc[2][2][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
              for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) acc[mi][ni][e] = D[(c0 + 8 * ni + 2 * q + e) * QLD + r0 + 8 * mi + g];
#pragma unroll
            for (int kk = 0; kk < 32; kk += 4) {
              const double* colp = D + (b0 + kk + q) * QLD;
              const double a0 = -colp[r0 + g], a1 = -colp[r0 + 8 + g];
              const double bb0 = colp[c0 + g], bb1 = colp[c0 + 8 + g];
              dmma(acc[0][0][0], acc[0][0][1], a0, bb0);
              dmma(acc[0][1][0], acc[0][1][1], a0, bb1);
              dmma(acc[1][0][0], acc[1][0][1], a1, bb0);
              dmma(acc[1][1][0], acc[1][1][1], a1, bb1);
            }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
              for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) D[(c0 + 8 * ni + 2 * q + e) * QLD + r0 + 8 * mi + g] = acc[mi][ni][e];
          }
      }
      bar_workers();
      if (below == 0 && warp == 14 && lane == 0) st_release_gpu(pb_flag, b + 1);      // every warp's share is stored and fenced
      TICK(4 + 3 * b);
    }
  }
  TICK(14);
  __syncthreads();
  TICK(15);
  TICK(16);
}

constexpr int DAG_THREADS = 512;
constexpr int UPD_STAGES = 4;
constexpr int UPD_SMEM = UPD_STAGES * KC * (LDS + LDS) * (int)sizeof(double);
constexpr int TU_SMEM = (16 * NB * 8 + 32 * (132 + 100 + 68 + 36)) * (int)sizeof(double);
constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int DAG_SMEM = cmax(cmax(TS_SMEM, POTRF_SMEM), cmax(UPD_SMEM, cmax(TU_SMEM, PROG_SMEM)));
enum { TASK_EXIT = -1, TASK_TRSM = 0, TASK_UPD = 1, TASK_INV = 2, TASK_UPD64 = 3 };
enum { F_HP = 0, F_LP = 32, F_ABORT = 64, F_NPAN = 96, F_UR = 128, F_PDONE = 160 };     // queue heads and the abort flag on their own 128-byte lines
#ifndef STBA_SPIN_SHIFT
#define STBA_SPIN_SHIFT 21
#endif
constexpr long long SPIN_LIMIT = 1ll << STBA_SPIN_SHIFT;     // ~ seconds; a broken schedule must never hang the device

struct DagParams {
  double* S;
  int ld, n, n_rows, T, Tr, R64;
  double* Linv;
  int* info;
  int* flags;               // [F_HP] [F_LP] [F_ABORT] [F_NPAN] . pdone[T] ver[R64 * T] pb[T] pinv[T]
  const int4* hp;           // (type, i or half, j, k | n_k << 16): n_k block columns k .. k + n_k - 1 in one update
  const int4* ur;
  const int4* lp;
  int n_hp, n_ur, n_lp;
  const int* hp_end;        // hp_end[k]: one past the last high-priority task of step k
  const int* ur_end;        // same for the urgent queue
  unsigned long long* trace;   // [T][16] globaltimer stamps of the chain CTAs (profiling runs only)
  long long* prof;          // per CTA: [0] wait cycles [1] potrf [2] trsm [3] upd [4] inv [5] tasks [6] total
};

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared (SASS UBLKCP), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
constexpr int LDC = NB + 2;      // staged C tile: (2 q LDC + g) mod 16 distinct over a half-warp -> conflict-free fragment reads

// C(i0.., j0..) -= A(i0.., k0..k0+K) A(j0.., k0..k0+K)^T on a 128 x 128 tile, 16 warps (4 x 4), warp tile 32 x 32.
// Four warps per scheduler keep the FP64 tensor pipe fed through the per-chunk barriers; the accumulators
// start as the C tile (store-only epilogue).  On a diagonal tile the warps strictly above the diagonal skip
// the arithmetic.
template <int MT>
__device__ __forceinline__ void upd_dev(double* __restrict__ S, int lda, int n_rows, int n_cols, int k0, int K, int i0, int j0,
                                        double* smem, long long* ph, unsigned long long* cbar, unsigned& cphase) {
  constexpr int ROWS = 32 * MT;      // 128 (bulk and diagonal tiles) or 64 (the latency-critical updates of block column k + 1)
  const long long c_start = clock64();
  double* As = smem;                               // [UPD_STAGES][KC][LDS]
  double* Bs = smem + UPD_STAGES * KC * LDS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const bool diag = MT == 4 && i0 == j0;
  // Off-diagonal tile: 4 x 4 warp grid.  Diagonal tile: only the 10 warp tiles on or below the diagonal carry
  // arithmetic; they go to warps 0..9 so that the four schedulers (warp & 3) get 3, 3, 2, 2 of them instead
  // of 4, 3, 2, 1 — this tile is on the critical path of the factorisation.
  int wm = warp >> 2, wn = warp & 3;
  bool active = true;
  if (diag) {
    // warp:            0  1  2  3  4  5  6  7  8  9
    active = warp < 10;
    switch (warp) {
      case 0: wm = 0; wn = 0; break;  case 1: wm = 1; wn = 0; break;  case 2: wm = 1; wn = 1; break;  case 3: wm = 2; wn = 0; break;
      case 4: wm = 2; wn = 1; break;  case 5: wm = 2; wn = 2; break;  case 6: wm = 3; wn = 0; break;  case 7: wm = 3; wn = 1; break;
      case 8: wm = 3; wn = 2; break;  case 9: wm = 3; wn = 3; break;  default: wm = 0; wn = 0; break;
    }
  }
  const double* Ag = S + (size_t)k0 * lda;
  // The accumulators start as the C tile.  Loads are unconditional from clamped addresses (entries outside
  // the matrix are never stored), so nothing consumes a loaded value before the first DMMA.
  // The accumulators start as the C tile.  It is staged through the (still idle) pipeline buffers by TMA bulk
  // copies, one 1 KB column each: per-lane 8-byte loads of the fragment layout touch half of four 128-byte
  // lines per instruction and are throttled by the SM's outstanding-request capacity (~10 k cycles per tile,
  // measured); the bulk engine streams whole lines.  Entries outside the matrix are never stored.
  double acc[MT][4][2];
  {
    const int ncol = min(NB, n_cols - j0);
    const unsigned col_bytes = (unsigned)min(ROWS, lda - i0) * 8u;       // lda, i0 even: a multiple of 16
    if (tid == 0) mbar_expect_tx(cbar, col_bytes * (unsigned)ncol);
    __syncthreads();
    if (tid < ncol) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic-proxy writes to these bytes
      bulk_g2s(smem + tid * LDC, S + (size_t)(j0 + tid) * lda + i0, col_bytes, cbar);
    }
    mbar_wait(cbar, cphase & 1);
    ++cphase;
    if (active) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) acc[mt][nt][e] = smem[(wn * 32 + nt * 8 + 2 * q + e) * LDC + wm * (8 * MT) + mt * 8 + g];
    }
    __syncthreads();      // every fragment is in registers before the pipeline stages overwrite the tile
  }
  const int n_chunks = (K + KC - 1) / KC;
  auto load_stage = [&](int chunk, int stage) {
    double* as = As + stage * KC * LDS;
    double* bs = Bs + stage * KC * LDS;
#pragma unroll
    for (int p = 0; p < KC * 64 / DAG_THREADS; ++p) {
      const int piece = tid + p * DAG_THREADS;
      const int kk = piece >> 6, r2 = (piece & 63) * 2;
      const int k = chunk * KC + kk;
      const bool kin = k < K;
      const int ra = i0 + r2, rb = j0 + r2;
      if (r2 < ROWS) cp_async16(as + kk * LDS + r2, Ag + (size_t)(kin ? k : 0) * lda + (ra < n_rows ? ra : 0), kin && ra < n_rows);
      if (!diag) cp_async16(bs + kk * LDS + r2, Ag + (size_t)(kin ? k : 0) * lda + (rb < n_rows ? rb : 0), kin && rb < n_rows);
    }
  };
#pragma unroll
  for (int s = 0; s < UPD_STAGES - 1; ++s) {
    if (s < n_chunks) load_stage(s, s);
    cp_async_commit();
  }
  const long long c_issue = clock64();
  long long c_first = 0;
  for (int c = 0; c < n_chunks; ++c) {
    cp_async_wait<UPD_STAGES - 2>();
    __syncthreads();
    if (c == 0) c_first = clock64();
    if (c + UPD_STAGES - 1 < n_chunks) load_stage(c + UPD_STAGES - 1, (c + UPD_STAGES - 1) % UPD_STAGES);
    cp_async_commit();
    if (!active) continue;
    const double* as = As + (c % UPD_STAGES) * KC * LDS + wm * (8 * MT) + g;
    const double* bs = (diag ? As : Bs) + (c % UPD_STAGES) * KC * LDS + wn * 32 + g;
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
      double a[MT], b[4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a[mt] = -as[(kk + q) * LDS + mt * 8];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) b[nt] = bs[(kk + q) * LDS + nt * 8];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
  }
  cp_async_wait<0>();
  const long long c_loop = clock64();
  if (active) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int r = i0 + wm * (8 * MT) + mt * 8 + g;
      if (r >= n_rows) continue;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = j0 + wn * 32 + nt * 8 + 2 * q + e;
          if (c < n_cols) S[(size_t)c * lda + r] = acc[mt][nt][e];
        }
    }
  }
  __syncthreads();      // the stages are reused by the CTA's next task
  if (tid == 0) { ph[0] += c_issue - c_start; ph[1] += c_first - c_issue; ph[2] += c_loop - c_first; ph[3] += clock64() - c_loop; }
}


__device__ __forceinline__ bool half_exists(const DagParams& P, int h) { return h * 64 < P.n_rows; }

// thread 0: are the inputs of the task final?  (relaxed polls; the caller fences once after claiming)
__device__ __forceinline__ bool task_ready(const DagParams& P, int4 t) {
  const int* pdone = P.flags + F_PDONE;
  const int* ver = pdone + P.T;
  const int klo = t.w & 0xffff, k = klo + (t.w >> 16) - 1;      // k: the last block column the task reads
  if (t.x == TASK_TRSM) return ld_relaxed(pdone + k) != 0 && ld_relaxed(ver + t.y * P.T + k) == k;
  if (t.x == TASK_INV) return ld_relaxed(pdone + k) != 0;
  if (t.x == TASK_UPD64) {      // rows of half-tile t.y, block column t.z
    const int h = t.y, j = t.z;
    const int a0 = ld_relaxed(ver + h * P.T + k), c0 = ld_relaxed(ver + h * P.T + j);
    const int b0 = ld_relaxed(ver + (2 * j) * P.T + k), b1 = half_exists(P, 2 * j + 1) ? ld_relaxed(ver + (2 * j + 1) * P.T + k) : k + 1;
    return a0 == k + 1 && b0 == k + 1 && b1 == k + 1 && c0 == k;
  }
  const int i = t.y, j = t.z;
  const bool i2 = half_exists(P, 2 * i + 1);
  // L rows of tile i and tile j final for block column k; the C tile has received update k - 1
  const int a0 = ld_relaxed(ver + (2 * i) * P.T + k), a1 = i2 ? ld_relaxed(ver + (2 * i + 1) * P.T + k) : k + 1;
  const int b0 = ld_relaxed(ver + (2 * j) * P.T + k), b1 = half_exists(P, 2 * j + 1) ? ld_relaxed(ver + (2 * j + 1) * P.T + k) : k + 1;
  const int c0 = ld_relaxed(ver + (2 * i) * P.T + j), c1 = i2 ? ld_relaxed(ver + (2 * i + 1) * P.T + j) : klo;
  return a0 == k + 1 && a1 == k + 1 && b0 == k + 1 && b1 == k + 1 && c0 == klo && c1 == klo;
}

__device__ __forceinline__ void task_publish(const DagParams& P, int4 t) {
  int* ver = P.flags + F_PDONE + P.T;
  const int k = (t.w & 0xffff) + (t.w >> 16) - 1;
  if (t.x == TASK_TRSM) {
    st_release_gpu(ver + t.y * P.T + k, k + 1);
  } else if (t.x == TASK_UPD64) {
    st_release_gpu(ver + t.y * P.T + t.z, k + 1);
  } else if (t.x == TASK_UPD) {
    st_release_gpu(ver + (2 * t.y) * P.T + t.z, k + 1);
    if (half_exists(P, 2 * t.y + 1)) st_release_gpu(ver + (2 * t.y + 1) * P.T + t.z, k + 1);
  }
}

__device__ __forceinline__ void dag_abort(const DagParams& P) {
  atomicCAS(P.info, 0, -1);
  st_release_gpu(P.flags + F_ABORT, 1);
}

// ---- merged solve + update of tile row k + 1 (DAG kernel, CTA 1) ------------------------------------
// X = A(k+1, k) L_kk^-T by forward substitution over the four 32-column blocks AS CTA 0 PUBLISHES THEM
// (pb / pinv), each warp carrying 8 of the 128 rows; after every block the rank-32 update of the diagonal
// tile (k+1, k+1), whose accumulators stay in registers (16 x 16 macro tiles of the lower triangle, 36 of
// them dealt to the 16 warps, 9 per scheduler).  When CTA 0 finishes block column 3 only the last
// quarter of the solve and of the update is left: the chain per block column is potrf + ~1/4 (solve +
// update) instead of potrf + solve + update.
// Shared memory: the X strips (16 warps x 128 columns x 8 rows) and L_kk as four trapezoidal block columns
// (rows 32 b .. 127 of columns 32 b .. 32 b + 31, the diagonal block replaced by its inverse).
__device__ __forceinline__ int tu_lb_off(int b) { return b == 0 ? 0 : b == 1 ? 32 * 132 : b == 2 ? 32 * (132 + 100) : 32 * (132 + 100 + 68); }
__device__ __forceinline__ int tu_lb_ld(int b) { return 132 - 32 * b; }      // = 4 (mod 16): conflict-free B fragments

// CTA-wide wait on one or two counters (thread 0 polls with relaxed loads; the successful observation is
// repeated as an acquire load, which is much cheaper than a membar).  Returns the first counter's value, or
// -1 after an abort.
__device__ __forceinline__ int tu_wait(const DagParams& P, const int* f0, int v0, const int* f1, int v1, bool ge, int* s_ok, long long* t_wait) {
  if (threadIdx.x == 0) {
    const long long c0 = clock64();
    long long spins = 0;
    int ok = 0;
    for (;;) {
      const int a = ld_relaxed(f0), b = f1 ? ld_relaxed(f1) : v1;
      if (ge ? (a >= v0 && b >= v1) : (a == v0 && b == v1)) {
        ok = ld_acquire_gpu(f0);
        if (f1) ok = min(ok, ld_acquire_gpu(f1));
        break;
      }
      if ((++spins & 255) == 0 && ld_relaxed(P.flags + F_ABORT)) { ok = -1; break; }
      if (spins > SPIN_LIMIT) { dag_abort(P); ok = -1; break; }
    }
    *s_ok = ok;
    *t_wait += clock64() - c0;
  }
  __syncthreads();
  const int ok = *s_ok;
  __syncthreads();
  return ok;
}

__device__ __forceinline__ bool tu_dev(const DagParams& P, int k, double* sm, int* s_ok, long long* prof) {
  const int T = P.T, ld = P.ld, n_rows = P.n_rows;
  double* __restrict__ S = P.S;
  int* pdone = P.flags + F_PDONE;
  int* ver = pdone + T;
  const int* pb = ver + P.R64 * T;
  const int* pinv = pb + T;
  const int i = k + 1, r0 = i * NB, k0 = k * NB, nb = min(NB, P.n - k0);
  const bool has_u1 = i < T;
  const bool h2 = half_exists(P, 2 * i + 1);
  const double* Linv = P.Linv + (size_t)k * NB * NB;
  double* Xs = sm;                        // Xs[(strip * NB + col) * 8 + row_in_strip]
  double* Lb = sm + 16 * NB * 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  TRACE(P, k, 2);

  // macro tiles of the update: t = mi (mi + 1) / 2 + mj, t = warp, warp + 16, warp + 32
  int mt_i[3], mt_j[3];
  bool mt_ok[3];
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const int t = warp + 16 * s;
    int mi = 0;
    while ((mi + 1) * (mi + 2) / 2 <= t) ++mi;
    mt_i[s] = mi; mt_j[s] = t - mi * (mi + 1) / 2;
    mt_ok[s] = has_u1 && t < 36;
  }
  double acc[3][2][2][2];
  if (has_u1) {
    // the diagonal tile (k+1, k+1) received update k - 1 from the bulk queue long ago: its loads fly while
    // the solve still waits for its own inputs
    if (tu_wait(P, ver + (2 * i) * T + i, k, h2 ? ver + (2 * i + 1) * T + i : nullptr, k, false, s_ok, prof) < 0) return false;
    TRACE(P, k, 4);
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int r = min(r0 + 16 * mt_i[s] + 8 * a + g, n_rows - 1), c = min(r0 + 16 * mt_j[s] + 8 * b + 2 * q + e, P.n - 1);
            acc[s][a][b][e] = mt_ok[s] ? __ldcg(S + (size_t)c * ld + r) : 0.0;
          }
  }

  // inputs of the solve: tile (k+1, k) has received update k - 1
  if (tu_wait(P, ver + (2 * i) * T + k, k, h2 ? ver + (2 * i + 1) * T + k : nullptr, k, false, s_ok, prof) < 0) return false;
  TRACE(P, k, 3);
  const long long c_begin = clock64();
  for (int c = warp; c < NB; c += 16) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int r2 = 2 * lane + 64 * s, row = r0 + r2;
      double* dst = Xs + ((r2 >> 3) * NB + c) * 8 + (r2 & 7);
      const double* src = S + (size_t)(k0 + c) * ld + row;
      if (c < nb && row + 1 < n_rows) {
        cp_async16(dst, src, true);
      } else {
        dst[0] = (c < nb && row < n_rows) ? __ldcg(src) : 0.0;
        dst[1] = 0.0;
      }
    }
  }
  cp_async_commit();

  double* Xw = Xs + warp * NB * 8;
  const int row = r0 + warp * 8 + g;
  int loaded = 0;                  // block columns of L_kk already requested
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (loaded <= j) {
      // block column j of L_kk and its inverse diagonal block are in memory — usually several are by now:
      // request all of them at once
      const int avail = tu_wait(P, pb + k, j + 1, pinv + k, j + 1, true, s_ok, prof);
      if (avail < 0) return false;
      for (int jb = loaded; jb < min(avail, 4); ++jb) {
        double* Lj = Lb + tu_lb_off(jb);
        const int ldj = tu_lb_ld(jb);
        for (int cl = warp; cl < 32; cl += 16) {
          const int c = 32 * jb + cl;
          double* dl = Lj + cl * ldj - 32 * jb;           // dl[r] = L(r, c), r >= 32 jb
          const bool col_ok = c < nb;
          for (int r2 = 32 * jb + 2 * lane; r2 < NB; r2 += 64) {
            const bool in_diag = r2 < 32 * jb + 32;
            const double* src = in_diag ? Linv + (size_t)c * NB + r2 : S + (size_t)(k0 + c) * ld + k0 + r2;
            if (col_ok && r2 >= c && r2 + 1 < nb) {
              cp_async16(dl + r2, src, true);
            } else {
              dl[r2] = (col_ok && r2 >= c && r2 < nb) ? __ldcg(src) : 0.0;
              dl[r2 + 1] = (col_ok && r2 + 1 >= c && r2 + 1 < nb) ? __ldcg(src + 1) : 0.0;
            }
          }
        }
      }
      loaded = min(avail, 4);
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
    }
    TRACE(P, k, 5 + j);
    {
      double a4[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) a4[nt][e] = Xw[(32 * j + 8 * nt + 2 * q + e) * 8 + g];
#pragma unroll
      for (int p = 0; p < j; ++p) {
        const double* Lp = Lb + tu_lb_off(p) + (32 * j - 32 * p) + g;
        const int ldp = tu_lb_ld(p);
#pragma unroll
        for (int kk = 0; kk < 32; kk += 4) {
          const double a = -Xw[(32 * p + kk + q) * 8 + g];
          const double* bs = Lp + (kk + q) * ldp;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) dmma(a4[nt][0], a4[nt][1], a, bs[nt * 8]);
        }
      }
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) Xw[(32 * j + 8 * nt + 2 * q + e) * 8 + g] = a4[nt][e];
      __syncwarp();
      double out[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
      const double* Ij = Lb + tu_lb_off(j) + g;
      const int ldj = tu_lb_ld(j);
#pragma unroll
      for (int kk = 0; kk < 32; kk += 4) {
        const double a = Xw[(32 * j + kk + q) * 8 + g];
        const double* bs = Ij + (kk + q) * ldj;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma(out[nt][0], out[nt][1], a, bs[nt * 8]);
      }
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = 32 * j + 8 * nt + 2 * q + e;
          Xw[c * 8 + g] = out[nt][e];
          if (row < n_rows && c < nb) S[(size_t)(k0 + c) * ld + row] = out[nt][e];
        }
    }
    if (j == 3) __threadfence();
    __syncthreads();
    if (j == 3 && tid == 0) {      // the tile row is solved: the updates of block column k + 1 may start
      st_release_gpu(ver + (2 * i) * T + k, k + 1);
      if (h2) st_release_gpu(ver + (2 * i + 1) * T + k, k + 1);
      TRACE(P, k, 9);
    }
    if (has_u1) {
      // rank-32 update of the diagonal tile, the (up to) three macro tiles of the warp interleaved: twelve
      // independent accumulator pairs per k-step
      const double* xa[3];
      const double* xb[3];
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        xa[s] = Xs + ((2 * mt_i[s]) * NB + 32 * j + q) * 8 + g;
        xb[s] = Xs + ((2 * mt_j[s]) * NB + 32 * j + q) * 8 + g;
      }
#pragma unroll
      for (int kk = 0; kk < 32; kk += 4) {
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          if (!mt_ok[s]) continue;
          const double a0 = -xa[s][kk * 8], a1 = -xa[s][NB * 8 + kk * 8];
          const double b0 = xb[s][kk * 8], b1 = xb[s][NB * 8 + kk * 8];
          dmma(acc[s][0][0][0], acc[s][0][0][1], a0, b0);
          dmma(acc[s][0][1][0], acc[s][0][1][1], a0, b1);
          dmma(acc[s][1][0][0], acc[s][1][0][1], a1, b0);
          dmma(acc[s][1][1][0], acc[s][1][1][1], a1, b1);
        }
      }
    }
  }
  if (has_u1) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      if (!mt_ok[s]) continue;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int r = r0 + 16 * mt_i[s] + 8 * a + g, c = r0 + 16 * mt_j[s] + 8 * b + 2 * q + e;
            if (r < n_rows && c < P.n) S[(size_t)c * ld + r] = acc[s][a][b][e];
          }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      st_release_gpu(ver + (2 * i) * T + i, k + 1);
      if (h2) st_release_gpu(ver + (2 * i + 1) * T + i, k + 1);
    }
  }
  __syncthreads();
  TRACE(P, k, 10);
  if (tid == 0) { prof[2] += clock64() - c_begin; prof[5] += 1; }
  return true;
}

// Worker scheduling (thread 0).  Tasks are claimed by TICKET (one atomicAdd on the queue head, many in flight)
// and their inputs are polled afterwards; a worker holds at most one ticket per queue and runs whichever of its
// two tasks becomes ready first.  Claiming only tasks whose inputs are final — a compare-and-swap on the head
// after a readiness check — serialises every claim behind ~1.3 us of dependent L2 round trips: 19 000 tasks,
// 25 ms (measured).  High-priority tickets are only taken inside the steps whose block column is factored
// (hp_end[npan - 1]), so a worker is never parked on a panel solve of a later step.
// Deadlock freedom: both queue orders extend to one topological order of the task graph (steps only depend on
// earlier steps); the first unfinished task in that order has all inputs final and is either on a dedicated
// CTA, or held by a worker (which polls it), or at the head of its queue with no ticket of that queue
// outstanding — and workers never block while holding a ticket.
struct WorkerState {
  int4 t_hp, t_ur, t_lp;
  bool have_hp = false, have_ur = false, have_lp = false, hp_done = false, ur_done = false, lp_done = false;
  bool new_hp = false, new_ur = false, new_lp = false;     // ticket taken since the last prefetch round
  int patience = 0;        // polls during which only the held high-priority ticket is considered
};

// All threads: pull the tile a held ticket will read-modify-write (C tile of an update, panel rows of a
// solve) into L2.  S (287 MB at n = 5988) does not stay in the 126 MB L2 between two updates of a tile; without
// this every task starts with ~11 k cycles of exposed HBM latency (measured).
__device__ __forceinline__ void prefetch_task(const DagParams& P, int4 t) {
  const int tid = threadIdx.x;
  int r0, c0, nr;
  if (t.x == TASK_UPD) { r0 = t.y * NB; c0 = t.z * NB; nr = NB; }
  else if (t.x == TASK_TRSM) { r0 = t.y * 64; c0 = (t.w & 0xffff) * NB; nr = 64; }
  else if (t.x == TASK_UPD64) { r0 = t.y * 64; c0 = t.z * NB; nr = 64; }
  else return;
  const int lines_per_col = nr / 16;                 // 128-byte lines
  for (int e = tid; e < NB * lines_per_col; e += DAG_THREADS) {
    const int c = c0 + e / lines_per_col, r = r0 + (e % lines_per_col) * 16;
    if (c < P.n && r < P.n_rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.S + (size_t)c * P.ld + r));
  }
}

// A ticket of the high-priority / urgent queue is only taken inside the steps whose block column is factored.
__device__ __forceinline__ void take_gated(const DagParams& P, int flag, const int4* q, int n_q, const int* q_end, int4& t, bool& have,
                                           bool& done, bool& fresh) {
  if (have || done) return;
  const int npan = ld_relaxed(P.flags + F_NPAN), head = ld_relaxed(P.flags + flag);
  const int limit = npan > 0 ? q_end[npan - 1] : 0;
  if (head >= n_q) {
    done = true;
  } else if (head < limit) {
    const int h = atomicAdd(P.flags + flag, 1);
    if (h < n_q) { t = q[h]; have = true; fresh = true; } else done = true;
  }
}
__device__ __forceinline__ void take_far(const DagParams& P, WorkerState& w) {
  if (w.have_lp || w.lp_done) return;
  const int l = atomicAdd(P.flags + F_LP, 1);
  if (l < P.n_lp) { w.t_lp = P.lp[l]; w.have_lp = true; w.new_lp = true; } else w.lp_done = true;
}

__device__ __forceinline__ int4 next_task(const DagParams& P, WorkerState& w) {
  long long spins = 0;
  for (;;) {
    if ((spins & 63) == 0 && ld_relaxed(P.flags + F_ABORT)) return make_int4(TASK_EXIT, 0, 0, 0);
    take_gated(P, F_HP, P.hp, P.n_hp, P.hp_end, w.t_hp, w.have_hp, w.hp_done, w.new_hp);
    if (w.have_hp && task_ready(P, w.t_hp)) { w.have_hp = false; w.new_hp = false; w.patience = 0; __threadfence(); return w.t_hp; }
    if (w.have_hp && w.patience > 0) {      // the update that follows this worker's panel solve: its last input (the solve of
      --w.patience;                         // tile row k + 1 on CTA 1) is microseconds away — do not start a long task now
      ++spins;
      __nanosleep(100);
      continue;
    }
    take_gated(P, F_UR, P.ur, P.n_ur, P.ur_end, w.t_ur, w.have_ur, w.ur_done, w.new_ur);
    if (w.have_ur && task_ready(P, w.t_ur)) { w.have_ur = false; w.new_ur = false; __threadfence(); return w.t_ur; }
    take_far(P, w);
    if (w.have_lp && task_ready(P, w.t_lp)) { w.have_lp = false; w.new_lp = false; __threadfence(); return w.t_lp; }
    if (w.hp_done && w.ur_done && w.lp_done && !w.have_hp && !w.have_ur && !w.have_lp) return make_int4(TASK_EXIT, 0, 0, 0);
    if (++spins > SPIN_LIMIT) { dag_abort(P); return make_int4(TASK_EXIT, 0, 0, 0); }
    __nanosleep(spins < 8 ? 100 : 400);
  }
}

__global__ void __launch_bounds__(DAG_THREADS, 1) k_chol_dag(const DagParams P) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int4 s_task;
  __shared__ int s_ok;
  __shared__ int s_sig;
  __shared__ int s_prog;
  const int tid = threadIdx.x, b = blockIdx.x;
  const int T = P.T;
  int* pdone = P.flags + F_PDONE;
  __shared__ unsigned long long s_cbar;     // completion barrier of the C-tile bulk copies
  unsigned cphase = 0;                      // uses of s_cbar so far (CTA-uniform)
  if (threadIdx.x == 0) mbar_init(&s_cbar, 1);
  __shared__ long long s_prof[16];      // thread 0 only: [0] wait [1] potrf [2] trsm [3] upd [4] inv [5] tasks [6] start
  if (tid == 0) {
    for (int i = 0; i < 16; ++i) s_prof[i] = 0;
    s_prof[6] = clock64();
  }
#define t_wait s_prof[0]
#define t_potrf s_prof[1]
#define t_trsm s_prof[2]
#define t_upd s_prof[3]
#define t_inv s_prof[4]
#define n_tasks s_prof[5]

  auto run = [&](int4 t) {     // all threads
    const long long c0 = clock64();
    const int k = t.w & 0xffff, nk = t.w >> 16, k0 = k * NB, nb = min(NB, P.n - k0);
    if (t.x == TASK_TRSM) trsm64_dev(P.S, P.ld, P.n_rows, k0, nb, t.y * 64, P.Linv + (size_t)k * NB * NB, sm, DAG_THREADS / 32);
    else if (t.x == TASK_UPD) upd_dev<4>(P.S, P.ld, P.n_rows, P.n, k0, NB * nk, t.y * NB, t.z * NB, sm, s_prof + 8, &s_cbar, cphase);
    else if (t.x == TASK_UPD64) upd_dev<2>(P.S, P.ld, P.n_rows, P.n, k0, NB, t.y * 64, t.z * NB, sm, s_prof + 12, &s_cbar, cphase);
    else if (t.x == TASK_INV) inv_offdiag_dev(P.S, P.ld, k0, nb, P.Linv + (size_t)k * NB * NB, sm);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      task_publish(P, t);
      const long long dt = clock64() - c0;
      if (t.x == TASK_TRSM) t_trsm += dt; else if (t.x == TASK_UPD || t.x == TASK_UPD64) t_upd += dt; else t_inv += dt;
      ++n_tasks;
    }
  };
  if (b == 0) {
    // ---- the factorisation chain: diagonal blocks ----
    for (int k = 0; k < T; ++k) {
      const int k0 = k * NB, nb = min(NB, P.n - k0);
      if (tid == 0) {
        const long long c0 = clock64();
        const int* ver = pdone + T;
        long long spins = 0;
        int ok = 1;
        for (;;) {
          const int v0 = ld_relaxed(ver + (2 * k) * T + k), v1 = half_exists(P, 2 * k + 1) ? ld_relaxed(ver + (2 * k + 1) * T + k) : k;
          if (v0 == k && v1 == k) break;
          if ((++spins & 255) == 0 && ld_relaxed(P.flags + F_ABORT)) { ok = 0; break; }
          if (spins > SPIN_LIMIT) { dag_abort(P); ok = 0; break; }
        }
        __threadfence();
        s_ok = ok;
        t_wait += clock64() - c0;
      }
      __syncthreads();
      const bool ok = s_ok != 0;
      __syncthreads();
      if (!ok) break;
      TRACE(P, k, 0);
      const long long c0 = clock64();
      {
        int* pb = P.flags + F_PDONE + T + P.R64 * T;
        potrf128_prog_dev(P.S, P.ld, k0, nb, P.Linv + (size_t)k * NB * NB, P.info, sm, pb + k, pb + T + k, &s_sig, &s_prog);
      }
      __threadfence();
      __syncthreads();
      TRACE(P, k, 1);
      if (tid == 0) { st_release_gpu(pdone + k, 1); st_release_gpu(P.flags + F_NPAN, k + 1); t_potrf += clock64() - c0; ++n_tasks; }
      // the right-hand-side row n lives inside the last diagonal tile when n is not a multiple of 128
      if (k == T - 1 && P.n_rows > P.n && P.n < T * NB) {
        trsm64_dev(P.S, P.ld, P.n_rows, k0, nb, P.n, P.Linv + (size_t)k * NB * NB, sm, DAG_THREADS / 32);
        __syncthreads();
      }
    }
  } else if (b == 1) {
    // ---- tile row k + 1: solve of tile (k+1, k) and update of tile (k+1, k+1), pipelined with CTA 0 ----
    for (int k = 0; k < T; ++k) {
      if (!half_exists(P, 2 * (k + 1))) break;
      if (!tu_dev(P, k, sm, &s_ok, s_prof)) break;
    }
  } else {
    WorkerState w;
    __shared__ int4 s_pref[3];
    for (;;) {
      if (tid == 0) {
        const long long c0 = clock64();
        const int4 t = next_task(P, w);
        s_task = t;
        // refill the tickets right away: their tiles are prefetched into L2 while this task runs
        if (t.x != TASK_EXIT) {
          take_gated(P, F_UR, P.ur, P.n_ur, P.ur_end, w.t_ur, w.have_ur, w.ur_done, w.new_ur);
          take_far(P, w);
        }
        s_pref[0] = w.new_hp ? w.t_hp : make_int4(TASK_EXIT, 0, 0, 0);
        s_pref[1] = w.new_lp ? w.t_lp : make_int4(TASK_EXIT, 0, 0, 0);
        s_pref[2] = w.new_ur ? w.t_ur : make_int4(TASK_EXIT, 0, 0, 0);
        w.new_hp = w.new_lp = w.new_ur = false;
        t_wait += clock64() - c0;
      }
      __syncthreads();
      const int4 t = s_task, p0 = s_pref[0], p1 = s_pref[1], p2 = s_pref[2];
      __syncthreads();
      if (t.x == TASK_EXIT) break;
      prefetch_task(P, p0);
      prefetch_task(P, p1);
      prefetch_task(P, p2);
      run(t);
      // A panel solve is followed on the same worker by the update of its rows in block column k + 1 (held
      // like a claimed ticket: it never blocks the worker).  The pair is what the chain waits for.
      if (tid == 0 && t.x == TASK_TRSM && (t.w & 0xffff) + 1 < T && !w.have_hp) {
        w.t_hp = make_int4(TASK_UPD64, t.y, (t.w & 0xffff) + 1, t.w);
        w.have_hp = true;
        w.new_hp = false;
        w.patience = 64;
      }
    }
  }
  if (P.prof && tid == 0) {
    long long* o = P.prof + (size_t)b * 16;
    for (int i = 0; i < 4; ++i) o[8 + i] = s_prof[8 + i];
    o[0] = t_wait; o[1] = t_potrf; o[2] = t_trsm; o[3] = t_upd; o[4] = t_inv; o[5] = n_tasks; o[6] = clock64() - s_prof[6];
  }
#undef t_wait
#undef t_potrf
#undef t_trsm
#undef t_upd
#undef t_inv
#undef n_tasks
}

// =================================================================================================
// DAG kernel, second scheduler (default): the same tile operations, but
//   * no queues: every worker SCANS a table of tile states (st[i][j] = block-column updates applied to tile
//     (i, j), sv[h] = block columns for which the 64-row half h holds final L) and claims, with one
//     compare-and-swap, the ready task NEAREST THE FACTORISATION FRONT: panel solves first, then updates of
//     block column npan, npan + 1, ...  An update takes every panel that is available for its tile (up to G),
//     so tiles near the front are served one panel at a time with minimal latency while far tiles collect
//     their updates lazily with a long K (one C-tile round trip per G panels).  The FIFO queues of the first
//     scheduler tied the front to the bulk: potrf(k) could never run more than W + 2 panels ahead of the
//     far queue, so the chain-bound and the throughput-bound phases added up instead of overlapping
//     (profiles/r2_dense_notes.md).
//   * five dedicated CTAs carry the chain:  0 = potrf(k);  1 = solve of tile row k+1;  2 = update of the
//     diagonal tile (k+1, k+1) with panel k;  3 = solve of tile row k+2;  4 = update of tile (k+2, k+1) with
//     panel k.  CTAs 1 and 3 follow CTA 0 block column by block column (pb / pinv), CTAs 2 and 4 follow the
//     solves the same way (xpub): when potrf(k) ends, a quarter of a solve and a rank-32 update are left.
// Each (tile, panel) pair and each (half, panel) solve is executed exactly once; the work units finished
// are counted in F2_DONE and the workers leave when the count reaches the total.
// =================================================================================================
enum { F2_ABORT = 0, F2_NPAN = 32, F2_DONE = 64, F2_JMIN = 96, F2_ARR = 128 };     // F2_JMIN: every column before it is complete (hint, monotone)
constexpr int ST_LOCK = 1 << 20;
enum { TASK2_EXIT = -1, TASK2_TRSM = 0, TASK2_UPD = 1, TASK2_INV = 2 };

struct Dag2Params {
  double* S;
  int ld, n, n_rows, T, Tr, R64;
  double* Linv;
  int* info;
  int* flags;               // [F2_ABORT] [F2_NPAN] [F2_DONE] [F2_JMIN] . pdone[T] pb[T] pinv[T] invst[T] xpub[2 Tr] sv[R64] st[Tr * T] hd[Tr * T]
  int total_units;
  int W, G;
  int D;                    // tile rows k + 1 .. k + D are carried by dedicated CTAs at step k
  const int* tiles;         // every tile (i | j << 16), column by column
  const int* col_start;     // col_start[j]: position of tile (j, j) in tiles[]
  int n_tiles;
  unsigned long long* trace;
  long long* prof;
};
__device__ __forceinline__ int* f2_pdone(const Dag2Params& P) { return P.flags + F2_ARR; }
__device__ __forceinline__ int* f2_pb(const Dag2Params& P) { return f2_pdone(P) + P.T; }
__device__ __forceinline__ int* f2_pinv(const Dag2Params& P) { return f2_pb(P) + P.T; }
__device__ __forceinline__ int* f2_invst(const Dag2Params& P) { return f2_pinv(P) + P.T; }
__device__ __forceinline__ int* f2_xpub(const Dag2Params& P) { return f2_invst(P) + P.T; }
__device__ __forceinline__ int* f2_sv(const Dag2Params& P) { return f2_xpub(P) + 2 * P.Tr; }
__device__ __forceinline__ int* f2_st(const Dag2Params& P) { return f2_sv(P) + P.R64; }
__device__ __forceinline__ int* f2_hd(const Dag2Params& P) { return f2_st(P) + (size_t)P.Tr * P.T; }
__device__ __forceinline__ bool half2_exists(const Dag2Params& P, int h) { return h * 64 < P.n_rows; }
// Panels the WORKERS owe tile (i, j): panels k >= i - D belong to the dedicated CTAs; with D = 1 the solve CTAs of tile
// row j + 1 also apply the last update (panel j - 1) of their own tile (j + 1, j) themselves (solve_row_dev).
__device__ __forceinline__ int lim2(const Dag2Params& P, int i, int j) {
  const int l = min(j, i - P.D);
  return (P.D == 1 && i - j == 1) ? l - 1 : l;
}

__device__ __forceinline__ void dag2_abort(const Dag2Params& P) {
  atomicCAS(P.info, 0, -1);
  st_release_gpu(P.flags + F2_ABORT, 1);
}

// CTA-wide wait until *f0 >= v0 (and *f1 >= v1).  Thread 0 polls with relaxed loads and repeats the successful
// observation as an acquire load.  Returns false after an abort.
__device__ __forceinline__ bool wait2(const Dag2Params& P, const int* f0, int v0, const int* f1, int v1, int* s_ok, long long* t_wait) {
  if (threadIdx.x == 0) {
    const long long c0 = clock64();
    long long spins = 0;
    int ok = 1;
    for (;;) {
      const int a = ld_relaxed(f0) & ~ST_LOCK, b = f1 ? (ld_relaxed(f1) & ~ST_LOCK) : v1;
      if (a >= v0 && b >= v1) {
        ld_acquire_gpu(f0);
        if (f1) ld_acquire_gpu(f1);
        break;
      }
      if ((++spins & 255) == 0 && ld_relaxed(P.flags + F2_ABORT)) { ok = 0; break; }
      if (spins > SPIN_LIMIT) { dag2_abort(P); ok = 0; break; }
    }
    *s_ok = ok;
    *t_wait += clock64() - c0;
  }
  __syncthreads();
  const int ok = *s_ok;
  __syncthreads();
  return ok != 0;
}

// CTA-wide wait until min(*f0, *f1) >= v; returns that minimum as thread 0 saw it (acquire), -1 after an abort.
__device__ __forceinline__ int wait2v(const Dag2Params& P, const int* f0, const int* f1, int v, int* s_ok, long long* t_wait) {
  if (threadIdx.x == 0) {
    const long long c0 = clock64();
    long long spins = 0;
    int ok = -1;
    for (;;) {
      if (min(ld_relaxed(f0), ld_relaxed(f1)) >= v) {
        ok = min(ld_acquire_gpu(f0), ld_acquire_gpu(f1));
        break;
      }
      if ((++spins & 255) == 0 && ld_relaxed(P.flags + F2_ABORT)) break;
      if (spins > SPIN_LIMIT) { dag2_abort(P); break; }
    }
    *s_ok = ok;
    *t_wait += clock64() - c0;
  }
  __syncthreads();
  const int ok = *s_ok;
  __syncthreads();
  return ok;
}

// CTA-wide wait until *f0 >= v, *f1 >= v and *f2 >= v (null pointers are skipped).  Returns false after an abort.
__device__ __forceinline__ bool wait3(const Dag2Params& P, const int* f0, const int* f1, const int* f2, int v, int* s_ok, long long* t_wait) {
  if (threadIdx.x == 0) {
    const long long c0 = clock64();
    long long spins = 0;
    int ok = 1;
    for (;;) {
      if ((!f0 || (ld_relaxed(f0) & ~ST_LOCK) >= v) && (!f1 || (ld_relaxed(f1) & ~ST_LOCK) >= v) && (!f2 || (ld_relaxed(f2) & ~ST_LOCK) >= v)) {      // (sv counters carry a lock bit while a panel solve is running)
        if (f0) ld_acquire_gpu(f0);
        if (f1) ld_acquire_gpu(f1);
        if (f2) ld_acquire_gpu(f2);
        break;
      }
      if ((++spins & 255) == 0 && ld_relaxed(P.flags + F2_ABORT)) { ok = 0; break; }
      if (spins > SPIN_LIMIT) { dag2_abort(P); ok = 0; break; }
    }
    *s_ok = ok;
    *t_wait += clock64() - c0;
  }
  __syncthreads();
  const int ok = *s_ok;
  __syncthreads();
  return ok != 0;
}

// ---- solve of one tile row against panel k, following CTA 0 block column by block column ----------------
// X = A(i, k) L_kk^-T by forward substitution over the four 32-column blocks as they are published (pb / pinv),
// each warp carrying 8 of the 128 rows (warp-private strips: no CTA-wide synchronisation inside a step).  Every
// finished block column is published (xpub[i] = 4 k + b + 1) for the update CTAs; at the end the two halves of
// the tile row are marked solved for block column k.
__device__ __forceinline__ bool solve_row_dev(const Dag2Params& P, int i, int half, int k, double* sm, int* s_ok, long long* prof, int trace_base) {
  const int T = P.T, ld = P.ld, n_rows = P.n_rows;
  double* __restrict__ S = P.S;
  if (!half2_exists(P, 2 * i + half)) return true;
  const int r0 = i * NB + 64 * half, k0 = k * NB, nb = min(NB, P.n - k0);
  const double* Linv = P.Linv + (size_t)k * NB * NB;
  double* Xs = sm;                        // Xs[(strip * NB + col) * 8 + row_in_strip]
  double* Lb = sm + 16 * NB * 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  int* sv = f2_sv(P);
  if (trace_base >= 0) TRACE(P, k, trace_base);
  // With D = 1 the LAST update of tile (k + 1, k) (panel k - 1) is applied right here instead of by a worker: on the
  // chain's cycle it replaced  [worker scan -> 128 x 128 x 128 update, 24 us -> flag -> this CTA]  by 64 rows of that
  // update on this SM, started the moment L(k + 1, k - 1) exists.
  const bool own_last = P.D == 1 && i == k + 1 && k >= 1;
  // tile (i, k) has received every update the workers owe it
  if (!wait2(P, f2_st(P) + (size_t)i * T + k, own_last ? k - 1 : k, nullptr, 0, s_ok, prof)) return false;
  if (trace_base >= 0) TRACE(P, k, trace_base + 1);
  const long long c_begin = clock64();
  for (int c = warp; c < NB; c += 16) {
    const int r2 = 2 * lane, row = r0 + r2;
    double* dst = Xs + ((r2 >> 3) * NB + c) * 8 + (r2 & 7);
    const double* src = S + (size_t)(k0 + c) * ld + row;
    if (c < nb && row + 1 < n_rows) {
      cp_async16(dst, src, true);
    } else {
      dst[0] = (c < nb && row < n_rows) ? __ldcg(src) : 0.0;
      dst[1] = 0.0;
    }
  }
  cp_async_commit();
  if (own_last) {
    // X -= L(i, k-1)[my 64 rows] L(k, k-1)^T, K = 128 in four chunks of 32 (double-buffered cp.async stages in the part
    // of shared memory the solve only uses later); all 16 warps: strip = warp & 7, column half = warp >> 3.
    constexpr int ALD = 64 + 4, BLD = NB + 4, STG = 32 * (ALD + BLD);
    double* stg = sm + 8 * NB * 8;
    const int kp0 = (k - 1) * NB, rk0 = k * NB;
    // L(i, k-1) of these rows (panel solve by a worker) and L(k, k-1) (both halves, the previous step of the solve CTAs)
    if (!wait3(P, sv + 2 * i + half, sv + 2 * k, half2_exists(P, 2 * k + 1) ? sv + 2 * k + 1 : nullptr, k, s_ok, prof)) return false;
    auto load_chunk = [&](int c4, int stage) {
      double* As = stg + stage * STG;
      double* Bs = As + 32 * ALD;
#pragma unroll
      for (int p = 0; p < 32 * 64 / DAG_THREADS; ++p) {
        const int piece = tid + p * DAG_THREADS;
        const int kk = piece >> 6, r2 = (piece & 63) * 2;
        const double* colp = S + (size_t)(kp0 + 32 * c4 + kk) * ld;
        const int ra = r0 + r2, rb = rk0 + r2;
        if (r2 < 64) cp_async16(As + kk * ALD + r2, colp + (ra < n_rows ? ra : 0), ra < n_rows);
        cp_async16(Bs + kk * BLD + r2, colp + (rb < n_rows ? rb : 0), rb < n_rows);
      }
      cp_async_commit();
    };
    load_chunk(0, 0);
    cp_async_wait<1>();          // the X strips (first group) have landed
    __syncthreads();
    const int strip = warp & 7, ch = warp >> 3;
    double* Xq = Xs + strip * NB * 8;
    double acc[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) acc[nt][e] = Xq[(64 * ch + 8 * nt + 2 * q + e) * 8 + g];
    for (int c4 = 0; c4 < 4; ++c4) {
      if (c4 + 1 < 4) load_chunk(c4 + 1, (c4 + 1) & 1);
      if (c4 + 1 < 4) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncthreads();
      const double* As = stg + (c4 & 1) * STG + strip * 8 + g;
      const double* Bs = stg + (c4 & 1) * STG + 32 * ALD + 64 * ch + g;
#pragma unroll
      for (int kk = 0; kk < 32; kk += 4) {
        const double a = -As[(kk + q) * ALD];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) dmma(acc[nt][0], acc[nt][1], a, Bs[(kk + q) * BLD + 8 * nt]);
      }
      __syncthreads();           // the stage is refilled two chunks later
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) Xq[(64 * ch + 8 * nt + 2 * q + e) * 8 + g] = acc[nt][e];
    __syncthreads();
  }
  const bool cw = warp < 8;             // 64 rows = 8 strips: warps 8..15 only help with the loads and the barriers
  double* Xw = Xs + (warp & 7) * NB * 8;
  const int row = r0 + (warp & 7) * 8 + g;
  int loaded = 0;                  // block columns of L_kk already requested
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (loaded <= j) {
      // (the count must be CTA-uniform: every thread derives its loads and the next wait from it)
      const int avail = wait2v(P, f2_pb(P) + k, f2_pinv(P) + k, j + 1, s_ok, prof);
      if (avail < 0) return false;
      for (int jb = loaded; jb < min(avail, 4); ++jb) {
        double* Lj = Lb + tu_lb_off(jb);
        const int ldj = tu_lb_ld(jb);
        for (int cl = warp; cl < 32; cl += 16) {
          const int c = 32 * jb + cl;
          double* dl = Lj + cl * ldj - 32 * jb;           // dl[r] = L(r, c), r >= 32 jb
          const bool col_ok = c < nb;
          for (int r2 = 32 * jb + 2 * lane; r2 < NB; r2 += 64) {
            const bool in_diag = r2 < 32 * jb + 32;
            const double* src = in_diag ? Linv + (size_t)c * NB + r2 : S + (size_t)(k0 + c) * ld + k0 + r2;
            if (col_ok && r2 >= c && r2 + 1 < nb) {
              cp_async16(dl + r2, src, true);
            } else {
              dl[r2] = (col_ok && r2 >= c && r2 < nb) ? __ldcg(src) : 0.0;
              dl[r2 + 1] = (col_ok && r2 + 1 >= c && r2 + 1 < nb) ? __ldcg(src + 1) : 0.0;
            }
          }
        }
      }
      loaded = min(avail, 4);
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
    }
    if (trace_base >= 0) TRACE(P, k, trace_base + 2 + j);
    if (cw) {
      // X_j = A_j Inv_jj^T: the contributions of the earlier block columns were subtracted as soon as they were
      // known (below), so that only this product separates "block column j published" from "X_j published"
      double out[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
      const double* Ij = Lb + tu_lb_off(j) + g;
      const int ldj = tu_lb_ld(j);
#pragma unroll
      for (int kk = 0; kk < 32; kk += 4) {
        const double a = Xw[(32 * j + kk + q) * 8 + g];
        const double* bs = Ij + (kk + q) * ldj;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma(out[nt][0], out[nt][1], a, bs[nt * 8]);
      }
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = 32 * j + 8 * nt + 2 * q + e;
          Xw[c * 8 + g] = out[nt][e];
          if (row < n_rows && c < nb) S[(size_t)(k0 + c) * ld + row] = out[nt][e];
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      st_release_gpu(f2_xpub(P) + 2 * i + half, 4 * k + j + 1);
      if (j == 3) {      // these 64 rows are solved for block column k
        st_release_gpu(sv + 2 * i + half, k + 1);
        if (trace_base >= 0) TRACE(P, k, trace_base + 6);
      }
    }
    // right-looking inside the tile row: A_jj -= X_j L_jj,j^T for the later block columns jj (block column j of
    // L_kk is in shared memory) — off the chain: CTA 0 is factoring block column j + 1 meanwhile
    __syncwarp();
#pragma unroll
    for (int jj = j + 1; cw && jj < 4; ++jj) {
      double a4[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) a4[nt][e] = Xw[(32 * jj + 8 * nt + 2 * q + e) * 8 + g];
      const double* Lp = Lb + tu_lb_off(j) + (32 * jj - 32 * j) + g;
      const int ldp = tu_lb_ld(j);
#pragma unroll
      for (int kk = 0; kk < 32; kk += 4) {
        const double a = -Xw[(32 * j + kk + q) * 8 + g];
        const double* bs = Lp + (kk + q) * ldp;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma(a4[nt][0], a4[nt][1], a, bs[nt * 8]);
      }
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) Xw[(32 * jj + 8 * nt + 2 * q + e) * 8 + g] = a4[nt][e];
      __syncwarp();
    }
  }
  if (tid == 0) { prof[2] += clock64() - c_begin; prof[5] += 1; }
  return true;
}

// ---- update of one near-diagonal tile with panel k, following the solves block column by block column ----
// C(i, j) -= X_i X_j^T where X_i = L(i, k), X_j = L(j, k) are being produced by the solve CTAs: the
// accumulators hold the C tile while the four rank-32 updates arrive (xpub).  16 warps, 4 x 4, warp tile 32 x 32;
// on a diagonal tile only the warp tiles on or below the diagonal work (same map as upd_dev).
constexpr int US_LD = NB + 4;
constexpr int LDC64 = 64 + 2;
__device__ __forceinline__ bool upd_stream_dev(const Dag2Params& P, int i, int j, int half, int k, double* smem, int* s_ok, long long* prof,
                                               unsigned long long* cbar, unsigned& cphase, int trace_slot) {
  // 64 rows of the tile (half), 128 columns: 16 warps as 4 (rows of 16) x 4 (columns of 32)
  if (!half2_exists(P, 2 * i + half)) return true;
  const int T = P.T, lda = P.ld, n_rows = P.n_rows, n_cols = P.n;
  double* __restrict__ S = P.S;
  const int i0 = i * NB + 64 * half, j0 = j * NB, k0 = k * NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  const bool diag = i == j;
  const int wm = warp >> 2, wn = warp & 3;
  // diagonal tile: warp tiles entirely above the diagonal carry no arithmetic
  const bool active = !(diag && 32 * wn > 64 * half + 16 * wm + 15);
  int* st = f2_st(P) + (size_t)i * T + j;
  if (!wait2(P, st, k, nullptr, 0, s_ok, prof)) return false;
  const long long c_begin = clock64();
  double acc[2][4][2];
  {
    const int ncol = min(NB, n_cols - j0);
    const unsigned col_bytes = (unsigned)min(64, lda - i0) * 8u;
    if (tid == 0) mbar_expect_tx(cbar, col_bytes * (unsigned)ncol);
    __syncthreads();
    if (tid < ncol) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      bulk_g2s(smem + tid * LDC64, S + (size_t)(j0 + tid) * lda + i0, col_bytes, cbar);
    }
    mbar_wait(cbar, cphase & 1);
    ++cphase;
    if (active) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) acc[mt][nt][e] = smem[(wn * 32 + nt * 8 + 2 * q + e) * LDC64 + wm * 16 + mt * 8 + g];
    }
    __syncthreads();
  }
  double* As = smem;                     // [32][US_LD]: 64 rows of tile row i
  double* Bs = smem + 32 * US_LD;        // [32][US_LD]: 128 rows of tile row j
  const int* xa = f2_xpub(P) + 2 * i + half;
  const int* xb0 = f2_xpub(P) + 2 * j;
  const int* xb1 = half2_exists(P, 2 * j + 1) ? f2_xpub(P) + 2 * j + 1 : nullptr;
  for (int b = 0; b < 4; ++b) {
    if (!wait3(P, xa, xb0, xb1, 4 * k + b + 1, s_ok, prof)) return false;
#pragma unroll
    for (int p = 0; p < 32 * 64 / DAG_THREADS; ++p) {
      const int piece = tid + p * DAG_THREADS;
      const int kk = piece >> 6, r2 = (piece & 63) * 2;
      const int kc = 32 * b + kk;
      const bool kin = k0 + kc < n_cols;
      const int ra = i0 + r2, rb = j0 + r2;
      const double* colp = S + (size_t)(kin ? k0 + kc : k0) * lda;
      if (r2 < 64) cp_async16(As + kk * US_LD + r2, colp + (ra < n_rows ? ra : 0), kin && ra < n_rows);
      cp_async16(Bs + kk * US_LD + r2, colp + (rb < n_rows ? rb : 0), kin && rb < n_rows);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (active) {
      const double* as = As + wm * 16 + g;
      const double* bs = Bs + wn * 32 + g;
#pragma unroll
      for (int kk = 0; kk < 32; kk += 4) {
        double a[2], bb[4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) a[mt] = -as[(kk + q) * US_LD + mt * 8];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) bb[nt] = bs[(kk + q) * US_LD + nt * 8];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], bb[nt]);
      }
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int r = i0 + wm * 16 + mt * 8 + g;
      if (r >= n_rows) continue;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = j0 + wn * 32 + nt * 8 + 2 * q + e;
          if (c < n_cols) S[(size_t)c * lda + r] = acc[mt][nt][e];
        }
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    // the second half to arrive publishes the tile (its fence orders the first half's stores, observed through the counter)
    const int need = half2_exists(P, 2 * i + 1) ? 2 : 1;
    const int old = atomicAdd(f2_hd(P) + (size_t)i * T + j, 1);
    if ((old + 1) % need == 0) {
      __threadfence();
      st_release_gpu(st, k + 1);
      if (trace_slot >= 0) TRACE(P, k, trace_slot);
    }
    prof[3] += clock64() - c_begin;
    prof[5] += 1;
  }
  __syncthreads();
  return true;
}

// ---- worker scheduling: warp 0 scans the state tables and claims the ready task nearest the front ----------
// Returns (type, h | i, j, a | nk << 16): TRSM of half h against panel a; UPD of tile (i, j) with panels a .. a + nk - 1.
// tiles[] lists every tile (i | j << 16) column by column (the scan order = distance from the front), col_start[j]
// the position of tile (j, j).  Four candidates per lane and round: the state loads of 128 tiles fly together.
constexpr int MAX_TR = 256;
__device__ __forceinline__ int4 find_task2(const Dag2Params& P, int wid, long long* t_wait, int* s_rowsv) {
  const int lane = threadIdx.x & 31;
  const int T = P.T, Tr = P.Tr;
  int* sv = f2_sv(P);
  int* st = f2_st(P);
  int* invst = f2_invst(P);
  const long long c0 = clock64();
  long long spins = 0;
  int4 res = make_int4(TASK2_EXIT, 0, 0, 0);
  for (;;) {
    if (ld_relaxed(P.flags + F2_ABORT)) break;
    const int np = ld_relaxed(P.flags + F2_NPAN);
    bool claimed = false;
    // ---- solved counts of every tile row; panel solves (tile rows k + 1 and k + 2 belong to CTAs 1 and 3) ----
    for (int i0 = 0; i0 < Tr; i0 += 32) {
      const int i = i0 + lane;
      int ka = 0, kb = 0;
      bool ok0 = false, ok1 = false;
      if (i < Tr) {
        ka = ld_relaxed(sv + 2 * i);
        const bool h1 = half2_exists(P, 2 * i + 1);
        kb = h1 ? ld_relaxed(sv + 2 * i + 1) : (1 << 19);
        s_rowsv[i] = min(ka & ~ST_LOCK, kb & ~ST_LOCK);
        const bool c0k = !(ka & ST_LOCK) && ka < np && ka + P.D < i;
        const bool c1k = h1 && !(kb & ST_LOCK) && kb < np && kb + P.D < i;
        const int s0 = c0k ? ld_relaxed(st + (size_t)i * T + ka) : -1;
        const int s1 = c1k ? ld_relaxed(st + (size_t)i * T + kb) : -1;
        ok0 = c0k && s0 == ka;
        ok1 = c1k && s1 == kb;
      }
      if (claimed) continue;
      const unsigned m = __ballot_sync(FULL, ok0 || ok1);
      if (m) {
        const int cnt = __popc(m);
        const int pick = __fns(m, 0, 1 + (wid % cnt));
        int got = 0, hh = 0, kk = 0;
        if (lane == pick) {
          if (ok0 && atomicCAS(sv + 2 * i, ka, ka | ST_LOCK) == ka) { got = 1; hh = 2 * i; kk = ka; }
          else if (ok1 && atomicCAS(sv + 2 * i + 1, kb, kb | ST_LOCK) == kb) { got = 1; hh = 2 * i + 1; kk = kb; }
        }
        got = __shfl_sync(FULL, got, pick);
        if (got) {
          res = make_int4(TASK2_TRSM, __shfl_sync(FULL, hh, pick), 0, __shfl_sync(FULL, kk, pick) | (1 << 16));
          claimed = true;
        }
      }
    }
    __syncwarp();
    // ---- updates, from the front outwards.  Columns within W of the front take whatever is available; farther
    //      columns wait for a full chunk of G panels; if nothing qualifies the first ready far tile is taken
    //      anyway (better than idling).
    if (!claimed) {
      int fb_i = -1, fb_j = 0, fb_a = 0, fb_nk = 0;
      // Columns behind the front can still owe updates (potrf(j) only needs tile (j, j)): the scan starts at the
      // first column not known to be complete and moves that hint forward as it goes.
      const int j_start = min(ld_relaxed(P.flags + F2_JMIN), T);
      int jinc = 1 << 20;
      int c0 = P.col_start[j_start];
      for (; c0 < P.n_tiles && !claimed; c0 += 128) {
        int ti[4], tj[4], ta[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + 32 * u + lane;
          const int e = c < P.n_tiles ? __ldg(P.tiles + c) : -1;
          ti[u] = e < 0 ? -1 : (e & 0xffff);
          tj[u] = e < 0 ? 0 : (e >> 16);
          ta[u] = e < 0 ? ST_LOCK : ld_relaxed(st + (size_t)ti[u] * T + tj[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = ti[u], j = tj[u];
          if (i >= 0 && (ta[u] & ~ST_LOCK) < lim2(P, i, j)) jinc = min(jinc, j);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (claimed) break;
          bool ok = false, fb = false;
          int nk = 0;
          const int i = ti[u], j = tj[u], a = ta[u];
          if (i >= 0 && !(a & ST_LOCK)) {
            const int lim = lim2(P, i, j);
            const int av = min(min(s_rowsv[i], s_rowsv[j]), lim);
            nk = min(av - a, P.G);
            if (nk > 0) {
              // the farther from the front, the longer the chunk a tile waits for: 1 panel inside the window W,
              // one more per block column of distance, G at most — tiles catch up gradually as the front approaches
              // instead of owing G - 1 panels when they become urgent
              ok = av - a >= min(P.G, max(1, j - np - P.W + 1));
              fb = !ok;
            }
          }
          const unsigned m = __ballot_sync(FULL, ok);
          if (m) {
            const int cnt = __popc(m);
            const int pick = __fns(m, 0, 1 + (wid % cnt));
            int got = 0;
            if (lane == pick) got = (atomicCAS(st + (size_t)i * T + j, a, a | ST_LOCK) == a);
            got = __shfl_sync(FULL, got, pick);
            if (got) {
              res = make_int4(TASK2_UPD, __shfl_sync(FULL, i, pick), __shfl_sync(FULL, j, pick), __shfl_sync(FULL, a, pick) | (__shfl_sync(FULL, nk, pick) << 16));
              claimed = true;
            }
          } else if (fb_i < 0) {
            const unsigned mf = __ballot_sync(FULL, fb);
            if (mf) {
              const int pick = __fns(mf, 0, 1 + (wid % __popc(mf)));
              fb_i = __shfl_sync(FULL, i, pick); fb_j = __shfl_sync(FULL, j, pick); fb_a = __shfl_sync(FULL, a, pick); fb_nk = __shfl_sync(FULL, nk, pick);
            }
          }
        }
      }
      {
        // every tile before position c0 has been looked at: the hint may move up to the first incomplete column seen
        const int j_unseen = c0 < P.n_tiles ? (__ldg(P.tiles + c0) >> 16) : T;
        const int j_new = min(__reduce_min_sync(FULL, jinc), j_unseen);
        if (lane == 0 && j_new > j_start) atomicMax(P.flags + F2_JMIN, j_new);
      }
      if (!claimed && fb_i >= 0) {
        int got = 0;
        if (lane == 0) got = (atomicCAS(st + (size_t)fb_i * T + fb_j, fb_a, fb_a | ST_LOCK) == fb_a);
        got = __shfl_sync(FULL, got, 0);
        if (got) { res = make_int4(TASK2_UPD, fb_i, fb_j, fb_a | (fb_nk << 16)); claimed = true; }
      }
    }
    // ---- completion of the diagonal inverses (needed by the backward substitution only) ----
    for (int k0 = 0; k0 < np && !claimed; k0 += 32) {
      const int k = k0 + lane;
      const bool ok = k < np && ld_relaxed(invst + k) == 0;
      const unsigned m = __ballot_sync(FULL, ok);
      if (m) {
        const int pick = __ffs(m) - 1;
        int got = 0;
        if (lane == pick) got = (atomicCAS(invst + k, 0, 1) == 0);
        got = __shfl_sync(FULL, got, pick);
        if (got) { res = make_int4(TASK2_INV, 0, 0, (k0 + pick) | (1 << 16)); claimed = true; }
      }
    }
    if (claimed) break;
    if (ld_relaxed(P.flags + F2_DONE) >= P.total_units) break;
    if (++spins > (SPIN_LIMIT >> 3)) { if (lane == 0) dag2_abort(P); break; }
    // Nothing ready.  In the tail of the factorisation a handful of tasks exist at any time and ~140 idle workers
    // scanning and CAS-ing the same few words slow down the ones that matter: back off, the more the longer nothing
    // was found, and park the workers the remaining trailing matrix (< (T - np)^2 tiles) cannot employ anyway.
    const int rem = max(T - np, 1);
    const bool surplus = wid >= 8 + 2 * rem * rem;
    __nanosleep(surplus ? 3000 : min(200u << min(spins, 4ll), 1600u));
  }
  __threadfence();
  if (lane == 0) *t_wait += clock64() - c0;
  return res;
}

__global__ void __launch_bounds__(DAG_THREADS, 1) k_chol_dag2(const Dag2Params P) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int4 s_task;
  __shared__ int s_ok;
  __shared__ int s_sig;
  __shared__ int s_prog;
  __shared__ unsigned long long s_cbar;
  __shared__ long long s_prof[16];      // [0] wait [1] potrf [2] trsm [3] upd [4] inv [5] tasks [6] start, [8..11] upd phases
  __shared__ int s_rowsv[MAX_TR];
  const int tid = threadIdx.x, b = blockIdx.x;
  const int T = P.T;
  unsigned cphase = 0;
  if (tid == 0) {
    mbar_init(&s_cbar, 1);
    for (int i = 0; i < 16; ++i) s_prof[i] = 0;
    s_prof[6] = clock64();
  }
  __syncthreads();
  if (b == 0) {
    // ---- diagonal blocks ----
    for (int k = 0; k < T; ++k) {
      const int k0 = k * NB, nb = min(NB, P.n - k0);
      // every update but the last (panel k - 1: streamed in by potrf128_prog_dev from the solve of tile row k)
      if (!wait2(P, f2_st(P) + (size_t)k * T + k, k - 1, nullptr, 0, &s_ok, s_prof)) break;
      TRACE(P, k, 0);
      const long long c0 = clock64();
      potrf128_prog_dev(P.S, P.ld, k0, nb, P.Linv + (size_t)k * NB * NB, P.info, sm, f2_pb(P) + k, f2_pinv(P) + k, &s_sig, &s_prog,
                        k > 0 ? f2_xpub(P) + 2 * k : nullptr, 4 * (k - 1), P.n_rows, P.flags + F2_ABORT,
                        (k > 0 && half2_exists(P, 2 * k + 1)) ? f2_xpub(P) + 2 * k + 1 : nullptr);
      __threadfence();
      __syncthreads();
      TRACE(P, k, 1);
      if (tid == 0) { st_release_gpu(f2_pdone(P) + k, 1); st_release_gpu(P.flags + F2_NPAN, k + 1); s_prof[1] += clock64() - c0; s_prof[5] += 1; }
      // the right-hand-side row n lives inside the last diagonal tile when n is not a multiple of 128
      if (k == T - 1 && P.n_rows > P.n && P.n < T * NB) {
        trsm64_dev(P.S, P.ld, P.n_rows, k0, nb, P.n, P.Linv + (size_t)k * NB * NB, sm, DAG_THREADS / 32);
        __syncthreads();
      }
    }
  } else if (b <= 2 * P.D) {
    // ---- solve of 64 rows of tile row k + di against panel k, following CTA 0 ----
    const int di = (b - 1) / 2 + 1, half = (b - 1) & 1;
    for (int k = 0; k < T; ++k) {
      if (k + di >= P.Tr) break;
      if (!solve_row_dev(P, k + di, half, k, sm, &s_ok, s_prof, b == 1 ? 2 : -1)) break;
    }
  } else if (b <= 2 * P.D + P.D * (P.D + 1) - 2) {
    // ---- update of 64 rows of tile (k + di, k + dj) with panel k, following the solves of its two tile rows; (1, 1),
    //      the last update of the diagonal tile, happens inside CTA 0 ----
    const int u = b - 2 * P.D - 1, half = u & 1;
    int r = u / 2 + 1, di = 1;
    while (r >= di) { r -= di; ++di; }
    const int dj = r + 1;
    const int slot = half ? -1 : (di == 2 && dj == 1) ? 10 : (di == 2 && dj == 2) ? 9 : -1;
    for (int k = 0; k < T; ++k) {
      if (k + dj >= T || k + di >= P.Tr) break;
      if (!upd_stream_dev(P, k + di, k + dj, half, k, sm, &s_ok, s_prof, &s_cbar, cphase, slot)) break;
    }
  } else {
    for (;;) {
      if (tid < 32) {
        const int4 t = find_task2(P, b, s_prof, s_rowsv);
        if (tid == 0) s_task = t;
      }
      __syncthreads();
      const int4 t = s_task;
      __syncthreads();
      if (t.x == TASK2_EXIT) break;
      const long long c0 = clock64();
      const int k = t.w & 0xffff, nk = t.w >> 16, k0 = k * NB, nb = min(NB, P.n - k0);
      if (t.x == TASK2_TRSM) trsm64_dev(P.S, P.ld, P.n_rows, k0, nb, t.y * 64, P.Linv + (size_t)k * NB * NB, sm, DAG_THREADS / 32);
      else if (t.x == TASK2_UPD) upd_dev<4>(P.S, P.ld, P.n_rows, P.n, k0, NB * nk, t.y * NB, t.z * NB, sm, s_prof + 8, &s_cbar, cphase);
      else inv_offdiag_dev(P.S, P.ld, k0, nb, P.Linv + (size_t)k * NB * NB, sm);
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        if (t.x == TASK2_TRSM) st_release_gpu(f2_sv(P) + t.y, k + 1);
        else if (t.x == TASK2_UPD) st_release_gpu(f2_st(P) + (size_t)t.y * T + t.z, k + nk);
        else st_release_gpu(f2_invst(P) + k, 2);
        atomicAdd(P.flags + F2_DONE, nk);
        const long long dt = clock64() - c0;
        if (t.x == TASK2_TRSM) s_prof[2] += dt; else if (t.x == TASK2_UPD) s_prof[3] += dt; else s_prof[4] += dt;
        s_prof[5] += 1;
      }
    }
  }
  if (P.prof && tid == 0) {
    long long* o = P.prof + (size_t)b * 16;
    for (int i = 0; i < 4; ++i) o[8 + i] = s_prof[8 + i];
    for (int i = 0; i < 6; ++i) o[i] = s_prof[i];
    o[6] = clock64() - s_prof[6];
  }
}

// ---- split factorisation (multi-GPU): the Schur-complement update of the trailing block ---------------------------
// C(i, j) -= L(i, 0 : K) L(j, 0 : K)^T for a LIST of 128 x 128 tiles (i | j << 16) of the trailing matrix, K = the
// width of the factored block-column range: the one embarrassingly parallel piece of a Cholesky factorisation, so the
// piece that is spread over the GPUs (chol_factor_solve_split).  Persistent CTAs, tiles round-robin, upd_dev<4>.
__global__ void __launch_bounds__(DAG_THREADS, 1) k_upd_list(double* __restrict__ S, int ld, int n_rows, int n, int K,
                                                             const int* __restrict__ tiles, int n_list) {
  extern __shared__ __align__(16) double sm[];
  __shared__ unsigned long long s_cbar;
  __shared__ long long s_prof[8];
  unsigned cphase = 0;
  if (threadIdx.x == 0) mbar_init(&s_cbar, 1);
  __syncthreads();
  for (int t = blockIdx.x; t < n_list; t += gridDim.x) {
    const int e = __ldg(tiles + t);
    upd_dev<4>(S, ld, n_rows, n, 0, K, (e & 0xffff) * NB, (e >> 16) * NB, sm, s_prof, &s_cbar, cphase);
    __syncthreads();
  }
}
// pack / unpack of trailing-matrix tiles (column-major 128 x 128 pieces of S, clipped to the n_rows x n matrix) into a
// contiguous buffer, tile t at t * NB * NB: what one rank computed goes out to the others in one piece
__global__ void k_tiles_pack(const double* __restrict__ S, int ld, int n_rows, int n, const int* __restrict__ tiles, int t0, double* __restrict__ buf, int unpack) {
  const int e = __ldg(tiles + t0 + blockIdx.x), i0 = (e & 0xffff) * NB, j0 = (e >> 16) * NB;
  double* b = buf + (size_t)(t0 + blockIdx.x) * NB * NB;
  for (int x = threadIdx.x; x < NB * NB; x += blockDim.x) {
    const int r = i0 + (x & (NB - 1)), c = j0 + (x >> 7);
    if (r < n_rows && c < n) {
      double* sp = const_cast<double*>(S) + (size_t)c * ld + r;
      if (unpack) *sp = b[x]; else b[x] = *sp;
    }
  }
}

// The one-launch substitution kernels spin on flags of CTAs with a smaller block index.  CUDA does not promise
// in-order dispatch once a grid exceeds the resident capacity, so the grid is cut into launches of at most one
// CTA per SM: inside a launch every CTA is resident, across launches the producers have already finished.
static int substitution_chunk() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 1;
  }
  return sms;
}
#define CKC(call)                                                                                           \
  do {                                                                                                      \
    cudaError_t e_ = (call);                                                                                \
    if (e_ != cudaSuccess) {                                                                                \
      fprintf(stderr, "[stba] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return STBA_ERR_CUDA;                                                                                 \
    }                                                                                                       \
  } while (0)

}  // namespace

}  // namespace stba
#include "stba_chol_small.cuh"
namespace stba {

struct CholPlan {
  double* S = nullptr;
  double* rhs = nullptr;
  int* info = nullptr;
  int n = 0, ld = 0;
  double* Linv = nullptr;     // T blocks of NB x NB
  double* ybuf = nullptr;     // intermediate vector of the triangular solves
  int2* tiles = nullptr;      // device tile lists
  int* flags = nullptr;       // x_k-ready flags of the one-launch backward substitution
  cudaGraphExec_t exec = nullptr;
  // DAG-scheduled persistent factorisation (default)
  bool dag = false;
  int4 *hp = nullptr, *ur = nullptr, *lp = nullptr;
  int* hp_end = nullptr;
  int n_hp = 0, n_ur = 0, n_lp = 0, grid = 0, R64 = 0, Tr = 0;
  int* dflags = nullptr;
  size_t n_dflags = 0;
  int dag_version = 2;        // 2: scan scheduler (default), 1: ticket queues (STBA_CHOL_DAG1=1)
  cudaStream_t pool_stream = nullptr;   // DAG 2: every buffer is stream-ordered (pooled): no cudaMalloc / cudaFree stalls per problem
  int* d_tiles = nullptr;     // DAG 2: tile list + column starts
  int n_tiles = 0, total_units = 0, W = 2, G = 8, D = 1;
  long long* prof = nullptr;
  unsigned long long* trace = nullptr;
  cudaStream_t side = nullptr, inv = nullptr;
  std::vector<cudaEvent_t> events;
  int launches = 0;
  bool rest64 = false;        // bulk trailing update in 64-row tiles, two CTAs per SM
  std::vector<size_t> off_strip, off_rest, off_panel;
  std::vector<int> n_strip, n_rest, n_panel;
};

static void destroy_plan(CholPlan* p) {
  if (!p) return;
  if (p->pool_stream) {      // DAG 2: stream-ordered buffers
    cudaStream_t st = p->pool_stream;
    if (p->Linv) cudaFreeAsync(p->Linv, st);
    if (p->ybuf) cudaFreeAsync(p->ybuf, st);
    if (p->flags) cudaFreeAsync(p->flags, st);
    if (p->dflags) cudaFreeAsync(p->dflags, st);
    if (p->d_tiles) cudaFreeAsync(p->d_tiles, st);
    if (p->prof) cudaFree(p->prof);
    if (p->trace) cudaFree(p->trace);
    delete p;
    return;
  }
  if (p->exec) cudaGraphExecDestroy(p->exec);
  if (p->Linv) cudaFree(p->Linv);
  if (p->ybuf) cudaFree(p->ybuf);
  if (p->tiles) cudaFree(p->tiles);
  if (p->flags) cudaFree(p->flags);
  if (p->hp) cudaFree(p->hp);
  if (p->lp) cudaFree(p->lp);
  if (p->ur) cudaFree(p->ur);
  if (p->hp_end) cudaFree(p->hp_end);
  if (p->dflags) cudaFree(p->dflags);
  if (p->d_tiles) cudaFree(p->d_tiles);
  if (p->prof) cudaFree(p->prof);
  if (p->trace) cudaFree(p->trace);
  if (p->side) cudaStreamDestroy(p->side);
  if (p->inv) cudaStreamDestroy(p->inv);
  for (auto e : p->events) cudaEventDestroy(e);
  delete p;
}

CholWorkspace::~CholWorkspace() { destroy_plan(plan); }
void CholWorkspace::reset() { destroy_plan(plan); plan = nullptr; }

// Enqueue the whole factor + solve schedule on `main` (and `side` for the look-ahead panels).
static int enqueue(CholPlan& P, cudaStream_t main, bool lookahead) {
  const int n = P.n, ld = P.ld, T = (n + NB - 1) / NB, n_rows = n + 1;
  double* S = P.S;
  P.launches = 0;
  const std::vector<size_t>&off_strip = P.off_strip, &off_rest = P.off_rest;
  const std::vector<int>&n_strip = P.n_strip, &n_rest = P.n_rest;
  size_t ev = 0;
  auto next_event = [&]() -> cudaEvent_t {
    if (ev == P.events.size()) {
      cudaEvent_t e;
      cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      P.events.push_back(e);
    }
    return P.events[ev++];
  };
  k_put_row<<<(n + 255) / 256, 256, 0, main>>>(S, ld, n, P.rhs);
  ++P.launches;
  // panel(k) = potrf + panel solve on stream s; the completion of the inverse block forks off to P.inv
  auto panel = [&](int k, cudaStream_t s) -> int {
    const int k0 = k * NB, nb = std::min(NB, n - k0);
    double* Li = P.Linv + (size_t)k * NB * NB;
    k_potrf128<<<1, PT, POTRF_SMEM, s>>>(S, ld, k0, nb, Li, P.info);
    ++P.launches;
    cudaEvent_t ep = next_event();
    CKC(cudaEventRecord(ep, s));
    CKC(cudaStreamWaitEvent(P.inv, ep, 0));
    k_inv_offdiag<<<1, PT, POTRF_SMEM, P.inv>>>(S, ld, k0, nb, Li);
    ++P.launches;
    const int first = k0 + NB;                       // rows below the diagonal block (the rhs row n is the last one)
    if (first < n_rows) {
      k_trsm_sub<<<(n_rows - first + TS_ROWS - 1) / TS_ROWS, TS_THREADS, TS_SMEM, s>>>(S, ld, n_rows, k0, nb, first, Li);
      ++P.launches;
    } else {
      k_trsm_sub<<<1, TS_THREADS, TS_SMEM, s>>>(S, ld, n_rows, k0, nb, n, Li);     // last block: only the rhs row
      ++P.launches;
    }
    return STBA_OK;
  };
  if (panel(0, main) != STBA_OK) return STBA_ERR_CUDA;
  for (int k = 0; k + 1 < T; ++k) {
    const int k0 = k * NB;
    // strip update of block column k+1, then its panel (look-ahead: on the side stream)
    k_gemm_nt<MODE_SYRK, 64><<<n_strip[k], GEMM_THREADS, gemm_smem(64), main>>>(S, ld, n_rows, n, k0, NB, nullptr, ld, P.tiles + off_strip[k]);
    ++P.launches;
    if (lookahead && n_rest[k]) {
      cudaEvent_t e1 = next_event(), e2 = next_event();
      CKC(cudaEventRecord(e1, main));
      CKC(cudaStreamWaitEvent(P.side, e1, 0));
      if (panel(k + 1, P.side) != STBA_OK) return STBA_ERR_CUDA;
      CKC(cudaEventRecord(e2, P.side));
      if (P.rest64) k_gemm_nt<MODE_SYRK, 64><<<n_rest[k], GEMM_THREADS, gemm_smem(64), main>>>(S, ld, n_rows, n, k0, NB, nullptr, ld, P.tiles + off_rest[k]);
      else k_gemm_nt<MODE_SYRK, 128><<<n_rest[k], GEMM_THREADS, gemm_smem(128), main>>>(S, ld, n_rows, n, k0, NB, nullptr, ld, P.tiles + off_rest[k]);
      ++P.launches;
      CKC(cudaStreamWaitEvent(main, e2, 0));
    } else {
      if (n_rest[k]) {
        if (P.rest64) k_gemm_nt<MODE_SYRK, 64><<<n_rest[k], GEMM_THREADS, gemm_smem(64), main>>>(S, ld, n_rows, n, k0, NB, nullptr, ld, P.tiles + off_rest[k]);
        else k_gemm_nt<MODE_SYRK, 128><<<n_rest[k], GEMM_THREADS, gemm_smem(128), main>>>(S, ld, n_rows, n, k0, NB, nullptr, ld, P.tiles + off_rest[k]);
        ++P.launches;
      }
      if (panel(k + 1, main) != STBA_OK) return STBA_ERR_CUDA;
    }
  }
  // y = L^-1 rhs now sits in row n; backward substitution into rhs (needs the completed inverse blocks)
  {
    cudaEvent_t ei = next_event();
    CKC(cudaEventRecord(ei, P.inv));
    CKC(cudaStreamWaitEvent(main, ei, 0));
  }
  k_get_row<<<(n + 255) / 256, 256, 0, main>>>(S, ld, n, P.ybuf);
  ++P.launches;
  CKC(cudaMemsetAsync(P.flags, 0, (size_t)T * sizeof(int), main));
  for (int b0 = 0; b0 < T; b0 += substitution_chunk())
    k_trsv_bwd_all<<<std::min(T - b0, substitution_chunk()), TBA_THREADS, TBA_SMEM, main>>>(S, ld, n, T, P.Linv, P.ybuf, P.rhs, P.flags, P.info, b0);
  ++P.launches;
  CKC(cudaGetLastError());
  return STBA_OK;
}

// Task queues of the DAG kernel (host, once per plan).  High priority: for every block column k the panel
// solves of the 64-row halves below tile (k+1, k), then the updates of block column k+1; low priority: the
// bulk trailing update of step k, block columns ascending (the next panel's column first), then the
// completion of the diagonal inverse of block k.  Queue order is a topological order of the task graph.
static int build_dag_plan(CholPlan& P) {
  const int n = P.n, T = (n + NB - 1) / NB, n_rows = n + 1;
  P.Tr = (n_rows + NB - 1) / NB;
  P.R64 = (n_rows + 63) / 64;
  // Three queues, each in a topological order of the task graph (sorted by the last block column read):
  //   high priority  step k: panel solves of the 64-row halves below tile row k + 1, then the 64-row updates of
  //                  block column k + 1 (tile row k + 1 itself belongs to CTA 1);
  //   urgent         step k: single-panel updates of block columns k + 2 .. k + 1 + W — everything the chain
  //                  needs from the trailing matrix, so the factorisation runs ahead of the bulk;
  //   far            block columns beyond the window receive their updates lazily, G block columns at a
  //                  time (K = 128 G: one C-tile round trip per G updates), the remainder when the column enters
  //                  the window; plus the completion of the diagonal inverses for the backward substitution.
  int W = 2, G = 3;
  if (const char* s = getenv("STBA_CHOL_WINDOW")) W = std::max(1, atoi(s));
  if (const char* s = getenv("STBA_CHOL_AGG")) G = std::max(1, std::min(16, atoi(s)));
  auto enc = [](int klo, int nk) { return klo | (nk << 16); };
  std::vector<int4> hp, ur;
  std::vector<std::vector<int4>> far_by_level(T);
  std::vector<int> hp_end(T, 0), ur_end(T, 0);
  for (int k = 0; k < T; ++k) {
    // (the update of the same rows in block column k + 1 follows each solve on the same worker)
    for (int h = 2 * (k + 2); h < P.R64; ++h) hp.push_back(make_int4(TASK_TRSM, h, 0, enc(k, 1)));
    hp_end[k] = (int)hp.size();
    // urgent: block column j receives panel k as a single update iff k >= j - 1 - W (and k <= j - 2); tile rows
    // nearest the diagonal first — those are the ones the chain waits for

    far_by_level[k].push_back(make_int4(TASK_INV, 0, 0, enc(k, 1)));
  }
  // far: block column j receives panels 0 .. j - 2 - W in aligned chunks of G, queued at the level of the chunk's last
  // panel.  The remainder (fewer than G panels, applied when the column enters the window) is URGENT: queued
  // behind the bulk it would stall the chain for a whole level of far work.
  std::vector<std::vector<int4>> rem_by_level(T);
  for (int j = 2; j < T; ++j) {
    const int last = j - 2 - W;        // last far panel of this block column
    for (int klo = 0; klo <= last; klo += G) {
      const int nk = std::min(G, last - klo + 1);
      auto& dst = (nk == G) ? far_by_level[klo + nk - 1] : rem_by_level[klo + nk - 1];
      for (int i = j; i < P.Tr; ++i) dst.push_back(make_int4(TASK_UPD, i, j, enc(klo, nk)));
    }
  }
  for (int k = 0; k < T; ++k) {
    for (int i = k + 2; i < P.Tr; ++i)
      for (int j = k + 2; j < T && j <= k + 1 + W && j <= i; ++j) ur.push_back(make_int4(TASK_UPD, i, j, enc(k, 1)));
    ur.insert(ur.end(), rem_by_level[k].begin(), rem_by_level[k].end());    // needed by the urgent updates of step k + 1
    ur_end[k] = (int)ur.size();
  }
  std::vector<int4> lp;
  for (int k = 0; k < T; ++k) {
    // within a level: block columns ascending (stable: the chunks were appended column by column), the inverse last
    std::stable_sort(far_by_level[k].begin(), far_by_level[k].end(), [](const int4& a, const int4& b) {
      const int ja = a.x == TASK_INV ? (1 << 30) : a.z, jb = b.x == TASK_INV ? (1 << 30) : b.z;
      return ja < jb;
    });
    lp.insert(lp.end(), far_by_level[k].begin(), far_by_level[k].end());
  }
  P.n_hp = (int)hp.size();
  P.n_ur = (int)ur.size();
  P.n_lp = (int)lp.size();
  CKC(cudaMalloc(&P.hp, std::max<size_t>(hp.size(), 1) * sizeof(int4)));
  CKC(cudaMalloc(&P.ur, std::max<size_t>(ur.size(), 1) * sizeof(int4)));
  CKC(cudaMalloc(&P.lp, std::max<size_t>(lp.size(), 1) * sizeof(int4)));
  CKC(cudaMemcpy(P.hp, hp.data(), hp.size() * sizeof(int4), cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(P.ur, ur.data(), ur.size() * sizeof(int4), cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(P.lp, lp.data(), lp.size() * sizeof(int4), cudaMemcpyHostToDevice));
  CKC(cudaMalloc(&P.hp_end, 2 * T * sizeof(int)));
  CKC(cudaMemcpy(P.hp_end, hp_end.data(), T * sizeof(int), cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(P.hp_end + T, ur_end.data(), T * sizeof(int), cudaMemcpyHostToDevice));
  P.n_dflags = (size_t)F_PDONE + T + (size_t)P.R64 * T + 2 * (size_t)T;
  CKC(cudaMalloc(&P.dflags, P.n_dflags * sizeof(int)));
  int dev = 0, sms = 0;
  CKC(cudaGetDevice(&dev));
  CKC(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  P.grid = sms;
  if (const char* g = getenv("STBA_CHOL_GRID")) P.grid = std::max(3, std::min(sms, atoi(g)));
  CKC(cudaFuncSetAttribute(k_chol_dag, cudaFuncAttributeMaxDynamicSharedMemorySize, DAG_SMEM));
  if (getenv("STBA_CHOL_PROF")) {
    CKC(cudaMalloc(&P.prof, (size_t)P.grid * 16 * sizeof(long long)));
    CKC(cudaMemset(P.prof, 0, (size_t)P.grid * 16 * sizeof(long long)));
    CKC(cudaMalloc(&P.trace, (size_t)T * 16 * sizeof(unsigned long long)));
    CKC(cudaMemset(P.trace, 0, (size_t)T * 16 * sizeof(unsigned long long)));
  }
  return STBA_OK;
}

static int run_dag(CholPlan& P, cudaStream_t stream) {
  const int n = P.n, ld = P.ld, T = (n + NB - 1) / NB;
  k_put_row<<<(n + 255) / 256, 256, 0, stream>>>(P.S, ld, n, P.rhs);
  CKC(cudaMemsetAsync(P.dflags, 0, P.n_dflags * sizeof(int), stream));
  DagParams dp;
  dp.S = P.S; dp.ld = ld; dp.n = n; dp.n_rows = n + 1; dp.T = T; dp.Tr = P.Tr; dp.R64 = P.R64;
  dp.Linv = P.Linv; dp.info = P.info; dp.flags = P.dflags;
  dp.hp = P.hp; dp.ur = P.ur; dp.lp = P.lp; dp.n_hp = P.n_hp; dp.n_ur = P.n_ur; dp.n_lp = P.n_lp; dp.hp_end = P.hp_end; dp.ur_end = P.hp_end + T; dp.prof = P.prof; dp.trace = P.trace;
  void* args[] = {&dp};
  CKC(cudaLaunchCooperativeKernel((const void*)k_chol_dag, dim3(P.grid), dim3(DAG_THREADS), args, DAG_SMEM, stream));
  k_get_row<<<(n + 255) / 256, 256, 0, stream>>>(P.S, ld, n, P.ybuf);
  CKC(cudaMemsetAsync(P.flags, 0, (size_t)T * sizeof(int), stream));
  for (int b0 = 0; b0 < T; b0 += substitution_chunk())
    k_trsv_bwd_all<<<std::min(T - b0, substitution_chunk()), TBA_THREADS, TBA_SMEM, stream>>>(P.S, ld, n, T, P.Linv, P.ybuf, P.rhs, P.flags, P.info, b0);
  CKC(cudaGetLastError());
  P.launches = 4;
  if (P.prof) {
    CKC(cudaStreamSynchronize(stream));
    std::vector<long long> h((size_t)P.grid * 16);
    CKC(cudaMemcpy(h.data(), P.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    auto row = [&](int b) {
      const long long* o = &h[(size_t)b * 16];
      fprintf(stderr, "[chol dag] cta %3d: wait %8lld potrf %8lld trsm %8lld upd %8lld inv %8lld tasks %5lld total %8lld | upd phases: issue %lld first %lld loop %lld store %lld\n", b, o[0], o[1], o[2], o[3],
              o[4], o[5], o[6], o[8], o[9], o[10], o[11]);
    };
    for (int b = 0; b < std::min(P.grid, 4); ++b) row(b);
    long long w = 0, tr = 0, up = 0, iv = 0, nt = 0, tot = 0;
    for (int b = 2; b < P.grid; ++b) { const long long* o = &h[(size_t)b * 16]; w += o[0]; tr += o[2]; up += o[3]; iv += o[4]; nt += o[5]; tot += o[6]; }
#ifdef STBA_CHOL_TIMING
    {
      long long hc[64];
      cudaMemcpyFromSymbol(hc, g_potrf_clk, sizeof(hc));
      long long ht[16];
      cudaMemcpyFromSymbol(ht, g_trsm_clk, sizeof(ht));
      fprintf(stderr, "[trsm clocks, last call of CTA 0]");
      for (int i = 1; i <= 9; ++i) fprintf(stderr, " %d:%lld", i, ht[i] - ht[i - 1]);
      fprintf(stderr, "\n[potrf128 clocks, last panel]");
      for (int i = 1; i <= 16; ++i) fprintf(stderr, " %d:%lld", i, hc[i] - hc[i - 1]);
      for (int b = 0; b < 4; ++b) fprintf(stderr, " potf2_%d:%lld", b, hc[20 + b] - hc[b ? 1 + 3 * b : 1]);
      fprintf(stderr, "\n");
    }
#endif
    if (P.trace) {
      std::vector<unsigned long long> tr((size_t)T * 16);
      CKC(cudaMemcpy(tr.data(), P.trace, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      const unsigned long long t0 = tr[0];
      fprintf(stderr, "[chol trace] us since potrf(0) start: k | P start, P end | TU in, X ready, C ready, step0..3 start, X published, TU end\n");
      fprintf(stderr, "[chol trace] potrf start times (us):");
      for (int k = 0; k < T; ++k) fprintf(stderr, " %.0f", (double)(tr[(size_t)k * 16] - t0) * 1e-3);
      fprintf(stderr, "\n");
      for (int k = 0; k < T; k += (k < 4 || k + 5 > T) ? 1 : 6) {
        fprintf(stderr, "[chol trace] %2d |", k);
        for (int s = 0; s <= 10; ++s) fprintf(stderr, " %8.1f%s", tr[(size_t)k * 16 + s] ? (double)(tr[(size_t)k * 16 + s] - t0) * 1e-3 : -1.0, (s == 1) ? " |" : "");
        fprintf(stderr, "\n");
      }
    }
    const double nw = std::max(1, P.grid - 2);
    fprintf(stderr, "[chol dag] workers (mean cycles): wait %.0f trsm %.0f upd %.0f inv %.0f tasks %.1f total %.0f\n", w / nw, tr / nw, up / nw,
            iv / nw, nt / nw, tot / nw);
  }
  return STBA_OK;
}


// DAG 2 (scan scheduler): state tables, the tile list in scan order and the number of work units the workers owe.
static int build_dag2_plan(CholPlan& P, cudaStream_t stream) {
  const int n = P.n, T = (n + NB - 1) / NB, n_rows = n + 1;
  P.Tr = (n_rows + NB - 1) / NB;
  P.R64 = (n_rows + 63) / 64;
  P.W = 2; P.G = 8; P.D = 1;
  if (const char* s = getenv("STBA_CHOL_DEPTH")) P.D = std::max(1, std::min(4, atoi(s)));
  if (const char* s = getenv("STBA_CHOL_WINDOW")) P.W = std::max(0, atoi(s));
  if (const char* s = getenv("STBA_CHOL_AGG")) P.G = std::max(1, std::min(16, atoi(s)));
  std::vector<int> tiles, col_start(T + 1, 0);
  long long units = T;                               // the inverse completions
  for (int j = 0; j < T; ++j) {
    col_start[j] = (int)tiles.size();
    for (int i = j; i < P.Tr; ++i) {
      tiles.push_back(i | (j << 16));
      units += std::max(std::min(j, i - P.D) - ((P.D == 1 && i - j == 1) ? 1 : 0), 0);      // = lim2() of the kernel
    }
  }
  col_start[T] = (int)tiles.size();
  for (int h = 0; h < P.R64; ++h) units += std::max(0, std::min((h >> 1) - P.D, T));     // panel solves of tile rows > k + D
  P.n_tiles = (int)tiles.size();
  P.total_units = (int)units;
  std::vector<int> both(tiles);
  both.insert(both.end(), col_start.begin(), col_start.end());
  CKC(cudaMallocAsync((void**)&P.d_tiles, both.size() * sizeof(int), stream));
  CKC(cudaMemcpyAsync(P.d_tiles, both.data(), both.size() * sizeof(int), cudaMemcpyHostToDevice, stream));    // pageable source: staged before the call returns
  P.n_dflags = (size_t)F2_ARR + 4 * (size_t)T + 2 * P.Tr + P.R64 + 2 * (size_t)P.Tr * T;
  CKC(cudaMallocAsync((void**)&P.dflags, P.n_dflags * sizeof(int), stream));
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    CKC(cudaGetDevice(&dev));
    CKC(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  P.grid = sms;
  const int n_ded = 2 * P.D + P.D * (P.D + 1) - 1;
  if (const char* g = getenv("STBA_CHOL_GRID")) P.grid = std::max(n_ded + 1, std::min(sms, atoi(g)));
  if (P.grid < n_ded + 1) return STBA_ERR_UNSUPPORTED;
  {
    static bool attr_set = false;      // per process (one device per process)
    if (!attr_set) { CKC(cudaFuncSetAttribute(k_chol_dag2, cudaFuncAttributeMaxDynamicSharedMemorySize, DAG_SMEM)); attr_set = true; }
  }
  if (getenv("STBA_CHOL_PROF")) {
    CKC(cudaMalloc(&P.prof, (size_t)P.grid * 16 * sizeof(long long)));
    CKC(cudaMemset(P.prof, 0, (size_t)P.grid * 16 * sizeof(long long)));
    CKC(cudaMalloc(&P.trace, (size_t)T * 16 * sizeof(unsigned long long)));
    CKC(cudaMemset(P.trace, 0, (size_t)T * 16 * sizeof(unsigned long long)));
  }
  return STBA_OK;
}

static int run_dag2(CholPlan& P, cudaStream_t stream) {
  const int n = P.n, ld = P.ld, T = (n + NB - 1) / NB;
  k_put_row<<<(n + 255) / 256, 256, 0, stream>>>(P.S, ld, n, P.rhs);
  CKC(cudaMemsetAsync(P.dflags, 0, P.n_dflags * sizeof(int), stream));
  Dag2Params dp;
  dp.S = P.S; dp.ld = ld; dp.n = n; dp.n_rows = n + 1; dp.T = T; dp.Tr = P.Tr; dp.R64 = P.R64;
  dp.Linv = P.Linv; dp.info = P.info; dp.flags = P.dflags;
  dp.total_units = P.total_units; dp.W = P.W; dp.G = P.G; dp.D = P.D;
  dp.tiles = P.d_tiles; dp.col_start = P.d_tiles + P.n_tiles; dp.n_tiles = P.n_tiles;
  dp.prof = P.prof; dp.trace = P.trace;
  void* args[] = {&dp};
  CKC(cudaLaunchCooperativeKernel((const void*)k_chol_dag2, dim3(P.grid), dim3(DAG_THREADS), args, DAG_SMEM, stream));
  k_get_row<<<(n + 255) / 256, 256, 0, stream>>>(P.S, ld, n, P.ybuf);
  CKC(cudaMemsetAsync(P.rhs, 0xFF, (size_t)n * sizeof(double), stream));      // NaN = "not computed yet" (k_trsv_bwd_all, poll_data)
  for (int b0 = 0; b0 < T; b0 += substitution_chunk())
    k_trsv_bwd_all<<<std::min(T - b0, substitution_chunk()), TBA_THREADS, TBA_SMEM, stream>>>(P.S, ld, n, T, P.Linv, P.ybuf, P.rhs, P.flags, P.info, b0, 1);
  CKC(cudaGetLastError());
  P.launches = 4;
  if (P.prof) {
    CKC(cudaStreamSynchronize(stream));
    std::vector<long long> h((size_t)P.grid * 16);
    CKC(cudaMemcpy(h.data(), P.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    const int n_ded = 2 * P.D + P.D * (P.D + 1) - 1;
    for (int b = 0; b < std::min(P.grid, n_ded + 2); ++b) {
      const long long* o = &h[(size_t)b * 16];
      fprintf(stderr, "[chol dag2] cta %3d: wait %8lld potrf %8lld trsm %8lld upd %8lld inv %8lld tasks %5lld total %8lld | upd phases: issue %lld first %lld loop %lld store %lld\n", b, o[0], o[1], o[2], o[3],
              o[4], o[5], o[6], o[8], o[9], o[10], o[11]);
    }
    long long w = 0, tr = 0, up = 0, iv = 0, nt = 0, tot = 0;
    for (int b = n_ded; b < P.grid; ++b) { const long long* o = &h[(size_t)b * 16]; w += o[0]; tr += o[2]; up += o[3]; iv += o[4]; nt += o[5]; tot += o[6]; }
    if (P.trace) {
      std::vector<unsigned long long> tr((size_t)T * 16);
      CKC(cudaMemcpy(tr.data(), P.trace, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      const unsigned long long t0 = tr[0];
      fprintf(stderr, "[chol trace] us since potrf(0) start: k | P start, P end | solve in, X ready, step0..3 start, X published, diag upd done, subdiag upd done\n");
      fprintf(stderr, "[chol trace] potrf start times (us):");
      for (int k = 0; k < T; ++k) fprintf(stderr, " %.0f", (double)(tr[(size_t)k * 16] - t0) * 1e-3);
      fprintf(stderr, "\n");
      const bool every = atoi(getenv("STBA_CHOL_PROF")) >= 2;
      for (int k = 0; k < T; k += (every || k < 4 || k + 5 > T) ? 1 : 6) {
        fprintf(stderr, "[chol trace] %2d |", k);
        for (int s = 0; s <= 10; ++s) fprintf(stderr, " %8.1f%s", tr[(size_t)k * 16 + s] ? (double)(tr[(size_t)k * 16 + s] - t0) * 1e-3 : -1.0, (s == 1) ? " |" : "");
        fprintf(stderr, "\n");
      }
    }
#ifdef STBA_CHOL_TIMING
    {
      static bool tick_set = false;
      if (!tick_set) { int v = getenv("STBA_TICK_K0") ? atoi(getenv("STBA_TICK_K0")) : -1; cudaMemcpyToSymbol(g_tick_k0, &v, sizeof(int)); tick_set = true; }
      long long hc[64];
      cudaMemcpyFromSymbol(hc, g_potrf_clk, sizeof(hc));
      fprintf(stderr, "[potrf128 clocks, last panel]");
      for (int i = 1; i <= 16; ++i) fprintf(stderr, " %d:%lld", i, hc[i] - hc[i - 1]);
      for (int b = 0; b < 4; ++b) fprintf(stderr, " potf2_%d:%lld", b, hc[20 + b] - hc[b ? 1 + 3 * b : 1]);
      fprintf(stderr, "\n");
    }
#endif
    const double nw = std::max(1, P.grid - n_ded);
    fprintf(stderr, "[chol dag2] workers (mean cycles): wait %.0f trsm %.0f upd %.0f inv %.0f tasks %.1f total %.0f\n", w / nw, tr / nw, up / nw,
            iv / nw, nt / nw, tot / nw);
  }
  return STBA_OK;
}

int chol_factor_solve(CholWorkspace& ws, double* S, int n, int ld, double* rhs, int* dev_info, cudaStream_t stream, int* n_launches) {
  if (n <= 0) return STBA_OK;
  if (ld % 2 || ld < n + 1) return STBA_ERR_UNSUPPORTED;   // 16-byte cp.async rows; room for the rhs row
  {
    // small systems: one launch of one CTA (stba_chol_small.cuh); STBA_CHOL_SMALL_N=0 sends everything to the DAG kernel
    static const int small_n = std::min(CS_MAXN, getenv("STBA_CHOL_SMALL_N") ? atoi(getenv("STBA_CHOL_SMALL_N")) : CS_DEFAULT_N);
    if (n <= small_n) {
      {
        // the shared-memory opt-in is a per-device attribute
        static std::mutex mu;
        static bool done[64] = {};
        int dev = 0;
        CKC(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= 64 || !done[dev]) {
          CKC(cudaFuncSetAttribute(k_chol_small, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM + CS_SMEM_CLK));
          if (dev >= 0 && dev < 64) done[dev] = true;
        }
      }
      k_chol_small<<<1, CS_THREADS, CS_SMEM + CS_SMEM_CLK, stream>>>(S, ld, n, rhs, dev_info);
      CKC(cudaGetLastError());
#ifdef STBA_CS_TIMING
      {
        cudaStreamSynchronize(stream);
        long long h[16];
        cudaMemcpyFromSymbol(h, g_cs_clk, sizeof(h));
        fprintf(stderr, "[chol small n=%d] cycles of warp 0: init+potf2 %lld | panel solve+barrier %lld | (unused %lld %lld) tile00 %lld potf2 %lld tiles %lld barrier %lld | bwd products+barrier %lld bwd solve+barrier %lld | panel solve: loads %lld chain %lld stores %lld\n",
                n, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10], h[11], h[12]);
      }
#endif
      if (n_launches) *n_launches += 1;
      return STBA_OK;
    }
  }
  CholPlan* P = ws.plan;
  if (!P || P->S != S || P->n != n || P->ld != ld || P->rhs != rhs || P->info != dev_info) {
    destroy_plan(P);
    ws.plan = P = new CholPlan();
    P->S = S; P->n = n; P->ld = ld; P->rhs = rhs; P->info = dev_info;
    const int T = (n + NB - 1) / NB;
    P->rest64 = getenv("STBA_SYRK_TM") ? atoi(getenv("STBA_SYRK_TM")) == 64 : false;
    P->dag = getenv("STBA_CHOL_GRAPH") == nullptr;
    P->dag_version = (getenv("STBA_CHOL_DAG1") || (n + 1 + NB - 1) / NB > MAX_TR) ? 1 : 2;
    if (P->dag && P->dag_version == 2) {
      P->pool_stream = stream;
      CKC(cudaMallocAsync((void**)&P->Linv, (size_t)T * NB * NB * sizeof(double), stream));
      CKC(cudaMemsetAsync(P->Linv, 0, (size_t)T * NB * NB * sizeof(double), stream));   // upper triangles stay zero forever
      CKC(cudaMallocAsync((void**)&P->ybuf, (size_t)T * NB * sizeof(double), stream));
      CKC(cudaMallocAsync((void**)&P->flags, (size_t)T * sizeof(int), stream));
    } else {
      CKC(cudaMalloc(&P->Linv, (size_t)T * NB * NB * sizeof(double)));
      CKC(cudaMemset(P->Linv, 0, (size_t)T * NB * NB * sizeof(double)));   // upper triangles stay zero forever
      CKC(cudaMalloc(&P->ybuf, (size_t)T * NB * sizeof(double)));
      CKC(cudaMalloc(&P->flags, (size_t)T * sizeof(int)));
    }
    {
      static bool attr_set = false;
      if (!attr_set) { CKC(cudaFuncSetAttribute(k_trsv_bwd_all, cudaFuncAttributeMaxDynamicSharedMemorySize, TBA_SMEM)); attr_set = true; }
    }
    if (P->dag) { if ((P->dag_version == 2 ? build_dag2_plan(*P, stream) : build_dag_plan(*P)) != STBA_OK) return STBA_ERR_CUDA; }
    if (!P->dag) {
    {
      // tile lists: for step k, [panel rows | column k+1 strip | the rest], stored back to back
      std::vector<int2> h;
      P->off_strip.resize(T); P->off_rest.resize(T); P->off_panel.resize(T);
      P->n_strip.resize(T); P->n_rest.resize(T); P->n_panel.resize(T);
      for (int k = 0; k < T; ++k) {
        P->off_panel[k] = h.size();
        const int T64 = (n + 1 + 63) / 64, Tr = (n + 1 + NB - 1) / NB;             // row tiles include the rhs row n
        for (int i = 2 * (k + 1); i < T64; ++i) h.push_back(make_int2(i, 0));          // 64-row tiles below block k
        P->n_panel[k] = (int)(h.size() - P->off_panel[k]);
        P->off_strip[k] = h.size();
        if (k + 1 < T) for (int i = 2 * (k + 1); i < T64; ++i) h.push_back(make_int2(i, k + 1));
        P->n_strip[k] = (int)(h.size() - P->off_strip[k]);
        P->off_rest[k] = h.size();
        for (int j = k + 2; j < T; ++j) {
          if (P->rest64) for (int i = 2 * j; i < T64; ++i) h.push_back(make_int2(i, j));
          else for (int i = j; i < Tr; ++i) h.push_back(make_int2(i, j));
        }
        P->n_rest[k] = (int)(h.size() - P->off_rest[k]);
      }
      CKC(cudaMalloc(&P->tiles, std::max<size_t>(h.size(), 1) * sizeof(int2)));
      CKC(cudaMemcpy(P->tiles, h.data(), h.size() * sizeof(int2), cudaMemcpyHostToDevice));
    }
    CKC(cudaStreamCreateWithFlags(&P->side, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&P->inv, cudaStreamNonBlocking));
    CKC(cudaFuncSetAttribute(k_gemm_nt<MODE_SYRK, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem(128)));
    CKC(cudaFuncSetAttribute(k_gemm_nt<MODE_SYRK, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem(64)));
    CKC(cudaFuncSetAttribute(k_potrf128, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_SMEM));
    CKC(cudaFuncSetAttribute(k_inv_offdiag, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_SMEM));
    CKC(cudaFuncSetAttribute(k_trsm_sub, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
    CKC(cudaFuncSetAttribute(k_trsv_bwd_all, cudaFuncAttributeMaxDynamicSharedMemorySize, TBA_SMEM));
    // capture the static schedule once
    cudaGraph_t graph = nullptr;
    CKC(cudaStreamSynchronize(stream));
    CKC(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    const int r = enqueue(*P, stream, true);
    cudaError_t ce = cudaStreamEndCapture(stream, &graph);
    if (r != STBA_OK || ce != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      fprintf(stderr, "[stba] Cholesky graph capture failed (%d, %s)\n", r, cudaGetErrorString(ce));
      return STBA_ERR_CUDA;
    }
    ce = cudaGraphInstantiate(&P->exec, graph, 0);
    cudaGraphDestroy(graph);
    CKC(ce);
    }
  }
  if (P->dag) {
    const int r = P->dag_version == 2 ? run_dag2(*P, stream) : run_dag(*P, stream);
    if (r != STBA_OK) return r;
    if (n_launches) *n_launches += P->launches;
    return STBA_OK;
  }
  CKC(cudaGraphLaunch(P->exec, stream));
#ifdef STBA_CHOL_TIMING
  {
    cudaStreamSynchronize(stream);
    long long h[64];
    cudaMemcpyFromSymbol(h, g_potrf_clk, sizeof(h));
    long long ht[16];
    cudaMemcpyFromSymbol(ht, g_trsm_clk, sizeof(ht));
    fprintf(stderr, "[trsm_sub clocks]");
    for (int i = 1; i <= 9; ++i) fprintf(stderr, " %d:%lld", i, ht[i] - ht[i - 1]);
    fprintf(stderr, "\n[potrf128 clocks]");
    for (int i = 1; i <= 16; ++i) fprintf(stderr, " %d:%lld", i, h[i] - h[i - 1]);
    for (int b = 0; b < 4; ++b) fprintf(stderr, " potf2_%d:%lld", b, h[20 + b] - h[b ? 1 + 3 * b : 1]);
    fprintf(stderr, "\n");
  }
#endif
  if (n_launches) *n_launches += P->launches;
  return STBA_OK;
}

// =================================================================================================
// Split factorisation for N GPUs that hold the SAME matrix (the reduced camera system after the all-reduce).
//   S = [A Bt; B C], A = the first m block columns.
//   1. every rank: DAG factorisation restricted to the first m block columns, ALL rows (L_A, X = B L_A^-T, the
//      right-hand-side row rides along) — k_chol_dag2 with T = m;
//   2. the Schur complement C' = C - X X^T is a list of independent 128 x 128 x (128 m) tile updates: rank r
//      computes tiles r, r + N, ... (k_upd_list), packs them, and the ranks exchange them (grouped ncclBroadcast);
//   3. every rank: DAG factorisation of C' (a sub-matrix of S with the same leading dimension, the right-hand-side
//      row still in row n), then ONE backward substitution over the whole factor.
// The factorisation chain (one 128-column panel after the other, ~55 us each) stays replicated — it is a latency
// chain, not work — the n^3/3-sized bulk between the two halves is what scales.  With one rank the same code runs
// without the exchange (STBA_CHOL_SPLIT=1: tests).
struct SplitPlan {
  double* S = nullptr; double* rhs = nullptr; int* info = nullptr;
  int n = 0, ld = 0, m = 0, T = 0, Tr = 0;
  cudaStream_t stream = nullptr;
  double *Linv = nullptr, *ybuf = nullptr, *pack = nullptr;
  int* flags = nullptr;
  int *tiles1 = nullptr, *tiles3 = nullptr, *dflags1 = nullptr, *dflags3 = nullptr, *upd_tiles = nullptr;
  int n_tiles1 = 0, n_tiles3 = 0, units1 = 0, units3 = 0, n_upd = 0, grid = 0, W = 2, G = 8, D = 1;
  size_t n_dflags1 = 0, n_dflags3 = 0;
  int nranks = 1, rank = 0;
  std::vector<int> first;       // first[r] .. first[r + 1]: tiles of rank r in upd_tiles
};
static void destroy_split(SplitPlan* p) {
  if (!p) return;
  cudaStream_t st = p->stream;
  void* bufs[] = {p->Linv, p->ybuf, p->pack, p->flags, p->tiles1, p->tiles3, p->dflags1, p->dflags3, p->upd_tiles};
  for (void* b : bufs) if (b) cudaFreeAsync(b, st);
  delete p;
}
SplitWorkspace::~SplitWorkspace() { reset(); }
void SplitWorkspace::reset() { destroy_split(plan); plan = nullptr; }

// tile list + column starts + unit count of a DAG factorisation of `T` block columns over `Tr` tile rows (= build_dag2_plan)
static int split_tables(int T, int Tr, int R64, int D, cudaStream_t stream, int** d_tiles, int* n_tiles, int* units_out, int** dflags, size_t* n_dflags) {
  std::vector<int> tiles, col_start(T + 1, 0);
  long long units = T;
  for (int j = 0; j < T; ++j) {
    col_start[j] = (int)tiles.size();
    for (int i = j; i < Tr; ++i) {
      tiles.push_back(i | (j << 16));
      units += std::max(std::min(j, i - D) - ((D == 1 && i - j == 1) ? 1 : 0), 0);
    }
  }
  col_start[T] = (int)tiles.size();
  for (int h = 0; h < R64; ++h) units += std::max(0, std::min((h >> 1) - D, T));
  *n_tiles = (int)tiles.size();
  *units_out = (int)units;
  std::vector<int> both(tiles);
  both.insert(both.end(), col_start.begin(), col_start.end());
  CKC(cudaMallocAsync((void**)d_tiles, both.size() * sizeof(int), stream));
  CKC(cudaMemcpyAsync(*d_tiles, both.data(), both.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
  CKC(cudaStreamSynchronize(stream));      // `both` is a local
  *n_dflags = (size_t)F2_ARR + 4 * (size_t)T + 2 * Tr + R64 + 2 * (size_t)Tr * T;
  CKC(cudaMallocAsync((void**)dflags, *n_dflags * sizeof(int), stream));
  return STBA_OK;
}

static int split_launch_dag(const SplitPlan& P, double* S, int n, int T, double* Linv, int* dflags, size_t n_dflags, int* d_tiles, int n_tiles, int units) {
  CKC(cudaMemsetAsync(dflags, 0, n_dflags * sizeof(int), P.stream));
  Dag2Params dp;
  dp.S = S; dp.ld = P.ld; dp.n = n; dp.n_rows = n + 1; dp.T = T; dp.Tr = (n + 1 + NB - 1) / NB; dp.R64 = (n + 1 + 63) / 64;
  dp.Linv = Linv; dp.info = P.info; dp.flags = dflags;
  dp.total_units = units; dp.W = P.W; dp.G = P.G; dp.D = P.D;
  dp.tiles = d_tiles; dp.col_start = d_tiles + n_tiles; dp.n_tiles = n_tiles;
  dp.prof = nullptr; dp.trace = nullptr;
  void* args[] = {&dp};
  CKC(cudaLaunchCooperativeKernel((const void*)k_chol_dag2, dim3(P.grid), dim3(DAG_THREADS), args, DAG_SMEM, P.stream));
  return STBA_OK;
}

int chol_factor_solve_split(SplitWorkspace& ws, double* S, int n, int ld, double* rhs, int* dev_info, cudaStream_t stream, void* nccl_comm, int rank,
                            int nranks, int* n_launches) {
  if (n <= 0) return STBA_OK;
  if (ld % 2 || ld < n + 1) return STBA_ERR_UNSUPPORTED;
  const int T = (n + NB - 1) / NB, Tr = (n + 1 + NB - 1) / NB;
  if (T < 8 || Tr > MAX_TR) return STBA_ERR_UNSUPPORTED;      // (callers fall back to chol_factor_solve)
  SplitPlan* P = ws.plan;
  if (!P || P->S != S || P->n != n || P->ld != ld || P->rhs != rhs || P->info != dev_info || P->nranks != nranks || P->rank != rank) {
    destroy_split(P);
    ws.plan = P = new SplitPlan();
    P->S = S; P->n = n; P->ld = ld; P->rhs = rhs; P->info = dev_info; P->stream = stream; P->T = T; P->Tr = Tr; P->nranks = nranks; P->rank = rank;
    P->m = T / 2 + 1;
    if (const char* e = getenv("STBA_CHOL_SPLIT_M")) P->m = std::max(2, std::min(T - 2, atoi(e)));
    int sms = 0, dev = 0;
    CKC(cudaGetDevice(&dev));
    CKC(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    P->grid = sms;
    if (P->grid < 4) return STBA_ERR_UNSUPPORTED;
    CKC(cudaMallocAsync((void**)&P->Linv, (size_t)T * NB * NB * sizeof(double), stream));
    CKC(cudaMemsetAsync(P->Linv, 0, (size_t)T * NB * NB * sizeof(double), stream));
    CKC(cudaMallocAsync((void**)&P->ybuf, (size_t)T * NB * sizeof(double), stream));
    CKC(cudaMallocAsync((void**)&P->flags, (size_t)T * sizeof(int), stream));
    const int m = P->m, n3 = n - m * NB, T3 = (n3 + NB - 1) / NB, Tr3 = (n3 + 1 + NB - 1) / NB;
    if (split_tables(m, Tr, (n + 1 + 63) / 64, P->D, stream, &P->tiles1, &P->n_tiles1, &P->units1, &P->dflags1, &P->n_dflags1) != STBA_OK) return STBA_ERR_CUDA;
    if (split_tables(T3, Tr3, (n3 + 1 + 63) / 64, P->D, stream, &P->tiles3, &P->n_tiles3, &P->units3, &P->dflags3, &P->n_dflags3) != STBA_OK) return STBA_ERR_CUDA;
    // trailing tiles, dealt round-robin in column-major order (every rank gets the same mix of near and far tiles);
    // stored rank by rank so that each rank's tiles are one contiguous piece of the pack buffer
    std::vector<std::vector<int>> mine(nranks);
    int cnt = 0;
    for (int j = m; j < T; ++j)
      for (int i = j; i < Tr; ++i) mine[(cnt++) % nranks].push_back(i | (j << 16));
    std::vector<int> all;
    P->first.assign(nranks + 1, 0);
    for (int r = 0; r < nranks; ++r) { P->first[r] = (int)all.size(); all.insert(all.end(), mine[r].begin(), mine[r].end()); }
    P->first[nranks] = P->n_upd = (int)all.size();
    CKC(cudaMallocAsync((void**)&P->upd_tiles, std::max<size_t>(all.size(), 1) * sizeof(int), stream));
    CKC(cudaMemcpyAsync(P->upd_tiles, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    CKC(cudaStreamSynchronize(stream));
    if (nranks > 1) CKC(cudaMallocAsync((void**)&P->pack, (size_t)P->n_upd * NB * NB * sizeof(double), stream));
    static bool attr_set = false;
    if (!attr_set) {
      CKC(cudaFuncSetAttribute(k_chol_dag2, cudaFuncAttributeMaxDynamicSharedMemorySize, DAG_SMEM));
      CKC(cudaFuncSetAttribute(k_upd_list, cudaFuncAttributeMaxDynamicSharedMemorySize, DAG_SMEM));
      CKC(cudaFuncSetAttribute(k_trsv_bwd_all, cudaFuncAttributeMaxDynamicSharedMemorySize, TBA_SMEM));
      attr_set = true;
    }
  }
  const int m = P->m, K = m * NB, n3 = n - K, T3 = (n3 + NB - 1) / NB;
  // 1. first m block columns, all rows
  k_put_row<<<(n + 255) / 256, 256, 0, stream>>>(S, ld, n, rhs);
  if (split_launch_dag(*P, S, n, m, P->Linv, P->dflags1, P->n_dflags1, P->tiles1, P->n_tiles1, P->units1) != STBA_OK) return STBA_ERR_CUDA;
  // 2. this rank's tiles of the Schur complement, then everybody's
  const int t0 = P->first[rank], nt = P->first[rank + 1] - t0;
  if (nt) k_upd_list<<<std::min(P->grid, nt), DAG_THREADS, DAG_SMEM, stream>>>(S, ld, n + 1, n, K, P->upd_tiles + t0, nt);
  int launches = 3;
  if (nranks > 1) {
    if (nt) k_tiles_pack<<<nt, 256, 0, stream>>>(S, ld, n + 1, n, P->upd_tiles, t0, P->pack, 0);
    ncclComm_t comm = static_cast<ncclComm_t>(nccl_comm);
    if (ncclGroupStart() != ncclSuccess) return STBA_ERR_COMM;
    for (int r = 0; r < nranks; ++r) {
      const size_t cnt = (size_t)(P->first[r + 1] - P->first[r]) * NB * NB;
      double* b = P->pack + (size_t)P->first[r] * NB * NB;
      if (cnt && ncclBroadcast(b, b, cnt, ncclDouble, r, comm, stream) != ncclSuccess) return STBA_ERR_COMM;
    }
    if (ncclGroupEnd() != ncclSuccess) return STBA_ERR_COMM;
    for (int r = 0; r < nranks; ++r) {
      const int c = P->first[r + 1] - P->first[r];
      if (r != rank && c) k_tiles_pack<<<c, 256, 0, stream>>>(S, ld, n + 1, n, P->upd_tiles, P->first[r], P->pack, 1);
    }
    launches += 1 + nranks;
  }
  // 3. the trailing block (same leading dimension, the right-hand-side row is still row n), then one backward substitution
  double* S3 = S + (size_t)K * ld + K;
  if (split_launch_dag(*P, S3, n3, T3, P->Linv + (size_t)m * NB * NB, P->dflags3, P->n_dflags3, P->tiles3, P->n_tiles3, P->units3) != STBA_OK) return STBA_ERR_CUDA;
  k_get_row<<<(n + 255) / 256, 256, 0, stream>>>(S, ld, n, P->ybuf);
  CKC(cudaMemsetAsync(rhs, 0xFF, (size_t)n * sizeof(double), stream));      // NaN = "not computed yet" (k_trsv_bwd_all, poll_data)
  for (int b0 = 0; b0 < T; b0 += substitution_chunk())
    k_trsv_bwd_all<<<std::min(T - b0, substitution_chunk()), TBA_THREADS, TBA_SMEM, stream>>>(S, ld, n, T, P->Linv, P->ybuf, rhs, P->flags, dev_info, b0, 1);
  CKC(cudaGetLastError());
  if (n_launches) *n_launches += launches + 3;
  return STBA_OK;
}

// Both triangular solves on a finished lower Cholesky factor (e.g. cusolverDnDpotrf's): 128 x 128 diagonal
// inverses (T CTAs side by side), then the one-launch forward and backward substitutions.  cusolverDnDpotrs
// runs two latency-bound trsv kernels (0.40 + 0.58 ms at n = 5988); this is three launches.
struct SolvePlan {
  int n = 0;
  double* Linv = nullptr;
  double* ybuf = nullptr;
  int* flags = nullptr;
  cudaStream_t stream = nullptr;      // allocations are stream-ordered (pooled): no cudaMalloc / cudaFree stalls per engine
};
static void destroy_solve(SolvePlan* p) {
  if (!p) return;
  if (p->Linv) cudaFreeAsync(p->Linv, p->stream);
  if (p->ybuf) cudaFreeAsync(p->ybuf, p->stream);
  if (p->flags) cudaFreeAsync(p->flags, p->stream);
  delete p;
}
SolveWorkspace::~SolveWorkspace() { reset(); }
void SolveWorkspace::reset() { destroy_solve(plan); plan = nullptr; }

int chol_solve_with_factor(SolveWorkspace& ws, const double* S, int n, int ld, double* rhs, int* dev_info, cudaStream_t stream,
                           int* n_launches) {
  if (n <= 0) return STBA_OK;
  if (ld % 2) return STBA_ERR_UNSUPPORTED;
  const int T = (n + NB - 1) / NB;
  SolvePlan* P = ws.plan;
  if (!P || P->n != n) {
    destroy_solve(P);
    ws.plan = P = new SolvePlan();
    P->n = n;
    P->stream = stream;
    CKC(cudaMallocAsync((void**)&P->Linv, (size_t)T * NB * NB * sizeof(double), stream));
    CKC(cudaMemsetAsync(P->Linv, 0, (size_t)T * NB * NB * sizeof(double), stream));
    CKC(cudaMallocAsync((void**)&P->ybuf, (size_t)T * NB * sizeof(double), stream));
    CKC(cudaMallocAsync((void**)&P->flags, 2 * (size_t)T * sizeof(int), stream));
    CKC(cudaFuncSetAttribute(k_inv128, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_SMEM));
    CKC(cudaFuncSetAttribute(k_trsv_fwd_all, cudaFuncAttributeMaxDynamicSharedMemorySize, TBA_SMEM));
    CKC(cudaFuncSetAttribute(k_trsv_bwd_all, cudaFuncAttributeMaxDynamicSharedMemorySize, TBA_SMEM));
  }
  CKC(cudaMemsetAsync(P->flags, 0, 2 * (size_t)T * sizeof(int), stream));
  k_inv128<<<T, PT, POTRF_SMEM, stream>>>(S, ld, n, P->Linv);
  for (int b0 = 0; b0 < T; b0 += substitution_chunk())
    k_trsv_fwd_all<<<std::min(T - b0, substitution_chunk()), TBA_THREADS, TBA_SMEM, stream>>>(S, ld, n, T, P->Linv, rhs, P->ybuf, P->flags, dev_info, b0);
  for (int b0 = 0; b0 < T; b0 += substitution_chunk())
    k_trsv_bwd_all<<<std::min(T - b0, substitution_chunk()), TBA_THREADS, TBA_SMEM, stream>>>(S, ld, n, T, P->Linv, P->ybuf, rhs, P->flags + T, dev_info, b0);
  CKC(cudaGetLastError());
  if (n_launches) *n_launches += 3;
  return STBA_OK;
}

}  // namespace stba
