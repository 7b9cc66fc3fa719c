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

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
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
#define TTICK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_trsm_clk[i] = clock64(); } while (0)
#else
#define TTICK(i) do {} while (0)
#endif
constexpr int TS_THREADS = 256;
constexpr int TS_ROWS = 64;
constexpr int TS_SMEM = (NB * LDS + 8 * NB * 8) * (int)sizeof(double);

__global__ void __launch_bounds__(TS_THREADS, 1)
k_trsm_sub(double* __restrict__ S, int ld, int n_rows, int k0, int nb, int row0, const double* __restrict__ Linv) {
  extern __shared__ __align__(16) double smem[];
  double* Ls = smem;                         // Ls[c * LDS + r]: L(r, c) below the diagonal blocks, Inv(r, c) inside them
  double* Xs = smem + NB * LDS;              // Xs[warp][col][8 rows]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  const int i0 = row0 + (int)blockIdx.x * TS_ROWS;
  TTICK(0);
  // Operands arrive in four cp.async groups, one per 32-column block b: the columns 32b..32b+31 of
  // the factor block (Inv_bb inside the diagonal block, L below it) and of the CTA's 64 rows.  Step
  // j only needs groups <= j, so the later groups land while the first steps compute.  One warp per
  // column, lanes along the rows: the index arithmetic is warp-uniform and cheap.
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    for (int c = 32 * b + warp; c < 32 * b + 32; c += TS_THREADS / 32) {
      double* dl = Ls + c * LDS;
      const bool col_ok = c < nb;
      // rows 32b .. 127 of column c, as pairs (r2, r2 + 1)
      for (int r2 = 32 * b + 2 * lane; r2 < NB; r2 += 64) {
        const bool in_diag = r2 < 32 * b + 32;
        const double* src = in_diag ? Linv + (size_t)c * NB + r2 : S + (size_t)(k0 + c) * ld + k0 + r2;
        if (col_ok && r2 >= c && r2 + 1 < nb) {
          cp_async16(dl + r2, src, true);
        } else {                       // diagonal / padding pairs: asynchronous too, so that no lane ever blocks its warp on a load
          if (col_ok && r2 >= c && r2 < nb) cp_async8(dl + r2, src); else dl[r2] = 0.0;
          if (col_ok && r2 + 1 >= c && r2 + 1 < nb) cp_async8(dl + r2 + 1, src + 1); else dl[r2 + 1] = 0.0;
        }
      }
      // the CTA's 64 rows of panel column c
      {
        const int r2 = 2 * lane, row = i0 + r2;
        double* dst = Xs + ((r2 >> 3) * NB + c) * 8 + (r2 & 7);
        if (col_ok && row + 1 < n_rows) {
          cp_async16(dst, S + (size_t)(k0 + c) * ld + row, true);
        } else {
          if (col_ok && row < n_rows) cp_async8(dst, S + (size_t)(k0 + c) * ld + row); else dst[0] = 0.0;
          dst[1] = 0.0;
        }
      }
    }
    cp_async_commit();
  }
  double* Xw = Xs + warp * NB * 8;
  const int row = i0 + warp * 8 + g;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    TTICK(1 + 2 * j);
    if (j == 0) cp_async_wait<3>(); else if (j == 1) cp_async_wait<2>(); else if (j == 2) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    TTICK(2 + 2 * j);
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

// ---- diagonal block: Cholesky + inverse of the factor, one CTA of 512 threads -----------------
constexpr int PT = 512;          // threads of the diagonal-block kernel
constexpr int PLD = NB + 1;      // odd leading dimension: column reads by consecutive lanes conflict-free
constexpr int TLD = 97;          // leading dimension of the 32 x 96 product scratch
constexpr int POTRF_SMEM = (NB * PLD + NB + 32 * TLD + 32 * 33) * (int)sizeof(double);
constexpr unsigned FULL = 0xffffffffu;
#ifdef STBA_CHOL_TIMING
__device__ long long g_potrf_clk[64];
#define TICK(i) do { if (threadIdx.x == 0) g_potrf_clk[i] = clock64(); } while (0)
#else
#define TICK(i) do {} while (0)
#endif

__global__ void __launch_bounds__(PT, 1)
k_potrf128(double* __restrict__ S, int ld, int k0, int nb, double* __restrict__ Linv, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  double* D = sm;                 // D[c * PLD + r]: lower triangle + diagonal = the factor L;
                                  // strict upper triangle = the inverse, transposed: X(r,c), r > c, at D[r * PLD + c]
  double* xd = sm + NB * PLD;     // diagonal of the inverse
  double* Tm = xd + NB;           // Tm[rr * TLD + cc]: 32 x (32 bi) product scratch
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  TICK(0);
  for (int e = tid; e < NB * NB; e += PT) {
    const int r = e % NB, c = e / NB;
    D[c * PLD + r] = (r < nb && c < nb && r >= c) ? S[(size_t)(k0 + c) * ld + k0 + r] : (r == c ? 1.0 : 0.0);
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

// Completes Linv_kk: off-diagonal 32 x 32 blocks X_ib = -X_ii (sum_{p=b}^{i-1} L_ip X_pb), from the
// factor (in S) and the diagonal inverse blocks k_potrf128 stored.  One CTA; runs concurrently with
// the panel solve / trailing update of the same step.
__global__ void __launch_bounds__(PT, 1)
k_inv_offdiag(const double* __restrict__ S, int ld, int k0, int nb, double* __restrict__ Linv) {
  extern __shared__ __align__(16) double sm[];
  double* D = sm;
  double* xd = sm + NB * PLD;
  double* Tm = xd + NB;
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += PT) {
    const int r = e % NB, c = e / NB;
    if (r >= c) D[c * PLD + r] = (r < nb && c < nb) ? S[(size_t)(k0 + c) * ld + k0 + r] : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();       // (lower part written before the upper slots of the same columns are filled)
  for (int e = tid; e < NB * 32; e += PT) {
    const int c = e / 32, r = (c & ~31) + (e % 32);
    if (r < c) continue;
    const double v = Linv[(size_t)c * NB + r];
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
               double* x, int* flags, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  double* Lb = sm;                          // Lb[c * NB + r] = L(k0 + r, j0 + c)
  double* U = Lb + NB * NB;                 // U[c (c + 1) / 2 + r] = Linv_jj(c, r), r <= c
  double* yj = U + NB * (NB + 1) / 2;
  double* xk = yj + NB;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int j = T - 1 - (int)blockIdx.x, j0 = j * NB;
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
    if (t == 0) {
      long long spins = 0;
      while (ld_acquire(flags + k) == 0) {
        if (++spins > (1ll << 26)) { atomicCAS(info, 0, -1); break; }    // never hang the device
      }
    }
    __syncthreads();
    if (t < NB) xk[t] = (t < nbk) ? __ldcg(x + k0 + t) : 0.0;
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
    if (h == 0 && j0 + r < n) x[j0 + r] = (s0 + s1) + xk[r];
  }
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
               double* y, int* flags, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  double* Lb = sm;                          // Lb[c * NB + r] = L(j0 + r, k0 + c)
  double* U = Lb + NB * NB;                 // U[c * NB - c (c - 1) / 2 + (r - c)] = Linv_jj(r, c), r >= c
  double* bj = U + NB * (NB + 1) / 2;
  double* yk = bj + NB;
  const int t = threadIdx.x;
  const int j = (int)blockIdx.x, j0 = j * NB, nbj = min(NB, n - j0);
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
#define CKC(call)                                                                                           \
  do {                                                                                                      \
    cudaError_t e_ = (call);                                                                                \
    if (e_ != cudaSuccess) {                                                                                \
      fprintf(stderr, "[stba] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return STBA_ERR_CUDA;                                                                                 \
    }                                                                                                       \
  } while (0)

}  // namespace

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
  cudaStream_t side = nullptr, inv = nullptr;
  std::vector<cudaEvent_t> events;
  int launches = 0;
  bool rest64 = false;        // bulk trailing update in 64-row tiles, two CTAs per SM
  std::vector<size_t> off_strip, off_rest, off_panel;
  std::vector<int> n_strip, n_rest, n_panel;
};

static void destroy_plan(CholPlan* p) {
  if (!p) return;
  if (p->exec) cudaGraphExecDestroy(p->exec);
  if (p->Linv) cudaFree(p->Linv);
  if (p->ybuf) cudaFree(p->ybuf);
  if (p->tiles) cudaFree(p->tiles);
  if (p->flags) cudaFree(p->flags);
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
  k_trsv_bwd_all<<<T, TBA_THREADS, TBA_SMEM, main>>>(S, ld, n, T, P.Linv, P.ybuf, P.rhs, P.flags, P.info);
  ++P.launches;
  CKC(cudaGetLastError());
  return STBA_OK;
}

int chol_factor_solve(CholWorkspace& ws, double* S, int n, int ld, double* rhs, int* dev_info, cudaStream_t stream, int* n_launches) {
  if (n <= 0) return STBA_OK;
  if (ld % 2 || ld < n + 1) return STBA_ERR_UNSUPPORTED;   // 16-byte cp.async rows; room for the rhs row
  CholPlan* P = ws.plan;
  if (!P || P->S != S || P->n != n || P->ld != ld || P->rhs != rhs || P->info != dev_info) {
    destroy_plan(P);
    ws.plan = P = new CholPlan();
    P->S = S; P->n = n; P->ld = ld; P->rhs = rhs; P->info = dev_info;
    const int T = (n + NB - 1) / NB;
    P->rest64 = getenv("STBA_SYRK_TM") ? atoi(getenv("STBA_SYRK_TM")) == 64 : false;
    CKC(cudaMalloc(&P->Linv, (size_t)T * NB * NB * sizeof(double)));
    CKC(cudaMemset(P->Linv, 0, (size_t)T * NB * NB * sizeof(double)));   // upper triangles stay zero forever
    CKC(cudaMalloc(&P->ybuf, (size_t)T * NB * sizeof(double)));
    CKC(cudaMalloc(&P->flags, (size_t)T * sizeof(int)));
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
  k_trsv_fwd_all<<<T, TBA_THREADS, TBA_SMEM, stream>>>(S, ld, n, T, P->Linv, rhs, P->ybuf, P->flags, dev_info);
  k_trsv_bwd_all<<<T, TBA_THREADS, TBA_SMEM, stream>>>(S, ld, n, T, P->Linv, P->ybuf, rhs, P->flags + T, dev_info);
  CKC(cudaGetLastError());
  if (n_launches) *n_launches += 3;
  return STBA_OK;
}

}  // namespace stba
