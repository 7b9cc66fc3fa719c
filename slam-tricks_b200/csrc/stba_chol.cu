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

#include <algorithm>
#include <vector>

namespace stba {

namespace {

constexpr int NB = 128;          // block size
constexpr int KC = 16;           // k-chunk per pipeline stage
constexpr int STAGES = 4;
constexpr int LDS = NB + 4;      // smem leading dimension (doubles): (q*LDS + g) mod 16 distinct
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_SMEM = STAGES * 2 * KC * LDS * (int)sizeof(double);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
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
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_gemm_nt(double* __restrict__ S, int lda, int n_rows, int n_cols, int k0, int K, const double* __restrict__ Bmat, int ldb,
          const int2* __restrict__ tiles) {
  constexpr int WM = TM / 2;                      // warp tile rows (2 warps along M, 4 along N)
  constexpr int MT = WM / 8;                      // 8-row mma tiles per warp
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                              // [STAGES][KC][LDS]
  double* Bs = smem + STAGES * KC * LDS;
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
    double* as = As + stage * KC * LDS;
    double* bs = Bs + stage * KC * LDS;
#pragma unroll
    for (int p = 0; p < KC * (TM / 2) / GEMM_THREADS; ++p) {       // A: KC columns x TM rows
      const int piece = tid + p * GEMM_THREADS;
      const int kk = piece / (TM / 2), r2 = (piece % (TM / 2)) * 2;
      const int k = chunk * KC + kk;
      const bool kin = k < K;
      const int ra = i0 + r2;
      cp_async16(as + kk * LDS + r2, Ag + (size_t)(kin ? k : 0) * lda + (ra < n_rows ? ra : 0), kin && ra < n_rows);
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
    const double* as = As + (c % STAGES) * KC * LDS + wm * WM + g;
    const double* bs = Bs + (c % STAGES) * KC * LDS + wn * 32 + g;
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
      double a[MT], b[4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a[mt] = (MODE == MODE_SYRK) ? -as[(kk + q) * LDS + mt * 8] : as[(kk + q) * LDS + mt * 8];
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
k_potrf128(double* __restrict__ S, int ld, int k0, int nb, double* __restrict__ Linv, double* __restrict__ LinvT,
           int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  double* D = sm;                 // D[c * PLD + r]: lower triangle + diagonal = the factor L;
                                  // strict upper triangle = the inverse, transposed: X(r,c), r > c, at D[r * PLD + c]
  double* xd = sm + NB * PLD;     // diagonal of the inverse
  double* Tm = xd + NB;           // Tm[rr * TLD + cc]: 32 x (32 bi) product scratch
  double* Ib = Tm + 32 * TLD;     // Ib[i * 33 + c]: inverse of the current 32 x 32 diagonal sub-block
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
      double* rd = Tm + 64;            // reciprocals of the diagonal of the factor
      bool bad = false;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        double* col = cb + (j & 1) * 32;
        col[lane] = a[j];
        __syncwarp();
        const double d = col[j];
        if (!(d > 0.0) && !bad) { bad = true; if (lane == 0 && b0 + j < nb) atomicCAS(info, 0, k0 + b0 + j + 1); }
        // the dependent chain only needs 1/d (a_ik -= a_ij a_kj / d_j); the 1/sqrt(d) that turns
        // column j into the Cholesky column is computed off the chain
        const double t = a[j] * __drcp_rn(d);
#pragma unroll
        for (int k = j + 1; k < 32; ++k) a[k] = fma(-t, col[k], a[k]);   // lanes < k: unused upper values
        const double inv = rsqrt(d);
        a[j] = (lane == j) ? d * inv : a[j] * inv;
        if (lane == j) rd[j] = inv;
      }
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c <= lane) D[(b0 + c) * PLD + b0 + lane] = a[c];
      __syncwarp();
      // inverse of the 32 x 32 factor: lane = column; L(i,p) read back as broadcasts
      double x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
        for (int p = 0; p < i; ++p) {
          const double lip = D[(b0 + p) * PLD + b0 + i];
          if (p & 1) s1 = fma(-lip, x[p], s1); else s0 = fma(-lip, x[p], s0);
        }
        x[i] = (s0 + s1) * rd[i];
      }
      const int c = b0 + lane;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        Ib[i * 33 + lane] = x[i];                               // Inv(i, lane); zero above the diagonal
        if (i == lane) xd[c] = x[i];
        else if (i > lane) D[(b0 + i) * PLD + c] = x[i];        // X(b0+i, c) in its transposed slot
      }
    }
    __syncthreads();
    TICK(2 + (b0 / 32) * 3);
    // (2) rows below the sub-block: L_rows = A_rows Inv^T, four threads per row (8 columns each)
    const int below = NB - b0 - 32;
    double xr[8];
    const int r_loc = below ? tid % below : 0, part = below ? tid / below : 4;
    if (part < 4) {
      const int r = b0 + 32 + r_loc;
      double ar[32];
#pragma unroll
      for (int p = 0; p < 32; ++p) ar[p] = D[(b0 + p) * PLD + r];
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        double s = 0.0;
#pragma unroll
        for (int p = 0; p < 32; ++p) s = fma(ar[p], Ib[(part * 8 + cc) * 33 + p], s);    // Inv(c, p) = 0 for p > c
        xr[cc] = s;
      }
    }
    __syncthreads();
    if (part < 4) {
      const int r = b0 + 32 + r_loc;
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) D[(b0 + part * 8 + cc) * PLD + r] = xr[cc];
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
  // ---- off-diagonal blocks of the inverse, one block row at a time:
  //      X_ib = -X_ii (sum_{p=b}^{i-1} L_ip X_pb)
  for (int bi = 1; bi < 4; ++bi) {
    const int w = 32 * bi;
    for (int e = tid; e < 32 * w; e += PT) {
      const int rr = e % 32, cc = e / 32;          // T[rr][cc] = sum_{p=cc}^{w-1} L[w+rr][p] X[p][cc]
      double s = D[cc * PLD + w + rr] * xd[cc], s2 = 0.0, s3 = 0.0, s4 = 0.0;
      int p = cc + 1;
      for (; p + 3 < w; p += 4) {
        s = fma(D[p * PLD + w + rr], D[p * PLD + cc], s);
        s2 = fma(D[(p + 1) * PLD + w + rr], D[(p + 1) * PLD + cc], s2);
        s3 = fma(D[(p + 2) * PLD + w + rr], D[(p + 2) * PLD + cc], s3);
        s4 = fma(D[(p + 3) * PLD + w + rr], D[(p + 3) * PLD + cc], s4);
      }
      for (; p < w; ++p) s = fma(D[p * PLD + w + rr], D[p * PLD + cc], s);
      Tm[rr * TLD + cc] = (s + s2) + (s3 + s4);
    }
    __syncthreads();
    for (int e = tid; e < 32 * w; e += PT) {
      const int rr = e % 32, cc = e / 32;          // X[w+rr][cc] = -sum_{p<=rr} X_ii[rr][p] T[p][cc]
      double s = xd[w + rr] * Tm[rr * TLD + cc], s2 = 0.0, s3 = 0.0, s4 = 0.0;
      int p = 0;
      for (; p + 3 < rr; p += 4) {
        s = fma(D[(w + rr) * PLD + w + p], Tm[p * TLD + cc], s);
        s2 = fma(D[(w + rr) * PLD + w + p + 1], Tm[(p + 1) * TLD + cc], s2);
        s3 = fma(D[(w + rr) * PLD + w + p + 2], Tm[(p + 2) * TLD + cc], s3);
        s4 = fma(D[(w + rr) * PLD + w + p + 3], Tm[(p + 3) * TLD + cc], s4);
      }
      for (; p < rr; ++p) s = fma(D[(w + rr) * PLD + w + p], Tm[p * TLD + cc], s);
      D[(w + rr) * PLD + cc] = -((s + s2) + (s3 + s4));
    }
    __syncthreads();
  }
  TICK(15);
  for (int e = tid; e < NB * NB; e += PT) {
    const int r = e % NB, c = e / NB;
    double v = 0.0;
    if (r < nb && c < nb) v = r > c ? D[r * PLD + c] : (r == c ? xd[r] : 0.0);
    Linv[(size_t)c * NB + r] = v;                  // column-major
  }
  for (int e = tid; e < NB * NB; e += PT) {
    const int c = e % NB, r = e / NB;
    double v = 0.0;
    if (r < nb && c < nb) v = r > c ? D[r * PLD + c] : (r == c ? xd[r] : 0.0);
    LinvT[(size_t)r * NB + c] = v;                 // row-major copy = Linv^T column-major
  }
  TICK(16);
}

// ---- block triangular solves with the stored inverses -----------------------------------------
constexpr int TRSV_THREADS = 256;

// v = M w for a 128 x 128 column-major M (lower or upper triangular, zeros stored), all 256 threads
__device__ __forceinline__ double matvec128(const double* __restrict__ M, const double* w, double* scratch) {
  const int t = threadIdx.x, row = t & 127, half = t >> 7;
  double s = 0.0;
#pragma unroll 16
  for (int c = half * 64; c < half * 64 + 64; ++c) s = fma(M[(size_t)c * NB + row], w[c], s);
  scratch[t] = s;
  __syncthreads();
  return scratch[row] + scratch[row + 128];
}

// backward step k:  x_k = Linv_kk^T y_k (recomputed per CTA; CTA 0 stores), then the CTA's column
// block j < k:  y_j -= L_kj^T x_k   (one warp per column, lanes along the 128 rows)
__global__ void __launch_bounds__(TRSV_THREADS)
k_trsv_bwd(const double* __restrict__ S, int ld, int n, int k, const double* __restrict__ LinvT, double* __restrict__ y,
           double* __restrict__ x) {
  __shared__ double yk[NB], xk[NB], scratch[TRSV_THREADS];
  const int t = threadIdx.x, k0 = k * NB, lane = t & 31, warp = t >> 5;
  const int nb = min(NB, n - k0);
  if (t < NB) yk[t] = t < nb ? y[k0 + t] : 0.0;
  __syncthreads();
  const double v = matvec128(LinvT + (size_t)k * NB * NB, yk, scratch);
  if (t < NB) xk[t] = v;
  __syncthreads();
  if (blockIdx.x == 0) {
    if (t < nb) x[k0 + t] = v;
    return;
  }
  const int j0 = (blockIdx.x - 1) * NB;
  const double x0 = xk[lane], x1 = xk[lane + 32], x2 = xk[lane + 64], x3 = xk[lane + 96];
#pragma unroll 4
  for (int cc = 0; cc < 16; ++cc) {
    const int c = j0 + warp * 16 + cc;
    const double* col = S + (size_t)c * ld + k0 + lane;
    double acc = 0.0;
    if (lane < nb) acc = col[0] * x0;
    if (lane + 32 < nb) acc = fma(col[32], x1, acc);
    if (lane + 64 < nb) acc = fma(col[64], x2, acc);
    if (lane + 96 < nb) acc = fma(col[96], x3, acc);
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(FULL, acc, sft);
    if (lane == 0) y[c] -= acc;
  }
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
// last panel: the only row below the (possibly partial) last diagonal block is the rhs row:
// y_k = Linv_kk a^T
__global__ void __launch_bounds__(TRSV_THREADS)
k_trsm_last_row(double* __restrict__ S, int ld, int n, int k0, int nb, const double* __restrict__ Linv) {
  __shared__ double a[NB], scratch[TRSV_THREADS];
  const int t = threadIdx.x;
  if (t < NB) a[t] = t < nb ? S[(size_t)(k0 + t) * ld + n] : 0.0;
  __syncthreads();
  const double v = matvec128(Linv, a, scratch);
  if (t < nb) S[(size_t)(k0 + t) * ld + n] = v;
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
  double* LinvT = nullptr;    // transposes of the inverse blocks (backward solve reads them coalesced)
  int2* tiles = nullptr;      // device tile lists
  cudaGraphExec_t exec = nullptr;
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t> events;
  int launches = 0;
  std::vector<size_t> off_strip, off_rest, off_panel;
  std::vector<int> n_strip, n_rest, n_panel;
};

static void destroy_plan(CholPlan* p) {
  if (!p) return;
  if (p->exec) cudaGraphExecDestroy(p->exec);
  if (p->Linv) cudaFree(p->Linv);
  if (p->ybuf) cudaFree(p->ybuf);
  if (p->LinvT) cudaFree(p->LinvT);
  if (p->tiles) cudaFree(p->tiles);
  if (p->side) cudaStreamDestroy(p->side);
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
  const std::vector<size_t>&off_strip = P.off_strip, &off_rest = P.off_rest, &off_panel = P.off_panel;
  const std::vector<int>&n_strip = P.n_strip, &n_rest = P.n_rest, &n_panel = P.n_panel;
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
  // panel(k) = potrf + trsm on stream s
  auto panel = [&](int k, cudaStream_t s) {
    const int k0 = k * NB, nb = std::min(NB, n - k0);
    double* Li = P.Linv + (size_t)k * NB * NB;
    k_potrf128<<<1, PT, POTRF_SMEM, s>>>(S, ld, k0, nb, Li, P.LinvT + (size_t)k * NB * NB, P.info);
    ++P.launches;
    if (k == T - 1) {
      k_trsm_last_row<<<1, TRSV_THREADS, 0, s>>>(S, ld, n, k0, nb, Li);
      ++P.launches;
    } else if (n_panel[k]) {
      k_gemm_nt<MODE_TRSM, 64><<<n_panel[k], GEMM_THREADS, GEMM_SMEM, s>>>(S, ld, n_rows, n, k0, nb, Li, NB, P.tiles + off_panel[k]);
      ++P.launches;
    }
  };
  panel(0, main);
  for (int k = 0; k + 1 < T; ++k) {
    const int k0 = k * NB;
    // strip update of block column k+1, then its panel (look-ahead: on the side stream)
    k_gemm_nt<MODE_SYRK, 64><<<n_strip[k], GEMM_THREADS, GEMM_SMEM, main>>>(S, ld, n_rows, n, k0, NB, nullptr, ld, P.tiles + off_strip[k]);
    ++P.launches;
    if (lookahead && n_rest[k]) {
      cudaEvent_t e1 = next_event(), e2 = next_event();
      CKC(cudaEventRecord(e1, main));
      CKC(cudaStreamWaitEvent(P.side, e1, 0));
      panel(k + 1, P.side);
      CKC(cudaEventRecord(e2, P.side));
      k_gemm_nt<MODE_SYRK, 128><<<n_rest[k], GEMM_THREADS, GEMM_SMEM, main>>>(S, ld, n_rows, n, k0, NB, nullptr, ld, P.tiles + off_rest[k]);
      ++P.launches;
      CKC(cudaStreamWaitEvent(main, e2, 0));
    } else {
      if (n_rest[k]) {
        k_gemm_nt<MODE_SYRK, 128><<<n_rest[k], GEMM_THREADS, GEMM_SMEM, main>>>(S, ld, n_rows, n, k0, NB, nullptr, ld, P.tiles + off_rest[k]);
        ++P.launches;
      }
      panel(k + 1, main);
    }
  }
  // y = L^-1 rhs now sits in row n; backward substitution into rhs
  k_get_row<<<(n + 255) / 256, 256, 0, main>>>(S, ld, n, P.ybuf);
  ++P.launches;
  for (int k = T - 1; k >= 0; --k) {
    k_trsv_bwd<<<k + 1, TRSV_THREADS, 0, main>>>(S, ld, n, k, P.LinvT, P.ybuf, P.rhs);
    ++P.launches;
  }
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
    CKC(cudaMalloc(&P->Linv, (size_t)T * NB * NB * sizeof(double)));
    CKC(cudaMalloc(&P->ybuf, (size_t)T * NB * sizeof(double)));
    CKC(cudaMalloc(&P->LinvT, (size_t)T * NB * NB * sizeof(double)));
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
        for (int j = k + 2; j < T; ++j)
          for (int i = j; i < Tr; ++i) h.push_back(make_int2(i, j));
        P->n_rest[k] = (int)(h.size() - P->off_rest[k]);
      }
      CKC(cudaMalloc(&P->tiles, std::max<size_t>(h.size(), 1) * sizeof(int2)));
      CKC(cudaMemcpy(P->tiles, h.data(), h.size() * sizeof(int2), cudaMemcpyHostToDevice));
    }
    CKC(cudaStreamCreateWithFlags(&P->side, cudaStreamNonBlocking));
    CKC(cudaFuncSetAttribute(k_gemm_nt<MODE_SYRK, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    CKC(cudaFuncSetAttribute(k_gemm_nt<MODE_SYRK, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    CKC(cudaFuncSetAttribute(k_gemm_nt<MODE_TRSM, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    CKC(cudaFuncSetAttribute(k_potrf128, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_SMEM));
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
    fprintf(stderr, "[potrf128 clocks]");
    for (int i = 1; i <= 16; ++i) fprintf(stderr, " %d:%lld", i, h[i] - h[i - 1]);
    fprintf(stderr, "\n");
  }
#endif
  if (n_launches) *n_launches += P->launches;
  return STBA_OK;
}

}  // namespace stba
