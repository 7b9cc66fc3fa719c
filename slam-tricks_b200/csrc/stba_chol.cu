// Own dense reduced-camera solve: blocked right-looking Cholesky (fp64, lower, column-major, in
// place) + block triangular solves.  The trailing-matrix update and the panel solve run on the
// FP64 tensor pipe (mma.sync.m8n8k4.f64 = DMMA); see DESIGN.md §3.5.
//
//   for each 128-wide block column k:
//     k_potrf128     one CTA: Cholesky of the diagonal block in shared memory + explicit inverse
//                    of the triangular factor (Linv_kk, kept in a side buffer for the solves)
//     k_gemm_nt<TRSM> panel:   L_ik = A_ik Linv_kk^T                (DMMA, one CTA per row tile)
//     k_gemm_nt<SYRK> trailing: A_ij -= L_ik L_jk^T for i >= j > k   (DMMA, one CTA per tile)
//   forward / backward substitution by 128-blocks with the stored Linv_kk (matrix-vector only).
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

// One 128x128 output tile:  acc = A[i0.., k0..k0+K) * B[j0.., ...)^T  (both "row x k" panels stored
// column-major), then
//   MODE_SYRK: C[i0.., j0..] -= acc            A = B = the factored panel of S, ldb = lda
//   MODE_TRSM: A[i0.., k0..k0+K) = acc         B = Linv (K x K, ld NB), j0 = 0
// tiles[] lists (i-tile, j-tile) pairs; rows >= n_rows / cols >= n_cols are masked.
template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_gemm_nt(double* __restrict__ S, int lda, int n_rows, int n_cols, int k0, int K, const double* __restrict__ Bmat, int ldb,
          const int2* __restrict__ tiles) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                              // [STAGES][KC][LDS]
  double* Bs = smem + STAGES * KC * LDS;
  const int2 tile = tiles[blockIdx.x];
  const int i0 = tile.x * NB, j0 = (MODE == MODE_TRSM) ? 0 : tile.y * NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;        // warp tile 64 (M) x 32 (N)
  const double* Ag = S + (size_t)k0 * lda;        // panel columns k0..k0+K
  const double* Bg = (MODE == MODE_TRSM) ? Bmat : S + (size_t)k0 * lda;
  const int b_rows = (MODE == MODE_TRSM) ? K : n_rows;

  double acc[8][4][2];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  const int n_chunks = (K + KC - 1) / KC;
  auto load_stage = [&](int chunk, int stage) {
    // KC columns x 128 rows = KC*64 16-byte pieces per operand; 256 threads -> KC/4 pieces each
    double* as = As + stage * KC * LDS;
    double* bs = Bs + stage * KC * LDS;
#pragma unroll
    for (int p = 0; p < KC * 64 / GEMM_THREADS; ++p) {
      const int piece = tid + p * GEMM_THREADS;
      const int kk = piece >> 6, r2 = (piece & 63) * 2;
      const int k = chunk * KC + kk;
      const bool kin = k < K;
      const int ra = i0 + r2, rb = j0 + r2;
      cp_async16(as + kk * LDS + r2, Ag + (size_t)(kin ? k : 0) * lda + (ra < n_rows ? ra : 0), kin && ra < n_rows);
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
    const double* as = As + (c % STAGES) * KC * LDS + wm * 64 + g;
    const double* bs = Bs + (c % STAGES) * KC * LDS + wn * 32 + g;
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
      double a[8], b[4];
#pragma unroll
      for (int mt = 0; mt < 8; ++mt) a[mt] = as[(kk + q) * LDS + mt * 8];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) b[nt] = bs[(kk + q) * LDS + nt * 8];
#pragma unroll
      for (int mt = 0; mt < 8; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();   // TRSM overwrites the panel it read: every warp must be done reading
  // ---- epilogue ----
#pragma unroll
  for (int mt = 0; mt < 8; ++mt) {
    const int r = i0 + wm * 64 + mt * 8 + g;
    if (r >= n_rows) continue;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cc = wn * 32 + nt * 8 + 2 * q + e;
        if (MODE == MODE_SYRK) {
          const int c = j0 + cc;
          if (c < n_cols) S[(size_t)c * lda + r] -= acc[mt][nt][e];
        } else {
          if (cc < K) S[(size_t)(k0 + cc) * lda + r] = acc[mt][nt][e];
        }
      }
    }
  }
}

// ---- diagonal block: Cholesky + inverse of the factor, one CTA of 512 threads -----------------
constexpr int PT = 512;          // threads of the diagonal-block kernel
constexpr int PLD = NB + 1;     // odd leading dimension: column reads by consecutive lanes conflict-free
constexpr int TLD = 97;         // leading dimension of the 32 x 96 product scratch
constexpr int POTRF_SMEM = (NB * PLD + NB + 32 * TLD) * (int)sizeof(double);

__global__ void __launch_bounds__(PT, 1)
k_potrf128(double* __restrict__ S, int ld, int k0, int nb, double* __restrict__ Linv, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  double* D = sm;                 // D[c * PLD + r]: lower triangle + diagonal = the factor L;
                                  // strict upper triangle = the inverse, transposed: X(r,c), r > c, at D[r * PLD + c]
  double* xd = sm + NB * PLD;     // diagonal of the inverse
  double* Tm = xd + NB;           // Tm[rr * TLD + cc]: 32 x (32 bi) product scratch
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < NB * NB; e += PT) {
    const int r = e % NB, c = e / NB;
    D[c * PLD + r] = (r < nb && c < nb && r >= c) ? S[(size_t)(k0 + c) * ld + k0 + r] : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  for (int b0 = 0; b0 < NB; b0 += 32) {
    // (1) warp 0: unblocked Cholesky of the 32x32 diagonal sub-block, lane = row
    if (warp == 0) {
      const int r = b0 + lane;
      for (int j = 0; j < 32; ++j) {
        const int cj = b0 + j;
        const double djj = D[cj * PLD + cj];
        if (!(djj > 0.0) && lane == 0 && cj < nb) atomicCAS(info, 0, k0 + cj + 1);
        const double dj = sqrt(djj), inv = 1.0 / dj;
        __syncwarp();
        if (lane == j) D[cj * PLD + r] = dj;
        else if (lane > j) D[cj * PLD + r] *= inv;
        __syncwarp();
        const double lrj = D[cj * PLD + r];
        for (int k = j + 1; k < 32; ++k) {
          const double lkj = D[cj * PLD + b0 + k];
          if (lane >= k) D[(b0 + k) * PLD + r] -= lrj * lkj;
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // (2) rows below the sub-block: forward substitution, one thread per row
    const int below = NB - b0 - 32;
    if (tid < below) {
      const int r = b0 + 32 + tid;
      double x[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        double s = D[(b0 + c) * PLD + r];
#pragma unroll
        for (int p = 0; p < c; ++p) s -= x[p] * D[(b0 + p) * PLD + b0 + c];
        x[c] = s / D[(b0 + c) * PLD + b0 + c];
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) D[(b0 + c) * PLD + r] = x[c];
    }
    __syncthreads();
    // (3) symmetric rank-32 update of the remaining lower triangle
    for (int e = tid; e < below * below; e += PT) {
      const int rr = e % below, cc = e / below;
      if (rr < cc) continue;
      const int r = b0 + 32 + rr, c = b0 + 32 + cc;
      double s = 0.0;
#pragma unroll 8
      for (int p = 0; p < 32; ++p) s += D[(b0 + p) * PLD + r] * D[(b0 + p) * PLD + c];
      D[c * PLD + r] -= s;
    }
    __syncthreads();
  }
  // write the factor back (lower triangle incl. diagonal)
  for (int e = tid; e < NB * NB; e += PT) {
    const int r = e % NB, c = e / NB;
    if (r < nb && c < nb && r >= c) S[(size_t)(k0 + c) * ld + k0 + r] = D[c * PLD + r];
  }
  // ---- inverse of the lower-triangular factor, by 32-blocks --------------------------------
  __syncthreads();
  // diagonal blocks: warps 0..3, lane = column of the inverse, forward substitution on e_c
  if (warp < 4) {
    const int b0 = warp * 32, c = b0 + lane;
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int p = 0; p < i; ++p) s -= D[(b0 + p) * PLD + b0 + i] * x[p];
      x[i] = s / D[(b0 + i) * PLD + b0 + i];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i == lane) xd[c] = x[i];
      else if (i > lane) D[(b0 + i) * PLD + c] = x[i];     // X(b0+i, c) in its transposed slot
    }
  }
  __syncthreads();
  // off-diagonal blocks, one block row at a time:  X_ib = -X_ii (sum_{p=b}^{i-1} L_ip X_pb)
  for (int bi = 1; bi < 4; ++bi) {
    const int w = 32 * bi;
    for (int e = tid; e < 32 * w; e += PT) {
      const int rr = e % 32, cc = e / 32;          // T[rr][cc] = sum_{p=cc}^{w-1} L[w+rr][p] X[p][cc]
      double s = D[cc * PLD + w + rr] * xd[cc];
      for (int p = cc + 1; p < w; ++p) s += D[p * PLD + w + rr] * D[p * PLD + cc];
      Tm[rr * TLD + cc] = s;
    }
    __syncthreads();
    for (int e = tid; e < 32 * w; e += PT) {
      const int rr = e % 32, cc = e / 32;          // X[w+rr][cc] = -sum_{p<=rr} X_ii[rr][p] T[p][cc]
      double s = xd[w + rr] * Tm[rr * TLD + cc];
      for (int p = 0; p < rr; ++p) s += D[(w + rr) * PLD + w + p] * Tm[p * TLD + cc];
      D[(w + rr) * PLD + cc] = -s;
    }
    __syncthreads();
  }
  for (int e = tid; e < NB * NB; e += PT) {
    const int r = e % NB, c = e / NB;
    double v = 0.0;
    if (r < nb && c < nb) v = r > c ? D[r * PLD + c] : (r == c ? xd[r] : 0.0);
    Linv[(size_t)c * NB + r] = v;
  }
}

// ---- block triangular solves with the stored inverses -----------------------------------------
// forward step k:  y_k = Linv_kk b_k (every CTA recomputes it; CTA 0 stores it), then the CTA's
// row tile i > k:  b_i -= L_ik y_k
__global__ void __launch_bounds__(NB)
k_trsv_fwd(const double* __restrict__ S, int ld, int n, int k, const double* __restrict__ Linv, double* __restrict__ b,
           double* __restrict__ y) {
  __shared__ double bk[NB], yk[NB];
  const int t = threadIdx.x, k0 = k * NB;
  const int nb = min(NB, n - k0);
  bk[t] = t < nb ? b[k0 + t] : 0.0;
  __syncthreads();
  const double* Li = Linv + (size_t)k * NB * NB;
  double s = 0.0;
  for (int c = 0; c <= t; ++c) s += Li[(size_t)c * NB + t] * bk[c];
  yk[t] = s;
  __syncthreads();
  if (blockIdx.x == 0) {
    if (t < nb) y[k0 + t] = s;      // a separate vector: other CTAs of this step still read b_k
    return;
  }
  const int r = (k + blockIdx.x) * NB + t;
  if (r >= n) return;
  double acc = 0.0;
#pragma unroll 8
  for (int c = 0; c < nb; ++c) acc += S[(size_t)(k0 + c) * ld + r] * yk[c];
  b[r] -= acc;
}

// backward step k:  x_k = Linv_kk^T y_k (recomputed per CTA; CTA 0 stores), then the CTA's column
// tile j < k:  y_j -= L_kj^T x_k
__global__ void __launch_bounds__(NB)
k_trsv_bwd(const double* __restrict__ S, int ld, int n, int k, const double* __restrict__ Linv, double* __restrict__ y,
           double* __restrict__ x) {
  __shared__ double yk[NB], xk[NB];
  __shared__ double tile[32][NB + 1];
  const int t = threadIdx.x, k0 = k * NB;
  const int nb = min(NB, n - k0);
  yk[t] = t < nb ? y[k0 + t] : 0.0;
  __syncthreads();
  const double* Li = Linv + (size_t)k * NB * NB;
  double s = 0.0;
  for (int r = t; r < nb; ++r) s += Li[(size_t)t * NB + r] * yk[r];   // column t of Linv = row t of Linv^T
  xk[t] = s;
  __syncthreads();
  if (blockIdx.x == 0) {
    if (t < nb) x[k0 + t] = s;
    return;
  }
  const int j0 = (blockIdx.x - 1) * NB;
  // y_j[c] -= sum_r L[k0+r][j0+c] x_k[r]; stage 32 columns at a time so global reads stay coalesced
  for (int c0 = 0; c0 < NB; c0 += 32) {
    for (int e = t; e < 32 * NB; e += NB) {
      const int r = e % NB, c = e / NB;
      tile[c][r] = r < nb ? S[(size_t)(j0 + c0 + c) * ld + k0 + r] : 0.0;
    }
    __syncthreads();
    if (t < 32) {
      double acc = 0.0;
#pragma unroll 8
      for (int r = 0; r < NB; ++r) acc += tile[t][r] * xk[r];
      y[j0 + c0 + t] -= acc;
    }
    __syncthreads();
  }
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
  int n = 0;
  double* Linv = nullptr;     // T blocks of NB x NB
  double* ybuf = nullptr;     // intermediate vector of the triangular solves
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
  if (p->tiles) cudaFree(p->tiles);
  if (p->side) cudaStreamDestroy(p->side);
  for (auto e : p->events) cudaEventDestroy(e);
  delete p;
}

CholWorkspace::~CholWorkspace() { destroy_plan(plan); }

// Enqueue the whole factor + solve schedule on `main` (and `side` for the look-ahead panels).
static int enqueue(CholPlan& P, cudaStream_t main, bool lookahead) {
  const int n = P.n, T = (n + NB - 1) / NB;
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
  // panel(k) = potrf + trsm on stream s
  auto panel = [&](int k, cudaStream_t s) {
    const int k0 = k * NB, nb = std::min(NB, n - k0);
    k_potrf128<<<1, PT, POTRF_SMEM, s>>>(S, n, k0, nb, P.Linv + (size_t)k * NB * NB, P.info);
    ++P.launches;
    if (n_panel[k]) {
      k_gemm_nt<MODE_TRSM><<<n_panel[k], GEMM_THREADS, GEMM_SMEM, s>>>(S, n, n, n, k0, nb, P.Linv + (size_t)k * NB * NB, NB,
                                                                        P.tiles + off_panel[k]);
      ++P.launches;
    }
  };
  panel(0, main);
  for (int k = 0; k + 1 < T; ++k) {
    const int k0 = k * NB;
    // strip update of block column k+1, then its panel (look-ahead: on the side stream)
    k_gemm_nt<MODE_SYRK><<<n_strip[k], GEMM_THREADS, GEMM_SMEM, main>>>(S, n, n, n, k0, NB, nullptr, n, P.tiles + off_strip[k]);
    ++P.launches;
    if (lookahead && n_rest[k]) {
      cudaEvent_t e1 = next_event(), e2 = next_event();
      CKC(cudaEventRecord(e1, main));
      CKC(cudaStreamWaitEvent(P.side, e1, 0));
      panel(k + 1, P.side);
      CKC(cudaEventRecord(e2, P.side));
      k_gemm_nt<MODE_SYRK><<<n_rest[k], GEMM_THREADS, GEMM_SMEM, main>>>(S, n, n, n, k0, NB, nullptr, n, P.tiles + off_rest[k]);
      ++P.launches;
      CKC(cudaStreamWaitEvent(main, e2, 0));
    } else {
      if (n_rest[k]) {
        k_gemm_nt<MODE_SYRK><<<n_rest[k], GEMM_THREADS, GEMM_SMEM, main>>>(S, n, n, n, k0, NB, nullptr, n, P.tiles + off_rest[k]);
        ++P.launches;
      }
      panel(k + 1, main);
    }
  }
  for (int k = 0; k < T; ++k) {
    k_trsv_fwd<<<T - k, NB, 0, main>>>(S, n, n, k, P.Linv, P.rhs, P.ybuf);
    ++P.launches;
  }
  for (int k = T - 1; k >= 0; --k) {
    k_trsv_bwd<<<k + 1, NB, 0, main>>>(S, n, n, k, P.Linv, P.ybuf, P.rhs);
    ++P.launches;
  }
  CKC(cudaGetLastError());
  return STBA_OK;
}

int chol_factor_solve(CholWorkspace& ws, double* S, int n, double* rhs, int* dev_info, cudaStream_t stream, int* n_launches) {
  if (n <= 0) return STBA_OK;
  if (n % 2) return STBA_ERR_UNSUPPORTED;   // 16-byte cp.async rows; n = 6 * cameras is always even
  CholPlan* P = ws.plan;
  if (!P || P->S != S || P->n != n || P->rhs != rhs || P->info != dev_info) {
    destroy_plan(P);
    ws.plan = P = new CholPlan();
    P->S = S; P->n = n; P->rhs = rhs; P->info = dev_info;
    const int T = (n + NB - 1) / NB;
    CKC(cudaMalloc(&P->Linv, (size_t)T * NB * NB * sizeof(double)));
    CKC(cudaMalloc(&P->ybuf, (size_t)T * NB * sizeof(double)));
    {
      // tile lists: for step k, [panel rows | column k+1 strip | the rest], stored back to back
      std::vector<int2> h;
      P->off_strip.resize(T); P->off_rest.resize(T); P->off_panel.resize(T);
      P->n_strip.resize(T); P->n_rest.resize(T); P->n_panel.resize(T);
      for (int k = 0; k < T; ++k) {
        P->off_panel[k] = h.size();
        for (int i = k + 1; i < T; ++i) h.push_back(make_int2(i, 0));
        P->n_panel[k] = (int)(h.size() - P->off_panel[k]);
        P->off_strip[k] = h.size();
        if (k + 1 < T) for (int i = k + 1; i < T; ++i) h.push_back(make_int2(i, k + 1));
        P->n_strip[k] = (int)(h.size() - P->off_strip[k]);
        P->off_rest[k] = h.size();
        for (int j = k + 2; j < T; ++j)
          for (int i = j; i < T; ++i) h.push_back(make_int2(i, j));
        P->n_rest[k] = (int)(h.size() - P->off_rest[k]);
      }
      CKC(cudaMalloc(&P->tiles, std::max<size_t>(h.size(), 1) * sizeof(int2)));
      CKC(cudaMemcpy(P->tiles, h.data(), h.size() * sizeof(int2), cudaMemcpyHostToDevice));
    }
    CKC(cudaStreamCreateWithFlags(&P->side, cudaStreamNonBlocking));
    CKC(cudaFuncSetAttribute(k_gemm_nt<MODE_SYRK>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    CKC(cudaFuncSetAttribute(k_gemm_nt<MODE_TRSM>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
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
  if (n_launches) *n_launches += P->launches;
  return STBA_OK;
}

}  // namespace stba
