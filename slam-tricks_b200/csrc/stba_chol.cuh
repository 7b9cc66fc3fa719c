// Dense reduced-camera solve (own back end).  Interface only in this file; kernels in stba_chol_impl.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/stba.h"

namespace stba {

struct CholPlan;
struct CholWorkspace {
  CholPlan* plan = nullptr;   // captured schedule + side buffers for one (S, n); built lazily
  ~CholWorkspace();
  void reset();               // drop the captured schedule (before its buffers are freed)
};

// In-place lower Cholesky of the column-major n x n matrix S (leading dimension ld >= n + 1, ld even:
// row n of S is scratch for the right-hand side) followed by the two triangular solves on rhs (n).  dev_info <- 0 on success, k > 0 if the leading minor k
// is not positive definite (LAPACK convention).  *n_launches += kernels launched.
int chol_factor_solve(CholWorkspace& ws, double* S, int n, int ld, double* rhs, int* dev_info, cudaStream_t stream,
                      int* n_launches);

// Split factorisation for `nranks` GPUs holding the SAME matrix (multi-GPU reduced solve; see stba_chol.cu): first half of
// the block columns on every rank, the Schur-complement update of the second half spread over the ranks (tiles exchanged
// with NCCL broadcasts on `nccl_comm`, an ncclComm_t), second half and the backward substitution on every rank.  Returns
// STBA_ERR_UNSUPPORTED for matrices too small to split (callers use chol_factor_solve).
struct SplitPlan;
struct SplitWorkspace {
  SplitPlan* plan = nullptr;
  ~SplitWorkspace();
  void reset();
};
int chol_factor_solve_split(SplitWorkspace& ws, double* S, int n, int ld, double* rhs, int* dev_info, cudaStream_t stream, void* nccl_comm, int rank,
                            int nranks, int* n_launches);

// The two triangular solves on a finished lower Cholesky factor (column-major, leading dimension ld even) with the
// one-launch flag-driven substitution kernels of the own back end; used behind cusolverDnDpotrf.
struct SolvePlan;
struct SolveWorkspace {
  SolvePlan* plan = nullptr;
  ~SolveWorkspace();
  void reset();               // release the stream-ordered buffers (before the stream they were allocated on is destroyed)
};
int chol_solve_with_factor(SolveWorkspace& ws, const double* S, int n, int ld, double* rhs, int* dev_info, cudaStream_t stream,
                           int* n_launches);

}  // namespace stba
