// Own dense Cholesky back end — placeholder until the blocked kernel lands (this round).
#include "stba_chol.cuh"

namespace stba {
int chol_factor_solve(CholWorkspace&, double*, int, double*, int*, cudaStream_t, int*) {
  return STBA_ERR_UNSUPPORTED;
}
}  // namespace stba
