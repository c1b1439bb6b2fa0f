// k2 fill kernels, f64 arithmetic (see bsq_kernels.cuh)
#include "bsq_dispatch.cuh"
int bsq_k2_f64_op(int cfg, int blocal, int gos, int op, int grid, cudaStream_t st, const BsqFillArgs* A) {
  return bsq_k2_op_t<ScF64>(cfg, blocal, gos, op, grid, st, A);
}
