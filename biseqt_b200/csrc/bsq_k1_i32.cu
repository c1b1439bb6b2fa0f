// k1 fill kernels, i32 arithmetic (see bsq_kernels.cuh)
#include "bsq_dispatch.cuh"
int bsq_k1_i32_op(int cfg, int ball, int eall, int gos, int op, int grid, cudaStream_t st, const BsqFillArgs* A) {
  return bsq_k1_op_t<ScI32>(cfg, ball, eall, gos, op, grid, st, A);
}
