// k3 fill kernels (wide bands, strip-pipelined), f64 arithmetic (see bsq_k3.cuh)
#include "bsq_k3_dispatch.cuh"
int bsq_k3_f64_op(int cfg, int btype, int gos, int grid, cudaStream_t st, const BsqFillArgs* A) {
  return bsq_k3_op_t<ScF64>(cfg, btype, gos, grid, st, A);
}
