// k1h fill kernels: two short problems per lane, packed 16-bit arithmetic (see bsq_k1h.cuh)
#include "bsq_k1h.cuh"

template <class KernelT>
static int bsq_k1h_do(KernelT kern, int op, int grid, cudaStream_t st, const BsqFillArgs* A) {
  if (op == BSQ_OP_OCCUPANCY) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, BSQ_FILL_THREADS, 0) != cudaSuccess) nb = 1;
    return nb < 1 ? 1 : nb;
  }
  BSQ_LAUNCH(kern, grid, BSQ_FILL_THREADS, 0, st, *A);
  (void)st;
  return 0;
}

int bsq_k1h_op(int cfg, int gos, int same, int op, int grid, cudaStream_t st, const BsqFillArgs* A) {
  if (same) {
    if (cfg == 0 && gos < 0) return bsq_k1h_do(bsq_k1h_fill<8, 19, -1, true>, op, grid, st, A);
    if (cfg == 0 && gos == 0) return bsq_k1h_do(bsq_k1h_fill<8, 19, 0, true>, op, grid, st, A);
    if (cfg == 1 && gos < 0) return bsq_k1h_do(bsq_k1h_fill<16, 16, -1, true>, op, grid, st, A);
    if (cfg == 1 && gos == 0) return bsq_k1h_do(bsq_k1h_fill<16, 16, 0, true>, op, grid, st, A);
  } else {
    if (cfg == 0 && gos < 0) return bsq_k1h_do(bsq_k1h_fill<8, 19, -1, false>, op, grid, st, A);
    if (cfg == 0 && gos == 0) return bsq_k1h_do(bsq_k1h_fill<8, 19, 0, false>, op, grid, st, A);
    if (cfg == 1 && gos < 0) return bsq_k1h_do(bsq_k1h_fill<16, 16, -1, false>, op, grid, st, A);
    if (cfg == 1 && gos == 0) return bsq_k1h_do(bsq_k1h_fill<16, 16, 0, false>, op, grid, st, A);
  }
  return -1;
}
