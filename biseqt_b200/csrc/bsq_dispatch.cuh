// bsq_dispatch.cuh -- turns runtime (cfg, flags) into a template instantiation.
#pragma once
#include "bsq_kernels.cuh"
#include "bsq_launch.h"

template <class KernelT>
static int bsq_do_op(KernelT kern, int op, int grid, cudaStream_t st, const BsqFillArgs* A) {
  if (op == BSQ_OP_OCCUPANCY) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, BSQ_FILL_THREADS, 0) != cudaSuccess) nb = 1;
    return nb < 1 ? 1 : nb;
  }
  BSQ_LAUNCH(kern, grid, BSQ_FILL_THREADS, 0, st, *A);
  (void)st;
  return 0;
}

template <class S, int G, int K>
static int bsq_k1_variant(int ball, int eall, int gos, int op, int grid, cudaStream_t st, const BsqFillArgs* A) {
#define BSQ_K1_CASE(B, E, GS) \
  if (ball == B && eall == E && gos == GS) return bsq_do_op(bsq_k1_fill<S, G, K, B != 0, E, GS>, op, grid, st, A);
  BSQ_K1_CASE(0, 0, -1) BSQ_K1_CASE(0, 0, 0) BSQ_K1_CASE(0, 0, 1)
  BSQ_K1_CASE(0, 1, -1) BSQ_K1_CASE(0, 1, 0) BSQ_K1_CASE(0, 1, 1)
  BSQ_K1_CASE(0, 2, -1) BSQ_K1_CASE(0, 2, 0) BSQ_K1_CASE(0, 2, 1)
  BSQ_K1_CASE(1, 0, -1) BSQ_K1_CASE(1, 0, 0) BSQ_K1_CASE(1, 0, 1)
  BSQ_K1_CASE(1, 1, -1) BSQ_K1_CASE(1, 1, 0) BSQ_K1_CASE(1, 1, 1)
  BSQ_K1_CASE(1, 2, -1) BSQ_K1_CASE(1, 2, 0) BSQ_K1_CASE(1, 2, 1)
#undef BSQ_K1_CASE
  return -1;
}

template <class S>
static int bsq_k1_op_t(int cfg, int ball, int eall, int gos, int op, int grid, cudaStream_t st, const BsqFillArgs* A) {
  switch (cfg) {
    case 0: return bsq_k1_variant<S, 8, 19>(ball, eall, gos, op, grid, st, A);
    case 1: return bsq_k1_variant<S, 16, 16>(ball, eall, gos, op, grid, st, A);
    case 2: return bsq_k1_variant<S, 32, 16>(ball, eall, gos, op, grid, st, A);
  }
  return -1;
}

template <class S, int G, int C>
static int bsq_k2_variant(int blocal, int gos, int op, int grid, cudaStream_t st, const BsqFillArgs* A) {
#define BSQ_K2_CASE(B, GS) \
  if (blocal == B && gos == GS) return bsq_do_op(bsq_k2_fill<S, G, C, B, GS>, op, grid, st, A);
  BSQ_K2_CASE(0, -1) BSQ_K2_CASE(0, 0) BSQ_K2_CASE(0, 1)
  BSQ_K2_CASE(1, -1) BSQ_K2_CASE(1, 0) BSQ_K2_CASE(1, 1)
  if constexpr (S::IS_INT) {   // blocal == 2: B_LOCAL with packed end-cell keys (see bsq_k2_fill)
    BSQ_K2_CASE(2, -1) BSQ_K2_CASE(2, 0) BSQ_K2_CASE(2, 1)
  }
#undef BSQ_K2_CASE
  return -1;
}

template <class S>
static int bsq_k2_op_t(int cfg, int blocal, int gos, int op, int grid, cudaStream_t st, const BsqFillArgs* A) {
  switch (cfg) {
    case 0: return bsq_k2_variant<S, 8, 8>(blocal, gos, op, grid, st, A);
    case 1: return bsq_k2_variant<S, 8, 18>(blocal, gos, op, grid, st, A);
    case 2: return bsq_k2_variant<S, 16, 10>(blocal, gos, op, grid, st, A);
    case 3: return bsq_k2_variant<S, 16, 16>(blocal, gos, op, grid, st, A);
    case 4: return bsq_k2_variant<S, 32, 16>(blocal, gos, op, grid, st, A);
    case 5: return bsq_k2_variant<S, 32, 32>(blocal, gos, op, grid, st, A);
  }
  return -1;
}
