// bsq_k3_dispatch.cuh -- runtime (cfg, type, sign of go) -> K3 instantiation.
#pragma once
#include "bsq_k3.cuh"
#include "bsq_launch.h"

// rows per time step of the three strip geometries (K = 2, 4, 8 columns per lane)
#ifndef BSQ_K3_R0
#define BSQ_K3_R0 1
#define BSQ_K3_R1 2
#define BSQ_K3_R2 2
#endif

template <class S, int K, int R>
static int bsq_k3_variant(int btype, int gos, int grid, cudaStream_t st, const BsqFillArgs* A) {
#define BSQ_K3_CASE(BT, GS) \
  if (btype == BT && gos == GS) { BSQ_LAUNCH((bsq_k3_fill<S, K, R, BT, GS>), grid, BSQ_K3_THREADS, 0, st, *A); (void)st; return 0; }
  BSQ_K3_CASE(0, -1) BSQ_K3_CASE(0, 0) BSQ_K3_CASE(0, 1)
  BSQ_K3_CASE(1, -1) BSQ_K3_CASE(1, 0) BSQ_K3_CASE(1, 1)
  BSQ_K3_CASE(2, -1) BSQ_K3_CASE(2, 0) BSQ_K3_CASE(2, 1)
#undef BSQ_K3_CASE
  return -1;
}

// cfg 0: the latency form (bsq_k3l_fill: 64-column strips, one row per step, uniform gated step)
template <class S>
static int bsq_k3l_variant(int btype, int gos, int grid, cudaStream_t st, const BsqFillArgs* A) {
#define BSQ_K3L_CASE(BT, GS) \
  if (btype == BT && gos == GS) { BSQ_LAUNCH((bsq_k3l_fill<S, BT, GS>), grid, BSQ_K3_THREADS, 0, st, *A); (void)st; return 0; }
  BSQ_K3L_CASE(0, -1) BSQ_K3L_CASE(0, 0) BSQ_K3L_CASE(0, 1)
  BSQ_K3L_CASE(1, -1) BSQ_K3L_CASE(1, 0) BSQ_K3L_CASE(1, 1)
  BSQ_K3L_CASE(2, -1) BSQ_K3L_CASE(2, 0) BSQ_K3L_CASE(2, 1)
#undef BSQ_K3L_CASE
  return -1;
}

template <class S>
static int bsq_k3_op_t(int cfg, int btype, int gos, int grid, cudaStream_t st, const BsqFillArgs* A) {
  switch (cfg) {
    case 0: return bsq_k3l_variant<S>(btype, gos, grid, st, A);
    case 1: return bsq_k3_variant<S, 4, BSQ_K3_R1>(btype, gos, grid, st, A);
    case 2: return bsq_k3_variant<S, 8, BSQ_K3_R2>(btype, gos, grid, st, A);
  }
  return -1;
}
