// bsq_launch.h -- kernel-family tables and the per-translation-unit launchers.
// The fill kernels are instantiated over (arithmetic, lane geometry, B rule,
// end rule, sign of go); each family/arithmetic lives in its own .cu so the
// build parallelises.
#pragma once
#include "bsq_common.h"

#define BSQ_FILL_THREADS 128

struct BsqK1Cfg { int G, K; };   // STD: G lanes x K columns per pass
struct BsqK2Cfg { int G, C; };   // BANDED: G lanes x C diagonals

struct BsqK3Cfg { int G, K; };   // BANDED, any width: strips of G*K columns, one warp each

#define BSQ_K1_NCFG 3
#define BSQ_K2_NCFG 6
static const BsqK1Cfg BSQ_K1_CFGS[BSQ_K1_NCFG] = {{8, 19}, {16, 16}, {32, 16}};
static const BsqK2Cfg BSQ_K2_CFGS[BSQ_K2_NCFG] = {{8, 8}, {8, 18}, {16, 10}, {16, 16}, {32, 16}, {32, 32}};

#define BSQ_K3_NCFG 3
static const BsqK3Cfg BSQ_K3_CFGS[BSQ_K3_NCFG] = {{32, 2}, {32, 4}, {32, 8}};
#define BSQ_K3_THREADS 64     // two independent strip workers per block

enum { BSQ_OP_OCCUPANCY = 0, BSQ_OP_LAUNCH = 1 };

// op == BSQ_OP_OCCUPANCY: returns resident blocks per SM (>= 1) for the kernel
// op == BSQ_OP_LAUNCH   : launches it on `grid` blocks, returns 0 or a negative
//                         error (unknown variant)
int bsq_k1_f64_op(int cfg, int ball, int eall, int gos, int op, int grid, cudaStream_t st, const BsqFillArgs* A);
int bsq_k1_i32_op(int cfg, int ball, int eall, int gos, int op, int grid, cudaStream_t st, const BsqFillArgs* A);
int bsq_k2_f64_op(int cfg, int blocal, int gos, int op, int grid, cudaStream_t st, const BsqFillArgs* A);
int bsq_k2_i32_op(int cfg, int blocal, int gos, int op, int grid, cudaStream_t st, const BsqFillArgs* A);
// K3 (cfg = index into BSQ_K3_CFGS): launches on `grid` blocks of BSQ_K3_THREADS
int bsq_k3_f64_op(int cfg, int btype, int gos, int grid, cudaStream_t st, const BsqFillArgs* A);
int bsq_k3_i32_op(int cfg, int btype, int gos, int grid, cudaStream_t st, const BsqFillArgs* A);
// K1h: STD LOCAL, two short problems per lane in packed 16-bit arithmetic (geometry of
// BSQ_K1_CFGS[cfg], cfg 0 or 1); gos <= 0; same != 0: all pairs of the chunk have one shape
int bsq_k1h_op(int cfg, int gos, int same, int op, int grid, cudaStream_t st, const BsqFillArgs* A);
